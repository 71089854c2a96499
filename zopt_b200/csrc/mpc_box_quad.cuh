// Box-constrained lqrMpc, shared problem definition, (n,m) = (12,4), fp32: FOUR threads per problem.
//
// Same ADMM, same tables and the same arithmetic per entry as the thread-per-problem kernel (mpc_box.cuh), split for
// latency: that kernel is bound by one warp's instruction latency (16,384 problems are 512 warps -- less than one per
// scheduler -- each issuing ~820 instructions per horizon step); here thread t of a quad owns state rows 3t..3t+2 and
// control row t, so a step is ~250 instructions per thread and there are 4x as many warps.
//   * operands: every thread keeps ITS rows of A (for A x), its columns of A (for A'p), its rows / column of B in
//     registers (104 words, loaded once from the kernel parameters) -- still no shared memory, no loads in the products;
//   * the 12-vectors p and x and the 4-vectors h and u are all-gathered inside the quad with shuffles (32 per step);
//   * gains: row t of K_k, rows 3t..3t+2 of K_k' and row t of G_k^-1 from the rho-grid table (7 128-bit loads per step);
//   * state: [chunk of 8 problems][step][13 words][32 lanes] -- w (3+1), lambda (3+1), kff (1), dlambda (3+1) of a thread;
//   * residuals are reduced over the quad on termination-check iterations, so all decisions are quad-uniform; quads of a
//     warp converge independently (shuffles and barriers use the quad's own mask).
// Small batches only pay here (measured crossover in zb_api.cu); fp64 would need 208 registers of operands and stays on
// the thread-per-problem kernel.
#pragma once
#include "mpc_box.cuh"

namespace zb {
namespace box {

constexpr int QW = 13;  // state words per thread per step: w[4], lambda[4], kff, dlambda[4]
constexpr int Q_W = 0, Q_L = 4, Q_K = 8, Q_D = 9;

inline long long quad_ws_elems(int N, long long Bsz) { return (Bsz + 7) / 8 * (long long)(N + 1) * QW * 32; }

struct QuadRes {
    float rp, rd, nz, nw, nl, ndl, S;
    bool cert_ok;
};

template <bool CHK>
__device__ __forceinline__ void qproject(float zi, float lb, float ub, float rho, float inv_rho, float alpha, float wo, float lo,
                                         float& wn, float& ln, float& dl, QuadRes& r) {
    const float INF = __int_as_float(0x7f800000);
    const float zr = alpha * zi + (1.f - alpha) * wo;
    wn = clampv<float>(zr + lo * inv_rho, lb, ub);
    dl = rho * (zr - wn);
    ln = lo + dl;
    if (!CHK) return;
    r.rp = fmaxf(r.rp, fabsf(zi - wn));
    r.rd = fmaxf(r.rd, rho * fabsf(wn - wo));
    r.nz = fmaxf(r.nz, fabsf(zi));
    r.nw = fmaxf(r.nw, fabsf(wn));
    r.nl = fmaxf(r.nl, fabsf(ln));
    r.ndl = fmaxf(r.ndl, fabsf(dl));
    if (dl > 0.f) { if (ub == INF) r.cert_ok = false; else r.S += ub * dl; }
    else if (dl < 0.f) { if (lb == -INF) r.cert_ok = false; else r.S += lb * dl; }
    r.S -= dl * zi;
}

// per-thread slice of the problem definition
struct QuadOps {
    float Arow[3][12], Acol[3][12], Brow[3][4], Bcol[12];
    float xlb[3], xub[3], ulb, uub;
};

__device__ __forceinline__ void gather12(const float (&own)[3], float (&full)[12], unsigned qmask, int qbase) {
#pragma unroll
    for (int s = 0; s < 4; ++s)
#pragma unroll
        for (int r = 0; r < 3; ++r) full[3 * s + r] = __shfl_sync(qmask, own[r], qbase + s);
}
__device__ __forceinline__ void gather4(float own, float (&full)[4], unsigned qmask, int qbase) {
#pragma unroll
    for (int s = 0; s < 4; ++s) full[s] = __shfl_sync(qmask, own, qbase + s);
}
__device__ __forceinline__ float qmax(float v, unsigned qmask) {
    v = fmaxf(v, __shfl_xor_sync(qmask, v, 1));
    return fmaxf(v, __shfl_xor_sync(qmask, v, 2));
}
__device__ __forceinline__ float qsum(float v, unsigned qmask) {
    v += __shfl_xor_sync(qmask, v, 1);
    return v + __shfl_xor_sync(qmask, v, 2);
}

// SMEM: the warp's ADMM state lives in shared memory (small batches: a wave of CTAs holds every problem's state on chip,
// ~30 cycles per access instead of an L2 round trip per horizon step); otherwise in the global workspace, L2-only.
template <bool SMEM> __device__ __forceinline__ float zq_ld(const float* p) { if constexpr (SMEM) return *p; else return __ldcg(p); }
template <bool SMEM> __device__ __forceinline__ void zq_st(float* p, float v) { if constexpr (SMEM) *p = v; else __stcg(p, v); }
#define ZQ(k, j) ws[(long long)(k) * (QW * 32) + (j) * 32]
#define ZQ_LD(k, j) zq_ld<SMEM>(&ZQ(k, j))
#define ZQ_ST(k, j, v) zq_st<SMEM>(&ZQ(k, j), v)

template <bool CHK, bool SMEM>
__device__ __forceinline__ void quad_forward(const QuadOps& O, float* ws, int N, const float* tab, const float (&x0)[3], float rho,
                                             float inv_rho, float alpha, float* zx, float* zu, int t, unsigned qmask, int qbase, QuadRes& r) {
    float x[3] = {x0[0], x0[1], x0[2]};
    float fw[4], fl[4], fk;  // this thread's w, lambda, kff of the step being processed, loaded one step ahead
#pragma unroll
    for (int j = 0; j < 4; ++j) { fw[j] = ZQ_LD(0, Q_W + j); fl[j] = ZQ_LD(0, Q_L + j); }
    fk = ZQ_LD(0, Q_K);
    for (int k = 0; k < N; ++k) {
        float Kr[12];
        ldrow<float, 12>(tab + (long long)k * TW + t * 12, Kr);  // row t of K_k
        float cw[4], cl[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { cw[j] = fw[j]; cl[j] = fl[j]; }
        const float ck = fk;
        float xf[12];
        gather12(x, xf, qmask, qbase);
#pragma unroll
        for (int j = 0; j < 4; ++j) { fw[j] = ZQ_LD(k + 1, Q_W + j); fl[j] = ZQ_LD(k + 1, Q_L + j); }
        if (k + 1 < N) fk = ZQ_LD(k + 1, Q_K);
        float xn[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 12; ++j) s += O.Arow[q][j] * xf[j];
            xn[q] = s;
        }
        float u = ck;
#pragma unroll
        for (int j = 0; j < 12; ++j) u += Kr[j] * xf[j];
        u = -u;
        float uf[4];
        gather4(u, uf, qmask, qbase);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float wn, ln, dl;
            qproject<CHK>(j < 3 ? x[j] : u, j < 3 ? O.xlb[j] : O.ulb, j < 3 ? O.xub[j] : O.uub, rho, inv_rho, alpha, cw[j], cl[j], wn, ln, dl, r);
            ZQ_ST(k, Q_W + j, wn);
            ZQ_ST(k, Q_L + j, ln);
            if (CHK) ZQ_ST(k, Q_D + j, dl);
        }
        if (CHK) {
#pragma unroll
            for (int q = 0; q < 3; ++q) zx[(long long)k * 12 + 3 * t + q] = x[q];
            zu[(long long)k * 4 + t] = u;
        }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
#pragma unroll
            for (int a = 0; a < 4; ++a) xn[q] += O.Brow[q][a] * uf[a];
            x[q] = xn[q];
        }
    }
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float wn, ln, dl;
        qproject<CHK>(x[j], O.xlb[j], O.xub[j], rho, inv_rho, alpha, fw[j], fl[j], wn, ln, dl, r);
        ZQ_ST(N, Q_W + j, wn);
        ZQ_ST(N, Q_L + j, ln);
        if (CHK) ZQ_ST(N, Q_D + j, dl);
    }
    if (CHK) {
#pragma unroll
        for (int q = 0; q < 3; ++q) zx[(long long)N * 12 + 3 * t + q] = x[q];
    }
}

// ADMM iterations of one problem by its quad; mirrors admm_solve<T> (mpc_box.cuh) decision for decision
template <bool SMEM>
__device__ __forceinline__ int quad_admm_solve(const QuadOps& O, const Params<float>& P, float* ws, const float (&x0)[3], float* zx, float* zu,
                                               int& level, int& it, int t, unsigned qmask, int qbase) {
    const int N = P.N;
    int status = 1;
    float rho = ldexpf(P.rho0, level - LEVEL0), inv_rho = 1.f / rho;
    int rho_gap = P.check_every, rho_next = 0;
    for (it = 1; it <= P.max_iter; ++it) {
        const float* tab = P.tab + (long long)level * N * TW;
        // ---- backward vector sweep ----
        float p[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) p[q] = ZQ_LD(N, Q_L + q) - rho * ZQ_LD(N, Q_W + q);
        float sl[4], sw[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) { sw[j] = ZQ_LD(N - 1, Q_W + j); sl[j] = ZQ_LD(N - 1, Q_L + j); }
        for (int k = N - 1; k >= 0; --k) {
            const float* row = tab + (long long)k * TW;
            float Gr[4], Kt[12];
            ldrow<float, 4>(row + 48 + t * 4, Gr);     // row t of G_k^-1
            ldrow<float, 12>(row + 64 + t * 12, Kt);   // rows 3t..3t+2 of K_k'
            float g[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) g[j] = sl[j] - rho * sw[j];
            if (k > 0) {
#pragma unroll
                for (int j = 0; j < 4; ++j) { sw[j] = ZQ_LD(k - 1, Q_W + j); sl[j] = ZQ_LD(k - 1, Q_L + j); }
            }
            float pf[12];
            gather12(p, pf, qmask, qbase);
            float pn[3];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                float s = g[q];
#pragma unroll
                for (int j = 0; j < 12; ++j) s += O.Acol[q][j] * pf[j];
                pn[q] = s;
            }
            float h = g[3];
#pragma unroll
            for (int i = 0; i < 12; ++i) h += O.Bcol[i] * pf[i];
            float hf[4];
            gather4(h, hf, qmask, qbase);
            float kf = 0.f;
#pragma unroll
            for (int c = 0; c < 4; ++c) kf += Gr[c] * hf[c];
            ZQ_ST(k, Q_K, kf);
#pragma unroll
            for (int q = 0; q < 3; ++q) {
#pragma unroll
                for (int a = 0; a < 4; ++a) pn[q] -= Kt[q * 4 + a] * hf[a];
                p[q] = pn[q];
            }
        }
        // ---- forward rollout + projection ----
        const bool chk = (it % P.check_every == 0) || it == P.max_iter;
        QuadRes r{0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, true};
        if (chk) quad_forward<true, SMEM>(O, ws, N, tab, x0, rho, inv_rho, P.alpha, zx, zu, t, qmask, qbase, r);
        else quad_forward<false, SMEM>(O, ws, N, tab, x0, rho, inv_rho, P.alpha, zx, zu, t, qmask, qbase, r);
        if (chk) {
            r.rp = qmax(r.rp, qmask); r.rd = qmax(r.rd, qmask); r.nz = qmax(r.nz, qmask); r.nw = qmax(r.nw, qmask);
            r.nl = qmax(r.nl, qmask); r.ndl = qmax(r.ndl, qmask); r.S = qsum(r.S, qmask);
            r.cert_ok = __all_sync(qmask, r.cert_ok);
            if (r.rp <= P.eps_abs + P.eps_rel * fmaxf(r.nz, r.nw) && r.rd <= P.eps_abs + P.eps_rel * r.nl) {
                status = 0;
                break;
            }
            if (r.cert_ok && r.ndl > P.eps_inf && r.S < -P.eps_inf * r.ndl) {  // primal infeasibility certificate (adjoint sweep)
                float mu[3], g = 0.f;
#pragma unroll
                for (int q = 0; q < 3; ++q) mu[q] = ZQ_LD(N, Q_D + q);
                for (int k = N - 1; k >= 0; --k) {
                    float mf[12];
                    gather12(mu, mf, qmask, qbase);
                    float s = ZQ_LD(k, Q_D + 3);
#pragma unroll
                    for (int i = 0; i < 12; ++i) s += O.Bcol[i] * mf[i];
                    g = fmaxf(g, fabsf(s));
#pragma unroll
                    for (int q = 0; q < 3; ++q) {
                        float m2 = ZQ_LD(k, Q_D + q);
#pragma unroll
                        for (int j = 0; j < 12; ++j) m2 += O.Acol[q][j] * mf[j];
                        mu[q] = m2;
                    }
                }
                g = qmax(g, qmask);
                if (g <= P.eps_inf * r.ndl) {
                    status = 2;
                    break;
                }
            }
            if (it < P.max_iter && it >= rho_next) {  // residual balancing on the rho grid, with back-off
                const float rpn = r.rp / fmaxf(fmaxf(r.nz, r.nw), 1e-10f), rdn = r.rd / fmaxf(r.nl, 1e-10f);
                const float ratio = sqrtf(rpn / fmaxf(rdn, 1e-30f));
                if ((ratio > 5.f || ratio < 0.2f) && rdn > 0.f) {
                    const float lg = log2f(ratio);
                    int nl = level + (int)(lg >= 0.f ? lg + 0.5f : lg - 0.5f);
                    nl = nl < 0 ? 0 : (nl > LEVELS - 1 ? LEVELS - 1 : nl);
                    if (nl != level) {
                        rho_gap *= 2;
                        rho_next = it + rho_gap;
                        level = nl;
                        rho = ldexpf(P.rho0, level - LEVEL0);
                        inv_rho = 1.f / rho;
                    }
                }
            }
        }
    }
    if (it > P.max_iter) it = P.max_iter;
    return status;
}

__device__ __forceinline__ void quad_load_ops(const Ops<float>& G, int t, QuadOps& O) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
#pragma unroll
        for (int j = 0; j < 12; ++j) { O.Arow[q][j] = G.A[(3 * t + q) * 12 + j]; O.Acol[q][j] = G.A[j * 12 + 3 * t + q]; }
#pragma unroll
        for (int a = 0; a < 4; ++a) O.Brow[q][a] = G.B[(3 * t + q) * 4 + a];
        O.xlb[q] = G.xlb[3 * t + q];
        O.xub[q] = G.xub[3 * t + q];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) O.Bcol[i] = G.B[i * 4 + t];
    O.ulb = G.ulb[t];
    O.uub = G.uub[t];
}

template <bool LOOP, bool SMEM>
__global__ void __launch_bounds__(64) k_mpc_box_quad(const __grid_constant__ Ops<float> G, const __grid_constant__ Params<float> P) {
    extern __shared__ float quad_state[];
    const int lane = threadIdx.x & 31, t = lane & 3, qbase = lane & 28;
    const unsigned qmask = 0xFu << qbase;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long b = warp * 8 + (lane >> 2);
    if (b >= P.Bsz) return;  // whole quads leave together
    const int N = P.N;
    QuadOps O;
    quad_load_ops(G, t, O);
    float* ws = (SMEM ? quad_state + (threadIdx.x >> 5) * ((long long)(N + 1) * QW * 32) : P.ws + warp * ((long long)(N + 1) * QW * 32)) + lane;
    float* zx = P.xTraj + b * (long long)(N + 1) * 12;
    float* zu = P.uTraj + b * (long long)N * 4;
    const float INF = __int_as_float(0x7f800000), NaN = INF - INF;
    float x[3];
#pragma unroll
    for (int q = 0; q < 3; ++q) x[q] = P.x0[b * 12 + 3 * t + q];
    for (int k = 0; k <= N; ++k)
#pragma unroll
        for (int j = 0; j < 8; ++j) ZQ_ST(k, j, 0.f);
    int level = LEVEL0;
    if (!LOOP) {
        bool bad = false;
#pragma unroll
        for (int q = 0; q < 3; ++q) bad |= !(x[q] >= O.xlb[q] - P.eps_abs && x[q] <= O.xub[q] + P.eps_abs);
        bad = __any_sync(qmask, bad);
        int status = 2, it = 0;
        if (!bad) status = quad_admm_solve<SMEM>(O, P, ws, x, zx, zu, level, it, t, qmask, qbase);
        __syncwarp(qmask);
        if (status == 2) {
            for (long long i = t; i < (long long)(N + 1) * 12; i += 4) zx[i] = NaN;
            for (long long i = t; i < (long long)N * 4; i += 4) zu[i] = NaN;
        }
        __syncwarp(qmask);
        P.u0[b * 4 + t] = zu[t];
        if (t == 0) {
            P.status[b] = (int8_t)status;
            if (P.iters) P.iters[b] = it;
        }
    } else {
        float* xs = P.xSim + b * (long long)(P.Tsim + 1) * 12;
        float* us = P.uSim + b * (long long)P.Tsim * 4;
        int worst = 0, ts = 0;
        long long total = 0;
        for (; ts < P.Tsim; ++ts) {
            bool bad = false;
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                x[q] = clampv<float>(x[q], O.xlb[q] + P.clip_margin, O.xub[q] - P.clip_margin);
                bad |= !(x[q] == x[q]);
                xs[(long long)ts * 12 + 3 * t + q] = x[q];
            }
            bad = __any_sync(qmask, bad);
            int it = 0;
            const int st = bad ? 2 : quad_admm_solve<SMEM>(O, P, ws, x, zx, zu, level, it, t, qmask, qbase);
            total += it;
            worst = st > worst ? st : worst;
            if (st == 2) break;
            __syncwarp(qmask);  // the plan written by the other threads of the quad is visible
            us[(long long)ts * 4 + t] = zu[t];
#pragma unroll
            for (int q = 0; q < 3; ++q) x[q] = zx[12 + 3 * t + q];
            for (int k = 0; k < N; ++k) {  // receding horizon: shift (w, lambda) one step; the last control keeps its own
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    if (!((j & 3) == 3 && k == N - 1)) ZQ_ST(k, j, ZQ_LD(k + 1, j));
            }
        }
        if (ts < P.Tsim) {
            for (long long i = (long long)(ts + 1) * 12 + t; i < (long long)(P.Tsim + 1) * 12; i += 4) xs[i] = NaN;
            for (long long i = (long long)ts * 4 + t; i < (long long)P.Tsim * 4; i += 4) us[i] = NaN;
        } else {
#pragma unroll
            for (int q = 0; q < 3; ++q) xs[(long long)P.Tsim * 12 + 3 * t + q] = x[q];
        }
        if (t == 0) {
            P.status[b] = (int8_t)worst;
            if (P.iters) P.iters[b] = (int32_t)(total > 2147483647LL ? 2147483647LL : total);
        }
    }
}

#undef ZQ
#undef ZQ_LD
#undef ZQ_ST

}  // namespace box
}  // namespace zb
