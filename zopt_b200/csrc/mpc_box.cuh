// Box-constrained lqrMpc (zopt/mpcUtils.py:12-81) for ONE problem definition and a batch of initial states, (n,m) = (12,4).
//
// This is the reference's usage: `lqrMpc(A,B,Q,R,N,x_lb,x_ub,u_lb,u_ub)` builds the QP once (mpcUtils.py:14-59), `solve(x0)`
// re-solves it for a new parameter x0 (mpcUtils.py:61-81, demos/lqrMpc.py:42-47).  Batched over x0, every problem shares
// A, B, Q, R, Qf and the box.  Same ADMM as the generic kernel (zb_problems.cuh::admm_problem: OSQP's splitting with the
// dynamics kept exact, over-relaxation, residual-balanced rho) re-organised for the machine:
//   * A, B and the bounds travel BY VALUE in the kernel parameters (constant bank): the 384 multiply-adds per horizon step
//     that involve A or B take their matrix operand straight from c[0][..] -- no load instruction, no shared memory.
//   * rho is kept on a geometric grid rho0 * 2^(j-12), j = 0..24.  The gains K_k(rho), G_k(rho)^-1 of the equality-constrained
//     LQ solve depend on rho only, so ONE table per grid point serves the whole batch (k_box_tables: a warp per level);
//     threads read their level's rows with broadcast loads from L1/L2 instead of factoring per problem.
//   * the per-problem ADMM state (w, lambda, feed-forward kff; 36 words per horizon step) lives in a problem-interleaved
//     workspace [step][word][problem]: every access of a warp is one coalesced 128-byte (fp32) line.
//   * forward rollout, relaxation, projection, multiplier update and residuals are fused step by step; z is only written
//     (to the public trajectory buffers) on termination-check iterations.
// One thread per problem; all per-problem vectors (p, h, x, u) in registers.  The body is __host__ __device__ so that
// tests/hostsim runs the same arithmetic on the CPU against the oracle.
#pragma once
#include <math.h>
#include <stdint.h>

#include "zb_math.cuh"

namespace zb {
namespace box {

constexpr int LEVELS = 25, LEVEL0 = 12;  // rho_j = rho0 * 2^(j - LEVEL0)
constexpr int SW = 52;                   // state words per horizon step: w_x 12, w_u 4, lam_x 12, lam_u 4, kff 4, dlam_x 12, dlam_u 4
constexpr int O_WX = 0, O_WU = 12, O_LX = 16, O_LU = 28, O_KFF = 32, O_DX = 36, O_DU = 48;
constexpr int TW = 112;                  // table words per (level, step): K 4x12, G^-1 4x4, K' 12x4 (for the 4-threads-per-problem kernel)

template <typename T>
struct Ops {  // the problem definition, shared by the batch (kernel parameter -> constant bank)
    T A[144], B[48], xlb[12], xub[12], ulb[4], uub[4];
};
template <typename T>
struct Costs {  // Hessian blocks for the table builder
    T Q[144], R[16], Qf[144];
};

template <typename T>
struct Params {
    long long Bsz;
    int N;
    const T* x0;
    T *u0, *xTraj, *uTraj;
    int8_t* status;
    int32_t* iters;
    T* ws;         // ceil(Bsz/32) chunks of (N+1) * SW * 32: [chunk][step][word][problem % 32]
    const T* tab;  // LEVELS * N * TW
    int max_iter, check_every;
    T rho0, alpha, eps_abs, eps_rel, eps_inf;
    // closed loop only
    int Tsim;
    T *xSim, *uSim;
    T clip_margin;
};

ZB_HD long long ws_elems(int N, long long Bsz) { return (Bsz + 31) / 32 * (long long)(N + 1) * SW * 32; }
ZB_HD long long tab_elems(int N) { return (long long)LEVELS * N * TW; }

template <typename T>
ZB_HD T clampv(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }

// accumulators of one ADMM iteration's residuals
template <typename T>
struct Res {
    T rp, rd, nz, nw, nl, ndl, S;
    bool cert_ok;
};

// relaxation + projection + multiplier update of one entry (admm_problem, zb_problems.cuh "relaxation, projection ...")
// CHK = termination-check iteration: only then are the residuals, norms and the certificate sums needed
template <typename T, bool CHK>
ZB_HD void project(T zi, T lb, T ub, T rho, T inv_rho, T alpha, T wo, T lo, T& wn, T& ln, T& dl, Res<T>& r) {
    const T INF = T(1) / T(0);
    const T zr = alpha * zi + (T(1) - alpha) * wo;
    wn = clampv<T>(zr + lo * inv_rho, lb, ub);
    dl = rho * (zr - wn);
    ln = lo + dl;
    if (!CHK) return;
    r.rp = fmax(r.rp, fabs(zi - wn));
    r.rd = fmax(r.rd, rho * fabs(wn - wo));
    r.nz = fmax(r.nz, fabs(zi));
    r.nw = fmax(r.nw, fabs(wn));
    r.nl = fmax(r.nl, fabs(ln));
    {
        r.ndl = fmax(r.ndl, fabs(dl));
        if (dl > T(0)) { if (ub == INF) r.cert_ok = false; else r.S += ub * dl; }
        else if (dl < T(0)) { if (lb == -INF) r.cert_ok = false; else r.S += lb * dl; }
        r.S -= dl * zi;
    }
}

#define ZB_WS(k, j) ws[(long long)(k) * (SW * 32) + (j) * 32]
// The ADMM state streams through once per sweep: keep it out of L1 (L2 only) so that L1 holds the gain tables.
#ifdef __CUDA_ARCH__
#define ZB_LD(k, j) __ldcg(&ZB_WS(k, j))
#define ZB_ST(k, j, v) __stcg(&ZB_WS(k, j), v)
#else
#define ZB_LD(k, j) ZB_WS(k, j)
#define ZB_ST(k, j, v) ZB_WS(k, j) = (v)
#endif
// CNT table words (16-byte aligned rows) -> registers with 128-bit read-only loads
template <typename T, int CNT>
ZB_HD void ldrow(const T* __restrict__ p, T (&r)[CNT]) {
#ifdef __CUDA_ARCH__
    if constexpr (sizeof(T) == 4) {
#pragma unroll
        for (int e = 0; e < CNT / 4; ++e) {
            const float4 v = __ldg(reinterpret_cast<const float4*>(p) + e);
            r[4 * e] = v.x; r[4 * e + 1] = v.y; r[4 * e + 2] = v.z; r[4 * e + 3] = v.w;
        }
    } else {
#pragma unroll
        for (int e = 0; e < CNT / 2; ++e) {
            const double2 v = __ldg(reinterpret_cast<const double2*>(p) + e);
            r[2 * e] = v.x; r[2 * e + 1] = v.y;
        }
    }
#else
    for (int e = 0; e < CNT; ++e) r[e] = p[e];
#endif
}

// z-update, part 2 (forward rollout) fused with relaxation / projection / multiplier update, one horizon step at a time
template <typename T, bool CHK>
ZB_HD void forward(const Ops<T>& O, T* ws, int N, const T* tab, const T (&x0)[12], T rho, T inv_rho, T alpha, T* zx, T* zu, Res<T>& r) {
    T x[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = x0[i];
    T fw[16], fl[16], fk[4];  // w, lambda, kff of the step being processed, loaded one step ahead
#pragma unroll
    for (int j = 0; j < 16; ++j) { fw[j] = ZB_LD(0, O_WX + j); fl[j] = ZB_LD(0, O_LX + j); }
#pragma unroll
    for (int a = 0; a < 4; ++a) fk[a] = ZB_LD(0, O_KFF + a);
    for (int k = 0; k < N; ++k) {
        const T* Kk = tab + (long long)k * TW;
        T Kr[48];  // issued first, consumed after the 144 multiply-adds of A x (covers the L1 latency)
        ldrow<T, 48>(Kk, Kr);
        T cw[16], cl[16], u[4], xn[12];
#pragma unroll
        for (int j = 0; j < 16; ++j) { cw[j] = fw[j]; cl[j] = fl[j]; }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            T s = T(0);
#pragma unroll
            for (int j = 0; j < 12; ++j) s += O.A[i * 12 + j] * x[j];
            xn[i] = s;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            T s = fk[a];
#pragma unroll
            for (int j = 0; j < 12; ++j) s += Kr[a * 12 + j] * x[j];
            u[a] = -s;
        }
        {   // prefetch step k+1 (the terminal step has no control part: its slots exist but stay zero)
#pragma unroll
            for (int j = 0; j < 16; ++j) { fw[j] = ZB_LD(k + 1, O_WX + j); fl[j] = ZB_LD(k + 1, O_LX + j); }
            if (k + 1 < N) {
#pragma unroll
                for (int a = 0; a < 4; ++a) fk[a] = ZB_LD(k + 1, O_KFF + a);
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) {
            T wn, ln, dl;
            project<T, CHK>(j < 12 ? x[j] : u[j - 12], j < 12 ? O.xlb[j] : O.ulb[j - 12], j < 12 ? O.xub[j] : O.uub[j - 12], rho, inv_rho, alpha, cw[j], cl[j], wn, ln, dl, r);
            ZB_ST(k, O_WX + j, wn);
            ZB_ST(k, O_LX + j, ln);
            if (CHK) ZB_ST(k, O_DX + j, dl);
        }
        if (CHK) {
#pragma unroll
            for (int i = 0; i < 12; ++i) zx[(long long)k * 12 + i] = x[i];
#pragma unroll
            for (int a = 0; a < 4; ++a) zu[(long long)k * 4 + a] = u[a];
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) {
#pragma unroll
            for (int a = 0; a < 4; ++a) xn[i] += O.B[i * 4 + a] * u[a];
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = xn[i];
    }
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        T wn, ln, dl;
        project<T, CHK>(x[i], O.xlb[i], O.xub[i], rho, inv_rho, alpha, fw[i], fl[i], wn, ln, dl, r);
        ZB_ST(N, O_WX + i, wn);
        ZB_ST(N, O_LX + i, ln);
        if (CHK) ZB_ST(N, O_DX + i, dl);
    }
    if (CHK) {
#pragma unroll
        for (int i = 0; i < 12; ++i) zx[(long long)N * 12 + i] = x[i];
    }
}

// ADMM iterations from the state (w, lambda) found in `ws` (zeros = cold start; the previous solve's = warm start), rho level
// carried in/out.  On return the plan z of the last iteration is in zx (N+1,12) / zu (N,4).  Returns 0 optimal, 1 max_iter
// reached, 2 infeasible (certificate); `it` = iterations run.
template <typename T>
ZB_HD int admm_solve(const Ops<T>& O, const Params<T>& P, T* ws, const T (&x0)[12], T* zx, T* zu, int& level, int& it) {
    const int N = P.N;
    int status = 1;
    T rho = ldexp(P.rho0, level - LEVEL0), inv_rho = T(1) / rho;
    int rho_gap = P.check_every, rho_next = 0;  // back-off: every rho move doubles the wait before the next one (no limit cycles)
    for (it = 1; it <= P.max_iter; ++it) {
        const T* tab = P.tab + (long long)level * N * TW;
        // ---- z-update, part 1: backward vector sweep with linear terms lam - rho*w ----
        T p[12];
#pragma unroll
        for (int i = 0; i < 12; ++i) p[i] = ZB_LD(N, O_LX + i) - rho * ZB_LD(N, O_WX + i);
        T sl[16], sw[16];  // lambda / w of the step being processed, loaded one step ahead of their use
#pragma unroll
        for (int j = 0; j < 16; ++j) { sw[j] = ZB_LD(N - 1, O_WX + j); sl[j] = ZB_LD(N - 1, O_LX + j); }
        for (int k = N - 1; k >= 0; --k) {
            const T* Kk = tab + (long long)k * TW;
            T Kr[64];  // K_k (48) and G_k^-1 (16): issued first, consumed after the 192 multiply-adds with A and B
            ldrow<T, 64>(Kk, Kr);
            T g[16];  // lam - rho*w: [0,12) state part, [12,16) control part
#pragma unroll
            for (int j = 0; j < 16; ++j) g[j] = sl[j] - rho * sw[j];
            if (k > 0) {
#pragma unroll
                for (int j = 0; j < 16; ++j) { sw[j] = ZB_LD(k - 1, O_WX + j); sl[j] = ZB_LD(k - 1, O_LX + j); }
            }
            T h[4], pn[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                T s = g[i];
#pragma unroll
                for (int j = 0; j < 12; ++j) s += O.A[j * 12 + i] * p[j];
                pn[i] = s;
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                T s = g[12 + a];
#pragma unroll
                for (int i = 0; i < 12; ++i) s += O.B[i * 4 + a] * p[i];
                h[a] = s;
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                T s = T(0);
#pragma unroll
                for (int c = 0; c < 4; ++c) s += Kr[48 + a * 4 + c] * h[c];
                ZB_ST(k, O_KFF + a, s);
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) {
#pragma unroll
                for (int a = 0; a < 4; ++a) pn[i] -= Kr[a * 12 + i] * h[a];
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) p[i] = pn[i];
        }
        // ---- z-update, part 2 + projection (forward<>) ----
        const bool chk = (it % P.check_every == 0) || it == P.max_iter;
        Res<T> r{T(0), T(0), T(0), T(0), T(0), T(0), T(0), true};
        if (chk) forward<T, true>(O, ws, N, tab, x0, rho, inv_rho, P.alpha, zx, zu, r);
        else forward<T, false>(O, ws, N, tab, x0, rho, inv_rho, P.alpha, zx, zu, r);
        if (chk) {
            if (r.rp <= P.eps_abs + P.eps_rel * fmax(r.nz, r.nw) && r.rd <= P.eps_abs + P.eps_rel * r.nl) {
                status = 0;
                break;
            }
            // primal infeasibility certificate: delta-lambda separates the box from the dynamics' affine set
            if (r.cert_ok && r.ndl > P.eps_inf && r.S < -P.eps_inf * r.ndl) {
                T mu[12], g = T(0);
#pragma unroll
                for (int i = 0; i < 12; ++i) mu[i] = ZB_LD(N, O_DX + i);
                for (int k = N - 1; k >= 0; --k) {
#pragma unroll
                    for (int a = 0; a < 4; ++a) {
                        T s = ZB_LD(k, O_DU + a);
#pragma unroll
                        for (int i = 0; i < 12; ++i) s += O.B[i * 4 + a] * mu[i];
                        g = fmax(g, fabs(s));
                    }
                    T mn[12];
#pragma unroll
                    for (int i = 0; i < 12; ++i) {
                        T s = ZB_LD(k, O_DX + i);
#pragma unroll
                        for (int j = 0; j < 12; ++j) s += O.A[j * 12 + i] * mu[j];
                        mn[i] = s;
                    }
#pragma unroll
                    for (int i = 0; i < 12; ++i) mu[i] = mn[i];
                }
                if (g <= P.eps_inf * r.ndl) {
                    status = 2;
                    break;
                }
            }
            // residual balancing as OSQP, on the rho grid: move by round(log2(ratio)) levels when the residuals are 5x apart
            if (it < P.max_iter && it >= rho_next) {
                const T rpn = r.rp / fmax(fmax(r.nz, r.nw), T(1e-10)), rdn = r.rd / fmax(r.nl, T(1e-10));
                const T ratio = sqrt(rpn / fmax(rdn, T(1e-30)));
                if ((ratio > T(5) || ratio < T(0.2)) && rdn > T(0)) {
                    const T lg = log2(ratio);
                    int nl = level + (int)(lg >= T(0) ? lg + T(0.5) : lg - T(0.5));
                    nl = nl < 0 ? 0 : (nl > LEVELS - 1 ? LEVELS - 1 : nl);
                    if (nl != level) {
                        rho_gap *= 2;
                        rho_next = it + rho_gap;
                        level = nl;
                        rho = ldexp(P.rho0, level - LEVEL0);
                        inv_rho = T(1) / rho;
                    }
                }
            }
        }
    }
    if (it > P.max_iter) it = P.max_iter;
    return status;
}

template <typename T>
ZB_HD void problem(const Ops<T>& O, const Params<T>& P, long long b) {
    const int N = P.N;
    // 32 problems share a chunk; within it a (step, word) row is 32 consecutive elements, so the word offsets inside a
    // step are compile-time immediates of the load / store instructions
    T* ws = P.ws + (b >> 5) * ((long long)(N + 1) * SW * 32) + (b & 31);
    T* zx = P.xTraj + b * (long long)(N + 1) * 12;
    T* zu = P.uTraj + b * (long long)N * 4;
    T* u0 = P.u0 + b * 4;
    const T INF = T(1) / T(0);
    T x0[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) x0[i] = P.x0[b * 12 + i];
    // x_0 = x0 is itself box-constrained in the reference QP (mpcUtils.py:56,58): outside the box = infeasible
    bool x0_bad = false;
#pragma unroll
    for (int i = 0; i < 12; ++i) x0_bad |= !(x0[i] >= O.xlb[i] - P.eps_abs && x0[i] <= O.xub[i] + P.eps_abs);
    int status = 2, it = 0;
    if (!x0_bad) {
        int level = LEVEL0;
        for (int k = 0; k <= N; ++k)
#pragma unroll
            for (int j = 0; j < 32; ++j) ZB_ST(k, j, T(0));
        status = admm_solve<T>(O, P, ws, x0, zx, zu, level, it);
    }
    if (status == 2) {
        const T nan = INF - INF;
        for (long long i = 0; i < (long long)(N + 1) * 12; ++i) zx[i] = nan;
        for (long long i = 0; i < (long long)N * 4; ++i) zu[i] = nan;
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) u0[a] = zu[a];
    P.status[b] = (int8_t)status;
    if (P.iters) P.iters[b] = it;
}

// Receding-horizon loop of demos/lqrMpc.py:42-47 for one problem: every simulation step clips the state into the box
// (x_lb + margin, x_ub - margin: the QP also constrains x_0), solves the MPC problem from it WARM-STARTED with the previous
// step's (w, lambda) and rho level (cvxpy re-solves with warm_start=True by default), applies the first move and takes the
// plan's own next state as the new state ("assume perfect tracking", demos/lqrMpc.py:47).  xSim (Tsim+1,12) holds the
// clipped states the solves started from (+ the final one), uSim (Tsim,4) the applied moves; status = worst over the steps,
// iters = total ADMM iterations; a step that is infeasible ends the loop with NaNs from there on.
template <typename T>
ZB_HD void closed_loop_problem(const Ops<T>& O, const Params<T>& P, long long b) {
    const int N = P.N;
    T* ws = P.ws + (b >> 5) * ((long long)(N + 1) * SW * 32) + (b & 31);
    T* zx = P.xTraj + b * (long long)(N + 1) * 12;  // scratch: the current plan
    T* zu = P.uTraj + b * (long long)N * 4;
    T* xs = P.xSim + b * (long long)(P.Tsim + 1) * 12;
    T* us = P.uSim + b * (long long)P.Tsim * 4;
    const T INF = T(1) / T(0);
    T x[12];
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = P.x0[b * 12 + i];
    for (int k = 0; k <= N; ++k)
#pragma unroll
        for (int j = 0; j < 32; ++j) ZB_ST(k, j, T(0));
    int level = LEVEL0, worst = 0;
    long long total = 0;
    int ts = 0;
    for (; ts < P.Tsim; ++ts) {
        bool bad = false;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            x[i] = clampv<T>(x[i], O.xlb[i] + P.clip_margin, O.xub[i] - P.clip_margin);
            bad |= !(x[i] == x[i]);
            xs[(long long)ts * 12 + i] = x[i];
        }
        int it = 0;
        const int st = bad ? 2 : admm_solve<T>(O, P, ws, x, zx, zu, level, it);
        total += it;
        worst = st > worst ? st : worst;
        if (st == 2) break;
#pragma unroll
        for (int a = 0; a < 4; ++a) us[(long long)ts * 4 + a] = zu[a];
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = zx[12 + i];
        // receding horizon: the next problem's step k is this one's step k+1 -- shift (w, lambda) one step down (the last
        // control keeps its own values, the terminal state its own)
        for (int k = 0; k < N; ++k) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
                const bool upart = (j & 15) >= 12;
                if (!(upart && k == N - 1)) ZB_ST(k, j, ZB_LD(k + 1, j));
            }
        }
    }
    if (ts < P.Tsim) {  // infeasible step: NaN from here on
        const T nan = INF - INF;
        for (long long i = (long long)(ts + 1) * 12; i < (long long)(P.Tsim + 1) * 12; ++i) xs[i] = nan;
        for (long long i = (long long)ts * 4; i < (long long)P.Tsim * 4; ++i) us[i] = nan;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i) xs[(long long)P.Tsim * 12 + i] = x[i];
    }
    P.status[b] = (int8_t)worst;
    if (P.iters) P.iters[b] = (int32_t)(total > 2147483647LL ? 2147483647LL : total);
}
#undef ZB_WS
#undef ZB_LD
#undef ZB_ST

// One level of the table: gains K_k = G^-1 B'PA and G_k^-1 of the Hessian-form LQ problem with weights
// (2Q + rho I, 2R + rho I, 2Qf + rho I) -- admm_factor (zb_problems.cuh) for rho = rho0 * 2^(level - LEVEL0).
// `lane`/`nlanes` split the element loops (device: the 32 lanes of a warp; host: 0/1); `sync` is __syncwarp on the device.
template <typename T, typename Sync>
ZB_HD void table_level(const Ops<T>& O, const Costs<T>& C, int N, T rho, T* tabl, T* sm /* 592 words scratch */, int lane, int nlanes, Sync sync) {
    T *Pm = sm, *W = sm + 144, *Acl = sm + 288, *BtP = sm + 432, *M4 = sm + 480, *Kk = sm + 528, *G = sm + 576;  // G 16 -> 592
    for (int e = lane; e < 144; e += nlanes) Pm[e] = T(2) * C.Qf[e] + ((e / 12 == e % 12) ? rho : T(0));
    sync();
    for (int k = N - 1; k >= 0; --k) {
        T* out = tabl + (long long)k * TW;
        for (int e = lane; e < 48; e += nlanes) {  // BtP[a][j] = sum_i B[i][a] P[i][j]
            const int a = e / 12, j = e % 12;
            T s = T(0);
            for (int i = 0; i < 12; ++i) s += O.B[i * 4 + a] * Pm[i * 12 + j];
            BtP[e] = s;
        }
        sync();
        for (int e = lane; e < 16; e += nlanes) {  // G = B'PB + 2R + rho I
            const int a = e / 4, c = e % 4;
            T s = T(2) * C.R[e] + (a == c ? rho : T(0));
            for (int i = 0; i < 12; ++i) s += BtP[a * 12 + i] * O.B[i * 4 + c];
            G[e] = s;
        }
        for (int e = lane; e < 48; e += nlanes) {  // M4 = B'PA
            const int a = e / 12, j = e % 12;
            T s = T(0);
            for (int i = 0; i < 12; ++i) s += BtP[a * 12 + i] * O.A[i * 12 + j];
            M4[e] = s;
        }
        sync();
        if (lane == 0) {  // G^-1 by LU with partial pivoting (4x4), into the table
            T Gc[16], Gi[16];
            for (int e = 0; e < 16; ++e) { Gc[e] = G[e]; Gi[e] = (e / 4 == e % 4) ? T(1) : T(0); }
            lu_solve(Gc, 4, Gi, 4);
            for (int e = 0; e < 16; ++e) { G[e] = Gi[e]; out[48 + e] = Gi[e]; }
        }
        sync();
        for (int e = lane; e < 48; e += nlanes) {  // K = G^-1 B'PA
            const int a = e / 12, j = e % 12;
            T s = T(0);
            for (int c = 0; c < 4; ++c) s += G[a * 4 + c] * M4[c * 12 + j];
            Kk[e] = s;
            out[e] = s;
            out[64 + j * 4 + a] = s;  // transposed copy
        }
        sync();
        for (int e = lane; e < 144; e += nlanes) {  // Acl = A - B K
            const int i = e / 12, j = e % 12;
            T s = O.A[e];
            for (int a = 0; a < 4; ++a) s -= O.B[i * 4 + a] * Kk[a * 12 + j];
            Acl[e] = s;
        }
        sync();
        for (int e = lane; e < 144; e += nlanes) {  // W = P Acl
            const int i = e / 12, j = e % 12;
            T s = T(0);
            for (int q = 0; q < 12; ++q) s += Pm[i * 12 + q] * Acl[q * 12 + j];
            W[e] = s;
        }
        sync();
        for (int e = lane; e < 144; e += nlanes) {  // P = A' W + 2Q + rho I  (Acl reused as the unsymmetrised result)
            const int i = e / 12, j = e % 12;
            T s = T(2) * C.Q[e] + (i == j ? rho : T(0));
            for (int q = 0; q < 12; ++q) s += O.A[q * 12 + i] * W[q * 12 + j];
            Acl[e] = s;
        }
        sync();
        for (int e = lane; e < 144; e += nlanes) {  // keep P exactly symmetric
            const int i = e / 12, j = e % 12;
            Pm[e] = T(0.5) * (Acl[i * 12 + j] + Acl[j * 12 + i]);
        }
        sync();
    }
}

#ifdef __CUDACC__
template <typename T>
__global__ void __launch_bounds__(32) k_mpc_box(const __grid_constant__ Ops<T> O, const __grid_constant__ Params<T> P) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < P.Bsz) problem<T>(O, P, b);
}

template <typename T>
__global__ void __launch_bounds__(32) k_mpc_box_closed_loop(const __grid_constant__ Ops<T> O, const __grid_constant__ Params<T> P) {
    const long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (b < P.Bsz) closed_loop_problem<T>(O, P, b);
}

template <typename T>
__global__ void __launch_bounds__(32) k_box_tables(const __grid_constant__ Ops<T> O, const __grid_constant__ Costs<T> C, int N, T rho0, T* tab) {
    __shared__ T sm[592];
    const int level = blockIdx.x;
    table_level<T>(O, C, N, ldexp(rho0, level - LEVEL0), tab + (long long)level * N * TW, sm, (int)threadIdx.x, 32, [] { __syncwarp(); });
}
#endif

}  // namespace box
}  // namespace zb
