// Fused 16-way line search of the iLQR / DDP solvers for the registered quadcopter model (n=12, m=4), fp32 and fp64.
//
// One launch = forwardPass2 (zopt/ilqrUtils.py:116-150) + the solver's bookkeeping (ilqrUtils.py:318-321) for every
// active problem: 16 lanes of a half-warp roll out the 16 step sizes alpha = 0.5^j (ilqrUtils.py:139-146) with the state,
// the control and the diagonal cost weights in registers, the costs are exchanged with shuffles, argmin (first minimum,
// a NaN wins -- jnp.argmin) picks the winner and the same 16 lanes commit its trajectory:
//   * the lanes of alpha = 1 and 1/2 (the usual winners) store their rollouts speculatively; if one of them won the
//     half-warp copies it over the current trajectory with 128-bit accesses while it is still in L2;
//   * otherwise the half-warp re-runs the winning step size (identical arithmetic) and lane 0 stores it in place.
// The per-step operands every lane of a problem needs -- L_k (4x12), l_k, xPrev_k, uPrev_k: 68 words -- are the only
// data the rollout reads.  They are streamed with cp.async into a 4-stage per-problem ring in shared memory (issued
// three steps ahead, no registers held) and read back with 128-bit broadcast loads; the generic kernel's dependent
// global loads of L_k were 64 % of its stall samples (profiles/r1s3_ilqr_forward_*).
// Arithmetic order equals rollout_quad (zb_problems.cuh), so parity with the oracle is unchanged.
#pragma once
#include "ilqr_params.cuh"

namespace zb {

__device__ __forceinline__ void fwd_cp16(void* dst_smem, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
__device__ __forceinline__ void fwd_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int K>
__device__ __forceinline__ void fwd_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(K) : "memory"); }

__device__ __forceinline__ void st4(float* p, const float* v) { *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]); }
__device__ __forceinline__ void st4(double* p, const double* v) {
    *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(v[2], v[3]);
}
__device__ __forceinline__ void ld4(const float* p, float* v) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
}
__device__ __forceinline__ void ld4(const double* p, double* v) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
}

constexpr int FWD_DEPTH = 4;   // ring stages (steps in flight: 3)
constexpr int FWD_STAGE = 68;  // words per step: L 48 | l 4 | xPrev 12 | uPrev 4
// resident 64-thread CTAs per SM asked of the compiler: fp64 8 (16 warps, caps the kernel at 128 registers), fp32 12

// QDIAG: the caller asserted diagonal Q, R (weights in registers); otherwise the dense quadratic forms are evaluated row by
// row from global memory, exactly as the generic rollout does.
template <typename T, bool QDIAG>
__global__ void __launch_bounds__(64, sizeof(T) == 8 ? 8 : 12) k_forward_quad(FwdQuadP P) {
    constexpr int n = 12, m = 4, D = FWD_DEPTH, ST = FWD_STAGE;
    constexpr int EPC = 16 / (int)sizeof(T);  // words per 16-byte chunk
    constexpr int NCH = ST / EPC;             // chunks per step: 34 (fp64) / 17 (fp32)
    constexpr int C_L = 48 / EPC, C_l = 4 / EPC, C_X = 12 / EPC;
    __shared__ __align__(16) T ring[4][D][ST];
    const int tid = threadIdx.x, hw = tid >> 4, j = tid & 15;
    const long long slot = (long long)blockIdx.x * 4 + hw;
    if (slot >= (P.act.perm ? (long long)*P.act.count : P.Bsz)) return;
    const long long b = P.act.perm ? (long long)P.act.perm[slot] : slot;
    if (P.S.converged && P.S.J && P.S.converged[b]) return;  // frozen problem (half-warp uniform)
    const unsigned hmask = 0xFFFFu << (tid & 16);
    const int N = P.N;
    const T dt = T(P.dt);
    const T* x0 = reinterpret_cast<const T*>(P.x0) + b * n;
    T* xCur = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * n;
    T* uCur = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * m;
    const long long per = (long long)(N + 1) * n + (long long)N * m;

    // this lane's (up to three) 16-byte chunks of a step: source pointer at k = 0, words per step, offset in the stage
    const T* csrc[3];
    int cstr[3], coff[3];
    bool cval[3];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
        const int c = j + 16 * r;
        cval[r] = c < NCH;
        coff[r] = c * EPC;
        if (c < C_L) {
            csrc[r] = reinterpret_cast<const T*>(P.L) + b * (long long)N * 48 + c * EPC;
            cstr[r] = 48;
        } else if (c < C_L + C_l) {
            csrc[r] = reinterpret_cast<const T*>(P.l) + b * (long long)N * 4 + (c - C_L) * EPC;
            cstr[r] = 4;
        } else if (c < C_L + C_l + C_X) {
            csrc[r] = xCur + (c - C_L - C_l) * EPC;
            cstr[r] = 12;
        } else {
            csrc[r] = uCur + (c - C_L - C_l - C_X) * EPC;
            cstr[r] = 4;
        }
    }
    T(*rg)[ST] = ring[hw];
    auto issue = [&](int k) {
        T* dst = rg[k % D];
#pragma unroll
        for (int r = 0; r < 3; ++r)
            if (cval[r]) fwd_cp16(dst + coff[r], csrc[r] + (long long)k * cstr[r]);
    };

    // running-cost weights: diagonals in registers (the caller asserted ZB_COST_DIAGONAL), else dense blocks through L1
    T qd[n], rd[m];
    const T* Qg = P.C.Q.at<T>(b);
    const T* Rg = P.C.R.at<T>(b);
    if (QDIAG) {
#pragma unroll
        for (int i = 0; i < n; ++i) qd[i] = Qg[i * 13];
#pragma unroll
        for (int i = 0; i < m; ++i) rd[i] = Rg[i * 5];
    }
    T alpha = T(1);
    for (int i = 0; i < j; ++i) alpha *= T(0.5);  // 0.5**j exactly (ilqrUtils.py:145)
    bool write = false;
    T *wx = nullptr, *wu = nullptr;
    if (P.spec != nullptr && j < 2) {
        T* base = reinterpret_cast<T*>(P.spec) + (long long)j * P.Bsz * per;
        wx = base + b * (long long)(N + 1) * n;
        wu = base + P.Bsz * (long long)(N + 1) * n + b * (long long)N * m;
        write = true;
    }
    int idx = 0;
    T Jn = T(0);
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
        for (int s = 0; s < D - 1; ++s) {
            if (s < N) issue(s);
            fwd_commit();
        }
        T x[n], u[m], xd[n];
        T J = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) x[i] = x0[i];
        for (int k = 0; k < N; ++k) {
            fwd_wait<D - 2>();
            __syncwarp(hmask);  // step k's operands have landed for every lane; stage (k-1)%D is free again
            if (k + D - 1 < N) issue(k + D - 1);
            fwd_commit();
            const T* st = rg[k % D];
            T dx[n];
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                T xp[4];
                ld4(st + 52 + 4 * q, xp);
#pragma unroll
                for (int c = 0; c < 4; ++c) dx[4 * q + c] = x[4 * q + c] - xp[c];
            }
            T lk[4], up[4];
            ld4(st + 48, lk);
            ld4(st + 64, up);
#pragma unroll
            for (int i = 0; i < m; ++i) {
                T s = T(0);
#pragma unroll
                for (int q = 0; q < 3; ++q) {
                    T Lr[4];
                    ld4(st + i * 12 + 4 * q, Lr);
#pragma unroll
                    for (int c = 0; c < 4; ++c) s += Lr[c] * dx[4 * q + c];
                }
                u[i] = (alpha * lk[i] + s) + up[i];
            }
            if (write) {
                st4(wx + (long long)k * n, x);
                st4(wx + (long long)k * n + 4, x + 4);
                st4(wx + (long long)k * n + 8, x + 8);
                st4(wu + (long long)k * m, u);
            }
            if (QDIAG) {
                T a = T(0), c = T(0);
#pragma unroll
                for (int i = 0; i < n; ++i) a += x[i] * (qd[i] * x[i]);
#pragma unroll
                for (int i = 0; i < m; ++i) c += u[i] * (rd[i] * u[i]);
                J += a + c;
            } else {
                J += quad_form<T>(Qg, x, n) + quad_form<T>(Rg, u, m);
            }
            QuadTrig<T> tr = quad_trig(x);
            quad_xdot(tr, x, u, xd);
#pragma unroll
            for (int i = 0; i < n; ++i) x[i] = x[i] + dt * xd[i];
        }
        if (write) {
            st4(wx + (long long)N * n, x);
            st4(wx + (long long)N * n + 4, x + 4);
            st4(wx + (long long)N * n + 8, x + 8);
        }
        J += quad_form<T>(P.C.Qf.at<T>(b), x, n);
        fwd_wait<0>();
        __syncwarp(hmask);  // speculative stores visible to the half-warp; ring idle
        if (pass == 1) break;
        // ---- argmin over the 16 step sizes (ilqrUtils.py:147) ----
        if (P.Jall) reinterpret_cast<T*>(P.Jall)[b * 16 + j] = J;
        T Ja[16];
#pragma unroll
        for (int q = 0; q < 16; ++q) Ja[q] = __shfl_sync(hmask, J, q, 16);
        idx = argmin16<T>(Ja);
        Jn = Ja[0];
#pragma unroll
        for (int q = 1; q < 16; ++q) Jn = (idx == q) ? Ja[q] : Jn;
        if (P.spec != nullptr && idx < 2) {
            // the winner was stored speculatively: copy it over the current trajectory, 16 bytes per lane
            const T* base = reinterpret_cast<const T*>(P.spec) + (long long)idx * P.Bsz * per;
            const int4* sx = reinterpret_cast<const int4*>(base + b * (long long)(N + 1) * n);
            const int4* su = reinterpret_cast<const int4*>(base + P.Bsz * (long long)(N + 1) * n + b * (long long)N * m);
            int4* dxp = reinterpret_cast<int4*>(xCur);
            int4* dup = reinterpret_cast<int4*>(uCur);
            const int nx4 = (N + 1) * n / EPC, nu4 = N * m / EPC;
            int i = j;
            for (; i + 48 < nx4; i += 64) {
                const int4 a0 = sx[i], a1 = sx[i + 16], a2 = sx[i + 32], a3 = sx[i + 48];
                dxp[i] = a0; dxp[i + 16] = a1; dxp[i + 32] = a2; dxp[i + 48] = a3;
            }
            for (; i < nx4; i += 16) dxp[i] = sx[i];
            i = j;
            for (; i + 48 < nu4; i += 64) {
                const int4 a0 = su[i], a1 = su[i + 16], a2 = su[i + 32], a3 = su[i + 48];
                dup[i] = a0; dup[i + 16] = a1; dup[i + 32] = a2; dup[i + 48] = a3;
            }
            for (; i < nu4; i += 16) dup[i] = su[i];
            break;
        }
        // another step size won: every lane re-runs it (same arithmetic), lane 0 stores in place.  Row k of the current
        // trajectory is overwritten only after its copy has landed in the ring.
        alpha = T(1);
        for (int i = 0; i < idx; ++i) alpha *= T(0.5);
        write = (j == 0);
        wx = xCur;
        wu = uCur;
    }
    if (j != 0) return;
    if (P.S.J) {
        T* Jc = reinterpret_cast<T*>(P.S.J);
        const T dJ = Jc[b] - Jn;
        P.S.converged[b] = (fabs(dJ) <= T(P.S.tol)) ? 1 : 0;  // NaN compares false, as in the reference
        Jc[b] = Jn;
        P.S.iters[b] = P.S.it + 1;
        if (P.S.alpha_log) P.S.alpha_log[b * (long long)P.S.maxIter + P.S.it] = idx;
        if (P.S.J_log) reinterpret_cast<T*>(P.S.J_log)[b * (long long)(P.S.maxIter + 1) + P.S.it + 1] = Jn;
    } else {
        reinterpret_cast<T*>(P.S.J_out)[b] = Jn;
        if (P.S.idx_out) P.S.idx_out[b] = idx;
    }
}

// Once per solve, quadcopter + diagonal costs: (i) conditionQuadraticCost / conditionValueFunction (ilqrUtils.py:222-234,
// 254-257) of a diagonal Hessian -- its eigenvectors are the identity, so V max(L, eps) V' = diag(max(2 q_i, eps)) exactly,
// which is also what the generic Jacobi routine returns (no rotation is ever applied); (ii) the initial rollout
// u_k = uGuess_k (ilqrUtils.py:292-298) with the state in registers; (iii) the solver state.  Same arithmetic as
// k_solve_prep + k_solve_init (zb_api.cu), 10x faster than their run-time-sized local-memory bodies.
template <typename T>
__global__ void __launch_bounds__(64) k_solve_setup_quad(SetupQuadP P) {
    constexpr int n = 12, m = 4;
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= P.Bsz) return;
    const int N = P.N;
    const T eps = T(P.eps), dt = T(P.dt);
    const T* Q = P.C.Q.at<T>(b);
    const T* R = P.C.R.at<T>(b);
    const T* Qf = P.C.Qf.at<T>(b);
    T qd[n], rd[m];
#pragma unroll
    for (int i = 0; i < n; ++i) qd[i] = Q[i * 13];
#pragma unroll
    for (int i = 0; i < m; ++i) rd[i] = R[i * 5];
    {
        T* cz = reinterpret_cast<T*>(P.Czz) + b * 256;
        for (int e = 0; e < 256; ++e) cz[e] = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) { const T h = qd[i] + qd[i]; cz[i * 17] = h > eps ? h : eps; }
#pragma unroll
        for (int i = 0; i < m; ++i) { const T h = rd[i] + rd[i]; cz[(n + i) * 17] = h > eps ? h : eps; }
        T* vf = reinterpret_cast<T*>(P.Vfxx) + b * 144;
        for (int e = 0; e < 144; ++e) vf[e] = T(0);
#pragma unroll
        for (int i = 0; i < n; ++i) { const T h = Qf[i * 13] + Qf[i * 13]; vf[i * 13] = h > eps ? h : eps; }
    }
    const T* x0 = reinterpret_cast<const T*>(P.x0) + b * n;
    const T* ug = reinterpret_cast<const T*>(P.uGuess) + b * (long long)N * m;
    T* xT = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * n;
    T* uT = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * m;
    T x[n], u[m], xd[n];
    T Jc = T(0);
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = x0[i];
    T un[m];
    if (N > 0) ld4(ug, un);
    for (int k = 0; k < N; ++k) {
#pragma unroll
        for (int i = 0; i < m; ++i) u[i] = un[i];
        if (k + 1 < N) ld4(ug + (long long)(k + 1) * m, un);
        st4(xT + (long long)k * n, x);
        st4(xT + (long long)k * n + 4, x + 4);
        st4(xT + (long long)k * n + 8, x + 8);
        st4(uT + (long long)k * m, u);
        {   // quad_form(Q, x) + quad_form(R, u) of diagonal Q, R: the off-diagonal terms of the dense form add exact zeros
            T a = T(0), c = T(0);
#pragma unroll
            for (int i = 0; i < n; ++i) a += x[i] * (qd[i] * x[i]);
#pragma unroll
            for (int i = 0; i < m; ++i) c += u[i] * (rd[i] * u[i]);
            Jc += a + c;
        }
        QuadTrig<T> tr = quad_trig(x);
        quad_xdot(tr, x, u, xd);
#pragma unroll
        for (int i = 0; i < n; ++i) x[i] = x[i] + dt * xd[i];
    }
    st4(xT + (long long)N * n, x);
    st4(xT + (long long)N * n + 4, x + 4);
    st4(xT + (long long)N * n + 8, x + 8);
    Jc += quad_form<T>(Qf, x, n);
    reinterpret_cast<T*>(P.J)[b] = Jc;
    P.converged[b] = 0;
    P.iters[b] = 0;
    if (P.alpha_log)
        for (int i = 0; i < P.maxIter; ++i) P.alpha_log[b * (long long)P.maxIter + i] = -1;
    if (P.J_log) {
        T* jl = reinterpret_cast<T*>(P.J_log) + b * (long long)(P.maxIter + 1);
        jl[0] = Jc;
        for (int i = 1; i <= P.maxIter; ++i) jl[i] = Jc * T(0) + T(NAN);
    }
}

__global__ void k_compact_active(long long Bsz, const uint8_t* __restrict__ converged, int32_t* __restrict__ perm, int32_t* __restrict__ count) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const bool act = b < Bsz && converged[b] == 0;
    // one atomic per warp: the active lanes take consecutive positions
    const unsigned m = __ballot_sync(0xffffffffu, act);
    if (m == 0u) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(count, __popc(m));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (act) perm[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)b;
}

int32_t compact_active_launch(long long Bsz, const uint8_t* converged, int32_t* perm_out, int32_t* count_out, cudaStream_t stream) {
    ZB_CUDA(cudaMemsetAsync(count_out, 0, sizeof(int32_t), stream));
    k_compact_active<<<(unsigned)((Bsz + 255) / 256), 256, 0, stream>>>(Bsz, converged, perm_out, count_out);
    ZB_CUDA(cudaGetLastError());
    return 0;
}

int32_t solve_setup_quad_launch(int32_t dtype, const SetupQuadP& P, cudaStream_t stream) {
    const unsigned grid = (unsigned)((P.Bsz + 63) / 64);
    if (dtype == ZB_F32) k_solve_setup_quad<float><<<grid, 64, 0, stream>>>(P);
    else k_solve_setup_quad<double><<<grid, 64, 0, stream>>>(P);
    ZB_CUDA(cudaGetLastError());
    return 0;
}

int32_t fwd_quad_launch(int32_t dtype, const FwdQuadP& P, cudaStream_t stream) {
    const unsigned grid = (unsigned)((P.Bsz + 3) / 4);
    if (P.cost_diagonal) {
        if (dtype == ZB_F32) k_forward_quad<float, true><<<grid, 64, 0, stream>>>(P);
        else k_forward_quad<double, true><<<grid, 64, 0, stream>>>(P);
    } else {
        if (dtype == ZB_F32) k_forward_quad<float, false><<<grid, 64, 0, stream>>>(P);
        else k_forward_quad<double, false><<<grid, 64, 0, stream>>>(P);
    }
    ZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace zb
