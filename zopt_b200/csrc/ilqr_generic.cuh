// Generic (run-time n <= 16, m <= 8, any registered model) kernels of the iLQR / DDP solvers and of trajectoryRollout /
// forwardPass2: one thread per problem (16 per problem in the line search), bodies in zb_problems.cuh.
// Included by zb_api.cu (the library) and by zb_user_model.cu (a user-defined model compiled as a plug-in, models.py).
#pragma once
#include "ilqr_params.cuh"

namespace zb {

template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_rollout(RollP P, double alpha) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= P.Bsz) return;
    T J = rollout_any<T>(P, b, T(alpha), true);
    if (P.J) reinterpret_cast<T*>(P.J)[b] = J;
}

// forwardPass2 phase 1: 16 threads per problem, one step size each (ilqrUtils.py:139-146).  All lanes compute the cost;
// the lanes of the two largest step sizes (alpha = 1, 1/2: the usual winners) also store their trajectories into the
// speculative buffers `spec` ((2,Bsz,(N+1)n + Nm)) so that the commit can copy the winner instead of re-running it.
constexpr int SPEC_N = 2;
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_forward_costs(RollP P, void* Jall, const uint8_t* done, void* spec) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long b = t >> 4;
    int j = (int)(t & 15);
    if (b >= P.Bsz) return;
    if (done && done[b]) return;
    T alpha = T(1);
    for (int i = 0; i < j; ++i) alpha *= T(0.5);  // 0.5**j exactly (ilqrUtils.py:145)
    // one call for all 16 lanes (no divergence): the stores are predicated on `write`
    const bool write = spec != nullptr && j < SPEC_N;
    RollP Q = P;  // same inputs; outputs redirected so that problem b lands at spec[j][b]
    if (write) {
        const long long per = (long long)(P.N + 1) * P.M.n + (long long)P.N * P.M.m;
        T* base = reinterpret_cast<T*>(spec) + (long long)j * P.Bsz * per;
        Q.xTraj = base;
        Q.uTraj = base + P.Bsz * (long long)(P.N + 1) * P.M.n;
    }
    const T J = rollout_any<T>(Q, b, alpha, write);
    reinterpret_cast<T*>(Jall)[b * 16 + j] = J;
}

// forwardPass2 phase 2: argmin over the 16 costs, re-run the winning rollout and store it
// (ilqrUtils.py:147-150), plus the solver's bookkeeping (ilqrUtils.py:318-321) when `S.J` is set (CommitP: ilqr_forward.cuh).

// 16 lanes per problem: if the winner's trajectory was stored speculatively the lanes copy it (coalesced), otherwise
// lane 0 re-runs the winning rollout in place (identical arithmetic, so identical values).
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_forward_commit(RollP P, const void* Jall, CommitP S, const void* spec) {
    long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    long long b = t >> 4;
    const int lane16 = (int)(t & 15);
    if (b >= P.Bsz) return;
    if (S.converged && S.J && S.converged[b]) return;
    const T* Ja = reinterpret_cast<const T*>(Jall) + b * 16;
    const int idx = argmin16<T>(Ja);
    const int n = ZB_N_OF(P.M.n), m = ZB_M_OF(P.M.m), N = P.N;
    if (spec && idx < SPEC_N) {
        const long long per = (long long)(N + 1) * n + (long long)N * m;
        const T* base = reinterpret_cast<const T*>(spec) + (long long)idx * P.Bsz * per;
        const T* sx = base + b * (long long)(N + 1) * n;
        const T* su = base + P.Bsz * (long long)(N + 1) * n + b * (long long)N * m;
        T* dx = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * n;
        T* du = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * m;
        for (int i = lane16; i < (N + 1) * n; i += 16) dx[i] = sx[i];
        for (int i = lane16; i < N * m; i += 16) du[i] = su[i];
    } else if (lane16 == 0) {
        T alpha = T(1);
        for (int i = 0; i < idx; ++i) alpha *= T(0.5);
        rollout_any<T>(P, b, alpha, true);
    }
    if (lane16 != 0) return;
    T Jn = Ja[idx];
    if (S.J) {
        T* J = reinterpret_cast<T*>(S.J);
        T dJ = J[b] - Jn;
        S.converged[b] = (fabs(dJ) <= T(S.tol)) ? 1 : 0;  // NaN compares false, as in the reference
        J[b] = Jn;
        S.iters[b] = S.it + 1;
        if (S.alpha_log) S.alpha_log[b * (long long)S.maxIter + S.it] = idx;
        if (S.J_log) reinterpret_cast<T*>(S.J_log)[b * (long long)(S.maxIter + 1) + S.it + 1] = Jn;
    } else {
        reinterpret_cast<T*>(S.J_out)[b] = Jn;
        if (S.idx_out) S.idx_out[b] = idx;
    }
}

// iLQR/DDP solve: conditioned cost blocks (ilqrUtils.py:312-313), constant for quadratic costs
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_solve_prep(long long Bsz, int n, int m, Cost C, double eps, void* Czz, void* Vfxx) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= Bsz) return;
#ifdef ZB_USER_COST
    return;  // a user-defined cost is expanded and conditioned per step inside the backward pass
#endif
    const int p = n + m;
    T S[ZB_PD_MAX * ZB_PD_MAX], W[ZB_PD_MAX * ZB_PD_MAX];
    const T* Q = C.Q.at<T>(b);
    const T* R = C.R.at<T>(b);
    const T* Qf = C.Qf.at<T>(b);
    for (int i = 0; i < p * p; ++i) S[i] = T(0);
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) S[i * p + j] = Q[i * n + j] + Q[j * n + i];
    for (int i = 0; i < m; ++i)
        for (int j = 0; j < m; ++j) S[(n + i) * p + n + j] = R[i * m + j] + R[j * m + i];
    pd_clamp<T>(S, W, p, T(eps));
    T* o = reinterpret_cast<T*>(Czz) + b * (long long)p * p;
    for (int i = 0; i < p * p; ++i) o[i] = S[i];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) S[i * n + j] = Qf[i * n + j] + Qf[j * n + i];
    pd_clamp<T>(S, W, n, T(eps));
    o = reinterpret_cast<T*>(Vfxx) + b * (long long)n * n;
    for (int i = 0; i < n * n; ++i) o[i] = S[i];
}

// initial rollout u_k = uGuess_k (ilqrUtils.py:292-298) and solver state initialisation
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_solve_init(RollP P, const void* uGuess, void* J, uint8_t* converged, int32_t* iters,
                             int32_t* alpha_log, void* J_log, int maxIter) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= P.Bsz) return;
    const int n = ZB_N_OF(P.M.n), m = ZB_M_OF(P.M.m), N = P.N;
    const T* x0 = reinterpret_cast<const T*>(P.x0) + b * n;
    const T* ug = reinterpret_cast<const T*>(uGuess) + b * (long long)N * m;
    T* xT = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * n;
    T* uT = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * m;
    const T* Q = P.C.Q.at<T>(b);
    const T* R = P.C.R.at<T>(b);
    T x[NX], u[NU];
    T Jc = T(0);
    for (int i = 0; i < n; ++i) x[i] = x0[i];
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < m; ++i) u[i] = ug[(long long)k * m + i];
        for (int i = 0; i < n; ++i) xT[(long long)k * n + i] = x[i];
        for (int i = 0; i < m; ++i) uT[(long long)k * m + i] = u[i];
#ifdef ZB_USER_COST
        Jc += user_cost<T>(x, u);
#else
        Jc += quad_form<T>(Q, x, n) + quad_form<T>(R, u, m);
#endif
        model_step<T>(P.M, b, x, u, x);
    }
    for (int i = 0; i < n; ++i) xT[(long long)N * n + i] = x[i];
#ifdef ZB_USER_COST
    Jc += user_tcost<T>(x);
    (void)Q; (void)R;
#else
    Jc += quad_form<T>(P.C.Qf.at<T>(b), x, n);
#endif
    reinterpret_cast<T*>(J)[b] = Jc;
    converged[b] = 0;
    iters[b] = 0;
    if (alpha_log)
        for (int i = 0; i < maxIter; ++i) alpha_log[b * (long long)maxIter + i] = -1;
    if (J_log) {
        T* jl = reinterpret_cast<T*>(J_log) + b * (long long)(maxIter + 1);
        jl[0] = Jc;
        for (int i = 1; i <= maxIter; ++i) jl[i] = Jc * T(0) + T(NAN);
    }
}

template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_solve_backward(SolveBackP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < P.Bsz) solve_backward_problem<T>(P, b);
}

}  // namespace zb
