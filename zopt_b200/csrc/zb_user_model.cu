// Plug-in translation unit: the generic iLQR / DDP solver, rollout and line search compiled around ONE user-defined
// dynamics model.  The reference's solvers take an arbitrary Python callable and differentiate it with JAX
// (zopt/ilqrUtils.py:260-268, pytrees.py:138-194); a kernel cannot trace a lambda, so zopt_b200/plugin.py turns a symbolic
// (sympy) definition of x+ = f(x, u) into CUDA -- f, its Jacobians and the costate-contracted Hessian, common
// subexpressions shared -- writes it to the header named by ZB_USER_MODEL_HEADER and compiles this file with nvcc for
// sm_100a into its own shared library.  Entry points mirror zb_ilqr_solve / zb_ilqr_rollout of include/zopt_b200.h.
#include <stdint.h>

#include "zb_math.cuh"
#define ZB_USER_MODEL 1
#include ZB_USER_MODEL_HEADER  // user_step<T>, user_lin<T>, user_hess<T>, ZB_USER_N, ZB_USER_M

#include "ilqr_generic.cuh"

using namespace zb;

#define ZB_DISPATCH(dtype, KERNEL, grid, block, stream, ...)                                                   \
    do {                                                                                                        \
        if ((dtype) == ZB_F32) KERNEL<float><<<(grid), (block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__);       \
        else KERNEL<double><<<(grid), (block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__);                        \
        ZB_CUDA(cudaGetLastError());                                                                            \
    } while (0)

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static int32_t user_common(int32_t dtype, int64_t Bsz, int32_t N, const zb_cost* cost, bool need_cost, RollP& P) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "dtype must be ZB_F32 or ZB_F64 (got %d)", dtype);
    ZB_ARG(Bsz >= 0 && N >= 0, "negative size");
    P.Bsz = Bsz;
    P.N = N;
    P.M = Model{};
    P.M.kind = 2;
    P.M.n = ZB_USER_N;
    P.M.m = ZB_USER_M;
    P.C = to_cost(cost);
    P.has_cost = (cost != nullptr);
    if (need_cost) ZB_ARG(cost && cost->Q.ptr && cost->R.ptr && cost->Qf.ptr, "cost (Q,R,Qf) required");
    return 0;
}

// x+ = f(x, u) for a batch of points (parity of the generated code against the oracle's evaluation of the same expressions)
template <typename T>
__global__ void k_user_step(long long Bsz, const void* x, const void* u, void* xn, void* fx, void* fu) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= Bsz) return;
    constexpr int n = ZB_USER_N, m = ZB_USER_M;
    const T* xp = reinterpret_cast<const T*>(x) + b * n;
    const T* up = reinterpret_cast<const T*>(u) + b * m;
    T xl[NX], ul[NU], o[NX], jx[NX * NX], ju[NX * NU];
    for (int i = 0; i < n; ++i) xl[i] = xp[i];
    for (int i = 0; i < m; ++i) ul[i] = up[i];
    user_step<T>(xl, ul, o);
    for (int i = 0; i < n; ++i) reinterpret_cast<T*>(xn)[b * n + i] = o[i];
    if (fx && fu) {
        user_lin<T>(xl, ul, jx, ju);
        for (int i = 0; i < n * n; ++i) reinterpret_cast<T*>(fx)[b * n * n + i] = jx[i];
        for (int i = 0; i < n * m; ++i) reinterpret_cast<T*>(fu)[b * n * m + i] = ju[i];
    }
}

// Z (Bsz,p,p), p = n+m: sum_i lam_i d2 f_i / dz2 at (x, u), z = [x; u]  (the contraction of QuadraticDynamics' f_xx, f_ux, f_uu)
template <typename T>
__global__ void k_user_hess(long long Bsz, const void* x, const void* u, const void* lam, void* Z) {
    const long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= Bsz) return;
    constexpr int n = ZB_USER_N, m = ZB_USER_M, p = n + m;
    T xl[NX], ul[NU], ll[NX], H[(NX + NU) * (NX + NU)];
    for (int i = 0; i < n; ++i) { xl[i] = reinterpret_cast<const T*>(x)[b * n + i]; ll[i] = reinterpret_cast<const T*>(lam)[b * n + i]; }
    for (int i = 0; i < m; ++i) ul[i] = reinterpret_cast<const T*>(u)[b * m + i];
    user_hess<T>(xl, ul, ll, H);
    for (int i = 0; i < p * p; ++i) reinterpret_cast<T*>(Z)[b * p * p + i] = H[i];
}

extern "C" {

__attribute__((visibility("default"))) int32_t zb_user_hess(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x,
                                                            const void* u, const void* lam, void* Z) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "bad dtype %d", dtype);
    ZB_ARG(Bsz >= 0 && (Bsz == 0 || (x && u && lam && Z)), "bad operand");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    ZB_DISPATCH(dtype, k_user_hess, gen_grid(Bsz), GEN_THREADS, stream, (long long)Bsz, x, u, lam, Z);
    return 0;
}

__attribute__((visibility("default"))) int32_t zb_user_dims(int32_t* n, int32_t* m) {
    *n = ZB_USER_N;
    *m = ZB_USER_M;
    return 0;
}

__attribute__((visibility("default"))) int32_t zb_user_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return -1;
    strncpy(buf, err_buf(), len - 1);
    buf[len - 1] = 0;
    return 0;
}

__attribute__((visibility("default"))) int32_t zb_user_step(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x,
                                                            const void* u, void* xn, void* fx, void* fu) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "bad dtype %d", dtype);
    ZB_ARG(Bsz >= 0 && (Bsz == 0 || (x && u && xn)), "bad operand");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    ZB_DISPATCH(dtype, k_user_step, gen_grid(Bsz), GEN_THREADS, stream, (long long)Bsz, x, u, xn, fx, fu);
    return 0;
}

// zb_ilqr_rollout (include/zopt_b200.h) for the user model
__attribute__((visibility("default"))) int32_t zb_user_rollout(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N,
                                                               const zb_cost* cost, const void* x0, const void* l, const void* L,
                                                               const void* xPrev, const void* uPrev, double alpha, void* xTraj,
                                                               void* uTraj, void* J_out) {
    RollP P;
    int32_t rc = user_common(dtype, Bsz, N, cost, J_out != nullptr, P);
    if (rc) return rc;
    ZB_ARG(x0 && xTraj && (N == 0 || (l && L && xPrev && uPrev && uTraj)), "NULL operand");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    P.x0 = x0; P.l = l; P.L = L; P.xPrev = xPrev; P.uPrev = uPrev;
    P.xTraj = xTraj; P.uTraj = uTraj; P.J = J_out;
    if (!J_out) P.has_cost = 0;
    ZB_DISPATCH(dtype, k_rollout, gen_grid(Bsz), GEN_THREADS, stream, P, alpha);
    return 0;
}

__attribute__((visibility("default"))) size_t zb_user_ilqr_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N) {
    const size_t e = dtype == ZB_F64 ? 8 : 4, n = ZB_USER_N, m = ZB_USER_M, p = n + m;
    const size_t per = (size_t)(N + 1) * n + (size_t)N * m;
    return align256(e * Bsz * N * m) + align256(e * Bsz * 16) + align256(e * Bsz * p * p) + align256(e * Bsz * n * n) +
           align256(e * SPEC_N * Bsz * per) + 256;
}

// zb_ilqr_solve (include/zopt_b200.h) for the user model: same arguments minus the model, same outputs
__attribute__((visibility("default"))) int32_t zb_user_ilqr_solve(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N,
                                                                  int32_t flags, const zb_cost* cost, const void* x0,
                                                                  const void* uGuess, int32_t maxIter, double tol, void* xTraj,
                                                                  void* uTraj, void* L_out, void* J_out, uint8_t* converged_out,
                                                                  int32_t* iters_out, int32_t* alpha_log, void* J_log,
                                                                  void* workspace, size_t workspace_bytes) {
    const int32_t second_order = flags & ZB_SECOND_ORDER;
    RollP P;
    int32_t rc = user_common(dtype, Bsz, N, cost, true, P);
    if (rc) return rc;
    ZB_ARG(maxIter >= 0, "negative maxIter");
    ZB_ARG(x0 && uGuess && xTraj && uTraj && L_out && J_out && converged_out && iters_out, "NULL operand");
    const int n = ZB_N_OF(P.M.n), m = ZB_M_OF(P.M.m), p = n + m;
    const size_t need = zb_user_ilqr_workspace_bytes(dtype, Bsz, N);
    ZB_ARG(workspace && workspace_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    const size_t e = dtype == ZB_F64 ? 8 : 4;
    char* w = reinterpret_cast<char*>(workspace);
    void* l_ws = w;  w += align256(e * Bsz * N * m);
    void* Jall = w;  w += align256(e * Bsz * 16);
    void* Czz = w;   w += align256(e * Bsz * p * p);
    void* Vfxx = w;  w += align256(e * Bsz * n * n);
    void* spec = w;
    cudaStream_t s = (cudaStream_t)stream;
    ZB_CUDA(cudaMemsetAsync(L_out, 0, e * Bsz * N * m * n, s));  // policy.L = 0 before the first iteration (ilqrUtils.py:293)
    P.x0 = x0; P.l = l_ws; P.L = L_out; P.xPrev = xTraj; P.uPrev = uTraj;
    P.xTraj = xTraj; P.uTraj = uTraj; P.J = nullptr;
    ZB_DISPATCH(dtype, k_solve_prep, gen_grid(Bsz), GEN_THREADS, stream, (long long)Bsz, n, m, P.C, 1e-3, Czz, Vfxx);
    ZB_DISPATCH(dtype, k_solve_init, gen_grid(Bsz), GEN_THREADS, stream, P, uGuess, J_out, converged_out, iters_out, alpha_log,
                J_log, (int)maxIter);
    SolveBackP Bk{Bsz, N, second_order, P.M, P.C, xTraj, uTraj, Czz, Vfxx, converged_out, l_ws, L_out, 1e-3};
    for (int it = 0; it < maxIter; ++it) {
        ZB_DISPATCH(dtype, k_solve_backward, gen_grid(Bsz), GEN_THREADS, stream, Bk);
        ZB_DISPATCH(dtype, k_forward_costs, gen_grid(Bsz * 16), GEN_THREADS, stream, P, Jall, (const uint8_t*)converged_out, spec);
        CommitP S{J_out, converged_out, iters_out, alpha_log, J_log, it, (int)maxIter, tol, nullptr, nullptr};
        ZB_DISPATCH(dtype, k_forward_commit, gen_grid(Bsz * 16), GEN_THREADS, stream, P, (const void*)Jall, S, (const void*)spec);
    }
    return 0;
}

}  // extern "C"
