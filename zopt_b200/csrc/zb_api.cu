// C ABI entry points (include/zopt_b200.h) and the generic one-thread-per-problem kernels.
// The (n,m)=(12,4) fast kernels live in lqr_fast.cuh / ilqr_fast.cuh and are dispatched from here.
#include <type_traits>

#include "zb_common.cuh"
#include "lqr_fast.cuh"
#include "ilqr_params.cuh"
#include "ilqr_generic.cuh"
#include "lqr_t1.cuh"
#include "lqr_s84.cuh"
#include "lqr_s84d.cuh"
#include "mpc_coop.cuh"
#include "mpc_box.cuh"
#include "mpc_box_quad.cuh"

using namespace zb;

namespace zb {
int32_t mpc_closed_loop_w9_launch(const t1::ClosedLoopP& P, cudaStream_t stream);  // zb_small_batch.cu
}

#define ZB_FORCE_RUNTIME_SIZES ZB_FORCE_GENERIC
#ifndef ZB_EXIT_CHECK
#define ZB_EXIT_CHECK 4  // zb_ilqr_solve: iterations between two "is anything still iterating" reads
#endif
#ifndef ZB_W9_MAX_PER_SM
#define ZB_W9_MAX_PER_SM 68  // problems per SM up to which the nine-lanes-per-problem closed-loop kernel is selected (measured crossover with
                            // the thread-per-problem kernel on B200: ~10 K problems; it is ahead of the 4-thread kernel at every size)
#endif

// =================================================================================================
// generic kernels: one thread per problem
// =================================================================================================
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_lqr_generic(LqrP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < P.Bsz) lqr_problem<T>(P, b);
}

template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_bilinear_generic(BilinP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < P.Bsz) bilinear_problem<T>(P, b);
}

// compile-time (n, m) variants of the two kernels above (zb_steps.cuh): the shapes of the reference's demos and tests
template <typename T, int N_, int M_>
__global__ void __launch_bounds__(GEN_THREADS) k_lqr_ct(LqrP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < P.Bsz) lqr_problem_ct<T, N_, M_>(P, b);
}
template <typename T, int N_, int M_>
__global__ void __launch_bounds__(GEN_THREADS) k_bilinear_ct(BilinP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < P.Bsz) bilinear_problem_ct<T, N_, M_>(P, b);
}
#define ZB_CT_SHAPES(X) X(8, 4) X(2, 1) X(2, 2)
// fp64 at n = 8 does not fit one thread's registers (V, A - BL and V(A - BL) alone are 384): measured slower than the
// run-time-size kernel (0.9 vs 1.2 M solves/s), so fp64 keeps the compile-time variant for the tiny shapes only
#define ZB_CT_F64_OK(N_) ((N_) <= 4)
template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_care_rk4(CareP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < P.Bsz) care_problem<T>(P, b);
}

struct QuadP {
    long long Bsz;
    const void *x, *u, *lam;
    void *o1, *o2;
    double dt, wind[3];
    int has_wind;
};

template <typename T>
__global__ void k_quad_dynamics(QuadP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= P.Bsz) return;
    T x[12], u[4], xd[12], w[3] = {T(P.wind[0]), T(P.wind[1]), T(P.wind[2])};
    for (int i = 0; i < 12; ++i) x[i] = reinterpret_cast<const T*>(P.x)[b * 12 + i];
    for (int i = 0; i < 4; ++i) u[i] = reinterpret_cast<const T*>(P.u)[b * 4 + i];
    quad_F<T>(x, u, w, P.has_wind != 0, xd);
    for (int i = 0; i < 12; ++i) reinterpret_cast<T*>(P.o1)[b * 12 + i] = xd[i];
}

template <typename T>
__global__ void k_quad_linearize(QuadP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= P.Bsz) return;
    T x[12], u[4], fx[144], w[3] = {T(P.wind[0]), T(P.wind[1]), T(P.wind[2])};
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = reinterpret_cast<const T*>(P.x)[b * 12 + i];
#pragma unroll
    for (int i = 0; i < 4; ++i) u[i] = reinterpret_cast<const T*>(P.u)[b * 4 + i];
    QuadTrig<T> tr = quad_trig(x);
    if (P.has_wind) quad_jac_x_wind(tr, x, u, w, fx);
    else quad_jac_x(tr, x, u, fx);
    const T s = (P.dt != 0.0) ? T(P.dt) : T(1), one = (P.dt != 0.0) ? T(1) : T(0);
    // 16-byte vector stores of the row-major blocks (A: 12x12, B: 12x4; both 16-byte aligned per problem)
    using V = typename std::conditional<sizeof(T) == 4, float4, double2>::type;
    constexpr int PER = 16 / sizeof(T);
    V* A = reinterpret_cast<V*>(reinterpret_cast<T*>(P.o1) + b * 144);
#pragma unroll
    for (int q = 0; q < 144 / PER; ++q) {
        T e[PER];
#pragma unroll
        for (int j = 0; j < PER; ++j) {
            const int idx = q * PER + j;
            e[j] = s * fx[idx] + ((idx / 12 == idx % 12) ? one : T(0));
        }
        if (sizeof(T) == 4) A[q] = *reinterpret_cast<V*>(e);
        else A[q] = *reinterpret_cast<V*>(e);
    }
    if (P.o2) {
        V* B = reinterpret_cast<V*>(reinterpret_cast<T*>(P.o2) + b * 48);
#pragma unroll
        for (int q = 0; q < 48 / PER; ++q) {
            T e[PER];
#pragma unroll
            for (int j = 0; j < PER; ++j) {
                const int idx = q * PER + j;  // f_u: [2,0] = -s, [3,1] = [4,2] = [5,3] = +s
                e[j] = (idx == 8) ? -s : ((idx == 13 || idx == 18 || idx == 23) ? s : T(0));
            }
            B[q] = *reinterpret_cast<V*>(e);
        }
    }
}

template <typename T>
__global__ void k_quad_hess(QuadP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= P.Bsz) return;
    T x[12], u[4], lam[12], H[144], w[3] = {T(P.wind[0]), T(P.wind[1]), T(P.wind[2])};
    for (int i = 0; i < 12; ++i) x[i] = reinterpret_cast<const T*>(P.x)[b * 12 + i];
    for (int i = 0; i < 4; ++i) u[i] = reinterpret_cast<const T*>(P.u)[b * 4 + i];
    for (int i = 0; i < 12; ++i) lam[i] = reinterpret_cast<const T*>(P.lam)[b * 12 + i];
    quad_hess<T>(x, u, w, P.has_wind != 0, T(P.dt), lam, H);
    T* o = reinterpret_cast<T*>(P.o1) + b * 144;
    for (int i = 0; i < 144; ++i) o[i] = H[i];
}

template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_backward(BackP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < P.Bsz) backward_problem<T>(P, b);
}

template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_pd_clamp(long long Bsz, int p, double eps, const void* in, void* out) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= Bsz) return;
    T S[ZB_PD_MAX * ZB_PD_MAX], W[ZB_PD_MAX * ZB_PD_MAX];
    const T* src = reinterpret_cast<const T*>(in) + b * (long long)p * p;
    for (int i = 0; i < p * p; ++i) S[i] = src[i];
    pd_clamp<T>(S, W, p, T(eps));
    T* dst = reinterpret_cast<T*>(out) + b * (long long)p * p;
    for (int i = 0; i < p * p; ++i) dst[i] = S[i];
}

// lqrMpc.solve without active bounds: exact Riccati sweep + linear rollout (mpcUtils.py:47-59)
struct MpcP {
    long long Bsz;
    int N, n, m;
    Arr A, B, Q, R, Qf;
    const void* x0;
    void *u0, *xTraj, *uTraj, *gains;  // gains: workspace (Bsz,N,m,n)
    int8_t* status;
    int32_t* iters;
};

template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_mpc_riccati(MpcP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b >= P.Bsz) return;
    const int n = P.n, m = P.m, N = P.N;
    const T *A = P.A.at<T>(b), *B = P.B.at<T>(b), *Q = P.Q.at<T>(b), *R = P.R.at<T>(b), *Qf = P.Qf.at<T>(b);
    T V[NX * NX], L[NU * NX];
    for (int i = 0; i < n * n; ++i) V[i] = Qf[i];
    T* G = reinterpret_cast<T*>(P.gains) + b * (long long)N * m * n;
    for (int k = N - 1; k >= 0; --k) {
        lqr_joseph_step<T>(n, m, A, B, Q, R, V, L);
        for (int i = 0; i < m * n; ++i) G[(long long)k * m * n + i] = L[i];
    }
    const T* x0 = reinterpret_cast<const T*>(P.x0) + b * n;
    T* xT = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * n;
    T* uT = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * m;
    T x[NX], u[NU], xn[NX];
    for (int i = 0; i < n; ++i) x[i] = x0[i];
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < n; ++i) xT[(long long)k * n + i] = x[i];
        for (int i = 0; i < m; ++i) {
            T s = T(0);
            for (int j = 0; j < n; ++j) s += G[((long long)k * m + i) * n + j] * x[j];
            u[i] = -s;
            uT[(long long)k * m + i] = u[i];
        }
        for (int i = 0; i < n; ++i) {
            T s = T(0);
            for (int j = 0; j < n; ++j) s += A[i * n + j] * x[j];
            for (int j = 0; j < m; ++j) s += B[i * m + j] * u[j];
            xn[i] = s;
        }
        for (int i = 0; i < n; ++i) x[i] = xn[i];
    }
    for (int i = 0; i < n; ++i) xT[(long long)N * n + i] = x[i];
    T* u0 = reinterpret_cast<T*>(P.u0) + b * m;
    for (int i = 0; i < m; ++i) u0[i] = uT[i];
    P.status[b] = 0;
    if (P.iters) P.iters[b] = 0;
}

template <typename T>
__global__ void __launch_bounds__(GEN_THREADS) k_mpc_admm(AdmmP P) {
    long long b = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (b < P.Bsz) admm_problem<T>(P, b);
}

// dependent-FMA throughput probe (roofline denominator)
template <typename T, int ILP>
__global__ void k_peak_fma(T* out, int iters) {
    T a[ILP];
    const T b = T(1.0000001), c = T(1e-7);
#pragma unroll
    for (int i = 0; i < ILP; ++i) a[i] = T(threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = a[i] * b + c;
    }
    T s = T(0);
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += a[i];
    if (s == T(-1)) out[0] = s;
}

// =================================================================================================
// C ABI
// =================================================================================================
#define ZB_DISPATCH(dtype, KERNEL, grid, block, stream, ...)                              \
    do {                                                                                    \
        if ((dtype) == ZB_F32) KERNEL<float><<<(grid), (block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__);  \
        else KERNEL<double><<<(grid), (block), 0, (cudaStream_t)(stream)>>>(__VA_ARGS__);   \
        ZB_CUDA(cudaGetLastError());                                                        \
    } while (0)

extern "C" {

int32_t zb_version(void) { return 100; }

int32_t zb_last_error(char* buf, size_t len) {
    if (!buf || len == 0) return -1;
    strncpy(buf, err_buf(), len - 1);
    buf[len - 1] = 0;
    return 0;
}

int32_t zb_device_info(int32_t device, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, size_t* total_mem) {
    cudaDeviceProp p;
    ZB_CUDA(cudaGetDeviceProperties(&p, device));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return 0;
}

int32_t zb_lqr_dfh(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t T, int32_t n,
                   int32_t m, const zb_arr* A, const zb_arr* B, const zb_arr* Q, const zb_arr* R, void* L_out,
                   void* V0_out) {
    return zb_lqr_dfh_flags(dtype, device, stream, Bsz, N, T, n, m, A, B, Q, R, 0, L_out, V0_out);
}

int32_t zb_lqr_dfh_flags(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t T, int32_t n,
                         int32_t m, const zb_arr* A, const zb_arr* B, const zb_arr* Q, const zb_arr* R, int32_t flags,
                         void* L_out, void* V0_out) {
    int32_t rc = check_dims(dtype, Bsz, n, m);
    if (rc) return rc;
    if (Bsz == 0) return 0;  // empty batch: nothing to read or write (pointers may be NULL)
    ZB_ARG(N >= 0 && T >= 1 && T >= N, "need T >= N >= 0 and T >= 1 (got N=%d, T=%d)", N, T);
    ZB_ARG(A && B && Q && R && A->ptr && B->ptr && Q->ptr && R->ptr, "A, B, Q, R must be non-NULL");
    ZB_ARG(L_out != nullptr || N == 0, "L_out is NULL");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    LqrP P{Bsz, N, T, n, m, to_arr(A), to_arr(B), to_arr(Q), to_arr(R), L_out, V0_out};
    if ((flags & ZB_FORCE_GENERIC) && n == 12) {  // e.g. non-symmetric weights: the (12,4) kernels read the lower triangle only
        ZB_DISPATCH(dtype, k_lqr_generic, gen_grid(Bsz), GEN_THREADS, stream, P);
        return 0;
    }
    // the demos' shape (8,4) in fp32, symmetric weights: register-resident thread-per-problem kernel (lqr_s84.cuh); non-symmetric
    // weights (ZB_FORCE_GENERIC) fall through to the as-written compile-time-size kernel below
    if (dtype == ZB_F32 && n == 8 && m == 4 && N >= 1 && !(flags & ZB_FORCE_GENERIC) && arr_ok(P.A) && arr_ok(P.B) && arr_ok(P.Q) &&
        arr_ok(P.R) && aligned16(P.L) && (!P.V0 || aligned16(P.V0)) && !getenv("ZB_FORCE_RUNTIME_SIZES") && !getenv("ZB_NO_S84")) {
        s84::S84P F{};
        F.Bsz = Bsz; F.N = N; F.T = T;
        F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R;
        F.L = reinterpret_cast<float*>(L_out); F.V0 = reinterpret_cast<float*>(V0_out);
        s84::k_riccati_s84<false><<<(unsigned)((Bsz + 31) / 32), 32, 0, (cudaStream_t)stream>>>(F);
        ZB_CUDA(cudaGetLastError());
        return 0;
    }
    // genuinely time-varying (12,4) fp32 problems: streamed thread-per-problem kernel (lqr_t1.cuh)
    if (dtype == ZB_F32 && n == 12 && m == 4 && N >= 2 && !lqr_fast_eligible(dtype, P) && arr_ok(P.A) && arr_ok(P.B) && arr_ok(P.Q) &&
        arr_ok(P.R) && aligned16(P.L) && (!P.V0 || aligned16(P.V0))) {
        FastP F{};
        F.Bsz = P.Bsz; F.N = P.N; F.T = P.T;
        F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R;
        F.gains = reinterpret_cast<float*>(P.L);
        F.V0 = reinterpret_cast<float*>(P.V0);
        return riccati_t1_tv_launch(F, (cudaStream_t)stream, (flags & ZB_TV_BULK_COPY) ? true : (ZB_TV_BULK != 0));
    }
    if (lqr_fast_eligible(dtype, P)) {
        if (P.Q.st == 0 || P.N == 1) {  // fully time-invariant: thread-per-problem kernel (lqr_t1.cuh)
            FastP F{};
            F.Bsz = P.Bsz; F.N = P.N; F.T = 1;
            F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R;
            F.Q.st = 0;
            if (P.N == 1 && P.Q.st != 0) { F.T = P.T; F.Q.st = P.Q.st; }
            F.gains = reinterpret_cast<float*>(P.L);
            F.V0 = reinterpret_cast<float*>(P.V0);
            if (F.T == 1) return riccati_t1_launch<false>(F, (cudaStream_t)stream);
        }
        return lqr_fast_launch(dtype, P, (cudaStream_t)stream);
    }
    // fp64 (12,4): cooperative four-threads-per-problem kernel (lqr_quad64.cuh)
    if (dtype == ZB_F64 && n == 12 && m == 4 && N >= 1 && arr_ok(P.A) && arr_ok(P.B) && arr_ok(P.Q) && arr_ok(P.R) && aligned16(P.L) &&
        (!P.V0 || aligned16(P.V0))) {
        LqrQuadP F{};
        F.Bsz = Bsz; F.N = N; F.T = T;
        F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R;
        F.L = L_out; F.V0 = V0_out;
        return riccati_quad_launch(dtype, F, (cudaStream_t)stream);
    }
    if (dtype == ZB_F64 && n == 8 && m == 4 && N >= 1 && !(flags & ZB_FORCE_GENERIC) && s84d::arr_ok(P.A) && s84d::arr_ok(P.B) &&
        s84d::arr_ok(P.Q) && s84d::arr_ok(P.R) && aligned16(P.L) && (!P.V0 || aligned16(P.V0)) && !getenv("ZB_FORCE_RUNTIME_SIZES") &&
        !getenv("ZB_NO_S84")) {  // the same in fp64 (lqr_s84d.cuh)
        s84d::S84DP F{};
        F.Bsz = Bsz; F.N = N; F.T = T;
        F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R;
        F.L = reinterpret_cast<double*>(L_out); F.V0 = reinterpret_cast<double*>(V0_out);
        return s84d::launch<false>(F, (cudaStream_t)stream);
    }
#define ZB_CT_LQR(N_, M_)                                                                                   \
    if (n == N_ && m == M_ && (dtype == ZB_F32 || ZB_CT_F64_OK(N_)) && !getenv("ZB_FORCE_RUNTIME_SIZES")) {   \
        if (dtype == ZB_F32) k_lqr_ct<float, N_, M_><<<gen_grid(Bsz), GEN_THREADS, 0, (cudaStream_t)stream>>>(P);   \
        else k_lqr_ct<double, (ZB_CT_F64_OK(N_) ? N_ : 2), (ZB_CT_F64_OK(N_) ? M_ : 1)><<<gen_grid(Bsz), GEN_THREADS, 0, (cudaStream_t)stream>>>(P); \
        ZB_CUDA(cudaGetLastError());                                                                        \
        return 0;                                                                                           \
    }
    ZB_CT_SHAPES(ZB_CT_LQR)
#undef ZB_CT_LQR
    ZB_DISPATCH(dtype, k_lqr_generic, gen_grid(Bsz), GEN_THREADS, stream, P);
    return 0;
}

int32_t zb_lqr_bilinear(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t T, int32_t n,
                        int32_t m, const zb_arr* A, const zb_arr* B, const zb_arr* d, const zb_arr* Q,
                        const zb_arr* R, const zb_arr* H, const zb_arr* q, const zb_arr* r, const zb_arr* q0,
                        void* L_out, void* l_out) {
    return zb_lqr_bilinear_flags(dtype, device, stream, Bsz, N, T, n, m, A, B, d, Q, R, H, q, r, q0, 0, L_out, l_out);
}

int32_t zb_lqr_bilinear_flags(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t T, int32_t n,
                              int32_t m, const zb_arr* A, const zb_arr* B, const zb_arr* d, const zb_arr* Q,
                              const zb_arr* R, const zb_arr* H, const zb_arr* q, const zb_arr* r, const zb_arr* q0,
                              int32_t flags, void* L_out, void* l_out) {
    int32_t rc = check_dims(dtype, Bsz, n, m);
    if (rc) return rc;
    if (Bsz == 0) return 0;  // empty batch: nothing to read or write (pointers may be NULL)
    ZB_ARG(N >= 0 && T >= 1 && T >= N, "need T >= N >= 0 and T >= 1 (got N=%d, T=%d)", N, T);
    ZB_ARG(A && B && d && Q && R && H && q && r && q0, "NULL operand");
    ZB_ARG(A->ptr && B->ptr && d->ptr && Q->ptr && R->ptr && H->ptr && q->ptr && r->ptr && q0->ptr, "NULL operand pointer");
    ZB_ARG((L_out && l_out) || N == 0, "NULL output");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    BilinP P{Bsz, N, T, n, m, to_arr(A), to_arr(B), to_arr(d), to_arr(Q), to_arr(R), to_arr(H), to_arr(q), to_arr(r),
             to_arr(q0), L_out, l_out};
    if (dtype == ZB_F32 && n == 8 && m == 4 && N >= 1 && !(flags & ZB_FORCE_GENERIC) && arr_ok(P.A) && arr_ok(P.B) && arr_ok(P.d) &&
        arr_ok(P.Q) && arr_ok(P.R) && arr_ok(P.H) && arr_ok(P.q) && arr_ok(P.r) && aligned16(L_out) && aligned16(l_out) &&
        !getenv("ZB_FORCE_RUNTIME_SIZES") && !getenv("ZB_NO_S84")) {  // lqr_s84.cuh (symmetric Q, R)
        s84::S84P F{};
        F.Bsz = Bsz; F.N = N; F.T = T;
        F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R; F.H = P.H; F.d = P.d; F.q = P.q; F.r = P.r;
        F.L = reinterpret_cast<float*>(L_out); F.l = reinterpret_cast<float*>(l_out);
        s84::k_riccati_s84<true><<<(unsigned)((Bsz + 31) / 32), 32, 0, (cudaStream_t)stream>>>(F);
        ZB_CUDA(cudaGetLastError());
        return 0;
    }
    if (dtype == ZB_F64 && n == 8 && m == 4 && N >= 1 && !(flags & ZB_FORCE_GENERIC) && s84d::arr_ok(P.A) && s84d::arr_ok(P.B) &&
        s84d::arr_ok(P.d) && s84d::arr_ok(P.Q) && s84d::arr_ok(P.R) && s84d::arr_ok(P.H) && s84d::arr_ok(P.q) && s84d::arr_ok(P.r) &&
        aligned16(L_out) && aligned16(l_out) && !getenv("ZB_FORCE_RUNTIME_SIZES") && !getenv("ZB_NO_S84")) {  // lqr_s84d.cuh
        s84d::S84DP F{};
        F.Bsz = Bsz; F.N = N; F.T = T;
        F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R; F.H = P.H; F.d = P.d; F.q = P.q; F.r = P.r;
        F.L = reinterpret_cast<double*>(L_out); F.l = reinterpret_cast<double*>(l_out);
        return s84d::launch<true>(F, (cudaStream_t)stream);
    }
#define ZB_CT_BIL(N_, M_)                                                                                        \
    if (n == N_ && m == M_ && (dtype == ZB_F32 || ZB_CT_F64_OK(N_)) && !getenv("ZB_FORCE_RUNTIME_SIZES")) {        \
        if (dtype == ZB_F32) k_bilinear_ct<float, N_, M_><<<gen_grid(Bsz), GEN_THREADS, 0, (cudaStream_t)stream>>>(P);   \
        else k_bilinear_ct<double, (ZB_CT_F64_OK(N_) ? N_ : 2), (ZB_CT_F64_OK(N_) ? M_ : 1)><<<gen_grid(Bsz), GEN_THREADS, 0, (cudaStream_t)stream>>>(P); \
        ZB_CUDA(cudaGetLastError());                                                                             \
        return 0;                                                                                                \
    }
    ZB_CT_SHAPES(ZB_CT_BIL)
#undef ZB_CT_BIL
    ZB_DISPATCH(dtype, k_bilinear_generic, gen_grid(Bsz), GEN_THREADS, stream, P);
    return 0;
}

int32_t zb_lqr_care_rk4(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t substeps, int32_t n, int32_t m,
                        double T_horizon, const zb_arr* A, const zb_arr* B, const zb_arr* Q, const zb_arr* Rinv, const zb_arr* Qf,
                        void* V_out) {
    int32_t rc = check_dims(dtype, Bsz, n, m);
    if (rc) return rc;
    ZB_ARG(N >= 2 && substeps >= 1, "need N >= 2 output points and substeps >= 1 (got %d, %d)", N, substeps);
    ZB_ARG(T_horizon > 0, "the horizon T must be positive");
    if (Bsz == 0) return 0;
    ZB_ARG(A && B && Q && Rinv && Qf && A->ptr && B->ptr && Q->ptr && Rinv->ptr && Qf->ptr && V_out, "NULL operand");
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    CareP P{Bsz, N, substeps, n, m, T_horizon / ((double)(N - 1) * substeps), to_arr(A), to_arr(B), to_arr(Q), to_arr(Rinv), to_arr(Qf), V_out};
    ZB_DISPATCH(dtype, k_care_rk4, gen_grid(Bsz), GEN_THREADS, stream, P);
    return 0;
}

static int32_t quad_common(int32_t dtype, int64_t Bsz, const void* x, const void* u, const double* wind, QuadP& P) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "bad dtype %d", dtype);
    ZB_ARG(Bsz >= 0, "negative batch size");
    ZB_ARG(x && u, "x/u NULL");
    P.Bsz = Bsz;
    P.x = x;
    P.u = u;
    P.lam = nullptr;
    P.o1 = P.o2 = nullptr;
    P.dt = 0;
    P.has_wind = 0;
    for (int i = 0; i < 3; ++i) {
        P.wind[i] = wind ? wind[i] : 0.0;
        if (P.wind[i] != 0.0) P.has_wind = 1;
    }
    return 0;
}

int32_t zb_quad_dynamics(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x, const void* u,
                         const double* wind_ned, void* xdot_out) {
    QuadP P;
    int32_t rc = quad_common(dtype, Bsz, x, u, wind_ned, P);
    if (rc) return rc;
    ZB_ARG(xdot_out, "output NULL");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    P.o1 = xdot_out;
    ZB_DISPATCH(dtype, k_quad_dynamics, (unsigned)((Bsz + 127) / 128), 128, stream, P);
    return 0;
}

int32_t zb_quad_linearize(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x, const void* u,
                          const double* wind_ned, double dt, void* A_out, void* B_out) {
    QuadP P;
    int32_t rc = quad_common(dtype, Bsz, x, u, wind_ned, P);
    if (rc) return rc;
    ZB_ARG(A_out, "A_out NULL");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    P.o1 = A_out;
    P.o2 = B_out;
    P.dt = dt;
    ZB_DISPATCH(dtype, k_quad_linearize, (unsigned)((Bsz + 63) / 64), 64, stream, P);
    return 0;
}

int32_t zb_quad_hess_contract(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x,
                              const void* u, const double* wind_ned, double dt, const void* lam, void* H_out) {
    QuadP P;
    int32_t rc = quad_common(dtype, Bsz, x, u, wind_ned, P);
    if (rc) return rc;
    ZB_ARG(lam && H_out, "lam/H_out NULL");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    P.o1 = H_out;
    P.lam = lam;
    P.dt = dt;
    ZB_DISPATCH(dtype, k_quad_hess, (unsigned)((Bsz + 63) / 64), 64, stream, P);
    return 0;
}

static int32_t roll_common(int32_t dtype, int64_t Bsz, int32_t N, const zb_model* model, const zb_cost* cost,
                           bool need_cost, RollP& P) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "bad dtype %d", dtype);
    ZB_ARG(Bsz >= 0 && N >= 0, "negative size");
    int32_t rc = to_model(model, P.M);
    if (rc) return rc;
    P.Bsz = Bsz;
    P.N = N;
    P.C = to_cost(cost);
    P.has_cost = (cost != nullptr);
    if (need_cost) ZB_ARG(cost && cost->Q.ptr && cost->R.ptr && cost->Qf.ptr, "cost (Q,R,Qf) required");
    return 0;
}

int32_t zb_ilqr_rollout(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, const zb_model* model,
                        const zb_cost* cost, const void* x0, const void* l, const void* L, const void* xPrev,
                        const void* uPrev, double alpha, void* xTraj, void* uTraj, void* J_out) {
    RollP P;
    int32_t rc = roll_common(dtype, Bsz, N, model, cost, J_out != nullptr, P);
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

int32_t zb_ilqr_forward_pass(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N,
                             const zb_model* model, const zb_cost* cost, const void* x0, const void* l,
                             const void* L, const void* xPrev, const void* uPrev, void* xTraj, void* uTraj,
                             void* J_out, int32_t* alpha_idx_out, void* Jall_out) {
    RollP P;
    int32_t rc = roll_common(dtype, Bsz, N, model, cost, true, P);
    if (rc) return rc;
    ZB_ARG(x0 && xTraj && J_out && Jall_out && (N == 0 || (l && L && xPrev && uPrev && uTraj)), "NULL operand");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    P.x0 = x0; P.l = l; P.L = L; P.xPrev = xPrev; P.uPrev = uPrev;
    P.xTraj = xTraj; P.uTraj = uTraj; P.J = nullptr;
    ZB_DISPATCH(dtype, k_forward_costs, gen_grid(Bsz * 16), GEN_THREADS, stream, P, Jall_out, (const uint8_t*)nullptr, (void*)nullptr);
    CommitP S{};
    S.J_out = J_out;
    S.idx_out = alpha_idx_out;
    ZB_DISPATCH(dtype, k_forward_commit, gen_grid(Bsz * 16), GEN_THREADS, stream, P, (const void*)Jall_out, S, (const void*)nullptr);
    return 0;
}

int32_t zb_ilqr_backward(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t n, int32_t m,
                         int32_t second_order, const zb_arr* f_x, const zb_arr* f_u, const zb_arr* f_xx,
                         const zb_arr* f_ux, const zb_arr* f_uu, const zb_arr* c, const zb_arr* c_x,
                         const zb_arr* c_u, const zb_arr* c_xx, const zb_arr* c_ux, const zb_arr* c_uu,
                         const zb_arr* v, const zb_arr* v_x, const zb_arr* v_xx, void* l_out, void* L_out,
                         void* v_out, void* vx_out, void* vxx_out) {
    int32_t rc = check_dims(dtype, Bsz, n, m);
    if (rc) return rc;
    if (Bsz == 0) return 0;  // empty batch: nothing to read or write (pointers may be NULL)
    ZB_ARG(N >= 0, "negative N");
    ZB_ARG(f_x && f_u && c && c_x && c_u && c_xx && c_ux && c_uu && v && v_x && v_xx, "NULL operand");
    ZB_ARG(!second_order || (f_xx && f_ux && f_uu && f_xx->ptr && f_ux->ptr && f_uu->ptr), "DDP needs f_xx, f_ux, f_uu");
    ZB_ARG((l_out && L_out) || N == 0, "NULL output");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    BackP P{Bsz, N, n, m, second_order, to_arr(f_x), to_arr(f_u), to_arr(f_xx), to_arr(f_ux), to_arr(f_uu), to_arr(c),
            to_arr(c_x), to_arr(c_u), to_arr(c_xx), to_arr(c_ux), to_arr(c_uu), to_arr(v), to_arr(v_x), to_arr(v_xx),
            l_out, L_out, v_out, vx_out, vxx_out, 1e-3};
    ZB_DISPATCH(dtype, k_backward, gen_grid(Bsz), GEN_THREADS, stream, P);
    return 0;
}

int32_t zb_pd_clamp(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t p, double eps, const void* in,
                    void* out) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "bad dtype %d", dtype);
    ZB_ARG(p >= 1 && p <= ZB_PD_MAX, "p must be in [1,%d] (got %d)", ZB_PD_MAX, p);
    ZB_ARG(Bsz >= 0 && in && out, "bad operand");
    if (Bsz == 0) return 0;
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    ZB_DISPATCH(dtype, k_pd_clamp, gen_grid(Bsz), GEN_THREADS, stream, (long long)Bsz, p, eps, in, out);
    return 0;
}

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

size_t zb_ilqr_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N, int32_t n, int32_t m) {
    size_t e = dtype == ZB_F64 ? 8 : 4;
    size_t p = (size_t)(n + m);
    const size_t per = (size_t)(N + 1) * n + (size_t)N * m;  // one trajectory
    return align256(e * Bsz * N * m) + align256(e * Bsz * 16) + align256(e * Bsz * p * p) + align256(e * Bsz * n * n) +
           align256(e * SPEC_N * Bsz * per) + 2 * align256(sizeof(int32_t) * (size_t)Bsz) + 512 +  // + two active lists, two counters
           align256(e * Bsz * 84);                                                                // + DDP eigenvector carry
}

int32_t zb_ilqr_solve(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t flags,
                      const zb_model* model, const zb_cost* cost, const void* x0, const void* uGuess,
                      int32_t maxIter, double tol, void* xTraj, void* uTraj, void* L_out, void* J_out,
                      uint8_t* converged_out, int32_t* iters_out, int32_t* alpha_log, void* J_log,
                      void* workspace, size_t workspace_bytes) {
    const int32_t second_order = flags & ZB_SECOND_ORDER;
    const bool cost_diagonal = (flags & ZB_COST_DIAGONAL) != 0;
    RollP P;
    int32_t rc = roll_common(dtype, Bsz, N, model, cost, true, P);
    if (rc) return rc;
    ZB_ARG(maxIter >= 0, "negative maxIter");
    ZB_ARG(x0 && uGuess && xTraj && uTraj && L_out && J_out && converged_out && iters_out, "NULL operand");
    const int n = P.M.n, m = P.M.m, p = n + m;
    size_t need = zb_ilqr_workspace_bytes(dtype, Bsz, N, n, m);
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
    void* spec = w;  w += align256(e * SPEC_N * Bsz * ((size_t)(N + 1) * n + (size_t)N * m));  // speculative trajectories of the two largest step sizes
    int32_t* perm2[2];
    perm2[0] = reinterpret_cast<int32_t*>(w);  w += align256(sizeof(int32_t) * (size_t)Bsz);
    perm2[1] = reinterpret_cast<int32_t*>(w);  w += align256(sizeof(int32_t) * (size_t)Bsz);
    int32_t* count2 = reinterpret_cast<int32_t*>(w);  w += 512;  // two counters, 128 bytes apart
    void* ev_ws = w;  // DDP: eigenvectors of the clamped block, carried from step to step (ilqr_fast.cuh)
    cudaStream_t s = (cudaStream_t)stream;
    // With a real tolerance problems freeze at different iterations (ilqrUtils.py:301-303,318): the still-iterating ones are
    // re-listed after every forward pass so that frozen problems occupy no lane, and the host stops enqueuing once the list is
    // empty (checked every ZB_EXIT_CHECK iterations: one 4-byte read + stream sync).  tol < 0 (forced iteration count, as the
    // benchmarks use) never converges: no lists, no check, the call stays fully asynchronous.
    const bool track = tol >= 0.0 && maxIter > 0;
    // policy.L = 0 before the first iteration (ilqrUtils.py:293)
    ZB_CUDA(cudaMemsetAsync(L_out, 0, e * Bsz * N * m * n, s));
    P.x0 = x0; P.l = l_ws; P.L = L_out; P.xPrev = xTraj; P.uPrev = uTraj;
    P.xTraj = xTraj; P.uTraj = uTraj; P.J = nullptr;
    const bool fast_fwd = N >= 1 && !(flags & ZB_GENERIC_FORWARD) && fwd_quad_eligible(P.M) && aligned16(xTraj) && aligned16(uTraj) && aligned16(L_out) && aligned16(x0);
    if (fast_fwd && cost_diagonal && aligned16(uGuess)) {  // diagonal costs + quadcopter: closed-form conditioning and a register-resident initial rollout
        SetupQuadP Sp{Bsz, N, (int)maxIter, P.M.dt, 1e-3, P.C, x0, uGuess, xTraj, uTraj, J_out, converged_out, iters_out, alpha_log, J_log, Czz, Vfxx};
        rc = solve_setup_quad_launch(dtype, Sp, s);
        if (rc) return rc;
    } else {
        ZB_DISPATCH(dtype, k_solve_prep, gen_grid(Bsz), GEN_THREADS, stream, (long long)Bsz, n, m, P.C, 1e-3, Czz, Vfxx);
        ZB_DISPATCH(dtype, k_solve_init, gen_grid(Bsz), GEN_THREADS, stream, P, uGuess, J_out, converged_out, iters_out,
                    alpha_log, J_log, (int)maxIter);
    }
    SolveBackP Bk{Bsz, N, second_order, P.M, P.C, xTraj, uTraj, Czz, Vfxx, converged_out, l_ws, L_out, 1e-3};
    const bool fast_bwd = N >= 1 && ilqr_fast_eligible(P.M, second_order, cost_diagonal) && aligned16(xTraj) && aligned16(uTraj) && aligned16(L_out);
    IlqrFastP Fb{Bsz, N, P.M.dt, P.C, xTraj, uTraj, Czz, Vfxx, converged_out, l_ws, L_out, 1e-3, ActiveP{nullptr, nullptr}, Bsz, 0, ev_ws};
    if (track) {
        rc = compact_active_launch(Bsz, converged_out, perm2[0], count2, s);
        if (rc) return rc;
    }
    for (int it = 0; it < maxIter; ++it) {
        const ActiveP act = track ? ActiveP{perm2[it & 1], count2 + 32 * (it & 1)} : ActiveP{nullptr, nullptr};
        Fb.act = act;
        if (fast_bwd) {
            rc = ilqr_fast_launch(dtype, Fb, s, cost_diagonal, second_order != 0);
            if (rc) return rc;
        } else
            ZB_DISPATCH(dtype, k_solve_backward, gen_grid(Bsz), GEN_THREADS, stream, Bk);
        CommitP S{J_out, converged_out, iters_out, alpha_log, J_log, it, (int)maxIter, tol, nullptr, nullptr};
        if (fast_fwd) {  // line search + commit + bookkeeping in one launch (ilqr_forward.cuh)
            FwdQuadP Fw{Bsz, N, P.M.dt, P.C, x0, l_ws, L_out, xTraj, uTraj, spec, nullptr, S, cost_diagonal ? 1 : 0, act};
            rc = fwd_quad_launch(dtype, Fw, s);
            if (rc) return rc;
        } else {
            ZB_DISPATCH(dtype, k_forward_costs, gen_grid(Bsz * 16), GEN_THREADS, stream, P, Jall, (const uint8_t*)converged_out, spec);
            ZB_DISPATCH(dtype, k_forward_commit, gen_grid(Bsz * 16), GEN_THREADS, stream, P, (const void*)Jall, S, (const void*)spec);
        }
        if (track && it + 1 < maxIter) {
            int32_t* cnt_next = count2 + 32 * ((it + 1) & 1);
            rc = compact_active_launch(Bsz, converged_out, perm2[(it + 1) & 1], cnt_next, s);
            if (rc) return rc;
            if ((it % ZB_EXIT_CHECK) == ZB_EXIT_CHECK - 1) {  // anything left to iterate on?
                int32_t left = 1;
                ZB_CUDA(cudaMemcpyAsync(&left, cnt_next, sizeof(int32_t), cudaMemcpyDeviceToHost, s));
                ZB_CUDA(cudaStreamSynchronize(s));
                if (left == 0) break;
                Fb.active_hint = left;  // the list only shrinks from here on: the backward pass sizes its grid and CTAs by it
            }
        }
    }
    return 0;
}

size_t zb_mpc_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N, int32_t n, int32_t m) {
    size_t e = dtype == ZB_F64 ? 8 : 4;
    const size_t Bpad = ((size_t)Bsz + 31) / 32 * 32;  // the (12,4) kernel stores gains per 32-problem group
    return align256(e * Bpad * (size_t)admm_ws_elems(N, n, m)) + align256(sizeof(int) * (Bpad / 32)) + 512;  // >= the gains buffer of the unconstrained path + its group flags
}

int32_t zb_mpc_lqr_solve(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t n, int32_t m,
                         const zb_arr* A, const zb_arr* B, const zb_arr* Q, const zb_arr* R, const zb_arr* Qf,
                         const zb_arr* x_lb, const zb_arr* x_ub, const zb_arr* u_lb, const zb_arr* u_ub,
                         int32_t flags, const void* x0, const zb_admm_opts* opts, void* u0_out, void* xTraj,
                         void* uTraj, int8_t* status_out, int32_t* iters_out, void* workspace,
                         size_t workspace_bytes) {
    int32_t rc = check_dims(dtype, Bsz, n, m);
    if (rc) return rc;
    if (Bsz == 0) return 0;  // empty batch: nothing to read or write (pointers may be NULL)
    ZB_ARG(N >= 1, "N must be >= 1");
    ZB_ARG(A && B && Q && R && Qf && A->ptr && B->ptr && Q->ptr && R->ptr && Qf->ptr, "NULL operand");
    ZB_ARG(x0 && u0_out && xTraj && uTraj && status_out, "NULL operand");
    size_t need = zb_mpc_workspace_bytes(dtype, Bsz, N, n, m);
    ZB_ARG(workspace && workspace_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    if (flags & ZB_MPC_BOUNDED) {
        ZB_ARG(x_lb && x_ub && u_lb && u_ub && x_lb->ptr && x_ub->ptr && u_lb->ptr && u_ub->ptr, "bounds are NULL");
        AdmmP P{};
        P.Bsz = Bsz; P.N = N; P.n = n; P.m = m;
        P.A = to_arr(A); P.B = to_arr(B); P.Q = to_arr(Q); P.R = to_arr(R); P.Qf = to_arr(Qf);
        P.xlb = to_arr(x_lb); P.xub = to_arr(x_ub); P.ulb = to_arr(u_lb); P.uub = to_arr(u_ub);
        P.x0 = x0; P.u0 = u0_out; P.xTraj = xTraj; P.uTraj = uTraj; P.status = status_out; P.iters = iters_out;
        P.ws = workspace; P.ws_stride = admm_ws_elems(N, n, m);
        P.max_iter = opts && opts->max_iter > 0 ? opts->max_iter : 4000;
        P.check_every = opts && opts->check_every > 0 ? opts->check_every : 25;
        P.rho = opts && opts->rho > 0 ? opts->rho : 0.1;
        P.alpha = opts && opts->alpha > 0 ? opts->alpha : 1.6;
        P.eps_abs = opts && opts->eps_abs > 0 ? opts->eps_abs : 1e-3;
        P.eps_rel = opts && opts->eps_rel > 0 ? opts->eps_rel : 1e-3;
        P.eps_inf = opts && opts->eps_prim_inf > 0 ? opts->eps_prim_inf : 1e-4;
        ZB_DISPATCH(dtype, k_mpc_admm, gen_grid(Bsz), GEN_THREADS, stream, P);
        return 0;
    }
    MpcP P{Bsz, N, n, m, to_arr(A), to_arr(B), to_arr(Q), to_arr(R), to_arr(Qf), x0, u0_out, xTraj, uTraj, workspace,
           status_out, iters_out};
    const bool fast_ok = !(flags & ZB_FORCE_GENERIC);
    if (fast_ok && dtype == ZB_F32 && n == 12 && m == 4 && arr_ok(P.A) && arr_ok(P.B) && arr_ok(P.Q) && arr_ok(P.R) && arr_ok(P.Qf) &&
        aligned16(x0) && aligned16(u0_out) && aligned16(xTraj) && aligned16(uTraj) && aligned16(workspace)) {
        FastP F{};
        F.Bsz = Bsz; F.N = N; F.T = 1;
        F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R; F.Qf = P.Qf;
        F.q_time_varying = 0;
        F.gains = reinterpret_cast<float*>(workspace);
        F.x0 = reinterpret_cast<const float*>(x0);
        F.u0 = reinterpret_cast<float*>(u0_out);
        F.xTraj = reinterpret_cast<float*>(xTraj);
        F.uTraj = reinterpret_cast<float*>(uTraj);
        F.status = status_out;
        F.iters = iters_out;
        const bool diag = (flags & ZB_COST_DIAGONAL) != 0;
        int sm_count = 0;
        ZB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
        // opt-in: sweep and plan rollout as two concurrent kernels (measured equal to the fused kernel, see lqr_t1.cuh)
        if ((flags & ZB_MPC_SPLIT_ROLLOUT) && diag && Bsz > (int64_t)sm_count * 4 * 32)
        {   // the group flags live behind the gains in the workspace
            const size_t gains_bytes = (((size_t)Bsz + 31) / 32) * 32 * (size_t)N * 48 * sizeof(float);
            int* flags = reinterpret_cast<int*>(reinterpret_cast<char*>(workspace) + ((gains_bytes + 255) & ~(size_t)255));
            return riccati_t1_mpc_split_launch(F, (cudaStream_t)stream, diag, device, sm_count, flags);
        }
        return riccati_t1_launch<true>(F, (cudaStream_t)stream, diag);
    }
    if (fast_ok && dtype == ZB_F64 && n == 12 && m == 4 && arr_ok(P.A) && arr_ok(P.B) && arr_ok(P.Q) && arr_ok(P.R) && arr_ok(P.Qf) &&
        aligned16(x0) && aligned16(u0_out) && aligned16(xTraj) && aligned16(uTraj) && aligned16(workspace)) {
        LqrQuadP F{};  // fp64: cooperative Riccati sweep (gains to the workspace) + quad plan rollout (lqr_quad64.cuh)
        F.Bsz = Bsz; F.N = N; F.T = 1;
        F.A = P.A; F.B = P.B; F.Q = P.Q; F.R = P.R; F.Qf = P.Qf;
        F.A.st = F.B.st = F.Q.st = F.R.st = 0;
        F.L = workspace; F.x0 = x0; F.u0 = u0_out; F.xTraj = xTraj; F.uTraj = uTraj;
        F.status = status_out; F.iters = iters_out;
        F.cost_diagonal = (flags & ZB_COST_DIAGONAL) ? 1 : 0;
        return riccati_quad_launch(dtype, F, (cudaStream_t)stream);
    }
    ZB_DISPATCH(dtype, k_mpc_riccati, gen_grid(Bsz), GEN_THREADS, stream, P);
    return 0;
}

}  // extern "C"

// ---- lqrMpc with finite bounds, one problem definition shared by the batch (mpc_box.cuh) ----------------------------
template <typename T>
static void box_fill_ops(box::Ops<T>& O, const double* A, const double* B, const double* xlb, const double* xub,
                         const double* ulb, const double* uub) {
    for (int i = 0; i < 144; ++i) O.A[i] = (T)A[i];
    for (int i = 0; i < 48; ++i) O.B[i] = (T)B[i];
    for (int i = 0; i < 12; ++i) { O.xlb[i] = xlb ? (T)xlb[i] : T(0); O.xub[i] = xub ? (T)xub[i] : T(0); }
    for (int i = 0; i < 4; ++i) { O.ulb[i] = ulb ? (T)ulb[i] : T(0); O.uub[i] = uub ? (T)uub[i] : T(0); }
}

template <typename T>
static int32_t box_build_tables(cudaStream_t s, int N, const double* A, const double* B, const double* Q, const double* R,
                                const double* Qf, double rho0, void* tables) {
    box::Ops<T> O;
    box::Costs<T> C;
    box_fill_ops<T>(O, A, B, nullptr, nullptr, nullptr, nullptr);
    for (int i = 0; i < 144; ++i) { C.Q[i] = (T)Q[i]; C.Qf[i] = (T)Qf[i]; }
    for (int i = 0; i < 16; ++i) C.R[i] = (T)R[i];
    box::k_box_tables<T><<<box::LEVELS, 32, 0, s>>>(O, C, N, (T)rho0, reinterpret_cast<T*>(tables));
    ZB_CUDA(cudaGetLastError());
    return 0;
}

template <typename T>
static int32_t box_solve(cudaStream_t s, int64_t Bsz, int N, const double* A, const double* B, const double* xlb,
                         const double* xub, const double* ulb, const double* uub, const void* tables, const void* x0,
                         const zb_admm_opts* opts, void* u0, void* xTraj, void* uTraj, int8_t* status, int32_t* iters,
                         void* workspace, int32_t flags, int device, int Tsim = -1, void* xSim = nullptr, void* uSim = nullptr,
                         double clip_margin = 0) {
    box::Ops<T> O;
    box_fill_ops<T>(O, A, B, xlb, xub, ulb, uub);
    box::Params<T> P{};
    P.Bsz = Bsz; P.N = N;
    P.x0 = reinterpret_cast<const T*>(x0);
    P.u0 = reinterpret_cast<T*>(u0); P.xTraj = reinterpret_cast<T*>(xTraj); P.uTraj = reinterpret_cast<T*>(uTraj);
    P.status = status; P.iters = iters;
    P.ws = reinterpret_cast<T*>(workspace); P.tab = reinterpret_cast<const T*>(tables);
    P.max_iter = opts && opts->max_iter > 0 ? opts->max_iter : 4000;
    P.check_every = opts && opts->check_every > 0 ? opts->check_every : 25;
    P.rho0 = (T)(opts && opts->rho > 0 ? opts->rho : 0.1);
    P.alpha = (T)(opts && opts->alpha > 0 ? opts->alpha : 1.6);
    P.eps_abs = (T)(opts && opts->eps_abs > 0 ? opts->eps_abs : 1e-3);
    P.eps_rel = (T)(opts && opts->eps_rel > 0 ? opts->eps_rel : 1e-3);
    P.eps_inf = (T)(opts && opts->eps_prim_inf > 0 ? opts->eps_prim_inf : 1e-4);
    if (Tsim >= 0) { P.Tsim = Tsim; P.xSim = reinterpret_cast<T*>(xSim); P.uSim = reinterpret_cast<T*>(uSim); P.clip_margin = (T)clip_margin; }
    if constexpr (sizeof(T) == 4) {
        int sm_count = 0;
        ZB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
        // One-warp CTAs (8 problems) with the ADMM state in shared memory, as many per SM as fit; larger batches run as
        // successive waves of such CTAs.  Measured on B200 (N = 25, eps 1e-3): 4.6-5.3 ms per wave of <= 5,920 problems
        // against 15-20 ms for the same problems with the state in L2, and still ahead of the thread-per-problem kernel
        // (all problems resident, state streamed from L2/HBM) at 65,536 problems (49 vs 53-70 ms).  So: whenever the
        // horizon's state fits in shared memory.
        const size_t smem = (size_t)(N + 1) * box::QW * 32 * sizeof(float);
        const bool fits = smem + 1024 <= 227u * 1024;
        (void)sm_count;
        const bool quad = (flags & ZB_VARIANT_QUAD) || (!(flags & ZB_VARIANT_THREAD) && fits);
        if (quad) {
            const bool on_chip = !(flags & ZB_BOX_STATE_GLOBAL) && fits;
            if (on_chip) {
                const unsigned grid = (unsigned)((Bsz + 7) / 8);
                if (Tsim >= 0) {
                    ZB_CUDA(cudaFuncSetAttribute(box::k_mpc_box_quad<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    box::k_mpc_box_quad<true, true><<<grid, 32, smem, s>>>(O, P);
                } else {
                    ZB_CUDA(cudaFuncSetAttribute(box::k_mpc_box_quad<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
                    box::k_mpc_box_quad<false, true><<<grid, 32, smem, s>>>(O, P);
                }
            } else {
                const unsigned grid = (unsigned)((Bsz + 15) / 16);
                if (Tsim >= 0) box::k_mpc_box_quad<true, false><<<grid, 64, 0, s>>>(O, P);
                else box::k_mpc_box_quad<false, false><<<grid, 64, 0, s>>>(O, P);
            }
            ZB_CUDA(cudaGetLastError());
            return 0;
        }
    }
    if (Tsim >= 0) {
        box::k_mpc_box_closed_loop<T><<<(unsigned)((Bsz + 31) / 32), 32, 0, s>>>(O, P);
    } else {
        box::k_mpc_box<T><<<(unsigned)((Bsz + 31) / 32), 32, 0, s>>>(O, P);
    }
    ZB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" {

size_t zb_mpc_box_tables_bytes(int32_t dtype, int32_t N) {
    return align256((dtype == ZB_F64 ? 8 : 4) * (size_t)box::tab_elems(N));
}

size_t zb_mpc_box_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N) {
    return align256((dtype == ZB_F64 ? 8 : 4) * (size_t)box::ws_elems(N, Bsz)) + 256;
}

size_t zb_mpc_box_closed_loop_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N) {  // ADMM state + one plan per problem
    const size_t e = dtype == ZB_F64 ? 8 : 4;
    return align256(e * (size_t)box::ws_elems(N, Bsz)) + align256(e * (size_t)Bsz * (N + 1) * 12) + align256(e * (size_t)Bsz * N * 4) + 256;
}

int32_t zb_mpc_box_closed_loop(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t Tsim, const double* A,
                               const double* B, const double* x_lb, const double* x_ub, const double* u_lb, const double* u_ub,
                               const void* tables, const void* x0, const zb_admm_opts* opts, int32_t flags, double clip_margin,
                               void* xSim_out, void* uSim_out, int8_t* status_out, int32_t* iters_out, void* workspace,
                               size_t workspace_bytes) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "dtype must be ZB_F32 or ZB_F64 (got %d)", dtype);
    ZB_ARG(Bsz >= 0 && N >= 1 && Tsim >= 0, "bad sizes");
    if (Bsz == 0) return 0;
    ZB_ARG(A && B && x_lb && x_ub && u_lb && u_ub && tables, "NULL operand");
    ZB_ARG(x0 && xSim_out && (uSim_out || Tsim == 0) && status_out, "NULL operand");
    size_t need = zb_mpc_box_closed_loop_workspace_bytes(dtype, Bsz, N);
    ZB_ARG(workspace && workspace_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    const size_t e = dtype == ZB_F64 ? 8 : 4;
    char* w = reinterpret_cast<char*>(workspace);
    void* state = w; w += align256(e * (size_t)box::ws_elems(N, Bsz));
    void* planx = w; w += align256(e * (size_t)Bsz * (N + 1) * 12);
    void* planu = w;
    return dtype == ZB_F32 ? box_solve<float>((cudaStream_t)stream, Bsz, N, A, B, x_lb, x_ub, u_lb, u_ub, tables, x0, opts, nullptr, planx, planu,
                                              status_out, iters_out, state, flags, device, Tsim, xSim_out, uSim_out, clip_margin)
                           : box_solve<double>((cudaStream_t)stream, Bsz, N, A, B, x_lb, x_ub, u_lb, u_ub, tables, x0, opts, nullptr, planx, planu,
                                               status_out, iters_out, state, flags, device, Tsim, xSim_out, uSim_out, clip_margin);
}

int32_t zb_mpc_box_build_tables(int32_t dtype, int32_t device, void* stream, int32_t N, const double* A, const double* B,
                                const double* Q, const double* R, const double* Qf, double rho0, void* tables,
                                size_t tables_bytes) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "dtype must be ZB_F32 or ZB_F64 (got %d)", dtype);
    ZB_ARG(N >= 1, "N must be >= 1");
    ZB_ARG(A && B && Q && R && Qf && tables, "NULL operand");
    ZB_ARG(rho0 > 0, "rho0 must be positive");
    ZB_ARG(tables_bytes >= zb_mpc_box_tables_bytes(dtype, N), "tables buffer too small: need %zu bytes, got %zu",
           zb_mpc_box_tables_bytes(dtype, N), tables_bytes);
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    return dtype == ZB_F32 ? box_build_tables<float>((cudaStream_t)stream, N, A, B, Q, R, Qf, rho0, tables)
                           : box_build_tables<double>((cudaStream_t)stream, N, A, B, Q, R, Qf, rho0, tables);
}

int32_t zb_mpc_box_solve(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, const double* A, const double* B,
                         const double* x_lb, const double* x_ub, const double* u_lb, const double* u_ub, const void* tables,
                         const void* x0, const zb_admm_opts* opts, int32_t flags, void* u0_out, void* xTraj, void* uTraj,
                         int8_t* status_out, int32_t* iters_out, void* workspace, size_t workspace_bytes) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "dtype must be ZB_F32 or ZB_F64 (got %d)", dtype);
    ZB_ARG(Bsz >= 0, "negative batch size");
    if (Bsz == 0) return 0;
    ZB_ARG(N >= 1, "N must be >= 1");
    ZB_ARG(A && B && x_lb && x_ub && u_lb && u_ub && tables, "NULL operand");
    ZB_ARG(x0 && u0_out && xTraj && uTraj && status_out, "NULL operand");
    size_t need = zb_mpc_box_workspace_bytes(dtype, Bsz, N);
    ZB_ARG(workspace && workspace_bytes >= need, "workspace too small: need %zu bytes, got %zu", need, workspace_bytes);
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    return dtype == ZB_F32 ? box_solve<float>((cudaStream_t)stream, Bsz, N, A, B, x_lb, x_ub, u_lb, u_ub, tables, x0, opts, u0_out,
                                              xTraj, uTraj, status_out, iters_out, workspace, flags, device)
                           : box_solve<double>((cudaStream_t)stream, Bsz, N, A, B, x_lb, x_ub, u_lb, u_ub, tables, x0, opts, u0_out,
                                               xTraj, uTraj, status_out, iters_out, workspace, flags, device);
}

}  // extern "C"

extern "C" {

int32_t zb_mpc_closed_loop_quad(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t Tsim, double dt,
                                const double* u_trim, const zb_arr* Q, const zb_arr* R, const zb_arr* Qf, int32_t flags,
                                const void* x0, void* xSim_out, void* uSim_out) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "bad dtype %d", dtype);
    ZB_ARG(Bsz >= 0 && N >= 1 && Tsim >= 0, "bad sizes");
    if (Bsz == 0) return 0;
    ZB_ARG(Q && R && Qf && Q->ptr && R->ptr && Qf->ptr && x0 && xSim_out && (uSim_out || Tsim == 0) && u_trim, "NULL operand");
    ZB_ARG(aligned16(x0) && aligned16(xSim_out) && aligned16(uSim_out), "x0 / outputs must be 16-byte aligned");
    ZB_ARG((flags & ZB_COST_DIAGONAL) || (arr_ok(to_arr(Q)) && arr_ok(to_arr(R)) && arr_ok(to_arr(Qf))), "cost blocks must be 16-byte aligned");
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    if (dtype == ZB_F64) {  // the reference's own precision: cooperative four-threads-per-problem kernel (lqr_quad64.cuh)
        ZB_ARG(arr_ok(to_arr(Q)) && arr_ok(to_arr(R)) && arr_ok(to_arr(Qf)), "cost blocks must be 16-byte aligned");
        ClosedLoopQuadP P64{};
        P64.Bsz = Bsz; P64.N = N; P64.Tsim = Tsim; P64.dt = dt;
        for (int i = 0; i < 4; ++i) P64.utrim[i] = u_trim[i];
        P64.Q = to_arr(Q); P64.R = to_arr(R); P64.Qf = to_arr(Qf);
        P64.x0 = x0; P64.xSim = xSim_out; P64.uSim = uSim_out;
        P64.cost_diagonal = (flags & ZB_COST_DIAGONAL) ? 1 : 0;
        return mpc_closed_loop_quad64_launch(P64, (cudaStream_t)stream);
    }
    t1::ClosedLoopP P{};
    P.Bsz = Bsz; P.N = N; P.Tsim = Tsim; P.dt = (float)dt;
    for (int i = 0; i < 4; ++i) P.utrim[i] = (float)u_trim[i];
    P.Q = to_arr(Q); P.R = to_arr(R); P.Qf = to_arr(Qf);
    P.x0 = reinterpret_cast<const float*>(x0);
    P.xSim = reinterpret_cast<float*>(xSim_out);
    P.uSim = reinterpret_cast<float*>(uSim_out);
    // small batches: the cooperative (4 threads per problem) variant shortens the sequential chain and fills more SMs
    int sm_count = 0;  // cudaDeviceGetAttribute is cheap; cudaGetDeviceProperties costs milliseconds per call
    ZB_CUDA(cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, device));
    const bool dense_ok = arr_ok(to_arr(Q)) && arr_ok(to_arr(R)) && arr_ok(to_arr(Qf));
    // smallest batches (cfg 3 sharded over 8 GPUs: ~14 problems per SM): nine lanes per problem, operands in registers (mpc_warp.cuh)
    const bool forced = (flags & (ZB_VARIANT_THREAD | ZB_VARIANT_QUAD | ZB_VARIANT_WARP)) != 0;
    const bool want_warp = (flags & ZB_VARIANT_WARP) || (!forced && Bsz <= (int64_t)sm_count * ZB_W9_MAX_PER_SM);
    if (want_warp) return mpc_closed_loop_w9_launch(P, (cudaStream_t)stream);
    const bool want_quad = (flags & ZB_VARIANT_QUAD) != 0;  // kept selectable; no longer picked automatically
    if (dense_ok && want_quad) return mpc_closed_loop_coop_launch(P, (cudaStream_t)stream);
    return mpc_closed_loop_launch(P, (cudaStream_t)stream, (flags & ZB_COST_DIAGONAL) != 0);
}

int32_t zb_peak_fma(int32_t dtype, int32_t device, double* flops_per_s_out, double* sm_clock_mhz_out) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "bad dtype %d", dtype);
    ZB_ARG(flops_per_s_out, "NULL output");
    DeviceGuard g(device);
    ZB_CUDA(g.err);
    cudaDeviceProp prop;
    ZB_CUDA(cudaGetDeviceProperties(&prop, device));
    void* out;
    ZB_CUDA(cudaMalloc(&out, 64));
    cudaEvent_t e0, e1;
    ZB_CUDA(cudaEventCreate(&e0));
    ZB_CUDA(cudaEventCreate(&e1));
    const int ILP = 8, threads = 256, blocks = prop.multiProcessorCount * 8;
    const int iters = dtype == ZB_F32 ? 4096 : 2048;
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        ZB_CUDA(cudaEventRecord(e0));
        if (dtype == ZB_F32) k_peak_fma<float, ILP><<<blocks, threads>>>((float*)out, iters);
        else k_peak_fma<double, ILP><<<blocks, threads>>>((double*)out, iters);
        ZB_CUDA(cudaEventRecord(e1));
        ZB_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        ZB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double flops = 2.0 * ILP * 8 * (double)iters * threads * (double)blocks;
        double f = flops / (ms * 1e-3);
        if (rep > 0 && f > best) best = f;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *flops_per_s_out = best;
    if (sm_clock_mhz_out) *sm_clock_mhz_out = prop.clockRate / 1000.0;
    return 0;
}

}  // extern "C"
