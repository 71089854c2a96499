// TEST INFRASTRUCTURE ONLY.  Compiles the per-problem bodies of the generic CUDA kernels
// (zopt_b200/csrc/zb_problems.cuh) for the host so that `pytest -m "not gpu"` can check their
// arithmetic against the oracle in a container without a GPU.  Never linked into, imported by or
// reachable from the product package; the product path has no CPU fallback.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
using std::fabs;
using std::sqrt;
using std::sin;
using std::cos;
using std::fmax;
using std::log2;
using std::ldexp;
#include "../../include/zopt_b200.h"
#include "../../zopt_b200/csrc/zb_problems.cuh"
#include "../../zopt_b200/csrc/mpc_box.cuh"

using namespace zb;

static Arr A_(const zb_arr* a) {
    Arr r;
    r.p = a ? a->ptr : nullptr;
    r.sb = a ? a->stride_b : 0;
    r.st = a ? a->stride_t : 0;
    return r;
}
static Model M_(const zb_model* m) {
    Model M;
    M.kind = m->kind; M.n = m->n; M.m = m->m; M.has_wind = m->has_wind; M.dt = m->dt;
    for (int i = 0; i < 3; ++i) M.wind[i] = m->wind[i];
    M.A = A_(&m->A); M.B = A_(&m->B);
    return M;
}
static Cost C_(const zb_cost* c) {
    Cost C;
    C.Q = A_(c ? &c->Q : nullptr); C.R = A_(c ? &c->R : nullptr); C.Qf = A_(c ? &c->Qf : nullptr);
    return C;
}

#define EXPORT extern "C" __attribute__((visibility("default")))

EXPORT int hs_lqr_dfh(int dtype, int64_t Bsz, int N, int T, int n, int m, const zb_arr* A, const zb_arr* B,
                      const zb_arr* Q, const zb_arr* R, void* L, void* V0) {
    LqrP P{Bsz, N, T, n, m, A_(A), A_(B), A_(Q), A_(R), L, V0};
    for (int64_t b = 0; b < Bsz; ++b) dtype ? lqr_problem<double>(P, b) : lqr_problem<float>(P, b);
    return 0;
}

EXPORT int hs_lqr_bilinear(int dtype, int64_t Bsz, int N, int T, int n, int m, const zb_arr* A, const zb_arr* B,
                           const zb_arr* d, const zb_arr* Q, const zb_arr* R, const zb_arr* H, const zb_arr* q,
                           const zb_arr* r, const zb_arr* q0, void* L, void* l) {
    BilinP P{Bsz, N, T, n, m, A_(A), A_(B), A_(d), A_(Q), A_(R), A_(H), A_(q), A_(r), A_(q0), L, l};
    for (int64_t b = 0; b < Bsz; ++b) dtype ? bilinear_problem<double>(P, b) : bilinear_problem<float>(P, b);
    return 0;
}

template <typename T>
static void quad_all(int64_t Bsz, const T* x, const T* u, const double* wind, double dt, const T* lam, T* xd, T* A,
                     T* Bm, T* H) {
    T w[3] = {T(wind ? wind[0] : 0), T(wind ? wind[1] : 0), T(wind ? wind[2] : 0)};
    bool hw = w[0] != 0 || w[1] != 0 || w[2] != 0;
    for (int64_t b = 0; b < Bsz; ++b) {
        if (xd) quad_F<T>(x + b * 12, u + b * 4, w, hw, xd + b * 12);
        if (A) quad_lin<T>(x + b * 12, u + b * 4, w, hw, T(dt), A + b * 144, Bm ? Bm + b * 48 : nullptr);
        if (H) quad_hess<T>(x + b * 12, u + b * 4, w, hw, T(dt), lam + b * 12, H + b * 144);
    }
}
EXPORT int hs_quad(int dtype, int64_t Bsz, const void* x, const void* u, const double* wind, double dt,
                   const void* lam, void* xd, void* A, void* B, void* H) {
    if (dtype) quad_all<double>(Bsz, (const double*)x, (const double*)u, wind, dt, (const double*)lam, (double*)xd,
                                (double*)A, (double*)B, (double*)H);
    else quad_all<float>(Bsz, (const float*)x, (const float*)u, wind, dt, (const float*)lam, (float*)xd, (float*)A,
                         (float*)B, (float*)H);
    return 0;
}

static RollP R_(int64_t Bsz, int N, const zb_model* model, const zb_cost* cost, const void* x0, const void* l,
                const void* L, const void* xPrev, const void* uPrev, void* xTraj, void* uTraj) {
    RollP P;
    P.Bsz = Bsz; P.N = N; P.M = M_(model); P.C = C_(cost); P.has_cost = cost != nullptr;
    P.x0 = x0; P.l = l; P.L = L; P.xPrev = xPrev; P.uPrev = uPrev; P.xTraj = xTraj; P.uTraj = uTraj; P.J = nullptr;
    return P;
}

EXPORT int hs_rollout(int dtype, int64_t Bsz, int N, const zb_model* model, const zb_cost* cost, const void* x0,
                      const void* l, const void* L, const void* xPrev, const void* uPrev, double alpha, void* xTraj,
                      void* uTraj, void* J) {
    RollP P = R_(Bsz, N, model, J ? cost : nullptr, x0, l, L, xPrev, uPrev, xTraj, uTraj);
    for (int64_t b = 0; b < Bsz; ++b) {
        if (dtype) { double v = rollout_any<double>(P, b, alpha, true); if (J) ((double*)J)[b] = v; }
        else { float v = rollout_any<float>(P, b, (float)alpha, true); if (J) ((float*)J)[b] = v; }
    }
    return 0;
}

template <typename T>
static void fwd(const RollP& P, T* J, int32_t* idx, T* Jall) {
    for (int64_t b = 0; b < P.Bsz; ++b) {
        for (int j = 0; j < 16; ++j) {
            T alpha = T(1);
            for (int i = 0; i < j; ++i) alpha *= T(0.5);
            Jall[b * 16 + j] = rollout_any<T>(P, b, alpha, false);
        }
        int k = argmin16<T>(Jall + b * 16);
        T alpha = T(1);
        for (int i = 0; i < k; ++i) alpha *= T(0.5);
        rollout_any<T>(P, b, alpha, true);
        J[b] = Jall[b * 16 + k];
        if (idx) idx[b] = k;
    }
}
EXPORT int hs_forward_pass(int dtype, int64_t Bsz, int N, const zb_model* model, const zb_cost* cost, const void* x0,
                           const void* l, const void* L, const void* xPrev, const void* uPrev, void* xTraj,
                           void* uTraj, void* J, int32_t* idx, void* Jall) {
    RollP P = R_(Bsz, N, model, cost, x0, l, L, xPrev, uPrev, xTraj, uTraj);
    if (dtype) fwd<double>(P, (double*)J, idx, (double*)Jall);
    else fwd<float>(P, (float*)J, idx, (float*)Jall);
    return 0;
}

EXPORT int hs_backward(int dtype, int64_t Bsz, int N, int n, int m, int second_order, const zb_arr* f_x,
                       const zb_arr* f_u, const zb_arr* f_xx, const zb_arr* f_ux, const zb_arr* f_uu, const zb_arr* c,
                       const zb_arr* c_x, const zb_arr* c_u, const zb_arr* c_xx, const zb_arr* c_ux,
                       const zb_arr* c_uu, const zb_arr* v, const zb_arr* v_x, const zb_arr* v_xx, void* l, void* L,
                       void* v_out, void* vx_out, void* vxx_out) {
    BackP P{Bsz, N, n, m, second_order, A_(f_x), A_(f_u), A_(f_xx), A_(f_ux), A_(f_uu), A_(c), A_(c_x), A_(c_u),
            A_(c_xx), A_(c_ux), A_(c_uu), A_(v), A_(v_x), A_(v_xx), l, L, v_out, vx_out, vxx_out, 1e-3};
    for (int64_t b = 0; b < Bsz; ++b) dtype ? backward_problem<double>(P, b) : backward_problem<float>(P, b);
    return 0;
}

template <typename T>
static void pdc(int64_t Bsz, int p, double eps, const T* in, T* out) {
    std::vector<T> S(p * p), W(p * p);
    for (int64_t b = 0; b < Bsz; ++b) {
        for (int i = 0; i < p * p; ++i) S[i] = in[b * p * p + i];
        pd_clamp<T>(S.data(), W.data(), p, T(eps));
        for (int i = 0; i < p * p; ++i) out[b * p * p + i] = S[i];
    }
}
EXPORT int hs_pd_clamp(int dtype, int64_t Bsz, int p, double eps, const void* in, void* out) {
    if (dtype) pdc<double>(Bsz, p, eps, (const double*)in, (double*)out);
    else pdc<float>(Bsz, p, eps, (const float*)in, (float*)out);
    return 0;
}

// mirror of the zb_ilqr_solve launch sequence (zb_api.cu), problem by problem
template <typename T>
static void solve(int64_t Bsz, int N, int second_order, const zb_model* model, const zb_cost* cost, const T* x0,
                  const T* uGuess, int maxIter, double tol, T* xTraj, T* uTraj, T* Lout, T* J, uint8_t* conv,
                  int32_t* iters, int32_t* alpha_log, T* J_log) {
    Model M = M_(model);
    Cost C = C_(cost);
    const int n = M.n, m = M.m, p = n + m;
    std::vector<T> l((size_t)Bsz * N * m), Jall((size_t)Bsz * 16), Czz((size_t)Bsz * p * p), Vf((size_t)Bsz * n * n);
    std::vector<T> zl((size_t)Bsz * N * m, T(0)), zx((size_t)Bsz * (N + 1) * n, T(0)), zu((size_t)Bsz * N * m, T(0));
    std::memset(Lout, 0, sizeof(T) * Bsz * N * m * n);
    for (int64_t b = 0; b < Bsz; ++b) {
        std::vector<T> S(p * p, T(0)), W(p * p);
        const T *Q = C.Q.at<T>(b), *R = C.R.at<T>(b), *Qf = C.Qf.at<T>(b);
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) S[i * p + j] = Q[i * n + j] + Q[j * n + i];
        for (int i = 0; i < m; ++i) for (int j = 0; j < m; ++j) S[(n + i) * p + n + j] = R[i * m + j] + R[j * m + i];
        pd_clamp<T>(S.data(), W.data(), p, T(1e-3));
        for (int i = 0; i < p * p; ++i) Czz[b * p * p + i] = S[i];
        for (int i = 0; i < n; ++i) for (int j = 0; j < n; ++j) S[i * n + j] = Qf[i * n + j] + Qf[j * n + i];
        pd_clamp<T>(S.data(), W.data(), n, T(1e-3));
        for (int i = 0; i < n * n; ++i) Vf[b * n * n + i] = S[i];
    }
    // initial rollout: l = uGuess, L = 0, previous trajectory = 0, alpha = 1
    RollP P0 = R_(Bsz, N, model, cost, x0, uGuess, Lout, zx.data(), zu.data(), xTraj, uTraj);
    for (int64_t b = 0; b < Bsz; ++b) {
        J[b] = rollout_any<T>(P0, b, T(1), true);
        conv[b] = 0; iters[b] = 0;
        if (alpha_log) for (int i = 0; i < maxIter; ++i) alpha_log[b * maxIter + i] = -1;
        if (J_log) { J_log[b * (maxIter + 1)] = J[b]; for (int i = 1; i <= maxIter; ++i) J_log[b * (maxIter + 1) + i] = NAN; }
    }
    RollP P = R_(Bsz, N, model, cost, x0, l.data(), Lout, xTraj, uTraj, xTraj, uTraj);
    SolveBackP Bk{Bsz, N, second_order, M, C, xTraj, uTraj, Czz.data(), Vf.data(), conv, l.data(), Lout, 1e-3};
    for (int it = 0; it < maxIter; ++it)
        for (int64_t b = 0; b < Bsz; ++b) {
            if (conv[b]) continue;
            solve_backward_problem<T>(Bk, b);
            for (int j = 0; j < 16; ++j) {
                T alpha = T(1);
                for (int i = 0; i < j; ++i) alpha *= T(0.5);
                Jall[b * 16 + j] = rollout_any<T>(P, b, alpha, false);
            }
            int k = argmin16<T>(&Jall[b * 16]);
            T alpha = T(1);
            for (int i = 0; i < k; ++i) alpha *= T(0.5);
            rollout_any<T>(P, b, alpha, true);
            T Jn = Jall[b * 16 + k];
            conv[b] = (fabs(J[b] - Jn) <= T(tol)) ? 1 : 0;
            J[b] = Jn;
            iters[b] = it + 1;
            if (alpha_log) alpha_log[b * maxIter + it] = k;
            if (J_log) J_log[b * (maxIter + 1) + it + 1] = Jn;
        }
}
EXPORT int hs_ilqr_solve(int dtype, int64_t Bsz, int N, int second_order, const zb_model* model, const zb_cost* cost,
                         const void* x0, const void* uGuess, int maxIter, double tol, void* xTraj, void* uTraj,
                         void* L, void* J, uint8_t* conv, int32_t* iters, int32_t* alpha_log, void* J_log) {
    if (dtype) solve<double>(Bsz, N, second_order, model, cost, (const double*)x0, (const double*)uGuess, maxIter, tol,
                             (double*)xTraj, (double*)uTraj, (double*)L, (double*)J, conv, iters, alpha_log, (double*)J_log);
    else solve<float>(Bsz, N, second_order, model, cost, (const float*)x0, (const float*)uGuess, maxIter, tol,
                      (float*)xTraj, (float*)uTraj, (float*)L, (float*)J, conv, iters, alpha_log, (float*)J_log);
    return 0;
}

EXPORT long long hs_admm_ws_elems(int N, int n, int m) { return admm_ws_elems(N, n, m); }

EXPORT int hs_mpc_admm(int dtype, int64_t Bsz, int N, int n, int m, const zb_arr* A, const zb_arr* B, const zb_arr* Q,
                       const zb_arr* R, const zb_arr* Qf, const zb_arr* xlb, const zb_arr* xub, const zb_arr* ulb,
                       const zb_arr* uub, const void* x0, int max_iter, int check_every, double rho, double alpha,
                       double eps_abs, double eps_rel, double eps_inf, void* u0, void* xTraj, void* uTraj,
                       int8_t* status, int32_t* iters, void* ws) {
    AdmmP P{};
    P.Bsz = Bsz; P.N = N; P.n = n; P.m = m;
    P.A = A_(A); P.B = A_(B); P.Q = A_(Q); P.R = A_(R); P.Qf = A_(Qf);
    P.xlb = A_(xlb); P.xub = A_(xub); P.ulb = A_(ulb); P.uub = A_(uub);
    P.x0 = x0; P.u0 = u0; P.xTraj = xTraj; P.uTraj = uTraj; P.status = status; P.iters = iters;
    P.ws = ws; P.ws_stride = admm_ws_elems(N, n, m);
    P.max_iter = max_iter; P.check_every = check_every; P.rho = rho; P.alpha = alpha;
    P.eps_abs = eps_abs; P.eps_rel = eps_rel; P.eps_inf = eps_inf;
    for (int64_t b = 0; b < Bsz; ++b) dtype ? admm_problem<double>(P, b) : admm_problem<float>(P, b);
    return 0;
}

// host run of the shared-definition box-constrained lqrMpc kernel body (zopt_b200/csrc/mpc_box.cuh): tables + per-problem ADMM
template <typename T>
static void box_run(int64_t Bsz, int N, const double* A, const double* B, const double* Q, const double* R, const double* Qf,
                    const double* xlb, const double* xub, const double* ulb, const double* uub, const void* x0, int max_iter,
                    int check_every, double rho, double alpha, double eps_abs, double eps_rel, double eps_inf, void* u0,
                    void* xTraj, void* uTraj, int8_t* status, int32_t* iters, int Tsim = -1, void* xSim = nullptr,
                    void* uSim = nullptr, double clip = 0) {
    box::Ops<T> O;
    box::Costs<T> C;
    for (int i = 0; i < 144; ++i) { O.A[i] = (T)A[i]; C.Q[i] = (T)Q[i]; C.Qf[i] = (T)Qf[i]; }
    for (int i = 0; i < 48; ++i) O.B[i] = (T)B[i];
    for (int i = 0; i < 16; ++i) C.R[i] = (T)R[i];
    for (int i = 0; i < 12; ++i) { O.xlb[i] = (T)xlb[i]; O.xub[i] = (T)xub[i]; }
    for (int i = 0; i < 4; ++i) { O.ulb[i] = (T)ulb[i]; O.uub[i] = (T)uub[i]; }
    std::vector<T> tab(box::tab_elems(N)), ws(box::ws_elems(N, Bsz)), sm(592);
    for (int lv = 0; lv < box::LEVELS; ++lv)
        box::table_level<T>(O, C, N, (T)std::ldexp(rho, lv - box::LEVEL0), tab.data() + (long long)lv * N * box::TW, sm.data(), 0, 1, [] {});
    box::Params<T> P{};
    P.Bsz = Bsz; P.N = N;
    P.x0 = (const T*)x0; P.u0 = (T*)u0; P.xTraj = (T*)xTraj; P.uTraj = (T*)uTraj; P.status = status; P.iters = iters;
    P.ws = ws.data(); P.tab = tab.data();
    P.max_iter = max_iter; P.check_every = check_every;
    P.rho0 = (T)rho; P.alpha = (T)alpha; P.eps_abs = (T)eps_abs; P.eps_rel = (T)eps_rel; P.eps_inf = (T)eps_inf;
    if (Tsim >= 0) {
        P.Tsim = Tsim; P.xSim = (T*)xSim; P.uSim = (T*)uSim; P.clip_margin = (T)clip;
        for (int64_t b = 0; b < Bsz; ++b) box::closed_loop_problem<T>(O, P, b);
    } else {
        for (int64_t b = 0; b < Bsz; ++b) box::problem<T>(O, P, b);
    }
}

EXPORT int hs_mpc_box(int dtype, int64_t Bsz, int N, const double* A, const double* B, const double* Q, const double* R,
                      const double* Qf, const double* xlb, const double* xub, const double* ulb, const double* uub,
                      const void* x0, int max_iter, int check_every, double rho, double alpha, double eps_abs, double eps_rel,
                      double eps_inf, void* u0, void* xTraj, void* uTraj, int8_t* status, int32_t* iters) {
    if (dtype) box_run<double>(Bsz, N, A, B, Q, R, Qf, xlb, xub, ulb, uub, x0, max_iter, check_every, rho, alpha, eps_abs, eps_rel, eps_inf, u0, xTraj, uTraj, status, iters);
    else box_run<float>(Bsz, N, A, B, Q, R, Qf, xlb, xub, ulb, uub, x0, max_iter, check_every, rho, alpha, eps_abs, eps_rel, eps_inf, u0, xTraj, uTraj, status, iters);
    return 0;
}

EXPORT int hs_mpc_box_closed_loop(int dtype, int64_t Bsz, int N, int Tsim, const double* A, const double* B, const double* Q,
                                  const double* R, const double* Qf, const double* xlb, const double* xub, const double* ulb,
                                  const double* uub, const void* x0, int max_iter, int check_every, double rho, double alpha,
                                  double eps_abs, double eps_rel, double eps_inf, double clip, void* xSim, void* uSim,
                                  int8_t* status, int32_t* iters) {
    if (dtype) {
        std::vector<double> px((size_t)Bsz * (N + 1) * 12), pu((size_t)Bsz * N * 4);
        box_run<double>(Bsz, N, A, B, Q, R, Qf, xlb, xub, ulb, uub, x0, max_iter, check_every, rho, alpha, eps_abs, eps_rel, eps_inf, nullptr, px.data(), pu.data(), status, iters, Tsim, xSim, uSim, clip);
    } else {
        std::vector<float> px((size_t)Bsz * (N + 1) * 12), pu((size_t)Bsz * N * 4);
        box_run<float>(Bsz, N, A, B, Q, R, Qf, xlb, xub, ulb, uub, x0, max_iter, check_every, rho, alpha, eps_abs, eps_rel, eps_inf, nullptr, px.data(), pu.data(), status, iters, Tsim, xSim, uSim, clip);
    }
    return 0;
}

// ---- the (8,4) fp32 thread-per-problem kernel (zopt_b200/csrc/lqr_s84.cuh), run lane by lane on the host -------------------
// The kernel body is compiled as is: CUDA's qualifiers are defined away, float4 / __ldg / rsqrtf / the packed FMA get host
// equivalents, the shared-memory slab becomes a static array (each lane touches its own column only, so lanes can run one after
// the other) and threadIdx / blockIdx are plain globals set by the driver loop below.
struct float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
#define __device__
#define __forceinline__ inline
#define __global__
#define __shared__ static
#define __launch_bounds__(x)
template <typename T> static inline T __ldg(const T* p) { return *p; }
static inline float rsqrtf(float x) { return 1.0f / std::sqrt(x); }
static inline double rsqrt(double x) { return 1.0 / std::sqrt(x); }
using std::fmaf;
struct HsDim { unsigned x; };
static HsDim threadIdx, blockIdx;
namespace zb {
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
namespace t1 {
constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
inline void fma2(float& d0, float& d1, float a, float b0, float b1) { d0 = std::fmaf(a, b0, d0); d1 = std::fmaf(a, b1, d1); }
}  // namespace t1
}  // namespace zb
#define ZB_F4(v, e) ((e) == 0 ? (v).x : (e) == 1 ? (v).y : (e) == 2 ? (v).z : (v).w)
#include "../../zopt_b200/csrc/lqr_s84.cuh"
struct double2 { double x, y; };
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
#define __host__
using std::fma;
#include "../../zopt_b200/csrc/lqr_s84d.cuh"

EXPORT int hs_riccati_s84(int bilinear, int64_t Bsz, int N, int T, const zb_arr* A, const zb_arr* B, const zb_arr* d, const zb_arr* Q,
                          const zb_arr* R, const zb_arr* H, const zb_arr* q, const zb_arr* r, float* L, float* l, float* V0) {
    s84::S84P P{};
    P.Bsz = Bsz; P.N = N; P.T = T;
    P.A = A_(A); P.B = A_(B); P.Q = A_(Q); P.R = A_(R);
    if (bilinear) { P.H = A_(H); P.d = A_(d); P.q = A_(q); P.r = A_(r); }
    P.L = L; P.l = l; P.V0 = V0;
    for (unsigned blk = 0; blk < (Bsz + 31) / 32; ++blk)
        for (unsigned lane = 0; lane < 32; ++lane) {
            blockIdx.x = blk;
            threadIdx.x = lane;
            if (bilinear) s84::k_riccati_s84<true>(P);
            else s84::k_riccati_s84<false>(P);
        }
    return 0;
}

EXPORT int hs_riccati_s84d(int bilinear, int64_t Bsz, int N, int T, const zb_arr* A, const zb_arr* B, const zb_arr* d, const zb_arr* Q,
                           const zb_arr* R, const zb_arr* H, const zb_arr* q, const zb_arr* r, double* L, double* l, double* V0) {
    s84d::S84DP P{};
    P.Bsz = Bsz; P.N = N; P.T = T;
    P.A = A_(A); P.B = A_(B); P.Q = A_(Q); P.R = A_(R);
    if (bilinear) { P.H = A_(H); P.d = A_(d); P.q = A_(q); P.r = A_(r); }
    P.L = L; P.l = l; P.V0 = V0;
    for (unsigned blk = 0; blk < (Bsz + 31) / 32; ++blk)
        for (unsigned lane = 0; lane < 32; ++lane) {
            blockIdx.x = blk;
            threadIdx.x = lane;
            if (bilinear) s84d::k_riccati_s84d<true>(P);
            else s84d::k_riccati_s84d<false>(P);
        }
    return 0;
}
