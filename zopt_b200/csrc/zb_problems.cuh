// Per-problem bodies of the generic kernels: problem `b` of a strided batch, run-time (n, m, N).
// A CUDA kernel maps one thread to one problem and calls these; tests/hostsim compiles the same
// bodies for the host to check the arithmetic against the oracle without a GPU.
#pragma once

#include <stdint.h>

#include "zb_steps.cuh"

namespace zb {

struct Arr {  // device-side mirror of zb_arr
    const void* p;
    long long sb, st;
    template <typename T>
    ZB_HD const T* at(long long b, long long k = 0) const {
        return reinterpret_cast<const T*>(p) + b * sb + k * st;
    }
    ZB_HD bool null() const { return p == nullptr; }
};

// -------------------------------------------------------------------------------------------------
struct LqrP {
    long long Bsz;
    int N, T, n, m;
    Arr A, B, Q, R;
    void* L;   // (Bsz,N,m,n)
    void* V0;  // (Bsz,n,n) or null
};

template <typename T>
ZB_HD void lqr_problem(const LqrP& P, long long b) {
    const int n = P.n, m = P.m;
    T V[NX * NX], L[NU * NX];
    const T* Qf = P.Q.at<T>(b, P.T - 1);  // lqrUtils.py:172: terminal value is Q[-1]
    for (int i = 0; i < n * n; ++i) V[i] = Qf[i];
    T* Lout = reinterpret_cast<T*>(P.L) + b * (long long)P.N * m * n;
    for (int k = P.N - 1; k >= 0; --k) {
        lqr_joseph_step<T>(n, m, P.A.at<T>(b, k), P.B.at<T>(b, k), P.Q.at<T>(b, k), P.R.at<T>(b, k), V, L);
        for (int i = 0; i < m * n; ++i) Lout[(long long)k * m * n + i] = L[i];
    }
    if (P.V0) {
        T* V0 = reinterpret_cast<T*>(P.V0) + b * n * n;
        for (int i = 0; i < n * n; ++i) V0[i] = V[i];
    }
}

// -------------------------------------------------------------------------------------------------
struct BilinP {
    long long Bsz;
    int N, T, n, m;
    Arr A, B, d, Q, R, H, q, r, q0;
    void *L, *l;
};

template <typename T>
ZB_HD void bilinear_problem(const BilinP& P, long long b) {
    const int n = P.n, m = P.m;
    T V[NX * NX], v[NX], L[NU * NX], l[NU];
    // lqrUtils.py:261: initial carry (Q[-1], q[-1], q0[-1])
    const T* Qf = P.Q.at<T>(b, P.T - 1);
    const T* qf = P.q.at<T>(b, P.T - 1);
    for (int i = 0; i < n * n; ++i) V[i] = Qf[i];
    for (int i = 0; i < n; ++i) v[i] = qf[i];
    T v0 = *P.q0.at<T>(b, P.T - 1);
    T* Lout = reinterpret_cast<T*>(P.L) + b * (long long)P.N * m * n;
    T* lout = reinterpret_cast<T*>(P.l) + b * (long long)P.N * m;
    for (int k = P.N - 1; k >= 0; --k) {
        bilinear_step<T>(n, m, P.A.at<T>(b, k), P.B.at<T>(b, k), P.d.at<T>(b, k), P.Q.at<T>(b, k), P.R.at<T>(b, k),
                         P.H.at<T>(b, k), P.q.at<T>(b, k), P.r.at<T>(b, k), *P.q0.at<T>(b, k), V, v, v0, L, l);
        for (int i = 0; i < m * n; ++i) Lout[(long long)k * m * n + i] = L[i];
        for (int i = 0; i < m; ++i) lout[(long long)k * m + i] = l[i];
    }
}

// -------------------------------------------------------------------------------------------------
// Registered dynamics / cost (device-side mirror of zb_model / zb_cost)
struct Model {
    int kind, n, m, has_wind;
    double dt;
    double wind[3];
    Arr A, B;
};
struct Cost {
    Arr Q, R, Qf;
};

template <typename T>
ZB_HD void model_step(const Model& M, long long b, const T* x, const T* u, T* xn) {
    if (M.kind == 1) {
        T w[3] = {T(M.wind[0]), T(M.wind[1]), T(M.wind[2])};
        quad_euler<T>(x, u, w, M.has_wind != 0, T(M.dt), xn);
    } else {
        const T* A = M.A.at<T>(b);
        const T* B = M.B.at<T>(b);
        T t[NX];
        for (int i = 0; i < M.n; ++i) {
            T s = T(0);
            for (int j = 0; j < M.n; ++j) s += A[i * M.n + j] * x[j];
            for (int j = 0; j < M.m; ++j) s += B[i * M.m + j] * u[j];
            t[i] = s;
        }
        for (int i = 0; i < M.n; ++i) xn[i] = t[i];
    }
}

// f_x (n x n), f_u (n x m) at (x,u)
template <typename T>
ZB_HD void model_lin(const Model& M, long long b, const T* x, const T* u, T* fx, T* fu) {
    if (M.kind == 1) {
        T w[3] = {T(M.wind[0]), T(M.wind[1]), T(M.wind[2])};
        quad_lin<T>(x, u, w, M.has_wind != 0, T(M.dt), fx, fu);
    } else {
        const T* A = M.A.at<T>(b);
        const T* B = M.B.at<T>(b);
        for (int i = 0; i < M.n * M.n; ++i) fx[i] = A[i];
        for (int i = 0; i < M.n * M.m; ++i) fu[i] = B[i];
    }
}

// -------------------------------------------------------------------------------------------------
// trajectoryRollout (ilqrUtils.py:33-66) with AffinePolicy (pytrees.py:215-220) and, optionally, the
// trajectory cost (pytrees.py:40-55).  `write` = store the trajectory.
struct RollP {
    long long Bsz;
    int N;
    Model M;
    Cost C;
    int has_cost;
    const void *x0, *l, *L, *xPrev, *uPrev;
    void *xTraj, *uTraj, *J;
};

template <typename T>
ZB_HD T rollout_core(const RollP& P, long long b, T alpha, bool write) {
    const int n = P.M.n, m = P.M.m, N = P.N;
    const T* x0 = reinterpret_cast<const T*>(P.x0) + b * n;
    const T* lp = reinterpret_cast<const T*>(P.l) + b * (long long)N * m;
    const T* Lp = reinterpret_cast<const T*>(P.L) + b * (long long)N * m * n;
    const T* xP = reinterpret_cast<const T*>(P.xPrev) + b * (long long)(N + 1) * n;
    const T* uP = reinterpret_cast<const T*>(P.uPrev) + b * (long long)N * m;
    T* xT = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * n;
    T* uT = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * m;
    const T* Q = P.has_cost ? P.C.Q.at<T>(b) : nullptr;
    const T* R = P.has_cost ? P.C.R.at<T>(b) : nullptr;
    T x[NX], u[NU], dx[NX];
    T J = T(0);
    for (int i = 0; i < n; ++i) x[i] = x0[i];
    // xTraj/uTraj may alias xPrev/uPrev (in-place update by the solver): row k is stored only after
    // row k of the previous trajectory has been consumed.
    for (int k = 0; k < N; ++k) {
        for (int i = 0; i < n; ++i) dx[i] = x[i] - xP[(long long)k * n + i];
        for (int i = 0; i < m; ++i) {
            T s = T(0);
            for (int j = 0; j < n; ++j) s += Lp[((long long)k * m + i) * n + j] * dx[j];
            u[i] = (alpha * lp[(long long)k * m + i] + s) + uP[(long long)k * m + i];
        }
        if (write) {
            for (int i = 0; i < n; ++i) xT[(long long)k * n + i] = x[i];
            for (int i = 0; i < m; ++i) uT[(long long)k * m + i] = u[i];
        }
        if (P.has_cost) J += quad_form<T>(Q, x, n) + quad_form<T>(R, u, m);
        model_step<T>(P.M, b, x, u, x);
    }
    if (write)
        for (int i = 0; i < n; ++i) xT[(long long)N * n + i] = x[i];
    if (P.has_cost) J += quad_form<T>(P.C.Qf.at<T>(b), x, n);
    return J;
}

// argmin with NumPy/JAX semantics: first minimum on ties, a NaN wins (ilqrUtils.py:147)
template <typename T>
ZB_HD int argmin16(const T* J) {
    int best = 0;
    for (int j = 0; j < 16; ++j)
        if (J[j] != J[j]) return j;
    for (int j = 1; j < 16; ++j)
        if (J[j] < J[best]) best = j;
    return best;
}

// -------------------------------------------------------------------------------------------------
// backwardPass_ilqr / backwardPass_ddp on explicit stacked pytrees (ilqrUtils.py:176-181, 209-214)
struct BackP {
    long long Bsz;
    int N, n, m, second_order;
    Arr f_x, f_u, f_xx, f_ux, f_uu, c, c_x, c_u, c_xx, c_ux, c_uu, v, v_x, v_xx;
    void *l, *L, *v_out, *vx_out, *vxx_out;
    double eps;
};

// conditionQuadraticDynamics (ilqrUtils.py:237-251): contraction with v_x then eigen-clamp of the (n+m) block
template <typename T>
ZB_HD void ddp_condition(int n, int m, const T* f_xx, const T* f_ux, const T* f_uu, const T* v_x, T eps, T* vf_xx,
                         T* vf_ux, T* vf_uu) {
    const int p = n + m;
    T Z[ZB_PD_MAX * ZB_PD_MAX], W[ZB_PD_MAX * ZB_PD_MAX];
    for (int j = 0; j < p; ++j)
        for (int k = 0; k < p; ++k) {
            T s = T(0);
            if (j < n && k < n) {
                for (int i = 0; i < n; ++i) s += v_x[i] * f_xx[(i * n + j) * n + k];
            } else if (j >= n && k < n) {
                for (int i = 0; i < n; ++i) s += v_x[i] * f_ux[(i * m + (j - n)) * n + k];
            } else if (j < n && k >= n) {  // upper-right block is vf_ux^T
                for (int i = 0; i < n; ++i) s += v_x[i] * f_ux[(i * m + (k - n)) * n + j];
            } else {
                for (int i = 0; i < n; ++i) s += v_x[i] * f_uu[(i * m + (j - n)) * m + (k - n)];
            }
            Z[j * p + k] = s;
        }
    pd_clamp<T>(Z, W, p, eps);
    for (int j = 0; j < n; ++j)
        for (int k = 0; k < n; ++k) vf_xx[j * n + k] = Z[j * p + k];
    for (int j = 0; j < m; ++j) {
        for (int k = 0; k < n; ++k) vf_ux[j * n + k] = Z[(n + j) * p + k];
        for (int k = 0; k < m; ++k) vf_uu[j * m + k] = Z[(n + j) * p + n + k];
    }
}

template <typename T>
ZB_HD void backward_problem(const BackP& P, long long b) {
    const int n = P.n, m = P.m, N = P.N;
    T v, v_x[NX], v_xx[NX * NX], l[NU], L[NU * NX];
    T vf_xx[NX * NX], vf_ux[NU * NX], vf_uu[NU * NU];
    v = *P.v.at<T>(b);
    for (int i = 0; i < n; ++i) v_x[i] = P.v_x.at<T>(b)[i];
    for (int i = 0; i < n * n; ++i) v_xx[i] = P.v_xx.at<T>(b)[i];
    T* lo = reinterpret_cast<T*>(P.l) + b * (long long)N * m;
    T* Lo = reinterpret_cast<T*>(P.L) + b * (long long)N * m * n;
    for (int k = N - 1; k >= 0; --k) {
        if (P.second_order)
            ddp_condition<T>(n, m, P.f_xx.at<T>(b, k), P.f_ux.at<T>(b, k), P.f_uu.at<T>(b, k), v_x, T(P.eps), vf_xx,
                             vf_ux, vf_uu);
        ilqr_step<T>(n, m, P.f_x.at<T>(b, k), P.f_u.at<T>(b, k), *P.c.at<T>(b, k), P.c_x.at<T>(b, k),
                     P.c_u.at<T>(b, k), P.c_xx.at<T>(b, k), P.c_ux.at<T>(b, k), P.c_uu.at<T>(b, k),
                     P.second_order ? vf_xx : nullptr, vf_ux, vf_uu, v, v_x, v_xx, l, L);
        for (int i = 0; i < m; ++i) lo[(long long)k * m + i] = l[i];
        for (int i = 0; i < m * n; ++i) Lo[(long long)k * m * n + i] = L[i];
    }
    if (P.v_out) reinterpret_cast<T*>(P.v_out)[b] = v;
    if (P.vx_out)
        for (int i = 0; i < n; ++i) reinterpret_cast<T*>(P.vx_out)[b * n + i] = v_x[i];
    if (P.vxx_out)
        for (int i = 0; i < n * n; ++i) reinterpret_cast<T*>(P.vxx_out)[b * n * n + i] = v_xx[i];
}

// -------------------------------------------------------------------------------------------------
// One iLQR / DDP backward pass for a registered model + quadratic cost, linearising on the fly along
// the current trajectory (ilqrUtils.py:308-315 / :378-385): nothing of the expansion is materialised.
//   c_x = (Q+Q^T) x, c_u = (R+R^T) u, c_zz = clampPD(blockdiag(Q+Q^T, R+R^T)) (precomputed, Czz),
//   Vf = (x_N'Qf x_N, (Qf+Qf^T) x_N, clampPD(Qf+Qf^T)) (Vfxx precomputed).
struct SolveBackP {
    long long Bsz;
    int N, second_order;
    Model M;
    Cost C;
    const void *xTraj, *uTraj;  // current trajectory
    const void* Czz;            // (Bsz, p, p) conditioned running-cost Hessian, p = n+m
    const void* Vfxx;           // (Bsz, n, n) conditioned terminal Hessian
    const uint8_t* done;        // (Bsz) frozen problems are skipped
    void *l, *L;
    double eps;
};

template <typename T>
ZB_HD void solve_backward_problem(const SolveBackP& P, long long b) {
    if (P.done && P.done[b]) return;
    const int n = P.M.n, m = P.M.m, N = P.N, p = n + m;
    const T* xT = reinterpret_cast<const T*>(P.xTraj) + b * (long long)(N + 1) * n;
    const T* uT = reinterpret_cast<const T*>(P.uTraj) + b * (long long)N * m;
    const T* Czz = reinterpret_cast<const T*>(P.Czz) + b * (long long)p * p;
    const T* Q = P.C.Q.at<T>(b);
    const T* R = P.C.R.at<T>(b);
    const T* Qf = P.C.Qf.at<T>(b);
    T v, v_x[NX], v_xx[NX * NX], l[NU], L[NU * NX];
    T fx[NX * NX], fu[NX * NU], c_x[NX], c_u[NU], c_xx[NX * NX], c_ux[NU * NX], c_uu[NU * NU];
    T vf_xx[NX * NX], vf_ux[NU * NX], vf_uu[NU * NU];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) c_xx[i * n + j] = Czz[i * p + j];
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j) c_ux[i * n + j] = Czz[(n + i) * p + j];
        for (int j = 0; j < m; ++j) c_uu[i * m + j] = Czz[(n + i) * p + n + j];
    }
    const T* xN = xT + (long long)N * n;
    v = quad_form<T>(Qf, xN, n);
    for (int i = 0; i < n; ++i) {
        T s = T(0);
        for (int j = 0; j < n; ++j) s += (Qf[i * n + j] + Qf[j * n + i]) * xN[j];
        v_x[i] = s;
    }
    const T* Vf = reinterpret_cast<const T*>(P.Vfxx) + b * (long long)n * n;
    for (int i = 0; i < n * n; ++i) v_xx[i] = Vf[i];
    T* lo = reinterpret_cast<T*>(P.l) + b * (long long)N * m;
    T* Lo = reinterpret_cast<T*>(P.L) + b * (long long)N * m * n;
    for (int k = N - 1; k >= 0; --k) {
        const T* x = xT + (long long)k * n;
        const T* u = uT + (long long)k * m;
        model_lin<T>(P.M, b, x, u, fx, fu);
        for (int i = 0; i < n; ++i) {
            T s = T(0);
            for (int j = 0; j < n; ++j) s += (Q[i * n + j] + Q[j * n + i]) * x[j];
            c_x[i] = s;
        }
        for (int i = 0; i < m; ++i) {
            T s = T(0);
            for (int j = 0; j < m; ++j) s += (R[i * m + j] + R[j * m + i]) * u[j];
            c_u[i] = s;
        }
        T c = quad_form<T>(Q, x, n) + quad_form<T>(R, u, m);
        if (P.second_order) {
            // conditionQuadraticDynamics (ilqrUtils.py:237-251); f_ux = f_uu = 0 for both registered models
            T Z[ZB_PD_MAX * ZB_PD_MAX], W[ZB_PD_MAX * ZB_PD_MAX];
            for (int i = 0; i < p * p; ++i) Z[i] = T(0);
            if (P.M.kind == 1) {
                T w[3] = {T(P.M.wind[0]), T(P.M.wind[1]), T(P.M.wind[2])};
                T H[144];
                quad_hess<T>(x, u, w, P.M.has_wind != 0, T(P.M.dt), v_x, H);
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) Z[i * p + j] = H[i * n + j];
            }
            pd_clamp<T>(Z, W, p, T(P.eps));
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) vf_xx[i * n + j] = Z[i * p + j];
            for (int i = 0; i < m; ++i) {
                for (int j = 0; j < n; ++j) vf_ux[i * n + j] = Z[(n + i) * p + j];
                for (int j = 0; j < m; ++j) vf_uu[i * m + j] = Z[(n + i) * p + n + j];
            }
        }
        ilqr_step<T>(n, m, fx, fu, c, c_x, c_u, c_xx, c_ux, c_uu, P.second_order ? vf_xx : nullptr, vf_ux, vf_uu, v,
                     v_x, v_xx, l, L);
        for (int i = 0; i < m; ++i) lo[(long long)k * m + i] = l[i];
        for (int i = 0; i < m * n; ++i) Lo[(long long)k * m * n + i] = L[i];
    }
}

}  // namespace zb
