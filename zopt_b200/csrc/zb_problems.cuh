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

template <typename T, int N_, int M_>
ZB_HD void lqr_problem_ct(const LqrP& P, long long b) {  // lqr_problem with compile-time (n, m): operands in registers
    T V[N_ * N_], L[M_ * N_];
    const T* Qf = P.Q.at<T>(b, P.T - 1);
    ZB_UNROLL
    for (int i = 0; i < N_ * N_; ++i) V[i] = Qf[i];
    T* Lout = reinterpret_cast<T*>(P.L) + b * (long long)P.N * M_ * N_;
    for (int k = P.N - 1; k >= 0; --k) {
        lqr_joseph_step_ct<T, N_, M_>(P.A.at<T>(b, k), P.B.at<T>(b, k), P.Q.at<T>(b, k), P.R.at<T>(b, k), V, L);
        ZB_UNROLL
        for (int i = 0; i < M_ * N_; ++i) Lout[(long long)k * M_ * N_ + i] = L[i];
    }
    if (P.V0) {
        T* V0 = reinterpret_cast<T*>(P.V0) + b * N_ * N_;
        ZB_UNROLL
        for (int i = 0; i < N_ * N_; ++i) V0[i] = V[i];
    }
}

// -------------------------------------------------------------------------------------------------
// Continuous-time finite-horizon LQR (zopt/lqrUtils.py:39-98): the Riccati differential equation of the LQR HJB,
//   -dV/dt = Q + V A + A' V - V B R^-1 B' V,   V(T) = Qf,
// integrated backward from T with the classical RK4 scheme on a uniform grid (`sub` steps per output interval; the reference
// uses jax's adaptive Dormand-Prince at rtol = atol = 1.4e-8 and reports on the same N-point grid).  The coefficient
// functions A(t), B(t), Q(t), R^-1(t) are arbitrary Python callables in the reference; here they arrive SAMPLED at the
// scheme's stage times (two samples per step + 1: sample j is time T - j h/2), shared by the batch or per problem.
struct CareP {
    long long Bsz;
    int N, sub, n, m;  // N output points, `sub` RK4 steps between two of them
    double h;          // step
    Arr A, B, Q, Rinv; // time series of 2 (N-1) sub + 1 samples (st = 0: constant)
    Arr Qf;
    void* V;           // (Bsz, N, n, n): V[i] = V(t_i), t_i = i T/(N-1)  (the reference's `out[::-1]`)
};

template <typename T>
ZB_HD void care_rhs(int n, int m, const T* A, const T* B, const T* Q, const T* Ri, const T* V, T* dV) {
    T VB[NX * NU], VBR[NX * NU], W[NX * NX];
    mm(VB, V, B, n, n, m);        // V B
    mm(VBR, VB, Ri, n, m, m);     // V B R^-1
    mm(W, V, A, n, n, n);         // V A
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            T s = Q[i * n + j] + W[i * n + j];
            for (int k = 0; k < n; ++k) s += A[k * n + i] * V[k * n + j];          // A' V
            for (int a = 0; a < m; ++a) s -= VBR[i * m + a] * VB[j * m + a];       // (V B R^-1)(V B)'   (V symmetric)
            dV[i * n + j] = s;
        }
}

template <typename T>
ZB_HD void care_problem(const CareP& P, long long b) {
    const int n = P.n, m = P.m, nn = n * n;
    T V[NX * NX], k1[NX * NX], k2[NX * NX], k3[NX * NX], k4[NX * NX], Vt[NX * NX];
    const T* Qf = P.Qf.at<T>(b);
    for (int i = 0; i < nn; ++i) V[i] = Qf[i];
    T* out = reinterpret_cast<T*>(P.V) + b * (long long)P.N * nn;
    for (int i = 0; i < nn; ++i) out[(long long)(P.N - 1) * nn + i] = V[i];  // V(T) = Qf
    const T h = T(P.h);
    long long j = 0;  // sample index: time T - j h/2
    for (int o = P.N - 2; o >= 0; --o) {
        for (int s = 0; s < P.sub; ++s, j += 2) {
            care_rhs<T>(n, m, P.A.at<T>(b, j), P.B.at<T>(b, j), P.Q.at<T>(b, j), P.Rinv.at<T>(b, j), V, k1);
            for (int i = 0; i < nn; ++i) Vt[i] = V[i] + T(0.5) * h * k1[i];
            care_rhs<T>(n, m, P.A.at<T>(b, j + 1), P.B.at<T>(b, j + 1), P.Q.at<T>(b, j + 1), P.Rinv.at<T>(b, j + 1), Vt, k2);
            for (int i = 0; i < nn; ++i) Vt[i] = V[i] + T(0.5) * h * k2[i];
            care_rhs<T>(n, m, P.A.at<T>(b, j + 1), P.B.at<T>(b, j + 1), P.Q.at<T>(b, j + 1), P.Rinv.at<T>(b, j + 1), Vt, k3);
            for (int i = 0; i < nn; ++i) Vt[i] = V[i] + h * k3[i];
            care_rhs<T>(n, m, P.A.at<T>(b, j + 2), P.B.at<T>(b, j + 2), P.Q.at<T>(b, j + 2), P.Rinv.at<T>(b, j + 2), Vt, k4);
            for (int i = 0; i < nn; ++i) V[i] += h / T(6) * (k1[i] + T(2) * k2[i] + T(2) * k3[i] + k4[i]);
        }
        for (int i = 0; i < nn; ++i) out[(long long)o * nn + i] = V[i];
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

template <typename T, int N_, int M_>
ZB_HD void bilinear_problem_ct(const BilinP& P, long long b) {  // bilinear_problem with compile-time (n, m)
    T V[N_ * N_], v[N_], L[M_ * N_], l[M_];
    const T* Qf = P.Q.at<T>(b, P.T - 1);
    const T* qf = P.q.at<T>(b, P.T - 1);
    ZB_UNROLL
    for (int i = 0; i < N_ * N_; ++i) V[i] = Qf[i];
    ZB_UNROLL
    for (int i = 0; i < N_; ++i) v[i] = qf[i];
    T* Lout = reinterpret_cast<T*>(P.L) + b * (long long)P.N * M_ * N_;
    T* lout = reinterpret_cast<T*>(P.l) + b * (long long)P.N * M_;
    for (int k = P.N - 1; k >= 0; --k) {
        bilinear_step_ct<T, N_, M_>(P.A.at<T>(b, k), P.B.at<T>(b, k), P.d.at<T>(b, k), P.Q.at<T>(b, k), P.R.at<T>(b, k),
                                    P.H.at<T>(b, k), P.q.at<T>(b, k), P.r.at<T>(b, k), V, v, L, l);
        ZB_UNROLL
        for (int i = 0; i < M_ * N_; ++i) Lout[(long long)k * M_ * N_ + i] = L[i];
        ZB_UNROLL
        for (int i = 0; i < M_; ++i) lout[(long long)k * M_ + i] = l[i];
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

// ZB_USER_MODEL: this translation unit is a plug-in compiled around a user-defined model (kind 2, csrc/zb_user_model.cu):
// user_step / user_lin / user_hess come from the generated header (zopt_b200/plugin.py: sympy -> CUDA).
template <typename T>
ZB_HD void model_step(const Model& M, long long b, const T* x, const T* u, T* xn) {
#ifdef ZB_USER_MODEL
    if (M.kind == 2) {
        T t[NX];
        user_step<T>(x, u, t);  // xn may alias x
        for (int i = 0; i < M.n; ++i) xn[i] = t[i];
        return;
    }
#endif
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
#ifdef ZB_USER_MODEL
    if (M.kind == 2) {
        user_lin<T>(x, u, fx, fu);
        return;
    }
#endif
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
    const int n = ZB_N_OF(P.M.n), m = ZB_M_OF(P.M.m), N = P.N;
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
#ifdef ZB_USER_COST  // plug-in with a user-defined (symbolic) cost: pytrees.py:40-55 with the generated c(x,u), cf(x)
        if (P.has_cost) J += user_cost<T>(x, u);
#else
        if (P.has_cost) J += quad_form<T>(Q, x, n) + quad_form<T>(R, u, m);
#endif
        model_step<T>(P.M, b, x, u, x);
    }
    if (write)
        for (int i = 0; i < n; ++i) xT[(long long)N * n + i] = x[i];
#ifdef ZB_USER_COST
    if (P.has_cost) J += user_tcost<T>(x);
#else
    if (P.has_cost) J += quad_form<T>(P.C.Qf.at<T>(b), x, n);
#endif
    return J;
}

// Quadcopter specialisation of rollout_core: n = 12, m = 4 known at compile time so the state, the control and the
// (diagonal) cost weights live in registers.  Identical arithmetic order to rollout_core; a dense Q or R falls back
// to the row-by-row quadratic form read through L1.
template <typename T>
ZB_HD T rollout_quad(const RollP& P, long long b, T alpha, bool write) {
    constexpr int n = 12, m = 4;
    const int N = P.N;
    const T* x0 = reinterpret_cast<const T*>(P.x0) + b * n;
    const T* lp = reinterpret_cast<const T*>(P.l) + b * (long long)N * m;
    const T* Lp = reinterpret_cast<const T*>(P.L) + b * (long long)N * m * n;
    const T* xP = reinterpret_cast<const T*>(P.xPrev) + b * (long long)(N + 1) * n;
    const T* uP = reinterpret_cast<const T*>(P.uPrev) + b * (long long)N * m;
    T* xT = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * n;
    T* uT = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * m;
    const T* Q = P.has_cost ? P.C.Q.at<T>(b) : nullptr;
    const T* R = P.has_cost ? P.C.R.at<T>(b) : nullptr;
    const T w[3] = {T(P.M.wind[0]), T(P.M.wind[1]), T(P.M.wind[2])};
    const bool has_wind = P.M.has_wind != 0;
    const T dt = T(P.M.dt);
    T qd[n], rd[m];
    bool diag = true;
    if (P.has_cost) {
#pragma unroll
        for (int i = 0; i < n; ++i)
#pragma unroll
            for (int j = 0; j < n; ++j) {
                if (i == j) qd[i] = Q[i * n + j];
                else diag &= (Q[i * n + j] == T(0));
            }
#pragma unroll
        for (int i = 0; i < m; ++i)
#pragma unroll
            for (int j = 0; j < m; ++j) {
                if (i == j) rd[i] = R[i * m + j];
                else diag &= (R[i * m + j] == T(0));
            }
    }
    T x[n], u[m], xd[n];
    T J = T(0);
#pragma unroll
    for (int i = 0; i < n; ++i) x[i] = x0[i];
    for (int k = 0; k < N; ++k) {
        const T* Lk = Lp + (long long)k * m * n;
#pragma unroll
        for (int i = 0; i < m; ++i) {
            T s = T(0);
#pragma unroll
            for (int j = 0; j < n; ++j) s += Lk[i * n + j] * (x[j] - xP[(long long)k * n + j]);
            u[i] = (alpha * lp[(long long)k * m + i] + s) + uP[(long long)k * m + i];
        }
        if (write) {
#pragma unroll
            for (int i = 0; i < n; ++i) xT[(long long)k * n + i] = x[i];
#pragma unroll
            for (int i = 0; i < m; ++i) uT[(long long)k * m + i] = u[i];
        }
        if (P.has_cost) {
            if (diag) {
                T a = T(0), c = T(0);
#pragma unroll
                for (int i = 0; i < n; ++i) a += x[i] * (qd[i] * x[i]);
#pragma unroll
                for (int i = 0; i < m; ++i) c += u[i] * (rd[i] * u[i]);
                J += a + c;
            } else {
                J += quad_form<T>(Q, x, n) + quad_form<T>(R, u, m);
            }
        }
        QuadTrig<T> tr = quad_trig(x);
        if (has_wind) quad_xdot_wind(tr, x, u, w, xd);
        else quad_xdot(tr, x, u, xd);
#pragma unroll
        for (int i = 0; i < n; ++i) x[i] = x[i] + dt * xd[i];
    }
    if (write) {
#pragma unroll
        for (int i = 0; i < n; ++i) xT[(long long)N * n + i] = x[i];
    }
    if (P.has_cost) J += quad_form<T>(P.C.Qf.at<T>(b), x, n);
    return J;
}

// dispatch: compile-time-sized quadcopter path or the generic run-time-sized one
template <typename T>
ZB_HD T rollout_any(const RollP& P, long long b, T alpha, bool write) {
    if (P.M.kind == 1) return rollout_quad<T>(P, b, alpha, write);
    return rollout_core<T>(P, b, alpha, write);
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
    const int n = ZB_N_OF(P.M.n), m = ZB_M_OF(P.M.m), N = P.N, p = n + m;
    const T* xT = reinterpret_cast<const T*>(P.xTraj) + b * (long long)(N + 1) * n;
    const T* uT = reinterpret_cast<const T*>(P.uTraj) + b * (long long)N * m;
    const T* Czz = reinterpret_cast<const T*>(P.Czz) + b * (long long)p * p;
    const T* Q = P.C.Q.at<T>(b);
    const T* R = P.C.R.at<T>(b);
    const T* Qf = P.C.Qf.at<T>(b);
    T v, v_x[NX], v_xx[NX * NX], l[NU], L[NU * NX];
    T fx[NX * NX], fu[NX * NU], c_x[NX], c_u[NU], c_xx[NX * NX], c_ux[NU * NX], c_uu[NU * NU];
    T vf_xx[NX * NX], vf_ux[NU * NX], vf_uu[NU * NU];
#ifndef ZB_USER_COST
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) c_xx[i * n + j] = Czz[i * p + j];
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j) c_ux[i * n + j] = Czz[(n + i) * p + j];
        for (int j = 0; j < m; ++j) c_uu[i * m + j] = Czz[(n + i) * p + n + j];
    }
#endif
    const T* xN = xT + (long long)N * n;
#ifdef ZB_USER_COST
    // user-defined terminal cost: value, gradient and eigen-clamped Hessian at x_N (pytrees.py:71-81, ilqrUtils.py:254-257)
    {
        T Wc[ZB_PD_MAX * ZB_PD_MAX];
        v = user_tcost<T>(xN);
        user_tcost_grad<T>(xN, v_x);
        user_tcost_hess<T>(xN, v_xx);
        pd_clamp<T>(v_xx, Wc, n, T(P.eps));
    }
    (void)Qf; (void)Q; (void)R; (void)Czz;
#else
    v = quad_form<T>(Qf, xN, n);
    for (int i = 0; i < n; ++i) {
        T s = T(0);
        for (int j = 0; j < n; ++j) s += (Qf[i * n + j] + Qf[j * n + i]) * xN[j];
        v_x[i] = s;
    }
    const T* Vf = reinterpret_cast<const T*>(P.Vfxx) + b * (long long)n * n;
    for (int i = 0; i < n * n; ++i) v_xx[i] = Vf[i];
#endif
    T* lo = reinterpret_cast<T*>(P.l) + b * (long long)N * m;
    T* Lo = reinterpret_cast<T*>(P.L) + b * (long long)N * m * n;
    for (int k = N - 1; k >= 0; --k) {
        const T* x = xT + (long long)k * n;
        const T* u = uT + (long long)k * m;
        model_lin<T>(P.M, b, x, u, fx, fu);
#ifdef ZB_USER_COST
        // user-defined running cost: Taylor data at (x_k, u_k) (pytrees.py:99-115) and the per-step eigen-clamp of the stacked
        // Hessian [[c_xx, c_ux'], [c_ux, c_uu]] (conditionQuadraticCost, ilqrUtils.py:222-234) -- not hoistable for a general cost
        T c;
        {
            T Zc[ZB_PD_MAX * ZB_PD_MAX], Wc[ZB_PD_MAX * ZB_PD_MAX];
            c = user_cost<T>(x, u);
            user_cost_grad<T>(x, u, c_x, c_u);
            user_cost_hess<T>(x, u, Zc);
            pd_clamp<T>(Zc, Wc, p, T(P.eps));
            for (int i = 0; i < n; ++i)
                for (int j = 0; j < n; ++j) c_xx[i * n + j] = Zc[i * p + j];
            for (int i = 0; i < m; ++i) {
                for (int j = 0; j < n; ++j) c_ux[i * n + j] = Zc[(n + i) * p + j];
                for (int j = 0; j < m; ++j) c_uu[i * m + j] = Zc[(n + i) * p + n + j];
            }
        }
#else
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
#endif
        if (P.second_order) {
            // conditionQuadraticDynamics (ilqrUtils.py:237-251); f_ux = f_uu = 0 for both registered models, and for
            // the quadcopter v_x.f_xx only touches states 0..8, so the (n+m)x(n+m) block is blockdiag(H9, 0): its
            // eigen-clamp is blockdiag(clampPD(H9), eps*I) exactly (block-diagonal matrices have block eigenvectors;
            // the result is a spectral function, independent of the basis chosen inside the null space).
            for (int i = 0; i < n * n; ++i) vf_xx[i] = T(0);
            for (int i = 0; i < m * n; ++i) vf_ux[i] = T(0);
            for (int i = 0; i < m * m; ++i) vf_uu[i] = T(0);
            int q = 0;  // size of the non-trivial leading block
            T Z[ZB_PD_MAX * ZB_PD_MAX], W[ZB_PD_MAX * ZB_PD_MAX];
#ifdef ZB_USER_MODEL
            if (P.M.kind == 2) {
                // general model: the full (n+m) x (n+m) block [[v_x.f_xx, (v_x.f_ux)'], [v_x.f_ux, v_x.f_uu]] (ilqrUtils.py:244-249)
                user_hess<T>(x, u, v_x, Z);
                pd_clamp<T>(Z, W, p, T(P.eps));
                for (int i = 0; i < n; ++i)
                    for (int j = 0; j < n; ++j) vf_xx[i * n + j] = Z[i * p + j];
                for (int i = 0; i < m; ++i) {
                    for (int j = 0; j < n; ++j) vf_ux[i * n + j] = Z[(n + i) * p + j];
                    for (int j = 0; j < m; ++j) vf_uu[i * m + j] = Z[(n + i) * p + n + j];
                }
                q = -1;  // all blocks are set
            }
#endif
            if (P.M.kind == 1) {
                T w[3] = {T(P.M.wind[0]), T(P.M.wind[1]), T(P.M.wind[2])};
                T H[144];
                quad_hess<T>(x, u, w, P.M.has_wind != 0, T(P.M.dt), v_x, H);
                q = 9;
                for (int i = 0; i < q; ++i)
                    for (int j = 0; j < q; ++j) Z[i * q + j] = H[i * n + j];
                pd_clamp<T>(Z, W, q, T(P.eps));
                for (int i = 0; i < q; ++i)
                    for (int j = 0; j < q; ++j) vf_xx[i * n + j] = Z[i * q + j];
            }
            if (q >= 0) {
                for (int i = q; i < n; ++i) vf_xx[i * n + i] = T(P.eps);
                for (int i = 0; i < m; ++i) vf_uu[i * m + i] = T(P.eps);
            }
        }
        ilqr_step<T>(n, m, fx, fu, c, c_x, c_u, c_xx, c_ux, c_uu, P.second_order ? vf_xx : nullptr, vf_ux, vf_uu, v,
                     v_x, v_xx, l, L);
        for (int i = 0; i < m; ++i) lo[(long long)k * m + i] = l[i];
        for (int i = 0; i < m * n; ++i) Lo[(long long)k * m * n + i] = L[i];
    }
}


// -------------------------------------------------------------------------------------------------
// Box-constrained lqrMpc (zopt/mpcUtils.py:47-59) by ADMM.  The reference hands this QP to cvxpy ->
// OSQP (third-party, not in tree); here the same splitting idea (ADMM with over-relaxation, OSQP's
// default alpha = 1.6, residual-balancing rho) is applied with the DYNAMICS kept exact inside the
// linear solve: each iteration minimises  sum x'Qx + u'Ru + (rho/2)||z - w + lam/rho||^2  subject to
// x+ = Ax + Bu, x_0 = x0 -- an LQ problem whose gains depend only on rho (one Riccati sweep per rho)
// and whose affine part is one backward vector sweep + one forward rollout per iteration -- then
// projects onto the box and updates the multipliers.
struct AdmmP {
    long long Bsz;
    int N, n, m;
    Arr A, B, Q, R, Qf, xlb, xub, ulb, uub;
    const void* x0;
    void *u0, *xTraj, *uTraj;
    int8_t* status;
    int32_t* iters;
    void* ws;
    long long ws_stride;  // elements of T per problem
    int max_iter, check_every;
    double rho, alpha, eps_abs, eps_rel, eps_inf;
};

ZB_HD long long admm_ws_elems(int N, int n, int m) {
    return (long long)N * m * n + (long long)N * m * m + (long long)N * m + 3LL * ((long long)(N + 1) * n + (long long)N * m);
}

// gains K_k = G^-1 B'PA and G_k^-1 for the Hessian-form LQ problem with weights (2Q + rho I, 2R + rho I, 2Qf + rho I)
template <typename T>
ZB_HD void admm_factor(int N, int n, int m, const T* A, const T* B, const T* Q, const T* R, const T* Qf, T rho, T* K, T* Gi) {
    T P[NX * NX], BtP[NU * NX], G[NU * NU], Ginv[NU * NU], Kk[NU * NX], BK[NX * NX], W[NX * NX];
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) P[i * n + j] = T(2) * Qf[i * n + j] + (i == j ? rho : T(0));
    for (int k = N - 1; k >= 0; --k) {
        mm_tn(BtP, B, P, m, n, n);
        mm(G, BtP, B, m, n, m);
        for (int i = 0; i < m; ++i)
            for (int j = 0; j < m; ++j) {
                G[i * m + j] += T(2) * R[i * m + j] + (i == j ? rho : T(0));
                Ginv[i * m + j] = (i == j) ? T(1) : T(0);
            }
        lu_solve(G, m, Ginv, m);
        mm(W, BtP, A, m, n, n);     // B'PA (m x n), reuse W storage
        mm(Kk, Ginv, W, m, m, n);   // K
        for (int i = 0; i < m * n; ++i) K[(long long)k * m * n + i] = Kk[i];
        for (int i = 0; i < m * m; ++i) Gi[(long long)k * m * m + i] = Ginv[i];
        mm(BK, B, Kk, n, m, n);
        for (int i = 0; i < n * n; ++i) BK[i] = A[i] - BK[i];  // Acl
        mm(W, P, BK, n, n, n);                                  // P Acl
        mm_tn(P, A, W, n, n, n);                                // A' P Acl
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) P[i * n + j] += T(2) * Q[i * n + j] + (i == j ? rho : T(0));
        for (int i = 0; i < n; ++i)  // keep P exactly symmetric
            for (int j = i + 1; j < n; ++j) {
                T a = T(0.5) * (P[i * n + j] + P[j * n + i]);
                P[i * n + j] = a;
                P[j * n + i] = a;
            }
    }
}

template <typename T>
ZB_HD T clampT(T v, T lo, T hi) { return v < lo ? lo : (v > hi ? hi : v); }

template <typename T>
ZB_HD void admm_problem(const AdmmP& P, long long b) {
    const int N = P.N, n = P.n, m = P.m;
    const T *A = P.A.at<T>(b), *B = P.B.at<T>(b), *Q = P.Q.at<T>(b), *R = P.R.at<T>(b), *Qf = P.Qf.at<T>(b);
    const T *xlb = P.xlb.at<T>(b), *xub = P.xub.at<T>(b), *ulb = P.ulb.at<T>(b), *uub = P.uub.at<T>(b);
    const T* x0 = reinterpret_cast<const T*>(P.x0) + b * n;
    T* zx = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * n;
    T* zu = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * m;
    T* ws = reinterpret_cast<T*>(P.ws) + b * P.ws_stride;
    const long long nx = (long long)(N + 1) * n, nu = (long long)N * m;
    T* K = ws;
    T* Gi = K + (long long)N * m * n;
    T* kff = Gi + (long long)N * m * m;
    T* wx = kff + nu;
    T* wu = wx + nx;
    T* lx = wu + nu;
    T* lu = lx + nx;
    T* dx = lu + nu;  // delta-lambda (infeasibility certificate)
    T* du = dx + nx;
    const T INF = T(1) / T(0);
    T* u0 = reinterpret_cast<T*>(P.u0) + b * m;

    // x_0 = x0 is itself box-constrained in the reference QP (mpcUtils.py:56,58): outside the box = infeasible
    bool x0_bad = false;
    for (int i = 0; i < n; ++i) x0_bad |= !(x0[i] >= xlb[i] - T(P.eps_abs) && x0[i] <= xub[i] + T(P.eps_abs));
    if (x0_bad) {
        const T nan = INF - INF;
        for (long long i = 0; i < nx; ++i) zx[i] = nan;
        for (long long i = 0; i < nu; ++i) zu[i] = nan;
        for (int i = 0; i < m; ++i) u0[i] = nan;
        P.status[b] = 2;
        if (P.iters) P.iters[b] = 0;
        return;
    }
    T rho = T(P.rho);
    const T alpha = T(P.alpha);
    admm_factor<T>(N, n, m, A, B, Q, R, Qf, rho, K, Gi);
    for (long long i = 0; i < nx; ++i) { wx[i] = T(0); lx[i] = T(0); }
    for (long long i = 0; i < nu; ++i) { wu[i] = T(0); lu[i] = T(0); }
    int status = 1, it = 0;
    T p[NX], h[NU], x[NX], u[NU], t1[NX];
    for (it = 1; it <= P.max_iter; ++it) {
        // ---- z-update: backward vector sweep (linear terms -rho*(w - lam/rho) = lam - rho*w), forward rollout
        for (int i = 0; i < n; ++i) p[i] = lx[(long long)N * n + i] - rho * wx[(long long)N * n + i];
        for (int k = N - 1; k >= 0; --k) {
            const T* Kk = K + (long long)k * m * n;
            mv_t(h, B, p, m, n);
            for (int i = 0; i < m; ++i) h[i] += lu[(long long)k * m + i] - rho * wu[(long long)k * m + i];
            mv(kff + (long long)k * m, Gi + (long long)k * m * m, h, m, m);
            mv_t(t1, A, p, n, n);
            for (int i = 0; i < n; ++i) {
                T s = T(0);
                for (int a = 0; a < m; ++a) s += Kk[a * n + i] * h[a];
                p[i] = (lx[(long long)k * n + i] - rho * wx[(long long)k * n + i]) + t1[i] - s;
            }
        }
        for (int i = 0; i < n; ++i) x[i] = x0[i];
        for (int k = 0; k < N; ++k) {
            const T* Kk = K + (long long)k * m * n;
            for (int a = 0; a < m; ++a) {
                T s = kff[(long long)k * m + a];
                for (int j = 0; j < n; ++j) s += Kk[a * n + j] * x[j];
                u[a] = -s;
            }
            for (int i = 0; i < n; ++i) zx[(long long)k * n + i] = x[i];
            for (int a = 0; a < m; ++a) zu[(long long)k * m + a] = u[a];
            for (int i = 0; i < n; ++i) {
                T s = T(0);
                for (int j = 0; j < n; ++j) s += A[i * n + j] * x[j];
                for (int a = 0; a < m; ++a) s += B[i * m + a] * u[a];
                t1[i] = s;
            }
            for (int i = 0; i < n; ++i) x[i] = t1[i];
        }
        for (int i = 0; i < n; ++i) zx[(long long)N * n + i] = x[i];
        // ---- relaxation, projection on the box, multiplier update, residuals
        const bool chk = (it % P.check_every == 0) || it == P.max_iter;
        T rp = T(0), rd = T(0), nz = T(0), nw = T(0), nl = T(0), ndl = T(0), S = T(0);
        bool cert_ok = true;
        for (int part = 0; part < 2; ++part) {
            const long long cnt = part ? nu : nx;
            const int dim = part ? m : n;
            T* z = part ? zu : zx;
            T* w = part ? wu : wx;
            T* l = part ? lu : lx;
            T* d = part ? du : dx;
            const T* lb = part ? ulb : xlb;
            const T* ub = part ? uub : xub;
            for (long long i = 0; i < cnt; ++i) {
                const int c = (int)(i % dim);
                const T zi = z[i], wo = w[i];
                const T zr = alpha * zi + (T(1) - alpha) * wo;
                const T wn = clampT<T>(zr + l[i] / rho, lb[c], ub[c]);
                const T dl = rho * (zr - wn);
                l[i] += dl;
                w[i] = wn;
                rp = fmax(rp, fabs(zi - wn));
                rd = fmax(rd, rho * fabs(wn - wo));
                nz = fmax(nz, fabs(zi));
                nw = fmax(nw, fabs(wn));
                nl = fmax(nl, fabs(l[i]));
                if (chk) {
                    d[i] = dl;
                    ndl = fmax(ndl, fabs(dl));
                    if (dl > T(0)) { if (ub[c] == INF) cert_ok = false; else S += ub[c] * dl; }
                    else if (dl < T(0)) { if (lb[c] == -INF) cert_ok = false; else S += lb[c] * dl; }
                    S -= dl * zi;
                }
            }
        }
        if (chk) {
            if (rp <= T(P.eps_abs) + T(P.eps_rel) * fmax(nz, nw) && rd <= T(P.eps_abs) + T(P.eps_rel) * nl) {
                status = 0;
                break;
            }
            // primal infeasibility certificate: delta-lambda separates the box from the dynamics' affine set
            if (cert_ok && ndl > T(P.eps_inf) && S < -T(P.eps_inf) * ndl) {
                T mu[NX], g = T(0);
                for (int i = 0; i < n; ++i) mu[i] = dx[(long long)N * n + i];
                for (int k = N - 1; k >= 0; --k) {
                    mv_t(h, B, mu, m, n);
                    for (int a = 0; a < m; ++a) g = fmax(g, fabs(h[a] + du[(long long)k * m + a]));
                    mv_t(t1, A, mu, n, n);
                    for (int i = 0; i < n; ++i) mu[i] = dx[(long long)k * n + i] + t1[i];
                }
                if (g <= T(P.eps_inf) * ndl) {
                    status = 2;
                    break;
                }
            }
            // residual balancing (OSQP adapts rho the same way): refactor when the residuals are 5x apart
            if (it < P.max_iter) {
                const T rpn = rp / fmax(fmax(nz, nw), T(1e-10)), rdn = rd / fmax(nl, T(1e-10));
                const T ratio = sqrt(rpn / fmax(rdn, T(1e-30)));
                if ((ratio > T(5) || ratio < T(0.2)) && rdn > T(0)) {
                    T rn = clampT<T>(rho * ratio, T(1e-6), T(1e6));
                    if (rn != rho) {
                        rho = rn;
                        admm_factor<T>(N, n, m, A, B, Q, R, Qf, rho, K, Gi);
                    }
                }
            }
        }
    }
    if (it > P.max_iter) it = P.max_iter;
    if (status == 2) {
        const T nan = INF - INF;
        for (long long i = 0; i < nx; ++i) zx[i] = nan;
        for (long long i = 0; i < nu; ++i) zu[i] = nan;
    }
    for (int i = 0; i < m; ++i) u0[i] = zu[i];
    P.status[b] = (int8_t)status;
    if (P.iters) P.iters[b] = it;
}

}  // namespace zb
