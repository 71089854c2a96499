// One-problem building blocks of the hot path with run-time (n, m): the Riccati steps of
// lqrUtils / ilqrUtils, the quadcopter Euler model and the quadratic cost.  Used by the generic
// one-thread-per-problem kernels (any n <= ZB_MAX_N, m <= ZB_MAX_M); the (12,4) fast kernels
// implement the same recursions cooperatively and are tested against these through the oracle.
#pragma once

#include "zb_math.cuh"
#include "quad_model_gen.cuh"

#ifndef ZB_MAX_N
#define ZB_MAX_N 16
#endif
#ifndef ZB_MAX_M
#define ZB_MAX_M 8
#endif

// A user-model plug-in (zb_user_model.cu) is compiled for ONE model: its (n, m) are compile-time constants there, so the
// per-problem arrays are exactly as large as the model needs and every loop over states / controls has a constant trip
// count (unrolled, arrays promoted to registers) -- the library's generic kernels size everything for (16, 8) in local memory.
#ifdef ZB_USER_MODEL
#define ZB_N_OF(x) (ZB_USER_N)
#define ZB_M_OF(x) (ZB_USER_M)
#else
#define ZB_N_OF(x) (x)
#define ZB_M_OF(x) (x)
#endif

namespace zb {

#ifdef ZB_USER_MODEL
constexpr int NX = ZB_USER_N;
constexpr int NU = ZB_USER_M;
#else
constexpr int NX = ZB_MAX_N;
constexpr int NU = ZB_MAX_M;
#endif

// ---- zopt/lqrUtils.py:167-170  riccatiStep (Joseph form, as written) -------------------------
// V (n x n) in/out; L (m x n) out.  A,B,Q,R may live in global memory.
template <typename T>
ZB_HD void lqr_joseph_step(int n, int m, const T* A, const T* B, const T* Q, const T* R, T* V, T* L) {
    T BtV[NU * NX], G[NU * NU], Acl[NX * NX], W[NX * NX], RL[NU * NX];
    mm_tn(BtV, B, V, m, n, n);    // B^T V            (m x n)
    mm(G, BtV, B, m, n, m);       // B^T V B          (m x m)
    for (int i = 0; i < m * m; ++i) G[i] += R[i];
    mm(L, BtV, A, m, n, n);       // B^T V A          (m x n)
    lu_solve(G, m, L, n);         // L = (R + B^T V B)^-1 B^T V A
    mm(Acl, B, L, n, m, n);       // B L
    for (int i = 0; i < n * n; ++i) Acl[i] = A[i] - Acl[i];
    mm(W, V, Acl, n, n, n);       // V Acl
    mm_tn(V, Acl, W, n, n, n);    // Acl^T V Acl
    mm(RL, R, L, m, m, n);        // R L
    mm_tn(W, L, RL, n, m, n);     // L^T R L
    for (int i = 0; i < n * n; ++i) V[i] += Q[i] + W[i];
}

// ---- zopt/lqrUtils.py:242-259  bilinearRiccatiStep ------------------------------------------------
template <typename T>
ZB_HD void bilinear_step(int n, int m, const T* A, const T* B, const T* d, const T* Q, const T* R, const T* H,
                         const T* q, const T* r, T q0, T* V, T* v, T& v0, T* L, T* l) {
    T Vd[NX], vVd[NX], BtV[NU * NX], Su[NU], Suu[NU * NU], Suu_f[NU * NU], Sux[NU * NX], W[NX * NX], rhs[NU * (NX + 1)];
    mv(Vd, V, d, n, n);
    for (int i = 0; i < n; ++i) vVd[i] = v[i] + Vd[i];
    mv_t(Su, B, vVd, m, n);  // B^T (v + V d)  ==  v^T B + d^T V B
    for (int i = 0; i < m; ++i) Su[i] += r[i];
    mm_tn(BtV, B, V, m, n, n);
    mm(Suu, BtV, B, m, n, m);
    for (int i = 0; i < m * m; ++i) Suu[i] += R[i];
    mm(Sux, BtV, A, m, n, n);
    for (int i = 0; i < m * n; ++i) Sux[i] += H[i];
    // [L | l] = Suu^-1 [Sux | Su]
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j) rhs[i * (n + 1) + j] = Sux[i * n + j];
        rhs[i * (n + 1) + n] = Su[i];
    }
    for (int i = 0; i < m * m; ++i) Suu_f[i] = Suu[i];
    lu_solve(Suu_f, m, rhs, n + 1);
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j) L[i * n + j] = rhs[i * (n + 1) + j];
        l[i] = rhs[i * (n + 1) + n];
    }
    // v0New = v0 + q0 + d^T v + 0.5 d^T V d - 0.5 l^T Su
    v0 = v0 + q0 + dot<T>(d, v, n) + T(0.5) * dot<T>(d, Vd, n) - T(0.5) * dot<T>(l, Su, m);
    // vNew = q + A^T (v + V d) - Sux^T l
    T t1[NX], t2[NX];
    mv_t(t1, A, vVd, n, n);
    mv_t(t2, Sux, l, n, m);
    for (int i = 0; i < n; ++i) v[i] = q[i] + t1[i] - t2[i];
    // VNew = Q + A^T V A - L^T Suu L
    T SL[NU * NX];
    mm(W, V, A, n, n, n);
    mm_tn(V, A, W, n, n, n);
    mm(SL, Suu, L, m, m, n);
    mm_tn(W, L, SL, n, m, n);
    for (int i = 0; i < n * n; ++i) V[i] += Q[i] - W[i];
}

// -------------------------------------------------------------------------------------------------------------
// The two LQR steps above with COMPILE-TIME (n, m): every loop unrolls and V, W, A, B ... live in registers instead of
// local memory (the run-time-sized bodies index per-thread arrays dynamically, so all of their operands sit in local
// memory: ~1-2 M solves/s).  Used for the shapes of the reference's demos and tests, (8,4) and (2,1)/(2,2)
// (demos/discreteFiniteHorizonLqr.py:29-35, demos/bilinearLqrControl.py:21-43, tests/test_lqrUtils.py:61-98).
// Same formulas in the same order; the m x m system is solved by Gaussian elimination WITHOUT pivoting (S_uu = R + B'VB is
// symmetric positive definite), which differs from LAPACK's pivoted LU at rounding level only.
#if defined(__CUDACC__)
#define ZB_UNROLL _Pragma("unroll")
#else
#define ZB_UNROLL
#endif

template <typename T, int M_, int NR>
ZB_HD void spd_solve_ct(T* G, T* X) {  // G (M_ x M_) SPD, X (M_ x NR) right-hand sides -> G^-1 X
    ZB_UNROLL
    for (int c = 0; c < M_; ++c) {
        const T inv = T(1) / G[c * M_ + c];
        ZB_UNROLL
        for (int i = c + 1; i < M_; ++i) {
            const T f = G[i * M_ + c] * inv;
            ZB_UNROLL
            for (int j = c + 1; j < M_; ++j) G[i * M_ + j] -= f * G[c * M_ + j];
            ZB_UNROLL
            for (int j = 0; j < NR; ++j) X[i * NR + j] -= f * X[c * NR + j];
        }
    }
    ZB_UNROLL
    for (int i = M_ - 1; i >= 0; --i) {
        const T inv = T(1) / G[i * M_ + i];
        ZB_UNROLL
        for (int j = 0; j < NR; ++j) {
            T s = X[i * NR + j];
            ZB_UNROLL
            for (int k = i + 1; k < M_; ++k) s -= G[i * M_ + k] * X[k * NR + j];
            X[i * NR + j] = s * inv;
        }
    }
}

// C (P_ x R_) = A (P_ x Q_) B (Q_ x R_)   /   C = A' B with A stored (Q_ x P_)
template <typename T, int P_, int Q_, int R_, bool TA>
ZB_HD void mm_ct(T* C, const T* A, const T* B) {
    ZB_UNROLL
    for (int i = 0; i < P_; ++i)
        ZB_UNROLL
        for (int j = 0; j < R_; ++j) {
            T s = T(0);
            ZB_UNROLL
            for (int k = 0; k < Q_; ++k) s += (TA ? A[k * P_ + i] : A[i * Q_ + k]) * B[k * R_ + j];
            C[i * R_ + j] = s;
        }
}

// zopt/lqrUtils.py:167-170 (Joseph form, as written).  A, B, Q, R: global pointers (copied to registers here).
template <typename T, int N_, int M_>
ZB_HD void lqr_joseph_step_ct(const T* gA, const T* gB, const T* gQ, const T* gR, T* V, T* L) {
    T A[N_ * N_], B[N_ * M_], R[M_ * M_], BtV[M_ * N_], G[M_ * M_], W[N_ * N_];
    ZB_UNROLL
    for (int i = 0; i < N_ * N_; ++i) A[i] = gA[i];
    ZB_UNROLL
    for (int i = 0; i < N_ * M_; ++i) B[i] = gB[i];
    ZB_UNROLL
    for (int i = 0; i < M_ * M_; ++i) R[i] = gR[i];
    mm_ct<T, M_, N_, N_, true>(BtV, B, V);   // B'V
    mm_ct<T, M_, N_, M_, false>(G, BtV, B);  // B'VB
    ZB_UNROLL
    for (int i = 0; i < M_ * M_; ++i) G[i] += R[i];
    mm_ct<T, M_, N_, N_, false>(L, BtV, A);  // B'VA
    spd_solve_ct<T, M_, N_>(G, L);           // L = (R + B'VB)^-1 B'VA
    mm_ct<T, N_, M_, N_, false>(W, B, L);    // B L
    ZB_UNROLL
    for (int i = 0; i < N_ * N_; ++i) A[i] -= W[i];  // Acl = A - B L
    mm_ct<T, N_, N_, N_, false>(W, V, A);    // V Acl
    mm_ct<T, N_, N_, N_, true>(V, A, W);     // Acl' V Acl
    mm_ct<T, M_, M_, N_, false>(BtV, R, L);  // R L
    mm_ct<T, N_, M_, N_, true>(W, L, BtV);   // L' R L
    ZB_UNROLL
    for (int i = 0; i < N_ * N_; ++i) V[i] += gQ[i] + W[i];
}

// zopt/lqrUtils.py:242-259 (the scalar carry v0 does not feed (L, l) and is not formed)
template <typename T, int N_, int M_>
ZB_HD void bilinear_step_ct(const T* gA, const T* gB, const T* d, const T* gQ, const T* gR, const T* gH, const T* q, const T* r,
                            T* V, T* v, T* L, T* l) {
    T A[N_ * N_], B[N_ * M_], vVd[N_], BtV[M_ * N_], Suu[M_ * M_], Sf[M_ * M_], Sux[M_ * N_], rhs[M_ * (N_ + 1)], W[N_ * N_];
    ZB_UNROLL
    for (int i = 0; i < N_ * N_; ++i) A[i] = gA[i];
    ZB_UNROLL
    for (int i = 0; i < N_ * M_; ++i) B[i] = gB[i];
    ZB_UNROLL
    for (int i = 0; i < N_; ++i) {  // v + V d
        T s = T(0);
        ZB_UNROLL
        for (int k = 0; k < N_; ++k) s += V[i * N_ + k] * d[k];
        vVd[i] = v[i] + s;
    }
    mm_ct<T, M_, N_, N_, true>(BtV, B, V);
    mm_ct<T, M_, N_, M_, false>(Suu, BtV, B);
    mm_ct<T, M_, N_, N_, false>(Sux, BtV, A);
    ZB_UNROLL
    for (int i = 0; i < M_ * M_; ++i) { Suu[i] += gR[i]; Sf[i] = Suu[i]; }
    ZB_UNROLL
    for (int i = 0; i < M_; ++i) {  // [Sux | Su], Su = r + B'(v + V d)
        T s = T(0);
        ZB_UNROLL
        for (int k = 0; k < N_; ++k) s += B[k * M_ + i] * vVd[k];
        ZB_UNROLL
        for (int j = 0; j < N_; ++j) { Sux[i * N_ + j] += gH[i * N_ + j]; rhs[i * (N_ + 1) + j] = Sux[i * N_ + j]; }
        rhs[i * (N_ + 1) + N_] = s + r[i];
    }
    spd_solve_ct<T, M_, N_ + 1>(Sf, rhs);
    ZB_UNROLL
    for (int i = 0; i < M_; ++i) {
        ZB_UNROLL
        for (int j = 0; j < N_; ++j) L[i * N_ + j] = rhs[i * (N_ + 1) + j];
        l[i] = rhs[i * (N_ + 1) + N_];
    }
    ZB_UNROLL
    for (int i = 0; i < N_; ++i) {  // vNew = q + A'(v + V d) - Sux' l
        T s = q[i];
        ZB_UNROLL
        for (int k = 0; k < N_; ++k) s += A[k * N_ + i] * vVd[k];
        ZB_UNROLL
        for (int a = 0; a < M_; ++a) s -= Sux[a * N_ + i] * l[a];
        v[i] = s;
    }
    mm_ct<T, N_, N_, N_, false>(W, V, A);    // V A
    mm_ct<T, N_, N_, N_, true>(V, A, W);     // A' V A
    mm_ct<T, M_, M_, N_, false>(BtV, Suu, L);  // Suu L
    mm_ct<T, N_, M_, N_, true>(W, L, BtV);     // L' Suu L
    ZB_UNROLL
    for (int i = 0; i < N_ * N_; ++i) V[i] += gQ[i] - W[i];
}

// ---- zopt/ilqrUtils.py:153-173 riccatiStep_ilqr  /  :184-206 riccatiStep_ddp -----------------------
// vf_* (optional, may be null): the eigen-clamped v_x.f_zz blocks of the DDP step.
// value (v, v_x, v_xx) in/out; policy (l (m), L (m x n)) out.
template <typename T>
ZB_HD void ilqr_step(int n, int m, const T* f_x, const T* f_u, T c, const T* c_x, const T* c_u, const T* c_xx,
                     const T* c_ux, const T* c_uu, const T* vf_xx, const T* vf_ux, const T* vf_uu, T& v, T* v_x,
                     T* v_xx, T* l, T* L) {
    T W[NX * NX], VB[NX * NU], Qxx[NX * NX], Quu[NU * NU], Quu_f[NU * NU], Qux[NU * NX], Qx[NX], Qu[NU];
    T rhs[NU * (NX + 1)];
    mm(W, v_xx, f_x, n, n, n);        // v_xx f_x
    mm_tn(Qxx, f_x, W, n, n, n);      // f_x^T v_xx f_x
    mm_tn(Qux, f_u, W, m, n, n);      // f_u^T v_xx f_x
    mm(VB, v_xx, f_u, n, n, m);
    mm_tn(Quu, f_u, VB, m, n, m);     // f_u^T v_xx f_u
    for (int i = 0; i < n * n; ++i) Qxx[i] += c_xx[i];
    for (int i = 0; i < m * n; ++i) Qux[i] += c_ux[i];
    for (int i = 0; i < m * m; ++i) Quu[i] += c_uu[i];
    if (vf_xx) {
        for (int i = 0; i < n * n; ++i) Qxx[i] += vf_xx[i];
        for (int i = 0; i < m * n; ++i) Qux[i] += vf_ux[i];
        for (int i = 0; i < m * m; ++i) Quu[i] += vf_uu[i];
    }
    mv_t(Qx, f_x, v_x, n, n);
    mv_t(Qu, f_u, v_x, m, n);
    for (int i = 0; i < n; ++i) Qx[i] += c_x[i];
    for (int i = 0; i < m; ++i) Qu[i] += c_u[i];
    // l = -Quu^-1 Qu ; L = -Quu^-1 Qux
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j) rhs[i * (n + 1) + j] = Qux[i * n + j];
        rhs[i * (n + 1) + n] = Qu[i];
    }
    for (int i = 0; i < m * m; ++i) Quu_f[i] = Quu[i];
    lu_solve(Quu_f, m, rhs, n + 1);
    for (int i = 0; i < m; ++i) {
        for (int j = 0; j < n; ++j) L[i * n + j] = -rhs[i * (n + 1) + j];
        l[i] = -rhs[i * (n + 1) + n];
    }
    // valueOut = (Q - 0.5 l^T Quu l, Q_x - L^T Quu l, Q_xx - L^T Quu L)
    T Ql[NU], QL[NU * NX];
    mv(Ql, Quu, l, m, m);
    mm(QL, Quu, L, m, m, n);
    v = (c + v) - T(0.5) * dot<T>(l, Ql, m);
    T t[NX];
    mv_t(t, L, Ql, n, m);
    for (int i = 0; i < n; ++i) v_x[i] = Qx[i] - t[i];
    mm_tn(W, L, QL, n, m, n);
    for (int i = 0; i < n * n; ++i) v_xx[i] = Qxx[i] - W[i];
}

// ---- quadcopter (zopt/quadcopter.py:116-144), forward Euler x + dt*F (demos/iterativeLqr.py:35) --
template <typename T>
ZB_HD QuadTrig<T> quad_trig(const T* x) {
    QuadTrig<T> tr;
#ifdef __CUDA_ARCH__
    sincos(x[6], &tr.sph, &tr.cph);  // overloads: sincos(double,..) / sincos(float,..) -> sincosf
    sincos(x[7], &tr.sth, &tr.cth);
    sincos(x[8], &tr.sps, &tr.cps);
#else
    tr.sph = sin(x[6]); tr.cph = cos(x[6]);
    tr.sth = sin(x[7]); tr.cth = cos(x[7]);
    tr.sps = sin(x[8]); tr.cps = cos(x[8]);
#endif
    tr.sec = T(1) / tr.cth;
    tr.tth = tr.sth * tr.sec;
    return tr;
}

template <typename T>
ZB_HD void quad_F(const T* x, const T* u, const T* wind, bool has_wind, T* xd) {
    QuadTrig<T> tr = quad_trig(x);
    if (has_wind) quad_xdot_wind(tr, x, u, wind, xd);
    else quad_xdot(tr, x, u, xd);
}

// x_next = x + dt * F(x,u)   (x_next may alias x)
template <typename T>
ZB_HD void quad_euler(const T* x, const T* u, const T* wind, bool has_wind, T dt, T* xn) {
    T xd[12];
    quad_F(x, u, wind, has_wind, xd);
    for (int i = 0; i < 12; ++i) xn[i] = x[i] + dt * xd[i];
}

// f_x = I*(dt!=0) + s*dF/dx (12x12), f_u = s*dF/du (12x4) with s = dt ? dt : 1
template <typename T>
ZB_HD void quad_lin(const T* x, const T* u, const T* wind, bool has_wind, T dt, T* fx, T* fu) {
    QuadTrig<T> tr = quad_trig(x);
    if (has_wind) quad_jac_x_wind(tr, x, u, wind, fx);
    else quad_jac_x(tr, x, u, fx);
    const T s = (dt != T(0)) ? dt : T(1);
    for (int i = 0; i < 144; ++i) fx[i] *= s;
    if (dt != T(0))
        for (int i = 0; i < 12; ++i) fx[i * 13] += T(1);
    if (fu) {
        for (int i = 0; i < 48; ++i) fu[i] = T(0);
        fu[2 * 4 + 0] = -s;
        fu[3 * 4 + 1] = s;
        fu[4 * 4 + 2] = s;
        fu[5 * 4 + 3] = s;
    }
}

// H (12x12, full symmetric) = s * sum_i lam_i d2F_i/dx2
template <typename T>
ZB_HD void quad_hess(const T* x, const T* u, const T* wind, bool has_wind, T dt, const T* lam, T* H) {
    QuadTrig<T> tr = quad_trig(x);
    T h9[81];
    if (has_wind) quad_hess_contract_wind(tr, x, u, wind, lam, h9);
    else quad_hess_contract(tr, x, u, lam, h9);
    const T s = (dt != T(0)) ? dt : T(1);
    for (int i = 0; i < 144; ++i) H[i] = T(0);
    for (int i = 0; i < 9; ++i)
        for (int j = 0; j <= i; ++j) {
            H[i * 12 + j] = s * h9[i * 9 + j];
            H[j * 12 + i] = s * h9[i * 9 + j];
        }
}

// running cost x'Qx + u'Ru and terminal x'Qf x (demos/iterativeLqr.py:12-13,36-37)
template <typename T>
ZB_HD T quad_form(const T* M, const T* x, int n) {
    T s = T(0);
    for (int i = 0; i < n; ++i) {
        T r = T(0);
        for (int j = 0; j < n; ++j) r += M[i * n + j] * x[j];
        s += x[i] * r;
    }
    return s;
}

}  // namespace zb
