/*
 * TEST INFRASTRUCTURE -- not part of the product.  Plain-C (C99 + OpenMP) restatement of the reference's algorithm for
 * the headline path, fp64 like the reference (zopt/quadcopter.py:7 enables x64):
 *
 *   zo_quad_inertial_dynamics   zopt/quadcopter.py:116-144 (inertialDynamics) -> :70-113 (rigidBodyDynamics),
 *                               rotation matrices :23-48, aero :51-67
 *   zo_quad_linearize           the 12-state counterpart of zopt/quadcopter.py:178-201 (linearize): exact Jacobians of the
 *                               dynamics by forward-mode differentiation (dual numbers carrying 16 tangents), which is what
 *                               jax.jacobian computes for the reference, then forward Euler A = I + dt dF/dx, B = dt dF/du
 *   zo_dfh_lqr                  zopt/lqrUtils.py:144-173 (discreteFiniteHorizonLqr): the Riccati step exactly as written
 *                               there (:167-170), terminal value Q[-1] (:172), any (n, m), time-varying operands
 *   zo_lqr_mpc_solve_batch      one receding-horizon step of zopt/mpcUtils.py:12-81 with inactive bounds for a batch of
 *                               quadcopter problems (BASELINE cfg 2): linearise at (xbar, ubar), finite-horizon gains with
 *                               terminal weight Qf, the optimal plan x+ = A x + B u, u = -L_k x from x0 = xbar (the QP of
 *                               :47-59 with no active bound has exactly this solution)
 *
 * Used ONLY by tests/ (checked against the Python oracle and the reference-generated goldens) and by bench.py's CPU legs
 * (`cpu_baseline`, `--impl reference`): problems are independent, so the batch is an OpenMP loop over problems with every
 * host thread the box has.  Nothing under zopt_b200/ links, loads or calls this file.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ZO_API __attribute__((visibility("default")))

/* ---------------------------------------------------------------------------------------------------------------
 * forward-mode differentiation: value + ND directional derivatives (12 states + 4 controls)
 * ------------------------------------------------------------------------------------------------------------- */
#define ND 16
typedef struct {
    double v, d[ND];
} dual;

static dual d_const(double c) {
    dual r;
    r.v = c;
    memset(r.d, 0, sizeof r.d);
    return r;
}
static dual d_var(double c, int k) {
    dual r = d_const(c);
    r.d[k] = 1.0;
    return r;
}
static dual d_add(dual a, dual b) {
    dual r;
    r.v = a.v + b.v;
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] + b.d[i];
    return r;
}
static dual d_sub(dual a, dual b) {
    dual r;
    r.v = a.v - b.v;
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] - b.d[i];
    return r;
}
static dual d_neg(dual a) {
    dual r;
    r.v = -a.v;
    for (int i = 0; i < ND; ++i) r.d[i] = -a.d[i];
    return r;
}
static dual d_mul(dual a, dual b) {
    dual r;
    r.v = a.v * b.v;
    for (int i = 0; i < ND; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i];
    return r;
}
static dual d_scale(double c, dual a) {
    dual r;
    r.v = c * a.v;
    for (int i = 0; i < ND; ++i) r.d[i] = c * a.d[i];
    return r;
}
static dual d_div(dual a, dual b) {
    dual r;
    const double inv = 1.0 / b.v;
    r.v = a.v * inv;
    for (int i = 0; i < ND; ++i) r.d[i] = (a.d[i] - r.v * b.d[i]) * inv;
    return r;
}
static dual d_sin(dual a) {
    dual r;
    const double c = cos(a.v);
    r.v = sin(a.v);
    for (int i = 0; i < ND; ++i) r.d[i] = c * a.d[i];
    return r;
}
static dual d_cos(dual a) {
    dual r;
    const double s = -sin(a.v);
    r.v = cos(a.v);
    for (int i = 0; i < ND; ++i) r.d[i] = s * a.d[i];
    return r;
}
static dual d_tan(dual a) {
    dual r;
    const double t = tan(a.v), s = 1.0 + t * t;
    r.v = t;
    for (int i = 0; i < ND; ++i) r.d[i] = s * a.d[i];
    return r;
}

/* zopt/quadcopter.py:116-144 with :70-113, :23-67 inlined; g = 9.807, m = 2.5, I = eye(3) (:15-18) */
static void quad_inertial_dynamics_dual(const dual* x, const dual* u, const double* wind_ned, dual* xdot) {
    const double g = 9.807, mass = 2.5;
    const dual *uvw = x, *pqr = x + 3;
    const dual phi = x[6], theta = x[7], psi = x[8];
    const dual cphi = d_cos(phi), sphi = d_sin(phi), cth = d_cos(theta), sth = d_sin(theta), cpsi = d_cos(psi), spsi = d_sin(psi);
    const dual tth = d_tan(theta);
    /* body-to-inertial rotation matrix (:23-38) */
    dual R[3][3];
    R[0][0] = d_mul(cth, cpsi);
    R[0][1] = d_sub(d_mul(d_mul(sphi, sth), cpsi), d_mul(cphi, spsi));
    R[0][2] = d_sub(d_mul(d_mul(cphi, sth), cpsi), d_mul(sphi, spsi)); /* as written at quadcopter.py:33 (the textbook matrix has + here) */
    R[1][0] = d_mul(cth, spsi);
    R[1][1] = d_add(d_mul(d_mul(sphi, sth), spsi), d_mul(cphi, cpsi));
    R[1][2] = d_sub(d_mul(d_mul(cphi, sth), spsi), d_mul(sphi, cpsi));
    R[2][0] = d_neg(sth);
    R[2][1] = d_mul(sphi, cth);
    R[2][2] = d_mul(cphi, cth);
    /* body rates -> Euler rates (:40-48) */
    dual E[3][3];
    E[0][0] = d_const(1.0); E[0][1] = d_mul(sphi, tth); E[0][2] = d_mul(cphi, tth);
    E[1][0] = d_const(0.0); E[1][1] = cphi;             E[1][2] = d_neg(sphi);
    E[2][0] = d_const(0.0); E[2][1] = d_div(sphi, cth); E[2][2] = d_div(cphi, cth);
    /* wind_body = R_b2i.T @ wind_ned (:136) */
    dual wind_body[3];
    for (int i = 0; i < 3; ++i) {
        wind_body[i] = d_const(0.0);
        if (wind_ned)
            for (int j = 0; j < 3; ++j) wind_body[i] = d_add(wind_body[i], d_scale(wind_ned[j], R[j][i]));
    }
    /* aero forces / moments (:51-67) */
    const double force_lin[3] = {-0.2, -0.2, -0.3}, force_quad[3] = {-0.05, -0.05, -0.1}, moment_lin[3] = {-0.1, -0.1, -0.05};
    dual force_aero[3], moment_aero[3];
    for (int i = 0; i < 3; ++i) {
        const dual ua = d_sub(uvw[i], wind_body[i]);
        force_aero[i] = d_add(d_scale(force_lin[i], ua), d_scale(force_quad[i], d_mul(ua, ua)));
        moment_aero[i] = d_scale(moment_lin[i], pqr[i]);
    }
    /* rigid-body dynamics (:88-113) */
    const dual d2xyz[3] = {d_neg(sth), d_mul(sphi, cth), d_mul(cphi, cth)};
    const dual force_control[3] = {d_const(0.0), d_const(0.0), d_scale(-mass, u[0])};
    dual cross_pqr_uvw[3];
    cross_pqr_uvw[0] = d_sub(d_mul(pqr[1], uvw[2]), d_mul(pqr[2], uvw[1]));
    cross_pqr_uvw[1] = d_sub(d_mul(pqr[2], uvw[0]), d_mul(pqr[0], uvw[2]));
    cross_pqr_uvw[2] = d_sub(d_mul(pqr[0], uvw[1]), d_mul(pqr[1], uvw[0]));
    for (int i = 0; i < 3; ++i) {
        const dual force_total = d_add(d_add(force_control[i], force_aero[i]), d_scale(mass * g, d2xyz[i]));
        xdot[i] = d_scale(1.0 / mass, d_sub(force_total, cross_pqr_uvw[i]));
        /* I = eye(3): cross(pqr, I pqr) = 0, I_inv = eye(3) */
        xdot[3 + i] = d_add(u[1 + i], moment_aero[i]);
    }
    /* Euler rates (:108 rows 0-1, :141 row 2) */
    for (int i = 0; i < 3; ++i) {
        dual s = d_const(0.0);
        for (int j = 0; j < 3; ++j) s = d_add(s, d_mul(E[i][j], pqr[j]));
        xdot[6 + i] = s;
    }
    /* xyzDot = R_b2i @ uvw (:142) */
    for (int i = 0; i < 3; ++i) {
        dual s = d_const(0.0);
        for (int j = 0; j < 3; ++j) s = d_add(s, d_mul(R[i][j], uvw[j]));
        xdot[9 + i] = s;
    }
}

ZO_API void zo_quad_inertial_dynamics(const double* x, const double* u, const double* wind_ned, double* xdot) {
    dual xd[12], ud[4], out[12];
    for (int i = 0; i < 12; ++i) xd[i] = d_const(x[i]);
    for (int i = 0; i < 4; ++i) ud[i] = d_const(u[i]);
    quad_inertial_dynamics_dual(xd, ud, wind_ned, out);
    for (int i = 0; i < 12; ++i) xdot[i] = out[i].v;
}

/* A (12x12), B (12x4) row-major; dt = 0: continuous-time Jacobians (quadcopter.py:195-199) */
ZO_API void zo_quad_linearize(const double* x, const double* u, const double* wind_ned, double dt, double* A, double* B) {
    dual xd[12], ud[4], out[12];
    for (int i = 0; i < 12; ++i) xd[i] = d_var(x[i], i);
    for (int i = 0; i < 4; ++i) ud[i] = d_var(u[i], 12 + i);
    quad_inertial_dynamics_dual(xd, ud, wind_ned, out);
    for (int i = 0; i < 12; ++i) {
        for (int j = 0; j < 12; ++j) A[i * 12 + j] = (dt != 0.0) ? ((i == j ? 1.0 : 0.0) + dt * out[i].d[j]) : out[i].d[j];
        for (int j = 0; j < 4; ++j) B[i * 4 + j] = (dt != 0.0) ? dt * out[i].d[12 + j] : out[i].d[12 + j];
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * small dense helpers (row-major)
 * ------------------------------------------------------------------------------------------------------------- */
#define ZO_MAXN 16
#define ZO_MAXM 8
static void mm(double* C, const double* A, const double* B, int p, int q, int r) { /* C (p x r) = A (p x q) B (q x r) */
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < r; ++j) {
            double s = 0.0;
            for (int k = 0; k < q; ++k) s += A[i * q + k] * B[k * r + j];
            C[i * r + j] = s;
        }
}
static void mtm(double* C, const double* A, const double* B, int p, int q, int r) { /* C (q x r) = A' B, A (p x q), B (p x r) */
    for (int i = 0; i < q; ++i)
        for (int j = 0; j < r; ++j) {
            double s = 0.0;
            for (int k = 0; k < p; ++k) s += A[k * q + i] * B[k * r + j];
            C[i * r + j] = s;
        }
}
/* X = G^-1 RHS by LU with partial pivoting (what jnp.linalg.solve does); G (m x m) and RHS (m x r) are overwritten */
static int lu_solve(double* G, double* RHS, int m, int r) {
    for (int c = 0; c < m; ++c) {
        int piv = c;
        for (int i = c + 1; i < m; ++i)
            if (fabs(G[i * m + c]) > fabs(G[piv * m + c])) piv = i;
        if (G[piv * m + c] == 0.0) return 1;
        if (piv != c) {
            for (int j = 0; j < m; ++j) { const double t = G[c * m + j]; G[c * m + j] = G[piv * m + j]; G[piv * m + j] = t; }
            for (int j = 0; j < r; ++j) { const double t = RHS[c * r + j]; RHS[c * r + j] = RHS[piv * r + j]; RHS[piv * r + j] = t; }
        }
        for (int i = c + 1; i < m; ++i) {
            const double f = G[i * m + c] / G[c * m + c];
            for (int j = c; j < m; ++j) G[i * m + j] -= f * G[c * m + j];
            for (int j = 0; j < r; ++j) RHS[i * r + j] -= f * RHS[c * r + j];
        }
    }
    for (int i = m - 1; i >= 0; --i)
        for (int j = 0; j < r; ++j) {
            double s = RHS[i * r + j];
            for (int k = i + 1; k < m; ++k) s -= G[i * m + k] * RHS[k * r + j];
            RHS[i * r + j] = s / G[i * m + i];
        }
    return 0;
}

/* zopt/lqrUtils.py:167-170, as written:
 *   L = solve(R + B'VB, B'VA);   V = Q + L'RL + (A - BL)' V (A - BL)          V (n x n) in/out, L (m x n) out */
static int riccati_step(int n, int m, const double* A, const double* B, const double* Q, const double* R, double* V, double* L) {
    double BtV[ZO_MAXM * ZO_MAXN], G[ZO_MAXM * ZO_MAXM], Acl[ZO_MAXN * ZO_MAXN], T1[ZO_MAXN * ZO_MAXN], RL[ZO_MAXM * ZO_MAXN];
    mtm(BtV, B, V, n, m, n);             /* B'V   (m x n) */
    mm(G, BtV, B, m, n, m);              /* B'VB  (m x m) */
    for (int i = 0; i < m * m; ++i) G[i] += R[i];
    mm(L, BtV, A, m, n, n);              /* B'VA  (m x n) */
    if (lu_solve(G, L, m, n)) return 1;
    mm(Acl, B, L, n, m, n);              /* BL */
    for (int i = 0; i < n * n; ++i) Acl[i] = A[i] - Acl[i];
    mm(T1, V, Acl, n, n, n);             /* V Acl */
    mm(RL, R, L, m, m, n);               /* R L */
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) {
            double s = Q[i * n + j];
            for (int a = 0; a < m; ++a) s += L[a * n + i] * RL[a * n + j];
            for (int k = 0; k < n; ++k) s += Acl[k * n + i] * T1[k * n + j];
            V[i * n + j] = s;
        }
    return 0;
}

/* discreteFiniteHorizonLqr for ONE problem: A (T,n,n), B (T,n,m), Q (T,n,n), R (T,m,m) with T >= N time samples; gains
 * L (N,m,n); V0 (n,n) optional: the value matrix after the last step.  Returns 0, or 1 if a pivot vanished. */
ZO_API int zo_dfh_lqr(int n, int m, int N, int T, const double* A, const double* B, const double* Q, const double* R, double* L, double* V0) {
    if (n < 1 || n > ZO_MAXN || m < 1 || m > ZO_MAXM || N < 0 || T < N || T < 1) return 2;
    double V[ZO_MAXN * ZO_MAXN];
    memcpy(V, Q + (size_t)(T - 1) * n * n, sizeof(double) * n * n); /* lqrUtils.py:172: the scan starts from Q[-1] */
    for (int k = N - 1; k >= 0; --k)
        if (riccati_step(n, m, A + (size_t)k * n * n, B + (size_t)k * n * m, Q + (size_t)k * n * n, R + (size_t)k * m * m, V,
                         L + (size_t)k * m * n))
            return 1;
    if (V0) memcpy(V0, V, sizeof(double) * n * n);
    return 0;
}

/* BASELINE cfg 2, one step for a batch: per problem b
 *   (A, B) = linearise(xbar[b], ubar[b], dt);  Q = diag(qdiag[b]), R = diag(rdiag[b]), Qf = qf_scale * Q
 *   gains L_0..L_{N-1} (terminal value Qf, as lqrMpc's cost sum_k (x'Qx + u'Ru) + x_N' Qf x_N, mpcUtils.py:51-53)
 *   plan from x0 = xbar[b]: u_k = -L_k x_k, x_{k+1} = A x_k + B u_k
 * outputs: u0 (Bsz,4), xTraj (Bsz,N+1,12), uTraj (Bsz,N,4).  threads <= 0: all the cores OpenMP sees.  Returns the number
 * of problems whose recursion hit a zero pivot (0 on success). */
ZO_API int zo_lqr_mpc_solve_batch(long long Bsz, int N, const double* xbar, const double* ubar, const double* qdiag, const double* rdiag,
                                  double qf_scale, double dt, double* u0, double* xTraj, double* uTraj, int threads) {
    int bad = 0;
#ifdef _OPENMP
    omp_set_num_threads(threads > 0 ? threads : omp_get_num_procs()); /* (not OMP_NUM_THREADS: launchers often pin it to 1) */
#else
    (void)threads;
#endif
#pragma omp parallel for schedule(static) reduction(+ : bad)
    for (long long b = 0; b < Bsz; ++b) {
        enum { n = 12, m = 4 };
        double A[n * n], B[n * m], Q[n * n], R[m * m], V[n * n];
        double* Ls = (double*)malloc(sizeof(double) * (size_t)N * m * n);
        zo_quad_linearize(xbar + b * n, ubar + b * m, NULL, dt, A, B);
        memset(Q, 0, sizeof Q);
        memset(R, 0, sizeof R);
        for (int i = 0; i < n; ++i) Q[i * n + i] = qdiag[b * n + i];
        for (int i = 0; i < m; ++i) R[i * m + i] = rdiag[b * m + i];
        for (int i = 0; i < n * n; ++i) V[i] = qf_scale * Q[i];
        int fail = 0;
        for (int k = N - 1; k >= 0 && !fail; --k) fail = riccati_step(n, m, A, B, Q, R, V, Ls + (size_t)k * m * n);
        double x[n], xn[n], u[m];
        memcpy(x, xbar + b * n, sizeof x);
        double* xT = xTraj + (size_t)b * (N + 1) * n;
        double* uT = uTraj + (size_t)b * N * m;
        memcpy(xT, x, sizeof x);
        for (int k = 0; k < N; ++k) {
            const double* L = Ls + (size_t)k * m * n;
            for (int a = 0; a < m; ++a) {
                double s = 0.0;
                for (int j = 0; j < n; ++j) s += L[a * n + j] * x[j];
                u[a] = -s;
            }
            for (int i = 0; i < n; ++i) {
                double s = 0.0;
                for (int j = 0; j < n; ++j) s += A[i * n + j] * x[j];
                for (int a = 0; a < m; ++a) s += B[i * m + a] * u[a];
                xn[i] = s;
            }
            memcpy(x, xn, sizeof x);
            memcpy(uT + (size_t)k * m, u, sizeof u);
            memcpy(xT + (size_t)(k + 1) * n, x, sizeof x);
        }
        if (N > 0) memcpy(u0 + b * m, uT, sizeof(double) * m);
        free(Ls);
        bad += fail;
    }
    return bad;
}

ZO_API int zo_num_threads(void) {
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
