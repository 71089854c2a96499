"""
Oracle (test infrastructure) for zopt/mpcUtils.py:12-81 `lqrMpc`.

The reference builds the QP below with cvxpy (mpcUtils.py:47-59) and hands it to a third-party solver
(cvxpy==1.7.2 -> OSQP 1.0.4 in demos/lqrMpc.py:32, the cvxpy default QP solver in tests/test_mpcUtils.py:21);
neither is installed here and no reference test pins a number (only `status == "optimal"`), so for this
path PARITY IS UNPINNED.  What the oracle does instead:

    min  sum_{k<N} x_k'Q x_k + u_k'R u_k + x_N'Qf x_N            (no 1/2; Qf defaults to Q, mpcUtils.py:44-45,52-54)
    s.t. x_{k+1} = A x_k + B u_k                                  (:55)
         x_lb <= x_k <= x_ub for k = 0..N,  u_lb <= u_k <= u_ub   (:56-57)
         x_0 = x0                                                 (:58)

  * `solve_qp`       -- condenses the QP onto u and solves it to ~1e-10 with a dense Mehrotra primal-dual
                        interior-point method (a different algorithm from the kernel's ADMM);
  * `kkt_residuals`  -- independent check of any candidate solution: stationarity, feasibility, complementarity;
  * `is_feasible`    -- LP feasibility of the constraint set with SciPy's HiGHS;
  * `riccati_plan`   -- the exact solution when no bound binds (Riccati sweep + rollout), the bridge to lqrUtils.
"""
import numpy as np


def condense(A, B, Q, R, Qf, N, x0):
    """x (stacked, (N+1)n incl. x_0) = Phi x0 + Gam u ; cost = 1/2 u'Hu + g'u + const."""
    n, m = B.shape
    Phi = np.zeros(((N + 1) * n, n))
    Gam = np.zeros(((N + 1) * n, N * m))
    Phi[:n] = np.eye(n)
    for k in range(N):
        Phi[(k + 1) * n:(k + 2) * n] = A @ Phi[k * n:(k + 1) * n]
        Gam[(k + 1) * n:(k + 2) * n] = A @ Gam[k * n:(k + 1) * n]
        Gam[(k + 1) * n:(k + 2) * n, k * m:(k + 1) * m] += B
    Qt = np.zeros(((N + 1) * n, (N + 1) * n))
    for k in range(N):
        Qt[k * n:(k + 1) * n, k * n:(k + 1) * n] = Q
    Qt[N * n:, N * n:] = Qf
    Rt = np.kron(np.eye(N), R)
    Qs, Rs = Qt + Qt.T, Rt + Rt.T
    H = Gam.T @ Qs @ Gam + Rs
    g = Gam.T @ Qs @ (Phi @ x0)
    return Phi, Gam, H, g


def constraints(Phi, Gam, N, n, m, x0, x_lb, x_ub, u_lb, u_ub):
    """finite rows of  G u <= h"""
    xfree = Phi @ x0
    G, h = [], []
    xl, xu = np.tile(x_lb, N + 1), np.tile(x_ub, N + 1)
    ul, uu = np.tile(u_lb, N), np.tile(u_ub, N)
    I = np.eye(N * m)
    for i in range((N + 1) * n):
        if np.isfinite(xu[i]):
            G.append(Gam[i]); h.append(xu[i] - xfree[i])
        if np.isfinite(xl[i]):
            G.append(-Gam[i]); h.append(-(xl[i] - xfree[i]))
    for i in range(N * m):
        if np.isfinite(uu[i]):
            G.append(I[i]); h.append(uu[i])
        if np.isfinite(ul[i]):
            G.append(-I[i]); h.append(-ul[i])
    if not G:
        return np.zeros((0, N * m)), np.zeros(0)
    return np.array(G), np.array(h)


def is_feasible(G, h):
    if G.shape[0] == 0:
        return True
    from scipy.optimize import linprog
    res = linprog(np.zeros(G.shape[1]), A_ub=G, b_ub=h, bounds=[(None, None)] * G.shape[1], method="highs")
    return res.status == 0


def _ipm(H, g, G, h, tol=1e-10, max_iter=100):
    """Mehrotra predictor-corrector for min 1/2 u'Hu + g'u s.t. Gu + s = h, s >= 0."""
    nv, mc = H.shape[0], G.shape[0]
    if mc == 0:
        return np.linalg.solve(H, -g), np.zeros(0)
    u = np.linalg.solve(H, -g)
    s = np.maximum(h - G @ u, 1.0)
    z = np.ones(mc)
    for _ in range(max_iter):
        rd = H @ u + g + G.T @ z
        rp = G @ u + s - h
        mu = s @ z / mc
        if max(np.abs(rd).max(), np.abs(rp).max(), mu) < tol:
            break
        W = z / s
        Hs = H + G.T @ (W[:, None] * G)
        try:
            c = np.linalg.cholesky(Hs)
        except np.linalg.LinAlgError:  # conditioning limit of the normal equations: stop at the current iterate
            break

        def solve(rc):
            # Newton system: H du + G'dz = -rd ; G du + ds = -rp ; z ds + s dz = -rc
            rhs = -rd - G.T @ ((-rc + z * rp) / s)
            du = np.linalg.solve(c.T, np.linalg.solve(c, rhs))
            ds = -rp - G @ du
            dz = (-rc - z * ds) / s
            return du, ds, dz

        du, ds, dz = solve(s * z)

        def step(v, dv):
            neg = dv < 0
            return min(1.0, (-v[neg] / dv[neg]).min()) if neg.any() else 1.0

        a_aff = min(step(s, ds), step(z, dz))
        mu_aff = (s + a_aff * ds) @ (z + a_aff * dz) / mc
        sigma = (mu_aff / mu)**3
        du, ds, dz = solve(s * z + ds * dz - sigma * mu)
        a = 0.99 * min(step(s, ds), step(z, dz))
        u, s, z = u + a * du, s + a * ds, z + a * dz
    rd = np.abs(H @ u + g + G.T @ z).max()
    rp = max(0.0, (G @ u - h).max())
    scale = max(1.0, np.abs(g).max())
    if not (rd < 1e-6 * scale and rp < 1e-7 and s @ z / mc < 1e-7 * scale):
        raise RuntimeError(f"oracle IPM did not converge: rd={rd:.2e} rp={rp:.2e} gap={s @ z / mc:.2e}")
    return u, z


def solve_qp(A, B, Q, R, N, x_lb, x_ub, u_lb, u_ub, x0, Qf=None):
    """Returns (u0 (m), xTraj (N+1,n), uTraj (N,m), status, info) with status in {"optimal", "infeasible"}."""
    A, B, Q, R, x0 = (np.asarray(a, dtype=np.float64) for a in (A, B, Q, R, x0))
    Qf = Q if Qf is None else np.asarray(Qf, dtype=np.float64)
    n, m = B.shape
    x_lb, x_ub, u_lb, u_ub = (np.asarray(a, dtype=np.float64) for a in (x_lb, x_ub, u_lb, u_ub))
    Phi, Gam, H, g = condense(A, B, Q, R, Qf, N, x0)
    G, h = constraints(Phi, Gam, N, n, m, x0, x_lb, x_ub, u_lb, u_ub)
    if not is_feasible(G, h):
        nan = np.full
        return nan(m, np.nan), nan((N + 1, n), np.nan), nan((N, m), np.nan), "infeasible", {}
    u, z = _ipm(H, g, G, h)
    x = (Phi @ x0 + Gam @ u).reshape(N + 1, n)
    uT = u.reshape(N, m)
    J = sum(x[k] @ Q @ x[k] + uT[k] @ R @ uT[k] for k in range(N)) + x[N] @ Qf @ x[N]
    return uT[0], x, uT, "optimal", dict(H=H, g=g, G=G, h=h, z=z, J=J)


def kkt_residuals(A, B, Q, R, N, x_lb, x_ub, u_lb, u_ub, x0, uTraj, Qf=None):
    """Independent optimality check of a candidate plan `uTraj`: least-squares multipliers on the active set.
    Returns dict(stationarity, primal_violation, min_multiplier) -- all ~0 / >= 0 at the optimum."""
    A, B, Q, R, x0 = (np.asarray(a, dtype=np.float64) for a in (A, B, Q, R, x0))
    Qf = Q if Qf is None else np.asarray(Qf, dtype=np.float64)
    n, m = B.shape
    Phi, Gam, H, g = condense(A, B, Q, R, Qf, N, x0)
    G, h = constraints(Phi, Gam, N, n, m, x0, *(np.asarray(a, dtype=np.float64) for a in (x_lb, x_ub, u_lb, u_ub)))
    u = np.asarray(uTraj, dtype=np.float64).reshape(-1)
    grad = H @ u + g
    if G.shape[0] == 0:
        return dict(stationarity=np.abs(grad).max(), primal_violation=0.0, min_multiplier=0.0)
    slack = h - G @ u
    scale = max(1.0, np.abs(h[np.isfinite(h)]).max() if len(h) else 1.0)
    act = slack < 1e-4 * scale
    viol = max(0.0, -slack.min())
    if act.any():
        from scipy.optimize import nnls
        lam, _ = nnls(G[act].T, -grad)
        stat = np.abs(grad + G[act].T @ lam).max()
        return dict(stationarity=stat, primal_violation=viol, min_multiplier=lam.min() if len(lam) else 0.0)
    return dict(stationarity=np.abs(grad).max(), primal_violation=viol, min_multiplier=0.0)


def riccati_plan(A, B, Q, R, N, x0, Qf=None):
    """Exact optimum with no active bound: Riccati sweep with terminal Qf + closed-loop rollout."""
    A, B, Q, R, x0 = (np.asarray(a, dtype=np.float64) for a in (A, B, Q, R, x0))
    Qf = Q if Qf is None else np.asarray(Qf, dtype=np.float64)
    V = Qf
    Ls = []
    for _ in range(N):
        L = np.linalg.solve(R + B.T @ V @ B, B.T @ V @ A)
        Acl = A - B @ L
        V = Q + L.T @ R @ L + Acl.T @ V @ Acl
        Ls.append(L)
    Ls = Ls[::-1]
    x, xs, us = x0, [x0], []
    for k in range(N):
        u = -Ls[k] @ x
        x = A @ x + B @ u
        xs.append(x)
        us.append(u)
    return np.array(xs), np.array(us)
