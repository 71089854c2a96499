"""
Oracle (test infrastructure): torch-CPU fp64 restatement of zopt/ilqrUtils.py.

  trajectoryRollout               <- ilqrUtils.py:33-66
  forwardPass2                    <- ilqrUtils.py:116-150
  riccatiStep_ilqr                <- ilqrUtils.py:153-173
  backwardPass_ilqr               <- ilqrUtils.py:176-181
  riccatiStep_ddp                 <- ilqrUtils.py:184-206
  backwardPass_ddp                <- ilqrUtils.py:209-214
  ensurePositiveDefinite          <- ilqrUtils.py:217-219
  conditionQuadraticCost          <- ilqrUtils.py:222-234
  conditionQuadraticDynamics      <- ilqrUtils.py:237-251
  conditionValueFunction          <- ilqrUtils.py:254-257
  iterativeLqr                    <- ilqrUtils.py:260-327
  differentialDynamicProgramming  <- ilqrUtils.py:330-397

`lax.scan` -> Python loop, `lax.while_loop` -> Python while, `jax.vmap` over the
16 step sizes -> explicit loop (identical arithmetic per step size),
`jnp.linalg.solve/eigh` -> `torch.linalg.solve/eigh` (same LAPACK routines).
The solvers additionally return a per-iteration log of every discrete decision
(step-size index, cost) so parity can be checked iteration by iteration.
The legacy `forwardPass` (ilqrUtils.py:69-113) is not called by either solver
and is not restated.
"""
import torch

from .pytrees import (AffineDynamics, AffinePolicy, CostFunction, QuadraticCostFunction, QuadraticDynamics,
                      QuadraticValueFunction, Trajectory)


def trajectoryRollout(x0, dynFun, policy, trajPrev, alpha=1):  # ilqrUtils.py:33-66
    xPrev, uPrev = trajPrev
    N = uPrev.shape[0]
    x = x0
    xs, us = [x0], []
    for k in range(N):
        dx = x - xPrev[k]
        u = policy(dx, k=k, alpha=alpha) + uPrev[k]  # :59-60
        x = dynFun(x, u)
        xs.append(x)
        us.append(u)
    return Trajectory(torch.stack(xs), torch.stack(us))


def forwardPass2(x0, dynFun, costFun, policy, trajPrev, return_all=False):  # ilqrUtils.py:116-150
    JArr, trajArr = [], []
    for j in range(16):
        alpha = 0.5**j  # :145
        t = trajectoryRollout(x0, dynFun, policy, trajPrev, alpha=alpha)
        JArr.append(costFun(t))
        trajArr.append(t)
    JArr = torch.stack([torch.as_tensor(J) for J in JArr])
    # jnp.argmin: first index on ties; NaN wins (NumPy semantics)
    if torch.isnan(JArr).any():
        idx = int(torch.nonzero(torch.isnan(JArr))[0, 0])
    else:
        idx = int(torch.argmin(JArr))
    if return_all:
        return trajArr[idx], JArr[idx], idx, JArr
    return trajArr[idx], JArr[idx]


def _riccati_core(c, c_x, c_u, Q_xx, Q_ux, Q_uu, f_x, f_u, v, v_x):
    Q = c + v
    Q_x = c_x + f_x.T @ v_x
    Q_u = c_u + f_u.T @ v_x
    l = -torch.linalg.solve(Q_uu, Q_u)
    L = -torch.linalg.solve(Q_uu, Q_ux)
    valueOut = QuadraticValueFunction(Q - 0.5 * l @ Q_uu @ l, Q_x - L.T @ Q_uu @ l, Q_xx - L.T @ Q_uu @ L)
    return valueOut, AffinePolicy(l, L)


def riccatiStep_ilqr(dynamics, cost, value):  # ilqrUtils.py:153-173
    _, f_x, f_u = dynamics
    c, c_x, c_u, c_xx, c_ux, c_uu = cost
    v, v_x, v_xx = value
    Q_xx = c_xx + f_x.T @ v_xx @ f_x
    Q_uu = c_uu + f_u.T @ v_xx @ f_u
    Q_ux = c_ux + f_u.T @ v_xx @ f_x
    return _riccati_core(c, c_x, c_u, Q_xx, Q_ux, Q_uu, f_x, f_u, v, v_x)


def backwardPass_ilqr(dynamics, cost, Vf):  # ilqrUtils.py:176-181
    N = len(cost.c)
    V = Vf
    ls, Ls = [None] * N, [None] * N
    for k in range(N - 1, -1, -1):
        V, pol = riccatiStep_ilqr(dynamics[k], cost[k], V)
        ls[k], Ls[k] = pol
    return AffinePolicy(torch.stack(ls), torch.stack(Ls))


def ensurePositiveDefinite(a, eps=1e-3):  # ilqrUtils.py:217-219
    # jnp.linalg.eigh defaults to symmetrize_input=True, i.e. it factors (a + a^T)/2
    w, v = torch.linalg.eigh(0.5 * (a + a.T))
    return (v * torch.clamp(w, min=eps)) @ v.T


def conditionQuadraticCost(quadratic_cost):  # ilqrUtils.py:222-234
    (c, c_x, c_u, c_xx, c_ux, c_uu) = quadratic_cost
    n = c_xx.shape[1]
    m = c_uu.shape[1]
    c_zz = torch.cat([torch.cat([c_xx, c_ux.transpose(1, 2)], dim=2), torch.cat([c_ux, c_uu], dim=2)], dim=1)
    c_zz = torch.stack([ensurePositiveDefinite(z) for z in c_zz])
    c_xx, c_uu, c_ux = c_zz[:, :n, :n], c_zz[:, -m:, -m:], c_zz[:, -m:, :n]
    return QuadraticCostFunction(c, c_x, c_u, c_xx, c_ux, c_uu)


def conditionQuadraticDynamics(quadratic_dynamics, v_x):  # ilqrUtils.py:237-251
    _, _, _, f_xx, f_ux, f_uu = quadratic_dynamics
    vf_xx = torch.einsum('i,ijk', v_x, f_xx)
    vf_uu = torch.einsum('i,ijk', v_x, f_uu)
    vf_ux = torch.einsum('i,ijk', v_x, f_ux)
    n = vf_xx.shape[0]
    m = vf_uu.shape[0]
    vf_zz = torch.cat([torch.cat([vf_xx, vf_ux.T], dim=1), torch.cat([vf_ux, vf_uu], dim=1)], dim=0)
    vf_zz = ensurePositiveDefinite(vf_zz)
    vf_xx, vf_uu, vf_ux = vf_zz[:n, :n], vf_zz[-m:, -m:], vf_zz[-m:, :n]
    return vf_xx, vf_ux, vf_uu


def conditionValueFunction(Vf):  # ilqrUtils.py:254-257
    v, v_x, v_xx = Vf
    return QuadraticValueFunction(v, v_x, ensurePositiveDefinite(v_xx))


def riccatiStep_ddp(dynamics, cost, value):  # ilqrUtils.py:184-206
    c, c_x, c_u, c_xx, c_ux, c_uu = cost
    v, v_x, v_xx = value
    _, f_x, f_u, _, _, _ = dynamics
    vf_xx, vf_ux, vf_uu = conditionQuadraticDynamics(dynamics, v_x)
    Q_xx = c_xx + f_x.T @ v_xx @ f_x + vf_xx
    Q_uu = c_uu + f_u.T @ v_xx @ f_u + vf_uu
    Q_ux = c_ux + f_u.T @ v_xx @ f_x + vf_ux
    return _riccati_core(c, c_x, c_u, Q_xx, Q_ux, Q_uu, f_x, f_u, v, v_x)


def backwardPass_ddp(dynamics, cost, Vf):  # ilqrUtils.py:209-214
    N = len(cost.c)
    V = Vf
    ls, Ls = [None] * N, [None] * N
    for k in range(N - 1, -1, -1):
        V, pol = riccatiStep_ddp(dynamics[k], cost[k], V)
        ls[k], Ls[k] = pol
    return AffinePolicy(torch.stack(ls), torch.stack(Ls))


def _solve(dynamics, runningCost, terminalCost, x0, uGuess, maxIter, tol, second_order, log):
    # ilqrUtils.py:289-327 (iLQR) / :359-397 (DDP): identical except for the dynamics expansion
    n = x0.shape[0]
    N, m = uGuess.shape
    cost = CostFunction(runningCost, terminalCost)
    policy = AffinePolicy(uGuess, torch.zeros((N, m, n), dtype=x0.dtype))
    traj_prev = Trajectory(torch.zeros((N + 1, n), dtype=x0.dtype), torch.zeros((N, m), dtype=x0.dtype))
    traj = trajectoryRollout(x0, dynamics, policy, traj_prev)  # :297
    J = cost(traj)
    converged, it = False, 0
    if log is not None:
        log.append(dict(iter=-1, J=float(J), alpha_idx=-1))
    while (not converged) and it < maxIter:  # :301-303
        if second_order:
            dyn = QuadraticDynamics.from_trajectory(dynamics, traj)  # :378
        else:
            dyn = AffineDynamics.from_trajectory(dynamics, traj)  # :308
        quadratic_cost = QuadraticCostFunction.from_trajectory(cost, traj)  # :309
        Vf = QuadraticValueFunction.fromTerminalCostFunction(cost, traj.xTraj[-1])  # :310
        quadratic_cost = conditionQuadraticCost(quadratic_cost)  # :312
        Vf = conditionValueFunction(Vf)  # :313
        policy = (backwardPass_ddp if second_order else backwardPass_ilqr)(dyn, quadratic_cost, Vf)  # :315
        traj_new, J_new, idx, JArr = forwardPass2(x0, dynamics, cost, policy, traj, return_all=True)  # :316
        converged = bool(abs(J - J_new) <= tol)  # :318
        traj, J = traj_new, J_new
        it += 1
        if log is not None:
            log.append(dict(iter=it - 1, J=float(J), alpha_idx=idx, JArr=JArr.clone(), l=policy.l.clone(),
                            L=policy.L.clone(), xTraj=traj.xTraj.clone(), uTraj=traj.uTraj.clone()))
    return traj, policy.L, J, converged


def iterativeLqr(dynamics, runningCost, terminalCost, x0, uGuess, maxIter=100, tol=1e-3, log=None):
    """ilqrUtils.py:260-327. Returns (Trajectory, L (N,m,n), J, converged)."""
    return _solve(dynamics, runningCost, terminalCost, x0, uGuess, maxIter, tol, False, log)


def differentialDynamicProgramming(dynamics, runningCost, terminalCost, x0, uGuess, maxIter=100, tol=1e-3, log=None):
    """ilqrUtils.py:330-397. Returns (Trajectory, L (N,m,n), J, converged)."""
    return _solve(dynamics, runningCost, terminalCost, x0, uGuess, maxIter, tol, True, log)
