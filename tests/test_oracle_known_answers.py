"""
The oracle against every known-answer value the reference's own tests hold for
the hot path.  Each test names the reference test it restates.
CPU only.
"""
import numpy as np
import pytest
import torch

from oracle import ilqr, lqr, pytrees
from oracle.quadcopter import Quadcopter

torch.set_default_dtype(torch.float64)
T = lambda x: torch.as_tensor(np.asarray(x, dtype=np.float64))


# ---------------------------------------------------------------- lqrUtils
def test_discreteFiniteHorizonLqr():  # reference tests/test_lqrUtils.py:61-69
    N = 2
    A = B = Q = R = np.repeat(np.eye(2)[None], N, axis=0)
    K = lqr.discreteFiniteHorizonLqr(A, B, Q, R, N)
    assert K[1] == pytest.approx(0.5 * np.eye(2))
    assert K[0] == pytest.approx(0.6 * np.eye(2))
    Kb = lqr.discreteFiniteHorizonLqr_batched(A[None], B[None], Q[None], R[None], N)
    assert Kb[0] == pytest.approx(K)


def test_bilinearAffineLqr():  # reference tests/test_lqrUtils.py:82-98
    N = 2
    A = B = Q = R = H = np.repeat(np.eye(2)[None], N, axis=0)
    d = q = r = np.ones((N, 2))
    q0 = np.ones(N)
    K, k = lqr.bilinearAffineLqr(A, B, d, Q, R, H, q, r, q0, N)
    assert K[1] == pytest.approx(np.eye(2))
    assert k[1] == pytest.approx(1.5 * np.ones(2))
    assert K[0] == pytest.approx(np.eye(2))
    assert k[0] == pytest.approx(np.ones(2))
    Kb, kb = lqr.bilinearAffineLqr_batched(A[None], B[None], d[None], Q[None], R[None], H[None], q[None], r[None],
                                           q0[None], N)
    assert Kb[0] == pytest.approx(K) and kb[0] == pytest.approx(k)


# ---------------------------------------------------------------- ilqrUtils
def test_trajectoryRollout():  # reference tests/test_ilqrUtils.py:7-22
    N = 3
    dynFun = lambda x, u: x + u
    policy = lambda x, k, alpha: T([alpha * k])
    trajPrev = (torch.zeros(N), torch.zeros(N))
    x0 = T([0.])
    xTraj, uTraj = ilqr.trajectoryRollout(x0, dynFun, policy, trajPrev)
    assert torch.all(xTraj == T([0, 0, 1, 3])[:, None])
    assert torch.all(uTraj == T([0, 1, 2])[:, None])
    xTraj, uTraj = ilqr.trajectoryRollout(x0, dynFun, policy, trajPrev, alpha=0.5)
    assert torch.all(xTraj == T([0, 0, 0.5, 1.5])[:, None])
    assert torch.all(uTraj == T([0, 0.5, 1])[:, None])


def test_forwardPass2():  # reference tests/test_ilqrUtils.py:41-53 (type-only smoke)
    x0 = T([1., 1])
    N = 3
    A = T([[1, 0], [1, 1]])
    B = T([[0], [1]])
    dynFun = lambda x, u: A @ x + B @ u
    costFun = lambda traj: torch.sum(traj.xTraj**2) + torch.sum(traj.uTraj**2)
    policy = lambda x, k, alpha: T([-10 * alpha])
    trajPrev = pytrees.Trajectory(x0[None, :].repeat(N + 1, 1), torch.zeros((N, 1)))
    traj, J = ilqr.forwardPass2(x0, dynFun, costFun, policy, trajPrev)
    assert isinstance(traj, pytrees.Trajectory)


def _i2_problem():
    A = torch.eye(2)
    B = torch.eye(2)
    f = torch.zeros(2)
    cost = (T(0.), torch.zeros(2), torch.zeros(2), torch.eye(2), torch.zeros((2, 2)), torch.eye(2))
    value = (T(0.), torch.zeros(2), torch.eye(2))
    return f, A, B, cost, value


def test_riccatiStep_ilqr():  # reference tests/test_ilqrUtils.py:56-81 (exact ==)
    f, A, B, cost, value = _i2_problem()
    valueOut, policy = ilqr.riccatiStep_ilqr((f, A, B), cost, value)
    assert valueOut.v == 0
    assert torch.all(valueOut.v_x == T([0, 0]))
    assert torch.all(valueOut.v_xx == 1.5 * torch.eye(2))
    assert torch.all(policy.l == T([0, 0]))
    assert torch.all(policy.L == -0.5 * torch.eye(2))


def test_riccatiStep_ddp():  # reference tests/test_ilqrUtils.py:110-135 (rel 1e-3: the clamp adds 1e-3 I)
    f, A, B, cost, value = _i2_problem()
    z = torch.zeros((2, 2, 2))
    valueOut, policy = ilqr.riccatiStep_ddp((f, A, B, z, z, z), cost, value)
    assert valueOut.v == 0
    assert valueOut.v_x.numpy() == pytest.approx(np.zeros(2))
    assert valueOut.v_xx.numpy() == pytest.approx(1.5 * np.eye(2), rel=1e-3)
    assert policy.l.numpy() == pytest.approx(np.zeros(2))
    assert policy.L.numpy() == pytest.approx(-0.5 * np.eye(2), rel=1e-3)
    # the 1e-3 clamp is added even with zero second-order terms
    assert float(policy.L[0, 0]) == pytest.approx(-1 / 2.001, rel=1e-12)


def _stacked_i2(N):
    rep = lambda a: a[None].repeat(N, *([1] * a.ndim))
    dyn = pytrees.AffineDynamics(torch.zeros((N, 2)), rep(torch.eye(2)), rep(torch.eye(2)))
    cost = pytrees.QuadraticCostFunction(torch.zeros(N), torch.zeros((N, 2)), torch.zeros((N, 2)), rep(torch.eye(2)),
                                         torch.zeros((N, 2, 2)), rep(torch.eye(2)))
    value = pytrees.QuadraticValueFunction(T(0.), torch.zeros(2), torch.eye(2))
    return dyn, cost, value


def test_backwardPass_ilqr():  # reference tests/test_ilqrUtils.py:84-107
    dyn, cost, value = _stacked_i2(2)
    policy = ilqr.backwardPass_ilqr(dyn, cost, value)
    assert isinstance(policy, pytrees.AffinePolicy)
    assert policy.L.shape == (2, 2, 2)
    # two steps of the I2 recursion: v_xx 1 -> 1.5 -> 1.6 ; L = -v/(1+v)
    assert policy.L[1].numpy() == pytest.approx(-0.5 * np.eye(2))
    assert policy.L[0].numpy() == pytest.approx(-0.6 * np.eye(2))


def test_backwardPass_ddp():  # reference tests/test_ilqrUtils.py:138-164
    N = 2
    dyn, cost, value = _stacked_i2(N)
    C = torch.stack([torch.eye(2), torch.eye(2)])[None].repeat(N, 1, 1, 1)
    dynq = pytrees.QuadraticDynamics(dyn.f, dyn.f_x, dyn.f_u, C, C.clone(), torch.zeros((N, 2, 2, 2)))
    policy = ilqr.backwardPass_ddp(dynq, cost, value)
    assert isinstance(policy, pytrees.AffinePolicy)


@pytest.mark.parametrize("solver", [ilqr.iterativeLqr, ilqr.differentialDynamicProgramming])
def test_solvers_converge(solver):  # reference tests/test_ilqrUtils.py:167-196
    A = B = Q = R = torch.eye(2)
    N = 3
    dynamics = lambda x, u: A @ x + B @ u
    runningCost = lambda x, u: x @ Q @ x + u @ R @ u
    terminalCost = lambda x: x @ Q @ x
    x0 = T([2., 1])
    uGuess = torch.zeros((N, 2))
    trajectory, L, J, converged = solver(dynamics, runningCost, terminalCost, x0, uGuess)
    assert converged
    assert trajectory.xTraj.shape == (N + 1, 2) and L.shape == (N, 2, 2)


# ---------------------------------------------------------------- pytrees (reference tests/test_pytrees.py)
def test_Trajectory():  # :6-19
    m, n, N = 2, 3, 4
    xTraj = torch.arange((N + 1) * n).reshape((N + 1, n))
    uTraj = torch.arange(N * m).reshape((N, m))
    traj = pytrees.Trajectory(xTraj, uTraj)
    for i in range(N):
        assert torch.all(traj[i].xTraj == xTraj[i]) and torch.all(traj[i].uTraj == uTraj[i])


def test_CostFunction():  # :22-32, :35-44
    runningCost = lambda x, u: x @ x + u @ u
    traj = pytrees.Trajectory(T([[1, 2], [3, 4]]), T([[1, 1]]))
    C = pytrees.CostFunction(runningCost, lambda x: 2 * x @ x)
    assert C(traj, k=0) == 7 and C(traj) == 57
    C = pytrees.CostFunction.runningOnly(runningCost, 2)
    assert C(traj, k=0) == 7 and C(traj) == 32


def test_QuadraticValueFunction():  # :47-56 (V(x)=851.5), :59-69
    V = pytrees.QuadraticValueFunction(T(1.), T([2., 3]), T([[4., 5], [6, 7]]))
    assert float(V(T([8., 9]))) == pytest.approx(1 + 16 + 27 + 807.5)
    c, c_x, c_xx = 1, T([1, 2]), torch.eye(2)
    costFun = pytrees.CostFunction(0, lambda x: c + c_x @ x + 0.5 * x @ c_xx @ x)
    value = pytrees.QuadraticValueFunction.fromTerminalCostFunction(costFun, torch.zeros(2))
    assert value.v == c and torch.all(value.v_x == c_x) and torch.all(value.v_xx == c_xx)


def test_QuadraticCostFunction():  # :72-90 (C(x,u)=51), :119-139, :142-172
    C = pytrees.QuadraticCostFunction(T(0.), T([1, 2]), T([2, 1]), torch.eye(2), torch.ones((2, 2)), torch.eye(2))
    assert float(C(T([1., 2]), T([3., 4]))) == pytest.approx(0. + 5 + 10 + 2.5 + 21 + 12.5)
    c, c_x, c_u, c_xx, c_ux, c_uu = 1, T([1, 2]), T([2, 1]), torch.eye(2), T([[1, 2], [3, 4]]), torch.eye(2)
    costFun = pytrees.CostFunction.runningOnly(
        lambda x, u: c + c_x @ x + c_u @ u + 0.5 * (x @ c_xx @ x + 2 * u @ c_ux @ x + u @ c_uu @ u), 2)
    Cq = pytrees.QuadraticCostFunction.from_function(costFun, torch.zeros(2), torch.zeros(2))
    assert Cq.c == c
    for got, exp in zip(tuple(Cq)[1:], (c_x, c_u, c_xx, c_ux, c_uu)):
        assert torch.all(got == exp)
    x0 = T([[0., 0], [1, 0], [1, 1]])
    traj = pytrees.Trajectory(x0, torch.zeros((2, 2)))
    Ct = pytrees.QuadraticCostFunction.from_trajectory(costFun, traj)
    C1 = Ct[1]
    assert C1.c == costFun(traj, k=1)
    assert torch.all(C1.c_x == c_x + c_xx @ x0[1]) and torch.all(C1.c_u == c_u + c_ux @ x0[1])
    assert torch.all(C1.c_ux == c_ux)


def test_AffineDynamics():  # :175-187, :207-219, :222-238
    f, f_x, f_u = T([1, 1]), T([[2, 3], [4, 5]]), T([[6], [7]])
    dyn = pytrees.AffineDynamics(f, f_x, f_u)
    assert dyn(T([1, 2]), T([2])).numpy() == pytest.approx(np.array([21, 29]))
    dynFun = lambda x, u: f + f_x @ x + f_u @ u + 0.5 * x @ x
    x0 = T([[0., 0], [1, 0], [2, 0]])
    traj = pytrees.Trajectory(x0, torch.zeros((2, 1)))
    d = pytrees.AffineDynamics.from_trajectory(dynFun, traj)
    assert torch.all(d[0].f == f) and torch.all(d[0].f_x == f_x) and torch.all(d[0].f_u == f_u)
    assert torch.all(d[1].f == dynFun(x0[1], traj.uTraj[1]))
    assert torch.all(d[1].f_x == f_x + x0[1]) and torch.all(d[1].f_u == f_u)


def test_QuadraticDynamics():  # :241-263, :266-281, :284-307
    f, f_x, f_u = torch.zeros(2), torch.eye(2), torch.eye(2)
    f_xx = torch.stack([torch.eye(2), 2 * torch.eye(2)])
    f_ux = torch.zeros((2, 2, 2))
    f_uu = torch.stack([torch.eye(2), 2 * torch.eye(2)])
    dyn = pytrees.QuadraticDynamics(f, f_x, f_u, f_xx, f_ux, f_uu)
    assert torch.all(dyn(T([1, 0]), T([0, 1])) == T([2., 3]))
    dq = pytrees.QuadraticDynamics.from_function(lambda x, u: dyn(x, u), torch.zeros(2), torch.zeros(2))
    for got, exp in zip(dq, dyn):
        assert torch.all(got == exp)
    traj = pytrees.Trajectory(torch.zeros((3, 2)), torch.zeros((2, 2)))
    dt = pytrees.QuadraticDynamics.from_trajectory(lambda x, u: dyn(x, u), traj)
    for got, exp in zip(dt[1], dyn):
        assert torch.all(got == exp)


def test_AffinePolicy():  # :310-313 ([6,13]), :316-329
    policy = pytrees.AffinePolicy(T([1, 2]), T([[1, 2], [3, 4]]))
    assert torch.all(policy(T([1, 2])) == T([6, 13]))
    n, m, N = 2, 3, 2
    l = torch.arange(0, N * m, dtype=torch.float64).reshape((N, m))
    L = torch.arange(0, N * m * n, dtype=torch.float64).reshape((N, m, n))
    policy = pytrees.AffinePolicy(l, L)
    x = torch.zeros(2)
    assert torch.all(policy(x, k=1) == l[1]) and torch.all(policy(x, k=0, alpha=0.5) == 0.5 * l[0])
    with pytest.raises(ValueError):
        policy(x)


def test_QuadraticDeltaCost():  # :332-337
    dJ = pytrees.QuadraticDeltaCost(1, 2)
    assert dJ(1) == 3 and dJ(0.5) == 1


# ---------------------------------------------------------------- quadcopter (reference tests/test_quadcopter.py)
def test_rotation_matrices():  # :12-43
    ac = Quadcopter()
    assert ac._bodyToInertialRotationMatrix(0., 0., 0.).numpy() == pytest.approx(np.eye(3))
    th = np.pi / 6
    c, s, t = np.cos(th), np.sin(th), np.tan(th)
    assert ac._bodyToInertialRotationMatrix(th, 0., 0.).numpy() == pytest.approx(np.array([[1, 0, 0], [0, c, -s], [0, s, c]]))
    assert ac._bodyToInertialRotationMatrix(0., th, 0.).numpy() == pytest.approx(np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]]))
    assert ac._bodyToInertialRotationMatrix(0., 0., th).numpy() == pytest.approx(np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]]))
    assert ac._bodyRatesToEulerRatesRotationMatrix(0., 0.).numpy() == pytest.approx(np.eye(3))
    assert ac._bodyRatesToEulerRatesRotationMatrix(th, 0.).numpy() == pytest.approx(np.array([[1, 0, 0], [0, c, -s], [0, s, c]]))
    assert ac._bodyRatesToEulerRatesRotationMatrix(0., th).numpy() == pytest.approx(np.array([[1, 0, t], [0, 1, 0], [0, 0, 1 / c]]))


def test_rigidBodyDynamics():  # :46-57
    ac = Quadcopter()
    xDot = ac.rigidBodyDynamics(np.zeros(9), np.zeros(4))
    assert xDot.numpy() == pytest.approx(np.array([0, 0, 9.807, 0, 0, 0, 0, 0]))
    xDot = ac.rigidBodyDynamics(np.zeros(9), np.array([9.807, 0, 0, 0]))
    assert xDot.numpy() == pytest.approx(np.zeros(8))


def test_inertialDynamics():  # :60-86
    ac = Quadcopter()
    control = np.array([9.807, 0, 0, 0])
    assert ac.inertialDynamics(np.zeros(12), control).numpy() == pytest.approx(np.zeros(12))
    uvw = np.array([0.1, 0.2, 0.3])
    state = np.zeros(12)
    state[0:3] = uvw
    assert ac.inertialDynamics(state, control)[9:].numpy() == pytest.approx(uvw)
    state[8] = np.pi / 2
    assert ac.inertialDynamics(state, control)[9:].numpy() == pytest.approx(np.array([-uvw[1], uvw[0], uvw[2]]))


def test_trim_and_linearize():  # :89-116
    ac = Quadcopter()
    x0, u0 = ac.trim(np.zeros(3))
    assert x0[0:3] == pytest.approx(np.zeros(3))
    assert ac.rigidBodyDynamics(x0, u0).numpy() == pytest.approx(np.zeros(8), abs=1e-3)
    assert u0 == pytest.approx(np.array([9.807, 0, 0, 0]), abs=1e-6)
    for dt in (0, 1):
        A, B = ac.linearize(np.zeros(8), np.array([9.807, 0, 0, 0]), dt=dt)
        assert A.shape == (8, 8) and B.shape == (8, 4)
        assert not (torch.any(torch.isnan(A)) or torch.any(torch.isnan(B)))
