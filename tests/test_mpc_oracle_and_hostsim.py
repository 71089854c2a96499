"""
lqrMpc (zopt/mpcUtils.py:12-81): the oracle QP solver against independent checks, and the ADMM kernel body
(host build) against the oracle.  Reference parity for this path is UNPINNED (the reference test asserts only
`status == "optimal"`, tests/test_mpcUtils.py:8-23); gates are KKT residuals, agreement with the tight-tolerance
oracle, and exact equality to the Riccati plan when no bound binds.
"""
import ctypes as C

import numpy as np
import pytest

from oracle import mpc as ompc
from tests import hostsim as H

INF = np.inf


def demo_problem():
    """demos/lqrMpc.py:11-31 with the hover linearisation (zero wind), N = 25"""
    import torch
    from oracle.quadcopter import Quadcopter
    ac = Quadcopter()
    A, B = (t.numpy() for t in ac.linearizeInertial(np.zeros(12), np.array([9.807, 0, 0, 0]), 0.1))
    Q, R = np.eye(12), np.eye(4)
    x_ub = np.array([1, 1, 1, 0.3, 0.3, 0.1, 0.5, 0.5, INF, INF, INF, INF])
    u_ub = np.array([3.0, 3, 3, 3])
    return A, B, Q, R, 25, -x_ub, x_ub, -u_ub, u_ub


def run_admm(A, B, Q, R, N, xlb, xub, ulb, uub, x0, Qf=None, dt=np.float64, **kw):
    Qf = Q if Qf is None else Qf
    x0 = np.atleast_2d(np.asarray(x0, dtype=dt))
    Bsz, n, m = x0.shape[0], B.shape[-2], B.shape[-1]
    mats = [np.ascontiguousarray(a, dtype=dt) for a in (A, B, Q, R, Qf)]
    vecs = [np.ascontiguousarray(a, dtype=dt) for a in (xlb, xub, ulb, uub)]
    zs = [H.arr(a, 2, False, a.ndim == 3) for a in mats] + [H.arr(a, 1, False, a.ndim == 2) for a in vecs]
    u0, xT, uT = np.zeros((Bsz, m), dtype=dt), np.zeros((Bsz, N + 1, n), dtype=dt), np.zeros((Bsz, N, m), dtype=dt)
    status, iters = np.zeros(Bsz, dtype=np.int8), np.zeros(Bsz, dtype=np.int32)
    ws = np.zeros(Bsz * H.hs.hs_admm_ws_elems(N, n, m), dtype=dt)
    H.hs.hs_mpc_admm(int(dt == np.float64), C.c_int64(Bsz), N, n, m, *[C.byref(z) for z in zs], H.P(x0),
                     int(kw.get("max_iter", 4000)), int(kw.get("check_every", 25)), C.c_double(kw.get("rho", 0.1)),
                     C.c_double(kw.get("alpha", 1.6)), C.c_double(kw.get("eps_abs", 1e-3)), C.c_double(kw.get("eps_rel", 1e-3)),
                     C.c_double(kw.get("eps_inf", 1e-4)), H.P(u0), H.P(xT), H.P(uT), H.P(status), H.P(iters), H.P(ws))
    return u0, xT, uT, status, iters


H.hs.hs_admm_ws_elems.restype = C.c_longlong


def test_oracle_reference_test_problem():
    """tests/test_mpcUtils.py:8-23: I2, N=2, bounds +-1, x0 = 1 -> "optimal"; bounds stay inactive, so the optimum is the
    Riccati plan u=[[-.6,-.6],[-.2,-.2]], x=[[1,1],[.4,.4],[.2,.2]], J*=3.2 (derived)"""
    I = np.eye(2)
    one = np.ones(2)
    u0, x, u, status, info = ompc.solve_qp(I, I, I, I, 2, -one, one, -one, one, one)
    assert status == "optimal"
    assert u == pytest.approx(np.array([[-0.6, -0.6], [-0.2, -0.2]]), abs=1e-8)
    assert x == pytest.approx(np.array([[1, 1], [0.4, 0.4], [0.2, 0.2]]), abs=1e-8)
    assert info["J"] == pytest.approx(3.2, abs=1e-8)
    xr, ur = ompc.riccati_plan(I, I, I, I, 2, one)
    assert ur == pytest.approx(u, abs=1e-8)


def test_oracle_constrained_two_methods():
    """active bounds: interior-point solution vs independent KKT check (NNLS multipliers on the active set)"""
    A, B, Q, R, N, xlb, xub, ulb, uub = demo_problem()
    x0 = np.zeros(12)
    x0[9:12] = [10, 10, 10]
    u0, x, u, status, info = ompc.solve_qp(A, B, Q, R, N, xlb, xub, ulb, uub, x0)
    assert status == "optimal"
    assert np.max(np.abs(x[:, :3])) > 0.999  # the velocity bound binds (10 m offset, |v| <= 1)
    k = ompc.kkt_residuals(A, B, Q, R, N, xlb, xub, ulb, uub, x0, u)
    assert k["stationarity"] < 1e-6 and k["primal_violation"] < 1e-8 and k["min_multiplier"] >= 0
    # infeasible start (x0 outside the box) is reported
    x0b = x0.copy()
    x0b[0] = 2.0
    assert ompc.solve_qp(A, B, Q, R, N, xlb, xub, ulb, uub, x0b)[3] == "infeasible"


def test_admm_body_reference_test_problem():
    I = np.eye(2)
    one = np.ones(2)
    u0, xT, uT, status, iters = run_admm(I, I, I, I, 2, -one, one, -one, one, one, eps_abs=1e-9, eps_rel=1e-9)
    assert status[0] == 0
    assert uT[0] == pytest.approx(np.array([[-0.6, -0.6], [-0.2, -0.2]]), abs=1e-6)
    assert xT[0] == pytest.approx(np.array([[1, 1], [0.4, 0.4], [0.2, 0.2]]), abs=1e-6)


@pytest.mark.parametrize("eps,tol,tolJ", [(1e-3, 1e-1, 2e-2), (1e-7, 5e-4, 1e-6)])  # OSQP-default and tight tolerances
def test_admm_body_constrained_vs_oracle(eps, tol, tolJ):
    A, B, Q, R, N, xlb, xub, ulb, uub = demo_problem()
    rng = np.random.default_rng(4)
    Bsz = 4
    x0 = np.zeros((Bsz, 12))
    x0[:, 9:12] = rng.uniform(-10, 10, (Bsz, 3))
    x0[0, 9:12] = [10, 10, 10]  # the demo's initial state
    u0, xT, uT, status, iters = run_admm(A, B, Q, R, N, xlb, xub, ulb, uub, x0, eps_abs=eps, eps_rel=eps, max_iter=20000)
    assert (status == 0).all(), (status, iters)
    for b in range(Bsz):
        ur0, xr, ur, st, info = ompc.solve_qp(A, B, Q, R, N, xlb, xub, ulb, uub, x0[b])
        assert st == "optimal"
        assert np.max(np.abs(uT[b] - ur)) < tol * max(1.0, np.max(np.abs(ur)))
        assert np.max(np.abs(xT[b] - xr)) < tol * max(1.0, np.max(np.abs(xr)))
        J = sum(xT[b, k] @ Q @ xT[b, k] + uT[b, k] @ R @ uT[b, k] for k in range(N)) + xT[b, N] @ Q @ xT[b, N]
        assert abs(J - info["J"]) < tolJ * info["J"]
        if eps < 1e-5:
            k = ompc.kkt_residuals(A, B, Q, R, N, xlb, xub, ulb, uub, x0[b], uT[b])
            assert k["stationarity"] < 1e-3 and k["primal_violation"] < 1e-5
        # the plan satisfies the dynamics exactly (they are never relaxed)
        assert np.max(np.abs(xT[b, 1:] - (xT[b, :-1] @ A.T + uT[b] @ B.T))) < 1e-12


def test_admm_body_inactive_bounds_equal_riccati_and_infeasible():
    A, B, Q, R, N, xlb, xub, ulb, uub = demo_problem()
    x0 = np.zeros(12)
    x0[9:12] = [0.05, -0.05, 0.02]  # small offset: no bound binds
    u0, xT, uT, status, iters = run_admm(A, B, Q, R, N, xlb, xub, ulb, uub, x0, eps_abs=1e-10, eps_rel=1e-10, max_iter=20000)
    xr, ur = ompc.riccati_plan(A, B, Q, R, N, x0)
    assert status[0] == 0 and np.max(np.abs(uT[0] - ur)) < 1e-7 and np.max(np.abs(xT[0] - xr)) < 1e-7
    # infeasible: x0 outside the state box
    x0b = x0.copy()
    x0b[0] = 2.0
    u0, xT, uT, status, iters = run_admm(A, B, Q, R, N, xlb, xub, ulb, uub, x0b)
    assert status[0] == 2 and np.isnan(uT).all()
    # infeasible through the dynamics: a state that cannot be held inside a tiny box (certificate path)
    A2 = np.array([[1.0, 1.0], [0.0, 1.0]])
    B2 = np.array([[0.0], [1.0]])
    lb, ub = np.array([-1.0, -1.0]), np.array([1.0, 1.0])
    x02 = np.array([0.9, 1.0])  # x1+ = 1.9 > 1 whatever u is
    assert ompc.solve_qp(A2, B2, np.eye(2), np.eye(1), 3, lb, ub, np.array([-0.1]), np.array([0.1]), x02)[3] == "infeasible"
    u0, xT, uT, status, iters = run_admm(A2, B2, np.eye(2), np.eye(1), 3, lb, ub, np.array([-0.1]), np.array([0.1]), x02)
    assert status[0] == 2
