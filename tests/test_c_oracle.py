"""
oracle/c/zopt_oracle.c -- the plain-C (C99 + OpenMP) restatement of the reference's headline path that bench.py times on
the host cores (`cpu_baseline`, `--impl reference`) -- pinned like the Python oracle: against the golden fixtures produced
by the UNMODIFIED reference (tests/golden/*.npz, scripts/gen_golden_from_reference.py) and against the Python oracle on
random inputs.  CPU only; the product never loads this library.
"""
import os

import numpy as np
import pytest
import torch

from oracle import c_oracle as co
from oracle import lqr as olqr
from oracle.quadcopter import Quadcopter as OQuadcopter

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
load = lambda name: np.load(os.path.join(G, name), allow_pickle=False)
T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
rep = lambda a, N: np.repeat(np.asarray(a)[None], N, axis=0)


def relerr(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a.astype(np.float64) - b)) / (den if den > 0 else 1.0))


def test_c_quadcopter_against_reference_goldens():
    """zopt/quadcopter.py:116-144 and its jax.jacobian: dynamics (with and without wind) and Jacobians at the golden points"""
    g = load("quadcopter_points.npz")
    for i in range(len(g["x"])):
        x, u = g["x"][i], g["u"][i]
        assert relerr(co.inertialDynamics(x, u), g["F"][i]) < 1e-13
        assert relerr(co.inertialDynamics(x, u, g["wind"]), g["F_wind"][i]) < 1e-13
        jx, ju = co.linearizeInertial(x, u, 0.0)
        assert relerr(jx, g["Jx"][i]) < 1e-13 and relerr(ju, g["Ju"][i]) < 1e-13
    # wind path of the Jacobian against the Python oracle's autodiff
    oac = OQuadcopter()
    rng = np.random.default_rng(5)
    for _ in range(4):
        x, u, w = 0.5 * rng.normal(size=12), rng.normal(size=4) + np.array([9.8, 0, 0, 0]), rng.normal(size=3) * 3
        jx, ju = torch.func.jacrev(lambda a, b: oac.inertialDynamics(a, b, T(w)), argnums=(0, 1))(T(x), T(u))
        A, B = co.linearizeInertial(x, u, 0.1, w)
        assert relerr(A, np.eye(12) + 0.1 * jx.numpy()) < 1e-13 and relerr(B, 0.1 * ju.numpy()) < 1e-13


def test_c_riccati_against_reference_goldens_and_python_oracle():
    """zopt/lqrUtils.py:144-173: the demo's (8,4) problem and 32 cfg 2 problems as the reference solved them; random
    time-varying problems of several shapes against the Python oracle, value matrix included"""
    g = load("lqr_demo_n8.npz")
    N = int(g["N"])
    assert relerr(co.discreteFiniteHorizonLqr(rep(g["A"], N), rep(g["B"], N), g["Qk"], g["Rk"], N), g["K"]) < 1e-12
    g = load("lqr_cfg2_32.npz")
    N = int(g["N"])
    for i in range(32):
        A, B = co.linearizeInertial(g["xbar"][i], g["ubar"][i], 0.1)
        assert relerr(A, g["A"][i]) < 1e-13 and relerr(B, g["B"][i]) < 1e-13
        Qk = rep(np.diag(g["qdiag"][i]), N + 1)
        Qk[N] *= 10
        assert relerr(co.discreteFiniteHorizonLqr(rep(A, N), rep(B, N), Qk, rep(np.diag(g["rdiag"][i]), N), N), g["L"][i]) < 1e-11
    rng = np.random.default_rng(7)
    for (n, m, N) in ((2, 1, 100), (8, 4, 30), (12, 4, 50), (16, 8, 9), (5, 3, 1)):
        A = np.eye(n) + 0.1 * rng.normal(size=(N, n, n))
        B = 0.3 * rng.normal(size=(N, n, m))
        Mq, Mr = 0.3 * rng.normal(size=(N + 1, n, n)), 0.3 * rng.normal(size=(N, m, m))
        Q = np.eye(n) + Mq @ Mq.transpose(0, 2, 1)
        R = np.eye(m) + Mr @ Mr.transpose(0, 2, 1)
        L, V0 = co.discreteFiniteHorizonLqr(A, B, Q, R, N, return_V0=True)
        Lo, Vo = olqr.discreteFiniteHorizonLqr(A, B, Q, R, N, return_value=True)
        assert relerr(L, Lo) < 1e-11 and relerr(V0, Vo) < 1e-11, (n, m, N)


@pytest.mark.parametrize("threads", [1, 0])
def test_c_lqr_mpc_batch_equals_python_port(threads):
    """the batch entry point bench.py times (one unconstrained lqrMpc step per problem, BASELINE cfg 2) against the same
    step composed from the Python oracle: linearise, gains with terminal weight 10 Q, the optimal plan"""
    import importlib.util
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("_zb_configs_t", os.path.join(root, "zopt_b200", "configs.py"))
    cfgs = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(cfgs)
    d = cfgs.cfg2(Bsz=37)
    N, dt = int(d["N"]), float(d["dt"])
    u0, xT, uT = co.lqrMpcSolveBatch(d["xbar"], d["ubar"], d["qdiag"], d["rdiag"], N, dt, 10.0, threads)
    oac = OQuadcopter()
    for b in (0, 11, 36):
        A, B = (t.numpy() for t in oac.linearizeInertial(d["xbar"][b], d["ubar"][b], dt))
        Qk = rep(np.diag(d["qdiag"][b]), N + 1)
        Qk[N] *= 10
        L = olqr.discreteFiniteHorizonLqr(rep(A, N), rep(B, N), Qk, rep(np.diag(d["rdiag"][b]), N), N)
        L = L.numpy() if isinstance(L, torch.Tensor) else np.asarray(L)
        x = d["xbar"][b].copy()
        for k in range(N):
            u = -L[k] @ x
            assert relerr(uT[b, k], u) < 1e-10 and relerr(xT[b, k], x) < 1e-10
            x = A @ x + B @ u
        assert relerr(xT[b, N], x) < 1e-10 and np.array_equal(u0[b], uT[b, 0])
