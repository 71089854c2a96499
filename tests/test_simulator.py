"""
Batched closed-loop simulator (zopt_b200/simulator.py) -- mirror of the discrete half of zopt/simulator.py:124-169.
CPU: the oracle restatement against hand-computed sequences (the reference has no test for simulator.py) and the host
logic of the mirror; GPU: the two closed loops the reference demos run (demos/iterativeLqr.py:44-56 tracking controller in
wind, demos/discreteFiniteHorizonLqr.py:38-49 8-state gains on the 12-state plant) against the oracle, batched.
"""
import numpy as np
import pytest
import torch

from oracle import simulator as osim
from oracle.quadcopter import Quadcopter as OQuadcopter


def test_oracle_simulator_known_sequence():
    """x+ = 0.5 x + u with u = -0.25 x (stateless controller): x_k = 0.25^k, u_k = -0.25^(k+1); shapes of simulator.py:160-169"""
    dyn = osim.SimBlock(lambda k, x, u: (None, 0.5 * x + u), np.array([1.0]), dt=0.5)
    ctrl = osim.SimBlock(lambda k, xc, x: (-0.25 * x, np.array([])), np.array([]), dt=0.5)
    t, x0A, x1A, y0A, y1A = osim.Simulator([ctrl, dyn], (0, 2.0)).simulate()
    assert t == pytest.approx([0, 0.5, 1.0, 1.5, 2.0]) and x0A.shape == (5, 0)
    assert x1A[:, 0] == pytest.approx(0.25 ** np.arange(5))
    assert y0A[:, 0] == pytest.approx(-0.25 ** (np.arange(4) + 1)) and y0A.shape == (4, 1)


def test_mirror_rejects_unregistered_blocks_without_gpu():
    from zopt_b200.models import QuadcopterEuler
    from zopt_b200.simulator import SimBlock, Simulator, TrackingController
    dyn = SimBlock(QuadcopterEuler(0.1), np.zeros(12), dt=0.1)
    with pytest.raises(TypeError):
        Simulator([SimBlock(lambda k, xc, x: (x, xc), np.array([]), dt=0.1), dyn], (0, 1.0))
    with pytest.raises(TypeError):
        Simulator([SimBlock(TrackingController(np.zeros((11, 12)), np.zeros((10, 4)), np.zeros((10, 4, 12))), np.array([]), dt=0.1),
                   SimBlock(lambda k, x, u: (None, x), np.zeros(12), dt=0.1)], (0, 1.0))
    with pytest.raises(NotImplementedError):
        Simulator([SimBlock(TrackingController(np.zeros((11, 12)), np.zeros((10, 4)), np.zeros((10, 4, 12))), np.array([]), dt=0),
                   SimBlock(QuadcopterEuler(0.1), np.zeros(12), dt=0)], (0, 1.0))
    # the registered controllers are callable like the demos' lambdas
    c = TrackingController(np.ones((3, 2)), np.zeros((2, 1)), np.ones((2, 1, 2)))
    assert c(1, np.array([]), np.array([2.0, 3.0]))[0] == pytest.approx([3.0])


def _plan(Bsz, N, rng):
    ac = OQuadcopter()
    xT = np.zeros((Bsz, N + 1, 12))
    xT[:, :, 9:12] = np.linspace(1, 0, N + 1)[None, :, None] * rng.uniform(-5, 5, (Bsz, 1, 3))
    uT = np.tile(np.array([9.807, 0, 0, 0]), (Bsz, N, 1)) + 0.05 * rng.normal(size=(Bsz, N, 4))
    L = -0.05 * rng.uniform(0.5, 1.5, (Bsz, N, 4, 12)) * (rng.uniform(size=(Bsz, N, 4, 12)) < 0.3)
    return ac, xT, uT, L


@pytest.mark.gpu
@pytest.mark.parametrize("dt_", [torch.float64, torch.float32])
def test_tracking_controller_in_wind_vs_oracle(dt_):
    """demos/iterativeLqr.py:44-56: u = L_k (x - xTraj_k) + uTraj_k on x + dt f(x, u, wind_ned=[3,1,0])"""
    from zopt_b200.models import QuadcopterEuler
    from zopt_b200.simulator import SimBlock, Simulator, TrackingController
    rng = np.random.default_rng(21)
    Bsz, N, dt, wind = 9, 25, 0.1, np.array([3.0, 1.0, 0.0])
    ac, xT, uT, L = _plan(Bsz, N, rng)
    x0 = xT[:, 0] + 0.1 * rng.normal(size=(Bsz, 12))
    cu = lambda a: torch.as_tensor(a, dtype=dt_, device="cuda")
    sim = Simulator([SimBlock(TrackingController(cu(xT), cu(uT), cu(L)), np.array([]), dt=dt, name="Controller"),
                     SimBlock(QuadcopterEuler(dt, wind), cu(x0), dt=dt, name="Dynamics")], (0, N * dt))
    tS, x0A, xS, uS, y1 = sim.simulate()
    assert xS.shape == (Bsz, N + 1, 12) and uS.shape == (Bsz, N, 4) and x0A.shape == (Bsz, N + 1, 0) and tS.shape == (N + 1,)
    tol = 1e-9 if dt_ == torch.float64 else 5e-5  # random (not stabilising) gains amplify rounding along the 25 steps
    for b in range(Bsz):
        dyn = osim.SimBlock(lambda k, x, u: (None, x + dt * ac.inertialDynamics(torch.as_tensor(x), torch.as_tensor(u), torch.as_tensor(wind)).numpy()), x0[b], dt=dt)
        ctrl = osim.SimBlock(lambda k, xc, x, b=b: (L[b, k] @ (x - xT[b, k]) + uT[b, k], np.array([])), np.array([]), dt=dt)
        t, _, xr, ur, _ = osim.Simulator([ctrl, dyn], (0, N * dt)).simulate()
        assert np.max(np.abs(xS[b].cpu().numpy() - xr)) <= tol * np.max(np.abs(xr))
        assert np.max(np.abs(uS[b].cpu().numpy() - ur)) <= tol * np.max(np.abs(ur))
    assert tS.cpu().numpy() == pytest.approx(t)


@pytest.mark.gpu
def test_lqr_demo_closed_loop_vs_oracle():
    """demos/discreteFiniteHorizonLqr.py:24-49: gains of the 8-state rigid-body hover linearisation drive the 12-state plant
    through x[:8]; un-batched call with reference shapes"""
    from oracle import lqr as olqr
    from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
    from zopt_b200.models import QuadcopterEuler
    from zopt_b200.simulator import ProportionalFeedbackController, SimBlock, Simulator
    ac = OQuadcopter()
    dt, T = 0.1, 10
    N = int(T / dt)
    xTrim, uTrim = np.zeros(8), np.array([9.807, 0, 0, 0])
    A, B = (t.numpy() for t in ac.linearize(xTrim, uTrim, dt))
    Q, R = np.eye(8), np.eye(4)
    Ak, Bk, Rk = (np.tile(M[None], (N, 1, 1)) for M in (A, B, R))
    Qk = np.concatenate([10 * Q[None], np.tile(Q[None], (N - 1, 1, 1))])
    K = discreteFiniteHorizonLqr(Ak, Bk, Qk, Rk, N)
    Kref = olqr.discreteFiniteHorizonLqr(Ak, Bk, Qk, Rk, N)
    x0 = np.array([0.5, 0, 0, 0, 0, 0, np.pi / 16, 0, 0, 0, 0, -10.0])
    sim = Simulator([SimBlock(ProportionalFeedbackController(xTrim, uTrim, K, ns=8), np.array([]), dt=dt),
                     SimBlock(QuadcopterEuler(dt), x0, dt=dt)], (0, T))
    tS, _, xS, uS, _ = sim.simulate()
    assert xS.shape == (N + 1, 12) and uS.shape == (N, 4)
    dyn = osim.SimBlock(lambda k, x, u: (None, x + dt * ac.inertialDynamics(torch.as_tensor(x), torch.as_tensor(u)).numpy()), x0, dt=dt)
    ctrl = osim.SimBlock(lambda k, xc, x: (-Kref[k] @ (x[:8] - xTrim) + uTrim, np.array([])), np.array([]), dt=dt)
    t, _, xr, ur, _ = osim.Simulator([ctrl, dyn], (0, T)).simulate()
    assert np.max(np.abs(xS.cpu().numpy() - xr)) <= 1e-9 * np.max(np.abs(xr))
    assert np.max(np.abs(uS.cpu().numpy() - ur)) <= 1e-9 * np.max(np.abs(ur))
