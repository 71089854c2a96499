"""
Golden fixtures (tests/golden/*.npz) produced by the UNMODIFIED reference sources run on oracle/jax_shim
(scripts/gen_golden_from_reference.py; jax itself is not installable here).  CPU tests pin the oracle against them;
GPU tests (-m gpu) pin the CUDA path against the same files.  fp64: 1e-10 max-norm relative.
"""
import os

import numpy as np
import pytest
import torch

from oracle import ilqr as oilqr
from oracle import lqr as olqr
from oracle import pytrees as opt
from oracle.quadcopter import Quadcopter as OQuadcopter
from zopt_b200 import configs

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
load = lambda name: np.load(os.path.join(G, name), allow_pickle=False)
T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))


def relerr(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = b.detach().cpu().numpy() if isinstance(b, torch.Tensor) else np.asarray(b)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a.astype(np.float64) - b)) / (den if den > 0 else 1.0))


def rep(a, N):
    return np.repeat(np.asarray(a)[None], N, axis=0)


# ------------------------------------------------------------------------------------------------ oracle vs goldens (CPU)
def test_oracle_lqr_goldens():
    g = load("lqr_demo_n8.npz")
    N = int(g["N"])
    assert relerr(olqr.discreteFiniteHorizonLqr(rep(g["A"], N), rep(g["B"], N), g["Qk"], g["Rk"], N), g["K"]) < 1e-12
    g = load("lqr_cfg2_32.npz")
    N = int(g["N"])
    oac = OQuadcopter()
    for i in range(32):
        A, B = (t.numpy() for t in oac.linearizeInertial(g["xbar"][i], g["ubar"][i], 0.1))
        assert relerr(A, g["A"][i]) < 1e-13 and relerr(B, g["B"][i]) < 1e-13
        Qk = rep(np.diag(g["qdiag"][i]), N + 1)
        Qk[N] *= 10
        L = olqr.discreteFiniteHorizonLqr(rep(A, N), rep(B, N), Qk, rep(np.diag(g["rdiag"][i]), N), N)
        assert relerr(L, g["L"][i]) < 1e-11
    g = load("bilinear_demo.npz")
    N = int(g["N"])
    L, l = olqr.bilinearAffineLqr(rep(g["A"], N), rep(g["B"], N), g["d"], rep(np.eye(8), N), rep(np.eye(4), N), g["H"], g["q"],
                                  g["r"], g["q0"], N)
    assert relerr(L, g["L"]) < 1e-11 and relerr(l, g["l"]) < 1e-11


def test_oracle_quadcopter_goldens():
    g = load("quadcopter_points.npz")
    oac = OQuadcopter()
    for i in range(len(g["x"])):
        x, u = T(g["x"][i]), T(g["u"][i])
        assert relerr(oac.inertialDynamics(x, u), g["F"][i]) < 1e-13
        assert relerr(oac.inertialDynamics(x, u, T(g["wind"])), g["F_wind"][i]) < 1e-13
        jx, ju = torch.func.jacrev(oac.inertialDynamics, argnums=(0, 1))(x, u)
        assert relerr(jx, g["Jx"][i]) < 1e-13 and relerr(ju, g["Ju"][i]) < 1e-13


def _solver_golden(name):
    g = load(name)
    ddp = name.startswith("ddp")
    oac = OQuadcopter()
    Q, R = torch.eye(12, dtype=torch.float64), T(g["R"])
    N = int(g["N"])
    return g, ddp, oac.eulerStep(0.1), (lambda x, u: x @ Q @ x + u @ R @ u), (lambda x: 10 * x @ Q @ x), N


@pytest.mark.parametrize("name", ["ilqr_demo_N40_it4.npz", "ddp_demo_N40_it3.npz"])
def test_oracle_solver_goldens(name):
    g, ddp, dyn, rc, tc, N = _solver_golden(name)
    log = []
    solver = oilqr.differentialDynamicProgramming if ddp else oilqr.iterativeLqr
    traj, L, J, conv = solver(dyn, rc, tc, T(g["x0"]), T(rep(configs.U_TRIM, N)), maxIter=int(g["maxIter"]), tol=-1.0, log=log)
    assert relerr(np.array([e["J"] for e in log]), g["J_per_iter"]) < 1e-11
    assert relerr(traj.xTraj, g["xTraj"]) < 1e-10 and relerr(traj.uTraj, g["uTraj"]) < 1e-10 and relerr(L, g["L"]) < 1e-10



def _sim_blocks_a(g, mod, conv=lambda a: a):
    """demos/iterativeLqr.py:44-56: tracking controller on the plant in wind [3,1,0]; blocks as plain callables"""
    oac = OQuadcopter()
    dt, wind = float(g["a_dt"]), g["a_wind"]
    L, xT, uT = g["a_L"], g["a_xTraj"], g["a_uTraj"]
    dyn = mod.SimBlock(lambda k, x, u: (None, x + dt * oac.inertialDynamics(T(x), T(u), T(wind)).numpy()), g["a_x0"], dt=dt)
    ctrl = mod.SimBlock(lambda k, xc, x: (L[k] @ (x - xT[k]) + uT[k], np.array([])), np.array([]), dt=dt)
    return [ctrl, dyn], (0, L.shape[0] * dt)


def test_oracle_simulator_goldens():
    """oracle/simulator.py against the reference's own Simulator (zopt/simulator.py:124-169, run unmodified on the shim) on
    the two closed loops of the demos (demos/iterativeLqr.py:44-56, demos/discreteFiniteHorizonLqr.py:38-49)."""
    from oracle import simulator as osim
    g = load("simulator_demos.npz")
    blocks, span = _sim_blocks_a(g, osim)
    t, x0A, xA, uA, _ = osim.Simulator(blocks, span).simulate()
    assert np.allclose(t, g["a_t"]) and relerr(xA, g["a_x"]) < 1e-12 and relerr(uA, g["a_u"]) < 1e-12 and x0A.shape == (len(t), 0)
    oac = OQuadcopter()
    K, uTrim, dt = g["b_K"], g["b_uTrim"], 0.1
    dyn = osim.SimBlock(lambda k, x, u: (None, x + dt * oac.inertialDynamics(T(x), T(u)).numpy()), g["b_x0"], dt=dt)
    ctrl = osim.SimBlock(lambda k, xc, x: (-K[k] @ (x[:8] - np.zeros(8)) + uTrim, np.array([])), np.array([]), dt=dt)
    t, _, xB, uB, _ = osim.Simulator([ctrl, dyn], (0, 10)).simulate()
    assert np.allclose(t, g["b_t"]) and relerr(xB, g["b_x"]) < 1e-12 and relerr(uB, g["b_u"]) < 1e-12


@pytest.mark.parametrize("name", ["ilqr_cfg4_N200_it10.npz", "ddp_cfg5_N100_it10.npz"])
def test_oracle_solver_goldens_bench_size(name):
    """BASELINE cfg 4 / cfg 5 size (N = 200 / 100, 10 iterations): the oracle against the reference run on problem 0 --
    cost after EVERY iteration, final trajectory and gains."""
    g, ddp, dyn, rc, tc, N = _solver_golden(name)
    log = []
    solver = oilqr.differentialDynamicProgramming if ddp else oilqr.iterativeLqr
    traj, L, J, conv = solver(dyn, rc, tc, T(g["x0"][0]), T(rep(configs.U_TRIM, N)), maxIter=10, tol=-1.0, log=log)
    assert relerr(np.array([e["J"] for e in log]), g["J_per_iter"][0]) < 1e-11
    assert [e["alpha_idx"] for e in log[1:]] == g["alpha_idx"][0].tolist()
    assert relerr(traj.xTraj, g["xTraj"][0]) < 1e-9 and relerr(traj.uTraj, g["uTraj"][0]) < 1e-9 and relerr(L, g["L"][0]) < 1e-9



def _care_problem():
    g = load("care_tv.npz")
    A = lambda t: g["A0"] + np.sin(2.0 * float(t)) * g["A1"]
    B = lambda t: g["B0"] + float(t) * g["B1"]
    Q = lambda t: (1.0 + 0.5 * float(t)) * g["Q"]
    Ri = lambda t: g["R_inv"]
    return g, A, B, Q, Ri


def test_oracle_continuous_finite_horizon_lqr():
    """oracle.lqr.finiteHorizonLqr (zopt/lqrUtils.py:39-98) against the reference's own test (tests/test_lqrUtils.py:31-44:
    K(T) = I exactly, K(0) = the analytic scalar Riccati solution) and against the reference run on the shim for a
    time-varying problem, at grid and off-grid times (both integrate with a Dormand-Prince pair at 1.4e-8)."""
    I2 = lambda t: np.eye(2)
    K = olqr.finiteHorizonLqr(I2, I2, I2, I2, np.eye(2), 1, N=4)
    assert np.allclose(K(1), np.eye(2), atol=1e-12)
    s2 = np.sqrt(2)
    K_exp = lambda t: ((1 + s2) * np.exp(2 * s2) - (s2 - 1) * np.exp(2 * s2 * t)) / (np.exp(2 * s2 * t) + np.exp(2 * s2))
    assert np.allclose(K(0), K_exp(0) * np.eye(2), rtol=1e-6)
    g, A, B, Q, Ri = _care_problem()
    K = olqr.finiteHorizonLqr(A, B, Q, Ri, g["Qf"], float(g["T"]), N=int(g["N"]))
    for tq, Kr in zip(g["tq"], g["K"]):
        assert relerr(K(tq), Kr) < 1e-9


def test_oracle_riccati_step_goldens():
    g = load("riccati_steps.npz")
    n = 5
    czz = g["czz"]
    cost = opt.QuadraticCostFunction(T(g["c"]), T(g["c_x"]), T(g["c_u"]), T(czz[:n, :n]), T(czz[n:, :n]), T(czz[n:, n:]))
    val = opt.QuadraticValueFunction(T(g["v"]), T(g["v_x"]), T(g["v_xx"]))
    vo, p = oilqr.riccatiStep_ilqr(opt.AffineDynamics(torch.zeros(n), T(g["f_x"]), T(g["f_u"])), cost, val)
    assert relerr(vo.v_xx, g["ilqr_vxx"]) < 1e-12 and relerr(vo.v_x, g["ilqr_vx"]) < 1e-12 and relerr(p.L, g["ilqr_L"]) < 1e-12
    assert relerr(p.l, g["ilqr_l"]) < 1e-12 and abs(float(vo.v) - float(g["ilqr_v"])) < 1e-12
    vo, p = oilqr.riccatiStep_ddp(opt.QuadraticDynamics(torch.zeros(n), T(g["f_x"]), T(g["f_u"]), T(g["f_xx"]), T(g["f_ux"]), T(g["f_uu"])),
                                  cost, val)
    assert relerr(vo.v_xx, g["ddp_vxx"]) < 1e-11 and relerr(p.L, g["ddp_L"]) < 1e-11 and relerr(p.l, g["ddp_l"]) < 1e-11
    assert relerr(oilqr.ensurePositiveDefinite(T(g["S"])), g["S_pd"]) < 1e-12


# ------------------------------------------------------------------------------------------------ CUDA path vs goldens
@pytest.mark.gpu
def test_gpu_lqr_goldens():
    from zopt_b200.lqrUtils import bilinearAffineLqr, discreteFiniteHorizonLqr
    from zopt_b200.quadcopter import Quadcopter
    g = load("lqr_demo_n8.npz")
    N = int(g["N"])
    assert relerr(discreteFiniteHorizonLqr(rep(g["A"], N), rep(g["B"], N), g["Qk"], g["Rk"], N), g["K"]) < 1e-10
    g = load("lqr_cfg2_32.npz")
    N = int(g["N"])
    A, B = Quadcopter().linearizeInertial(g["xbar"], g["ubar"], 0.1)
    assert relerr(A, g["A"]) < 1e-12 and relerr(B, g["B"]) < 1e-12
    Q = configs.diag_embed(g["qdiag"])
    Qk = np.repeat(Q[:, None], N + 1, axis=1)
    Qk[:, N] *= 10
    Rk = torch.as_tensor(configs.diag_embed(g["rdiag"]), device="cuda")[:, None].expand(-1, N, -1, -1)
    for dt, tol in ((torch.float64, 1e-10), (torch.float32, 1e-5)):
        L = discreteFiniteHorizonLqr(A[:, None].expand(-1, N, -1, -1).to(dt), B[:, None].expand(-1, N, -1, -1).to(dt),
                                     torch.as_tensor(Qk, dtype=dt, device="cuda"), Rk.to(dt), N)
        err = np.max(np.abs(L.double().cpu().numpy() - g["L"]), axis=(1, 2, 3)) / np.max(np.abs(g["L"]), axis=(1, 2, 3))
        assert err.max() < tol
    g = load("bilinear_demo.npz")
    N = int(g["N"])
    L, l = bilinearAffineLqr(rep(g["A"], N), rep(g["B"], N), g["d"], rep(np.eye(8), N), rep(np.eye(4), N), g["H"], g["q"], g["r"],
                             g["q0"], N)
    assert relerr(L, g["L"]) < 1e-10 and relerr(l, g["l"]) < 1e-10


@pytest.mark.gpu
def test_gpu_quadcopter_goldens():
    from zopt_b200.quadcopter import Quadcopter, quad_hess_contract
    g = load("quadcopter_points.npz")
    ac = Quadcopter()
    assert relerr(ac.inertialDynamics(g["x"], g["u"]), g["F"]) < 1e-12
    assert relerr(ac.inertialDynamics(g["x"], g["u"], g["wind"]), g["F_wind"]) < 1e-12
    A, B = ac.linearizeInertial(g["x"], g["u"], dt=0)
    assert relerr(A, g["Jx"]) < 1e-12 and relerr(B, g["Ju"]) < 1e-12
    rng = np.random.default_rng(0)
    lam = rng.normal(size=(len(g["x"]), 12))
    assert relerr(quad_hess_contract(g["x"], g["u"], None, 0.0, lam), np.einsum('bi,bijk->bjk', lam, g["Hxx"])) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ilqr_demo_N40_it4.npz", "ddp_demo_N40_it3.npz"])
def test_gpu_solver_goldens(name):
    from zopt_b200 import ilqrUtils
    from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost
    g, ddp, _, _, _, N = _solver_golden(name)
    solver = ilqrUtils.differentialDynamicProgramming if ddp else ilqrUtils.iterativeLqr
    traj, L, J, conv, log = solver(QuadcopterEuler(0.1), QuadraticCost(np.eye(12), g["R"]), QuadraticTerminalCost(10 * np.eye(12)),
                                   g["x0"], rep(configs.U_TRIM, N), maxIter=int(g["maxIter"]), tol=-1.0, return_log=True)
    assert relerr(log["J"], g["J_per_iter"]) < 1e-10
    assert relerr(traj.xTraj, g["xTraj"]) < 1e-10 and relerr(traj.uTraj, g["uTraj"]) < 1e-10 and relerr(L, g["L"]) < 1e-10
    assert abs(float(J) - g["J_per_iter"][-1]) < 1e-10 * g["J_per_iter"][-1]


@pytest.mark.gpu
def test_gpu_riccati_step_goldens():
    from zopt_b200 import ilqrUtils, pytrees
    g = load("riccati_steps.npz")
    n = 5
    czz = g["czz"]
    cost = pytrees.QuadraticCostFunction(g["c"], g["c_x"], g["c_u"], czz[:n, :n], czz[n:, :n], czz[n:, n:])
    val = pytrees.QuadraticValueFunction(g["v"], g["v_x"], g["v_xx"])
    vo, p = ilqrUtils.riccatiStep_ilqr(pytrees.AffineDynamics(np.zeros(n), g["f_x"], g["f_u"]), cost, val)
    assert relerr(vo.v_xx, g["ilqr_vxx"]) < 1e-11 and relerr(vo.v_x, g["ilqr_vx"]) < 1e-11 and relerr(p.L, g["ilqr_L"]) < 1e-11
    assert relerr(p.l, g["ilqr_l"]) < 1e-11 and abs(float(vo.v) - float(g["ilqr_v"])) < 1e-11 * max(1, abs(float(g["ilqr_v"])))
    vo, p = ilqrUtils.riccatiStep_ddp(pytrees.QuadraticDynamics(np.zeros(n), g["f_x"], g["f_u"], g["f_xx"], g["f_ux"], g["f_uu"]), cost, val)
    assert relerr(vo.v_xx, g["ddp_vxx"]) < 1e-10 and relerr(p.L, g["ddp_L"]) < 1e-10 and relerr(p.l, g["ddp_l"]) < 1e-10
    assert relerr(ilqrUtils.ensurePositiveDefinite(g["S"]), g["S_pd"]) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["ilqr_cfg4_N200_it10.npz", "ddp_cfg5_N100_it10.npz"])
def test_gpu_solver_goldens_bench_size(name):
    """BASELINE cfg 4 / cfg 5 size (N = 200 / 100, 10 iterations, four problems of the configs' own distributions, one batched
    call): cost after every iteration, the step-size index of every iteration, final x, u, L against the reference run."""
    from zopt_b200 import ilqrUtils
    from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost
    g, ddp, _, _, _, N = _solver_golden(name)
    solver = ilqrUtils.differentialDynamicProgramming if ddp else ilqrUtils.iterativeLqr
    traj, L, J, conv, log = solver(QuadcopterEuler(0.1), QuadraticCost(np.eye(12), g["R"]), QuadraticTerminalCost(10 * np.eye(12)),
                                   g["x0"], rep(configs.U_TRIM, N), maxIter=10, tol=-1.0, return_log=True)
    assert log["alpha_idx"].cpu().numpy().tolist() == g["alpha_idx"].tolist()
    for b in range(g["x0"].shape[0]):
        assert relerr(log["J"][b], g["J_per_iter"][b]) < 1e-10
        assert relerr(traj.xTraj[b], g["xTraj"][b]) < 1e-9 and relerr(traj.uTraj[b], g["uTraj"][b]) < 1e-9 and relerr(L[b], g["L"][b]) < 1e-9
        assert abs(float(J[b]) - g["J_per_iter"][b][-1]) < 1e-10 * g["J_per_iter"][b][-1]


@pytest.mark.gpu
def test_gpu_simulator_goldens():
    """zopt_b200.simulator (one launch of the rollout kernel) against the reference's Simulator run unmodified on the shim:
    the tracking controller in wind (demos/iterativeLqr.py:44-56) and the 8-state LQR gains on the 12-state plant
    (demos/discreteFiniteHorizonLqr.py:38-49); reference shapes, fp64 1e-10."""
    from zopt_b200.models import QuadcopterEuler
    from zopt_b200.simulator import ProportionalFeedbackController, SimBlock, Simulator, TrackingController
    g = load("simulator_demos.npz")
    dt = float(g["a_dt"])
    N = g["a_L"].shape[0]
    sim = Simulator([SimBlock(TrackingController(g["a_xTraj"], g["a_uTraj"], g["a_L"]), np.array([]), dt=dt, name="Controller"),
                     SimBlock(QuadcopterEuler(dt, g["a_wind"]), g["a_x0"], dt=dt, name="Dynamics")], (0, N * dt))
    tS, x0A, xS, uS, _ = sim.simulate()
    assert np.allclose(tS.cpu().numpy(), g["a_t"]) and xS.shape == g["a_x"].shape and uS.shape == g["a_u"].shape
    assert relerr(xS, g["a_x"]) < 1e-10 and relerr(uS, g["a_u"]) < 1e-10
    sim = Simulator([SimBlock(ProportionalFeedbackController(np.zeros(8), g["b_uTrim"], g["b_K"], ns=8), np.array([]), dt=0.1),
                     SimBlock(QuadcopterEuler(0.1), g["b_x0"], dt=0.1)], (0, 10))
    tS, _, xS, uS, _ = sim.simulate()
    assert np.allclose(tS.cpu().numpy(), g["b_t"]) and relerr(xS, g["b_x"]) < 1e-10 and relerr(uS, g["b_u"]) < 1e-10


@pytest.mark.gpu
def test_gpu_continuous_finite_horizon_lqr():
    """zopt_b200.lqrUtils.finiteHorizonLqr (fixed-step RK4 kernel, coefficients sampled at the stage times) against the
    reference's test values, the reference run on the shim (time-varying coefficients, grid and off-grid times, clipping
    outside [0, T]) and the oracle; batched call with per-problem Qf; fp64 1e-7 (the reference's integrator runs at 1.4e-8),
    fp32 1e-4."""
    from zopt_b200.lqrUtils import finiteHorizonLqr
    I2 = lambda t: np.eye(2)
    K = finiteHorizonLqr(I2, I2, I2, I2, np.eye(2), 1, N=4)
    assert relerr(K(1), np.eye(2)) < 1e-14
    s2 = np.sqrt(2)
    K_exp = lambda t: ((1 + s2) * np.exp(2 * s2) - (s2 - 1) * np.exp(2 * s2 * t)) / (np.exp(2 * s2 * t) + np.exp(2 * s2))
    assert relerr(K(0), K_exp(0) * np.eye(2)) < 1e-8
    assert relerr(finiteHorizonLqr(np.eye(2), np.eye(2), np.eye(2), np.eye(2), np.eye(2), 1, N=4)(0), K_exp(0) * np.eye(2)) < 1e-8  # constant matrices
    g, A, B, Q, Ri = _care_problem()
    T_, N_ = float(g["T"]), int(g["N"])
    K = finiteHorizonLqr(A, B, Q, Ri, g["Qf"], T_, N=N_)
    assert K.V.shape == (N_, 4, 4) and K(0.3).shape == (2, 4)
    for tq, Kr in zip(g["tq"], g["K"]):
        assert relerr(K(tq), Kr) < 1e-7
    Ko = olqr.finiteHorizonLqr(A, B, Q, Ri, g["Qf"], T_, N=N_)
    assert relerr(K.V, Ko.V) < 1e-7
    # batched: per-problem terminal cost and a batched callable; every problem equals its own un-batched solve
    rng = np.random.default_rng(2)
    Qfb = np.stack([s * g["Qf"] for s in (0.5, 1.0, 3.0)])
    Ab = lambda t: np.stack([A(t), 0.5 * A(t), A(t) + 0.1 * np.eye(4)])
    Kb = finiteHorizonLqr(Ab, B, Q, Ri, Qfb, T_, N=N_)
    assert Kb.V.shape == (3, N_, 4, 4) and Kb(0.4).shape == (3, 2, 4)
    for i, (Ai, s) in enumerate(zip((A, lambda t: 0.5 * A(t), lambda t: A(t) + 0.1 * np.eye(4)), (0.5, 1.0, 3.0))):
        Ki = olqr.finiteHorizonLqr(Ai, B, Q, Ri, s * g["Qf"], T_, N=N_)
        assert relerr(Kb(0.4)[i], Ki(0.4)) < 1e-7
    K32 = finiteHorizonLqr(lambda t: A(t).astype(np.float32), lambda t: B(t).astype(np.float32), lambda t: Q(t).astype(np.float32),
                           lambda t: Ri(t).astype(np.float32), g["Qf"].astype(np.float32), T_, N=N_)
    assert K32.V.dtype == torch.float32 and relerr(K32(0.37), g["K"][2]) < 1e-4
    with pytest.raises(ValueError):
        finiteHorizonLqr(A, B, Q, Ri, np.eye(3), T_, N=N_)
