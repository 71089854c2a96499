"""
User-defined symbolic dynamics (zopt_b200/plugin.py): sympy -> CUDA code generation and the nvcc build of the
solver plug-in (CPU, no GPU needed: nvcc cross-compiles), then on a GPU the generated code against torch autodiff of the
same expressions and iLQR / DDP / trajectoryRollout against the oracle, which differentiates the model's own callable the
way the reference does with JAX (zopt/pytrees.py:138-194).
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from tests import plugin_models


@pytest.mark.parametrize("name", ["pendulum", "car"])
def test_codegen_and_build(name):
    mdl = plugin_models.build(name)
    f, n, m = plugin_models.MODELS[name]
    assert (mdl.n, mdl.m) == (n, m)
    for fn in ("user_step", "user_lin", "user_hess"):
        assert f"ZB_HD void {fn}(" in mdl.source
    assert os.path.exists(mdl.so_path)
    lib = C.CDLL(mdl.so_path)
    import re
    hdr = open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "include", "zopt_b200_plugin.h")).read()
    declared = re.findall(r"^ZB_API\s+\w+\s+(zb_\w+)\(", hdr, flags=re.M)
    assert len(declared) == 7
    for sym in declared:  # every symbol include/zopt_b200_plugin.h declares is exported by the built plug-in
        assert hasattr(lib, sym), sym
    # the object is callable like the lambda it replaces, and differentiable (this is what the oracle uses)
    x, u = torch.linspace(0.1, 0.4, n, dtype=torch.float64), torch.linspace(-0.3, 0.2, m, dtype=torch.float64)
    xn = mdl(x, u)
    assert xn.shape == (n,) and xn.dtype == torch.float64
    fx = torch.func.jacrev(mdl, argnums=0)(x, u)
    assert fx.shape == (n, n) and torch.isfinite(fx).all()
    xb = x[None].repeat(3, 1)
    assert torch.allclose(mdl(xb, u[None].repeat(3, 1))[1], xn)
    # a model that is not a function of (x, u) only, or too large, is refused up front
    import sympy as sp
    from zopt_b200.plugin import SymbolicDynamics
    with pytest.raises(ValueError):
        SymbolicDynamics(lambda x, u: [x[0] + sp.Symbol("k")], 1, 1, build=False)
    with pytest.raises(ValueError):
        SymbolicDynamics(lambda x, u: list(x), 17, 1, build=False)
    # no CPU fallback: solving without a CUDA device raises
    if not torch.cuda.is_available():
        from zopt_b200 import ilqrUtils
        from zopt_b200.models import QuadraticCost, QuadraticTerminalCost
        with pytest.raises(RuntimeError):
            ilqrUtils.iterativeLqr(mdl, QuadraticCost(np.eye(n), np.eye(m)), QuadraticTerminalCost(np.eye(n)), np.zeros(n), np.zeros((5, m)))


def _relerr(a, b):
    a = a.detach().cpu().numpy().astype(np.float64) if isinstance(a, torch.Tensor) else np.asarray(a, dtype=np.float64)
    b = b.detach().cpu().numpy().astype(np.float64) if isinstance(b, torch.Tensor) else np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pendulum", "car"])
@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_generated_model_vs_autodiff(name, dt):
    mdl = plugin_models.build(name)
    rng = np.random.default_rng(3)
    x, u = rng.normal(size=(37, mdl.n)), rng.normal(size=(37, mdl.m))
    xn, fx, fu = mdl.step(torch.as_tensor(x, dtype=dt, device="cuda"), torch.as_tensor(u, dtype=dt, device="cuda"), linearize=True)
    xt, ut = torch.as_tensor(x), torch.as_tensor(u)
    ref = mdl(xt, ut)
    Jx, Ju = torch.func.vmap(torch.func.jacrev(mdl, argnums=(0, 1)))(xt, ut)
    tol = 1e-12 if dt == torch.float64 else 2e-6
    assert _relerr(xn, ref) < tol and _relerr(fx, Jx) < tol and _relerr(fu, Ju) < tol


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pendulum", "car"])
@pytest.mark.parametrize("second_order", [False, True])
def test_solvers_with_symbolic_model_vs_oracle(name, second_order):
    """iLQR / DDP on a user-defined model, fp64: step-size sequence first, then x, u, L, J at 1e-10 against the oracle, whose
    derivatives come from torch autodiff of the same callable (DDP exercises the full (n+m)^2 second-order block)."""
    from oracle import ilqr as oilqr
    from zopt_b200 import ilqrUtils
    from zopt_b200.models import QuadraticCost, QuadraticTerminalCost
    mdl = plugin_models.build(name)
    n, m, N, Bsz, iters = mdl.n, mdl.m, 25, 5, 3
    rng = np.random.default_rng(11 + int(second_order))
    x0 = rng.uniform(-1, 1, (Bsz, n))
    uG = 0.1 * rng.normal(size=(N, m))
    Q, R, Qf = np.eye(n), 0.5 * np.eye(m), 10 * np.eye(n)
    solver = ilqrUtils.differentialDynamicProgramming if second_order else ilqrUtils.iterativeLqr
    traj, L, J, conv, log = solver(mdl, QuadraticCost(Q, R), QuadraticTerminalCost(Qf), torch.as_tensor(x0, device="cuda"), uG,
                                   maxIter=iters, tol=-1.0, return_log=True)
    assert traj.xTraj.shape == (Bsz, N + 1, n) and L.shape == (Bsz, N, m, n)
    Qt, Rt, Qft = torch.as_tensor(Q), torch.as_tensor(R), torch.as_tensor(Qf)
    rc, tc = (lambda x, u: x @ Qt @ x + u @ Rt @ u), (lambda x: x @ Qft @ x)
    osolver = oilqr.differentialDynamicProgramming if second_order else oilqr.iterativeLqr
    for b in range(Bsz):
        olog = []
        tr, Lr, Jr, _ = osolver(mdl, rc, tc, torch.as_tensor(x0[b]), torch.as_tensor(uG), maxIter=iters, tol=-1.0, log=olog)
        assert [e["alpha_idx"] for e in olog[1:]] == log["alpha_idx"][b].tolist()
        assert _relerr(traj.xTraj[b], tr.xTraj) < 1e-10 and _relerr(traj.uTraj[b], tr.uTraj) < 1e-10
        assert _relerr(L[b], Lr) < 1e-10 and abs(float(J[b]) - float(Jr)) < 1e-10 * abs(float(Jr))
    # un-batched call: reference shapes; trajectoryRollout with the same model reproduces the solver's trajectory
    t1, L1, J1, c1 = solver(mdl, QuadraticCost(Q, R), QuadraticTerminalCost(Qf), torch.as_tensor(x0[0], device="cuda"), uG, maxIter=iters, tol=-1.0)
    assert t1.xTraj.shape == (N + 1, n) and torch.equal(t1.xTraj, traj.xTraj[0])
    from zopt_b200.pytrees import AffinePolicy, Trajectory
    zero = torch.zeros((N, m), dtype=torch.float64, device="cuda")
    t2 = ilqrUtils.trajectoryRollout(torch.as_tensor(x0[0], device="cuda"), mdl, AffinePolicy(zero, L1), Trajectory(t1.xTraj, t1.uTraj), alpha=1)
    assert _relerr(t2.xTraj, t1.xTraj) < 1e-12 and _relerr(t2.uTraj, t1.uTraj) < 1e-12
    with pytest.raises(TypeError):
        solver(lambda x, u: x, QuadraticCost(Q, R), QuadraticTerminalCost(Qf), x0[0], uG)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pendulum", "car"])
def test_pytree_constructors_with_symbolic_model(name):
    """AffineDynamics / QuadraticDynamics .from_function / .from_trajectory (zopt/pytrees.py:138-153, 179-194) for a
    user-defined model: the generated derivatives against torch autodiff of the model's callable (what JAX does in the
    reference), then the DDP backward pass on those explicit pytrees against the oracle's."""
    from oracle import ilqr as oilqr
    from oracle import pytrees as opt
    from zopt_b200 import ilqrUtils
    from zopt_b200.pytrees import AffineDynamics, QuadraticDynamics, Trajectory
    mdl = plugin_models.build(name)
    n, m, N = mdl.n, mdl.m, 6
    rng = np.random.default_rng(2)
    xT, uT = rng.normal(size=(N + 1, n)), rng.normal(size=(N, m))
    qd = QuadraticDynamics.from_trajectory(mdl, Trajectory(torch.as_tensor(xT, device="cuda"), torch.as_tensor(uT, device="cuda")))
    assert qd.f_xx.shape == (N, n, n, n) and qd.f_ux.shape == (N, n, m, n) and qd.f_uu.shape == (N, n, m, m)
    xt, ut = torch.as_tensor(xT[:-1]), torch.as_tensor(uT)
    Jx, Ju = torch.func.vmap(torch.func.jacrev(mdl, argnums=(0, 1)))(xt, ut)
    (Hxx, Hxu), (Hux, Huu) = torch.func.vmap(torch.func.hessian(mdl, argnums=(0, 1)))(xt, ut)
    assert _relerr(qd.f, mdl(xt, ut)) < 1e-12 and _relerr(qd.f_x, Jx) < 1e-12 and _relerr(qd.f_u, Ju) < 1e-12
    assert _relerr(qd.f_xx, Hxx) < 1e-12 and _relerr(qd.f_uu, Huu) < 1e-12
    assert _relerr(qd.f_ux, Hux) < 1e-12 or float(Hux.abs().max()) == 0.0
    ad = AffineDynamics.from_function(mdl, torch.as_tensor(xT[0], device="cuda"), torch.as_tensor(uT[0], device="cuda"))
    assert ad.f_x.shape == (n, n) and _relerr(ad.f_x, Jx[0]) < 1e-12


@pytest.mark.parametrize("name", ["pendulum", "car"])
def test_symbolic_cost_codegen_and_build(name):
    """plugin.SymbolicCost: generated cost code compiled with the model into one plug-in (nvcc, no GPU needed); the cost parts
    are callable on torch tensors like the reference's lambdas, and mismatched pairs are refused."""
    from zopt_b200.models import symbolic_cost_of
    from zopt_b200.plugin import SymbolicCost
    mdl, bound, cost = plugin_models.build_with_cost(name)
    assert bound is mdl.with_cost(cost) and os.path.exists(bound.so_path) and bound.so_path != mdl.so_path
    for fn in ("user_cost", "user_cost_grad", "user_cost_hess", "user_tcost", "user_tcost_grad", "user_tcost_hess"):
        assert f" {fn}(" in bound.source
    assert "#define ZB_USER_COST 1" in bound.source and "ZB_USER_COST" not in mdl.source
    x, u = torch.linspace(0.1, 0.4, mdl.n, dtype=torch.float64), torch.linspace(-0.3, 0.2, mdl.m, dtype=torch.float64)
    c, cf = cost.running(x, u), cost.terminal(x)
    assert c.shape == () and cf.shape == () and torch.isfinite(c) and torch.isfinite(cf)
    g = torch.func.grad(cost.running, argnums=0)(x, u)
    assert g.shape == (mdl.n,)
    assert symbolic_cost_of(cost.running, cost.terminal) is cost
    other = SymbolicCost(lambda x, u: x[0] ** 2 + u[0] ** 2, lambda x: x[0] ** 2, mdl.n, mdl.m)
    with pytest.raises(TypeError):
        symbolic_cost_of(cost.running, other.terminal)
    with pytest.raises(ValueError):
        mdl.with_cost(SymbolicCost(lambda x, u: x[0] ** 2, lambda x: x[0] ** 2, mdl.n + 1, mdl.m))


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["pendulum", "car"])
@pytest.mark.parametrize("second_order", [False, True])
def test_solvers_with_symbolic_cost_vs_oracle(name, second_order):
    """iLQR / DDP with a user-defined NON-quadratic cost (and model), fp64, against the oracle, which differentiates the cost's
    own callables with torch autodiff the way the reference does with JAX (pytrees.py:99-115, 71-81) and eigen-clamps the
    stacked cost Hessian at every time step (ilqrUtils.py:222-234): step-size sequence, then x, u, L, J."""
    from oracle import ilqr as oilqr
    from zopt_b200 import ilqrUtils
    from zopt_b200.models import QuadraticCost, QuadraticTerminalCost
    mdl, bound, cost = plugin_models.build_with_cost(name)
    n, m, N, Bsz, iters = mdl.n, mdl.m, 20, 4, 3
    rng = np.random.default_rng(19 + int(second_order))
    x0 = rng.uniform(-1, 1, (Bsz, n))
    uG = 0.1 * rng.normal(size=(N, m))
    solver = ilqrUtils.differentialDynamicProgramming if second_order else ilqrUtils.iterativeLqr
    traj, L, J, conv, log = solver(mdl, cost.running, cost.terminal, torch.as_tensor(x0, device="cuda"), uG, maxIter=iters, tol=-1.0,
                                   return_log=True)
    osolver = oilqr.differentialDynamicProgramming if second_order else oilqr.iterativeLqr
    for b in range(Bsz):
        olog = []
        tr, Lr, Jr, _ = osolver(mdl, cost.running, cost.terminal, torch.as_tensor(x0[b]), torch.as_tensor(uG), maxIter=iters, tol=-1.0, log=olog)
        assert [e["alpha_idx"] for e in olog[1:]] == log["alpha_idx"][b].tolist()
        assert _relerr(log["J"][b], np.array([e["J"] for e in olog])) < 1e-10
        assert _relerr(traj.xTraj[b], tr.xTraj) < 1e-9 and _relerr(traj.uTraj[b], tr.uTraj) < 1e-9
        assert _relerr(L[b], Lr) < 1e-8 and abs(float(J[b]) - float(Jr)) < 1e-10 * abs(float(Jr))
    # a symbolic cost needs a symbolic model (it is compiled together with it)
    from zopt_b200.models import QuadcopterEuler
    with pytest.raises(TypeError):
        solver(QuadcopterEuler(0.1), cost.running, cost.terminal, np.zeros(12), np.zeros((N, 4)))
    with pytest.raises(TypeError):
        solver(mdl, cost.running, QuadraticTerminalCost(np.eye(n)), x0[0], uG)
