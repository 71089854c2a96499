"""
CPU-side checks: the C-ABI library loads and exports every symbol include/zopt_b200.h declares, the
ctypes table mirrors the header, argument errors are reported without touching a GPU, and the
product path refuses to run without CUDA (no CPU fallback).
"""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    src = open(os.path.join(ROOT, "include", "zopt_b200.h")).read()
    return re.findall(r"^ZB_API\s+\w+\s+(zb_\w+)\(", src, flags=re.M)


def test_library_exports_every_declared_symbol():
    from zopt_b200 import _lib
    names = _declared()
    assert len(names) >= 15
    for name in names:
        assert hasattr(_lib.lib, name), f"{name} declared in include/zopt_b200.h but not exported"
    assert set(names) == set(_lib.SIGNATURES), "ctypes table and header disagree"
    assert _lib.lib.zb_version() >= 100


def test_struct_layouts_match_header():
    from zopt_b200 import _lib
    assert C.sizeof(_lib.ZbArr) == 24
    assert C.sizeof(_lib.ZbModel) == 4 * 4 + 8 + 24 + 2 * 24
    assert C.sizeof(_lib.ZbCost) == 72
    assert C.sizeof(_lib.ZbAdmmOpts) == 8 + 6 * 8


def test_argument_errors_are_reported_without_a_gpu():
    from zopt_b200 import _lib
    a = _lib.ZbArr(1, 0, 0)
    rc = _lib.lib.zb_lqr_dfh(0, 0, None, 1, 2, 2, 17, 1, C.byref(a), C.byref(a), C.byref(a), C.byref(a), None, None)
    assert rc < 0 and "n must be" in _lib.last_error()
    with pytest.raises(ValueError):
        _lib.check(rc)
    rc = _lib.lib.zb_lqr_dfh(5, 0, None, 1, 2, 2, 2, 1, C.byref(a), C.byref(a), C.byref(a), C.byref(a), None, None)
    assert rc < 0 and "dtype" in _lib.last_error()
    rc = _lib.lib.zb_pd_clamp(1, 0, None, 1, 99, 1e-3, None, None)
    assert rc < 0
    m = _lib.ZbModel()
    m.kind = 7
    rc = _lib.lib.zb_ilqr_rollout(1, 0, None, 1, 1, C.byref(m), None, None, None, None, None, None, 1.0, None, None, None)
    assert rc == -2
    with pytest.raises(TypeError):
        _lib.check(rc)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
    I = np.repeat(np.eye(2)[None], 2, axis=0)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        discreteFiniteHorizonLqr(I, I, I, I, 2)


def test_unregistered_callables_raise_typeerror():
    from zopt_b200.models import LinearDynamics, QuadraticCost, require_cost, require_model
    with pytest.raises(TypeError):
        require_model(lambda x, u: x + u)
    with pytest.raises(TypeError):
        require_cost(lambda x, u: x @ x, lambda x: x @ x)
    dyn = LinearDynamics(np.eye(2), np.eye(2))
    assert require_model(dyn) is dyn
    # registered objects stay callable like the lambdas they replace
    x, u = torch.tensor([1., 2.], dtype=torch.float64), torch.tensor([3., 4.], dtype=torch.float64)
    assert torch.equal(dyn(x, u), x + u)
    assert float(QuadraticCost(np.eye(2), 2 * np.eye(2))(x, u)) == 5 + 50


def test_pytree_semantics():  # container half of reference tests/test_pytrees.py (no GPU needed)
    from zopt_b200 import pytrees
    T = lambda v: torch.as_tensor(np.asarray(v, dtype=np.float64))
    traj = pytrees.Trajectory(torch.arange(15).reshape(5, 3), torch.arange(8).reshape(4, 2))
    assert torch.equal(traj[2].xTraj, torch.arange(15).reshape(5, 3)[2])
    V = pytrees.QuadraticValueFunction(T(1.), T([2., 3]), T([[4., 5], [6, 7]]))
    assert float(V(T([8., 9]))) == pytest.approx(851.5)
    Cq = pytrees.QuadraticCostFunction(T(0.), T([1, 2]), T([2, 1]), torch.eye(2, dtype=torch.float64),
                                       torch.ones((2, 2), dtype=torch.float64), torch.eye(2, dtype=torch.float64))
    assert float(Cq(T([1., 2]), T([3., 4]))) == pytest.approx(51.0)
    dyn = pytrees.AffineDynamics(T([1, 1]), T([[2, 3], [4, 5]]), T([[6], [7]]))
    assert dyn(T([1, 2]), T([2])).numpy() == pytest.approx(np.array([21, 29]))
    qd = pytrees.QuadraticDynamics(torch.zeros(2, dtype=torch.float64), torch.eye(2, dtype=torch.float64),
                                   torch.eye(2, dtype=torch.float64), torch.stack([torch.eye(2), 2 * torch.eye(2)]).double(),
                                   torch.zeros((2, 2, 2), dtype=torch.float64),
                                   torch.stack([torch.eye(2), 2 * torch.eye(2)]).double())
    assert torch.equal(qd(T([1, 0]), T([0, 1])), T([2., 3]))
    pol = pytrees.AffinePolicy(T([1, 2]), T([[1, 2], [3, 4]]))
    assert torch.equal(pol(T([1, 2])), T([6, 13]))
    l = torch.arange(6, dtype=torch.float64).reshape(2, 3)
    L = torch.arange(12, dtype=torch.float64).reshape(2, 3, 2)
    pol = pytrees.AffinePolicy(l, L)
    assert torch.equal(pol(torch.zeros(2, dtype=torch.float64), k=0, alpha=0.5), 0.5 * l[0])
    for tree, args in ((pol, (torch.zeros(2, dtype=torch.float64),)),):
        with pytest.raises(ValueError):
            tree(*args)
    dJ = pytrees.QuadraticDeltaCost(1, 2)
    assert dJ(1) == 3 and dJ(0.5) == 1


def test_quadcopter_jacobian_structure_table_matches_generated_model():
    """csrc/ilqr_fast.cuh::quad_jx_kind (which entries of dF/dx are zero / constant / state-dependent: the backward kernel
    rewrites only the chunks of f_x that vary) must agree with the sympy-generated csrc/quad_model_gen.cuh::quad_jac_x."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gen = open(os.path.join(root, "zopt_b200", "csrc", "quad_model_gen.cuh")).read()
    a = gen.index("ZB_HD void quad_jac_x(const QuadTrig<T>& tr")
    body = gen[a:gen.index("\n}\n", a)]
    kind = {}
    for m in re.finditer(r"J\[(\d+)\] = (.*);", body):
        rhs = m.group(2).strip()
        kind[int(m.group(1))] = 0 if rhs == "T(0)" else (2 if re.search(r"\b(t\d+|x\[|u\[|tr\.)", rhs) else 1)
    assert sorted(kind) == list(range(144))
    fast = open(os.path.join(root, "zopt_b200", "csrc", "ilqr_fast.cuh")).read()
    a = fast.index("constexpr char K[12][13] = {")
    rows = re.findall(r'"([012]{12})"', fast[a:fast.index("};", a)])
    assert len(rows) == 12
    assert [int(c) for r in rows for c in r] == [kind[i] for i in range(144)]
    # 17 of the 36 four-column chunks vary along a trajectory (the count the kernel's comments and DESIGN.md quote)
    assert sum(any(kind[i * 12 + 4 * q + c] == 2 for c in range(4)) for i in range(12) for q in range(3)) == 17


def test_batch_and_shape_validation_host_side():
    """every batched operand is shared (1) or carries exactly Bsz problems, and block shapes are checked against
    (n, m), before a raw pointer can reach a kernel (pure host logic, no device needed)."""
    from zopt_b200.models import LinearDynamics, QuadraticCost, QuadraticTerminalCost, cost_batch, reconcile_batch
    assert reconcile_batch(1, 8, 1, 8) == 8 and reconcile_batch(1, 1) == 1
    with pytest.raises(ValueError):
        reconcile_batch(8, 4)
    rc, tc = QuadraticCost(np.zeros((4, 12, 12)), np.zeros((4, 4))), QuadraticTerminalCost(np.zeros((12, 12)))
    assert cost_batch(rc, tc, 12, 4) == 4
    with pytest.raises(ValueError):  # Q batched by 4, Qf by 8
        cost_batch(rc, QuadraticTerminalCost(np.zeros((8, 12, 12))), 12, 4)
    with pytest.raises(ValueError):  # R is not (m, m)
        cost_batch(QuadraticCost(np.zeros((12, 12)), np.zeros((3, 3))), tc, 12, 4)
    with pytest.raises(ValueError):  # Q is not (n, n) for the model
        cost_batch(QuadraticCost(np.zeros((8, 8)), np.zeros((4, 4))), QuadraticTerminalCost(np.zeros((8, 8))), 12, 4)
    assert LinearDynamics(np.zeros((5, 2, 2)), np.zeros((2, 1))).batch() == 5
    with pytest.raises(ValueError):
        LinearDynamics(np.zeros((5, 2, 2)), np.zeros((3, 2, 1)))
    with pytest.raises(ValueError):
        LinearDynamics(np.zeros((2, 3)), np.zeros((2, 1)))


def test_hostbind_helpers_without_gpu():
    """zopt_b200.hostbind: cpulist parsing and the topology record of the bench (pure host logic)"""
    from zopt_b200 import hostbind
    assert hostbind._parse_cpulist("0-3,8,10-11") == {0, 1, 2, 3, 8, 10, 11} and hostbind._parse_cpulist("") == set()
    topo = hostbind.host_topology()
    assert topo["cpus_allowed"] >= 1 and topo["cpu_count"] >= topo["cpus_allowed"]
    info = hostbind.bind_to_gpu(0, enable=False)
    assert info["bound"] is False and info["how"] == "unbound"
