"""
Out-of-bounds write check without compute-sanitizer (closed on this pool): every output and workspace buffer of the
fast kernels is carved out of a larger allocation whose head and tail guard zones are filled with a sentinel; after
the call through the raw C ABI the guards must be untouched.  Ragged batch sizes exercise the partial last CTA / warp.
"""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from zopt_b200 import _lib, configs  # noqa: E402
from zopt_b200._lib import View, ZbAdmmOpts, check, lib, ptr, stream_ptr  # noqa: E402

GUARD = 4096  # elements
SENT = 12345.0


class Guarded:
    def __init__(self, shape, dtype, dev):
        n = int(np.prod(shape))
        self.raw = torch.full((n + 2 * GUARD,), SENT, dtype=torch.float64, device=dev).to(dtype) if dtype != torch.uint8 \
            else torch.full((n + 2 * GUARD,), 77, dtype=torch.uint8, device=dev)
        self.sent = self.raw[0].clone()
        self.t = self.raw[GUARD:GUARD + n].view(shape)

    def ok(self):
        return bool((self.raw[:GUARD] == self.sent).all()) and bool((self.raw[-GUARD:] == self.sent).all())


def _cfg(Bsz, dev, dt=torch.float32):
    from zopt_b200.quadcopter import Quadcopter
    d = configs.cfg2(Bsz=Bsz)
    xbar = torch.as_tensor(d["xbar"], dtype=dt, device=dev)
    ubar = torch.as_tensor(d["ubar"], dtype=dt, device=dev)
    A, B = Quadcopter().linearizeInertial(xbar, ubar, 0.1)
    Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=dt, device=dev))
    R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=dt, device=dev))
    return d, xbar, A, B, Q, R


@pytest.mark.parametrize("Bsz", [1, 33, 77])
@pytest.mark.parametrize("diag", [True, False])
@pytest.mark.parametrize("f32", [torch.float32, torch.float64])  # fp64: cooperative Riccati + quad plan rollout (csrc/lqr_quad64.cuh)
def test_mpc_and_closed_loop_guards(Bsz, diag, f32):
    dev = torch.device("cuda", 0)
    code = 0 if f32 == torch.float32 else 1
    d, xbar, A, B, Q, R = _cfg(Bsz, dev, f32)
    if not diag:
        Q = Q + 0.01
        R = R + 0.01
    N = 17
    Qf = (10 * Q).contiguous()
    views = [View(t, 2, False, True) for t in (A, B, Q, R, Qf)]
    inf = [View(torch.full((k,), s * float("inf"), dtype=f32, device=dev), 1, False, False) for k, s in ((12, -1), (12, 1), (4, -1), (4, 1))]
    u0, xT, uT = Guarded((Bsz, 4), f32, dev), Guarded((Bsz, N + 1, 12), f32, dev), Guarded((Bsz, N, 4), f32, dev)
    st, it = Guarded((Bsz,), torch.uint8, dev), Guarded((Bsz,), torch.float32, dev)
    wsb = lib.zb_mpc_workspace_bytes(code, Bsz, N, 12, 4)
    ws = Guarded((wsb,), torch.uint8, dev)
    opts = ZbAdmmOpts(4000, 25, 0.1, 1e-6, 1.6, 1e-3, 1e-3, 1e-4)
    check(lib.zb_mpc_lqr_solve(code, 0, stream_ptr(dev), Bsz, N, 12, 4, *[v.ref() for v in views], *[v.ref() for v in inf],
                               2 if diag else 0, ptr(xbar), C.byref(opts), ptr(u0.t), ptr(xT.t), ptr(uT.t), ptr(st.t), ptr(it.t),
                               ptr(ws.t), wsb))
    torch.cuda.synchronize()
    assert all(g.ok() for g in (u0, xT, uT, st, it, ws))
    assert torch.isfinite(xT.t).all() and torch.isfinite(uT.t).all() and (st.t == 0).all()
    # closed loop
    Ts = 5
    xS, uS = Guarded((Bsz, Ts + 1, 12), f32, dev), Guarded((Bsz, Ts, 4), f32, dev)
    ut = (C.c_double * 4)(*configs.U_TRIM)
    x0 = xbar.clone()
    x0[:, 9:12] *= 0.2
    for variant in ((4, 8) if code == 0 else (0,)):  # fp32: thread-per-problem, quad; fp64: the cooperative kernel
        check(lib.zb_mpc_closed_loop_quad(code, 0, stream_ptr(dev), Bsz, N, Ts, 0.1, ut, views[2].ref(), views[3].ref(), views[4].ref(),
                                          (2 if diag else 0) | variant, ptr(x0), ptr(xS.t), ptr(uS.t)))
        torch.cuda.synchronize()
        assert xS.ok() and uS.ok() and torch.isfinite(xS.t).all()


@pytest.mark.parametrize("Bsz", [1, 45])
@pytest.mark.parametrize("mode", ["views", "q_series", "materialised"])
@pytest.mark.parametrize("f32", [torch.float32, torch.float64])  # fp64: the cooperative kernel of csrc/lqr_quad64.cuh
def test_lqr_dfh_guards(Bsz, mode, f32):
    dev = torch.device("cuda", 0)
    d, xbar, A, B, Q, R = _cfg(Bsz, dev, f32)
    N = 9
    ex = lambda t: t[:, None].expand(-1, N, -1, -1)
    Ak, Bk, Qk, Rk = ex(A), ex(B), ex(Q), ex(R)
    if mode != "views":
        Qk = Qk.contiguous()
    if mode == "materialised":
        Ak, Bk, Rk = Ak.contiguous(), Bk.contiguous(), Rk.contiguous()
    views = [View(t, 2, True, True) for t in (Ak, Bk, Qk, Rk)]
    L, V0 = Guarded((Bsz, N, 4, 12), f32, dev), Guarded((Bsz, 12, 12), f32, dev)
    check(lib.zb_lqr_dfh(0 if f32 == torch.float32 else 1, 0, stream_ptr(dev), Bsz, N, N, 12, 4, *[v.ref() for v in views], ptr(L.t), ptr(V0.t)))
    torch.cuda.synchronize()
    assert L.ok() and V0.ok() and torch.isfinite(L.t).all() and torch.isfinite(V0.t).all()


@pytest.mark.parametrize("second_order", [0, 1])
@pytest.mark.parametrize("dt", [torch.float64, torch.float32])
def test_ilqr_solve_guards(second_order, dt):
    from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost, cost_spec
    dev = torch.device("cuda", 0)
    Bsz, N, maxIter = 21, 11, 2
    d = configs.cfg4(Bsz=Bsz, N=N)
    x0 = torch.as_tensor(d["x0"], dtype=dt, device=dev)
    uG = torch.as_tensor(d["uGuess"], dtype=dt, device=dev)[None].expand(Bsz, N, 4).contiguous()
    mspec, _ = QuadcopterEuler(0.1).spec(dt, dev)
    cspec, keep = cost_spec(QuadraticCost(d["Q"], d["R"]), QuadraticTerminalCost(d["Qf"]), dt, dev)
    xT, uT, L = Guarded((Bsz, N + 1, 12), dt, dev), Guarded((Bsz, N, 4), dt, dev), Guarded((Bsz, N, 4, 12), dt, dev)
    J, conv = Guarded((Bsz,), dt, dev), Guarded((Bsz,), torch.uint8, dev)
    iters = torch.zeros((Bsz,), dtype=torch.int32, device=dev)
    code = 0 if dt == torch.float32 else 1
    wsb = lib.zb_ilqr_workspace_bytes(code, Bsz, N, 12, 4)
    ws = Guarded((wsb,), torch.uint8, dev)
    check(lib.zb_ilqr_solve(code, 0, stream_ptr(dev), Bsz, N, second_order | 2, C.byref(mspec), C.byref(cspec), ptr(x0), ptr(uG),
                            maxIter, -1.0, ptr(xT.t), ptr(uT.t), ptr(L.t), ptr(J.t), ptr(conv.t), ptr(iters), None, None, ptr(ws.t), wsb))
    torch.cuda.synchronize()
    assert all(g.ok() for g in (xT, uT, L, J, conv, ws))
    assert torch.isfinite(xT.t).all() and torch.isfinite(L.t).all() and (iters == maxIter).all()


def test_python_mirror_rejects_inconsistent_batches_and_shapes():
    """operands whose batch size is neither 1 nor Bsz, short time axes and wrong block shapes raise ValueError in
    the Python mirror instead of reaching a kernel as an out-of-bounds read."""
    import numpy as np
    import torch
    from zopt_b200 import ilqrUtils
    from zopt_b200.lqrUtils import bilinearAffineLqr
    from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost
    from zopt_b200.mpcUtils import lqrMpc, quadcopterClosedLoopMpc
    x0 = torch.zeros((8, 12), dtype=torch.float64, device="cuda")
    uG = np.tile(np.array([9.807, 0, 0, 0]), (5, 1))
    with pytest.raises(ValueError):  # x0 batch 8, Q batch 4
        ilqrUtils.iterativeLqr(QuadcopterEuler(0.1), QuadraticCost(np.tile(np.eye(12), (4, 1, 1)), np.eye(4)), QuadraticTerminalCost(np.eye(12)),
                               x0, uG, maxIter=1)
    with pytest.raises(ValueError):  # R is 3x3
        ilqrUtils.iterativeLqr(QuadcopterEuler(0.1), QuadraticCost(np.eye(12), np.eye(3)), QuadraticTerminalCost(np.eye(12)), x0, uG, maxIter=1)
    n, m, N = 3, 2, 6
    ops = dict(A=np.zeros((N, n, n)), B=np.zeros((N, n, m)), d=np.zeros((N, n)), Q=np.tile(np.eye(n), (N, 1, 1)), R=np.tile(np.eye(m), (N, 1, 1)),
               H=np.zeros((N, m, n)), q=np.zeros((N, n)), r=np.zeros((N, m)), q0=np.zeros(N))
    bilinearAffineLqr(*ops.values(), N)
    for bad in ("A", "B", "d", "R", "H", "r"):  # time axis shorter than the horizon
        o = dict(ops)
        o[bad] = o[bad][:N - 2]
        with pytest.raises(ValueError):
            bilinearAffineLqr(*o.values(), N)
    o = dict(ops)
    o["H"] = np.zeros((N, n, m))  # transposed block
    with pytest.raises(ValueError):
        bilinearAffineLqr(*o.values(), N)
    with pytest.raises(ValueError):  # x0 is not a 12-state
        quadcopterClosedLoopMpc(torch.zeros((4, 8), device="cuda"), np.eye(12), np.eye(4), 5, 2)
    with pytest.raises(ValueError):  # Q batch 3, x0 batch 4
        quadcopterClosedLoopMpc(torch.zeros((4, 12), device="cuda"), np.tile(np.eye(12), (3, 1, 1)), np.eye(4), 5, 2)
    tr = quadcopterClosedLoopMpc(np.zeros((2, 12)), np.eye(12), np.eye(4), 5, 2)  # NumPy float64 input -> fp64, like everywhere else
    assert tr.xTraj.dtype == torch.float64
    inf12, inf4 = np.full(12, np.inf), np.full(4, np.inf)
    with pytest.raises(ValueError):
        lqrMpc(np.eye(12), np.zeros((12, 4)), np.eye(12), np.eye(3), 5, -inf12, inf12, -inf4, inf4)
    with pytest.raises(ValueError):
        lqrMpc(np.eye(12), np.zeros((12, 4)), np.eye(12), np.eye(4), 5, -inf12, inf12, -inf4, inf4).solve(np.zeros(8))


def test_non_symmetric_weights_take_the_generic_kernels():
    """the (12,4) fast kernels read the lower triangle of Q, R; the reference uses the weights as given
    (lqrUtils.py:168-169).  With a non-symmetric Q the Python mirror must select the generic kernel: gains equal the oracle's
    as-written recursion, and differ from what the lower triangle alone would give."""
    import numpy as np
    import torch
    from oracle import lqr as olqr
    from oracle.quadcopter import Quadcopter as OQ
    from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
    rng = np.random.default_rng(5)
    N = 12
    A, B = (t.numpy() for t in OQ().linearizeInertial(np.zeros(12), np.array([9.807, 0, 0, 0]), 0.1))
    Q = np.eye(12) + 0.3 * np.triu(rng.normal(size=(12, 12)), 1)  # upper triangle only: invisible to a lower-triangle reader
    R = np.eye(4)
    rep = lambda M: np.tile(M[None], (N, 1, 1))
    Lref = olqr.discreteFiniteHorizonLqr(rep(A), rep(B), rep(Q), rep(R), N)
    Llow = olqr.discreteFiniteHorizonLqr(rep(A), rep(B), rep(np.tril(Q) + np.tril(Q, -1).T), rep(R), N)
    assert np.max(np.abs(Lref - Llow)) > 1e-3
    for dt, tol in ((torch.float64, 1e-10), (torch.float32, 2e-5)):
        c = lambda M: torch.as_tensor(rep(M), dtype=dt, device="cuda")
        L = discreteFiniteHorizonLqr(c(A), c(B), c(Q), c(R), N)
        assert np.max(np.abs(L.double().cpu().numpy() - Lref)) < tol * np.max(np.abs(Lref))
