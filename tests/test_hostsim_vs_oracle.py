"""
CPU pre-check of the CUDA kernels' arithmetic: the per-problem bodies of the generic kernels
(zopt_b200/csrc/zb_problems.cuh), compiled for the host by tests/hostsim, against the oracle.
The same comparisons run on the real kernels in tests/test_gpu_*.py (-m gpu).
Tolerances: per-array max-norm relative error, fp64 1e-10, fp32 1e-5.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import ilqr as oilqr
from oracle import lqr as olqr
from oracle import pytrees as opt
from oracle.quadcopter import Quadcopter
from tests import hostsim as H
from zopt_b200 import configs

torch.set_default_dtype(torch.float64)


def relerr(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    return float(np.max(np.abs(a - b)) / (den if den > 0 else 1.0))


TOL = {np.float64: 1e-10, np.float32: 1e-5}


def quad_linearized(xbar, ubar, dt=0.1):
    ac = Quadcopter()
    f = ac.eulerStep(dt)
    A, B = torch.func.vmap(torch.func.jacrev(f, argnums=(0, 1)))(torch.as_tensor(xbar), torch.as_tensor(ubar))
    return A.numpy(), B.numpy()


def run_lqr(A, B, Q, R, N, dt):
    A, B, Q, R = (np.ascontiguousarray(a, dtype=dt) for a in (A, B, Q, R))
    Bsz, T, n, m = A.shape[0], Q.shape[1], B.shape[-2], B.shape[-1]
    L = np.zeros((Bsz, N, m, n), dtype=dt)
    V0 = np.zeros((Bsz, n, n), dtype=dt)
    zs = [H.arr(a, 2) for a in (A, B, Q, R)]
    H.hs.hs_lqr_dfh(int(dt == np.float64), C.c_int64(Bsz), N, T, n, m, *[C.byref(z) for z in zs], H.P(L), H.P(V0))
    return L, V0


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_lqr_known_answer(dt):  # reference tests/test_lqrUtils.py:61-69
    N = 2
    I = np.repeat(np.eye(2)[None, None], N, axis=1)
    L, _ = run_lqr(I, I, I, I, N, dt)
    assert L[0, 1] == pytest.approx(0.5 * np.eye(2))
    assert L[0, 0] == pytest.approx(0.6 * np.eye(2))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_lqr_cfg2_subset(dt):
    d = configs.cfg2(Bsz=32)
    A, B = quad_linearized(d["xbar"], d["ubar"])
    N = d["N"]
    Q = np.repeat(configs.diag_embed(d["qdiag"])[:, None], N + 1, axis=1)
    Q[:, N] *= 10
    R = np.repeat(configs.diag_embed(d["rdiag"])[:, None], N, axis=1)
    Ak = np.repeat(A[:, None], N, axis=1)
    Bk = np.repeat(B[:, None], N, axis=1)
    Lref, Vref = olqr.discreteFiniteHorizonLqr_batched(Ak, Bk, Q, R, N, return_value=True)
    L, V0 = run_lqr(Ak, Bk, Q, R, N, dt)
    for b in range(A.shape[0]):
        assert relerr(L[b], Lref[b]) < TOL[dt]
        assert relerr(V0[b], Vref[b]) < TOL[dt]


def test_lqr_double_integrator_and_time_varying():
    Ak, Bk, Qk, Rk, N = configs.cfg1_double_integrator()
    Lref = olqr.discreteFiniteHorizonLqr(Ak, Bk, Qk, Rk, N)
    L, _ = run_lqr(Ak[None], Bk[None], Qk[None], Rk[None], N, np.float64)
    assert relerr(L[0], Lref) < 1e-10
    # genuinely time-varying random problem, n=5, m=3, T > N (only rows 0..N-1 and -1 of Q are read)
    rng = np.random.default_rng(0)
    n, m, N, T = 5, 3, 7, 9
    A = rng.normal(size=(T, n, n)) * 0.5
    B = rng.normal(size=(T, n, m))
    Q = np.stack([(lambda M: M @ M.T + np.eye(n))(rng.normal(size=(n, n))) for _ in range(T)])
    R = np.stack([(lambda M: M @ M.T + np.eye(m))(rng.normal(size=(m, m))) for _ in range(T)])
    Lref = olqr.discreteFiniteHorizonLqr(A, B, Q, R, N)
    L, _ = run_lqr(A[None], B[None], Q[None], R[None], N, np.float64)
    assert relerr(L[0], Lref) < 1e-10


def test_bilinear():
    # known answer: reference tests/test_lqrUtils.py:82-98
    N = 2
    I = np.repeat(np.eye(2)[None, None], N, axis=1)
    one = np.ones((1, N, 2))
    L = np.zeros((1, N, 2, 2))
    l = np.zeros((1, N, 2))
    ops = [(I, 2), (I, 2), (one, 1), (I, 2), (I, 2), (I, 2), (one, 1), (one, 1), (np.ones((1, N)), 0)]
    zs = [H.arr(np.ascontiguousarray(a), k) for a, k in ops]
    H.hs.hs_lqr_bilinear(1, C.c_int64(1), N, N, 2, 2, *[C.byref(z) for z in zs], H.P(L), H.P(l))
    assert L[0, 1] == pytest.approx(np.eye(2)) and L[0, 0] == pytest.approx(np.eye(2))
    assert l[0, 1] == pytest.approx(1.5 * np.ones(2)) and l[0, 0] == pytest.approx(np.ones(2))
    # random problem vs oracle
    rng = np.random.default_rng(1)
    n, m, N, Bsz = 8, 4, 20, 3
    A = rng.normal(size=(Bsz, N, n, n)) * 0.3
    Bm = rng.normal(size=(Bsz, N, n, m))
    d = rng.normal(size=(Bsz, N, n)) * 0.1
    Q = np.stack([[np.eye(n) * (1 + rng.uniform()) for _ in range(N)] for _ in range(Bsz)])
    R = np.stack([[np.eye(m) * (1 + rng.uniform()) for _ in range(N)] for _ in range(Bsz)])
    Hm = 0.2 * rng.normal(size=(Bsz, N, m, n))
    q = rng.normal(size=(Bsz, N, n))
    r = rng.normal(size=(Bsz, N, m))
    q0 = rng.normal(size=(Bsz, N))
    Lr, lr = olqr.bilinearAffineLqr_batched(A, Bm, d, Q, R, Hm, q, r, q0, N)
    for b in range(Bsz):
        L1, l1 = olqr.bilinearAffineLqr(A[b], Bm[b], d[b], Q[b], R[b], Hm[b], q[b], r[b], q0[b], N)
        assert relerr(Lr[b], L1) < 1e-12 and relerr(lr[b], l1) < 1e-12
    L = np.zeros((Bsz, N, m, n))
    l = np.zeros((Bsz, N, m))
    zs = [H.arr(a, k) for a, k in ((A, 2), (Bm, 2), (d, 1), (Q, 2), (R, 2), (Hm, 2), (q, 1), (r, 1), (q0, 0))]
    H.hs.hs_lqr_bilinear(1, C.c_int64(Bsz), N, N, n, m, *[C.byref(z) for z in zs], H.P(L), H.P(l))
    assert relerr(L, Lr) < 1e-10 and relerr(l, lr) < 1e-10


@pytest.mark.parametrize("wind", [None, (3.0, 1.0, 0.0)])
@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_quadcopter_model_vs_autodiff(dt, wind):
    """analytic (sympy-generated) F, dF/dx, dF/du and costate-contracted Hessian vs torch autodiff of the oracle"""
    rng = np.random.default_rng(7)
    Bsz = 16
    x = configs.quad_states(rng, Bsz)
    u = np.tile(configs.U_TRIM, (Bsz, 1)) + rng.normal(size=(Bsz, 4))
    lam = rng.normal(size=(Bsz, 12))
    ac = Quadcopter()
    wt = None if wind is None else torch.tensor(wind)
    F = lambda xx, uu: ac.inertialDynamics(xx, uu, wt)
    xt, ut = torch.as_tensor(x), torch.as_tensor(u)
    Fref = torch.func.vmap(F)(xt, ut).numpy()
    Jx, Ju = (t.numpy() for t in torch.func.vmap(torch.func.jacrev(F, argnums=(0, 1)))(xt, ut))
    Hx = torch.func.vmap(torch.func.hessian(F, argnums=0))(xt, ut).numpy()  # (B,12,12,12)
    Href = np.einsum('bi,bijk->bjk', lam, Hx)
    xd = np.zeros((Bsz, 12), dtype=dt)
    A = np.zeros((Bsz, 12, 12), dtype=dt)
    Bm = np.zeros((Bsz, 12, 4), dtype=dt)
    Hh = np.zeros((Bsz, 12, 12), dtype=dt)
    w = (C.c_double * 3)(*(wind or (0, 0, 0)))
    xc, uc, lc = x.astype(dt), u.astype(dt), lam.astype(dt)  # keep alive across the calls
    H.hs.hs_quad(int(dt == np.float64), C.c_int64(Bsz), H.P(xc), H.P(uc), w, C.c_double(0.0),
                 H.P(lc), H.P(xd), H.P(A), H.P(Bm), H.P(Hh))
    tol = 1e-12 if dt == np.float64 else 2e-6
    assert relerr(xd, Fref) < tol and relerr(A, Jx) < tol and relerr(Bm, Ju) < tol and relerr(Hh, Href) < tol
    # discrete form used by the solvers: I + dt*J, dt*B
    H.hs.hs_quad(int(dt == np.float64), C.c_int64(Bsz), H.P(xc), H.P(uc), w, C.c_double(0.1),
                 H.P(lc), None, H.P(A), H.P(Bm), H.P(Hh))
    assert relerr(A, np.eye(12) + 0.1 * Jx) < tol and relerr(Bm, 0.1 * Ju) < tol and relerr(Hh, 0.1 * Href) < tol
    # structure facts of SURVEY 8a-a11
    assert np.count_nonzero(Ju[0]) == 4 and np.all(Hx[:, :, 9:, :] == 0)


def test_quadcopter_known_answers():  # reference tests/test_quadcopter.py:46-86
    x = np.zeros((3, 12))
    u = np.tile(configs.U_TRIM, (3, 1))
    u[0] = 0
    x[2, 0:3] = [0.1, 0.2, 0.3]
    x[2, 8] = np.pi / 2
    xd = np.zeros((3, 12))
    H.hs.hs_quad(1, C.c_int64(3), H.P(x), H.P(u), None, C.c_double(0.0), None, H.P(xd), None, None, None)
    assert xd[0] == pytest.approx(np.array([0, 0, 9.807] + [0] * 9))
    assert xd[1] == pytest.approx(np.zeros(12))
    assert xd[2, 9:] == pytest.approx(np.array([-0.2, 0.1, 0.3]))


@pytest.mark.parametrize("p", [1, 4, 12, 16])
def test_pd_clamp(p):
    rng = np.random.default_rng(p)
    Bsz = 8
    S = rng.normal(size=(Bsz, p, p))
    S = S + np.swapaxes(S, 1, 2)
    S[0] = 0  # all-zero block -> eps * I
    if p >= 4:
        S[1] = np.diag(np.arange(p) - 1.0)  # already diagonal with negative / zero entries
    out = np.zeros_like(S)
    H.hs.hs_pd_clamp(1, C.c_int64(Bsz), p, C.c_double(1e-3), H.P(S), H.P(out))
    for b in range(Bsz):
        ref = oilqr.ensurePositiveDefinite(torch.as_tensor(S[b])).numpy()
        assert relerr(out[b], ref) < 1e-11
    assert out[0] == pytest.approx(1e-3 * np.eye(p))


def _stack_views(tree, blocks):
    return [H.arr(np.ascontiguousarray(t), k) for t, k in zip(tree, blocks)]


@pytest.mark.parametrize("second_order", [0, 1])
def test_backward_pass_explicit_pytrees(second_order):
    """zb_ilqr_backward body vs oracle backwardPass_ilqr / backwardPass_ddp on random stacked pytrees"""
    rng = np.random.default_rng(3 + second_order)
    Bsz, N, n, m = 2, 6, 5, 3
    spd = lambda k: (lambda M: M @ M.T + np.eye(k))(rng.normal(size=(k, k)))
    f_x = rng.normal(size=(Bsz, N, n, n)) * 0.5
    f_u = rng.normal(size=(Bsz, N, n, m))
    f_xx = rng.normal(size=(Bsz, N, n, n, n)) * 0.1
    f_xx = f_xx + np.swapaxes(f_xx, -1, -2)
    f_ux = rng.normal(size=(Bsz, N, n, m, n)) * 0.1
    f_uu = rng.normal(size=(Bsz, N, n, m, m)) * 0.1
    f_uu = f_uu + np.swapaxes(f_uu, -1, -2)
    c = rng.normal(size=(Bsz, N))
    c_x = rng.normal(size=(Bsz, N, n))
    c_u = rng.normal(size=(Bsz, N, m))
    czz = np.stack([[spd(n + m) for _ in range(N)] for _ in range(Bsz)])
    c_xx, c_ux, c_uu = czz[..., :n, :n].copy(), czz[..., n:, :n].copy(), czz[..., n:, n:].copy()
    v = rng.normal(size=(Bsz,))
    v_x = rng.normal(size=(Bsz, n))
    v_xx = np.stack([spd(n) for _ in range(Bsz)])
    l = np.zeros((Bsz, N, m))
    L = np.zeros((Bsz, N, m, n))
    vo, vxo, vxxo = np.zeros(Bsz), np.zeros((Bsz, n)), np.zeros((Bsz, n, n))
    time_ops = [(f_x, 2), (f_u, 2), (f_xx, 3), (f_ux, 3), (f_uu, 3), (c, 0), (c_x, 1), (c_u, 1), (c_xx, 2), (c_ux, 2),
                (c_uu, 2)]
    zs = [H.arr(a, k) for a, k in time_ops] + [H.arr(v, 0, False), H.arr(v_x, 1, False), H.arr(v_xx, 2, False)]
    H.hs.hs_backward(1, C.c_int64(Bsz), N, n, m, second_order, *[C.byref(z) for z in zs], H.P(l), H.P(L), H.P(vo),
                     H.P(vxo), H.P(vxxo))
    T = torch.as_tensor
    for b in range(Bsz):
        cost = opt.QuadraticCostFunction(T(c[b]), T(c_x[b]), T(c_u[b]), T(c_xx[b]), T(c_ux[b]), T(c_uu[b]))
        Vf = opt.QuadraticValueFunction(T(v[b]), T(v_x[b]), T(v_xx[b]))
        if second_order:
            dyn = opt.QuadraticDynamics(torch.zeros(N, n), T(f_x[b]), T(f_u[b]), T(f_xx[b]), T(f_ux[b]), T(f_uu[b]))
            pol = oilqr.backwardPass_ddp(dyn, cost, Vf)
        else:
            dyn = opt.AffineDynamics(torch.zeros(N, n), T(f_x[b]), T(f_u[b]))
            pol = oilqr.backwardPass_ilqr(dyn, cost, Vf)
        assert relerr(l[b], pol.l.numpy()) < 1e-9
        assert relerr(L[b], pol.L.numpy()) < 1e-9


def test_riccati_step_known_answers():  # reference tests/test_ilqrUtils.py:56-81, :110-135
    n = m = 2
    I, Z2, z = np.eye(2)[None, None], np.zeros((1, 1, 2, 2)), np.zeros((1, 1, 2))
    z3 = np.zeros((1, 1, 2, 2, 2))
    for so, Lexp, vexp in ((0, -0.5, 1.5), (1, -1 / 2.001, 2.001 - 1 / 2.001)):
        l, L = np.zeros((1, 1, 2)), np.zeros((1, 1, 2, 2))
        vo, vxo, vxxo = np.zeros(1), np.zeros((1, 2)), np.zeros((1, 2, 2))
        ops = [(I, 2), (I, 2), (z3, 3), (z3, 3), (z3, 3), (np.zeros((1, 1)), 0), (z, 1), (z, 1), (I, 2), (Z2, 2), (I, 2)]
        zs = [H.arr(a.copy(), k) for a, k in ops] + [H.arr(np.zeros(1), 0, False), H.arr(np.zeros((1, 2)), 1, False),
                                                     H.arr(np.eye(2)[None].copy(), 2, False)]
        H.hs.hs_backward(1, C.c_int64(1), 1, n, m, so, *[C.byref(a) for a in zs], H.P(l), H.P(L), H.P(vo), H.P(vxo),
                         H.P(vxxo))
        assert vo[0] == 0 and np.all(vxo == 0) and np.all(l == 0)
        assert L[0, 0] == pytest.approx(Lexp * np.eye(2), rel=1e-12)
        assert vxxo[0] == pytest.approx(vexp * np.eye(2), rel=1e-12)
        if so == 0:  # exact equality in the reference test
            assert np.all(L[0, 0] == -0.5 * np.eye(2)) and np.all(vxxo[0] == 1.5 * np.eye(2))


def _quad_problem(N=30, Bsz=3, seed=11, R_scale=1.0, spread=10.0):
    rng = np.random.default_rng(seed)
    x0 = np.zeros((Bsz, 12))
    x0[:, 9:12] = rng.uniform(-spread, spread, (Bsz, 3))
    uG = np.tile(configs.U_TRIM, (Bsz, N, 1))
    Q, R, Qf = np.eye(12), R_scale * np.eye(4), 10 * np.eye(12)
    return x0, uG, Q, R, Qf


def _oracle_fns(Q, R, Qf, dt=0.1):
    ac = Quadcopter()
    Qt, Rt, Qft = torch.as_tensor(Q), torch.as_tensor(R), torch.as_tensor(Qf)
    return ac.eulerStep(dt), (lambda x, u: x @ Qt @ x + u @ Rt @ u), (lambda x: x @ Qft @ x)


def test_rollout_and_forward_pass_quadcopter():
    N, Bsz = 25, 3
    x0, uG, Q, R, Qf = _quad_problem(N, Bsz)
    rng = np.random.default_rng(5)
    l = rng.normal(size=(Bsz, N, 4)) * 0.3
    L = rng.normal(size=(Bsz, N, 4, 12)) * 0.05
    xP = rng.normal(size=(Bsz, N + 1, 12)) * 0.1
    xP[:, :, 9:12] += x0[:, None, 9:12]
    uP = uG + rng.normal(size=(Bsz, N, 4)) * 0.1
    dyn, rc, tc = _oracle_fns(Q, R, Qf)
    M = H.model_quad(0.1)
    cst = H.cost(Q[None], R[None], Qf[None], True)
    T = torch.as_tensor
    for alpha in (1.0, 0.25):
        xT, uT, J = np.zeros((Bsz, N + 1, 12)), np.zeros((Bsz, N, 4)), np.zeros(Bsz)
        H.hs.hs_rollout(1, C.c_int64(Bsz), N, C.byref(M), C.byref(cst), H.P(x0), H.P(l), H.P(L), H.P(xP), H.P(uP),
                        C.c_double(alpha), H.P(xT), H.P(uT), H.P(J))
        for b in range(Bsz):
            tr = oilqr.trajectoryRollout(T(x0[b]), dyn, opt.AffinePolicy(T(l[b]), T(L[b])),
                                         opt.Trajectory(T(xP[b]), T(uP[b])), alpha=alpha)
            Jr = opt.CostFunction(rc, tc)(tr)
            assert relerr(xT[b], tr.xTraj.numpy()) < 1e-11 and relerr(uT[b], tr.uTraj.numpy()) < 1e-11
            assert abs(J[b] - float(Jr)) < 1e-11 * abs(float(Jr))
    xT, uT, J = np.zeros((Bsz, N + 1, 12)), np.zeros((Bsz, N, 4)), np.zeros(Bsz)
    idx = np.zeros(Bsz, dtype=np.int32)
    Jall = np.zeros((Bsz, 16))
    H.hs.hs_forward_pass(1, C.c_int64(Bsz), N, C.byref(M), C.byref(cst), H.P(x0), H.P(l), H.P(L), H.P(xP), H.P(uP),
                         H.P(xT), H.P(uT), H.P(J), H.P(idx), H.P(Jall))
    for b in range(Bsz):
        tr, Jr, ir, Jar = oilqr.forwardPass2(T(x0[b]), dyn, opt.CostFunction(rc, tc), opt.AffinePolicy(T(l[b]), T(L[b])),
                                             opt.Trajectory(T(xP[b]), T(uP[b])), return_all=True)
        assert idx[b] == ir
        assert relerr(Jall[b], Jar.numpy()) < 1e-11
        assert relerr(xT[b], tr.xTraj.numpy()) < 1e-11 and relerr(uT[b], tr.uTraj.numpy()) < 1e-11


def test_rollout_known_answer_linear():
    """reference tests/test_ilqrUtils.py:7-22: f = x + u, policy alpha*k -> x=[0,0,1,3], u=[0,1,2]"""
    N = 3
    A = np.eye(1)[None]
    M = H.model_linear(A, A.copy(), True)
    l = np.arange(3.0).reshape(1, 3, 1)
    L = np.zeros((1, 3, 1, 1))
    for alpha, xe, ue in ((1.0, [0, 0, 1, 3], [0, 1, 2]), (0.5, [0, 0, 0.5, 1.5], [0, 0.5, 1])):
        xT, uT = np.zeros((1, N + 1, 1)), np.zeros((1, N, 1))
        x0, xP, uP = np.zeros((1, 1)), np.zeros((1, N + 1, 1)), np.zeros((1, N, 1))
        H.hs.hs_rollout(1, C.c_int64(1), N, C.byref(M), None, H.P(x0), H.P(l), H.P(L), H.P(xP), H.P(uP),
                        C.c_double(alpha), H.P(xT), H.P(uT), None)
        assert np.all(xT[0, :, 0] == np.array(xe)) and np.all(uT[0, :, 0] == np.array(ue))


def _run_solve(second_order, model, cst, x0, uG, N, n, m, maxIter, tol, dt=np.float64):
    Bsz = x0.shape[0]
    xT, uT = np.zeros((Bsz, N + 1, n), dtype=dt), np.zeros((Bsz, N, m), dtype=dt)
    L, J = np.zeros((Bsz, N, m, n), dtype=dt), np.zeros(Bsz, dtype=dt)
    conv, iters = np.zeros(Bsz, dtype=np.uint8), np.zeros(Bsz, dtype=np.int32)
    alog, Jlog = np.zeros((Bsz, maxIter), dtype=np.int32), np.zeros((Bsz, maxIter + 1), dtype=dt)
    x0c, uGc = np.ascontiguousarray(x0, dtype=dt), np.ascontiguousarray(uG, dtype=dt)
    H.hs.hs_ilqr_solve(int(dt == np.float64), C.c_int64(Bsz), N, second_order, C.byref(model), C.byref(cst),
                       H.P(x0c), H.P(uGc), maxIter, C.c_double(tol), H.P(xT), H.P(uT), H.P(L),
                       H.P(J), H.P(conv), H.P(iters), H.P(alog), H.P(Jlog))
    return xT, uT, L, J, conv, iters, alog, Jlog


@pytest.mark.parametrize("second_order", [0, 1])
def test_solve_linear_known_answer(second_order):  # reference tests/test_ilqrUtils.py:167-196: converged
    I = np.eye(2)[None]
    M = H.model_linear(I, I.copy(), True)
    cst = H.cost(I.copy(), I.copy(), I.copy(), True)
    x0 = np.array([[2.0, 1.0]])
    uG = np.zeros((1, 3, 2))
    xT, uT, L, J, conv, iters, alog, Jlog = _run_solve(second_order, M, cst, x0, uG, 3, 2, 2, 100, 1e-3)
    assert conv[0] == 1
    solver = oilqr.differentialDynamicProgramming if second_order else oilqr.iterativeLqr
    A = B = Q = R = torch.eye(2)
    log = []
    tr, Lr, Jr, cr = solver(lambda x, u: A @ x + B @ u, lambda x, u: x @ Q @ x + u @ R @ u, lambda x: x @ Q @ x,
                            torch.as_tensor(x0[0]), torch.as_tensor(uG[0]), log=log)
    assert cr and iters[0] == len(log) - 1
    assert [e["alpha_idx"] for e in log[1:]] == list(alog[0, :iters[0]])
    assert relerr(xT[0], tr.xTraj.numpy()) < 1e-10 and relerr(L[0], Lr.numpy()) < 1e-10
    assert abs(J[0] - float(Jr)) < 1e-10 * abs(float(Jr))


@pytest.mark.parametrize("second_order,R_scale,spread", [(0, 1.0, 10.0), (1, 0.2, 5.0)])
def test_solve_quadcopter_vs_oracle(second_order, R_scale, spread):
    """3 iterations of iLQR / DDP on the quadcopter (demo-shaped problems, N=30): step-size sequence first,
    then trajectories, gains, cost at fp64 1e-10."""
    N, Bsz, iters_n = 30, 2, 3
    x0, uG, Q, R, Qf = _quad_problem(N, Bsz, seed=21 + second_order, R_scale=R_scale, spread=spread)
    M = H.model_quad(0.1)
    cst = H.cost(Q[None], R[None], Qf[None], True)
    xT, uT, L, J, conv, iters, alog, Jlog = _run_solve(second_order, M, cst, x0, uG, N, 12, 4, iters_n, -1.0)
    dyn, rc, tc = _oracle_fns(Q, R, Qf)
    solver = oilqr.differentialDynamicProgramming if second_order else oilqr.iterativeLqr
    for b in range(Bsz):
        log = []
        tr, Lr, Jr, cr = solver(dyn, rc, tc, torch.as_tensor(x0[b]), torch.as_tensor(uG[b]), maxIter=iters_n, tol=-1.0,
                                log=log)
        assert [e["alpha_idx"] for e in log[1:]] == list(alog[b])
        assert relerr(Jlog[b], np.array([e["J"] for e in log])) < 1e-10
        assert relerr(xT[b], tr.xTraj.numpy()) < 1e-10 and relerr(uT[b], tr.uTraj.numpy()) < 1e-10
        assert relerr(L[b], Lr.numpy()) < 1e-10
        assert not conv[b] and iters[b] == iters_n


# ------------------------------------------------------------------------------------------------ (8,4) fp32 kernel body
def _s84_problem(rng, Bsz, T, tv=True):
    n, m = 8, 4
    A = rng.normal(size=(Bsz, T, n, n)) * 0.5 / np.sqrt(n)
    Bm = rng.normal(size=(Bsz, T, n, m))
    spd = lambda k: (lambda M: M @ M.T / k + np.eye(k))(rng.normal(size=(k, k)))
    Q = np.stack([[spd(n) for _ in range(T)] for _ in range(Bsz)])
    R = np.stack([[spd(m) for _ in range(T)] for _ in range(Bsz)])
    Hm = 0.2 * rng.normal(size=(Bsz, T, m, n))
    d = 0.1 * rng.normal(size=(Bsz, T, n))
    q, r, q0 = rng.normal(size=(Bsz, T, n)), rng.normal(size=(Bsz, T, m)), rng.normal(size=(Bsz, T))
    ops = [A, Bm, d, Q, R, Hm, q, r, q0]
    if not tv:  # every operand constant in time (stride_t = 0 on the device: staged once)
        ops = [np.repeat(a[:, :1], T, axis=1) for a in ops]
    return ops


def _run_s84(bilinear, ops, N, time_invariant=False, dt=np.float32):
    A, Bm, d, Q, R, Hm, q, r, _ = (np.ascontiguousarray(a, dtype=dt) for a in ops)
    Bsz, T = Q.shape[0], Q.shape[1]
    cut = (lambda a: np.ascontiguousarray(a[:, :1])) if time_invariant else (lambda a: a)
    zs = [H.arr(cut(a), k) for a, k in ((A, 2), (Bm, 2), (d, 1), (Q, 2), (R, 2), (Hm, 2), (q, 1), (r, 1))]
    L, l, V0 = np.zeros((Bsz, N, 4, 8), dt), np.zeros((Bsz, N, 4), dt), np.zeros((Bsz, 8, 8), dt)
    fn = H.hs.hs_riccati_s84 if dt == np.float32 else H.hs.hs_riccati_s84d  # lqr_s84.cuh / lqr_s84d.cuh
    fn(int(bilinear), C.c_int64(Bsz), N, 1 if time_invariant else T, *[C.byref(z) for z in zs], H.P(L), H.P(l), H.P(V0))
    return L, l, V0


def _pre(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float((np.abs(a - b).reshape(a.shape[0], -1).max(1) / np.abs(b).reshape(b.shape[0], -1).max(1)).max())


@pytest.mark.parametrize("dt,tol", [(np.float32, 1e-5), (np.float64, 1e-10)])
@pytest.mark.parametrize("tv", [True, False])
def test_s84_kernel_body_vs_oracle(tv, dt, tol):
    """lqr_s84.cuh (fp32) and lqr_s84d.cuh (fp64) compiled for the host against the fp64 oracle: time-varying operands with T > N
    and a ragged last warp, and fully time-invariant operands staged once; the gates are BASELINE.md section 6's fp32 1e-5
    (measured: below 1e-6) and fp64 1e-10."""
    rng = np.random.default_rng(7)
    T, N, Bsz = 9, 7, 37
    ops = _s84_problem(rng, Bsz, T, tv)
    A, Bm, d, Q, R, Hm, q, r, q0 = ops
    Lref, Vref = olqr.discreteFiniteHorizonLqr_batched(A, Bm, Q, R, N, return_value=True)
    L, _, V0 = _run_s84(False, ops, N, time_invariant=not tv, dt=dt)
    assert _pre(L, Lref) < tol and _pre(V0, Vref) < tol
    Lr, lr = olqr.bilinearAffineLqr_batched(A, Bm, d, Q, R, Hm, q, r, q0, N)
    L, l, _ = _run_s84(True, ops, N, time_invariant=not tv, dt=dt)
    assert _pre(L, Lr) < tol and _pre(l, lr) < tol


@pytest.mark.parametrize("dt,tol", [(np.float32, 1e-5), (np.float64, 1e-10)])
def test_s84_kernel_body_demo_horizon(dt, tol):
    """the demos' problems at N = 100 (demos/discreteFiniteHorizonLqr.py:29-35 incl. the Q[-1] terminal quirk,
    demos/bilinearLqrControl.py:21-43 with a seeded H): fp32 within 1e-5, fp64 within 1e-10 of the fp64 oracle"""
    N = 100
    A0, B0 = (t.numpy() for t in Quadcopter().linearize(np.zeros(8), configs.U_TRIM, dt=0.1))
    Qk, Rk = (np.asarray(a) for a in configs.cfg1_demo_weights(N))
    rep = lambda a: np.repeat(a[None, None], N, axis=1)
    Kref = olqr.discreteFiniteHorizonLqr(rep(A0)[0], rep(B0)[0], Qk, Rk, N)
    z = np.zeros((1, N, 1))
    L, _, _ = _run_s84(False, [rep(A0), rep(B0), z, Qk[None], Rk[None], z, z, z, z], N, dt=dt)
    assert _pre(L, np.asarray(Kref)[None]) < tol
    rng = np.random.default_rng(1)
    Bsz = 5
    bat = lambda a: np.repeat(a, Bsz, axis=0)
    d, Hm = rng.normal(size=(Bsz, N, 8)) * 0.01, 0.2 * rng.normal(size=(Bsz, N, 4, 8))
    q = 0.1 * bat(rep(np.array([1., -1, 0, 0, 0, 0, 0, 0])))
    r, q0 = rng.normal(size=(Bsz, N, 4)) * 0.1, rng.normal(size=(Bsz, N))
    ops = [bat(rep(A0)), bat(rep(B0)), d, bat(rep(np.eye(8))), bat(rep(np.eye(4))), Hm, q, r, q0]
    Lr, lr = olqr.bilinearAffineLqr_batched(*ops, N)
    L, l, _ = _run_s84(True, ops, N, dt=dt)
    assert _pre(L, Lr) < tol and _pre(l, lr) < tol
