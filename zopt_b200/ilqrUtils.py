"""
Iterative LQR / differential dynamic programming -- mirror of zopt/ilqrUtils.py.

Same function names, argument order and return pytrees as the reference.  Differences forced by
the GPU:
  * `dynamics` / `dynFun` must be a registered model (zopt_b200.models.LinearDynamics or
    QuadcopterEuler) or a symbolic definition compiled into a plug-in (zopt_b200.plugin.SymbolicDynamics),
    and costs `QuadraticCost` / `QuadraticTerminalCost`; an arbitrary callable raises TypeError
    (no CPU fallback).
  * every array may carry one leading batch axis; `J` and `converged` then have shape (Bsz,).
  * the legacy `forwardPass` (ilqrUtils.py:69-113) is not provided: neither solver calls it, and its
    loop never terminates once alpha <= alphaMin.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import View, ZbArr, check, dcode, lib, null_arr, pick_device, pick_dtype, ptr, stream_ptr, to_dev
from .models import (QuadraticCost, QuadraticTerminalCost, cost_batch, cost_spec, reconcile_batch, require_cost, require_model,
                     symbolic_cost_of)
from .pytrees import (AffineDynamics, AffinePolicy, CostFunction, QuadraticCostFunction, QuadraticDynamics,
                      QuadraticValueFunction, Trajectory)


def _batch_of(ts, core_ndims):
    Bsz, any_b = 1, False
    for t, nd in zip(ts, core_ndims):
        if t.ndim == nd + 1:
            any_b = True
            if t.shape[0] != 1:
                if Bsz != 1 and Bsz != t.shape[0]:
                    raise ValueError(f"inconsistent batch sizes {Bsz} and {t.shape[0]}")
                Bsz = t.shape[0]
        elif t.ndim != nd:
            raise ValueError(f"expected {nd} or {nd + 1} dimensions, got shape {tuple(t.shape)}")
    return Bsz, any_b


def _full(t, nd, Bsz):
    """materialise an operand as (Bsz, ...) contiguous (trajectory-sized operands only)"""
    if t.ndim == nd:
        t = t.unsqueeze(0)
    return t.expand((Bsz,) + tuple(t.shape[1:])).contiguous()


def _rollout_args(x0, dynFun, policy, trajPrev, costFun):
    model = require_model(dynFun)
    l, L = policy
    xPrev, uPrev = trajPrev
    device = pick_device(x0, l, L, xPrev, uPrev)
    dtype = pick_dtype(x0, l, L, xPrev, uPrev)
    x0, l, L, xPrev, uPrev = (to_dev(t, dtype, device) for t in (x0, l, L, xPrev, uPrev))
    Bsz, any_b = _batch_of([x0, l, L, xPrev, uPrev], [1, 2, 3, 2, 2])
    Bsz = reconcile_batch(Bsz, model.batch())  # every operand: shared, or exactly Bsz problems
    cspec, ckeep = None, None
    if costFun is not None:
        rc, tc = require_cost(*costFun)
        Bsz = reconcile_batch(Bsz, cost_batch(rc, tc, model.n, model.m))
        cspec, ckeep = cost_spec(rc, tc, dtype, device, model.n, model.m)
    any_b = any_b or Bsz > 1
    x0, l, L, xPrev, uPrev = (_full(t, nd, Bsz) for t, nd in zip((x0, l, L, xPrev, uPrev), (1, 2, 3, 2, 2)))
    N, m = l.shape[1], l.shape[2]
    n = x0.shape[1]
    if (n, m) != (model.n, model.m):
        raise ValueError(f"policy/state dimensions ({n},{m}) do not match the model ({model.n},{model.m})")
    mspec, mkeep = (model, None) if getattr(model, "is_plugin", False) else model.spec(dtype, device)
    return device, dtype, Bsz, any_b, N, n, m, x0, l, L, xPrev, uPrev, mspec, mkeep, cspec, ckeep


def trajectoryRollout(x0, dynFun, policy, trajPrev, alpha=1):
    """
    Roll out a trajectory from the initial state with the affine policy (zopt/ilqrUtils.py:33-66):
    `u_k = alpha*l_k + L_k (x_k - xPrev_k) + uPrev_k`, `x_{k+1} = dynFun(x_k, u_k)`.
    """
    (device, dtype, Bsz, any_b, N, n, m, x0, l, L, xPrev, uPrev, mspec, mkeep, _, _) = \
        _rollout_args(x0, dynFun, policy, trajPrev, None)
    xTraj = torch.empty((Bsz, N + 1, n), dtype=dtype, device=device)
    uTraj = torch.empty((Bsz, N, m), dtype=dtype, device=device)
    if getattr(mspec, "is_plugin", False):  # user-defined symbolic model (plugin.py)
        mspec._need()
        mspec._check(mspec._lib.zb_user_rollout(dcode(dtype), device.index, stream_ptr(device), Bsz, N, None, ptr(x0), ptr(l), ptr(L),
                                                ptr(xPrev), ptr(uPrev), float(alpha), ptr(xTraj), ptr(uTraj), None))
    else:
        check(lib.zb_ilqr_rollout(dcode(dtype), device.index, stream_ptr(device), Bsz, N, C.byref(mspec), None, ptr(x0), ptr(l),
                                  ptr(L), ptr(xPrev), ptr(uPrev), float(alpha), ptr(xTraj), ptr(uTraj), None))
    return Trajectory(xTraj, uTraj) if any_b else Trajectory(xTraj[0], uTraj[0])


def forwardPass2(x0, dynFun, costFun, policy, trajPrev, return_index=False):
    """
    Simplified iLQR forward pass (zopt/ilqrUtils.py:116-150): roll out the 16 step sizes 0.5**j, take the argmin of the
    cost (no acceptance test; first index on ties; a NaN cost wins).  Returns (Trajectory, J).
    """
    (device, dtype, Bsz, any_b, N, n, m, x0, l, L, xPrev, uPrev, mspec, mkeep, cspec, ckeep) = \
        _rollout_args(x0, dynFun, policy, trajPrev, costFun)
    if getattr(mspec, "is_plugin", False):
        raise TypeError("forwardPass2 as a stand-alone building block is available for the registered models; a "
                        "SymbolicDynamics model runs it inside iterativeLqr / differentialDynamicProgramming")
    xTraj = torch.empty((Bsz, N + 1, n), dtype=dtype, device=device)
    uTraj = torch.empty((Bsz, N, m), dtype=dtype, device=device)
    J = torch.empty((Bsz,), dtype=dtype, device=device)
    idx = torch.empty((Bsz,), dtype=torch.int32, device=device)
    Jall = torch.empty((Bsz, 16), dtype=dtype, device=device)
    check(lib.zb_ilqr_forward_pass(dcode(dtype), device.index, stream_ptr(device), Bsz, N, C.byref(mspec), C.byref(cspec),
                                   ptr(x0), ptr(l), ptr(L), ptr(xPrev), ptr(uPrev), ptr(xTraj), ptr(uTraj), ptr(J),
                                   ptr(idx), ptr(Jall)))
    if not any_b:
        xTraj, uTraj, J, idx, Jall = xTraj[0], uTraj[0], J[0], idx[0], Jall[0]
    if return_index:
        return Trajectory(xTraj, uTraj), J, idx, Jall
    return Trajectory(xTraj, uTraj), J


# ------------------------------------------------------------------------------------------------ backward pass
def _backward(dynamics, cost, value, second_order, stacked):
    """Shared driver of riccatiStep_* (stacked=False: leaves without a time axis) and backwardPass_*."""
    if second_order:
        _, f_x, f_u, f_xx, f_ux, f_uu = dynamics
    else:
        _, f_x, f_u = dynamics
        f_xx = f_ux = f_uu = None
    c, c_x, c_u, c_xx, c_ux, c_uu = cost
    v, v_x, v_xx = value
    leaves = [f_x, f_u, c, c_x, c_u, c_xx, c_ux, c_uu, v, v_x, v_xx] + ([f_xx, f_ux, f_uu] if second_order else [])
    device = pick_device(*leaves)
    dtype = pick_dtype(*leaves)
    leaves = [to_dev(t, dtype, device) for t in leaves]
    f_x, f_u, c, c_x, c_u, c_xx, c_ux, c_uu, v, v_x, v_xx = leaves[:11]
    t_ = 1 if stacked else 0  # number of time axes on the stacked leaves
    core = [2 + t_, 2 + t_, 0 + t_, 1 + t_, 1 + t_, 2 + t_, 2 + t_, 2 + t_, 0, 1, 2] + ([3 + t_] * 3 if second_order else [])
    Bsz, any_b = _batch_of(leaves, core)
    n, m = f_u.shape[-2], f_u.shape[-1]
    N = f_x.shape[-3] if stacked else 1
    blocks = [2, 2, 0, 1, 1, 2, 2, 2, 0, 1, 2] + ([3, 3, 3] if second_order else [])
    views = []
    for i, (t, k, nd) in enumerate(zip(leaves, blocks, core)):
        is_value = i in (8, 9, 10)
        has_time = stacked and not is_value
        views.append(View(t if (has_time or is_value or not stacked) else t, k, has_time, t.ndim == nd + 1))
    fx_v, fu_v, c_v, cx_v, cu_v, cxx_v, cux_v, cuu_v, v_v, vx_v, vxx_v = views[:11]
    if second_order:
        fxx_r, fux_r, fuu_r = (w.ref() for w in views[11:])
    else:
        fxx_r = fux_r = fuu_r = null_arr()
    l = torch.empty((Bsz, N, m), dtype=dtype, device=device)
    L = torch.empty((Bsz, N, m, n), dtype=dtype, device=device)
    vo = torch.empty((Bsz,), dtype=dtype, device=device)
    vxo = torch.empty((Bsz, n), dtype=dtype, device=device)
    vxxo = torch.empty((Bsz, n, n), dtype=dtype, device=device)
    check(lib.zb_ilqr_backward(dcode(dtype), device.index, stream_ptr(device), Bsz, N, n, m, int(second_order), fx_v.ref(),
                               fu_v.ref(), fxx_r, fux_r, fuu_r, c_v.ref(), cx_v.ref(), cu_v.ref(), cxx_v.ref(),
                               cux_v.ref(), cuu_v.ref(), v_v.ref(), vx_v.ref(), vxx_v.ref(), ptr(l), ptr(L), ptr(vo),
                               ptr(vxo), ptr(vxxo)))
    if not stacked:
        l, L = l[:, 0], L[:, 0]
    if not any_b:
        l, L, vo, vxo, vxxo = l[0], L[0], vo[0], vxo[0], vxxo[0]
    return QuadraticValueFunction(vo, vxo, vxxo), AffinePolicy(l, L)


def riccatiStep_ilqr(dynamics, cost, value):
    """One step of the backward Riccati recursion (zopt/ilqrUtils.py:153-173) -> (valueOut, policy)."""
    return _backward(dynamics, cost, value, False, False)


def backwardPass_ilqr(dynamics, cost, Vf):
    """Backward pass of iLQR over stacked pytrees (zopt/ilqrUtils.py:176-181) -> AffinePolicy."""
    return _backward(dynamics, cost, Vf, False, True)[1]


def riccatiStep_ddp(dynamics, cost, value):
    """One DDP Riccati step with the eigen-clamped v_x.f_zz block (zopt/ilqrUtils.py:184-206)."""
    return _backward(dynamics, cost, value, True, False)


def backwardPass_ddp(dynamics, cost, Vf):
    """Backward pass of DDP over stacked pytrees (zopt/ilqrUtils.py:209-214) -> AffinePolicy."""
    return _backward(dynamics, cost, Vf, True, True)[1]


# ------------------------------------------------------------------------------------------------ conditioning
def ensurePositiveDefinite(a, eps=1e-3):
    """`V max(Lambda, eps) V^T` of symmetric (p,p) blocks, any leading axes (zopt/ilqrUtils.py:217-219)."""
    device = pick_device(a)
    dtype = pick_dtype(a)
    a = to_dev(a, dtype, device)
    p = a.shape[-1]
    flat = a.reshape(-1, p, p).contiguous()
    out = torch.empty_like(flat)
    check(lib.zb_pd_clamp(dcode(dtype), device.index, stream_ptr(device), flat.shape[0], p, float(eps), ptr(flat), ptr(out)))
    return out.reshape(a.shape)


def conditionQuadraticCost(quadratic_cost):
    """Make the stacked cost Hessian [[c_xx, c_ux^T],[c_ux, c_uu]] positive definite (zopt/ilqrUtils.py:222-234)."""
    (c, c_x, c_u, c_xx, c_ux, c_uu) = quadratic_cost
    n, m = c_xx.shape[-1], c_uu.shape[-1]
    c_zz = torch.cat([torch.cat([c_xx, c_ux.transpose(-1, -2)], dim=-1), torch.cat([c_ux, c_uu], dim=-1)], dim=-2)
    c_zz = ensurePositiveDefinite(c_zz)
    return QuadraticCostFunction(c, c_x, c_u, c_zz[..., :n, :n], c_zz[..., -m:, :n], c_zz[..., -m:, -m:])


def conditionQuadraticDynamics(quadratic_dynamics, v_x):
    """Eigen-clamped contraction of the second-order dynamics with v_x (zopt/ilqrUtils.py:237-251)."""
    _, _, _, f_xx, f_ux, f_uu = quadratic_dynamics
    v_x = v_x.to(f_xx)
    vf_xx = torch.einsum('...i,...ijk->...jk', v_x, f_xx)
    vf_uu = torch.einsum('...i,...ijk->...jk', v_x, f_uu)
    vf_ux = torch.einsum('...i,...ijk->...jk', v_x, f_ux)
    n, m = vf_xx.shape[-1], vf_uu.shape[-1]
    vf_zz = torch.cat([torch.cat([vf_xx, vf_ux.transpose(-1, -2)], dim=-1), torch.cat([vf_ux, vf_uu], dim=-1)], dim=-2)
    vf_zz = ensurePositiveDefinite(vf_zz)
    return vf_zz[..., :n, :n], vf_zz[..., -m:, :n], vf_zz[..., -m:, -m:]


def conditionValueFunction(Vf):
    """zopt/ilqrUtils.py:254-257"""
    v, v_x, v_xx = Vf
    return QuadraticValueFunction(v, v_x, ensurePositiveDefinite(v_xx))


# ------------------------------------------------------------------------------------------------ solvers
# test hook: force the two-kernel line search (k_forward_costs + k_forward_commit) instead of the fused quadcopter kernel
_GENERIC_FORWARD = False


def _bind_symbolic_cost(model, runningCost, terminalCost):
    """costs given symbolically (plugin.SymbolicCost): compile them together with the symbolic model into one plug-in and hand
    the kernels placeholder quadratic weights (never read: the plug-in evaluates the generated cost code)"""
    sc = symbolic_cost_of(runningCost, terminalCost)
    if sc is None:
        return model, runningCost, terminalCost
    if not getattr(model, "is_plugin", False):
        raise TypeError("symbolic costs are compiled together with a symbolic model: define the dynamics with "
                        "zopt_b200.plugin.SymbolicDynamics(f, n, m) as well")
    import numpy as np
    return model.with_cost(sc), QuadraticCost(np.eye(model.n), np.eye(model.m)), QuadraticTerminalCost(np.eye(model.n))


def _solve(dynamics, runningCost, terminalCost, x0, uGuess, maxIter, tol, second_order, return_log):
    model = require_model(dynamics)
    model, runningCost, terminalCost = _bind_symbolic_cost(model, runningCost, terminalCost)
    rc, tc = require_cost(runningCost, terminalCost)
    device = pick_device(x0, uGuess)
    dtype = pick_dtype(x0, uGuess)
    x0, uGuess = to_dev(x0, dtype, device), to_dev(uGuess, dtype, device)
    Bsz, any_b = _batch_of([x0, uGuess], [1, 2])
    Bsz = reconcile_batch(Bsz, model.batch(), cost_batch(rc, tc, model.n, model.m))  # shared (1) or exactly Bsz, nothing else
    any_b = any_b or Bsz > 1
    x0, uGuess = _full(x0, 1, Bsz), _full(uGuess, 2, Bsz)
    n, (N, m) = x0.shape[1], uGuess.shape[1:]
    if (n, m) != (model.n, model.m):
        raise ValueError(f"x0/uGuess dimensions ({n},{m}) do not match the model ({model.n},{model.m})")
    maxIter = int(maxIter)
    cspec, ckeep = cost_spec(rc, tc, dtype, device, model.n, model.m)
    if getattr(model, "is_plugin", False):  # user-defined symbolic model: its own compiled solver library (plugin.py)
        xTraj, uTraj, L, J, conv, iters, alog, Jlog = model.solve(cspec, dtype, device, Bsz, N, x0, uGuess, maxIter, tol,
                                                                  second_order, return_log)
        return _pack_solution(xTraj, uTraj, L, J, conv, iters, alog, Jlog, any_b, return_log, maxIter)
    mspec, mkeep = model.spec(dtype, device)
    xTraj = torch.empty((Bsz, N + 1, n), dtype=dtype, device=device)
    uTraj = torch.empty((Bsz, N, m), dtype=dtype, device=device)
    L = torch.empty((Bsz, N, m, n), dtype=dtype, device=device)
    J = torch.empty((Bsz,), dtype=dtype, device=device)
    conv = torch.empty((Bsz,), dtype=torch.uint8, device=device)
    iters = torch.empty((Bsz,), dtype=torch.int32, device=device)
    alog = torch.empty((Bsz, max(maxIter, 1)), dtype=torch.int32, device=device) if return_log else None
    Jlog = torch.empty((Bsz, maxIter + 1), dtype=dtype, device=device) if return_log else None
    wsb = lib.zb_ilqr_workspace_bytes(dcode(dtype), Bsz, N, n, m)
    ws = torch.empty((wsb,), dtype=torch.uint8, device=device)
    from .mpcUtils import _is_diagonal
    Qd, Rd, Qfd = (vw.t for vw in ckeep)
    flags = (1 if second_order else 0) | (2 if all(_is_diagonal(t) for t in (Qd, Rd, Qfd)) else 0)
    if _GENERIC_FORWARD:
        flags |= 32  # ZB_GENERIC_FORWARD
    check(lib.zb_ilqr_solve(dcode(dtype), device.index, stream_ptr(device), Bsz, N, flags, C.byref(mspec),
                            C.byref(cspec), ptr(x0), ptr(uGuess), maxIter, float(tol), ptr(xTraj), ptr(uTraj), ptr(L),
                            ptr(J), ptr(conv), ptr(iters), ptr(alog), ptr(Jlog), ptr(ws), wsb))
    return _pack_solution(xTraj, uTraj, L, J, conv, iters, alog, Jlog, any_b, return_log, maxIter)


def _pack_solution(xTraj, uTraj, L, J, conv, iters, alog, Jlog, any_b, return_log, maxIter):
    conv = conv.bool()
    if not any_b:
        xTraj, uTraj, L, J, conv, iters = xTraj[0], uTraj[0], L[0], J[0], conv[0], iters[0]
        if return_log:
            alog, Jlog = alog[0], Jlog[0]
    out = (Trajectory(xTraj, uTraj), L, J, conv)
    if return_log:
        return out + (dict(iters=iters, alpha_idx=alog[..., :maxIter], J=Jlog),)
    return out


def iterativeLqr(dynamics, runningCost, terminalCost, x0, uGuess, maxIter=100, tol=1e-3, return_log=False):
    """
    Iterative LQR (zopt/ilqrUtils.py:260-327).

    Returns
    -------
    traj : Trajectory(xTraj (N+1,n), uTraj (N,m))
    L : feedback gains (N,m,n): `u[k] = L[k] @ (x[k]-xTraj[k]) + uTraj[k]`
    J : cost
    converged : `abs(J_prev - J) <= tol` reached within maxIter iterations
    """
    return _solve(dynamics, runningCost, terminalCost, x0, uGuess, maxIter, tol, False, return_log)


def differentialDynamicProgramming(dynamics, runningCost, terminalCost, x0, uGuess, maxIter=100, tol=1e-3,
                                   return_log=False):
    """Differential dynamic programming (zopt/ilqrUtils.py:330-397); same returns as iterativeLqr."""
    return _solve(dynamics, runningCost, terminalCost, x0, uGuess, maxIter, tol, True, return_log)
