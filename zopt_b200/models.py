"""
Registered dynamics / cost objects: the kernel-side counterpart of the Python callables the reference
passes to iLQR/DDP (`dynamics`, `runningCost`, `terminalCost`, zopt/ilqrUtils.py:260-268).

The reference differentiates arbitrary lambdas with JAX; a CUDA kernel cannot trace a lambda, so the
drop-in accepts these objects instead.  Each one is still *callable* like the lambda it replaces
(`dynFun(x, u)`, `costFun(x, u)`, `terminalCostFun(x)`), so reference-style user code keeps working,
and each carries the spec the kernels need.  Anything else raises TypeError -- there is no CPU
fallback.

    LinearDynamics(A, B)              x+ = A x + B u                 (tests/test_ilqrUtils.py:167-196)
    plugin.SymbolicDynamics(f, n, m)  x+ = f(x, u) given in sympy, compiled into a solver plug-in (zopt_b200/plugin.py)
    QuadcopterEuler(dt, wind_ned)     x+ = x + dt*inertialDynamics   (demos/iterativeLqr.py:35, zopt/quadcopter.py:116-144)
    QuadraticCost(Q, R)               c(x,u) = x'Qx + u'Ru           (demos/iterativeLqr.py:12-13)
    QuadraticTerminalCost(Qf)         cf(x) = x'Qf x                 (demos/iterativeLqr.py:37)
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (MODEL_LINEAR, MODEL_QUADCOPTER, View, ZbArr, ZbCost, ZbModel, check, dcode, lib, pick_device, ptr,
                   stream_ptr, to_dev)


def _mat(x):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


def _block_batch(t, shape, name):
    """Batch size of an operand holding `shape` blocks: `shape` itself (shared by the batch -> 1) or one leading batch axis.
    Anything else is rejected here, before a raw pointer reaches a kernel."""
    shape = tuple(shape)
    if tuple(t.shape) == shape:
        return 1
    if t.ndim == len(shape) + 1 and tuple(t.shape[1:]) == shape:
        return int(t.shape[0])
    raise ValueError(f"{name} must have shape {shape} or (Bsz,) + {shape}, got {tuple(t.shape)}")


def reconcile_batch(*sizes):
    """One batch size for a call: every operand is either shared (1) or carries exactly that many problems."""
    Bsz = 1
    for b in sizes:
        b = int(b)
        if b != 1:
            if Bsz != 1 and Bsz != b:
                raise ValueError(f"inconsistent batch sizes {Bsz} and {b}")
            Bsz = b
    return Bsz


class LinearDynamics:
    """x+ = A x + B u; A (n,n) or (Bsz,n,n), B (n,m) or (Bsz,n,m)."""

    def __init__(self, A, B):
        self.A, self.B = _mat(A), _mat(B)
        if self.B.ndim not in (2, 3) or self.A.ndim not in (2, 3):
            raise ValueError(f"A must be (n,n) or (Bsz,n,n) and B (n,m) or (Bsz,n,m), got {tuple(self.A.shape)}, {tuple(self.B.shape)}")
        self.n, self.m = self.B.shape[-2], self.B.shape[-1]
        self.batch()  # shapes and batch sizes are checked once, here

    def __call__(self, x, u):
        A, B = self.A.to(x), self.B.to(x)
        return (A @ x.unsqueeze(-1)).squeeze(-1) + (B @ u.unsqueeze(-1)).squeeze(-1)

    def spec(self, dtype, device):
        A, B = to_dev(self.A, dtype, device), to_dev(self.B, dtype, device)
        vA, vB = View(A, 2, False, A.ndim == 3), View(B, 2, False, B.ndim == 3)
        M = ZbModel()
        M.kind, M.n, M.m, M.has_wind, M.dt = MODEL_LINEAR, self.n, self.m, 0, 0.0
        M.A, M.B = vA.arr, vB.arr
        return M, (vA, vB)

    def batch(self):
        return reconcile_batch(_block_batch(self.A, (self.n, self.n), "A"), _block_batch(self.B, (self.n, self.m), "B"))


class QuadcopterEuler:
    """Forward-Euler quadcopter step x + dt * inertialDynamics(x, u, wind_ned) (n=12, m=4)."""
    n, m = 12, 4

    def __init__(self, dt, wind_ned=None):
        self.dt = float(dt)
        self.wind = [0.0, 0.0, 0.0] if wind_ned is None else [float(v) for v in np.asarray(wind_ned).reshape(3)]

    def __call__(self, x, u):
        from .quadcopter import Quadcopter
        return x + self.dt * Quadcopter().inertialDynamics(x, u, self.wind)

    def spec(self, dtype, device):
        M = ZbModel()
        M.kind, M.n, M.m, M.dt = MODEL_QUADCOPTER, 12, 4, self.dt
        M.has_wind = int(any(w != 0.0 for w in self.wind))
        for i in range(3):
            M.wind[i] = self.wind[i]
        return M, ()

    def batch(self):
        return 1


class QuadraticCost:
    """Running cost x'Qx + u'Ru; Q (n,n) or (Bsz,n,n), R (m,m) or (Bsz,m,m)."""

    def __init__(self, Q, R):
        self.Q, self.R = _mat(Q), _mat(R)

    def __call__(self, x, u):
        Q, R = self.Q.to(x), self.R.to(x)
        return torch.einsum('...i,...ij,...j->...', x, Q, x) + torch.einsum('...i,...ij,...j->...', u, R, u)


class QuadraticTerminalCost:
    """Terminal cost x'Qf x."""

    def __init__(self, Qf):
        self.Qf = _mat(Qf)

    def __call__(self, x):
        return torch.einsum('...i,...ij,...j->...', x, self.Qf.to(x), x)


def require_model(dynFun):
    if isinstance(dynFun, (LinearDynamics, QuadcopterEuler)) or getattr(dynFun, "is_plugin", False):
        return dynFun
    raise TypeError(
        "zopt_b200 cannot differentiate or roll out an arbitrary Python callable on the GPU (the reference traces it "
        "with JAX). Pass a registered model: zopt_b200.models.LinearDynamics(A, B) or QuadcopterEuler(dt, wind_ned), or "
        "define the dynamics symbolically: zopt_b200.plugin.SymbolicDynamics(f, n, m). There is no CPU fallback.")


def symbolic_cost_of(runningCost, terminalCost):
    """the SymbolicCost both callables belong to (plugin.SymbolicCost(...).running / .terminal), else None"""
    if getattr(runningCost, "is_symbolic_cost", False) or getattr(terminalCost, "is_symbolic_cost", False):
        ro, to = getattr(runningCost, "owner", None), getattr(terminalCost, "owner", None)
        if ro is None or ro is not to:
            raise TypeError("a symbolic running cost goes with the terminal cost of the SAME plugin.SymbolicCost (cost.running, cost.terminal)")
        return ro
    return None


def require_cost(runningCost, terminalCost):
    if not isinstance(runningCost, QuadraticCost) or not isinstance(terminalCost, QuadraticTerminalCost):
        raise TypeError(
            "zopt_b200 needs registered costs: zopt_b200.models.QuadraticCost(Q, R) and QuadraticTerminalCost(Qf), or costs "
            "defined symbolically together with a symbolic model: plugin.SymbolicCost(c, cf, n, m).running / .terminal "
            "(the reference autodiffs arbitrary callables with JAX; there is no CPU fallback).")
    return runningCost, terminalCost


def cost_spec(runningCost, terminalCost, dtype, device, n=None, m=None):
    runningCost, terminalCost = require_cost(runningCost, terminalCost)
    cost_batch(runningCost, terminalCost, n, m)
    Q, R, Qf = (to_dev(t, dtype, device) for t in (runningCost.Q, runningCost.R, terminalCost.Qf))
    views = [View(t, 2, False, t.ndim == 3) for t in (Q, R, Qf)]
    c = ZbCost()
    c.Q, c.R, c.Qf = views[0].arr, views[1].arr, views[2].arr
    return c, views


def cost_batch(runningCost, terminalCost, n=None, m=None):
    """Batch size of the cost operands; Q, Qf must be (n,n) blocks and R (m,m) (n, m from the model when given), each shared
    or batched, all batched ones with the same batch size."""
    Q, R, Qf = runningCost.Q, runningCost.R, terminalCost.Qf
    n = int(Q.shape[-1]) if n is None else int(n)
    m = int(R.shape[-1]) if m is None else int(m)
    return reconcile_batch(_block_batch(Q, (n, n), "Q"), _block_batch(R, (m, m), "R"), _block_batch(Qf, (n, n), "Qf"))


# ------------------------------------------------------------------------------------------------
# analytic Taylor expansions behind the pytree constructors (zopt/pytrees.py:71-81,99-115,138-153,179-194)
def _flat(x, k):
    """(..., k) -> (P, k) contiguous, plus the leading shape"""
    lead = x.shape[:-1]
    return x.reshape(-1, k).contiguous(), lead


def expand_dynamics(dynFun, x, u, second_order):
    model = require_model(dynFun)
    device = pick_device(x, u)
    dtype = _lib.pick_dtype(x, u)
    x, u = to_dev(x, dtype, device), to_dev(u, dtype, device)
    n, m = model.n, model.m
    xf, lead = _flat(x, n)
    uf, _ = _flat(u, m)
    P = xf.shape[0]
    if getattr(model, "is_plugin", False):  # user-defined symbolic model: the generated CUDA code (plugin.py)
        xf, uf = xf.contiguous(), uf.contiguous()
        f, f_x, f_u = model.step(xf, uf, linearize=True)
        if second_order:
            f_xx, f_ux, f_uu = model.second_order(xf, uf)
    elif isinstance(model, QuadcopterEuler):
        from .quadcopter import quad_linearize, quad_xdot, quad_hess_contract
        f = xf + model.dt * quad_xdot(xf, uf, model.wind)
        f_x, f_u = quad_linearize(xf, uf, model.wind, model.dt)
        if second_order:
            eye = torch.eye(12, dtype=dtype, device=device)
            f_xx = torch.stack([quad_hess_contract(xf, uf, model.wind, model.dt, eye[i].expand(P, 12).contiguous())
                                for i in range(12)], dim=1)
            f_ux = torch.zeros((P, n, m, n), dtype=dtype, device=device)
            f_uu = torch.zeros((P, n, m, m), dtype=dtype, device=device)
    else:
        A, B = to_dev(model.A, dtype, device), to_dev(model.B, dtype, device)
        if A.ndim == 3 or B.ndim == 3:
            raise ValueError("pytree constructors take an un-batched LinearDynamics")
        f = xf @ A.T + uf @ B.T
        f_x, f_u = A.expand(P, n, n).clone(), B.expand(P, n, m).clone()
        if second_order:
            f_xx = torch.zeros((P, n, n, n), dtype=dtype, device=device)
            f_ux = torch.zeros((P, n, m, n), dtype=dtype, device=device)
            f_uu = torch.zeros((P, n, m, m), dtype=dtype, device=device)
    out = [f.reshape(lead + (n,)), f_x.reshape(lead + (n, n)), f_u.reshape(lead + (n, m))]
    if second_order:
        out += [f_xx.reshape(lead + (n, n, n)), f_ux.reshape(lead + (n, m, n)), f_uu.reshape(lead + (n, m, m))]
    return tuple(out)


def expand_cost(runningCost, x, u):
    if not isinstance(runningCost, QuadraticCost):
        raise TypeError("expected zopt_b200.models.QuadraticCost (arbitrary callables cannot be differentiated here)")
    device = pick_device(x, u)
    dtype = _lib.pick_dtype(x, u)
    x, u = to_dev(x, dtype, device), to_dev(u, dtype, device)
    Q, R = to_dev(runningCost.Q, dtype, device), to_dev(runningCost.R, dtype, device)
    Qs, Rs = Q + Q.transpose(-1, -2), R + R.transpose(-1, -2)
    lead = x.shape[:-1]
    n, m = x.shape[-1], u.shape[-1]
    c = runningCost(x, u)
    c_x = torch.einsum('ij,...j->...i', Qs, x)
    c_u = torch.einsum('ij,...j->...i', Rs, u)
    c_xx = Qs.expand(lead + (n, n)).clone()
    c_uu = Rs.expand(lead + (m, m)).clone()
    c_ux = torch.zeros(lead + (m, n), dtype=dtype, device=device)
    return c, c_x, c_u, c_xx, c_ux, c_uu


def expand_terminal(terminalCost, xf):
    if not isinstance(terminalCost, QuadraticTerminalCost):
        raise TypeError("expected zopt_b200.models.QuadraticTerminalCost")
    device = pick_device(xf)
    dtype = _lib.pick_dtype(xf)
    xf = to_dev(xf, dtype, device)
    Qf = to_dev(terminalCost.Qf, dtype, device)
    Qs = Qf + Qf.transpose(-1, -2)
    return terminalCost(xf), torch.einsum('ij,...j->...i', Qs, xf), Qs.expand(xf.shape[:-1] + Qs.shape).clone()
