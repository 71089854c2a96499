"""
Pytree containers of the hot path -- mirror of zopt/pytrees.py:6-236.

Same `NamedTuple` names, field names and field order as the reference, so code that unpacks or
indexes them keeps working.  Leaves are torch tensors (CUDA when they come out of the solvers);
`__getitem__(k)` slices every leaf (pytrees.py:11-12 and alike), `__call__` evaluates the Taylor
form with the reference's error behaviour (ValueError when a stacked pytree is called without k).

The autodiff constructors (`from_function`, `from_trajectory`, `fromTerminalCostFunction`) of the
reference differentiate arbitrary Python callables with JAX.  Here they accept the registered
model / cost objects of `zopt_b200.models` and evaluate the analytic expansions on the GPU;
anything else raises TypeError (no CPU fallback).
"""
from typing import Callable, NamedTuple

import torch


def _map(fn, tup):
    return type(tup)(*(fn(leaf) for leaf in tup))


class Trajectory(NamedTuple):
    """Trajectory tuple: (xTraj, uTraj)  -- pytrees.py:6-12"""
    xTraj: torch.Tensor
    uTraj: torch.Tensor

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class CostFunction(NamedTuple):
    """Cost function tuple: (runningCost, terminalCost); J = terminalCost(x[N]) + sum(runningCost(x[i],u[i]))
    -- pytrees.py:15-55"""
    runningCost: Callable
    terminalCost: Callable

    @classmethod
    def runningOnly(cls, runningCost, m: int = 1):
        from .models import QuadraticCost, QuadraticTerminalCost
        if isinstance(runningCost, QuadraticCost):
            return cls(runningCost, QuadraticTerminalCost(runningCost.Q))
        return cls(runningCost, lambda x: runningCost(x, torch.zeros(m, dtype=x.dtype, device=x.device)))

    def __call__(self, traj, k=None):
        runningCost, terminalCost = self
        xTraj, uTraj = traj
        if k is None:
            run = torch.stack([runningCost(xTraj[i], uTraj[i]) for i in range(uTraj.shape[0])])
            return torch.sum(run) + terminalCost(xTraj[-1])
        return runningCost(xTraj[k], uTraj[k])


class QuadraticValueFunction(NamedTuple):
    """(v, v_x, v_xx): V(x) = v + v_x.T x + 0.5 x.T v_xx x  -- pytrees.py:58-81"""
    v: torch.Tensor
    v_x: torch.Tensor
    v_xx: torch.Tensor

    def __call__(self, x):
        v, v_x, v_xx = self
        return v + v_x @ x + 0.5 * x @ v_xx @ x

    @classmethod
    def fromTerminalCostFunction(cls, costFun, xf):
        from .models import expand_terminal
        return cls(*expand_terminal(costFun.terminalCost, xf))


class QuadraticCostFunction(NamedTuple):
    """(c, c_x, c_u, c_xx, c_ux, c_uu)  -- pytrees.py:84-126"""
    c: torch.Tensor
    c_x: torch.Tensor
    c_u: torch.Tensor
    c_xx: torch.Tensor
    c_ux: torch.Tensor
    c_uu: torch.Tensor

    @classmethod
    def from_function(cls, costFun, x0, u0):
        from .models import expand_cost
        return cls(*expand_cost(costFun.runningCost, x0, u0))

    @classmethod
    def from_trajectory(cls, costFun, traj):
        from .models import expand_cost
        xTraj, uTraj = traj
        return cls(*expand_cost(costFun.runningCost, xTraj[..., :-1, :], uTraj))

    def __call__(self, x, u, k=None):
        c, c_x, c_u, c_xx, c_ux, c_uu = self
        if k is None and c.ndim != 0:
            raise ValueError("Must specify index for multi-dimensional cost")
        if k is None:
            return c + c_x @ x + c_u @ u + 0.5 * (x @ c_xx @ x + 2 * u @ c_ux @ x + u @ c_uu @ u)
        return self[k](x, u)

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class AffineDynamics(NamedTuple):
    """(f, f_x, f_u): xOut = f + f_x x + f_u u  -- pytrees.py:129-162"""
    f: torch.Tensor
    f_x: torch.Tensor
    f_u: torch.Tensor

    @classmethod
    def from_function(cls, dynFun, x0, u0):
        from .models import expand_dynamics
        return cls(*expand_dynamics(dynFun, x0, u0, second_order=False))

    @classmethod
    def from_trajectory(cls, dynFun, traj):
        from .models import expand_dynamics
        xTraj, uTraj = traj
        return cls(*expand_dynamics(dynFun, xTraj[..., :-1, :], uTraj, second_order=False))

    def __call__(self, x, u, k=None):
        f, f_x, f_u = self
        if k is None and f.ndim != 1:
            raise ValueError("Must specify index for multi-dimensional dynamics")
        return f + f_x @ x + f_u @ u if k is None else self[k](x, u)

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class QuadraticDynamics(NamedTuple):
    """(f, f_x, f_u, f_xx, f_ux, f_uu)  -- pytrees.py:165-204"""
    f: torch.Tensor
    f_x: torch.Tensor
    f_u: torch.Tensor
    f_xx: torch.Tensor
    f_ux: torch.Tensor
    f_uu: torch.Tensor

    @classmethod
    def from_function(cls, dynFun, x0, u0):
        from .models import expand_dynamics
        return cls(*expand_dynamics(dynFun, x0, u0, second_order=True))

    @classmethod
    def from_trajectory(cls, dynFun, traj):
        from .models import expand_dynamics
        xTraj, uTraj = traj
        return cls(*expand_dynamics(dynFun, xTraj[..., :-1, :], uTraj, second_order=True))

    def __call__(self, x, u, k=None):
        f, f_x, f_u, f_xx, f_ux, f_uu = self
        if k is None and f.ndim != 1:
            raise ValueError("Must specify index for trajectories")
        if k is None:
            quad = (torch.einsum('j,ijk,k->i', x, f_xx, x) + 2 * torch.einsum('j,ijk,k->i', u, f_ux, x) +
                    torch.einsum('j,ijk,k->i', u, f_uu, u))
            return f + f_x @ x + f_u @ u + 0.5 * quad
        return self[k](x, u)

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class AffinePolicy(NamedTuple):
    """(l, L): u = l + L x  -- pytrees.py:207-223"""
    l: torch.Tensor
    L: torch.Tensor

    def __call__(self, x, k=None, alpha=1):
        l, L = self
        if k is None and l.ndim != 1:
            raise ValueError("Must specify index for multi-dimensional policy")
        return alpha * l + L @ x if k is None else self[k](x, alpha=alpha)

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class QuadraticDeltaCost(NamedTuple):
    """dJ_exp = alpha * dJ_lin + alpha**2 * dJ_quad  -- pytrees.py:226-236"""
    dJ_lin: float
    dJ_quad: float

    def __call__(self, alpha):
        dJ_lin, dJ_quad = self
        return alpha * (dJ_lin + alpha * dJ_quad)
