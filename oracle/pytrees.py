"""
Oracle (test infrastructure): torch-CPU restatement of zopt/pytrees.py:1-236.

Field names, field order, `__getitem__` (slices every leaf) and `__call__`
semantics follow the reference one for one.  `jax.jacobian/hessian/grad/vmap`
are replaced by `torch.func.jacrev/hessian/grad/vmap` on fp64 CPU tensors.
"""
from typing import Callable, NamedTuple

import torch
from torch.func import grad, hessian, jacrev, vmap


def _map(fn, tup):
    return type(tup)(*(fn(leaf) for leaf in tup))


class Trajectory(NamedTuple):  # pytrees.py:6-12
    xTraj: torch.Tensor
    uTraj: torch.Tensor

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class CostFunction(NamedTuple):  # pytrees.py:15-55
    runningCost: Callable
    terminalCost: Callable

    @classmethod
    def runningOnly(cls, runningCost, m=1):  # pytrees.py:27-38
        return cls(runningCost, lambda x: runningCost(x, torch.zeros(m, dtype=x.dtype)))

    def __call__(self, traj, k=None):  # pytrees.py:40-55
        runningCost, terminalCost = self
        xTraj, uTraj = traj
        if k is None:
            return torch.sum(vmap(runningCost)(xTraj[:-1], uTraj)) + terminalCost(xTraj[-1])
        return runningCost(xTraj[k], uTraj[k])


class QuadraticValueFunction(NamedTuple):  # pytrees.py:58-81
    v: torch.Tensor
    v_x: torch.Tensor
    v_xx: torch.Tensor

    def __call__(self, x):
        v, v_x, v_xx = self
        return v + v_x @ x + 0.5 * x @ v_xx @ x

    @classmethod
    def fromTerminalCostFunction(cls, costFun, xf):  # pytrees.py:71-81
        cf = costFun.terminalCost
        return cls(cf(xf), grad(cf)(xf), hessian(cf)(xf))


class QuadraticCostFunction(NamedTuple):  # pytrees.py:84-126
    c: torch.Tensor
    c_x: torch.Tensor
    c_u: torch.Tensor
    c_xx: torch.Tensor
    c_ux: torch.Tensor
    c_uu: torch.Tensor

    @classmethod
    def from_function(cls, costFun, x0, u0):  # pytrees.py:99-107
        rc = costFun.runningCost
        c = rc(x0, u0)
        c_x, c_u = jacrev(rc, argnums=(0, 1))(x0, u0)
        ((c_xx, _), (c_ux, c_uu)) = hessian(rc, argnums=(0, 1))(x0, u0)
        return cls(c, c_x, c_u, c_xx, c_ux, c_uu)

    @classmethod
    def from_trajectory(cls, costFun, traj):  # pytrees.py:109-115
        xTraj, uTraj = traj
        return cls(*vmap(lambda x0, u0: tuple(cls.from_function(costFun, x0, u0)))(xTraj[:-1], uTraj))

    def __call__(self, x, u, k=None):  # pytrees.py:117-123
        c, c_x, c_u, c_xx, c_ux, c_uu = self
        if k is None and c.ndim != 0:
            raise ValueError("Must specify index for multi-dimensional cost")
        if k is None:
            return c + c_x @ x + c_u @ u + 0.5 * (x @ c_xx @ x + 2 * u @ c_ux @ x + u @ c_uu @ u)
        return self[k](x, u)

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class AffineDynamics(NamedTuple):  # pytrees.py:129-162
    f: torch.Tensor
    f_x: torch.Tensor
    f_u: torch.Tensor

    @classmethod
    def from_function(cls, dynFun, x0, u0):  # pytrees.py:138-144
        return cls(dynFun(x0, u0), jacrev(dynFun, 0)(x0, u0), jacrev(dynFun, 1)(x0, u0))

    @classmethod
    def from_trajectory(cls, dynFun, traj):  # pytrees.py:146-153
        xTraj, uTraj = traj
        return cls(*vmap(lambda x0, u0: tuple(cls.from_function(dynFun, x0, u0)))(xTraj[:-1], uTraj))

    def __call__(self, x, u, k=None):  # pytrees.py:155-159
        f, f_x, f_u = self
        if k is None and f.ndim != 1:
            raise ValueError("Must specify index for multi-dimensional dynamics")
        return f + f_x @ x + f_u @ u if k is None else self[k](x, u)

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class QuadraticDynamics(NamedTuple):  # pytrees.py:165-204
    f: torch.Tensor
    f_x: torch.Tensor
    f_u: torch.Tensor
    f_xx: torch.Tensor
    f_ux: torch.Tensor
    f_uu: torch.Tensor

    @classmethod
    def from_function(cls, dynFun, x0, u0):  # pytrees.py:179-185
        f = dynFun(x0, u0)
        f_x, f_u = jacrev(dynFun, argnums=(0, 1))(x0, u0)
        ((f_xx, _), (f_ux, f_uu)) = hessian(dynFun, argnums=(0, 1))(x0, u0)
        return cls(f, f_x, f_u, f_xx, f_ux, f_uu)

    @classmethod
    def from_trajectory(cls, dynFun, traj):  # pytrees.py:187-194
        xTraj, uTraj = traj
        return cls(*vmap(lambda x0, u0: tuple(cls.from_function(dynFun, x0, u0)))(xTraj[:-1], uTraj))

    def __call__(self, x, u, k=None):  # pytrees.py:196-201
        f, f_x, f_u, f_xx, f_ux, f_uu = self
        if k is None and f.ndim != 1:
            raise ValueError("Must specify index for trajectories")
        if k is None:
            # x.T @ f_xx @ x with f_xx (n,n,n): numpy/jax matmul semantics -> vector indexed by the FIRST axis
            quad = torch.einsum('j,ijk,k->i', x, f_xx, x) + 2 * torch.einsum('j,ijk,k->i', u, f_ux, x) \
                + torch.einsum('j,ijk,k->i', u, f_uu, u)
            return f + f_x @ x + f_u @ u + 0.5 * quad
        return self[k](x, u)

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class AffinePolicy(NamedTuple):  # pytrees.py:207-223
    l: torch.Tensor
    L: torch.Tensor

    def __call__(self, x, k=None, alpha=1):
        l, L = self
        if k is None and l.ndim != 1:
            raise ValueError("Must specify index for multi-dimensional policy")
        return alpha * l + L @ x if k is None else self[k](x, alpha=alpha)

    def __getitem__(self, k):
        return _map(lambda x: x[k], self)


class QuadraticDeltaCost(NamedTuple):  # pytrees.py:226-236
    dJ_lin: float
    dJ_quad: float

    def __call__(self, alpha):
        dJ_lin, dJ_quad = self
        return alpha * (dJ_lin + alpha * dJ_quad)
