"""Minimal stand-in for the JAX entry points used by zopt's hot path (see ../README.md).  Test infrastructure only."""
import torch
import torch.utils._pytree as _pt
from torch import func as _F

torch.set_default_dtype(torch.float64)  # the reference always runs with jax_enable_x64 (zopt/quadcopter.py:7)

# numpy-style Tensor.transpose(*axes)
_orig_transpose = torch.Tensor.transpose


def _np_transpose(self, *axes):
    if len(axes) == 2:
        return _orig_transpose(self, *axes)
    if len(axes) == 1 and isinstance(axes[0], (tuple, list)):
        axes = tuple(axes[0])
    return self.permute(*axes)


torch.Tensor.transpose = _np_transpose

from . import numpy  # noqa: E402,F401
from . import lax  # noqa: E402,F401
from . import tree  # noqa: E402,F401
from . import experimental  # noqa: E402,F401


class _Config:
    def update(self, *a, **k):
        pass


config = _Config()


def jit(fun=None, static_argnames=None, static_argnums=None, **kw):
    if fun is None:
        return lambda f: f
    return fun


def vmap(fun, in_axes=0, out_axes=0):
    import numpy as _np
    mapped = _F.vmap(fun, in_dims=in_axes, out_dims=out_axes)
    return lambda *a, **k: mapped(*[torch.as_tensor(_np.ascontiguousarray(x)) if isinstance(x, _np.ndarray) else x for x in a], **k)


def jacobian(fun, argnums=0):
    return _F.jacrev(fun, argnums=argnums)


jacrev = jacobian


def jacfwd(fun, argnums=0):
    return _F.jacfwd(fun, argnums=argnums)


def hessian(fun, argnums=0):
    return _F.jacfwd(_F.jacrev(fun, argnums=argnums), argnums=argnums)


def grad(fun, argnums=0):
    return _F.grad(fun, argnums=argnums)
