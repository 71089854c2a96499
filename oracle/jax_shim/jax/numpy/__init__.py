"""jax.numpy subset with NumPy semantics on torch-CPU fp64 tensors."""
import numpy as _np
import torch

from . import linalg  # noqa: F401

ndarray = torch.Tensor
pi = _np.pi
inf = _np.inf


def _t(x):
    if isinstance(x, torch.Tensor):
        return x
    a = _np.asarray(x)
    if a.dtype.kind == "f":
        a = a.astype(_np.float64)
    return torch.as_tensor(a)


def _has_tensor(x):
    if isinstance(x, torch.Tensor):
        return True
    if isinstance(x, (list, tuple)):
        return any(_has_tensor(e) for e in x)
    return False


def array(x, dtype=None):
    if isinstance(x, (list, tuple)) and _has_tensor(x):
        elems = [array(e) for e in x]
        elems = [e.to(torch.float64) if not e.is_floating_point() else e for e in elems]
        return torch.stack(elems)
    return _t(x)


asarray = array


def zeros(shape, dtype=None):
    return torch.zeros(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=torch.float64)


def ones(shape, dtype=None):
    return torch.ones(shape if isinstance(shape, (tuple, list)) else (shape,), dtype=torch.float64)


def eye(n):
    return torch.eye(n, dtype=torch.float64)


def arange(*a):
    return torch.arange(*a)


def linspace(a, b, num=50):
    return torch.linspace(a, b, num, dtype=torch.float64)


def interp(x, xp, fp, left=None, right=None, period=None):
    """np.interp for a scalar query (clipped at the ends), written so that jax.vmap over `fp` rows works"""
    xp, fp = _t(xp), _t(fp)
    xq = min(max(float(x), float(xp[0])), float(xp[-1]))
    i = min(max(int(torch.searchsorted(xp, torch.tensor(xq, dtype=xp.dtype), right=True)) - 1, 0), len(xp) - 2)
    w = (xq - float(xp[i])) / (float(xp[i + 1]) - float(xp[i]))
    return fp[i] * (1.0 - w) + fp[i + 1] * w


sin, cos, tan = (lambda x: torch.sin(_t(x))), (lambda x: torch.cos(_t(x))), (lambda x: torch.tan(_t(x)))


def concatenate(xs, axis=0):
    return torch.cat([_t(x) for x in xs], dim=axis)


def cross(a, b):
    return torch.linalg.cross(_t(a), _t(b))


def sum(x, axis=None):  # noqa: A001
    return torch.sum(_t(x)) if axis is None else torch.sum(_t(x), dim=axis)


def einsum(spec, *ops):
    return torch.einsum(spec, *[_t(o) for o in ops])


def maximum(a, b):
    if not isinstance(b, torch.Tensor):
        return torch.clamp(_t(a), min=b)
    return torch.maximum(_t(a), b)


def argmin(x):
    x = _t(x)
    nan = torch.isnan(x)
    if bool(nan.any()):  # NumPy/JAX: a NaN wins
        return torch.nonzero(nan)[0, 0]
    return torch.argmin(x)


def logical_not(x):
    if isinstance(x, torch.Tensor):
        return torch.logical_not(x)
    return not x


def block(rows):
    return torch.cat([torch.cat([_t(b) for b in row], dim=-1) for row in rows], dim=-2)


def repeat(x, n, axis=None):
    return torch.repeat_interleave(_t(x), n, dim=axis)


def all(x):  # noqa: A001
    return torch.all(_t(x))
