import torch
import torch.utils._pytree as _pt


def _items(xs):
    if isinstance(xs, torch.Tensor) and xs.ndim == 1 and not xs.is_floating_point():
        return [int(v) for v in xs]
    return list(xs)


def scan(f, init, xs=None, length=None, reverse=False):
    items = _items(xs) if xs is not None else [None] * length
    order = range(len(items) - 1, -1, -1) if reverse else range(len(items))
    carry, ys = init, [None] * len(items)
    for i in order:
        carry, ys[i] = f(carry, items[i])
    if len(ys) == 0 or ys[0] is None:
        return carry, None
    stacked = _pt.tree_map(lambda *leaves: torch.stack([torch.as_tensor(l) for l in leaves]), *ys)
    return carry, stacked


def while_loop(cond_fun, body_fun, init_val):
    val = init_val
    while bool(cond_fun(val)):
        val = body_fun(val)
    return val
