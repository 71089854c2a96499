import torch.utils._pytree as _pt


def map(f, tree, *rest):  # noqa: A001
    return _pt.tree_map(f, tree, *rest)
