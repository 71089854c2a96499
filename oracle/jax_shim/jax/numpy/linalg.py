import torch


def solve(a, b):
    return torch.linalg.solve(a, b)


def inv(a):
    return torch.linalg.inv(a)


def eigh(a):
    # jnp.linalg.eigh defaults to symmetrize_input=True
    return torch.linalg.eigh(0.5 * (a + a.transpose(-1, -2)))
