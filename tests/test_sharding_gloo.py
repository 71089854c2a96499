"""
Multi-GPU path on CPU: world_size-2 gloo.  The batch is split by rank with no data-path collective; each rank solves
its shard (here with the host build of the generic LQR kernel body, tests/hostsim -- test infrastructure), the results
are gathered and must equal the single-process solve bit for bit.
"""
import ctypes as C
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zopt_b200.sharding import gather, shard, shard_range


def test_shard_range_partitions_exactly():
    for Bsz in (0, 1, 7, 16, 65536, 16385):
        for world in (1, 2, 3, 8):
            spans = [shard_range(Bsz, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == Bsz
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _problem(Bsz):
    rng = np.random.default_rng(5)
    n, m, N = 4, 2, 6
    A = rng.normal(size=(Bsz, N, n, n)) * 0.4
    B = rng.normal(size=(Bsz, N, n, m))
    Q = np.repeat(np.repeat(np.eye(n)[None, None], N, 1), Bsz, 0) * (1 + rng.uniform(size=(Bsz, 1, 1, 1)))
    R = np.repeat(np.repeat(np.eye(m)[None, None], N, 1), Bsz, 0)
    return A, B, Q, R, N, n, m


def _solve(A, B, Q, R, N, n, m):
    from tests import hostsim as H
    Bsz = A.shape[0]
    A, B, Q, R = (np.ascontiguousarray(a) for a in (A, B, Q, R))
    L = np.zeros((Bsz, N, m, n))
    if Bsz:
        zs = [H.arr(a, 2) for a in (A, B, Q, R)]
        H.hs.hs_lqr_dfh(1, C.c_int64(Bsz), N, N, n, m, *[C.byref(z) for z in zs], H.P(L), None)
    return L


def _worker(rank, world, port, Bsz, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    A, B, Q, R, N, n, m = _problem(Bsz)
    mine = [shard(torch.as_tensor(a), rank, world).numpy() for a in (A, B, Q, R)]
    L_local = torch.as_tensor(_solve(*mine, N, n, m))
    L_all = gather(L_local, Bsz)
    # timing reduction used by bench.py: max over ranks
    t = torch.tensor([float(rank + 1)])
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        torch.save(dict(L=L_all, tmax=float(t)), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("Bsz", [9, 16])
def test_two_rank_sharded_solve_equals_single_process(tmp_path, Bsz):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "gathered.pt")
    mp.spawn(_worker, args=(2, port, Bsz, out), nprocs=2, join=True)
    got = torch.load(out)
    ref = _solve(*_problem(Bsz))
    assert got["tmax"] == 2.0
    assert got["L"].shape == ref.shape and np.array_equal(got["L"].numpy(), ref)
