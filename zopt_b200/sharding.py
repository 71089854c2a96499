"""
Batch sharding across the GPUs of one box.  Every problem of the hot path is independent, so the batch axis
is split contiguously by rank and there is NO collective on the data path; the only communication is the optional final
gather of results (costs, trajectories), done here with `torch.distributed.all_gather` (NCCL over NVLink on the GPU box,
gloo in the CPU tests).  Host logic only -- no arithmetic.

Two ways to use the 8 GPUs of a box:
  * one process per GPU (`torchrun`, what bench.py does): `shard_range` / `shard` pick the rank's slice, `gather` collects;
  * ONE process driving every visible GPU: `solve_sharded(fn, args, ...)` -- a host thread and a CUDA stream per device, the
    batch split contiguously, optionally in chunks so that the device->host copy of chunk c overlaps the solve of chunk
    c+1, results landing in pinned host buffers in batch order (the "final gather" as host-side concatenation).
"""
import threading

import torch
import torch.distributed as dist


def shard_range(Bsz, rank, world):
    """Contiguous split of range(Bsz): rank r owns [lo, hi); sizes differ by at most one, earlier ranks get the extra."""
    base, rem = divmod(int(Bsz), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard(t, rank, world, dim=0):
    """This rank's slice of a batched array (a view, no copy)."""
    lo, hi = shard_range(t.shape[dim], rank, world)
    return t.narrow(dim, lo, hi - lo) if isinstance(t, torch.Tensor) else t[(slice(None),) * dim + (slice(lo, hi),)]


def gather(local, Bsz, group=None):
    """Concatenate per-rank results (ragged shards allowed) in rank order on every rank."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    sizes = [shard_range(Bsz, r, world) for r in range(world)]
    maxn = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((maxn,) + tuple(local.shape[1:]))
    pad[:local.shape[0]] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(outs, sizes)], dim=0)


# ------------------------------------------------------------------------------------------------ one process, all GPUs
def _flatten(out):
    """leaves of a (nested) tuple / list / NamedTuple of tensors, plus a function that rebuilds the structure"""
    if isinstance(out, torch.Tensor):
        return [out], lambda leaves: leaves[0]
    if isinstance(out, (tuple, list)):
        parts = [_flatten(o) for o in out]
        sizes = [len(p[0]) for p in parts]
        leaves = [l for p in parts for l in p[0]]

        def rebuild(ls, parts=parts, sizes=sizes, typ=type(out)):
            res, i = [], 0
            for (_, rb), k in zip(parts, sizes):
                res.append(rb(ls[i:i + k]))
                i += k
            return typ(*res) if hasattr(typ, "_fields") else typ(res)
        return leaves, rebuild
    raise TypeError(f"solve_sharded: unsupported output leaf {type(out)}")


def solve_sharded(fn, batched_args, shared_args=(), devices=None, chunks=1, bind_cpus=True):
    """
    Run `fn(*batched_slice, *shared)` on every visible GPU from ONE process and return its outputs for the whole batch as
    pinned host tensors in batch order (same pytree structure as `fn` returns; every output leaf must be batched).

    batched_args : tensors / arrays with the same leading batch size (host, ideally pinned, or on any device); each
                   device gets a contiguous slice (`shard_range`), moved with non-blocking copies on that device's stream
    shared_args  : passed to every call unchanged (scalars, models, un-batched weights)
    chunks       : split every device's shard in this many pieces; the D2H copy of piece c overlaps the solve of piece c+1
    bind_cpus    : pin each worker thread to the CPUs local to its GPU before it touches pinned memory (hostbind)

    Problems are independent, so there is no collective: per device a stream, a thread, and nothing shared but the output
    buffers (disjoint row ranges).  Timing such a call: wall clock around it, or CUDA events per device stream.
    """
    if not torch.cuda.is_available():
        raise RuntimeError("zopt_b200 needs a CUDA device; there is no CPU fallback")
    devices = list(range(torch.cuda.device_count())) if devices is None else [int(d) for d in devices]
    batched_args = [a if isinstance(a, torch.Tensor) else torch.as_tensor(a) for a in batched_args]
    Bsz = int(batched_args[0].shape[0])
    if any(int(a.shape[0]) != Bsz for a in batched_args):
        raise ValueError("solve_sharded: batched arguments disagree on the batch size")
    world = len(devices)
    state = {"out": None, "rebuild": None, "err": None}
    lock = threading.Lock()

    def worker(rank, d):
        try:
            if bind_cpus:
                from . import hostbind
                hostbind.bind_to_gpu(d)
            torch.cuda.set_device(d)
            dev = torch.device("cuda", d)
            lo, hi = shard_range(Bsz, rank, world)
            stream, copy_stream = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            nch = max(1, min(int(chunks), hi - lo)) if hi > lo else 0
            with torch.cuda.stream(stream):
                for c in range(nch):
                    clo, chi = shard_range(hi - lo, c, nch)
                    clo, chi = lo + clo, lo + chi
                    args = [a[clo:chi].to(dev, non_blocking=True) for a in batched_args]
                    leaves, rebuild = _flatten(fn(*args, *shared_args))
                    with lock:
                        if state["out"] is None:
                            state["out"] = [torch.empty((Bsz,) + tuple(l.shape[1:]), dtype=l.dtype).pin_memory() for l in leaves]
                            state["rebuild"] = rebuild
                    ready = torch.cuda.Event()
                    ready.record(stream)
                    copy_stream.wait_event(ready)
                    with torch.cuda.stream(copy_stream):
                        for h, l in zip(state["out"], leaves):
                            if l.shape[0] != chi - clo:
                                raise ValueError("solve_sharded: every output leaf must carry the batch axis first")
                            l.record_stream(copy_stream)
                            h[clo:chi].copy_(l, non_blocking=True)
            stream.synchronize()
            copy_stream.synchronize()
        except BaseException as e:  # surfaced by the caller
            state["err"] = e

    threads = [threading.Thread(target=worker, args=(r, d)) for r, d in enumerate(devices)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if state["err"] is not None:
        raise state["err"]
    return state["rebuild"](state["out"])
