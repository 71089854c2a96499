#!/usr/bin/env python
"""
Host-side copy roof of the end-to-end path (VERDICT r1 task 1): how fast can N ranks move the step's D2H payload
(214 MB of pinned memory per rank) at the same time, and what does CPU/NUMA placement change?

  torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 scripts/probe_pcie.py

Every rank: (a) unbound, (b) bound to its GPU's NUMA node before allocating (zopt_b200.hostbind), each with one 214 MB
pinned buffer pair; D2H and H2D, all ranks concurrently (barrier, CUDA events, max over ranks).  Rank 0 prints one JSON line
per configuration plus the host topology.  No solver kernels run here.
"""
import json
import os
import subprocess
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")
        dist.init_process_group("nccl", device_id=dev)
    from zopt_b200 import hostbind
    nbytes = 213975040  # d2h bytes per step of the bench (65,536 problems, fp32)
    src = torch.empty(nbytes, dtype=torch.uint8, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def measure(tag, ranks_active=None):
        act = ranks_active is None or rank in ranks_active
        hb = [torch.empty(nbytes, dtype=torch.uint8).pin_memory() for _ in range(2)] if act else None
        res = {}
        for direction in ("d2h", "h2d"):
            ms = 0.0
            if act:
                for h in hb:
                    (h.copy_(src, non_blocking=True) if direction == "d2h" else src.copy_(h, non_blocking=True))
            barrier()
            if act:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for i in range(10):
                    (hb[i % 2].copy_(src, non_blocking=True) if direction == "d2h" else src.copy_(hb[i % 2], non_blocking=True))
                e1.record()
            barrier()
            if act:
                ms = e0.elapsed_time(e1) / 10
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            allt = [torch.zeros_like(t) for _ in range(world)]
            if world > 1:
                dist.all_gather(allt, t)
            else:
                allt = [t]
            per = [float(x) for x in allt]
            n_act = sum(1 for p in per if p > 0)
            res[direction] = {"per_rank_gbs": [round(nbytes / (p * 1e-3) / 1e9, 1) if p > 0 else None for p in per],
                              "aggregate_gbs": round(n_act * nbytes / (max(per) * 1e-3) / 1e9, 1)}
        if rank == 0:
            print(json.dumps({"config": tag, "ranks": world, **res}), flush=True)
        del hb

    if rank == 0:
        topo = hostbind.host_topology()
        for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
            try:
                out = subprocess.run(cmd, capture_output=True, text=True, timeout=20).stdout
                print(out if cmd[0] == "nvidia-smi" else "\n".join(l for l in out.splitlines() if any(k in l for k in ("Model name", "Socket", "NUMA", "Thread", "Core", "CPU(s):"))), flush=True)
            except Exception as e:
                print("no", cmd, e)
        print(json.dumps({"host": topo}), flush=True)
    nodes = [None] * world
    me = {"rank": rank, "numa": hostbind.gpu_numa_node(local), "bus": hostbind.pci_bus_id(local), "cpus_allowed": len(os.sched_getaffinity(0))}
    if world > 1:
        dist.all_gather_object(nodes, me)
    else:
        nodes = [me]
    if rank == 0:
        print(json.dumps({"gpus": nodes}), flush=True)
    measure("unbound")
    if world > 1:
        measure("unbound, rank 0 alone", ranks_active=[0])
        measure("unbound, ranks 0-1", ranks_active=[0, 1])
        if world >= 8:
            measure("unbound, ranks 0-3", ranks_active=[0, 1, 2, 3])
            measure("unbound, ranks 0,2,4,6", ranks_active=[0, 2, 4, 6])
    b = hostbind.bind_to_gpu(local)
    allb = [None] * world
    if world > 1:
        dist.all_gather_object(allb, b)
    else:
        allb = [b]
    if rank == 0:
        print(json.dumps({"binding": allb}), flush=True)
    measure("bound to the GPU's NUMA node before pinned allocation")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
