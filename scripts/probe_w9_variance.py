#!/usr/bin/env python
"""Run-to-run spread of the nine-lane closed-loop kernel at one shard of cfg 3 (2,048 problems, 200 steps, horizon 50): default
(work rotation) against the static mapping (ZB_W9_WORKERS_PER_SCHED=0); every launch timed on its own (CUDA events around the call
and the host time of the call), with and without a device synchronisation between launches."""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.mpcUtils import quadcopterClosedLoopMpc
dev = torch.device("cuda", 0); f32 = torch.float32
d = configs.cfg3(Bsz=2048)
x = torch.as_tensor(d["xbar"], dtype=f32, device=dev)
Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
REPS = int(sys.argv[1]) if len(sys.argv) > 1 else 100
for label, env, sync in (("default", None, True), ("default", None, False), ("static", "0", True), ("static", "0", False)):
    if env is None: os.environ.pop("ZB_W9_WORKERS_PER_SCHED", None)
    else: os.environ["ZB_W9_WORKERS_PER_SCHED"] = env
    quadcopterClosedLoopMpc(x, Q, R, 50, 200, dt=0.1, Qf=10 * Q); torch.cuda.synchronize()
    ev, host = [], []
    for _ in range(REPS):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(); quadcopterClosedLoopMpc(x, Q, R, 50, 200, dt=0.1, Qf=10 * Q); e1.record()
        host.append((time.perf_counter() - t0) * 1e3)
        if sync: torch.cuda.synchronize()
        ev.append((e0, e1))
    torch.cuda.synchronize()
    ts = sorted((a.elapsed_time(b), h) for (a, b), h in zip(ev, host))
    med = ts[len(ts) // 2][0]
    out = [f"{t:.2f} (host {h:.2f})" for t, h in ts if t > 1.15 * med]
    print(f"{label} sync={sync}: median {med:.2f} ms, min {ts[0][0]:.2f}, max {ts[-1][0]:.2f}, host median {sorted(host)[len(host) // 2]:.2f} ms; outliers: {out}", flush=True)
