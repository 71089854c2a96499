#!/usr/bin/env python
"""cfg 3 closed loop in fp64 (fused cooperative kernel) vs fp32: 16,384 problems x 200 steps, horizon 50."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.mpcUtils import quadcopterClosedLoopMpc
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
d = configs.cfg3(Bsz=Bsz)
for dt in (torch.float32, torch.float64):
    x0 = torch.as_tensor(d["xbar"], dtype=dt, device="cuda"); x0[:, 9:12] *= 0.2
    Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=dt, device="cuda")); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=dt, device="cuda"))
    t0 = time.time()
    while time.time() - t0 < 0.7:
        out = quadcopterClosedLoopMpc(x0, Q, R, 50, 200, dt=0.1, Qf=10 * Q); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): out = quadcopterClosedLoopMpc(x0, Q, R, 50, 200, dt=0.1, Qf=10 * Q)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    print(f"{dt}: {ms:.2f} ms, {Bsz * 200 / ms * 1e3:.3e} MPC solves/s")
