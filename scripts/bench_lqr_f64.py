#!/usr/bin/env python
"""discreteFiniteHorizonLqr / lqrMpc.solve throughput in fp64 vs fp32 on cfg 2 (time-invariant operands as stride-0 views)."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
from zopt_b200.mpcUtils import lqrMpc
from zopt_b200.quadcopter import Quadcopter
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = 50
d = configs.cfg2(Bsz=Bsz)
dev = torch.device("cuda", 0)
inf_n, inf_m = torch.full((12,), float("inf")), torch.full((4,), float("inf"))
for dt in (torch.float32, torch.float64):
    xbar = torch.as_tensor(d["xbar"], dtype=dt, device=dev); ubar = torch.as_tensor(d["ubar"], dtype=dt, device=dev)
    A, B = Quadcopter().linearizeInertial(xbar, ubar, 0.1)
    Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=dt, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=dt, device=dev))
    ex = lambda t: t[:, None].expand(-1, N, -1, -1)
    fns = {"discreteFiniteHorizonLqr": lambda: discreteFiniteHorizonLqr(ex(A), ex(B), ex(Q), ex(R), N),
           "lqrMpc.solve": lambda: lqrMpc(A, B, Q, R, N, -inf_n, inf_n, -inf_m, inf_m, Qf=10 * Q).solve(xbar)}
    for name, fn in fns.items():
        t0 = time.time()
        while time.time() - t0 < 0.5:
            out = fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"{name} {dt}: {ms:.3f} ms, {Bsz / ms * 1e3:.3e} solves/s")
