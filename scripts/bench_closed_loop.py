#!/usr/bin/env python
"""Closed-loop LQR-MPC (cfg 3) timing for the three fp32 kernel variants over the batch size."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.mpcUtils import quadcopterClosedLoopMpc
dev = torch.device("cuda", 0); f32 = torch.float32
for Bsz in (512, 2048, 4096, 8192, 16384, 65536):
    torch.manual_seed(0)
    d = configs.cfg3(Bsz=Bsz)
    x = torch.as_tensor(d["xbar"], dtype=f32, device=dev); x[:, 9:12] *= 0.2
    Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
    Qf = 10 * Q
    for variant in ("thread", "quad", "warp"):
        if variant == "warp" and Bsz > 16384:
            continue
        quadcopterClosedLoopMpc(x, Q, R, 50, 20, Qf=Qf, variant=variant); torch.cuda.synchronize()
        ms = 1e9
        for _ in range(3):  # best of three
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); quadcopterClosedLoopMpc(x, Q, R, 50, 200, Qf=Qf, variant=variant); e1.record(); torch.cuda.synchronize()
            ms = min(ms, e0.elapsed_time(e1))
        print(f"Bsz={Bsz:6d} {variant:6s}: {ms:8.2f} ms  {Bsz * 200 / ms * 1e3:.3e} MPC solves/s")
