#!/usr/bin/env python
"""One 2,048-problem closed loop (the shard one GPU gets from cfg 3 on eight GPUs) for ncu: nine-lane kernel with work rotation."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.mpcUtils import quadcopterClosedLoopMpc
dev = torch.device("cuda", 0); f32 = torch.float32
d = configs.cfg3(Bsz=2048)
x = torch.as_tensor(d["xbar"], dtype=f32, device=dev); x[:, 9:12] *= 0.2
Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
Qf = 10 * Q
for _ in range(2):
    quadcopterClosedLoopMpc(x, Q, R, 50, 40, Qf=Qf, variant="warp")
torch.cuda.synchronize()
