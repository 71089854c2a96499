#!/usr/bin/env python
"""Short single-GPU driver for ncu: one closed-loop LQR-MPC launch (cfg 3) -- argv: batch, variant, sim steps, dtype."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs  # noqa: E402
from zopt_b200.mpcUtils import quadcopterClosedLoopMpc  # noqa: E402

Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
variant = sys.argv[2] if len(sys.argv) > 2 else "warp"
Tsim = int(sys.argv[3]) if len(sys.argv) > 3 else 20
dt_ = torch.float64 if (len(sys.argv) > 4 and sys.argv[4] == "f64") else torch.float32
dev = torch.device("cuda", 0)
d = configs.cfg3(Bsz=Bsz)
x = torch.as_tensor(d["xbar"], dtype=dt_, device=dev)
x[:, 9:12] *= 0.2
Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=dt_, device=dev))
R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=dt_, device=dev))
for _ in range(2):
    tr = quadcopterClosedLoopMpc(x, Q, R, 50, Tsim, Qf=10 * Q, variant=variant)
torch.cuda.synchronize()
print("ok", float(tr.uTraj.abs().sum()))
