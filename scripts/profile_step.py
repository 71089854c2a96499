#!/usr/bin/env python
"""Short single-GPU driver for ncu: a few cfg-2 steps (linearise + lqrMpc.solve) at the bench batch size."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs  # noqa: E402
from zopt_b200.mpcUtils import lqrMpc  # noqa: E402
from zopt_b200.quadcopter import Quadcopter  # noqa: E402

Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
d = configs.cfg2(Bsz=Bsz)
dev = torch.device("cuda", 0)
f32 = torch.float32
xbar = torch.as_tensor(d["xbar"], dtype=f32, device=dev)
ubar = torch.as_tensor(d["ubar"], dtype=f32, device=dev)
Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev))
R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
inf_n, inf_m = torch.full((12,), float("inf")), torch.full((4,), float("inf"))
ac = Quadcopter()
for _ in range(steps):
    A, B = ac.linearizeInertial(xbar, ubar, d["dt"])
    u, traj, status = lqrMpc(A, B, Q, R, d["N"], -inf_n, inf_n, -inf_m, inf_m, Qf=10 * Q).solve(xbar)
torch.cuda.synchronize()
print("ok", float(u.abs().sum()))
