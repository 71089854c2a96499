#!/usr/bin/env python
"""cfg 4 / cfg 5 problems with the reference's own stopping rule (maxIter=100, tol=1e-3): early exit + active-list compaction.
usage: bench_ilqr_defaults.py [ilqr|ddp] [Bsz]"""
import os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs, ilqrUtils
from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost
kind = sys.argv[1] if len(sys.argv) > 1 else "ilqr"
Bsz = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
d = configs.cfg4(Bsz=Bsz) if kind == "ilqr" else configs.cfg5(Bsz=Bsz)
dev = torch.device("cuda", 0)
x0 = torch.as_tensor(d["x0"], dtype=torch.float64, device=dev); uG = torch.as_tensor(d["uGuess"], dtype=torch.float64, device=dev)
solver = ilqrUtils.iterativeLqr if kind == "ilqr" else ilqrUtils.differentialDynamicProgramming
args = (QuadcopterEuler(d["dt"]), QuadraticCost(d["Q"], d["R"]), QuadraticTerminalCost(d["Qf"]))
for _ in range(2):
    *out, log = solver(*args, x0, uG, return_log=True); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
out = solver(*args, x0, uG)
e1.record(); torch.cuda.synchronize()
it = log["iters"].float()
print(f"{kind} defaults Bsz={Bsz}: {e0.elapsed_time(e1):.1f} ms, iterations mean {float(it.mean()):.1f} max {int(it.max())}, converged {float(out[3].float().mean()):.4f}")
