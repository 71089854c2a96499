#!/usr/bin/env python
"""iLQR / DDP throughput probe (BASELINE cfg 4 / 5): problem-iterations per second.
usage: bench_ilqr.py [ilqr|ddp] [Bsz] [N] [iters] [f64|f32]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs, ilqrUtils  # noqa: E402
from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "ilqr"
Bsz = int(sys.argv[2]) if len(sys.argv) > 2 else 16384
N = int(sys.argv[3]) if len(sys.argv) > 3 else 200
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 10
dt = torch.float32 if (len(sys.argv) > 5 and sys.argv[5] == "f32") else torch.float64
d = configs.cfg4(Bsz=Bsz, N=N) if kind == "ilqr" else configs.cfg5(Bsz=Bsz, N=N)
dev = torch.device("cuda", 0)
x0 = torch.as_tensor(d["x0"], dtype=dt, device=dev)
uG = torch.as_tensor(d["uGuess"], dtype=dt, device=dev)
solver = ilqrUtils.iterativeLqr if kind == "ilqr" else ilqrUtils.differentialDynamicProgramming
args = (QuadcopterEuler(d["dt"]), QuadraticCost(d["Q"], d["R"]), QuadraticTerminalCost(d["Qf"]))
reps = int(os.environ.get("REPS", 2))
t_warm = time.time()
while time.time() - t_warm < float(os.environ.get("WARM_S", 1.0)):  # let the clocks ramp up: a fresh process starts at idle clocks
    out = solver(*args, x0, uG, maxIter=iters, tol=-1.0)
    torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    out = solver(*args, x0, uG, maxIter=iters, tol=-1.0)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{kind} Bsz={Bsz} N={N} iters={iters} {dt}: {ms:.2f} ms/solve-batch, {Bsz * iters / (ms * 1e-3):.3e} problem-iterations/s, "
      f"J mean {float(out[2].mean()):.6f}")
