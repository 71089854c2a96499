#!/usr/bin/env python
"""User-defined (symbolic) model through the solver plug-in: iLQR / DDP problem-iterations per second.
usage: bench_plugin.py [pendulum|car] [ilqr|ddp] [Bsz] [N] [iters] [f64|f32]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tests import plugin_models
from zopt_b200 import ilqrUtils
from zopt_b200.models import QuadraticCost, QuadraticTerminalCost
name = sys.argv[1] if len(sys.argv) > 1 else "car"
kind = sys.argv[2] if len(sys.argv) > 2 else "ilqr"
Bsz = int(sys.argv[3]) if len(sys.argv) > 3 else 65536
N = int(sys.argv[4]) if len(sys.argv) > 4 else 50
iters = int(sys.argv[5]) if len(sys.argv) > 5 else 5
dt = torch.float32 if (len(sys.argv) > 6 and sys.argv[6] == "f32") else torch.float64
mdl = plugin_models.build(name)
n, m = mdl.n, mdl.m
rng = np.random.default_rng(3)
x0 = torch.as_tensor(rng.uniform(-1, 1, (Bsz, n)), dtype=dt, device="cuda")
uG = torch.zeros((N, m), dtype=dt, device="cuda")
solver = ilqrUtils.iterativeLqr if kind == "ilqr" else ilqrUtils.differentialDynamicProgramming
args = (mdl, QuadraticCost(np.eye(n), 0.5 * np.eye(m)), QuadraticTerminalCost(10 * np.eye(n)))
t0 = time.time()
while time.time() - t0 < 1.0:
    out = solver(*args, x0, uG, maxIter=iters, tol=-1.0); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): out = solver(*args, x0, uG, maxIter=iters, tol=-1.0)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
print(f"{name} (n={n}, m={m}) {kind} Bsz={Bsz} N={N} iters={iters} {dt}: {ms:.2f} ms, {Bsz * iters / ms * 1e3:.3e} problem-iterations/s, J mean {float(out[2].mean()):.6f}")
