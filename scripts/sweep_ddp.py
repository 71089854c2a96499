#!/usr/bin/env python
"""BASELINE cfg 5: throughput sweep of batched DDP (and iLQR) over the batch size on one GPU.
usage: sweep_ddp.py [N=100] [iters=5] [f64|f32] [max_log2=20]  -> one JSON line per batch size"""
import json, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs, ilqrUtils
from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost
N = int(sys.argv[1]) if len(sys.argv) > 1 else 100
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dt = torch.float32 if (len(sys.argv) > 3 and sys.argv[3] == "f32") else torch.float64
max_log2 = int(sys.argv[4]) if len(sys.argv) > 4 else 20
dev = torch.device("cuda", 0)
import time
_d = configs.cfg5(Bsz=16384, N=N)
_t0 = time.time()
while time.time() - _t0 < 1.5:  # a fresh process starts at idle clocks: ramp them up before the first (smallest, shortest) point
    ilqrUtils.iterativeLqr(QuadcopterEuler(_d["dt"]), QuadraticCost(_d["Q"], _d["R"]), QuadraticTerminalCost(_d["Qf"]),
                           torch.as_tensor(_d["x0"], dtype=dt, device=dev), torch.as_tensor(_d["uGuess"], dtype=dt, device=dev), maxIter=2, tol=-1.0)
    torch.cuda.synchronize()
for kind, solver in (("ddp", ilqrUtils.differentialDynamicProgramming), ("ilqr", ilqrUtils.iterativeLqr)):
    for lg in range(10, max_log2 + 1, 2):
        Bsz = 1 << lg
        d = configs.cfg5(Bsz=Bsz, N=N)
        x0 = torch.as_tensor(d["x0"], dtype=dt, device=dev)
        uG = torch.as_tensor(d["uGuess"], dtype=dt, device=dev)
        args = (QuadcopterEuler(d["dt"]), QuadraticCost(d["Q"], d["R"]), QuadraticTerminalCost(d["Qf"]))
        try:
            out = solver(*args, x0[: min(Bsz, 1024)], uG, maxIter=1, tol=-1.0)  # warm-up
            torch.cuda.synchronize()
            ms = float("inf")
            for _ in range(2):  # best of two: the first call at a new size pays the allocator's cudaMalloc of outputs and workspace
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                out = solver(*args, x0, uG, maxIter=iters, tol=-1.0)
                e1.record()
                torch.cuda.synchronize()
                ms = min(ms, e0.elapsed_time(e1))
            print(json.dumps({"solver": kind, "dtype": str(dt).split(".")[1], "N": N, "iters": iters, "batch": Bsz, "ms": ms,
                              "problem_iterations_per_s": Bsz * iters / (ms * 1e-3), "J_mean": float(out[2].mean())}), flush=True)
            del out
        except torch.OutOfMemoryError:
            print(json.dumps({"solver": kind, "batch": Bsz, "error": "out of memory"}), flush=True)
        torch.cuda.empty_cache()
