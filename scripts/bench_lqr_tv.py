#!/usr/bin/env python
"""discreteFiniteHorizonLqr throughput: time-invariant views vs materialised (time-varying) operands, (12,4) fp32."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
from zopt_b200.quadcopter import Quadcopter
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 32768
N = 50
d = configs.cfg2(Bsz=Bsz)
dev = torch.device("cuda", 0); f32 = torch.float32
xbar = torch.as_tensor(d["xbar"], dtype=f32, device=dev); ubar = torch.as_tensor(d["ubar"], dtype=f32, device=dev)
A, B = Quadcopter().linearizeInertial(xbar, ubar, 0.1)
Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
ex = lambda t: t[:, None].expand(-1, N, -1, -1)
def timeit(name, args):
    import time
    t0 = time.time()
    while time.time() - t0 < 0.5:  # clock ramp-up
        L = discreteFiniteHorizonLqr(*args, N); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): L = discreteFiniteHorizonLqr(*args, N)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name}: {ms:.3f} ms, {Bsz / ms * 1e3:.3e} solves/s")
    return L
L0 = timeit("time-invariant views (t1 kernel)", (ex(A), ex(B), ex(Q), ex(R)))
Qk = ex(Q).contiguous()
L1 = timeit("A,B,R views + materialised Q series", (ex(A), ex(B), Qk, ex(R)))
L2 = timeit("all operands materialised (streamed tv kernel)", (ex(A).contiguous(), ex(B).contiguous(), Qk, ex(R).contiguous()))
print("max |dL| vs t1:", float((L1 - L0).abs().max()), float((L2 - L0).abs().max()))
