#!/usr/bin/env python
"""Time-varying thread-per-problem Riccati kernel: static mapping against work rotation (ZB_T1_ROTATE, ZB_T1_CHUNK); gains must be
bit-identical.  (The same rotation on the time-invariant headline kernel was measured and dropped: DESIGN.md 4.1.)"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
from zopt_b200.quadcopter import Quadcopter
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
N = 50
d = configs.cfg2(Bsz=Bsz)
dev = torch.device("cuda", 0); f32 = torch.float32
xbar = torch.as_tensor(d["xbar"], dtype=f32, device=dev); ubar = torch.as_tensor(d["ubar"], dtype=f32, device=dev)
A, B = Quadcopter().linearizeInertial(xbar, ubar, 0.1)
Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
ex = lambda t: t[:, None].expand(-1, N, -1, -1).contiguous()
args = (ex(A), ex(B), ex(Q), ex(R))
def run(rot, chunk=None):
    os.environ["ZB_T1_ROTATE"] = str(rot)
    if chunk: os.environ["ZB_T1_CHUNK"] = str(chunk)
    else: os.environ.pop("ZB_T1_CHUNK", None)
    for _ in range(3): L = discreteFiniteHorizonLqr(*args, N)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): L = discreteFiniteHorizonLqr(*args, N)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / 5, L
ms0, L0 = run(0)
print(f"time-varying, Bsz={Bsz}: static {ms0:.3f} ms")
for chunk in (25, 13, 10, 5):
    ms, L = run(1, chunk)
    print(f"   rotate chunk={chunk}: {ms:.3f} ms  {'bit-identical' if torch.equal(L, L0) else 'DIFFERENT'}")

