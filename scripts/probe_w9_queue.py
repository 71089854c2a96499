#!/usr/bin/env python
"""Nine-lane closed-loop kernel: static mapping against the work-rotating mapping (ZB_W9_WORKERS_PER_SCHED, ZB_W9_CHUNK)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.mpcUtils import quadcopterClosedLoopMpc
dev = torch.device("cuda", 0); f32 = torch.float32
def run(Bsz, per, chunk=None, T=200):
    os.environ["ZB_W9_WORKERS_PER_SCHED"] = str(per)
    if chunk: os.environ["ZB_W9_CHUNK"] = str(chunk)
    else: os.environ.pop("ZB_W9_CHUNK", None)
    d = configs.cfg3(Bsz=Bsz)
    x = torch.as_tensor(d["xbar"], dtype=f32, device=dev); x[:, 9:12] *= 0.2
    Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
    out = quadcopterClosedLoopMpc(x, Q, R, 50, 20, Qf=10 * Q, variant="warp"); torch.cuda.synchronize()
    ms = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = quadcopterClosedLoopMpc(x, Q, R, 50, T, Qf=10 * Q, variant="warp"); e1.record(); torch.cuda.synchronize()
        ms = min(ms, e0.elapsed_time(e1))
    return ms, out
for Bsz in (1779, 2048, 2400, 3000, 4096):
    ms0, o0 = run(Bsz, 0)
    line = f"Bsz={Bsz}: static {ms0:.2f} ms"
    for per, chunk in ((1, 5), (1, 10), (1, 2), (2, 5)):
        if per == 2 and Bsz <= 3552: continue
        ms, o = run(Bsz, per, chunk)
        same = all(torch.equal(a, b) for a, b in zip(o0[:2], o[:2]))
        line += f" | per={per} chunk={chunk}: {ms:.2f} ms {'bit-identical' if same else 'DIFFERENT'}"
    print(line)
