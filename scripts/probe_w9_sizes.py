import os, sys
import torch
sys.path.insert(0, '/root/repo')
from zopt_b200 import configs
from zopt_b200.mpcUtils import quadcopterClosedLoopMpc
dev = torch.device("cuda", 0); f32 = torch.float32
for Bsz in (444, 888, 1332, 1776, 1779, 2048, 2220, 2664, 3552):
    d = configs.cfg3(Bsz=Bsz)
    x = torch.as_tensor(d["xbar"], dtype=f32, device=dev); x[:, 9:12] *= 0.2
    Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
    Qf = 10 * Q
    quadcopterClosedLoopMpc(x, Q, R, 50, 20, Qf=Qf, variant="warp"); torch.cuda.synchronize()
    ms = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); quadcopterClosedLoopMpc(x, Q, R, 50, 200, Qf=Qf, variant="warp"); e1.record(); torch.cuda.synchronize()
        ms = min(ms, e0.elapsed_time(e1))
    print(f"Bsz={Bsz:6d} warps={(Bsz+2)//3:5d} per-SM {((Bsz+2)//3)/148:.2f}: {ms:8.2f} ms  cycles/step {ms*1e-3*1.965e9/10000:.0f}")
