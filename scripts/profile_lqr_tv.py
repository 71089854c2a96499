import os, sys, torch
sys.path.insert(0, '/root/repo')
from zopt_b200 import configs
from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
from zopt_b200.quadcopter import Quadcopter
Bsz, N = 65536, 50
d = configs.cfg2(Bsz=Bsz)
dev = torch.device("cuda", 0); f32 = torch.float32
xbar = torch.as_tensor(d["xbar"], dtype=f32, device=dev); ubar = torch.as_tensor(d["ubar"], dtype=f32, device=dev)
A, B = Quadcopter().linearizeInertial(xbar, ubar, 0.1)
Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev)); R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
ex = lambda t: t[:, None].expand(-1, N, -1, -1).contiguous()
args = (ex(A), ex(B), ex(Q), ex(R))
for _ in range(3):
    L = discreteFiniteHorizonLqr(*args, N); torch.cuda.synchronize()
