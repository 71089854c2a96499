#!/usr/bin/env python
"""Small invocations of every kernel touched in the third session, for compute-sanitizer (memcheck / racecheck)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs, ilqrUtils
from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost
from zopt_b200.mpcUtils import lqrMpc, quadcopterClosedLoopMpc
from zopt_b200.quadcopter import Quadcopter
dev = "cuda"
for dt in (torch.float64, torch.float32):
    for kind in ("ilqr", "ddp"):
        d = configs.cfg4(Bsz=21, N=9) if kind == "ilqr" else configs.cfg5(Bsz=21, N=9)
        solver = ilqrUtils.iterativeLqr if kind == "ilqr" else ilqrUtils.differentialDynamicProgramming
        out = solver(QuadcopterEuler(d["dt"]), QuadraticCost(d["Q"], d["R"]), QuadraticTerminalCost(d["Qf"]),
                     torch.as_tensor(d["x0"], dtype=dt, device=dev), torch.as_tensor(d["uGuess"], dtype=dt, device=dev), maxIter=2, tol=-1.0)
        print(kind, dt, float(out[2].mean()))
    d = configs.cfg2(Bsz=45)
    N = 7
    xbar, ubar = torch.as_tensor(d["xbar"], dtype=dt, device=dev), torch.as_tensor(d["ubar"], dtype=dt, device=dev)
    A, B = Quadcopter().linearizeInertial(xbar, ubar, 0.1)
    Q, R = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=dt, device=dev)), torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=dt, device=dev))
    ex = lambda t: t[:, None].expand(-1, N, -1, -1)
    L1 = discreteFiniteHorizonLqr(ex(A), ex(B), ex(Q), ex(R), N)
    L2 = discreteFiniteHorizonLqr(ex(A).contiguous(), ex(B).contiguous(), ex(Q).contiguous(), ex(R).contiguous(), N)
    inf_n, inf_m = torch.full((12,), float("inf")), torch.full((4,), float("inf"))
    for QQ in (Q, Q + 0.01):
        u, traj, st = lqrMpc(A, B, QQ, R, N, -inf_n, inf_n, -inf_m, inf_m, Qf=10 * QQ).solve(xbar)
        x0 = xbar.clone(); x0[:, 9:12] *= 0.2
        cl = quadcopterClosedLoopMpc(x0, QQ, R, N, 4, dt=0.1, Qf=10 * QQ)
    print("lqr/mpc", dt, float(L1.abs().sum()), float(L2.abs().sum()), float(u.abs().sum()), float(cl.xTraj.abs().sum()))
torch.cuda.synchronize()
print("done")
