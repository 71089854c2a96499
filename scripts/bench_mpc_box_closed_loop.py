#!/usr/bin/env python
"""Closed-loop box-constrained lqrMpc (demos/lqrMpc.py:42-47, batched, fused, warm-started): MPC solves per second.
usage: bench_mpc_box_closed_loop.py [Bsz] [f32|f64] [eps] [Tsim] [N] [check_termination] [auto|thread|quad|quad_global]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.mpcUtils import lqrMpc
from zopt_b200.quadcopter import Quadcopter
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dt = torch.float64 if (len(sys.argv) > 2 and sys.argv[2] == "f64") else torch.float32
eps = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-2
Tsim = int(sys.argv[4]) if len(sys.argv) > 4 else 200
N = int(sys.argv[5]) if len(sys.argv) > 5 else 25
dev = torch.device("cuda", 0)
A, B = Quadcopter().linearizeInertial(np.zeros(12), configs.U_TRIM, 0.1)
x_ub = np.array([1, 1, 1, 0.3, 0.3, 0.1, 0.5, 0.5, np.inf, np.inf, np.inf, np.inf]); u_ub = np.full(4, 3.0)
rng = np.random.default_rng(0)
x0 = np.zeros((Bsz, 12)); x0[:, 9:12] = rng.uniform(-10, 10, (Bsz, 3))
x0 = torch.as_tensor(x0, dtype=dt, device=dev)
prob = lqrMpc(A.to(dt), B.to(dt), torch.eye(12, dtype=dt, device=dev), torch.eye(4, dtype=dt, device=dev), N, -x_ub, x_ub, -u_ub, u_ub)
chk = int(sys.argv[6]) if len(sys.argv) > 6 else 25
kern = sys.argv[7] if len(sys.argv) > 7 else 'auto'
kw = dict(eps_abs=eps, eps_rel=eps, check_termination=chk, kernel=kern)
traj, st = prob.closedLoop(x0, Tsim, **kw); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); traj, st = prob.closedLoop(x0, Tsim, **kw); e1.record(); torch.cuda.synchronize()
el = e0.elapsed_time(e1) * 1e-3
it = prob.iters.float()
print(f"closed-loop box lqrMpc Bsz={Bsz} N={N} Tsim={Tsim} {dt} eps={eps} check={chk} kernel={kern}: {el*1e3:.1f} ms, {Bsz*Tsim/el:.3e} MPC solves/s, ADMM iters/solve mean {float(it.mean())/Tsim:.1f}, "
      f"final |pos| mean {float(traj.xTraj[:, -1, 9:12].norm(dim=1).mean()):.3f}, status counts {torch.bincount(st.long(), minlength=3).tolist()}")
