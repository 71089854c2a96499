#!/usr/bin/env python
"""Box-constrained lqrMpc (ADMM tier) throughput probe: demos/lqrMpc.py problem, batched initial states.
usage: bench_mpc_bounded.py [Bsz] [f64|f32] [eps] [auto|generic|thread|quad|quad_global] [N]"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200 import configs
from zopt_b200.mpcUtils import lqrMpc
from zopt_b200.quadcopter import Quadcopter
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
dt = torch.float32 if (len(sys.argv) > 2 and sys.argv[2] == "f32") else torch.float64
eps = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-3
kernel = sys.argv[4] if len(sys.argv) > 4 else "auto"
N = int(sys.argv[5]) if len(sys.argv) > 5 else 25
dev = torch.device("cuda", 0)
ac = Quadcopter()
A, B = ac.linearizeInertial(np.zeros(12), configs.U_TRIM, 0.1)
A, B = A.to(dt), B.to(dt)
x_ub = np.array([1, 1, 1, 0.3, 0.3, 0.1, 0.5, 0.5, np.inf, np.inf, np.inf, np.inf]); u_ub = np.full(4, 3.0)
rng = np.random.default_rng(0)
x0 = np.zeros((Bsz, 12)); x0[:, 9:12] = rng.uniform(-10, 10, (Bsz, 3))
x0 = torch.as_tensor(x0, dtype=dt, device=dev)
prob = lqrMpc(A, B, torch.eye(12, dtype=dt, device=dev), torch.eye(4, dtype=dt, device=dev), N, -x_ub, x_ub, -u_ub, u_ub)
kw = dict(eps_abs=eps, eps_rel=eps, kernel=kernel)
u, traj, st = prob.solve(x0, **kw); torch.cuda.synchronize()
reps = 3
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps): u, traj, st = prob.solve(x0, **kw)
e1.record(); torch.cuda.synchronize(); el = e0.elapsed_time(e1) * 1e-3 / reps
it = prob.iters.float()
print(f"bounded lqrMpc [{kernel}] Bsz={Bsz} N={N} {dt} eps={eps}: {el*1e3:.2f} ms, {Bsz/el:.3e} solves/s, ADMM iters mean {float(it.mean()):.0f} max {int(it.max())}, "
      f"{float(it.sum())/el:.3e} problem-iterations/s, status counts {torch.bincount(st.long(), minlength=3).tolist()}")
