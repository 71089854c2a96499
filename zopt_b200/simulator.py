"""
Closed-loop simulation -- mirror of the DISCRETE half of zopt/simulator.py (SimBlock :9-39, Simulator :48-169) for batches.

The reference wires two blocks, `(u, xCtrl') = controller.update(k, xCtrl, x)` and `(y, x') = dynamics.update(k, x, u)`
(simulator.py:66-80, :124-138), and steps them in a Python loop.  Its block functions are arbitrary Python callables; a
kernel cannot trace those, so -- like the iLQR models -- the blocks here are registered kinds that are still callable the way
the demos' lambdas are:

  * `TrackingController(xTraj, uTraj, LArr)`         u = LArr[k] (x - xTraj[k]) + uTraj[k]      demos/iterativeLqr.py:16-17,47-52
  * `ProportionalFeedbackController(x0, u0, K, ns)`  u = -K[k] (x[:ns] - x0) + u0               zopt/lqrUtils.py:266-269,
                                                                                                  demos/discreteFiniteHorizonLqr.py:41-47
  * the dynamics block wraps a registered model: `QuadcopterEuler(dt, wind_ned)` (demos/iterativeLqr.py:46) or `LinearDynamics`.

Both controllers are stateless (xCtrl0 = [] in every demo), so the loop is `trajectoryRollout` (zopt/ilqrUtils.py:33-66) with
l = 0: one launch of the rollout kernel (zb_ilqr_rollout) simulates every problem of the batch.  The continuous-time path
(solve_ivp, simulator.py:153-157) is out of scope.
"""
import numpy as np
import torch

from ._lib import pick_device, pick_dtype, to_dev
from .ilqrUtils import trajectoryRollout
from .models import require_model
from .pytrees import AffinePolicy, Trajectory


class TrackingController:
    """u = LArr[k] @ (x - xTraj[k]) + uTraj[k]  (demos/iterativeLqr.py:16-17); every array optionally batched"""

    def __init__(self, xTraj, uTraj, LArr):
        self.xTraj, self.uTraj, self.LArr = xTraj, uTraj, LArr

    def __call__(self, k, xCtrl, x):
        return self.LArr[..., k, :, :] @ (x - self.xTraj[..., k, :]) + self.uTraj[..., k, :], xCtrl

    def arrays(self):
        return (self.xTraj, self.uTraj, self.LArr)

    def as_policy(self, n, N, dtype, device):
        xT, uT, L = (to_dev(t, dtype, device) for t in (self.xTraj, self.uTraj, self.LArr))
        return L[..., :N, :, :], xT[..., :N + 1, :], uT[..., :N, :]


class ProportionalFeedbackController:
    """u = -K[k] @ (x[:ns] - x0) + u0 (zopt/lqrUtils.py:266-269 as used in demos/discreteFiniteHorizonLqr.py:41-47, where
    an 8-state gain drives the 12-state plant through `x[:8]`); K (N,m,ns), x0 (ns,), u0 (m,), optionally batched"""

    def __init__(self, x0, u0, K, ns=None):
        self.x0, self.u0, self.K = x0, u0, K
        self.ns = int(ns) if ns is not None else int(K.shape[-1])

    def __call__(self, k, xCtrl, x):
        return -self.K[..., k, :, :] @ (x[..., :self.ns] - self.x0) + self.u0, xCtrl

    def arrays(self):
        return (self.x0, self.u0, self.K)

    def as_policy(self, n, N, dtype, device):
        x0, u0, K = (to_dev(t, dtype, device) for t in (self.x0, self.u0, self.K))
        K = K[..., :N, :, :]
        L = torch.zeros(K.shape[:-1] + (n,), dtype=dtype, device=device)
        L[..., :self.ns] = -K
        xref = torch.zeros(x0.shape[:-1] + (n,), dtype=dtype, device=device)
        xref[..., :self.ns] = x0
        xT = xref.unsqueeze(-2).expand(xref.shape[:-1] + (N + 1, n))
        uT = u0.unsqueeze(-2).expand(u0.shape[:-1] + (N, u0.shape[-1]))
        return L, xT, uT


class SimBlock():

    def __init__(self, fun, x0, dt=0, jittable=True, name=None):
        """
        Create a simulation block (zopt/simulator.py:11-39).  `fun` is a registered controller (above) or a registered
        dynamics model (zopt_b200.models); x0 the initial block state, (nx,) or (Bsz,nx); dt the sample time (> 0).
        """
        self.update = fun
        self.dt = dt
        self.jittable = jittable
        self.x0 = x0
        self.nx = int(np.shape(x0)[-1]) if np.ndim(x0) else 0
        self.name = name


class Simulator():

    def __init__(self, blocks, t_span, method="RK45", t_eval=None):
        """Two blocks, [controller, dynamics] with state feedback (zopt/simulator.py:50-86); discrete time only."""
        assert len(blocks) == 2, "Currently only supports 2 simBlocks."
        assert len(set([block.dt for block in blocks])) == 1, "Multi-sample time not implemented yet."
        self.blocks, self.t_span = blocks, t_span
        self.dt = blocks[0].dt
        if self.dt == 0:
            raise NotImplementedError("continuous-time simulation (scipy solve_ivp, zopt/simulator.py:153-157) is out of scope")
        ctrl, dyn = blocks
        if not hasattr(ctrl.update, "as_policy"):
            raise TypeError("controller block must be a TrackingController or ProportionalFeedbackController "
                            "(arbitrary callables cannot run in a kernel; there is no CPU fallback)")
        if ctrl.nx != 0:
            raise NotImplementedError("controllers with state are not supported (every reference demo uses xCtrl0 = [])")
        self.model = require_model(dyn.update)

    def simulate(self):
        """
        Run the simulation (zopt/simulator.py:140-169).  Returns (tArr (N+1,), x0Arr ([Bsz,]N+1,0), x1Arr ([Bsz,]N+1,n),
        y0Arr ([Bsz,]N,m), y1Arr None) -- the discrete-time shapes of the reference, batch axis first when any input has one.
        """
        ctrl, dyn = self.blocks
        N = int(np.ceil(self.t_span[1] / self.dt))
        n = self.model.n
        x0 = dyn.x0
        device = pick_device(x0, *ctrl.update.arrays())
        dtype = pick_dtype(x0, *ctrl.update.arrays())
        x0 = to_dev(x0, dtype, device)
        L, xT, uT = ctrl.update.as_policy(n, N, dtype, device)
        if L.shape[-3] < N:
            raise ValueError(f"controller holds {L.shape[-3]} gains, the simulation needs {N}")
        l = torch.zeros(L.shape[:-1], dtype=dtype, device=device)
        traj = trajectoryRollout(x0, dyn.update, AffinePolicy(l, L), Trajectory(xT, uT), 1)
        tArr = torch.arange(0, N + 1, dtype=dtype, device=device) * self.dt
        x0Arr = traj.xTraj.new_zeros(traj.xTraj.shape[:-1] + (0,))
        return tArr, x0Arr, traj.xTraj, traj.uTraj, None
