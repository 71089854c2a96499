"""
LQR-MPC -- mirror of zopt/mpcUtils.py:12-81 (the plotting half, :84-202, is out of scope).

`lqrMpc(A, B, Q, R, N, x_lb, x_ub, u_lb, u_ub, Qf=None).solve(x0)` solves

    min  sum_{k<N} x_k'Q x_k + u_k'R u_k + x_N'Qf x_N
    s.t. x_{k+1} = A x_k + B u_k,  x_lb <= x_k <= x_ub (k = 0..N),  u_lb <= u_k <= u_ub,  x_0 = x0

exactly as the reference builds it with cvxpy (mpcUtils.py:47-59), for a whole batch of problems:
every constructor array and `x0` may carry one leading batch axis.  The reference hands the QP to a
third-party solver (cvxpy -> OSQP/Clarabel); here the solve runs in CUDA kernels.  When no bound is
finite the optimum is the Riccati sweep + linear rollout (exact); with finite bounds it is the ADMM
splitting OSQP uses, with the KKT solve done as a Riccati sweep.

Returns `(u, Trajectory(xTraj, uTraj), status)`; un-batched calls return cvxpy's status string,
batched calls an int8 tensor of codes (`STATUS[code]` gives the string).
"""
import ctypes as C
import weakref

import numpy as np
import torch

from ._lib import (View, ZbAdmmOpts, cached_matrix_flag, check, dcode, is_symmetric, lib, null_arr, pick_device, pick_dtype, ptr,
                   stream_ptr, to_dev)
from .pytrees import Trajectory

STATUS = ("optimal", "optimal_inaccurate", "infeasible")


_INF_CACHE = {}


def _isinf_all(t, sign):
    """every entry is +inf (sign > 0) / -inf; device tensors are cached by object identity (one sync per tensor)"""
    if not isinstance(t, torch.Tensor):
        t = torch.as_tensor(np.asarray(t))
    elif t.is_cuda:
        ent = _INF_CACHE.get((id(t), sign))
        if ent is not None and ent[0]() is t and ent[1] == t._version:
            return ent[2]
        if len(_INF_CACHE) > 256:
            _INF_CACHE.clear()
        res = bool(torch.all(torch.isinf(t) & ((t > 0) if sign > 0 else (t < 0))))
        _INF_CACHE[(id(t), sign)] = (weakref.ref(t), t._version, res)
        return res
    return bool(torch.all(torch.isinf(t) & ((t > 0) if sign > 0 else (t < 0))))


def _is_diagonal(t):
    """Exact diagonality of a batch of square matrices (cached per tensor object, see _lib.cached_matrix_flag)."""
    return cached_matrix_flag(t, "diag", lambda a: bool((a - torch.diag_embed(torch.diagonal(a, dim1=-2, dim2=-1))).abs().max() == 0))


class lqrMpc():

    def __init__(self, A, B, Q, R, N, x_lb, x_ub, u_lb, u_ub, Qf=None):
        """
        Setup an LQR MPC problem (zopt/mpcUtils.py:14-59).  Shapes as the reference -- A (n,n), B (n,m), Q (n,n),
        R (m,m), bounds (n,) / (m,), Qf defaults to Q -- each optionally with a leading batch axis.
        """
        if Qf is None:
            Qf = Q
        ops = [A, B, Q, R, Qf, x_lb, x_ub, u_lb, u_ub]
        self.device = pick_device(*ops)
        self.dtype = pick_dtype(A, B, Q, R, Qf)
        self.ops = [to_dev(t, self.dtype, self.device) for t in ops]
        self.N = int(N)
        self.n, self.m = self.ops[1].shape[-2], self.ops[1].shape[-1]
        core = [2, 2, 2, 2, 2, 1, 1, 1, 1]
        self.Bsz, self.batched = 1, False
        for t, nd in zip(self.ops, core):
            if t.ndim == nd + 1:
                self.batched = True
                if t.shape[0] != 1:
                    if self.Bsz not in (1, t.shape[0]):
                        raise ValueError("inconsistent batch sizes")
                    self.Bsz = t.shape[0]
            elif t.ndim != nd:
                raise ValueError(f"expected {nd} or {nd + 1} dimensions, got shape {tuple(t.shape)}")
        self.views = [View(t, nd, False, t.ndim == nd + 1) for t, nd in zip(self.ops, core)]
        self.bounded = not (_isinf_all(x_lb, -1) and _isinf_all(x_ub, +1) and _isinf_all(u_lb, -1) and
                            _isinf_all(u_ub, +1))
        # diagonal costs: lets the fp32 (12,4) kernel keep less cost data on chip.  The check needs a device->host
        # read, so it is cached per tensor (storage, shape, in-place version): re-building the problem every MPC step
        # around the same Q, R, Qf (demos/lqrMpc.py:31 builds once; a re-linearising loop rebuilds) costs nothing.
        self.cost_diagonal = all(_is_diagonal(t) for t in (self.ops[2], self.ops[3], self.ops[4]))
        # the (12,4) Riccati kernels read the lower triangle of the (symmetric) weights; non-symmetric weights are legal input
        # for the reference's cvxpy quad_form only up to symmetrisation, so they take the generic kernels, weights as given
        self.cost_symmetric = self.cost_diagonal or all(is_symmetric(t) for t in (self.ops[2], self.ops[3], self.ops[4]))
        shapes = [(self.n, self.n), (self.n, self.m), (self.n, self.n), (self.m, self.m), (self.n, self.n), (self.n,), (self.n,), (self.m,), (self.m,)]
        for name, t, nd, shp in zip(("A", "B", "Q", "R", "Qf", "x_lb", "x_ub", "u_lb", "u_ub"), self.ops, core, shapes):
            if tuple(t.shape[-nd:]) != shp:
                raise ValueError(f"{name}: expected trailing shape {shp}, got {tuple(t.shape)}")
        self.iters = None
        # Finite bounds, (n,m) = (12,4), one problem definition for every x0 (the reference's usage: the QP is built once,
        # mpcUtils.py:14-59, and re-solved per x0): the definition is kept on the host and handed to the kernel by value,
        # the rho-grid gain tables are built at the first solve and reused (zb_mpc_box_*).
        self._box = None
        if self.bounded and (self.n, self.m) == (12, 4) and all(t.ndim == nd or t.shape[0] == 1 for t, nd in zip(self.ops, core)):
            host = [np.ascontiguousarray(t.detach().to("cpu", torch.float64).numpy().reshape(t.shape[-nd:]))
                    for t, nd in zip(self.ops, core)]
            self._box = {"host": host, "tables": None, "rho0": None}

    def solve(self, x0, **kwargs):
        """
        Solve the MPC step at state x0 (zopt/mpcUtils.py:61-81).  Keyword arguments follow OSQP's names as passed through
        cvxpy in the reference demo; honoured: `eps_abs`, `eps_rel`, `max_iter`, `rho` (initial, then residual-balanced),
        `alpha` (over-relaxation), `check_termination`, `eps_prim_inf`.  Others are accepted and ignored -- in particular
        `sigma`: OSQP regularises its KKT system with it, the Riccati-structured solve here needs no regularisation.
        `kernel="split"` (bounds inactive, fp32 (12,4), diagonal costs, multi-wave batches) runs the sweep and the plan rollout as two
        concurrent kernels instead of the fused one (an experiment that measured equal; same bits).
        `kernel="generic"` forces the generic one-thread-per-problem ADMM kernel where the shared-definition (12,4) kernels would
        run; `"thread"` / `"quad"` pick the shared-definition kernel with one / four threads per problem (default: by batch size).

        Returns
        -------
            u : optimal control at the current time step, (m,) or (Bsz,m)
            traj : Trajectory(xTraj (N+1,n), uTraj (N,m)) of the plan
            status : "optimal" / "optimal_inaccurate" / "infeasible" (int8 codes when batched)
        """
        x0 = to_dev(x0, self.dtype, self.device)
        if x0.ndim not in (1, 2) or x0.shape[-1] != self.n:
            raise ValueError(f"x0 must be ({self.n},) or (Bsz,{self.n}), got {tuple(x0.shape)}")
        batched = self.batched or x0.ndim == 2
        Bsz = self.Bsz
        if x0.ndim == 2 and x0.shape[0] != 1:
            if Bsz not in (1, x0.shape[0]):
                raise ValueError("inconsistent batch sizes")
            Bsz = x0.shape[0]
        x0 = (x0 if x0.ndim == 2 else x0[None]).expand(Bsz, self.n).contiguous()
        N, n, m, dt, dev = self.N, self.n, self.m, self.dtype, self.device
        u0 = torch.empty((Bsz, m), dtype=dt, device=dev)
        xTraj = torch.empty((Bsz, N + 1, n), dtype=dt, device=dev)
        uTraj = torch.empty((Bsz, N, m), dtype=dt, device=dev)
        status = torch.empty((Bsz,), dtype=torch.int8, device=dev)
        iters = torch.empty((Bsz,), dtype=torch.int32, device=dev)
        opts = ZbAdmmOpts(int(kwargs.get("max_iter", 4000)), int(kwargs.get("check_termination", 25)),
                          float(kwargs.get("rho", 0.1)), float(kwargs.get("sigma", 1e-6)), float(kwargs.get("alpha", 1.6)),
                          float(kwargs.get("eps_abs", 1e-3)), float(kwargs.get("eps_rel", 1e-3)),
                          float(kwargs.get("eps_prim_inf", 1e-4)))
        if self._box is not None and kwargs.get("kernel", "auto") != "generic":
            self._solve_box(x0, Bsz, opts, u0, xTraj, uTraj, status, iters, {"auto": 0, "thread": 4, "quad": 8, "quad_global": 24}[kwargs.get("kernel", "auto")])
            self.iters = iters
            if not batched:
                return u0[0], Trajectory(xTraj[0], uTraj[0]), STATUS[int(status[0])]
            return u0, Trajectory(xTraj, uTraj), status
        wsb = lib.zb_mpc_workspace_bytes(dcode(dt), Bsz, N, n, m)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        check(lib.zb_mpc_lqr_solve(dcode(dt), dev.index, stream_ptr(dev), Bsz, N, n, m, *[v.ref() for v in self.views],
                                   (1 if self.bounded else 0) | (2 if self.cost_diagonal else 0) | (0 if self.cost_symmetric else 128) |
                                   (512 if kwargs.get("kernel") == "split" else 0), ptr(x0), C.byref(opts), ptr(u0), ptr(xTraj), ptr(uTraj), ptr(status),
                                   ptr(iters), ptr(ws), wsb))
        self.iters = iters
        if not batched:
            return u0[0], Trajectory(xTraj[0], uTraj[0]), STATUS[int(status[0])]
        return u0, Trajectory(xTraj, uTraj), status

    def _box_tables(self, opts):
        """rho-grid gain tables of the shared problem definition, built once per (object, rho0)"""
        dt, dev, N, box = self.dtype, self.device, self.N, self._box
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        A, B, Q, R, Qf = box["host"][:5]
        if box["tables"] is None or box["rho0"] != opts.rho:
            tb = lib.zb_mpc_box_tables_bytes(dcode(dt), N)
            tables = torch.empty((tb,), dtype=torch.uint8, device=dev)
            check(lib.zb_mpc_box_build_tables(dcode(dt), dev.index, stream_ptr(dev), N, dp(A), dp(B), dp(Q), dp(R), dp(Qf),
                                              opts.rho, ptr(tables), tb))
            box["tables"], box["rho0"] = tables, opts.rho
        return box["tables"]

    def _solve_box(self, x0, Bsz, opts, u0, xTraj, uTraj, status, iters, variant=0):
        """bounded (12,4) problem shared by the batch -> zb_mpc_box_solve (csrc/mpc_box.cuh)"""
        dt, dev, N, box = self.dtype, self.device, self.N, self._box
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        A, B, Q, R, Qf, xlb, xub, ulb, uub = box["host"]
        self._box_tables(opts)
        wsb = lib.zb_mpc_box_workspace_bytes(dcode(dt), Bsz, N)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        check(lib.zb_mpc_box_solve(dcode(dt), dev.index, stream_ptr(dev), Bsz, N, dp(A), dp(B), dp(xlb), dp(xub), dp(ulb), dp(uub),
                                   ptr(box["tables"]), ptr(x0), C.byref(opts), variant, ptr(u0), ptr(xTraj), ptr(uTraj), ptr(status),
                                   ptr(iters), ptr(ws), wsb))

    def closedLoop(self, x0, Tsim, clip=1e-6, **kwargs):
        """
        The receding-horizon loop of demos/lqrMpc.py:42-47 for a batch of initial states, fused into one kernel
        (zb_mpc_box_closed_loop): per simulation step `x = clip(x, x_lb + clip, x_ub - clip)`, `solve(x)` warm-started from
        the previous step, `x = traj.xTraj[1]` ("assume perfect tracking").  Needs the shared-definition (12,4) bounded
        problem (one A, B, Q, R, bounds for the batch); otherwise loop over `solve` as the demo does.

        Returns (Trajectory(xTraj (Bsz,Tsim+1,12), uTraj (Bsz,Tsim,4)), status (Bsz,) int8 worst over the steps); total ADMM
        iterations per problem in `self.iters`.
        """
        if self._box is None:
            raise NotImplementedError("closedLoop needs finite bounds, (n,m) = (12,4) and one problem definition for the batch; "
                                      "compose solve() step by step otherwise")
        dt, dev, N, box = self.dtype, self.device, self.N, self._box
        x0 = to_dev(x0, dt, dev)
        if x0.ndim not in (1, 2) or x0.shape[-1] != 12:
            raise ValueError(f"x0 must be (12,) or (Bsz,12), got {tuple(x0.shape)}")
        batched = x0.ndim == 2
        x0 = (x0 if batched else x0[None]).contiguous()
        Bsz, Tsim = x0.shape[0], int(Tsim)
        opts = ZbAdmmOpts(int(kwargs.get("max_iter", 4000)), int(kwargs.get("check_termination", 25)),
                          float(kwargs.get("rho", 0.1)), float(kwargs.get("sigma", 1e-6)), float(kwargs.get("alpha", 1.6)),
                          float(kwargs.get("eps_abs", 1e-3)), float(kwargs.get("eps_rel", 1e-3)),
                          float(kwargs.get("eps_prim_inf", 1e-4)))
        tables = self._box_tables(opts)
        dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
        A, B, Q, R, Qf, xlb, xub, ulb, uub = box["host"]
        xS = torch.empty((Bsz, Tsim + 1, 12), dtype=dt, device=dev)
        uS = torch.empty((Bsz, Tsim, 4), dtype=dt, device=dev)
        status = torch.empty((Bsz,), dtype=torch.int8, device=dev)
        iters = torch.empty((Bsz,), dtype=torch.int32, device=dev)
        wsb = lib.zb_mpc_box_closed_loop_workspace_bytes(dcode(dt), Bsz, N)
        ws = torch.empty((wsb,), dtype=torch.uint8, device=dev)
        check(lib.zb_mpc_box_closed_loop(dcode(dt), dev.index, stream_ptr(dev), Bsz, N, Tsim, dp(A), dp(B), dp(xlb), dp(xub),
                                         dp(ulb), dp(uub), ptr(tables), ptr(x0), C.byref(opts),
                                         {"auto": 0, "thread": 4, "quad": 8, "quad_global": 24}[kwargs.get("kernel", "auto")], float(clip), ptr(xS), ptr(uS),
                                         ptr(status), ptr(iters), ptr(ws), wsb))
        self.iters = iters
        if not batched:
            return Trajectory(xS[0], uS[0]), STATUS[int(status[0])]
        return Trajectory(xS, uS), status

    @staticmethod
    def status_str(status, i=0):
        return STATUS[int(status[i])]


def quadcopterClosedLoopMpc(x0, Q, R, N, Tsim, dt=0.1, Qf=None, uTrim=(9.807, 0.0, 0.0, 0.0), variant="auto"):
    """
    Receding-horizon LQR-MPC of the quadcopter in closed loop with the nonlinear plant (BASELINE cfg 3) -- the loop of
    demos/lqrMpc.py:42-47, batched and fused into one kernel: every simulation step re-linearises the Euler quadcopter at
    the current state (`A = I + dt dF/dx(x_t, uTrim)`, `B = dt dF/du`, zopt/quadcopter.py:116-144), solves the
    unconstrained `lqrMpc(A, B, Q, R, N, Qf=Qf)` problem from `x_t` (a full Riccati sweep, nothing cached), applies the
    first move (`uTrim + u_t`) to the plant `x + dt * inertialDynamics(x, u)`.  Bounds inactive; fp32, or fp64 when `x0` is an
    fp64 tensor.

    x0 (Bsz,12) or (12,); Q (12,12), R (4,4), Qf (12,12, default Q) optionally batched.
    Returns Trajectory(xTraj (Bsz,Tsim+1,12), uTraj (Bsz,Tsim,4)) with uTraj the deviation from uTrim.
    `variant`: "auto" (by batch size: nine lanes per problem for the smallest batches, four threads per problem for small
    ones, one thread per problem otherwise), "thread", "quad", "warp" (fp32 only).
    With finite bounds, compose `Quadcopter.linearizeInertial`, `lqrMpc(...).solve` and
    `Quadcopter.inertialDynamics` step by step instead.
    """
    if Qf is None:
        Qf = Q
    device = pick_device(x0, Q, R, Qf)
    # dtype as everywhere else (pick_dtype): fp32 only when every floating input is fp32; NumPy / Python numbers are fp64, the
    # reference's precision (cooperative fp64 kernel of csrc/lqr_quad64.cuh)
    f32 = pick_dtype(x0, Q, R, Qf)
    x0, Q, R, Qf = (to_dev(t, f32, device) for t in (x0, Q, R, Qf))
    if x0.ndim not in (1, 2) or x0.shape[-1] != 12:
        raise ValueError(f"x0 must be (12,) or (Bsz,12), got {tuple(x0.shape)}")
    batched = x0.ndim == 2
    x0 = (x0 if batched else x0[None]).contiguous()
    Bsz = x0.shape[0]
    for name, t, blk in (("Q", Q, (12, 12)), ("R", R, (4, 4)), ("Qf", Qf, (12, 12))):
        if tuple(t.shape[-2:]) != blk or t.ndim not in (2, 3) or (t.ndim == 3 and t.shape[0] not in (1, Bsz)):
            raise ValueError(f"{name} must be {blk} or (Bsz,) + {blk} with Bsz = {Bsz}, got {tuple(t.shape)}")
    views = [View(t, 2, False, t.ndim == 3) for t in (Q, R, Qf)]
    diag = all(_is_diagonal(t) for t in (Q, R, Qf))
    if not diag and not all(is_symmetric(t) for t in (Q, R, Qf)):
        raise ValueError("Q, R, Qf must be symmetric (the fused kernels read their lower triangle; cvxpy's quad_form in the "
                         "reference, zopt/mpcUtils.py:52-54, rejects non-symmetric weights as well)")
    xS = torch.empty((Bsz, Tsim + 1, 12), dtype=f32, device=device)
    uS = torch.empty((Bsz, Tsim, 4), dtype=f32, device=device)
    ut = (C.c_double * 4)(*[float(v) for v in uTrim])
    check(lib.zb_mpc_closed_loop_quad(0 if f32 == torch.float32 else 1, device.index, stream_ptr(device), Bsz, int(N), int(Tsim), float(dt), ut,
                                      *[v.ref() for v in views],
                                      (2 if diag else 0) | {"auto": 0, "thread": 4, "quad": 8, "warp": 64}[variant], ptr(x0), ptr(xS), ptr(uS)))
    return Trajectory(xS, uS) if batched else Trajectory(xS[0], uS[0])
