#!/usr/bin/env python
"""
Benchmark of the hot path -- contract in the task statement / DESIGN.md "Measurement".

Workload (BASELINE.json configs[1], the configuration the metric is quoted on that fits one GPU):
  batched quadcopter LQR-MPC solve, n=12 m=4 N=50, 65,536 problems per GPU, fp32.
One "step" = for every problem i: linearise the Euler quadcopter at xbar_i (CUDA kernel), then
`lqrMpc(A_i,B_i,Q_i,R_i,N,Qf=10 Q_i).solve(x0_i)` with infinite bounds = a full Riccati sweep (gains
recomputed every step, nothing cached) + the closed-loop plan rollout.  Metric: solves/s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]

N>1 is launched by torchrun, one rank per GPU; the batch is sharded (each rank owns its own 65,536
problems: weak scaling), no collective on the data path; timing = barrier + sync on both sides, max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_HORIZON, NX, NU = 50, 12, 4
FLOP_PER_SOLVE = 657067.0  # 50 x 13,141 dense as-written flop of the Riccati step (no symmetry discount)
METRIC = "quadcopter_lqr_mpc_solves_per_s"
UNIT = "solves/s"


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            f = [s.strip() for s in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_problem(Bsz, rank):
    from zopt_b200 import configs
    d = configs.cfg2(Bsz=Bsz, seed=1234 + 2 + 1000 * rank)
    return d


# ------------------------------------------------------------------------------------------------ CPU baseline
def cpu_step(xbar, ubar, qd, rd, N, dt, threads):
    """The oracle (port of the reference's algorithm) on host cores: autodiff linearisation (the reference uses
    jax.jacobian), batched Riccati recursion, linear closed-loop rollout.  torch-CPU fp64, all host threads."""
    from oracle.quadcopter import Quadcopter
    torch.set_num_threads(threads)
    ac = Quadcopter()
    f = ac.eulerStep(dt)
    xb, ub = torch.as_tensor(xbar), torch.as_tensor(ubar)
    A, B = torch.func.vmap(torch.func.jacrev(f, argnums=(0, 1)))(xb, ub)
    Q, R = torch.diag_embed(torch.as_tensor(qd)), torch.diag_embed(torch.as_tensor(rd))
    V = 10 * Q
    Ls = []
    Bt = B.transpose(1, 2)
    for k in range(N):  # lqrUtils.py:167-170
        BtV = Bt @ V
        L = torch.linalg.solve(R + BtV @ B, BtV @ A)
        Acl = A - B @ L
        V = Q + L.transpose(1, 2) @ R @ L + Acl.transpose(1, 2) @ V @ Acl
        Ls.append(L)
    Ls = Ls[::-1]
    x = xb
    xs, us = [x], []
    for k in range(N):
        u = -(Ls[k] @ x.unsqueeze(-1)).squeeze(-1)
        x = (A @ x.unsqueeze(-1)).squeeze(-1) + (B @ u.unsqueeze(-1)).squeeze(-1)
        xs.append(x)
        us.append(u)
    return torch.stack(xs, 1), torch.stack(us, 1)


def time_cpu(sample, steps, warmup, rank=0):
    d = make_problem(sample, rank)
    threads = os.cpu_count() or 1
    args = (d["xbar"], d["ubar"], d["qdiag"], d["rdiag"], d["N"], d["dt"], threads)
    for _ in range(warmup):
        cpu_step(*args)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(*args)
    el = time.perf_counter() - t0
    return sample * steps / el, el / steps, threads


def try_real_reference(sample, steps, warmup):
    """The reference's own JAX-CPU path, if `jax` and the reference package are importable (they are not in this image:
    DESIGN.md "Reference install").  jax.jit(jax.vmap(...)) over zopt.lqrUtils.discreteFiniteHorizonLqr + the plan rollout."""
    try:
        sys.path.insert(0, os.path.join(ROOT, "baseline", "_ref"))
        os.environ.setdefault("JAX_PLATFORMS", "cpu")
        import jax
        import jax.numpy as jnp
        from zopt.lqrUtils import discreteFiniteHorizonLqr
        from zopt.quadcopter import Quadcopter
    except Exception:
        return None
    d = make_problem(sample, 0)
    ac, N, dt = Quadcopter(), d["N"], d["dt"]

    def one(xbar, ubar, qd, rd):
        Aw, Bw = jax.jacobian(ac.inertialDynamics, argnums=(0, 1))(xbar, ubar)
        A, B = jnp.eye(12) + dt * Aw, dt * Bw
        Q, R = jnp.diag(qd), jnp.diag(rd)
        Qk = jnp.concatenate([jnp.repeat(Q[None], N, 0), 10 * Q[None]], 0)
        L = discreteFiniteHorizonLqr(jnp.repeat(A[None], N, 0), jnp.repeat(B[None], N, 0), Qk, jnp.repeat(R[None], N, 0), N)

        def stepf(x, Lk):
            u = -Lk @ x
            return A @ x + B @ u, (x, u)

        _, (xs, us) = jax.lax.scan(stepf, xbar, L)
        return xs, us

    f = jax.jit(jax.vmap(one))
    args_ = tuple(jnp.asarray(d[k]) for k in ("xbar", "ubar", "qdiag", "rdiag"))
    for _ in range(max(warmup, 1)):
        jax.block_until_ready(f(*args_))
    t0 = time.perf_counter()
    for _ in range(steps):
        jax.block_until_ready(f(*args_))
    el = time.perf_counter() - t0
    return sample * steps / el, el / steps, os.cpu_count() or 1


def run_reference(args):
    rank, _, world = dist_env()
    if rank != 0:
        return
    # bounded sample of the workload per step, scaled so that the whole run is ~20 full samples of CPU work whatever K is
    sample = max(256, min(args.cpu_sample, args.cpu_sample * 20 // max(args.steps, 1)))
    real = try_real_reference(sample, args.steps, min(args.warmup, 1))
    kind = "reference" if real else "port"
    val, sec, threads = real if real else time_cpu(sample, args.steps, min(args.warmup, 1))
    cb = {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
          "sample": f"{sample} problems of the same cfg-2 workload per step (torch-CPU fp64 oracle port; JAX is not installed)"}
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic", "config": workload_config(args.batch), "cpu_baseline": cb,
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(Bsz):
    return {"workload": "cfg2: batched quadcopter LQR-MPC solve (linearise + Riccati sweep N=50 + plan rollout), n=12 m=4, "
                        f"{Bsz} problems per GPU", "n": NX, "m": NU, "N": N_HORIZON, "batch_per_gpu": Bsz,
            "parallelism": "batch sharded over GPUs, no collective",
            "l2": "working set per step (gains 9.6 KB + plan 3.2 KB + A,B,Q,R 1.4 KB per problem) is ~0.9 GB >> 126 MB L2"}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    rank, local_rank, world = dist_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        # NCCL's version banner / INFO lines must not share stdout with the JSON line
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        dist.init_process_group("nccl", device_id=dev)
    from zopt_b200 import _lib
    from zopt_b200.mpcUtils import lqrMpc
    from zopt_b200.quadcopter import Quadcopter

    Bsz = args.batch
    d = make_problem(Bsz, rank)
    f32 = torch.float32
    # host (pinned) inputs of one step: the states to solve from
    x0_host = torch.as_tensor(d["xbar"], dtype=f32).pin_memory()
    xbar = torch.as_tensor(d["xbar"], dtype=f32, device=dev)
    ubar = torch.as_tensor(d["ubar"], dtype=f32, device=dev)
    Q = torch.diag_embed(torch.as_tensor(d["qdiag"], dtype=f32, device=dev))
    R = torch.diag_embed(torch.as_tensor(d["rdiag"], dtype=f32, device=dev))
    Qf = 10 * Q
    inf_n, inf_m = torch.full((NX,), float("inf"), device=dev), torch.full((NU,), float("inf"), device=dev)
    ninf_n, ninf_m = -inf_n, -inf_m
    ac = Quadcopter()
    N, dt = d["N"], d["dt"]

    def step(x0_dev):
        A, B = ac.linearizeInertial(xbar, ubar, dt)
        prob = lqrMpc(A, B, Q, R, N, ninf_n, inf_n, ninf_m, inf_m, Qf=Qf)
        return prob.solve(x0_dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        out = step(xbar)
    barrier()

    # --- kernel-resident timing: inputs already in HBM -----------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        out = step(xbar)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)

    # --- dominant kernel alone (the Riccati sweep + rollout launch), events on the launching stream
    A, B = ac.linearizeInertial(xbar, ubar, dt)
    prob = lqrMpc(A, B, Q, R, N, ninf_n, inf_n, ninf_m, inf_m, Qf=Qf)
    kms = []
    for _ in range(max(3, min(args.steps, 20))):
        k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        k0.record()
        prob.solve(xbar)
        k1.record()
        torch.cuda.synchronize(dev)
        kms.append(k0.elapsed_time(k1))
    k_ms = float(np.mean(kms))

    if os.environ.get("BENCH_SAMPLE_E2E", "0") != "1":
        clocks = sampler.stop() if rank == 0 else None  # sampled every 100 ms over the device-timed loops above

    # --- end to end through the public API with host buffers -----------------------------------
    # every step: x0 pinned host -> device, linearise + solve, then the call's whole return value (u, plan, status)
    # device -> pinned host.  The D2H of step i runs on a copy stream while step i+1 computes (two host buffer sets).
    def host_bufs():
        return [torch.empty((Bsz, NU), dtype=f32).pin_memory(), torch.empty((Bsz, N + 1, NX), dtype=f32).pin_memory(),
                torch.empty((Bsz, N, NU), dtype=f32).pin_memory(), torch.empty((Bsz,), dtype=torch.int8).pin_memory()]

    hb = [host_bufs(), host_bufs()]
    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)

    def e2e_loop(steps, full):
        for i in range(steps):
            x0d = x0_host.to(dev, non_blocking=True)
            u, traj, status = step(x0d)
            outs = (u, traj.xTraj, traj.uTraj, status) if full else (u, status)
            dst = hb[i % 2] if full else [hb[i % 2][0], hb[i % 2][3]]
            ready = torch.cuda.Event()
            ready.record(main)
            copy_stream.wait_event(ready)
            with torch.cuda.stream(copy_stream):
                for h, t in zip(dst, outs):
                    t.record_stream(copy_stream)
                    h.copy_(t, non_blocking=True)
        copy_stream.synchronize()

    e2e_loop(2, True)
    barrier()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    e2e_loop(args.steps, True)
    main.wait_stream(copy_stream)
    t1.record()
    barrier()
    e2e_ms = t0.elapsed_time(t1)
    h2d = x0_host.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in hb[0])
    t2, t3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t2.record()
    e2e_loop(args.steps, False)
    main.wait_stream(copy_stream)
    t3.record()
    barrier()
    e2e_u_ms = t2.elapsed_time(t3)
    if os.environ.get("BENCH_SAMPLE_E2E", "0") == "1":
        clocks = sampler.stop() if rank == 0 else None

    # --- secondary workloads of BASELINE.json (reported under "extra"; the headline stays cfg 2) -------------------
    # cfg 3: closed-loop LQR-MPC, N=50 horizon x 200 sim steps, 16,384 problems in total SHARDED over the ranks (strong)
    # cfg 4: iLQR, N=200, 10 forced iterations with the 16-way line search, 16,384 problems in total, fp64
    from zopt_b200 import configs, ilqrUtils
    from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost
    from zopt_b200.mpcUtils import quadcopterClosedLoopMpc
    from zopt_b200.sharding import shard_range
    extra_ms = [0.0]
    if not args.no_extras:
        lo, hi = shard_range(16384, rank, world)
        d3 = configs.cfg3(Bsz=16384)
        x3 = torch.as_tensor(d3["xbar"][lo:hi], dtype=f32, device=dev)
        x3[:, 9:12] *= 0.2
        Q3 = torch.diag_embed(torch.as_tensor(d3["qdiag"][lo:hi], dtype=f32, device=dev))
        R3 = torch.diag_embed(torch.as_tensor(d3["rdiag"][lo:hi], dtype=f32, device=dev))
        Qf3 = 10 * Q3
        quadcopterClosedLoopMpc(x3, Q3, R3, 50, 200, dt=0.1, Qf=Qf3)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            quadcopterClosedLoopMpc(x3, Q3, R3, 50, 200, dt=0.1, Qf=Qf3)
        c1.record()
        barrier()
        extra_ms[0] = c0.elapsed_time(c1) / 3
        # the same with 16,384 problems PER GPU (weak scaling): 2,048 problems leave most SMs of a B200 idle
        if world > 1:
            d3w = configs.cfg3(Bsz=16384, seed=1234 + 3 + 1000 * rank)
            x3w = torch.as_tensor(d3w["xbar"], dtype=f32, device=dev)
            x3w[:, 9:12] *= 0.2
            Q3w = torch.diag_embed(torch.as_tensor(d3w["qdiag"], dtype=f32, device=dev))
            R3w = torch.diag_embed(torch.as_tensor(d3w["rdiag"], dtype=f32, device=dev))
            Qf3w = 10 * Q3w
            quadcopterClosedLoopMpc(x3w, Q3w, R3w, 50, 200, dt=0.1, Qf=Qf3w)
            barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record()
            for _ in range(3):
                quadcopterClosedLoopMpc(x3w, Q3w, R3w, 50, 200, dt=0.1, Qf=Qf3w)
            c1.record()
            barrier()
            extra_ms.append(c0.elapsed_time(c1) / 3)
        else:
            extra_ms.append(extra_ms[0])
        # cfg 3 in fp64, the reference's own precision: fused cooperative kernel (csrc/lqr_quad64.cuh)
        x3d, Q3d, R3d = x3.double(), Q3.double(), R3.double()
        quadcopterClosedLoopMpc(x3d, Q3d, R3d, 50, 200, dt=0.1, Qf=10 * Q3d)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(2):
            quadcopterClosedLoopMpc(x3d, Q3d, R3d, 50, 200, dt=0.1, Qf=10 * Q3d)
        c1.record()
        barrier()
        cl64_ms = c0.elapsed_time(c1) / 2
        del x3d, Q3d, R3d
        d4 = configs.cfg4(Bsz=16384)
        x4 = torch.as_tensor(d4["x0"][lo:hi], dtype=torch.float64, device=dev)
        uG = torch.as_tensor(d4["uGuess"], dtype=torch.float64, device=dev)
        margs = (QuadcopterEuler(d4["dt"]), QuadraticCost(d4["Q"], d4["R"]), QuadraticTerminalCost(d4["Qf"]))
        ilqrUtils.iterativeLqr(*margs, x4, uG, maxIter=10, tol=-1.0)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(2):
            ilqrUtils.iterativeLqr(*margs, x4, uG, maxIter=10, tol=-1.0)
        c1.record()
        barrier()
        extra_ms.append(c0.elapsed_time(c1) / 2)
        # cfg 5: DDP (second-order dynamics terms), N=100, 10 forced iterations, 16,384 problems in total, fp64
        d5 = configs.cfg5(Bsz=16384, N=100)
        x5 = torch.as_tensor(d5["x0"][lo:hi], dtype=torch.float64, device=dev)
        uG5 = torch.as_tensor(d5["uGuess"], dtype=torch.float64, device=dev)
        m5 = (QuadcopterEuler(d5["dt"]), QuadraticCost(d5["Q"], d5["R"]), QuadraticTerminalCost(d5["Qf"]))
        ilqrUtils.differentialDynamicProgramming(*m5, x5, uG5, maxIter=10, tol=-1.0)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        ilqrUtils.differentialDynamicProgramming(*m5, x5, uG5, maxIter=10, tol=-1.0)
        c1.record()
        barrier()
        extra_ms.append(c0.elapsed_time(c1))
        # a1 with genuinely time-varying operands: A[k], B[k], Q[k], R[k] materialised per step and streamed from HBM
        from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
        Atv = A[:, None].expand(-1, N, -1, -1).contiguous()
        Btv = B[:, None].expand(-1, N, -1, -1).contiguous()
        Qtv = Q[:, None].expand(-1, N, -1, -1).contiguous()
        Rtv = R[:, None].expand(-1, N, -1, -1).contiguous()
        Ltv = discreteFiniteHorizonLqr(Atv, Btv, Qtv, Rtv, N)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            Ltv = discreteFiniteHorizonLqr(Atv, Btv, Qtv, Rtv, N)
        c1.record()
        barrier()
        extra_ms.append(c0.elapsed_time(c1) / 5)
        del Atv, Btv, Qtv, Rtv, Ltv
        # the headline workload in the reference's own precision (fp64): cooperative four-threads-per-problem Riccati kernel
        f64 = torch.float64
        xb64, ub64 = xbar.to(f64), ubar.to(f64)
        Q64, R64 = Q.to(f64), R.to(f64)
        Qf64 = 10 * Q64

        def step64():
            A64, B64 = ac.linearizeInertial(xb64, ub64, dt)
            return lqrMpc(A64, B64, Q64, R64, N, ninf_n, inf_n, ninf_m, inf_m, Qf=Qf64).solve(xb64)

        step64()
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(5):
            step64()
        c1.record()
        barrier()
        extra_ms.append(c0.elapsed_time(c1) / 5)
        del xb64, ub64, Q64, R64, Qf64
        # cfg 3 tier B: the reference demo's box-constrained lqrMpc (demos/lqrMpc.py:11-32: hover linearisation, N=25, bounds
        # |uvw|<=1, |pq|<=0.3, |r|<=0.1, |phi,theta|<=0.5, |u|<=3, OSQP eps 1e-2), one solve per initial state, bounds bind
        from zopt_b200.quadcopter import Quadcopter as _Q
        Ab, Bb = _Q().linearizeInertial(np.zeros(12), configs.U_TRIM, 0.1)
        x_ub = np.array([1, 1, 1, 0.3, 0.3, 0.1, 0.5, 0.5, np.inf, np.inf, np.inf, np.inf])
        u_ub = np.full(4, 3.0)
        xb = np.zeros((16384, 12))
        xb[:, 9:12] = np.random.default_rng(1234 + 3).uniform(-10, 10, (16384, 3))
        xb = torch.as_tensor(xb[lo:hi], dtype=f32, device=dev)
        pb = lqrMpc(Ab.to(f32), Bb.to(f32), torch.eye(12, dtype=f32, device=dev), torch.eye(4, dtype=f32, device=dev), 25,
                    -x_ub, x_ub, -u_ub, u_ub)
        pb.solve(xb, eps_abs=1e-2, eps_rel=1e-2)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for _ in range(3):
            _, _, stb = pb.solve(xb, eps_abs=1e-2, eps_rel=1e-2)
        c1.record()
        barrier()
        extra_ms.append(c0.elapsed_time(c1) / 3)
        box_iters = float(pb.iters.float().mean())
        box_opt = float((stb == 0).float().mean())
        # ... and the demo's receding-horizon loop around it (demos/lqrMpc.py:42-47): 200 steps, warm-started, one fused kernel
        pb.closedLoop(xb, 200, eps_abs=1e-2, eps_rel=1e-2)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        _, stc = pb.closedLoop(xb, 200, eps_abs=1e-2, eps_rel=1e-2)
        c1.record()
        barrier()
        extra_ms.append(c0.elapsed_time(c1))
        boxcl_iters = float(pb.iters.float().mean()) / 200
        boxcl_opt = float((stc == 0).float().mean())
        # same loop, termination checked every 5 ADMM iterations instead of OSQP's default 25 (a warm-started solve converges in ~5)
        pb.closedLoop(xb, 200, eps_abs=1e-2, eps_rel=1e-2, check_termination=5)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        _, stc5 = pb.closedLoop(xb, 200, eps_abs=1e-2, eps_rel=1e-2, check_termination=5)
        c1.record()
        barrier()
        extra_ms.append(c0.elapsed_time(c1))
        boxcl5_iters = float(pb.iters.float().mean()) / 200
        boxcl5_opt = float((stc5 == 0).float().mean())
    while len(extra_ms) < 9:
        extra_ms.append(0.0)
    extra_ms.append(0.0 if args.no_extras else cl64_ms)

    times = torch.tensor([ms, e2e_ms, k_ms, e2e_u_ms] + extra_ms, dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    ms, e2e_ms, k_ms, e2e_u_ms, cl_ms, clw_ms, il_ms, ddp_ms, tv_ms, f64_ms, box_ms, boxcl_ms, boxcl5_ms, cl64_ms = (float(v) for v in times.cpu())

    if rank == 0:
        total = Bsz * world
        value = total * args.steps / (ms * 1e-3)
        e2e_val = total * args.steps / (e2e_ms * 1e-3)
        import ctypes as C
        peak32, clk = C.c_double(0), C.c_double(0)
        _lib.check(_lib.lib.zb_peak_fma(0, local_rank, C.byref(peak32), C.byref(clk)))
        peak64 = C.c_double(0)
        _lib.check(_lib.lib.zb_peak_fma(1, local_rank, C.byref(peak64), C.byref(clk)))
        ach_tf = FLOP_PER_SOLVE * Bsz / (k_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = (1456 + 9600 * 2 + 3248) * Bsz  # in: A,B,Q,R,Qf,x0; gains written + re-read; plan written (fp32)
        traffic = None
        try:  # DRAM bytes of the dominant kernel from the committed ncu capture (same kernel, same batch)
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("batch") == Bsz and tj.get("N") == N:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
        except Exception:
            pass
        roof = {"bound": "fp32_fma", "achieved": ach_tf, "peak": peak32.value / 1e12, "unit": "TFLOP/s",
                "frac": ach_tf / (peak32.value / 1e12), "traffic": traffic, "algorithmic_bytes": alg_bytes,
                "peak_source": "zb_peak_fma dependent-FMA probe measured in this run (MEASURED_PEAKS.json has no FP32 CUDA-core figure)",
                "kernel": "lqrMpc.solve launch (Riccati sweep + rollout)", "kernel_ms": k_ms,
                "algorithmic_flop_per_solve": FLOP_PER_SOLVE,
                "hbm": {"achieved": alg_bytes / (k_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                        "frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                        "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}}
        cpu_val, cpu_sec, threads = time_cpu(args.cpu_sample, 3, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(Bsz), "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "returns": "u (Bsz,4), plan xTraj (Bsz,51,12) + uTraj (Bsz,50,4), status -- PCIe-bound",
                    "control_only": {"value": total * args.steps / (e2e_u_ms * 1e-3), "unit": UNIT,
                                     "d2h_bytes_per_step": Bsz * (NU * 4 + 1), "returns": "u and status only"}},
            "gpu_launches": 2 * args.steps, "roofline": roof,
            "extra": None if args.no_extras else {
                "cfg3_closed_loop_mpc": {"value": 16384 * 200 / (cl_ms * 1e-3), "unit": "MPC solves/s", "ms": cl_ms,
                                         "workload": "16,384 problems total (sharded over ranks), 200 sim steps, horizon 50, fp32, "
                                                     "re-linearised every step, bounds inactive, one fused kernel",
                                         "scaling": "strong"},
                "cfg3_closed_loop_mpc_weak": {"value": 16384 * world * 200 / (clw_ms * 1e-3), "unit": "MPC solves/s", "ms": clw_ms,
                                              "workload": "as cfg3 but 16,384 problems PER GPU", "scaling": "weak"},
                "cfg3_closed_loop_mpc_fp64": {"value": 16384 * 200 / (cl64_ms * 1e-3), "unit": "MPC solves/s", "ms": cl64_ms,
                                              "workload": "cfg3 (16,384 problems total, sharded over ranks) in fp64, the reference's own precision: "
                                                          "one fused cooperative kernel", "scaling": "strong",
                                              "roofline": {"bound": "fp64_fma", "achieved": 16384 * 200 * FLOP_PER_SOLVE / (cl64_ms * 1e-3) / 1e12,
                                                           "peak": peak64.value / 1e12, "unit": "TFLOP/s",
                                                           "frac": 16384 * 200 * FLOP_PER_SOLVE / (cl64_ms * 1e-3) / peak64.value}},
                "cfg4_ilqr": {"value": 16384 * 10 / (il_ms * 1e-3), "unit": "problem-iterations/s", "ms": il_ms,
                              "workload": "16,384 problems total (sharded over ranks), N=200, 10 iterations, 16-way line search, fp64",
                              "scaling": "strong", "algorithmic_flop_per_problem_iteration": 4.66e6,
                              "roofline": {"bound": "fp64_fma", "achieved": 16384 * 10 * 4.66e6 / (il_ms * 1e-3) / 1e12,
                                           "peak": peak64.value / 1e12, "unit": "TFLOP/s",
                                           "frac": 16384 * 10 * 4.66e6 / (il_ms * 1e-3) / peak64.value,
                                           "peak_source": "zb_peak_fma(f64) dependent-FMA probe measured in this run"}},
                "cfg5_ddp": {"value": 16384 * 10 / (ddp_ms * 1e-3), "unit": "problem-iterations/s", "ms": ddp_ms,
                             "workload": "16,384 problems total (sharded over ranks), N=100, 10 iterations, eigen-clamped second-order "
                                         "terms every step, fp64", "scaling": "strong",
                             "algorithmic_flop_per_problem_iteration": 7.33e6,
                             "roofline": {"bound": "fp64_fma", "achieved": 16384 * 10 * 7.33e6 / (ddp_ms * 1e-3) / 1e12,
                                          "peak": peak64.value / 1e12, "unit": "TFLOP/s",
                                          "frac": 16384 * 10 * 7.33e6 / (ddp_ms * 1e-3) / peak64.value}},
                "cfg2_fp64": {"value": Bsz * world / (f64_ms * 1e-3), "unit": "solves/s", "ms": f64_ms,
                              "workload": f"the headline step (linearise + lqrMpc.solve, N={N}) in fp64, the reference's own precision, "
                                          f"{Bsz} problems per GPU", "scaling": "weak",
                              "roofline": {"bound": "fp64_fma", "achieved": FLOP_PER_SOLVE * Bsz / (f64_ms * 1e-3) / 1e12,
                                           "peak": peak64.value / 1e12, "unit": "TFLOP/s",
                                           "frac": FLOP_PER_SOLVE * Bsz / (f64_ms * 1e-3) / peak64.value}},
                "lqr_time_varying": {"value": Bsz * world / (tv_ms * 1e-3), "unit": "solves/s", "ms": tv_ms,
                                     "workload": f"discreteFiniteHorizonLqr with A[k], B[k], Q[k], R[k] materialised per step "
                                                 f"(streamed from HBM), {Bsz} problems per GPU, N={N}, fp32", "scaling": "weak",
                                     "roofline": {"bound": "hbm", "achieved": Bsz * N * 1600 / (tv_ms * 1e-3) / 1e9,
                                                  "peak": hbm_peak, "unit": "GB/s",
                                                  "frac": Bsz * N * 1600 / (tv_ms * 1e-3) / 1e9 / hbm_peak,
                                                  "algorithmic_bytes_per_problem_step": 1600,
                                                  "note": "SURVEY 8d: A 576 + B 192 + Q 576 + R 64 in, gains 192 out (fp32); the kernel "
                                                          "fetches only the lower-triangle chunks of Q (384 B)"}},
                "cfg3_box_constrained_mpc": {"value": 16384 / (box_ms * 1e-3), "unit": "solves/s", "ms": box_ms,
                                             "admm_iterations_mean_rank0": box_iters, "optimal_fraction_rank0": box_opt,
                                             "workload": "16,384 initial states total (sharded over ranks), the reference demo's box-constrained "
                                                         "lqrMpc (hover linearisation, N=25, demo bounds, eps 1e-2 as demos/lqrMpc.py:32), fp32, "
                                                         "bounds bind (10 m offsets, |v| <= 1)", "scaling": "strong"},
                "cfg3_box_constrained_mpc_closed_loop": {"value": 16384 * 200 / (boxcl_ms * 1e-3), "unit": "MPC solves/s", "ms": boxcl_ms,
                                                         "admm_iterations_per_solve_rank0": boxcl_iters, "optimal_fraction_rank0": boxcl_opt,
                                                         "workload": "the same problem in the demo's receding-horizon loop (demos/lqrMpc.py:42-47): "
                                                                     "16,384 initial states total, 200 steps, clip + warm-started solve + perfect "
                                                                     "tracking, one fused kernel, fp32, eps 1e-2", "scaling": "strong"},
                "cfg3_box_constrained_mpc_closed_loop_check5": {"value": 16384 * 200 / (boxcl5_ms * 1e-3), "unit": "MPC solves/s", "ms": boxcl5_ms,
                                                                "admm_iterations_per_solve_rank0": boxcl5_iters, "optimal_fraction_rank0": boxcl5_opt,
                                                                "workload": "as above with check_termination=5 (OSQP default is 25)", "scaling": "strong"}},
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{args.cpu_sample} problems of the same workload per step, 3 steps of {cpu_sec:.2f} s after 1 warm-up, torch-CPU fp64 oracle port (JAX not installed)"},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=65536, help="problems per GPU")
    ap.add_argument("--cpu-sample", type=int, default=16384)
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary cfg3/cfg4 measurements")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
