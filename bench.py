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


def load_configs():
    """zopt_b200/configs.py (NumPy-only problem generators) loaded BY PATH: importing the package would map
    libzopt_b200.so into the process, and the reference arm must not touch the product's native code."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("_zb_configs", os.path.join(ROOT, "zopt_b200", "configs.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def make_problem(Bsz, rank):
    return load_configs().cfg2(Bsz=Bsz, seed=1234 + 2 + 1000 * rank)


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


def c_oracle_or_none():
    """oracle/c/zopt_oracle.c through ctypes (compiled for this host on first use), or None when it cannot be built here"""
    try:
        from oracle import c_oracle
        c_oracle.load()
        return c_oracle
    except Exception as e:  # no gcc / no OpenMP on this box: the torch port below still runs
        print(f"bench: C oracle unavailable ({type(e).__name__}: {str(e)[:200]}); falling back to the torch-CPU port", file=sys.stderr)
        return None


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def time_cpu(sample, steps, warmup, rank=0):
    """-> (solves/s, seconds per step, threads, description).  The compiled restatement (plain C + OpenMP over problems,
    fp64) when it can be built on this host, else the torch-CPU port; both are checked against the reference's goldens
    (tests/test_c_oracle.py, tests/test_golden.py)."""
    d = make_problem(sample, rank)
    threads = host_threads()
    co = c_oracle_or_none()
    if co is not None:
        step = lambda: co.lqrMpcSolveBatch(d["xbar"], d["ubar"], d["qdiag"], d["rdiag"], d["N"], d["dt"], 10.0, threads)
        how = "plain-C + OpenMP restatement of the reference's algorithm (oracle/c/zopt_oracle.c: forward-mode linearisation, Riccati step as written, plan rollout), fp64"
    else:
        args = (d["xbar"], d["ubar"], d["qdiag"], d["rdiag"], d["N"], d["dt"], threads)
        step = lambda: cpu_step(*args)
        how = "torch-CPU fp64 oracle port (a Python loop of batched small products)"
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    el = time.perf_counter() - t0
    return sample * steps / el, el / steps, threads, how


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
    # bounded sample of the workload per step: sized from a probe so that the K steps take about a minute of CPU time, and
    # never more than the full step (args.batch problems)
    co = c_oracle_or_none()
    probe = 2048 if co is not None else 512
    rate = time_cpu(probe, 1, 1)[0]
    sample = int(max(256, min(args.batch, rate * 60.0 / max(args.steps, 1))))
    real = try_real_reference(sample, args.steps, min(args.warmup, 1))
    kind = "reference" if real else "port"
    if real:
        val, sec, threads = real
        how = "the reference's own jax.jit(jax.vmap(...)) on XLA:CPU"
    else:
        val, sec, threads, how = time_cpu(sample, args.steps, min(args.warmup, 1))
    cb = {"value": val, "unit": UNIT, "cores": threads, "kind": kind,
          "sample": f"{sample} problems of the same cfg-2 workload per step, {args.steps} steps; {how}; JAX is not installed"}
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
    from zopt_b200 import _lib, hostbind
    from zopt_b200.mpcUtils import lqrMpc
    from zopt_b200.quadcopter import Quadcopter

    # pin this rank to the CPUs of its GPU's NUMA node BEFORE any pinned allocation (the end-to-end loop is PCIe/host bound)
    binding = hostbind.bind_to_gpu(local_rank, enable=os.environ.get("BENCH_NO_BIND", "0") != "1")

    Bsz = args.batch
    d = make_problem(Bsz, rank)
    f32, f64 = torch.float32, torch.float64
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

    def make_step(xb, ub, Qm, Rm, Qfm):
        def step(x0_dev):
            A, B = ac.linearizeInertial(xb, ub, dt)
            prob = lqrMpc(A, B, Qm, Rm, N, ninf_n, inf_n, ninf_m, inf_m, Qf=Qfm)
            return prob.solve(x0_dev)
        return step

    step = make_step(xbar, ubar, Q, R, Qf)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, reps):
        """barrier + sync, `reps` back-to-back calls between ONE event pair on the launching stream, barrier + sync"""
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        barrier()
        return e0.elapsed_time(e1) / reps

    for _ in range(max(args.warmup, 3)):
        out = step(xbar)
    barrier()

    # --- kernel-resident timing: inputs already in HBM -----------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms = timed(lambda: step(xbar), args.steps) * args.steps

    # --- dominant kernel alone (the Riccati sweep + rollout launch): back-to-back launches between one event pair,
    #     no synchronisation in between, so launch latency is hidden behind the previous launch
    A, B = ac.linearizeInertial(xbar, ubar, dt)
    prob = lqrMpc(A, B, Q, R, N, ninf_n, inf_n, ninf_m, inf_m, Qf=Qf)
    prob.solve(xbar)
    k_reps = max(10, min(args.steps, 100))
    k_ms = timed(lambda: prob.solve(xbar), k_reps)
    # the same launch WITHOUT the diagonal-cost variant (dense symmetric Q, R, Qf: 27 instead of 4 float4 of cost data on chip)
    off = 1e-3 * torch.ones((NX, NX), device=dev, dtype=f32)
    Qd = Q + off * (1 - torch.eye(NX, device=dev))
    Rd = R + 1e-3 * (1 - torch.eye(NU, device=dev))
    prob_dense = lqrMpc(A, B, Qd, Rd, N, ninf_n, inf_n, ninf_m, inf_m, Qf=10 * Qd)
    prob_dense.solve(xbar)
    kd_ms = timed(lambda: prob_dense.solve(xbar), max(5, k_reps // 2))
    del prob_dense, Qd, Rd

    if os.environ.get("BENCH_SAMPLE_E2E", "0") != "1":
        clocks = sampler.stop() if rank == 0 else None  # sampled every 100 ms over the device-timed loops above

    # --- end to end through the public API with host buffers -----------------------------------
    # every step: x0 pinned host -> device, linearise + solve, then the call's whole return value (u, plan, status)
    # device -> pinned host.  The D2H of step i runs on a copy stream while step i+1 computes (two host buffer sets).
    def host_bufs(dtype):
        return [torch.empty((Bsz, NU), dtype=dtype).pin_memory(), torch.empty((Bsz, N + 1, NX), dtype=dtype).pin_memory(),
                torch.empty((Bsz, N, NU), dtype=dtype).pin_memory(), torch.empty((Bsz,), dtype=torch.int8).pin_memory()]

    copy_stream = torch.cuda.Stream(dev)
    main = torch.cuda.current_stream(dev)

    def e2e_loop(steps, full, stepf, x0h, hb):
        for i in range(steps):
            x0d = x0h.to(dev, non_blocking=True)
            u, traj, status = stepf(x0d)
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

    def e2e_time(steps, full, stepf, x0h, hb):
        e2e_loop(2, full, stepf, x0h, hb)
        barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        e2e_loop(steps, full, stepf, x0h, hb)
        main.wait_stream(copy_stream)
        t1.record()
        barrier()
        return t0.elapsed_time(t1)

    hb = [host_bufs(f32), host_bufs(f32)]
    e2e_ms = e2e_time(args.steps, True, step, x0_host, hb)
    h2d = x0_host.numel() * 4
    d2h = sum(t.numel() * t.element_size() for t in hb[0])
    e2e_u_ms = e2e_time(args.steps, False, step, x0_host, hb)

    # --- copy roof: the step's D2H payload alone, pinned, every rank at the same time (what PCIe + host memory allow) ---
    def copy_roof(hbufs, reps):
        srcs = [torch.empty(h.shape, dtype=h.dtype, device=dev) for h in hbufs[0]]
        for h, s in zip(hbufs[0], srcs):
            h.copy_(s, non_blocking=True)
        barrier()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        for i in range(reps):
            for h, s in zip(hbufs[i % 2], srcs):
                h.copy_(s, non_blocking=True)
        c1.record()
        barrier()
        return c0.elapsed_time(c1) / reps

    roof_ms = copy_roof(hb, max(5, min(args.steps, 20)))
    del hb

    # --- the same end-to-end loop in fp64, the reference's own precision (cooperative fp64 kernels) ------------------
    e2e64_ms, d2h64, h2d64 = 0.0, 0, 0
    if not args.no_extras:
        xb64, ub64, Q64, R64 = xbar.to(f64), ubar.to(f64), Q.to(f64), R.to(f64)
        step64 = make_step(xb64, ub64, Q64, R64, 10 * Q64)
        x0_host64 = torch.as_tensor(d["xbar"], dtype=f64).pin_memory()
        hb64 = [host_bufs(f64), host_bufs(f64)]
        n64 = max(3, min(args.steps, 20))
        e2e64_ms = e2e_time(n64, True, step64, x0_host64, hb64) / n64
        d2h64 = sum(t.numel() * t.element_size() for t in hb64[0])
        h2d64 = x0_host64.numel() * 8
        del hb64, x0_host64
    if os.environ.get("BENCH_SAMPLE_E2E", "0") == "1":
        clocks = sampler.stop() if rank == 0 else None

    # --- final gather of the step's costs-sized result over NCCL (the only collective, off the hot path) --
    gather_ms = 0.0
    if world > 1:
        from zopt_b200.sharding import gather
        u_last, traj_last, _ = out
        gather(u_last, Bsz * world)
        gather_ms = timed(lambda: (gather(u_last, Bsz * world), gather(traj_last.xTraj, Bsz * world), gather(traj_last.uTraj, Bsz * world)), 3)

    # --- secondary workloads of BASELINE.json (reported under "extra"; the headline stays cfg 2) -------------------
    # cfg 3: closed-loop LQR-MPC, N=50 horizon x 200 sim steps, 16,384 problems in total SHARDED over the ranks (strong)
    # cfg 4: iLQR, N=200, 10 forced iterations with the 16-way line search, 16,384 problems in total, fp64
    from zopt_b200 import configs, ilqrUtils
    from zopt_b200.models import QuadcopterEuler, QuadraticCost, QuadraticTerminalCost
    from zopt_b200.mpcUtils import quadcopterClosedLoopMpc
    from zopt_b200.sharding import shard_range
    X = {}  # name -> ms (max over ranks below)
    info = {}
    if not args.no_extras:
        lo, hi = shard_range(16384, rank, world)
        d3 = configs.cfg3(Bsz=16384)
        x3 = torch.as_tensor(d3["xbar"][lo:hi], dtype=f32, device=dev)
        x3[:, 9:12] *= 0.2
        Q3 = torch.diag_embed(torch.as_tensor(d3["qdiag"][lo:hi], dtype=f32, device=dev))
        R3 = torch.diag_embed(torch.as_tensor(d3["rdiag"][lo:hi], dtype=f32, device=dev))
        Qf3 = 10 * Q3
        quadcopterClosedLoopMpc(x3, Q3, R3, 50, 200, dt=0.1, Qf=Qf3)
        X["cl"] = timed(lambda: quadcopterClosedLoopMpc(x3, Q3, R3, 50, 200, dt=0.1, Qf=Qf3), 3)
        # the same with 16,384 problems PER GPU (weak scaling)
        if world > 1:
            d3w = configs.cfg3(Bsz=16384, seed=1234 + 3 + 1000 * rank)
            x3w = torch.as_tensor(d3w["xbar"], dtype=f32, device=dev)
            x3w[:, 9:12] *= 0.2
            Q3w = torch.diag_embed(torch.as_tensor(d3w["qdiag"], dtype=f32, device=dev))
            R3w = torch.diag_embed(torch.as_tensor(d3w["rdiag"], dtype=f32, device=dev))
            Qf3w = 10 * Q3w
            quadcopterClosedLoopMpc(x3w, Q3w, R3w, 50, 200, dt=0.1, Qf=Qf3w)
            X["clw"] = timed(lambda: quadcopterClosedLoopMpc(x3w, Q3w, R3w, 50, 200, dt=0.1, Qf=Qf3w), 3)
            del x3w, Q3w, R3w, Qf3w
        else:
            X["clw"] = X["cl"]
        # what ONE GPU does when cfg 3 is sharded over eight (2,048 problems): nine-lane kernel with work rotation (csrc/mpc_warp.cuh)
        if world == 1:
            xs_, Qs_, Rs_ = x3[:2048].contiguous(), Q3[:2048].contiguous(), R3[:2048].contiguous()
            Qfs_ = 10 * Qs_
            quadcopterClosedLoopMpc(xs_, Qs_, Rs_, 50, 200, dt=0.1, Qf=Qfs_)
            X["cl_shard"] = timed(lambda: quadcopterClosedLoopMpc(xs_, Qs_, Rs_, 50, 200, dt=0.1, Qf=Qfs_), 3)
            del xs_, Qs_, Rs_, Qfs_
        else:
            X["cl_shard"] = X["cl"]
        # cfg 3 in fp64, the reference's own precision: fused cooperative kernel (csrc/lqr_quad64.cuh)
        x3d, Q3d, R3d = x3.double(), Q3.double(), R3.double()
        quadcopterClosedLoopMpc(x3d, Q3d, R3d, 50, 200, dt=0.1, Qf=10 * Q3d)
        X["cl64"] = timed(lambda: quadcopterClosedLoopMpc(x3d, Q3d, R3d, 50, 200, dt=0.1, Qf=10 * Q3d), 2)
        del x3d, Q3d, R3d
        d4 = configs.cfg4(Bsz=16384)
        x4 = torch.as_tensor(d4["x0"][lo:hi], dtype=f64, device=dev)
        uG = torch.as_tensor(d4["uGuess"], dtype=f64, device=dev)
        margs = (QuadcopterEuler(d4["dt"]), QuadraticCost(d4["Q"], d4["R"]), QuadraticTerminalCost(d4["Qf"]))
        ilqrUtils.iterativeLqr(*margs, x4, uG, maxIter=10, tol=-1.0)
        X["il"] = timed(lambda: ilqrUtils.iterativeLqr(*margs, x4, uG, maxIter=10, tol=-1.0), 2)
        # the reference's defaults (maxIter=100, tol=1e-3, ilqrUtils.py:267-268): cost must follow the iterations actually taken
        _, _, _, conv_d, log_d = ilqrUtils.iterativeLqr(*margs, x4, uG, return_log=True)
        X["il_default"] = timed(lambda: ilqrUtils.iterativeLqr(*margs, x4, uG), 1)
        info["il_default_iters_mean"] = float(log_d["iters"].float().mean())
        info["il_default_iters_max"] = int(log_d["iters"].max())
        info["il_default_converged"] = float(conv_d.float().mean())
        # fp32 vs fp64 step-size sequences on a 1,024-problem subset (mismatches counted and reported)
        xs = x4[:min(1024, x4.shape[0])]
        _, _, _, _, lg64 = ilqrUtils.iterativeLqr(*margs, xs, uG, maxIter=10, tol=-1.0, return_log=True)
        _, _, _, _, lg32 = ilqrUtils.iterativeLqr(*margs, xs.float(), uG.float(), maxIter=10, tol=-1.0, return_log=True)
        info["ilqr_fp32_alpha_mismatch"] = int((lg64["alpha_idx"] != lg32["alpha_idx"]).any(dim=1).sum())
        info["ilqr_fp32_alpha_subset"] = int(xs.shape[0])
        # cfg 5: DDP (second-order dynamics terms), N=100, 10 forced iterations, 16,384 problems in total, fp64
        d5 = configs.cfg5(Bsz=16384, N=100)
        x5 = torch.as_tensor(d5["x0"][lo:hi], dtype=f64, device=dev)
        uG5 = torch.as_tensor(d5["uGuess"], dtype=f64, device=dev)
        m5 = (QuadcopterEuler(d5["dt"]), QuadraticCost(d5["Q"], d5["R"]), QuadraticTerminalCost(d5["Qf"]))
        ilqrUtils.differentialDynamicProgramming(*m5, x5, uG5, maxIter=10, tol=-1.0)
        X["ddp"] = timed(lambda: ilqrUtils.differentialDynamicProgramming(*m5, x5, uG5, maxIter=10, tol=-1.0), 1)
        # a1 with genuinely time-varying operands: A[k], B[k], Q[k], R[k] materialised per step and streamed from HBM
        from zopt_b200.lqrUtils import discreteFiniteHorizonLqr
        Atv = A[:, None].expand(-1, N, -1, -1).contiguous()
        Btv = B[:, None].expand(-1, N, -1, -1).contiguous()
        Qtv = Q[:, None].expand(-1, N, -1, -1).contiguous()
        Rtv = R[:, None].expand(-1, N, -1, -1).contiguous()
        discreteFiniteHorizonLqr(Atv, Btv, Qtv, Rtv, N)
        X["tv"] = timed(lambda: discreteFiniteHorizonLqr(Atv, Btv, Qtv, Rtv, N), 5)
        del Atv, Btv, Qtv, Rtv
        # cfg 1's shape (the demos' (n, m) = (8, 4), N = 100) as a batch: discreteFiniteHorizonLqr and bilinearAffineLqr with per-problem
        # A, B constant in time and H, d, q, r series shared by the batch, fp32 (k_riccati_s84) and fp64 (k_riccati_s84d)
        from zopt_b200.lqrUtils import bilinearAffineLqr
        g8 = torch.Generator(device="cpu").manual_seed(8)
        for dt8, tag in ((f32, "f32"), (f64, "f64")):
            c8 = lambda t: t.to(dtype=dt8, device=dev)
            A8 = c8(torch.eye(8) + 0.1 * torch.randn(Bsz, 1, 8, 8, generator=g8)).expand(-1, 100, -1, -1)
            B8 = c8(0.3 * torch.randn(Bsz, 1, 8, 4, generator=g8)).expand(-1, 100, -1, -1)
            Q8, R8 = c8(torch.eye(8))[None, None].expand(Bsz, 100, -1, -1), c8(torch.eye(4))[None, None].expand(Bsz, 100, -1, -1)
            H8 = c8(0.1 * torch.randn(1, 100, 4, 8, generator=g8)).expand(Bsz, -1, -1, -1)
            d8, q8 = (c8(s8 * torch.randn(1, 100, 8, generator=g8)).expand(Bsz, -1, -1) for s8 in (0.01, 0.1))
            r8, q08 = c8(0.05 * torch.randn(1, 100, 4, generator=g8)).expand(Bsz, -1, -1), c8(torch.zeros(1, 100)).expand(Bsz, -1)
            discreteFiniteHorizonLqr(A8, B8, Q8, R8, 100)
            X["dfh8_" + tag] = timed(lambda: discreteFiniteHorizonLqr(A8, B8, Q8, R8, 100), 5)
            bilinearAffineLqr(A8, B8, d8, Q8, R8, H8, q8, r8, q08, 100)
            X["bil8_" + tag] = timed(lambda: bilinearAffineLqr(A8, B8, d8, Q8, R8, H8, q8, r8, q08, 100), 5)
        del A8, B8, Q8, R8, H8, d8, q8, r8, q08
        # the headline workload in the reference's own precision (fp64): cooperative four-threads-per-problem Riccati kernel
        step64(xb64)
        X["f64"] = timed(lambda: step64(xb64), 5)
        del xb64, ub64, Q64, R64
        # cfg 3 tier B: the reference demo's box-constrained lqrMpc (demos/lqrMpc.py:11-32: hover linearisation, N=25, bounds
        # |uvw|<=1, |pq|<=0.3, |r|<=0.1, |phi,theta|<=0.5, |u|<=3, OSQP eps 1e-2), one solve per initial state, bounds bind
        Ab, Bb = Quadcopter().linearizeInertial(np.zeros(12), configs.U_TRIM, 0.1)
        x_ub = np.array([1, 1, 1, 0.3, 0.3, 0.1, 0.5, 0.5, np.inf, np.inf, np.inf, np.inf])
        u_ub = np.full(4, 3.0)
        xb = np.zeros((16384, 12))
        xb[:, 9:12] = np.random.default_rng(1234 + 3).uniform(-10, 10, (16384, 3))
        xb = torch.as_tensor(xb[lo:hi], dtype=f32, device=dev)
        pb = lqrMpc(Ab.to(f32), Bb.to(f32), torch.eye(12, dtype=f32, device=dev), torch.eye(4, dtype=f32, device=dev), 25,
                    -x_ub, x_ub, -u_ub, u_ub)
        _, _, stb = pb.solve(xb, eps_abs=1e-2, eps_rel=1e-2)
        X["box"] = timed(lambda: pb.solve(xb, eps_abs=1e-2, eps_rel=1e-2), 3)
        info["box_iters"] = float(pb.iters.float().mean())
        info["box_opt"] = float((stb == 0).float().mean())
        # ... and the demo's receding-horizon loop around it (demos/lqrMpc.py:42-47): 200 steps, warm-started, one fused kernel
        _, stc = pb.closedLoop(xb, 200, eps_abs=1e-2, eps_rel=1e-2)
        X["boxcl"] = timed(lambda: pb.closedLoop(xb, 200, eps_abs=1e-2, eps_rel=1e-2), 1)
        info["boxcl_iters"] = float(pb.iters.float().mean()) / 200
        info["boxcl_opt"] = float((stc == 0).float().mean())
        # same loop, termination checked every 5 ADMM iterations instead of OSQP's default 25 (a warm-started solve converges in ~5)
        _, stc5 = pb.closedLoop(xb, 200, eps_abs=1e-2, eps_rel=1e-2, check_termination=5)
        X["boxcl5"] = timed(lambda: pb.closedLoop(xb, 200, eps_abs=1e-2, eps_rel=1e-2, check_termination=5), 1)
        info["boxcl5_iters"] = float(pb.iters.float().mean()) / 200
        info["boxcl5_opt"] = float((stc5 == 0).float().mean())

    names = sorted(X)
    base = [ms, e2e_ms, k_ms, e2e_u_ms, kd_ms, roof_ms, e2e64_ms, gather_ms]
    times = torch.tensor(base + [X[k] for k in names], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    vals = [float(v) for v in times.cpu()]
    ms, e2e_ms, k_ms, e2e_u_ms, kd_ms, roof_ms, e2e64_ms, gather_ms = vals[:len(base)]
    X = dict(zip(names, vals[len(base):]))

    if rank == 0:
        total = Bsz * world
        value = total * args.steps / (ms * 1e-3)
        e2e_val = total * args.steps / (e2e_ms * 1e-3)
        import ctypes as C
        peak32, clk = C.c_double(0), C.c_double(0)
        _lib.check(_lib.lib.zb_peak_fma(0, local_rank, C.byref(peak32), C.byref(clk)))
        peak64 = C.c_double(0)
        _lib.check(_lib.lib.zb_peak_fma(1, local_rank, C.byref(peak64), C.byref(clk)))
        p32, p64 = peak32.value / 1e12, peak64.value / 1e12  # per GPU
        ach_tf = FLOP_PER_SOLVE * Bsz / (k_ms * 1e-3) / 1e12
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = peaks.get("hbm_gbs", 6650.0)
        alg_bytes = (1456 + 9600 * 2 + 3248) * Bsz  # in: A,B,Q,R,Qf,x0; gains written + re-read; plan written (fp32)
        traffic, traffic_src = None, None
        try:  # DRAM bytes of the dominant kernel from the committed ncu capture (same kernel, same batch); not measured in this run
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            if tj.get("batch") == Bsz and tj.get("N") == N:
                traffic = tj["dram_bytes_read"] + tj["dram_bytes_write"]
                traffic_src = "profiles/traffic.json (ncu --set full capture of this kernel at this batch; not re-measured in this run)"
        except Exception:
            pass
        roof = {"bound": "fp32_fma", "achieved": ach_tf, "peak": p32, "unit": "TFLOP/s",
                "frac": ach_tf / p32, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes": alg_bytes,
                "peak_source": "zb_peak_fma dependent-FMA probe measured in this run (MEASURED_PEAKS.json has no FP32 CUDA-core figure)",
                "kernel": "k_riccati_t1<MPC,QDIAG>: lqrMpc.solve launch (Riccati sweep + rollout)", "kernel_ms": k_ms,
                "kernel_timing": f"{k_reps} back-to-back launches between one CUDA-event pair on the launching stream",
                "algorithmic_flop_per_solve": FLOP_PER_SOLVE,
                "dense_cost_kernel_ms": kd_ms, "dense_cost_frac": FLOP_PER_SOLVE * Bsz / (kd_ms * 1e-3) / 1e12 / p32,
                "hbm_achieved_gbs": alg_bytes / (k_ms * 1e-3) / 1e9, "hbm_peak_gbs": hbm_peak,
                "hbm_frac": alg_bytes / (k_ms * 1e-3) / 1e9 / hbm_peak,
                "hbm_peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback"}
        # the full step (Bsz problems) through the compiled restatement, about ten seconds of CPU work in all; the torch port
        # (no compiler on the box) keeps its 16,384-problem sample
        if c_oracle_or_none() is not None:
            r0 = time_cpu(4096, 1, 1)[0]
            cpu_n = Bsz
            cpu_steps = int(max(3, min(50, 10.0 * r0 / cpu_n)))
        else:
            cpu_n, cpu_steps = args.cpu_sample, 3
        cpu_val, cpu_sec, threads, cpu_how = time_cpu(cpu_n, cpu_steps, 1)
        # copy roof: all ranks moving the step's D2H payload at once; e2e as a fraction of it
        roof_gbs = d2h * world / (roof_ms * 1e-3) / 1e9
        e2e_gbs = e2e_val * (d2h / Bsz) / 1e9
        e2e = {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "returns": "u (Bsz,4), plan xTraj (Bsz,51,12) + uTraj (Bsz,50,4), status -- PCIe-bound",
               "d2h_gbs": e2e_gbs, "pcie_roof_gbs": roof_gbs, "frac_of_pcie_roof": e2e_gbs / roof_gbs,
               "pcie_roof_how": "the step's D2H payload alone (pinned, same buffers), every rank at once, CUDA events, max over ranks; aggregate GB/s",
               "control_only_value": total * args.steps / (e2e_u_ms * 1e-3), "control_only_d2h_bytes_per_step": Bsz * (NU * 4 + 1),
               "fp64_value": (total / (e2e64_ms * 1e-3)) if e2e64_ms else None, "fp64_d2h_bytes_per_step": d2h64, "fp64_h2d_bytes_per_step": h2d64,
               "gather_ms": gather_ms, "gather_bytes": (Bsz * world * (NU + (N + 1) * NX + N * NU) * 4) if world > 1 else 0,
               "cpu_binding": binding["how"], "cpu_binding_numa_node": binding["numa_node"], "cpus_bound": binding["cpus"]}
        extra = None
        tail = ""
        if not args.no_extras:
            def fl(units, flop, t_ms, peak):  # strong-scaled extras: the job's flops against `world` GPUs' peak
                a = units * flop / (t_ms * 1e-3) / 1e12
                return {"achieved": a, "peak": peak * world, "unit": "TFLOP/s", "frac": a / (peak * world), "n_gpus": world}
            cl, clw, cl64, il, ild, ddp, tv, f64t = X["cl"], X["clw"], X["cl64"], X["il"], X["il_default"], X["ddp"], X["tv"], X["f64"]
            extra = {
                "cfg3_closed_loop_mpc": {"value": 16384 * 200 / (cl * 1e-3), "unit": "MPC solves/s", "ms": cl,
                                         "workload": "16,384 problems total (sharded over ranks), 200 sim steps, horizon 50, fp32, "
                                                     "re-linearised every step, bounds inactive, one fused kernel",
                                         "scaling": "strong", "roofline": dict(bound="fp32_fma", **fl(16384 * 200, FLOP_PER_SOLVE, cl, p32))},
                "cfg3_closed_loop_mpc_weak": {"value": 16384 * world * 200 / (clw * 1e-3), "unit": "MPC solves/s", "ms": clw,
                                              "workload": "as cfg3 but 16,384 problems PER GPU", "scaling": "weak"},
                "cfg3_one_shard_of_eight": {"value": 2048 * 200 / (X["cl_shard"] * 1e-3), "unit": "MPC solves/s", "ms": X["cl_shard"],
                                            "workload": "the 2,048 problems ONE GPU gets when cfg3 is sharded over eight (measured at N=1 only; at N>1 it is "
                                                        "cfg3_closed_loop_mpc itself): nine lanes per problem, work rotation over one-warp workers",
                                            "strong_scaling_1_to_8_implied": cl / X["cl_shard"] if world == 1 else None, "scaling": "strong"},
                "cfg3_closed_loop_mpc_fp64": {"value": 16384 * 200 / (cl64 * 1e-3), "unit": "MPC solves/s", "ms": cl64,
                                              "workload": "cfg3 (16,384 problems total, sharded over ranks) in fp64, the reference's own precision: "
                                                          "one fused cooperative kernel", "scaling": "strong",
                                              "roofline": dict(bound="fp64_fma", **fl(16384 * 200, FLOP_PER_SOLVE, cl64, p64))},
                "cfg4_ilqr": {"value": 16384 * 10 / (il * 1e-3), "unit": "problem-iterations/s", "ms": il,
                              "workload": "16,384 problems total (sharded over ranks), N=200, 10 iterations, 16-way line search, fp64",
                              "scaling": "strong", "algorithmic_flop_per_problem_iteration": 4.66e6,
                              "fp32_alpha_sequence_mismatches": info["ilqr_fp32_alpha_mismatch"], "fp32_alpha_subset": info["ilqr_fp32_alpha_subset"],
                              "roofline": dict(bound="fp64_fma", **fl(16384 * 10, 4.66e6, il, p64))},
                "cfg4_ilqr_reference_defaults": {"ms": ild, "iterations_mean_rank0": info["il_default_iters_mean"],
                                                 "iterations_max_rank0": info["il_default_iters_max"], "converged_fraction_rank0": info["il_default_converged"],
                                                 "value": 16384 * info["il_default_iters_mean"] / (ild * 1e-3), "unit": "problem-iterations/s",
                                                 "workload": "cfg4 problems with the reference's defaults maxIter=100, tol=1e-3 (ilqrUtils.py:267-268): "
                                                             "early exit on the device, cost follows the iterations taken", "scaling": "strong"},
                "cfg5_ddp": {"value": 16384 * 10 / (ddp * 1e-3), "unit": "problem-iterations/s", "ms": ddp,
                             "workload": "16,384 problems total (sharded over ranks), N=100, 10 iterations, eigen-clamped second-order "
                                         "terms every step, fp64", "scaling": "strong",
                             "algorithmic_flop_per_problem_iteration": 7.33e6,
                             "roofline": dict(bound="fp64_fma", **fl(16384 * 10, 7.33e6, ddp, p64))},
                "cfg2_fp64": {"value": Bsz * world / (f64t * 1e-3), "unit": "solves/s", "ms": f64t,
                              "workload": f"the headline step (linearise + lqrMpc.solve, N={N}) in fp64, the reference's own precision, "
                                          f"{Bsz} problems per GPU", "scaling": "weak",
                              "roofline": {"bound": "fp64_fma", "achieved": FLOP_PER_SOLVE * Bsz / (f64t * 1e-3) / 1e12,
                                           "peak": p64, "unit": "TFLOP/s", "frac": FLOP_PER_SOLVE * Bsz / (f64t * 1e-3) / 1e12 / p64}},
                "lqr_time_varying": {"value": Bsz * world / (tv * 1e-3), "unit": "solves/s", "ms": tv,
                                     "workload": f"discreteFiniteHorizonLqr with A[k], B[k], Q[k], R[k] materialised per step "
                                                 f"(streamed from HBM), {Bsz} problems per GPU, N={N}, fp32", "scaling": "weak",
                                     "roofline": {"bound": "hbm", "achieved": Bsz * N * 1600 / (tv * 1e-3) / 1e9,
                                                  "peak": hbm_peak, "unit": "GB/s",
                                                  "frac": Bsz * N * 1600 / (tv * 1e-3) / 1e9 / hbm_peak,
                                                  "algorithmic_bytes_per_problem_step": 1600,
                                                  "note": "SURVEY 8d: A 576 + B 192 + Q 576 + R 64 in, gains 192 out (fp32); the kernel "
                                                          "fetches only the lower-triangle chunks of Q (384 B)"}},
                **{f"cfg1_shape_8x4_{name}_{tag}": {
                    "value": Bsz * world / (X[key + tag] * 1e-3), "unit": "solves/s", "ms": X[key + tag],
                    "workload": f"{fn} at the demos' shape (n, m) = (8, 4), N = 100, {Bsz} problems per GPU, {tag}: per-problem A, B constant in "
                                "time" + (", H, d, q, r series shared by the batch" if name == "bilinear" else ""), "scaling": "weak",
                    "roofline": {"bound": tag.replace("f", "fp") + "_fma", "achieved": Bsz * 100 * flop / (X[key + tag] * 1e-3) / 1e12, "peak": pk,
                                 "unit": "TFLOP/s", "frac": Bsz * 100 * flop / (X[key + tag] * 1e-3) / 1e12 / pk,
                                 "frac_executed": Bsz * 100 * flop_exec / (X[key + tag] * 1e-3) / 1e12 / pk,
                                 "note": f"SURVEY 8d dense count at (8,4): {flop} flop per problem-step; the kernel executes {flop_exec} (symmetric V "
                                         "as a lower triangle, V' = Q + A'W - M'L): frac_executed is the FMA pipe's share"}}
                   for name, fn, key, flop, fe32, fe64 in (("dfh", "discreteFiniteHorizonLqr", "dfh8_", 5093, 3284, 3404),
                                                           ("bilinear", "bilinearAffineLqr", "bil8_", 5493, 3692, 3812))
                   for tag, pk, flop_exec in (("f32", p32, fe32), ("f64", p64, fe64))},
                "cfg3_box_constrained_mpc": {"value": 16384 / (X["box"] * 1e-3), "unit": "solves/s", "ms": X["box"],
                                             "admm_iterations_mean_rank0": info["box_iters"], "optimal_fraction_rank0": info["box_opt"],
                                             "workload": "16,384 initial states total (sharded over ranks), the reference demo's box-constrained "
                                                         "lqrMpc (hover linearisation, N=25, demo bounds, eps 1e-2 as demos/lqrMpc.py:32), fp32, "
                                                         "bounds bind (10 m offsets, |v| <= 1)", "scaling": "strong"},
                "cfg3_box_constrained_mpc_closed_loop": {"value": 16384 * 200 / (X["boxcl"] * 1e-3), "unit": "MPC solves/s", "ms": X["boxcl"],
                                                         "admm_iterations_per_solve_rank0": info["boxcl_iters"], "optimal_fraction_rank0": info["boxcl_opt"],
                                                         "workload": "the same problem in the demo's receding-horizon loop (demos/lqrMpc.py:42-47): "
                                                                     "16,384 initial states total, 200 steps, clip + warm-started solve + perfect "
                                                                     "tracking, one fused kernel, fp32, eps 1e-2", "scaling": "strong"},
                "cfg3_box_constrained_mpc_closed_loop_check5": {"value": 16384 * 200 / (X["boxcl5"] * 1e-3), "unit": "MPC solves/s", "ms": X["boxcl5"],
                                                                "admm_iterations_per_solve_rank0": info["boxcl5_iters"], "optimal_fraction_rank0": info["boxcl5_opt"],
                                                                "workload": "as above with check_termination=5 (OSQP default is 25)", "scaling": "strong"}}
            # compact digest as the LAST key of the line: the driver's record keeps the tail of stdout
            shard_note = f" (one shard of eight: {X['cl_shard']:.2f} ms)" if world == 1 else ""
            tail = (f"N={world} e2e {e2e_val / 1e6:.1f}M/s ({e2e_gbs:.0f} of {roof_gbs:.0f} GB/s copy roof) | cfg3 strong 16384x200: {cl:.2f} ms "
                    f"{16384 * 200 / cl / 1e3:.1f}M MPC/s{shard_note}; weak {clw:.2f} ms; fp64 {cl64:.1f} ms | cfg4 iLQR f64 {il:.1f} ms {16384 * 10 / il / 1e3:.2f}M it/s; "
                    f"defaults {ild:.1f} ms @{info['il_default_iters_mean']:.1f} it | cfg5 DDP {ddp:.1f} ms | tv {tv:.2f} ms | f64 {f64t:.2f} ms | "
                    f"(8,4) N=100 dfh/bilinear f32 {X['dfh8_f32']:.2f}/{X['bil8_f32']:.2f} ms f64 {X['dfh8_f64']:.2f}/{X['bil8_f64']:.2f} ms")
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(Bsz), "clocks": clocks, "e2e": e2e,
            "gpu_launches": 2 * args.steps, "roofline": roof, "extra": extra,
            "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": threads, "kind": "port",
                             "sample": f"{cpu_n} problems of the same workload per step, {cpu_steps} steps of {cpu_sec:.2f} s after 1 warm-up; {cpu_how}; JAX is not installed"},
            "host": hostbind.host_topology(), "digest": tail,
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
