#!/usr/bin/env python
"""bilinearAffineLqr / discreteFiniteHorizonLqr throughput at the reference demos' shape (n=8, m=4, N=100): the fp32 register-resident
kernel k_riccati_s84 (default), the as-written compile-time-size kernels k_bilinear_ct / k_lqr_ct (ZB_NO_S84=1) and the run-time-size
generic kernels (ZB_FORCE_RUNTIME_SIZES=1).  Second argument "tv": every operand a genuine time series shared by the batch
(stride_t != 0: re-staged at every step) instead of per-problem matrices constant in time."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200.lqrUtils import bilinearAffineLqr, discreteFiniteHorizonLqr
Bsz = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
TV = len(sys.argv) > 2 and sys.argv[2] == "tv"
TVFULL = len(sys.argv) > 2 and sys.argv[2] == "tvfull"  # every operand a per-problem time series materialised in HBM (streamed)
n, m, N = 8, 4, 100
rng = np.random.default_rng(3)
for dt in ((torch.float32,) if TVFULL else (torch.float32, torch.float64)):
    c = lambda a: torch.as_tensor(a, dtype=dt, device="cuda")
    A = c(np.eye(n) + 0.1 * rng.normal(size=(Bsz, 1, n, n))).expand(-1, N, -1, -1)
    B = c(0.3 * rng.normal(size=(Bsz, 1, n, m))).expand(-1, N, -1, -1)
    Q = c(np.eye(n))[None, None].expand(Bsz, N, -1, -1); R = c(np.eye(m))[None, None].expand(Bsz, N, -1, -1)
    H = c(0.1 * rng.normal(size=(1, N, m, n))).expand(Bsz, -1, -1, -1)
    d = c(0.01 * rng.normal(size=(1, N, n))).expand(Bsz, -1, -1); q = c(0.1 * rng.normal(size=(1, N, n))).expand(Bsz, -1, -1)
    r = c(0.05 * rng.normal(size=(1, N, m))).expand(Bsz, -1, -1); q0 = c(np.zeros((1, N))).expand(Bsz, -1)
    if TV:
        A = c(np.eye(n) + 0.1 * rng.normal(size=(1, N, n, n))).expand(Bsz, -1, -1, -1); B = c(0.3 * rng.normal(size=(1, N, n, m))).expand(Bsz, -1, -1, -1)
        Q = c(np.eye(n) * (1 + rng.uniform(size=(1, N, 1, 1)))).expand(Bsz, -1, -1, -1); R = c(np.eye(m) * (1 + rng.uniform(size=(1, N, 1, 1)))).expand(Bsz, -1, -1, -1)
        d = c(0.01 * rng.normal(size=(Bsz, N, n))); r = c(0.05 * rng.normal(size=(Bsz, N, m)))
    if TVFULL:
        g = torch.Generator(device="cuda").manual_seed(5)
        rn = lambda *sh: torch.randn(*sh, generator=g, device="cuda", dtype=dt)
        A = torch.eye(n, device="cuda", dtype=dt) + 0.1 * rn(Bsz, N, n, n); B = 0.3 * rn(Bsz, N, n, m)
        Q = (torch.eye(n, device="cuda", dtype=dt) * (1 + torch.rand(Bsz, N, 1, 1, generator=g, device="cuda", dtype=dt))).contiguous()
        R = (torch.eye(m, device="cuda", dtype=dt) * (1 + torch.rand(Bsz, N, 1, 1, generator=g, device="cuda", dtype=dt))).contiguous()
        H = 0.1 * rn(Bsz, N, m, n); d = 0.01 * rn(Bsz, N, n); q = 0.1 * rn(Bsz, N, n); r = 0.05 * rn(Bsz, N, m); q0 = torch.zeros(Bsz, N, device="cuda", dtype=dt)
    for name, fn in (("bilinearAffineLqr", lambda: bilinearAffineLqr(A, B, d, Q, R, H, q, r, q0, N)), ("discreteFiniteHorizonLqr", lambda: discreteFiniteHorizonLqr(A, B, Q, R, N))):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        flop = (5093 if name[0] == "d" else 5093 + 400) * N  # SURVEY 8d count at (8,4)
        print(f"{name} {str(dt)[6:]} Bsz={Bsz} N={N}{' tv' if TV else ''}{' tvfull' if TVFULL else ''}{' ZB_NO_S84' if os.environ.get('ZB_NO_S84') else ''}: {ms:.2f} ms  {Bsz / ms * 1e3:.3e} solves/s  {Bsz * flop / ms / 1e9:.2f} TFLOP/s (dense count)" + (f"  {Bsz * N * 4 * ((64 + 32 + 64 + 16 + 32) + (32 + 8 + 8 + 4 + 4 if name[0] == 'b' else 0)) / ms / 1e6:.0f} GB/s (operands in + gains out)" if TVFULL else ""))
