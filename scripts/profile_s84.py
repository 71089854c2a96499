#!/usr/bin/env python
"""Two launches of k_riccati_s84 (discreteFiniteHorizonLqr and bilinearAffineLqr at (8,4), N=100, 65,536 problems, fp32) for ncu:
ncu --set full --clock-control none --import-source on -k regex:k_riccati_s84 -o gpurun_out/s84 python scripts/profile_s84.py"""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from zopt_b200.lqrUtils import bilinearAffineLqr, discreteFiniteHorizonLqr
Bsz, n, m, N = 65536, 8, 4, 100
rng = np.random.default_rng(3)
DT = torch.float64 if (len(sys.argv) > 1 and sys.argv[1] == "f64") else torch.float32  # f64: k_riccati_s84d
c = lambda a: torch.as_tensor(a, dtype=DT, device="cuda")
A = c(np.eye(n) + 0.1 * rng.normal(size=(Bsz, 1, n, n))).expand(-1, N, -1, -1)
B = c(0.3 * rng.normal(size=(Bsz, 1, n, m))).expand(-1, N, -1, -1)
Q = c(np.eye(n))[None, None].expand(Bsz, N, -1, -1); R = c(np.eye(m))[None, None].expand(Bsz, N, -1, -1)
H = c(0.1 * rng.normal(size=(1, N, m, n))).expand(Bsz, -1, -1, -1)
d = c(0.01 * rng.normal(size=(1, N, n))).expand(Bsz, -1, -1); q = c(0.1 * rng.normal(size=(1, N, n))).expand(Bsz, -1, -1)
r = c(0.05 * rng.normal(size=(1, N, m))).expand(Bsz, -1, -1); q0 = c(np.zeros((1, N))).expand(Bsz, -1)
discreteFiniteHorizonLqr(A, B, Q, R, N)
bilinearAffineLqr(A, B, d, Q, R, H, q, r, q0, N)
torch.cuda.synchronize()
