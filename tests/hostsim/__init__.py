"""
TEST INFRASTRUCTURE ONLY: host build of the per-problem bodies of the generic CUDA kernels
(zopt_b200/csrc/zb_problems.cuh), used by `-m "not gpu"` tests to check the kernels' arithmetic
against the oracle where no GPU exists.  Never imported by the product package.
"""
import ctypes as C
import os
import subprocess

import numpy as np

from zopt_b200._lib import ZbArr, ZbCost, ZbModel

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libhostsim.so")
_SRC = os.path.join(_HERE, "hostsim.cpp")
_DEPS = [_SRC] + [os.path.join(_HERE, "..", "..", "zopt_b200", "csrc", f)
                  for f in ("zb_problems.cuh", "zb_steps.cuh", "zb_math.cuh", "quad_model_gen.cuh", "mpc_box.cuh", "lqr_s84.cuh", "lqr_s84d.cuh")]


def _build():
    if os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(d) for d in _DEPS):
        return
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", _SO, _SRC])


_build()
hs = C.CDLL(_SO)


def f64(x):
    return np.ascontiguousarray(np.asarray(x, dtype=np.float64))


def arr(a, block_ndim, has_time=True, batched=True):
    """zb_arr over a C-contiguous numpy array shaped ([Bsz,] [T,] block...)."""
    a = np.ascontiguousarray(a)
    es = a.itemsize
    lead = a.ndim - block_ndim
    sb = st = 0
    if has_time:
        st = a.strides[lead - 1] // es if a.shape[lead - 1] > 1 else 0
        if batched:
            sb = a.strides[lead - 2] // es if a.shape[lead - 2] > 1 else 0
    elif batched:
        sb = a.strides[lead - 1] // es if a.shape[lead - 1] > 1 else 0
    z = ZbArr(a.ctypes.data, sb, st)
    z._keep = a
    return z


def P(a):
    return C.c_void_p(a.ctypes.data) if a is not None else C.c_void_p(0)


def model_linear(A, B, batched):
    n, m = B.shape[-2:]
    M = ZbModel()
    M.kind, M.n, M.m, M.has_wind, M.dt = 0, n, m, 0, 0.0
    M.A = arr(A, 2, False, batched)
    M.B = arr(B, 2, False, batched)
    M._keep = (A, B)
    return M


def model_quad(dt, wind=None):
    M = ZbModel()
    M.kind, M.n, M.m, M.dt = 1, 12, 4, float(dt)
    w = [0.0, 0.0, 0.0] if wind is None else [float(v) for v in wind]
    M.has_wind = int(any(v != 0 for v in w))
    for i in range(3):
        M.wind[i] = w[i]
    return M


def cost(Q, R, Qf, batched):
    c = ZbCost()
    c.Q, c.R, c.Qf = arr(Q, 2, False, batched), arr(R, 2, False, batched), arr(Qf, 2, False, batched)
    c._keep = (Q, R, Qf)
    return c
