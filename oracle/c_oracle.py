"""
TEST INFRASTRUCTURE -- ctypes binding of oracle/c/zopt_oracle.c, the plain-C (C99 + OpenMP, fp64) restatement of the
reference's headline path: quadcopter dynamics and exact linearisation (zopt/quadcopter.py:23-144, 178-201), the Riccati
recursion as written (zopt/lqrUtils.py:144-173) and one unconstrained lqrMpc step for a batch (zopt/mpcUtils.py:12-81).
Imported only by tests/ and by bench.py's CPU legs; never by the product.  `make -C oracle/c` builds it (so does
`__graft_entry__.build()`); `load()` builds on demand when gcc is there.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_lib = None


def _cpu_key():
    """the library is compiled with -march=native, so it is keyed by the host CPU: a prebuilt copy that travelled to another
    machine (gpurun ships built .so files) is not loaded there, a fresh one is compiled instead"""
    import hashlib
    try:
        with open("/proc/cpuinfo") as fh:
            lines = [ln for ln in fh if ln.startswith(("model name", "flags"))][:2]
    except OSError:
        lines = []
    return hashlib.sha1("".join(lines).encode()).hexdigest()[:10]


SO = os.path.join(_HERE, "_build", f"libzopt_oracle_{_cpu_key()}.so")


def build():
    """compile to a private name and publish with an atomic rename: several processes (ranks, xdist workers) may build at once"""
    tmp = f"{SO}.{os.getpid()}.tmp"
    r = subprocess.run(["make", "-B", "-C", os.path.join(_HERE, "c"), f"OUT={tmp}"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("building oracle/c failed:\n" + r.stdout[-2000:] + r.stderr[-2000:])
    os.replace(tmp, SO)
    return SO


def load():
    global _lib
    if _lib is None:
        src = os.path.join(_HERE, "c", "zopt_oracle.c")
        if not os.path.exists(SO) or (os.path.exists(src) and os.path.getmtime(src) > os.path.getmtime(SO)):
            build()
        L = C.CDLL(SO)
        dp = C.POINTER(C.c_double)
        L.zo_quad_inertial_dynamics.argtypes = [dp, dp, dp, dp]
        L.zo_quad_linearize.argtypes = [dp, dp, dp, C.c_double, dp, dp]
        L.zo_dfh_lqr.argtypes = [C.c_int] * 4 + [dp] * 6
        L.zo_dfh_lqr.restype = C.c_int
        L.zo_lqr_mpc_solve_batch.argtypes = [C.c_longlong, C.c_int, dp, dp, dp, dp, C.c_double, C.c_double, dp, dp, dp, C.c_int]
        L.zo_lqr_mpc_solve_batch.restype = C.c_int
        L.zo_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _c(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def inertialDynamics(x, u, wind_ned=None):
    x, u = _c(x), _c(u)
    w = _c(wind_ned) if wind_ned is not None else None
    out = np.empty(12)
    load().zo_quad_inertial_dynamics(_p(x), _p(u), _p(w), _p(out))
    return out


def linearizeInertial(x, u, dt=0.0, wind_ned=None):
    x, u = _c(x), _c(u)
    w = _c(wind_ned) if wind_ned is not None else None
    A, B = np.empty((12, 12)), np.empty((12, 4))
    load().zo_quad_linearize(_p(x), _p(u), _p(w), float(dt), _p(A), _p(B))
    return A, B


def discreteFiniteHorizonLqr(A, B, Q, R, N, return_V0=False):
    A, B, Q, R = _c(A), _c(B), _c(Q), _c(R)
    T, n, m = Q.shape[0], B.shape[-2], B.shape[-1]
    L, V0 = np.empty((N, m, n)), np.empty((n, n))
    rc = load().zo_dfh_lqr(n, m, N, T, _p(A), _p(B), _p(Q), _p(R), _p(L), _p(V0))
    if rc:
        raise RuntimeError(f"zo_dfh_lqr failed ({rc})")
    return (L, V0) if return_V0 else L


def lqrMpcSolveBatch(xbar, ubar, qdiag, rdiag, N, dt, qf_scale=10.0, threads=0):
    """one unconstrained lqrMpc step for every row of (xbar, ubar): returns u0 (B,4), xTraj (B,N+1,12), uTraj (B,N,4)"""
    xbar, ubar, qdiag, rdiag = _c(xbar), _c(ubar), _c(qdiag), _c(rdiag)
    Bsz = xbar.shape[0]
    u0, xT, uT = np.empty((Bsz, 4)), np.empty((Bsz, N + 1, 12)), np.empty((Bsz, N, 4))
    bad = load().zo_lqr_mpc_solve_batch(Bsz, int(N), _p(xbar), _p(ubar), _p(qdiag), _p(rdiag), float(qf_scale), float(dt), _p(u0), _p(xT),
                                        _p(uT), int(threads))
    if bad:
        raise RuntimeError(f"{bad} problems hit a zero pivot")
    return u0, xT, uT


def num_threads():
    return int(load().zo_num_threads())
