"""
Discrete finite-horizon LQR -- mirror of zopt/lqrUtils.py:144-173 and :207-262.

Same names, argument order and return values as the reference.  Batched convention:
every array may carry one extra leading axis `Bsz`; operands without it are shared by the batch.
Returns torch CUDA tensors; dtype follows the inputs (fp32 only when every input is fp32).
"""
import numpy as np
import torch

from ._lib import View, check, dcode, is_symmetric, lib, pick_device, pick_dtype, ptr, stream_ptr, to_dev


def _prep(arrs, core_ndims):
    """Move to device, find the batch size.  core_ndims[i] = ndim of the un-batched operand."""
    device = pick_device(*arrs)
    dtype = pick_dtype(*arrs)
    ts = [to_dev(a, dtype, device) for a in arrs]
    Bsz, batched = None, []
    for t, nd in zip(ts, core_ndims):
        if t.ndim == nd:
            batched.append(False)
        elif t.ndim == nd + 1:
            batched.append(True)
            if t.shape[0] != 1:
                if Bsz is not None and Bsz != t.shape[0]:
                    raise ValueError(f"inconsistent batch sizes {Bsz} and {t.shape[0]}")
                Bsz = t.shape[0]
        else:
            raise ValueError(f"expected {nd} or {nd + 1} dimensions, got shape {tuple(t.shape)}")
    any_batched = any(batched)
    if Bsz is None:
        Bsz = 1
    return device, dtype, ts, batched, any_batched, Bsz


def discreteFiniteHorizonLqr(A, B, Q, R, N, return_value=False, kernel_flags=0):
    """
    Finite-horizon LQR gains by the backward Riccati recursion (zopt/lqrUtils.py:144-173).

    Arguments
    ---------
        A : (T,n,n) or (Bsz,T,n,n), time along the axis before the matrix: `A[k]`
        B : (T,n,m) or (Bsz,T,n,m)
        Q : (T,n,n) or (Bsz,T,n,n); the terminal value is `Q[-1]` (lqrUtils.py:172)
        R : (T,m,m) or (Bsz,T,m,m); Q, R are used as given -- symmetric weights (the normal case) run on the fast
            (12,4) kernels, non-symmetric ones on the generic kernel, with the reference's result either way
        N : horizon (T >= N)

    Returns
    -------
        L : (N,m,n) or (Bsz,N,m,n) gains, `u = -L[k] x`
    """
    device, dtype, (A, B, Q, R), batched, any_b, Bsz = _prep([A, B, Q, R], [3, 3, 3, 3])
    n, m = B.shape[-2], B.shape[-1]
    N = int(N)
    Ts = [t.shape[-3] for t in (A, B, Q, R)]
    T = Q.shape[-3]
    if A.shape[-2:] != (n, n) or Q.shape[-2:] != (n, n) or R.shape[-2:] != (m, m):
        raise ValueError("inconsistent matrix shapes")
    if min(Ts[0], Ts[1], Ts[3]) < N or T < max(N, 1):
        raise ValueError(f"time axis shorter than the horizon N={N}")
    vA, vB, vQ, vR = (View(t, 2, True, b) for t, b in zip((A, B, Q, R), batched))
    L = torch.empty((Bsz, N, m, n), dtype=dtype, device=device)
    V0 = torch.empty((Bsz, n, n), dtype=dtype, device=device) if return_value else None
    # The (12,4) fast kernels keep the symmetric value matrix as a lower triangle and read the lower triangle of Q, R.  The
    # reference uses the weights exactly as given (lqrUtils.py:168-169), so non-symmetric ones take the generic kernel.
    sym_kernel = (n, m) in ((12, 4), (8, 4))
    flags = 0 if (not sym_kernel or (is_symmetric(Q) and is_symmetric(R))) else 128  # ZB_FORCE_GENERIC
    flags |= int(kernel_flags)  # e.g. 256 = ZB_TV_BULK_COPY (tests / experiments)
    check(lib.zb_lqr_dfh_flags(dcode(dtype), device.index, stream_ptr(device), Bsz, N, T, n, m, vA.ref(), vB.ref(), vQ.ref(),
                               vR.ref(), flags, ptr(L), ptr(V0)))
    if not any_b:
        L = L[0]
        V0 = V0[0] if return_value else None
    return (L, V0) if return_value else L


def bilinearAffineLqr(A, B, d, Q, R, H, q, r, q0, N):
    """
    Finite-horizon LQR with bilinear cost and affine dynamics (zopt/lqrUtils.py:207-262).

    Shapes as the reference, each optionally with a leading batch axis:
    A (N,n,n) B (N,n,m) d (N,n) Q (N,n,n) R (N,m,m) H (N,m,n) q (N,n) r (N,m) q0 (N,)

    Returns
    -------
        L : (N,m,n) gains, l : (N,m) offsets
    """
    device, dtype, ts, batched, any_b, Bsz = _prep([A, B, d, Q, R, H, q, r, q0], [3, 3, 2, 3, 3, 3, 2, 2, 1])
    A, B, d, Q, R, H, q, r, q0 = ts
    n, m = B.shape[-2], B.shape[-1]
    N = int(N)
    T = Q.shape[-3]
    if q.shape[-2] != T or q0.shape[-1] != T:
        raise ValueError("Q, q and q0 must have the same time length (the initial carry is their last row)")
    # every operand is indexed up to step N-1 by the kernel: time axes and block shapes are checked here, as for
    # discreteFiniteHorizonLqr (the C entry point only sees T from Q)
    want = {"A": (A, (n, n)), "B": (B, (n, m)), "d": (d, (n,)), "Q": (Q, (n, n)), "R": (R, (m, m)), "H": (H, (m, n)),
            "q": (q, (n,)), "r": (r, (m,))}
    for name, (t, blk) in want.items():
        if tuple(t.shape[t.ndim - len(blk):]) != blk:
            raise ValueError(f"{name}: expected blocks of shape {blk}, got {tuple(t.shape)}")
        if t.shape[t.ndim - len(blk) - 1] < N:
            raise ValueError(f"{name}: time axis ({t.shape[t.ndim - len(blk) - 1]}) shorter than the horizon N={N}")
    if T < max(N, 1):
        raise ValueError(f"time axis shorter than the horizon N={N}")
    blocks = [2, 2, 1, 2, 2, 2, 1, 1, 0]
    views = [View(t, k, True, b) for t, k, b in zip(ts, blocks, batched)]
    L = torch.empty((Bsz, N, m, n), dtype=dtype, device=device)
    l = torch.empty((Bsz, N, m), dtype=dtype, device=device)
    # the (8,4) kernels read the lower triangles of Q and R (lqr_s84.cuh, lqr_s84d.cuh); non-symmetric weights take the as-written kernel
    flags = 0 if ((n, m) != (8, 4) or (is_symmetric(Q) and is_symmetric(R))) else 128  # ZB_FORCE_GENERIC
    check(lib.zb_lqr_bilinear_flags(dcode(dtype), device.index, stream_ptr(device), Bsz, N, T, n, m, *[v.ref() for v in views],
                                    flags, ptr(L), ptr(l)))
    if not any_b:
        return L[0], l[0]
    return L, l


def _sample_coefficient(f, times, dtype, device):
    """A coefficient of finiteHorizonLqr on the integrator's stage times: `f` is a callable of time returning a matrix
    ((p,q) or (Bsz,p,q), NumPy or torch) as in the reference, or a constant matrix.  Returns (tensor, has_time, batched)."""
    if not callable(f):
        t = to_dev(f, dtype, device)
        return t, False, t.ndim == 3
    first = f(float(times[0]))
    vals = [to_dev(first, dtype, device)] + [to_dev(f(float(tq)), dtype, device) for tq in times[1:]]
    batched = vals[0].ndim == 3
    return torch.stack(vals, dim=1 if batched else 0).contiguous(), True, batched


def _care_solve(A, B, Q, R_inv, Qf, T, N, substeps, dtype, device):
    """one fixed-step integration: sample the coefficients at the RK4 stage times, launch zb_lqr_care_rk4"""
    S = (N - 1) * substeps
    times = T - np.arange(2 * S + 1) * (T / S / 2)  # sample j <-> time T - j h/2
    ops = [_sample_coefficient(f, times, dtype, device) for f in (A, B, Q, R_inv)]
    n, m = ops[1][0].shape[-2], ops[1][0].shape[-1]
    for name, (t, _, _), blk in zip(("A", "B", "Q", "R_inv"), ops, ((n, n), (n, m), (n, n), (m, m))):
        if tuple(t.shape[-2:]) != blk:
            raise ValueError(f"{name}: expected {blk} matrices, got {tuple(t.shape)}")
    if tuple(Qf.shape[-2:]) != (n, n) or Qf.ndim not in (2, 3):
        raise ValueError(f"Qf: expected ({n},{n}), got {tuple(Qf.shape)}")
    sizes = [t.shape[0] for t, _, b in ops if b] + ([Qf.shape[0]] if Qf.ndim == 3 else [])
    Bsz = 1
    for b in sizes:
        if b != 1:
            if Bsz not in (1, b):
                raise ValueError(f"inconsistent batch sizes {Bsz} and {b}")
            Bsz = b
    views = [View(t, 2, ht, b) for t, ht, b in ops] + [View(Qf, 2, False, Qf.ndim == 3)]
    V = torch.empty((Bsz, N, n, n), dtype=dtype, device=device)
    check(lib.zb_lqr_care_rk4(dcode(dtype), device.index, stream_ptr(device), Bsz, N, substeps, n, m, T, *[v.ref() for v in views], ptr(V)))
    return V, len(sizes) > 0


def finiteHorizonLqr(A, B, Q, R_inv, Qf, T, N=50, substeps=None, rtol=1.4e-8):
    """
    Finite-horizon LQR gains in continuous time, by integrating the LQR HJB / Riccati differential equation backward from
    `V(T) = Qf` (zopt/lqrUtils.py:39-98):  `-dV/dt = Q + V A + A'V - V B R^-1 B'V`,  `K(t) = R^-1(t) B(t)' V(t)`.

    Arguments (as the reference; every matrix optionally with one leading batch axis)
    ---------
        A, B, Q, R_inv : callables of time `A(t)` ... returning (n,n), (n,m), (n,n), (m,m) matrices -- or constant matrices
        Qf : terminal state cost (n,n)
        T : time horizon;  N : number of output time points (`linspace(0, T, N)`, lqrUtils.py:87)
        substeps, rtol : the reference integrates with jax's adaptive Dormand-Prince at rtol = atol = 1.4e-8.  A CUDA kernel
            cannot call back into Python, so the callables are SAMPLED on the host at the stage times of a fixed-step RK4
            scheme (`substeps` steps between two output points) and the kernel integrates the whole batch.  With
            `substeps=None` (default) the step is halved until two successive solutions agree to `rtol` (Richardson estimate,
            relative to max|V|, fp64; one pass at ~200 steps over the horizon in fp32): the same error control, applied
            globally.  `K.substeps`, `K.err_estimate` report what was used.

    Returns
    -------
        K : callable `K(t)` -> (m,n) or (Bsz,m,n) gains, `V` linearly interpolated between the grid points and clipped
            outside [0, T] exactly as the reference's `interpMapped` (np.interp) does; `K.V` (.., N, n, n), `K.t` (N,) are kept
    """
    N = int(N)
    if N < 2 or (substeps is not None and int(substeps) < 1):
        raise ValueError("need N >= 2 and substeps >= 1")
    T = float(T)
    probe = [f(0.0) if callable(f) else f for f in (A, B, Q, R_inv)] + [Qf]
    device = pick_device(*probe)
    dtype = pick_dtype(*probe)
    Qf = to_dev(Qf, dtype, device)
    err = None
    if substeps is not None:
        sub = int(substeps)
        V, any_b = _care_solve(A, B, Q, R_inv, Qf, T, N, sub, dtype, device)
    else:
        sub = max(4, -(-100 // (N - 1)))
        V, any_b = _care_solve(A, B, Q, R_inv, Qf, T, N, sub, dtype, device)
        while dtype == torch.float64 or err is None:
            Vf, _ = _care_solve(A, B, Q, R_inv, Qf, T, N, 2 * sub, dtype, device)
            err = float((Vf - V).abs().max() / Vf.abs().max()) / 15.0  # RK4: the finer solution carries 1/16 of the difference
            V, sub = Vf, 2 * sub
            if err <= rtol or sub * (N - 1) >= 1 << 16 or dtype != torch.float64:
                break
    tgrid = torch.linspace(0.0, T, N, dtype=dtype, device=device)
    if not any_b:
        V = V[0]

    def K(t):
        tq = min(max(float(t), 0.0), T)  # np.interp clips outside the grid
        pos = tq / T * (N - 1)
        i0 = min(int(np.floor(pos)), N - 2)
        w = pos - i0
        Vt = (1.0 - w) * V[..., i0, :, :] + w * V[..., i0 + 1, :, :]
        Rt = to_dev(R_inv(t) if callable(R_inv) else R_inv, dtype, device)
        Bt = to_dev(B(t) if callable(B) else B, dtype, device)
        return Rt @ Bt.transpose(-1, -2) @ Vt

    K.V, K.t, K.substeps, K.err_estimate = V, tgrid, sub, err
    return K


def proportionalFeedbackController(x, x0, u0, K):
    """zopt/lqrUtils.py:266-269: `u = -K (x - x0) + u0`; no controller states."""
    control = -K @ (x - x0) + u0
    return control, control.new_zeros(0)
