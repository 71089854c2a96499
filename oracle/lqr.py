"""
Oracle (test infrastructure): NumPy fp64 restatement of the discrete Riccati
recursions of the reference.

  discreteFiniteHorizonLqr  <- zopt/lqrUtils.py:144-173 (step :167-170)
  bilinearAffineLqr         <- zopt/lqrUtils.py:207-262 (step :242-259)

Every function accepts the reference's single-problem shapes; the `*_batched`
variants take one extra leading axis and run the same arithmetic with NumPy
broadcasting (used for parity at batch sizes and as the timed CPU baseline).
`np.linalg.solve` is LAPACK LU with partial pivoting, the same primitive
`jnp.linalg.solve` lowers to on CPU.
"""
import numpy as np


def _T(x):
    return np.swapaxes(x, -1, -2)


def discreteFiniteHorizonLqr(A, B, Q, R, N, return_value=False):
    """zopt/lqrUtils.py:144-173.  A (T,n,n) B (T,n,m) Q (T,n,n) R (T,m,m), T >= N.

    Terminal value is Q[-1] -- the LAST row of the Q array whatever N is
    (lqrUtils.py:172) -- and Q[N-1] is also the stage cost at k=N-1.
    Returns L (N,m,n) in natural time order (u = -L x).
    """
    A = np.asarray(A)
    B = np.asarray(B)
    Q = np.asarray(Q)
    R = np.asarray(R)
    n, m = B.shape[-2:]
    L = np.zeros((N, m, n), dtype=np.result_type(A, B, Q, R, np.float32))
    V = Q[-1]
    for k in range(N - 1, -1, -1):
        # lqrUtils.py:168
        Lk = np.linalg.solve(R[k] + B[k].T @ V @ B[k], B[k].T @ V @ A[k])
        # lqrUtils.py:169 (Joseph form, as written)
        Acl = A[k] - B[k] @ Lk
        V = Q[k] + Lk.T @ R[k] @ Lk + Acl.T @ V @ Acl
        L[k] = Lk
    if return_value:
        return L, V
    return L


def discreteFiniteHorizonLqr_batched(A, B, Q, R, N, return_value=False):
    """Same recursion over a leading batch axis: A (Bsz,T,n,n) ... -> L (Bsz,N,m,n)."""
    A = np.asarray(A)
    B = np.asarray(B)
    Q = np.asarray(Q)
    R = np.asarray(R)
    Bsz = A.shape[0]
    n, m = B.shape[-2:]
    L = np.zeros((Bsz, N, m, n), dtype=np.result_type(A, B, Q, R, np.float32))
    V = Q[:, -1]
    for k in range(N - 1, -1, -1):
        Ak, Bk, Qk, Rk = A[:, k], B[:, k], Q[:, k], R[:, k]
        BtV = _T(Bk) @ V
        Lk = np.linalg.solve(Rk + BtV @ Bk, BtV @ Ak)
        Acl = Ak - Bk @ Lk
        V = Qk + _T(Lk) @ Rk @ Lk + _T(Acl) @ V @ Acl
        L[:, k] = Lk
    if return_value:
        return L, V
    return L


def bilinearAffineLqr(A, B, d, Q, R, H, q, r, q0, N, return_value=False):
    """zopt/lqrUtils.py:207-262.  Returns (L (N,m,n), l (N,m)); u = -L x - l."""
    A, B, d, Q, R, H, q, r, q0 = (np.asarray(a) for a in (A, B, d, Q, R, H, q, r, q0))
    n, m = B.shape[1:]
    dt = np.result_type(A, B, d, Q, R, H, q, r, q0, np.float32)
    LArr = np.zeros((N, m, n), dtype=dt)
    lArr = np.zeros((N, m), dtype=dt)
    # lqrUtils.py:261: initial carry (Q[-1], q[-1], q0[-1])
    V, v, v0 = Q[-1], q[-1], q0[-1]
    for k in range(N - 1, -1, -1):
        # lqrUtils.py:244-246
        Su = r[k] + v.T @ B[k] + d[k].T @ V @ B[k]
        Suu = R[k] + B[k].T @ V @ B[k]
        Sux = H[k] + B[k].T @ V @ A[k]
        # :248-249
        L = np.linalg.solve(Suu, Sux)
        l = np.linalg.solve(Suu, Su)
        # :251-253
        VNew = Q[k] + A[k].T @ V @ A[k] - L.T @ Suu @ L
        vNew = q[k] + A[k].T @ (v + V @ d[k]) - Sux.T @ l
        v0New = v0 + q0[k] + d[k].T @ v + 0.5 * d[k].T @ V @ d[k] - 0.5 * l.T @ Su
        V, v, v0 = VNew, vNew, v0New
        LArr[k] = L
        lArr[k] = l
    if return_value:
        return LArr, lArr, (V, v, v0)
    return LArr, lArr


def bilinearAffineLqr_batched(A, B, d, Q, R, H, q, r, q0, N):
    """Leading batch axis on every array; same arithmetic as bilinearAffineLqr."""
    A, B, d, Q, R, H, q, r, q0 = (np.asarray(a) for a in (A, B, d, Q, R, H, q, r, q0))
    Bsz = A.shape[0]
    n, m = B.shape[-2:]
    dt = np.result_type(A, B, d, Q, R, H, q, r, q0, np.float32)
    LArr = np.zeros((Bsz, N, m, n), dtype=dt)
    lArr = np.zeros((Bsz, N, m), dtype=dt)
    V, v, v0 = Q[:, -1], q[:, -1], q0[:, -1]
    for k in range(N - 1, -1, -1):
        Ak, Bk, dk, Qk, Rk, Hk, qk, rk = A[:, k], B[:, k], d[:, k], Q[:, k], R[:, k], H[:, k], q[:, k], r[:, k]
        Vd = np.einsum('bij,bj->bi', V, dk)
        Su = rk + np.einsum('bi,bij->bj', v, Bk) + np.einsum('bi,bij->bj', Vd, Bk)
        Suu = Rk + _T(Bk) @ V @ Bk
        Sux = Hk + _T(Bk) @ V @ Ak
        L = np.linalg.solve(Suu, Sux)
        l = np.linalg.solve(Suu, Su[..., None])[..., 0]
        VNew = Qk + _T(Ak) @ V @ Ak - _T(L) @ Suu @ L
        vNew = qk + np.einsum('bji,bj->bi', Ak, v + Vd) - np.einsum('bji,bj->bi', Sux, l)
        v0 = v0 + q0[:, k] + np.einsum('bi,bi->b', dk, v) + 0.5 * np.einsum('bi,bi->b', dk, Vd) \
            - 0.5 * np.einsum('bi,bi->b', l, Su)
        V, v = VNew, vNew
        LArr[:, k] = L
        lArr[:, k] = l
    return LArr, lArr


def finiteHorizonLqr(A, B, Q, R_inv, Qf, T, N=50):
    """zopt/lqrUtils.py:55-98 (with _lqrHjb :39-52): integrate `dV/ds = Q - V B R^-1 B' V + V A + A' V` in reversed time
    s = T - t from V = Qf, report on `linspace(0, T, N)`, return `K(t) = R^-1(t) B(t)' interp(V)(t)`.
    jax's `odeint` is an adaptive Dormand-Prince 5(4) with rtol = atol = 1.4e-8; SciPy's `RK45` is the same pair and is run
    at the same tolerances (so the two agree to the integrator's tolerance, not bit for bit)."""
    import scipy.integrate as spi
    Qf = np.asarray(Qf, dtype=float)
    n = np.asarray(A(0)).shape[0]
    t = np.linspace(0, T, num=N)

    def hjb(tq, V):  # lqrUtils.py:49-52
        V = V.reshape((n, n))
        At, Bt = np.asarray(A(tq), dtype=float), np.asarray(B(tq), dtype=float)
        dV = -np.asarray(Q(tq), dtype=float) + V @ Bt @ np.asarray(R_inv(tq), dtype=float) @ Bt.T @ V - V @ At - At.T @ V
        return dV.reshape(-1)

    sol = spi.solve_ivp(lambda s, V: -hjb(T - s, V), (0, T), Qf.reshape(-1), method="RK45", t_eval=t, rtol=1.4e-8, atol=1.4e-8)
    out = sol.y.T                      # (N, n*n), out[i] = V(T - t_i)
    V = out[::-1].T                    # lqrUtils.py:94
    Vfun = lambda tq: np.array([np.interp(tq, t, row) for row in V])  # interpMapped (jaxUtils.py:7-24)
    K = lambda tq: np.asarray(R_inv(tq), dtype=float) @ np.asarray(B(tq), dtype=float).T @ Vfun(tq).reshape((n, n))
    K.V = out[::-1].reshape(N, n, n)
    return K
