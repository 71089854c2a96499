"""
Synthetic problem generators for the BASELINE.json configurations.

NumPy only; seeds are `numpy.random.default_rng(1234 + config_index)`; everything is generated in
fp64 and cast by the caller.  dt = 0.1 everywhere.  Used by tests and bench.py.
"""
import numpy as np

DT = 0.1
G = 9.807
U_TRIM = np.array([G, 0.0, 0.0, 0.0])


def quad_states(rng, Bsz):
    """Linearisation points / initial states of cfg 2 and 3: uvw~U(-1,1), pqr~U(-.3,.3),
    phi,theta~U(-.5,.5), psi~U(-pi,pi), xyz~U(-10,10)."""
    x = np.zeros((Bsz, 12))
    x[:, 0:3] = rng.uniform(-1, 1, (Bsz, 3))
    x[:, 3:6] = rng.uniform(-0.3, 0.3, (Bsz, 3))
    x[:, 6:8] = rng.uniform(-0.5, 0.5, (Bsz, 2))
    x[:, 8] = rng.uniform(-np.pi, np.pi, Bsz)
    x[:, 9:12] = rng.uniform(-10, 10, (Bsz, 3))
    return x


def cfg1_double_integrator(N=100):
    """cfg 1b: double integrator n=2, m=1 (defined here; the reference has none).  Q has N+1 rows with
    Q[N] = 10 I, so the terminal value Q[-1] is 10 I."""
    A = np.array([[1.0, DT], [0.0, 1.0]])
    B = np.array([[DT**2 / 2], [DT]])
    Ak = np.repeat(A[None], N, axis=0)
    Bk = np.repeat(B[None], N, axis=0)
    Qk = np.repeat(np.eye(2)[None], N + 1, axis=0)
    Qk[N] = 10 * np.eye(2)
    Rk = np.repeat(np.eye(1)[None], N, axis=0)
    return Ak, Bk, Qk, Rk, N


def cfg1_demo_weights(N=100):
    """cfg 1a: cost arrays of demos/discreteFiniteHorizonLqr.py:29-34 (n=8, m=4): Qk = [10 I, I x 99]."""
    Qk = np.repeat(np.eye(8)[None], N, axis=0)
    Qk[0] = 10 * np.eye(8)
    Rk = np.repeat(np.eye(4)[None], N, axis=0)
    return Qk, Rk


def cfg2(Bsz=65536, seed=1234 + 2):
    """cfg 2: per-problem linearisation point xbar (=x0), ubar = hover, Q_i = diag(10^U(-1,1)),
    R_i = diag(10^U(-1,1)), terminal 10 Q_i, N = 50.  Returns dict of fp64 arrays."""
    rng = np.random.default_rng(seed)
    xbar = quad_states(rng, Bsz)
    ubar = np.tile(U_TRIM, (Bsz, 1))
    qd = 10.0**rng.uniform(-1, 1, (Bsz, 12))
    rd = 10.0**rng.uniform(-1, 1, (Bsz, 4))
    return dict(xbar=xbar, ubar=ubar, qdiag=qd, rdiag=rd, N=50, dt=DT)


def cfg3(Bsz=16384, seed=1234 + 3):
    d = cfg2(Bsz, seed)
    d["sim_steps"] = 200
    return d


def cfg4(Bsz=16384, seed=1234 + 4, N=200):
    """cfg 4: iLQR, x0 = (0.., xyz~U(-10,10)), uGuess = hover, Q=I12, R=I4, terminal 10 I (demos/iterativeLqr.py:22-38)."""
    rng = np.random.default_rng(seed)
    x0 = np.zeros((Bsz, 12))
    x0[:, 9:12] = rng.uniform(-10, 10, (Bsz, 3))
    return dict(x0=x0, uGuess=np.tile(U_TRIM, (N, 1)), Q=np.eye(12), R=np.eye(4), Qf=10 * np.eye(12), N=N, dt=DT)


def cfg5(Bsz=16384, seed=1234 + 5, N=100):
    """cfg 5: DDP, R = 0.2 I4, x0 xyz~U(-5,5) (demos/differentialDynamicProgramming.py:22-38)."""
    rng = np.random.default_rng(seed)
    x0 = np.zeros((Bsz, 12))
    x0[:, 9:12] = rng.uniform(-5, 5, (Bsz, 3))
    return dict(x0=x0, uGuess=np.tile(U_TRIM, (N, 1)), Q=np.eye(12), R=0.2 * np.eye(4), Qf=10 * np.eye(12), N=N,
                dt=DT)


def diag_embed(d):
    out = np.zeros(d.shape + (d.shape[-1],))
    idx = np.arange(d.shape[-1])
    out[..., idx, idx] = d
    return out
