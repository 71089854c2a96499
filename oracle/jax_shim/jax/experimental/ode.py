"""jax.experimental.ode.odeint stand-in (TEST INFRASTRUCTURE): jax integrates with an adaptive Dormand-Prince 5(4) at
rtol = atol = 1.4e-8; SciPy's RK45 is the same pair, run here at the same tolerances on torch/NumPy fp64 states."""
import numpy as np
import scipy.integrate as spi
import torch


def odeint(func, y0, t, *args, rtol=1.4e-8, atol=1.4e-8, **kw):
    t = np.asarray(t, dtype=float)
    y0 = np.asarray(y0, dtype=float)
    f = lambda s, y: np.asarray(func(torch.as_tensor(y), s, *args), dtype=float)
    sol = spi.solve_ivp(f, (t[0], t[-1]), y0, method="RK45", t_eval=t, rtol=rtol, atol=atol)
    return sol.y.T.copy()  # NumPy: the reference reverses it with out[::-1] (lqrUtils.py:94)
