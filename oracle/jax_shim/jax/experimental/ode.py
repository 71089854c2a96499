def odeint(*a, **k):
    raise NotImplementedError("jax.experimental.ode.odeint is outside the hot path (continuous-time LQR, SURVEY 2 #7)")
