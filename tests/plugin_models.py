"""Symbolic test models for zopt_b200.plugin (shared by the CPU build test, the GPU parity tests and __graft_entry__.build)."""
import sympy as sp

DT = 0.1


def pendulum(x, u):
    """damped pendulum with a state-dependent torque arm: n = 2, m = 1"""
    return [x[0] + DT * x[1], x[1] + DT * (u[0] * (1 - sp.Rational(1, 10) * x[1] ** 2) - sp.sin(x[0]) - sp.Rational(1, 20) * x[1])]


def car(x, u):
    """planar car [px, py, heading, speed], controls [acceleration, turn rate]; f_ux and f_uu are non-zero: n = 4, m = 2"""
    px, py, th, v = x
    a, w = u
    return [px + DT * v * sp.cos(th), py + DT * v * sp.sin(th), th + DT * w * (1 + sp.Rational(1, 10) * v),
            v + DT * (a - sp.Rational(1, 20) * v ** 2 - sp.Rational(1, 10) * a ** 2 * sp.tanh(v))]


MODELS = {"pendulum": (pendulum, 2, 1), "car": (car, 4, 2)}


# ---- user-defined (non-quadratic) costs, compiled together with a model (plugin.SymbolicCost) ----
def car_cost(x, u):
    """smooth-abs distance to the origin, a soft speed limit, a heading term and control effort with a cross term (c_ux != 0)"""
    px, py, th, v = x
    a, w = u
    return (sp.sqrt(px ** 2 + py ** 2 + sp.Rational(1, 10)) + sp.Rational(1, 2) * (1 - sp.cos(th)) + sp.Rational(1, 5) * sp.log(1 + sp.exp(4 * (v - 2)))
            + sp.Rational(1, 10) * a ** 2 + sp.Rational(1, 5) * w ** 2 + sp.Rational(1, 20) * a * v)


def car_terminal(x):
    px, py, th, v = x
    return 10 * (px ** 2 + py ** 2) + 2 * (1 - sp.cos(th)) + v ** 2 + sp.Rational(1, 2) * px ** 2 * py ** 2


def pendulum_cost(x, u):
    """swing-up style cost: 1 - cos (indefinite Hessian away from the bottom: the eigen-clamp is active) + effort"""
    return (1 + sp.cos(x[0])) + sp.Rational(1, 10) * x[1] ** 2 + sp.Rational(1, 20) * u[0] ** 2


def pendulum_terminal(x):
    return 5 * (1 + sp.cos(x[0])) + x[1] ** 2


COSTS = {"pendulum": (pendulum_cost, pendulum_terminal), "car": (car_cost, car_terminal)}
_COST_CACHE = {}


def build(name):
    from zopt_b200.plugin import SymbolicDynamics
    f, n, m = MODELS[name]
    return SymbolicDynamics(f, n, m)


def build_with_cost(name):
    """(model bound to its cost, the SymbolicCost) -- one plug-in library per (model, cost) pair"""
    from zopt_b200.plugin import SymbolicCost
    if name not in _COST_CACHE:
        mdl = build(name)
        c, cf = COSTS[name]
        cost = SymbolicCost(c, cf, mdl.n, mdl.m)
        _COST_CACHE[name] = (mdl, mdl.with_cost(cost), cost)
    return _COST_CACHE[name]
