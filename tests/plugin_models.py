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


def build(name):
    from zopt_b200.plugin import SymbolicDynamics
    f, n, m = MODELS[name]
    return SymbolicDynamics(f, n, m)
