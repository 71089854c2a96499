"""
Quadcopter model -- mirror of zopt/quadcopter.py:10-201.

Same class, method names and argument meaning as the reference.  The dynamics, its Jacobian and
the costate-contracted Hessian are analytic CUDA device code (csrc/quad_model_gen.cuh, generated
offline with sympy and checked against autodiff of the oracle); every method accepts one extra
leading batch axis.  States: [u,v,w,p,q,r,phi,theta,psi,x,y,z]; controls [thrust,mx,my,mz].
"""
import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import check, dcode, lib, pick_device, ptr, stream_ptr, to_dev


def _wind(w):
    if w is None:
        return None
    if isinstance(w, torch.Tensor):
        w = w.detach().cpu().numpy()
    w = np.asarray(w, dtype=np.float64).reshape(3)
    return (C.c_double * 3)(*w)


def _prep(x, u):
    device = pick_device(x, u)
    dtype = _lib.pick_dtype(x, u)
    return to_dev(x, dtype, device).contiguous(), to_dev(u, dtype, device).contiguous(), dtype, device


def quad_xdot(x, u, wind_ned=None):
    """xdot (P,12) = inertialDynamics(x (P,12), u (P,4), wind_ned)"""
    x, u, dtype, device = _prep(x, u)
    out = torch.empty_like(x)
    check(lib.zb_quad_dynamics(dcode(dtype), device.index, stream_ptr(device), x.shape[0], ptr(x), ptr(u), _wind(wind_ned),
                               ptr(out)))
    return out


def quad_linearize(x, u, wind_ned=None, dt=0.0):
    """(A (P,12,12), B (P,12,4)): continuous Jacobians when dt == 0, else I + dt*A, dt*B"""
    x, u, dtype, device = _prep(x, u)
    P = x.shape[0]
    A = torch.empty((P, 12, 12), dtype=dtype, device=device)
    B = torch.empty((P, 12, 4), dtype=dtype, device=device)
    check(lib.zb_quad_linearize(dcode(dtype), device.index, stream_ptr(device), P, ptr(x), ptr(u), _wind(wind_ned),
                                float(dt), ptr(A), ptr(B)))
    return A, B


def quad_hess_contract(x, u, wind_ned, dt, lam):
    """H (P,12,12) = s * sum_i lam_i d2F_i/dx2, s = dt or 1 (the v_x . f_xx block of ilqrUtils.py:240)"""
    x, u, dtype, device = _prep(x, u)
    lam = to_dev(lam, dtype, device).contiguous()
    P = x.shape[0]
    H = torch.empty((P, 12, 12), dtype=dtype, device=device)
    check(lib.zb_quad_hess_contract(dcode(dtype), device.index, stream_ptr(device), P, ptr(x), ptr(u), _wind(wind_ned),
                                    float(dt), ptr(lam), ptr(H)))
    return H


def _rot_b2i(phi, theta, psi):
    cphi, sphi, cth, sth, cpsi, spsi = (torch.cos(phi), torch.sin(phi), torch.cos(theta), torch.sin(theta),
                                        torch.cos(psi), torch.sin(psi))
    rows = [
        [cth * cpsi, sphi * sth * cpsi - cphi * spsi, cphi * sth * cpsi - sphi * spsi],
        [cth * spsi, sphi * sth * spsi + cphi * cpsi, cphi * sth * spsi - sphi * cpsi],
        [-sth, sphi * cth, cphi * cth],
    ]
    return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)


class Quadcopter():
    """Quadcopter object (zopt/quadcopter.py:10-20)"""

    def __init__(self):
        self.g = 9.807  # gravity (m / s**2)
        self.m = 2.5  # mass (kg)
        self.I = torch.eye(3, dtype=torch.float64)  # inertia tensor
        self.I_inv = torch.eye(3, dtype=torch.float64)

    # quadcopter.py:23-38 / :41-48 -- small host-side helpers (not on the hot path)
    def _bodyToInertialRotationMatrix(self, phi, theta, psi):
        t = lambda v: torch.as_tensor(v, dtype=torch.float64) if not isinstance(v, torch.Tensor) else v
        return _rot_b2i(t(phi), t(theta), t(psi))

    def _bodyRatesToEulerRatesRotationMatrix(self, phi, theta):
        t = lambda v: torch.as_tensor(v, dtype=torch.float64) if not isinstance(v, torch.Tensor) else v
        phi, theta = t(phi), t(theta)
        sphi, cphi, cth, tth = torch.sin(phi), torch.cos(phi), torch.cos(theta), torch.tan(theta)
        one, zero = torch.ones_like(phi), torch.zeros_like(phi)
        rows = [[one, sphi * tth, cphi * tth], [zero, cphi, -sphi], [zero, sphi / cth, cphi / cth]]
        return torch.stack([torch.stack(r, dim=-1) for r in rows], dim=-2)

    def inertialDynamics(self, state, control, wind_ned=None):
        """xDot = f(x,u) with position states (quadcopter.py:116-144); state (12,) or (Bsz,12)."""
        x, u, dtype, device = _prep(state, control)
        if x.shape[-1] != 12 or u.shape[-1] != 4:
            raise ValueError("inertialDynamics expects state (...,12) and control (...,4)")
        lead = x.shape[:-1]
        out = quad_xdot(x.reshape(-1, 12), u.reshape(-1, 4).expand(x.reshape(-1, 12).shape[0], 4), wind_ned)
        return out.reshape(lead + (12,))

    def rigidBodyDynamics(self, state, control, wind_body=None):
        """Rigid-body dynamics (quadcopter.py:70-113); state [u,v,w,p,q,r,phi,theta] (8 or 9 entries used: first 8).
        Evaluated as the first 8 rows of the 12-state kernel with psi = 0 (they do not depend on psi, x, y, z);
        a body-frame wind is rotated to NED with psi = 0 for that purpose (un-batched input only)."""
        x, u, dtype, device = _prep(state, control)
        lead = x.shape[:-1]
        x12 = torch.zeros(lead + (12,), dtype=dtype, device=device)
        x12[..., :8] = x[..., :8]
        wind_ned = None
        if wind_body is not None:
            wb = torch.as_tensor(np.asarray(wind_body, dtype=np.float64)) if not isinstance(wind_body, torch.Tensor) \
                else wind_body.detach().to("cpu", torch.float64)
            if bool(torch.any(wb != 0)):
                if len(lead) != 0:
                    raise NotImplementedError("rigidBodyDynamics: a non-zero wind_body is supported for un-batched input")
                xs = x.detach().to("cpu", torch.float64)
                wind_ned = _rot_b2i(xs[6], xs[7], torch.zeros(())) @ wb
        return self.inertialDynamics(x12, u, wind_ned)[..., :8]

    def trim(self, uvwTrim):
        """Trim at the given body velocities (quadcopter.py:146-177): BFGS on sum(rigidBodyDynamics**2) from
        z0 = [0]*5 + [g,0,0,0].  Runs once per script on the host; every objective evaluation is a kernel call."""
        import scipy.optimize as spo
        uvw = np.asarray(uvwTrim.detach().cpu() if isinstance(uvwTrim, torch.Tensor) else uvwTrim, dtype=np.float64)
        nxz = 5

        def _getXu(z):
            return np.concatenate([uvw, z[:nxz]]), z[nxz:]

        def trimFunc(z):
            x, u = _getXu(z)
            return float(torch.sum(self.rigidBodyDynamics(x, u)**2))

        z0 = np.concatenate([np.zeros(nxz), [self.g, 0, 0, 0]])
        out = spo.minimize(trimFunc, z0, method="BFGS")
        if not out.success:
            raise RuntimeError("Trim failed")
        return _getXu(out.x)

    def linearize(self, x0, u0, dt=0):
        """Jacobian linearisation of rigidBodyDynamics about (x0,u0), forward-Euler discretised when dt != 0
        (quadcopter.py:179-201).  Returns A (8,8), B (8,4) (leading batch axis allowed)."""
        x, u, dtype, device = _prep(x0, u0)
        lead = x.shape[:-1]
        x12 = torch.zeros(lead + (12,), dtype=dtype, device=device)
        x12[..., :8] = x[..., :8]
        A, B = self.linearizeInertial(x12, u, dt)
        return A[..., :8, :8].contiguous(), B[..., :8, :].contiguous()

    def linearizeInertial(self, x0, u0, dt=0, wind_ned=None):
        """12-state counterpart used by demos/lqrMpc.py:26-28: A = I + dt*dF/dx, B = dt*dF/du of inertialDynamics."""
        x, u, dtype, device = _prep(x0, u0)
        lead = x.shape[:-1]
        xf = x.reshape(-1, 12)
        A, B = quad_linearize(xf, u.reshape(-1, 4).expand(xf.shape[0], 4), wind_ned, dt)
        return A.reshape(lead + (12, 12)), B.reshape(lead + (12, 4))
