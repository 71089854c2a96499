"""
Oracle (test infrastructure): torch-CPU restatement of zopt/quadcopter.py.

  Quadcopter.__init__                      <- quadcopter.py:13-20
  _bodyToInertialRotationMatrix            <- quadcopter.py:23-38
  _bodyRatesToEulerRatesRotationMatrix     <- quadcopter.py:41-48
  _getAeroForceMomemnts                    <- quadcopter.py:51-67
  rigidBodyDynamics                        <- quadcopter.py:70-113
  inertialDynamics                         <- quadcopter.py:116-144
  linearize                                <- quadcopter.py:179-201

Written with torch ops so that `torch.func.jacrev/hessian/vmap` differentiate the
SAME callable the way the reference's `jax.jacobian/hessian/vmap` do; the CUDA
kernels' analytic Jacobians/Hessians are checked against this autodiff.
`trim` (quadcopter.py:146-177) is BFGS from u0=[g,0,0,0]; restated with SciPy.
"""
import numpy as np
import torch


def _t(x, like=None):
    if isinstance(x, torch.Tensor):
        return x
    return torch.as_tensor(np.asarray(x, dtype=np.float64))


class Quadcopter:
    def __init__(self):
        # quadcopter.py:15-18
        self.g = 9.807
        self.m = 2.5
        self.I = torch.eye(3, dtype=torch.float64)
        self.I_inv = torch.linalg.inv(self.I)

    # quadcopter.py:23-38
    def _bodyToInertialRotationMatrix(self, phi, theta, psi):
        phi, theta, psi = _t(phi), _t(theta), _t(psi)
        cphi, sphi = torch.cos(phi), torch.sin(phi)
        cth, sth = torch.cos(theta), torch.sin(theta)
        cpsi, spsi = torch.cos(psi), torch.sin(psi)
        rows = [
            [cth * cpsi, sphi * sth * cpsi - cphi * spsi, cphi * sth * cpsi - sphi * spsi],
            [cth * spsi, sphi * sth * spsi + cphi * cpsi, cphi * sth * spsi - sphi * cpsi],
            [-sth, sphi * cth, cphi * cth],
        ]
        return torch.stack([torch.stack(r) for r in rows])

    # quadcopter.py:41-48
    def _bodyRatesToEulerRatesRotationMatrix(self, phi, theta):
        phi, theta = _t(phi), _t(theta)
        sphi, cphi = torch.sin(phi), torch.cos(phi)
        cth, tth = torch.cos(theta), torch.tan(theta)
        one, zero = torch.ones_like(phi), torch.zeros_like(phi)
        rows = [[one, sphi * tth, cphi * tth], [zero, cphi, -sphi], [zero, sphi / cth, cphi / cth]]
        return torch.stack([torch.stack(r) for r in rows])

    # quadcopter.py:51-67
    def _getAeroForceMomemnts(self, state, windBody=None):
        uvw = state[0:3]
        pqr = state[3:6]
        if windBody is None:
            windBody = torch.zeros(3, dtype=state.dtype)
        force_lin = torch.tensor([-0.2, -0.2, -0.3], dtype=state.dtype)
        force_quad = torch.tensor([-0.05, -0.05, -0.1], dtype=state.dtype)
        moment_lin = torch.tensor([-0.1, -0.1, -0.05], dtype=state.dtype)
        uvw_aero = uvw - windBody
        force_aero = force_lin * uvw_aero + force_quad * uvw_aero**2
        moment_aero = moment_lin * pqr
        return force_aero, moment_aero

    # quadcopter.py:70-113
    def rigidBodyDynamics(self, state, control, wind_body=None):
        state, control = _t(state), _t(control)
        uvw = state[0:3]
        pqr = state[3:6]
        phi, theta = state[6], state[7]
        thrust = control[0]
        mxyz = control[1:4]
        d2xyz = torch.stack([-torch.sin(theta), torch.sin(phi) * torch.cos(theta), torch.cos(phi) * torch.cos(theta)])
        R_rates2Eul = self._bodyRatesToEulerRatesRotationMatrix(phi, theta)
        force_aero, moment_aero = self._getAeroForceMomemnts(state, wind_body)
        zero = torch.zeros_like(thrust)
        force_control = self.m * torch.stack([zero, zero, -thrust])
        force_gravity = self.m * self.g * d2xyz
        force_total = force_control + force_aero + force_gravity
        I = self.I.to(state.dtype)
        moment_control = I @ mxyz
        moment_total = moment_control + moment_aero
        # quadcopter.py:108-110 (note: the Coriolis term is divided by m as well)
        uvwDot = (1 / self.m) * (-torch.linalg.cross(pqr, uvw) + force_total)
        pqrDot = self.I_inv.to(state.dtype) @ (-torch.linalg.cross(pqr, I @ pqr) + moment_total)
        phiThetaDot = R_rates2Eul[0:2, :] @ pqr
        return torch.cat([uvwDot, pqrDot, phiThetaDot])

    # quadcopter.py:116-144
    def inertialDynamics(self, state, control, wind_ned=None):
        state, control = _t(state), _t(control)
        uvw = state[0:3]
        pqr = state[3:6]
        phi, theta, psi = state[6], state[7], state[8]
        R_b2i = self._bodyToInertialRotationMatrix(phi, theta, psi)
        R_rates2Eul = self._bodyRatesToEulerRatesRotationMatrix(phi, theta)
        if wind_ned is None:
            wind_ned = torch.zeros(3, dtype=state.dtype)
        wind_body = R_b2i.T @ _t(wind_ned).to(state.dtype)
        xDot_rb = self.rigidBodyDynamics(state[:9], control, wind_body=wind_body)
        psiDot = (R_rates2Eul[2, :] @ pqr).reshape(1)
        xyzDot = R_b2i @ uvw
        return torch.cat([xDot_rb, psiDot, xyzDot])

    # quadcopter.py:146-177
    def trim(self, uvwTrim):
        import scipy.optimize as spo
        uvwTrim = np.asarray(uvwTrim, dtype=np.float64)
        nxz = 5

        def _getXu(z):
            return np.concatenate([uvwTrim, z[:nxz]]), z[nxz:]

        z0 = np.concatenate([np.zeros(nxz), [self.g, 0, 0, 0]])

        def trimFunc(z):
            zt = torch.as_tensor(z, dtype=torch.float64)
            x = torch.cat([torch.as_tensor(uvwTrim), zt[:nxz]])
            val = torch.sum(self.rigidBodyDynamics(x, zt[nxz:])**2)
            return float(val)

        out = spo.minimize(trimFunc, z0, method="BFGS")
        if not out.success:
            raise RuntimeError("Trim failed")
        return _getXu(out.x)

    # quadcopter.py:179-201 (jacobian of rigidBodyDynamics, forward-Euler discretised)
    def linearize(self, x0, u0, dt=0):
        x0, u0 = _t(x0), _t(u0)
        A, B = torch.func.jacrev(self.rigidBodyDynamics, argnums=(0, 1))(x0, u0)
        if dt != 0:
            A = torch.eye(A.shape[0], dtype=A.dtype) + dt * A
            B = dt * B
        return A, B

    # Caller-side discretisations used by the demos (not methods of the reference class):
    #   demos/iterativeLqr.py:35   dynFun = x + dt * inertialDynamics(x, u)
    #   demos/lqrMpc.py:26-28      A = I + dt * d(inertialDynamics)/dx, B = dt * d/du
    def eulerStep(self, dt, wind_ned=None):
        return lambda x, u: x + dt * self.inertialDynamics(x, u, wind_ned)

    def linearizeInertial(self, x0, u0, dt):
        x0, u0 = _t(x0), _t(u0)
        Aw, Bw = torch.func.jacrev(self.inertialDynamics, argnums=(0, 1))(x0, u0)
        return torch.eye(12, dtype=Aw.dtype) + dt * Aw, dt * Bw
