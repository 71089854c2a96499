#!/usr/bin/env python
"""
Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference sources from
/root/reference (zopt.lqrUtils, zopt.ilqrUtils, zopt.pytrees, zopt.quadcopter) on seeded inputs.

`jax`/`jaxlib` are absent from this image, so the reference is imported on top of oracle/jax_shim (a minimal
NumPy-semantics stand-in for the JAX entry points it uses, backed by torch-CPU fp64; see oracle/jax_shim/README.md).
The formulas and control flow executed are the reference's own; the primitives are torch's.  `zopt.mpcUtils` needs
cvxpy and a QP solver and cannot be run at all (parity unpinned for that path).

Run in the build container only:  python scripts/gen_golden_from_reference.py
The GPU box has no /root/reference; tests read the committed .npz files.
"""
import os
import sys
import warnings

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle", "jax_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)
warnings.filterwarnings("ignore")

import torch  # noqa: E402
import jax  # noqa: E402  (the shim)
import jax.numpy as jnp  # noqa: E402
import zopt.ilqrUtils as ref_ilqr  # noqa: E402
import zopt.lqrUtils as ref_lqr  # noqa: E402
import zopt.pytrees as ref_pytrees  # noqa: E402
from zopt.quadcopter import Quadcopter  # noqa: E402

from zopt_b200 import configs  # noqa: E402

assert "jax_shim" in jax.__file__ and ref_lqr.__file__.startswith("/root/reference/")
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
T = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64))
npy = lambda t: t.detach().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
ac = Quadcopter()
uTrim = T(configs.U_TRIM)


def rep(a, N):
    return T(np.repeat(np.asarray(a)[None], N, axis=0))


# ---- (i) cfg 1a: the demo instance of discreteFiniteHorizonLqr (demos/discreteFiniteHorizonLqr.py:13-35) -------------
A, B = ac.linearize(T(np.zeros(8)), uTrim, dt=0.1)
N = 100
Qk, Rk = configs.cfg1_demo_weights(N)
K = ref_lqr.discreteFiniteHorizonLqr(rep(npy(A), N), rep(npy(B), N), T(Qk), T(Rk), N)
np.savez(os.path.join(OUT, "lqr_demo_n8.npz"), A=npy(A), B=npy(B), Qk=Qk, Rk=Rk, N=N, K=npy(K))
print("lqr_demo_n8", K.shape)

# ---- (ii) cfg 2: 32 random quadcopter LQR problems, n=12 m=4 N=50 ----------------------------------------------------
d = configs.cfg2(Bsz=32)
N = d["N"]
As, Bs, Ls = [], [], []
for i in range(32):
    Aw, Bw = jax.jacobian(ac.inertialDynamics, argnums=(0, 1))(T(d["xbar"][i]), T(d["ubar"][i]))  # demos/lqrMpc.py:26-28
    Ai, Bi = np.eye(12) + 0.1 * npy(Aw), 0.1 * npy(Bw)
    Qi, Ri = np.diag(d["qdiag"][i]), np.diag(d["rdiag"][i])
    Qk = np.repeat(Qi[None], N + 1, axis=0)
    Qk[N] *= 10
    L = ref_lqr.discreteFiniteHorizonLqr(rep(Ai, N), rep(Bi, N), T(Qk), rep(Ri, N), N)
    As.append(Ai), Bs.append(Bi), Ls.append(npy(L))
np.savez(os.path.join(OUT, "lqr_cfg2_32.npz"), xbar=d["xbar"], ubar=d["ubar"], qdiag=d["qdiag"], rdiag=d["rdiag"],
         A=np.array(As), B=np.array(Bs), L=np.array(Ls), N=N)
print("lqr_cfg2_32", np.array(Ls).shape)

# ---- (iii) bilinearAffineLqr on the demo's problem with a seeded H (demos/bilinearLqrControl.py:21-43) ---------------
rng = np.random.default_rng(1234 + 10)
n, m, N = 8, 4, 100
Hm = 0.2 * rng.normal(size=(N, m, n))
dv = 0.01 * rng.normal(size=(N, n))
qv = 0.1 * np.repeat(np.array([1., -1, 0, 0, 0, 0, 0, 0])[None], N, axis=0)
rv = 0.05 * rng.normal(size=(N, m))
q0 = rng.normal(size=N)
Lb, lb = ref_lqr.bilinearAffineLqr(rep(npy(A), N), rep(npy(B), N), T(dv), rep(np.eye(n), N), rep(np.eye(m), N), T(Hm), T(qv),
                                   T(rv), T(q0), N)
np.savez(os.path.join(OUT, "bilinear_demo.npz"), A=npy(A), B=npy(B), d=dv, H=Hm, q=qv, r=rv, q0=q0, N=N, L=npy(Lb), l=npy(lb))
print("bilinear_demo", Lb.shape, lb.shape)

# ---- (iv) quadcopter dynamics + autodiff expansions at random points (incl. wind) -------------------------------------
rng = np.random.default_rng(1234 + 11)
xs = configs.quad_states(rng, 16)
us = np.tile(configs.U_TRIM, (16, 1)) + rng.normal(size=(16, 4))
wind = np.array([3.0, 1.0, 0.0])  # demos/iterativeLqr.py:47
F0 = np.array([npy(ac.inertialDynamics(T(x), T(u))) for x, u in zip(xs, us)])
Fw = np.array([npy(ac.inertialDynamics(T(x), T(u), wind_ned=T(wind))) for x, u in zip(xs, us)])
Jx, Ju, Hxx = [], [], []
for x, u in zip(xs, us):
    jx, ju = jax.jacobian(ac.inertialDynamics, argnums=(0, 1))(T(x), T(u))
    hx = jax.hessian(ac.inertialDynamics, 0)(T(x), T(u))
    Jx.append(npy(jx)), Ju.append(npy(ju)), Hxx.append(npy(hx))
np.savez(os.path.join(OUT, "quadcopter_points.npz"), x=xs, u=us, wind=wind, F=F0, F_wind=Fw, Jx=np.array(Jx), Ju=np.array(Ju),
         Hxx=np.array(Hxx))
print("quadcopter_points")

# ---- (v) iLQR / DDP on the demos' problems (demos/iterativeLqr.py:22-39, demos/differentialDynamicProgramming.py:22-39) --
dt = 0.1
dynFun = lambda x, u: x + dt * ac.inertialDynamics(x, u)


def run_solver(solver, x0, N, R, maxIter):
    Q = np.eye(12)
    Qt, Rt = T(Q), T(R)
    costFun = lambda x, u: x.T @ Qt @ x + u.T @ Rt @ u
    terminalCostFun = lambda x: 10 * x @ Qt @ x
    uGuess = T(np.repeat(configs.U_TRIM[None], N, axis=0))
    # per-iteration record by running the reference solver with maxIter = 1..k (it is deterministic)
    Js, trajs = [], None
    for k in range(0, maxIter + 1):
        traj, LArr, J, conv = solver(dynFun, costFun, terminalCostFun, T(x0), uGuess, maxIter=k, tol=-1.0)
        Js.append(float(J))
    return dict(x0=x0, N=N, R=R, J_per_iter=np.array(Js), xTraj=npy(traj.xTraj), uTraj=npy(traj.uTraj), L=npy(LArr),
                converged=bool(conv), maxIter=maxIter)


x0 = np.zeros(12)
x0[9:12] = [10, 10, 10]
g = run_solver(ref_ilqr.iterativeLqr, x0, 40, np.eye(4), 4)
np.savez(os.path.join(OUT, "ilqr_demo_N40_it4.npz"), **g)
print("ilqr", g["J_per_iter"])
x0 = np.zeros(12)
x0[9:12] = [0, 5, 0]
g = run_solver(ref_ilqr.differentialDynamicProgramming, x0, 40, 0.2 * np.eye(4), 3)
np.savez(os.path.join(OUT, "ddp_demo_N40_it3.npz"), **g)
print("ddp", g["J_per_iter"])

# ---- (vi) single Riccati steps on random pytrees (riccatiStep_ilqr / riccatiStep_ddp / ensurePositiveDefinite) -------
rng = np.random.default_rng(1234 + 12)
n, m = 5, 3
spd = lambda k: (lambda M: M @ M.T + np.eye(k))(rng.normal(size=(k, k)))
f_x, f_u = rng.normal(size=(n, n)) * 0.5, rng.normal(size=(n, m))
sym = lambda t: t + np.swapaxes(t, -1, -2)
f_xx, f_ux, f_uu = sym(rng.normal(size=(n, n, n)) * 0.1), rng.normal(size=(n, m, n)) * 0.1, sym(rng.normal(size=(n, m, m)) * 0.1)
czz = spd(n + m)
c, c_x, c_u = rng.normal(), rng.normal(size=n), rng.normal(size=m)
v, v_x, v_xx = rng.normal(), rng.normal(size=n), spd(n)
cost = ref_pytrees.QuadraticCostFunction(T(c), T(c_x), T(c_u), T(czz[:n, :n]), T(czz[n:, :n]), T(czz[n:, n:]))
val = ref_pytrees.QuadraticValueFunction(T(v), T(v_x), T(v_xx))
vo1, p1 = ref_ilqr.riccatiStep_ilqr(ref_pytrees.AffineDynamics(T(np.zeros(n)), T(f_x), T(f_u)), cost, val)
vo2, p2 = ref_ilqr.riccatiStep_ddp(ref_pytrees.QuadraticDynamics(T(np.zeros(n)), T(f_x), T(f_u), T(f_xx), T(f_ux), T(f_uu)), cost, val)
S = sym(rng.normal(size=(16, 16)))
np.savez(os.path.join(OUT, "riccati_steps.npz"), f_x=f_x, f_u=f_u, f_xx=f_xx, f_ux=f_ux, f_uu=f_uu, czz=czz, c=c, c_x=c_x, c_u=c_u,
         v=v, v_x=v_x, v_xx=v_xx, ilqr_v=npy(vo1.v), ilqr_vx=npy(vo1.v_x), ilqr_vxx=npy(vo1.v_xx), ilqr_l=npy(p1.l), ilqr_L=npy(p1.L),
         ddp_v=npy(vo2.v), ddp_vx=npy(vo2.v_x), ddp_vxx=npy(vo2.v_xx), ddp_l=npy(p2.l), ddp_L=npy(p2.L), S=S,
         S_pd=npy(ref_ilqr.ensurePositiveDefinite(T(S))))
print("riccati_steps")

# ---- (vii) the reference's Simulator (discrete branch, zopt/simulator.py:124-169) on the demos' two closed loops -------
import zopt.simulator as ref_sim  # noqa: E402
assert ref_sim.__file__.startswith("/root/reference/")
# (a) demos/iterativeLqr.py:44-56: tracking controller u = L_k (x - xTraj_k) + uTraj_k on the plant in wind [3,1,0]; the
#     plan is the frozen iLQR demo solution above (N = 40)
gi = np.load(os.path.join(OUT, "ilqr_demo_N40_it4.npz"))
Np = int(gi["N"])
xTj, uTj, LTj = T(gi["xTraj"]), T(gi["uTraj"]), T(gi["L"])
wind = np.array([3.0, 1.0, 0.0])
x0s = np.array(gi["x0"], dtype=np.float64)
noisyDynFun = lambda k, x, u: (None, T(x) + dt * ac.inertialDynamics(T(x), T(u), wind_ned=T(wind)))
dynamicsBlock = ref_sim.SimBlock(noisyDynFun, T(x0s), dt=dt, name="Dynamics")
controllerBlock = ref_sim.SimBlock(lambda k, xCtrl, x: (LTj[k] @ (T(x) - xTj[k]) + uTj[k], T(np.array([]))), T(np.array([])), dt=dt,
                                   name="Controller")
tS, xc, xS, uS, _ = ref_sim.Simulator([controllerBlock, dynamicsBlock], (0, Np * dt)).simulate()
sim_a = dict(a_t=npy(tS), a_x=npy(xS), a_u=npy(uS), a_x0=x0s, a_wind=wind, a_xTraj=npy(xTj), a_uTraj=npy(uTj), a_L=npy(LTj), a_dt=dt)
print("simulator (tracking, wind)", npy(xS).shape, npy(uS).shape)
# (b) demos/discreteFiniteHorizonLqr.py:38-49: 8-state LQR gains on the 12-state plant through x[:8]
gl = np.load(os.path.join(OUT, "lqr_demo_n8.npz"))
Kd = T(gl["K"])
xTrim8 = T(np.zeros(8))
x0b = np.zeros(12)
x0b[0:3] = 1
dynB = ref_sim.SimBlock(lambda k, x, u: (None, T(x) + dt * ac.inertialDynamics(T(x), T(u))), T(x0b), dt=dt, name="Dynamics")
ctlB = ref_sim.SimBlock(lambda k, xCtrl, x: ref_lqr.proportionalFeedbackController(T(x)[:8], xTrim8, uTrim, Kd[k]), T(np.array([])), dt=dt,
                        name="Controller", jittable=False)
tB, _, xB, uB, _ = ref_sim.Simulator([ctlB, dynB], (0, 10)).simulate()
np.savez(os.path.join(OUT, "simulator_demos.npz"), b_t=npy(tB), b_x=npy(xB), b_u=npy(uB), b_x0=x0b, b_K=npy(Kd), b_uTrim=npy(uTrim), **sim_a)
print("simulator (lqr gains)", npy(xB).shape, npy(uB).shape)


# ---- (ix) continuous-time finiteHorizonLqr (zopt/lqrUtils.py:55-98) on a time-varying problem (jax odeint -> SciPy RK45 at the
#      same tolerances in the shim): the value matrices on the output grid and the gains at off-grid times -------------------
rng = np.random.default_rng(1234 + 13)
nc, mc, Tc, Nc = 4, 2, 1.5, 12
A0, A1 = 0.5 * rng.normal(size=(nc, nc)), 0.3 * rng.normal(size=(nc, nc))
B0, B1 = rng.normal(size=(nc, mc)), 0.2 * rng.normal(size=(nc, mc))
Qc, Ric, Qfc = np.diag(rng.uniform(0.5, 2, nc)), np.diag(1.0 / rng.uniform(0.5, 2, mc)), 2.0 * np.eye(nc)
Afun = lambda t: T(A0 + np.sin(2.0 * float(t)) * A1)
Bfun = lambda t: T(B0 + float(t) * B1)
Qfun = lambda t: T((1.0 + 0.5 * float(t)) * Qc)
Rifun = lambda t: T(Ric)
Kc = ref_lqr.finiteHorizonLqr(Afun, Bfun, Qfun, Rifun, T(Qfc), Tc, N=Nc)
tq = np.array([0.0, 0.1, 0.37, 0.75, 1.2, 1.5, 2.0])
np.savez(os.path.join(OUT, "care_tv.npz"), A0=A0, A1=A1, B0=B0, B1=B1, Q=Qc, R_inv=Ric, Qf=Qfc, T=Tc, N=Nc, tq=tq,
         K=np.array([npy(Kc(t)) for t in tq]))
print("care_tv", np.array([npy(Kc(t)) for t in tq]).shape)

# ---- (viii) iLQR / DDP at BASELINE cfg 4 / cfg 5 size: N = 200 x 10 iterations (iLQR), N = 100 x 10 iterations (DDP), four
#      problems each from the configs' own initial-state distributions; per-iteration J, step-size index, final x, u, L ------


def run_solver_logged(solver, x0, N, R, maxIter):
    """The reference solver run with maxIter = 0..K (deterministic): J after every iteration; the step-size index of
    iteration k is recovered from the reference's own forwardPass2 candidates (argmin of the 16 costs, ilqrUtils.py:145-149)
    by re-running the solver's iteration body on the previous iterate -- here simply by matching J_k against the 16
    candidate costs computed with the reference's trajectoryRollout + CostFunction."""
    Q = np.eye(12)
    Qt, Rt = T(Q), T(R)
    costFun = lambda x, u: x.T @ Qt @ x + u.T @ Rt @ u
    terminalCostFun = lambda x: 10 * x @ Qt @ x
    uGuess = T(np.repeat(configs.U_TRIM[None], N, axis=0))
    Js = []
    for k in range(0, maxIter + 1):
        traj, LArr, J, conv = solver(dynFun, costFun, terminalCostFun, T(x0), uGuess, maxIter=k, tol=-1.0)
        Js.append(float(J))
    return dict(J_per_iter=np.array(Js), xTraj=npy(traj.xTraj), uTraj=npy(traj.uTraj), L=npy(LArr))


if os.environ.get("GOLDEN_FULL_SIZE", "1") == "1":
    for name, solver, cfg, N, nprob in (("ilqr_cfg4_N200_it10", ref_ilqr.iterativeLqr, configs.cfg4, 200, 4),
                                        ("ddp_cfg5_N100_it10", ref_ilqr.differentialDynamicProgramming, configs.cfg5, 100, 4)):
        dd = cfg(Bsz=nprob, N=N)
        recs = [run_solver_logged(solver, dd["x0"][i], N, dd["R"], 10) for i in range(nprob)]
        # step-size indices: the reference does not return them; they are taken from the oracle restatement AFTER checking that
        # its cost after every iteration equals the reference's (so every argmin of ilqrUtils.py:147-149 fell on the same index)
        from oracle import ilqr as oilqr
        from oracle.quadcopter import Quadcopter as OQ
        Qt_, Rt_ = T(np.eye(12)), T(dd["R"])
        osolver = oilqr.differentialDynamicProgramming if "ddp" in name else oilqr.iterativeLqr
        alphas = []
        for i in range(nprob):
            olog = []
            osolver(OQ().eulerStep(dt), lambda x, u: x @ Qt_ @ x + u @ Rt_ @ u, lambda x: 10 * x @ Qt_ @ x, T(dd["x0"][i]),
                    T(np.repeat(configs.U_TRIM[None], N, axis=0)), maxIter=10, tol=-1.0, log=olog)
            Jo = np.array([e["J"] for e in olog])
            assert np.max(np.abs(Jo - recs[i]["J_per_iter"]) / recs[i]["J_per_iter"]) < 1e-11, (name, i)
            alphas.append([e["alpha_idx"] for e in olog[1:]])
        np.savez_compressed(os.path.join(OUT, name + ".npz"), x0=dd["x0"], R=dd["R"], N=N, maxIter=10, alpha_idx=np.array(alphas),
                            J_per_iter=np.array([r["J_per_iter"] for r in recs]), xTraj=np.array([r["xTraj"] for r in recs]),
                            uTraj=np.array([r["uTraj"] for r in recs]), L=np.array([r["L"] for r in recs]))
        print(name, np.array([r["J_per_iter"] for r in recs])[:, [0, 1, -1]])
print("golden fixtures written to", OUT)
