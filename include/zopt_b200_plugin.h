/*
 * zopt_b200 -- C ABI of a USER-MODEL PLUG-IN library.
 *
 * The reference's iLQR / DDP solvers take an arbitrary Python callable `dynamics(x, u)` and differentiate it with JAX
 * (zopt/ilqrUtils.py:260-268, 330-338; zopt/pytrees.py:138-153, 179-194).  Here a model outside the registered ones is
 * defined symbolically (zopt_b200/plugin.py: sympy -> CUDA for f, f_x, f_u and sum_i lam_i d2f_i/dz2) and
 * zopt_b200/csrc/zb_user_model.cu -- the library's generic solver kernels compiled around that generated header -- is
 * built with nvcc for sm_100a into ONE shared library per model, exporting the entry points below.  They mirror
 * zb_ilqr_rollout / zb_ilqr_solve of zopt_b200.h minus the `zb_model` argument: same pointers, sizes, strides, stream,
 * error convention (0 OK, < 0 argument error, > 0 CUDA error; message through zb_user_last_error), caller-owned buffers,
 * no synchronisation.  Limits: n <= ZB_MAX_N, m <= ZB_MAX_M.
 *
 * User-defined COSTS (plugin.SymbolicCost: `runningCost(x, u)`, `terminalCost(x)` of zopt/ilqrUtils.py:260-268, differentiated by
 * JAX in the reference, pytrees.py:71-81, 99-115): when the generated header also defines ZB_USER_COST (c, (c_x, c_u), the
 * stacked Hessian, cf, cf_x, cf_xx), the same entry points evaluate THAT cost -- in the rollouts, and expanded and eigen-clamped
 * per time step in the backward pass (ilqrUtils.py:222-234) -- and the `zb_cost` argument is a placeholder (non-NULL, never read).
 */
#ifndef ZOPT_B200_PLUGIN_H
#define ZOPT_B200_PLUGIN_H

#include "zopt_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* dimensions the plug-in was generated for */
ZB_API int32_t zb_user_dims(int32_t* n, int32_t* m);
ZB_API int32_t zb_user_last_error(char* buf, size_t len);

/* x+ = f(x, u) at Bsz points: x (Bsz,n), u (Bsz,m) -> xn (Bsz,n); optionally f_x (Bsz,n,n), f_u (Bsz,n,m) (both or neither).
 * Replaces: calling the reference's `dynamics` callable / AffineDynamics.from_function (pytrees.py:138-147). */
ZB_API int32_t zb_user_step(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x, const void* u, void* xn,
                            void* fx, void* fu);

/* Z (Bsz,n+m,n+m) = sum_i lam_i d2 f_i / dz2 at (x, u), z = [x; u]: the costate contraction of QuadraticDynamics' f_xx, f_ux,
 * f_uu (pytrees.py:179-194; with lam = e_i it returns the slices themselves).  x (Bsz,n), u (Bsz,m), lam (Bsz,n). */
ZB_API int32_t zb_user_hess(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x, const void* u, const void* lam,
                            void* Z);

/* zopt/ilqrUtils.py:33-66 trajectoryRollout with the user model (arguments as zb_ilqr_rollout) */
ZB_API int32_t zb_user_rollout(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, const zb_cost* cost,
                               const void* x0, const void* l, const void* L, const void* xPrev, const void* uPrev, double alpha,
                               void* xTraj, void* uTraj, void* J_out);

/* zopt/ilqrUtils.py:260-327 iterativeLqr, :330-397 differentialDynamicProgramming (flags & ZB_SECOND_ORDER) with the user
 * model: arguments and outputs as zb_ilqr_solve.  DDP uses the model's full second-order block
 * [[v_x.f_xx, (v_x.f_ux)'], [v_x.f_ux, v_x.f_uu]], eigen-clamped per step (ilqrUtils.py:237-251). */
ZB_API size_t zb_user_ilqr_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N);
ZB_API int32_t zb_user_ilqr_solve(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t flags,
                                  const zb_cost* cost, const void* x0, const void* uGuess, int32_t maxIter, double tol,
                                  void* xTraj, void* uTraj, void* L_out, void* J_out, uint8_t* converged_out, int32_t* iters_out,
                                  int32_t* alpha_log, void* J_log, void* workspace, size_t workspace_bytes);

#ifdef __cplusplus
}
#endif
#endif
