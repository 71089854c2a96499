/*
 * zopt_b200 -- C ABI of the B200-native LQR / iLQR / DDP / LQR-MPC solver.
 *
 * This is the drop-in boundary for the hot path of zprihoda/zopt.  The reference
 * is pure Python (JAX); it has no FFI of its own, so each entry point below names
 * the reference *function* it replaces (file:line under the reference tree).  The
 * Python mirror in zopt_b200/*.py binds these with ctypes.
 *
 * Conventions
 *   - Every function returns int32: 0 = OK, <0 = argument error, >0 = CUDA error
 *     code.  zb_last_error() returns the thread-local message of the last failure.
 *     Numerical non-success (converged=false, NaN) is data, not an error.
 *   - All pointers are DEVICE pointers on `device`; work is enqueued on `stream`
 *     (a cudaStream_t) and the call returns without synchronising.  The caller
 *     owns every buffer; the library keeps no pointer after return.
 *   - dtype: ZB_F32 or ZB_F64; every floating-point buffer of a call has that type.
 *   - Dense matrices are row-major and contiguous; batches / time series are
 *     described by element strides (zb_arr), so a time-invariant or batch-shared
 *     operand is passed with stride 0 and never materialised.
 *   - Public layouts follow the reference index order with one leading batch
 *     axis: L is (Bsz,N,m,n), xTraj (Bsz,N+1,n), uTraj (Bsz,N,m).
 */
#ifndef ZOPT_B200_H
#define ZOPT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define ZB_API __attribute__((visibility("default")))
#else
#define ZB_API
#endif

#define ZB_F32 0
#define ZB_F64 1

#define ZB_MAX_N 16 /* generic kernels: state dimension limit */
#define ZB_MAX_M 8  /* generic kernels: control dimension limit */

/* strided view of a batch (and optionally a time series) of dense row-major blocks */
typedef struct zb_arr {
    const void* ptr;  /* device pointer to block [b=0][k=0] */
    int64_t stride_b; /* elements between consecutive problems; 0 = shared by the whole batch */
    int64_t stride_t; /* elements between consecutive time steps; 0 = time-invariant */
} zb_arr;

/* dynamics model of the iLQR/DDP/rollout entry points (the reference takes a Python callable,
 * ilqrUtils.py:260-268; a kernel cannot trace a lambda, so models are registered kinds) */
#define ZB_MODEL_LINEAR 0    /* x+ = A x + B u            (tests/test_ilqrUtils.py:167-196) */
#define ZB_MODEL_QUADCOPTER 1 /* x+ = x + dt*inertialDynamics(x,u,wind)  (zopt/quadcopter.py:116-144, demos/iterativeLqr.py:35) */

typedef struct zb_model {
    int32_t kind;  /* ZB_MODEL_* */
    int32_t n, m;  /* state / control dimension (quadcopter: 12, 4) */
    int32_t has_wind;
    double dt;         /* quadcopter: Euler step */
    double wind[3];    /* quadcopter: wind_ned */
    zb_arr A, B;       /* linear: (n,n) and (n,m) blocks, stride_t ignored */
} zb_model;

/* cost of the iLQR/DDP entry points: running x'Qx + u'Ru, terminal x'Qf x (demos/iterativeLqr.py:12-13,36-37) */
typedef struct zb_cost {
    zb_arr Q, R, Qf; /* (n,n), (m,m), (n,n) blocks, stride_t ignored */
} zb_cost;

ZB_API int32_t zb_version(void);
ZB_API int32_t zb_last_error(char* buf, size_t len);
/* number of SMs / device name probe (used by bench.py to size batches) */
ZB_API int32_t zb_device_info(int32_t device, int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor, size_t* total_mem);

/* ---- zopt/lqrUtils.py:144-173  discreteFiniteHorizonLqr(A,B,Q,R,N) -> L -------------------------
 * A,B,Q,R are time series with T >= N rows; terminal value is row T-1 of Q (lqrUtils.py:172).
 * L_out (Bsz,N,m,n); V0_out (Bsz,n,n) optional (NULL to skip): value matrix after the last step. */
ZB_API int32_t zb_lqr_dfh(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t T, int32_t n,
                   int32_t m, const zb_arr* A, const zb_arr* B, const zb_arr* Q, const zb_arr* R, void* L_out,
                   void* V0_out);

/* Same with flags: ZB_FORCE_GENERIC = take the as-written kernels, which use Q and R exactly as given (the (12,4) and the
 * (8,4) kernels read only the lower triangle of the symmetric weights; the Python mirror passes this flag when a weight is not
 * symmetric, so that results equal the reference's for any input, lqrUtils.py:168-169). */
ZB_API int32_t zb_lqr_dfh_flags(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t T, int32_t n,
                         int32_t m, const zb_arr* A, const zb_arr* B, const zb_arr* Q, const zb_arr* R, int32_t flags,
                         void* L_out, void* V0_out);

/* ---- zopt/lqrUtils.py:39-98  finiteHorizonLqr (continuous time): the Riccati differential equation of the LQR HJB,
 *   -dV/dt = Q + V A + A'V - V B R^-1 B'V,  V(T) = Qf,  integrated backward with RK4, `substeps` steps between two of the N
 * output points t_i = i T/(N-1) (the reference: jax odeint, adaptive Dormand-Prince at 1.4e-8, same output grid).
 * A (n,n), B (n,m), Q (n,n), Rinv (m,m) are time series SAMPLED at the scheme's stage times: 2 (N-1) substeps + 1 samples,
 * sample j at time T - j h/2, h = T / ((N-1) substeps); stride_t = 0 for constant coefficients.  Qf (n,n) blocks.
 * V_out (Bsz,N,n,n): V_out[i] = V(t_i).  The gains K(t) = R^-1(t) B(t)' V(t) are formed by the caller (lqrUtils.py:95-97). */
ZB_API int32_t zb_lqr_care_rk4(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t substeps, int32_t n, int32_t m,
                        double T_horizon, const zb_arr* A, const zb_arr* B, const zb_arr* Q, const zb_arr* Rinv, const zb_arr* Qf,
                        void* V_out);

/* ---- zopt/lqrUtils.py:207-262  bilinearAffineLqr(A,B,d,Q,R,H,q,r,q0,N) -> (L,l) ------------------
 * L_out (Bsz,N,m,n), l_out (Bsz,N,m). */
ZB_API int32_t zb_lqr_bilinear(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t T, int32_t n,
                        int32_t m, const zb_arr* A, const zb_arr* B, const zb_arr* d, const zb_arr* Q,
                        const zb_arr* R, const zb_arr* H, const zb_arr* q, const zb_arr* r, const zb_arr* q0,
                        void* L_out, void* l_out);

/* Same with flags: ZB_FORCE_GENERIC = Q, R used exactly as given (the (8,4) kernels read their lower triangles only;
 * zb_lqr_bilinear = flags 0). */
ZB_API int32_t zb_lqr_bilinear_flags(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t T, int32_t n,
                              int32_t m, const zb_arr* A, const zb_arr* B, const zb_arr* d, const zb_arr* Q,
                              const zb_arr* R, const zb_arr* H, const zb_arr* q, const zb_arr* r, const zb_arr* q0,
                              int32_t flags, void* L_out, void* l_out);

/* ---- zopt/quadcopter.py:116-144 inertialDynamics, :179-201 linearize (12-state form used by
 * demos/lqrMpc.py:26-28 and demos/iterativeLqr.py:35) -------------------------------------------
 * x (Bsz,12), u (Bsz,4) contiguous.  xdot_out (Bsz,12) = F(x,u,wind).
 * linearize: A_out (Bsz,12,12) = I*(dt!=0) + (dt?dt:1)*dF/dx ; B_out (Bsz,12,4) = (dt?dt:1)*dF/du.
 * hess: H_out (Bsz,12,12) = sum_i lam_i d2F_i/dx2 (times dt if dt != 0). */
ZB_API int32_t zb_quad_dynamics(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x, const void* u,
                         const double* wind_ned /* host, 3 values or NULL */, void* xdot_out);
ZB_API int32_t zb_quad_linearize(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x, const void* u,
                          const double* wind_ned, double dt, void* A_out, void* B_out);
ZB_API int32_t zb_quad_hess_contract(int32_t dtype, int32_t device, void* stream, int64_t Bsz, const void* x,
                              const void* u, const double* wind_ned, double dt, const void* lam, void* H_out);

/* ---- zopt/ilqrUtils.py:33-66 trajectoryRollout + pytrees.py:215-220 AffinePolicy.__call__ --------
 * u_k = alpha*l_k + L_k (x_k - xPrev_k) + uPrev_k ; x_{k+1} = f(x_k,u_k).
 * x0 (Bsz,n); l (Bsz,N,m); L (Bsz,N,m,n); xPrev (Bsz,N+1,n); uPrev (Bsz,N,m); alpha host scalar.
 * Outputs xTraj (Bsz,N+1,n), uTraj (Bsz,N,m); J_out optional (Bsz): cost of the rollout under `cost`
 * (pytrees.py:40-55), cost may be NULL when J_out is NULL. */
ZB_API int32_t zb_ilqr_rollout(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, const zb_model* model,
                        const zb_cost* cost, const void* x0, const void* l, const void* L, const void* xPrev,
                        const void* uPrev, double alpha, void* xTraj, void* uTraj, void* J_out);

/* ---- zopt/ilqrUtils.py:116-150 forwardPass2: 16 rollouts at alpha = 0.5^j, argmin of the cost ----
 * J_out (Bsz), alpha_idx_out (Bsz) int32 (optional), Jall_out (Bsz,16) optional. */
ZB_API int32_t zb_ilqr_forward_pass(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N,
                             const zb_model* model, const zb_cost* cost, const void* x0, const void* l,
                             const void* L, const void* xPrev, const void* uPrev, void* xTraj, void* uTraj,
                             void* J_out, int32_t* alpha_idx_out, void* Jall_out);

/* ---- zopt/ilqrUtils.py:153-181 riccatiStep_ilqr/backwardPass_ilqr, :184-214 _ddp ------------------
 * Explicit-pytree form (stacked AffineDynamics / QuadraticCostFunction / QuadraticValueFunction leaves,
 * pytrees.py:84-204).  second_order != 0 adds the eigen-clamped v_x.f_zz block (ilqrUtils.py:237-251);
 * f_xx (n,n,n), f_ux (n,m,n), f_uu (n,m,m) blocks then required.
 * Outputs: l_out (Bsz,N,m), L_out (Bsz,N,m,n); optional value function after the last step:
 * v_out (Bsz), vx_out (Bsz,n), vxx_out (Bsz,n,n). */
ZB_API int32_t zb_ilqr_backward(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t n, int32_t m,
                         int32_t second_order, const zb_arr* f_x, const zb_arr* f_u, const zb_arr* f_xx,
                         const zb_arr* f_ux, const zb_arr* f_uu, const zb_arr* c, const zb_arr* c_x,
                         const zb_arr* c_u, const zb_arr* c_xx, const zb_arr* c_ux, const zb_arr* c_uu,
                         const zb_arr* v, const zb_arr* v_x, const zb_arr* v_xx, void* l_out, void* L_out,
                         void* v_out, void* vx_out, void* vxx_out);

/* ---- zopt/ilqrUtils.py:217-219 ensurePositiveDefinite: V max(eig,eps) V' of symmetric (p,p) blocks,
 * p <= ZB_MAX_N + ZB_MAX_M.  in/out (Bsz,p,p) contiguous (may alias). */
ZB_API int32_t zb_pd_clamp(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t p, double eps, const void* in,
                    void* out);

/* ---- zopt/ilqrUtils.py:260-327 iterativeLqr, :330-397 differentialDynamicProgramming -------------
 * Whole solve for a registered model/cost: initial rollout, then up to maxIter iterations of
 * {linearise along the trajectory, condition, backward pass, 16-way forward pass}; per-problem
 * convergence |J - J_new| <= tol freezes that problem (ilqrUtils.py:301-303,318).
 * x0 (Bsz,n), uGuess (Bsz,N,m).  Outputs xTraj (Bsz,N+1,n), uTraj (Bsz,N,m), L_out (Bsz,N,m,n),
 * J_out (Bsz), converged_out (Bsz) uint8, iters_out (Bsz) int32,
 * alpha_log (Bsz,maxIter) int32 optional (-1 = iteration not run), J_log (Bsz,maxIter+1) optional.
 * flags: ZB_SECOND_ORDER = DDP (ilqrUtils.py:330-397) instead of iLQR; ZB_COST_DIAGONAL = the caller asserts Q, R, Qf diagonal
 * (the quadcopter backward kernel then keeps twice as many problems per SM and the per-solve setup is closed-form).  The 16-way
 * line search + commit of a quadcopter solve runs as ONE fused launch (csrc/ilqr_forward.cuh); ZB_GENERIC_FORWARD forces the
 * two-kernel path.
 * workspace: device scratch of zb_ilqr_workspace_bytes(...) bytes. */
ZB_API size_t zb_ilqr_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N, int32_t n, int32_t m);
ZB_API int32_t zb_ilqr_solve(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t flags,
                      const zb_model* model, const zb_cost* cost, const void* x0, const void* uGuess,
                      int32_t maxIter, double tol, void* xTraj, void* uTraj, void* L_out, void* J_out,
                      uint8_t* converged_out, int32_t* iters_out, int32_t* alpha_log, void* J_log,
                      void* workspace, size_t workspace_bytes);

/* ---- zopt/mpcUtils.py:12-81 lqrMpc(A,B,Q,R,N,x_lb,x_ub,u_lb,u_ub,Qf).solve(x0) --------------------
 * min sum_{k<N} x'Qx + u'Ru + x_N'Qf x_N  s.t. x+ = Ax + Bu, x_lb<=x_k<=x_ub (k=0..N), u_lb<=u_k<=u_ub.
 * Bounds may be +-inf.  When no bound can bind the solve is the exact Riccati sweep + rollout;
 * otherwise ADMM (over-relaxed, residual-balanced rho, as OSQP) with the dynamics kept exact: the linear solve of
 * every iteration is a vector Riccati sweep + rollout against gains factored once per rho.
 * flags: ZB_MPC_BOUNDED = at least one bound is finite (ADMM path); ZB_COST_DIAGONAL = the caller asserts that Q, R
 * and Qf are diagonal (the fp32 (12,4) kernel then keeps 4 instead of 27 float4 of cost data per problem on chip).
 * status_out (Bsz) int8: 0 optimal, 1 optimal_inaccurate (max_iter hit), 2 infeasible.
 * iters_out (Bsz) int32 ADMM iterations (0 on the unconstrained path). */
#define ZB_MPC_BOUNDED 1   /* zb_mpc_lqr_solve */
#define ZB_SECOND_ORDER 1  /* zb_ilqr_solve */
#define ZB_COST_DIAGONAL 2 /* both */
#define ZB_VARIANT_THREAD 4 /* zb_mpc_closed_loop_quad, zb_mpc_box_*: force the thread-per-problem kernel */
#define ZB_VARIANT_QUAD 8   /* ... force the 4-threads-per-problem kernel (default for small batches) */
#define ZB_VARIANT_WARP 64  /* zb_mpc_closed_loop_quad (fp32): force the nine-lanes-per-problem register-tiled kernel (default for the smallest batches) */
#define ZB_FORCE_GENERIC 128 /* zb_lqr_dfh_flags, zb_mpc_lqr_solve: use the generic kernels (weights taken as given, not assumed symmetric) */
#define ZB_TV_BULK_COPY 256 /* zb_lqr_dfh_flags, time-varying fp32 (12,4): operands by cp.async.bulk + mbarrier (TMA engine) instead of per-lane cp.async */
#define ZB_MPC_SPLIT_ROLLOUT 512 /* zb_mpc_lqr_solve, fp32 (12,4), diagonal costs, bounds inactive, multi-wave batches: sweep and plan rollout as two concurrent kernels (rollout on a side stream, following the sweep group by group through flags) instead of the fused kernel; measured equal (0.88 ms per 65,536 solves), kept for experiments */
#define ZB_GENERIC_FORWARD 32 /* zb_ilqr_solve: force the two-kernel line search (k_forward_costs + k_forward_commit) instead of the fused quadcopter kernel */
#define ZB_BOX_STATE_GLOBAL 16 /* zb_mpc_box_*: keep the 4-threads-per-problem kernel's ADMM state in the global workspace even when it would fit on chip */

typedef struct zb_admm_opts {
    int32_t max_iter;   /* default 4000 */
    int32_t check_every; /* default 25 */
    double rho, sigma, alpha; /* 0.1, 1e-6, 1.6 */
    double eps_abs, eps_rel;   /* 1e-3, 1e-3 */
    double eps_prim_inf;       /* 1e-4: primal-infeasibility certificate tolerance */
} zb_admm_opts;
ZB_API size_t zb_mpc_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N, int32_t n, int32_t m);
ZB_API int32_t zb_mpc_lqr_solve(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t n, int32_t m,
                         const zb_arr* A, const zb_arr* B, const zb_arr* Q, const zb_arr* R, const zb_arr* Qf,
                         const zb_arr* x_lb, const zb_arr* x_ub, const zb_arr* u_lb, const zb_arr* u_ub,
                         int32_t flags, const void* x0, const zb_admm_opts* opts, void* u0_out, void* xTraj,
                         void* uTraj, int8_t* status_out, int32_t* iters_out, void* workspace,
                         size_t workspace_bytes);

/* ---- lqrMpc with finite bounds, ONE problem definition for the whole batch, (n,m) = (12,4) ---------------------------
 * The reference object `lqrMpc(A,B,Q,R,N,x_lb,x_ub,u_lb,u_ub,Qf)` builds the QP once (zopt/mpcUtils.py:14-59) and
 * `solve(x0)` re-solves it per initial state (mpcUtils.py:61-81; receding-horizon loop demos/lqrMpc.py:42-47).  Here the
 * problem definition is passed on the HOST (row-major fp64, converted to `dtype`; bounds may be +-inf) and travels to the
 * kernel by value; only x0 and the outputs are device buffers (Bsz leading).  Same ADMM as zb_mpc_lqr_solve's bounded path
 * with rho kept on the grid rho0 * 2^j, j = -12..12, whose LQ gains are tabulated once per problem definition:
 *   zb_mpc_box_build_tables -> `tables` (device, zb_mpc_box_tables_bytes), valid for (A,B,Q,R,Qf,N,rho0,dtype);
 *   zb_mpc_box_solve        -> any number of solves against them; opts->rho must equal the rho0 the tables were built with.
 * Outputs and status codes as zb_mpc_lqr_solve. */
ZB_API size_t zb_mpc_box_tables_bytes(int32_t dtype, int32_t N);
ZB_API size_t zb_mpc_box_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N);
ZB_API int32_t zb_mpc_box_build_tables(int32_t dtype, int32_t device, void* stream, int32_t N, const double* A /* host 12x12 */,
                                const double* B /* host 12x4 */, const double* Q, const double* R, const double* Qf,
                                double rho0, void* tables, size_t tables_bytes);
ZB_API int32_t zb_mpc_box_solve(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, const double* A, const double* B,
                         const double* x_lb, const double* x_ub, const double* u_lb, const double* u_ub /* host */,
                         const void* tables, const void* x0, const zb_admm_opts* opts,
                         int32_t flags /* 0 = pick by batch size | ZB_VARIANT_THREAD | ZB_VARIANT_QUAD (fp32 only) */, void* u0_out,
                         void* xTraj, void* uTraj, int8_t* status_out, int32_t* iters_out, void* workspace, size_t workspace_bytes);

/* The receding-horizon loop of demos/lqrMpc.py:42-47 around that problem, fused into one kernel: every simulation step clips
 * the state into [x_lb + clip_margin, x_ub - clip_margin] (demo: 1e-6; the QP constrains x_0 too), solves from it warm-started
 * with the previous step's ADMM state (cvxpy re-solves warm), applies the first move and continues from the plan's own next
 * state ("assume perfect tracking").  xSim_out (Bsz,Tsim+1,12): the states the solves started from + the final state;
 * uSim_out (Bsz,Tsim,4); status_out (Bsz): worst status over the steps (an infeasible step ends that problem's loop, NaN
 * afterwards); iters_out (Bsz): total ADMM iterations. */
ZB_API size_t zb_mpc_box_closed_loop_workspace_bytes(int32_t dtype, int64_t Bsz, int32_t N);
ZB_API int32_t zb_mpc_box_closed_loop(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t Tsim, const double* A,
                               const double* B, const double* x_lb, const double* x_ub, const double* u_lb, const double* u_ub,
                               const void* tables, const void* x0, const zb_admm_opts* opts, int32_t flags /* as zb_mpc_box_solve */,
                               double clip_margin, void* xSim_out, void* uSim_out, int8_t* status_out, int32_t* iters_out,
                               void* workspace, size_t workspace_bytes);

/* ---- closed-loop LQR-MPC of the quadcopter (BASELINE cfg 3; the receding-horizon loop of demos/lqrMpc.py:42-47 with a
 * nonlinear plant), fp32 or fp64 (the fp64 kernel is the cooperative one of csrc/lqr_quad64.cuh; the ZB_VARIANT_* flags select
 * between the two fp32 kernels), bounds inactive.  Per simulation step t: A_t = I + dt dF/dx(x_t,u_trim), B = dt dF/du, a full
 * Riccati sweep of horizon N from Qf (zopt/mpcUtils.py:47-59 with infinite bounds), u_t = first move of the plan,
 * x_{t+1} = x_t + dt F(x_t, u_trim + u_t) (zopt/quadcopter.py:116-144).  One fused kernel; only the trajectory is written:
 * xSim_out (Bsz,Tsim+1,12), uSim_out (Bsz,Tsim,4) (deviation from u_trim).  Q,R,Qf: (12,12),(4,4),(12,12) blocks (symmetric). */
ZB_API int32_t zb_mpc_closed_loop_quad(int32_t dtype, int32_t device, void* stream, int64_t Bsz, int32_t N, int32_t Tsim, double dt,
                                const double* u_trim /* host, 4 */, const zb_arr* Q, const zb_arr* R, const zb_arr* Qf,
                                int32_t flags /* ZB_COST_DIAGONAL | ZB_VARIANT_* */, const void* x0, void* xSim_out, void* uSim_out);

/* ---- roofline denominators: dependent-FMA throughput probe; returns achieved FLOP/s ------------ */
ZB_API int32_t zb_peak_fma(int32_t dtype, int32_t device, double* flops_per_s_out, double* sm_clock_mhz_out);

#ifdef __cplusplus
}
#endif
#endif /* ZOPT_B200_H */
