// Parameter blocks and host launchers of the quadcopter iLQR / DDP kernels (ilqr_fast.cuh, ilqr_forward.cuh).
// The kernels are compiled in their own translation unit (zb_ilqr_kernels.cu); zb_api.cu only sees this header.
#pragma once
#include "zb_common.cuh"

namespace zb {

// Compaction of the problems that are still iterating (frozen problems must not keep lanes idle): `perm[0..*count)`
// lists them; thread groups map slot -> perm[slot] and groups past *count leave at once.  perm == nullptr: identity.
struct ActiveP {
    const int32_t* perm;
    const int32_t* count;
};

struct IlqrFastP {
    long long Bsz;
    int N;
    double dt;
    Cost C;
    const void *xTraj, *uTraj;
    const void* Czz;   // (Bsz,16,16)
    const void* Vfxx;  // (Bsz,12,12)
    const uint8_t* done;
    void *l, *L;
    double eps;  // ensurePositiveDefinite threshold (1e-3)
    ActiveP act;
    long long active_hint;  // host-side upper bound on *act.count (Bsz when unknown): sizes the grid and the CTAs
    int cold_start;  // DDP: every eigen-solve starts from the identity (test hook ZB_DDP_COLD_START; the warm start is the default)
    void* ev;    // DDP fp64: (Bsz,84) scratch, eigenvectors of the clamped 9x9 block carried from step to step (null: cold start every step)
};

// solver bookkeeping shared by the generic commit kernel and the fused kernel
struct CommitP {
    void* J;             // (Bsz) current cost, updated in place (null for the bare forwardPass2 entry point)
    uint8_t* converged;  // (Bsz)
    int32_t* iters;      // (Bsz)
    int32_t* alpha_log;  // (Bsz,maxIter) or null
    void* J_log;         // (Bsz,maxIter+1) or null
    int it, maxIter;
    double tol;
    void* J_out;         // bare entry point: (Bsz)
    int32_t* idx_out;    // bare entry point: (Bsz) or null
};

struct FwdQuadP {
    long long Bsz;
    int N;
    double dt;
    Cost C;
    const void *x0, *l, *L;
    void *xTraj, *uTraj;  // current trajectory (xPrev, uPrev), replaced in place by the winner
    void* spec;           // (2, Bsz, (N+1)*12 + N*4): speculative rollouts of alpha = 1, 1/2
    void* Jall;           // (Bsz,16) or null
    CommitP S;
    int cost_diagonal;    // the caller asserted diagonal Q, R (ZB_COST_DIAGONAL)
    ActiveP act;
};

// once per solve: conditioned cost blocks of a DIAGONAL quadratic cost + the initial rollout u_k = uGuess_k
struct SetupQuadP {
    long long Bsz;
    int N, maxIter;
    double dt, eps;
    Cost C;
    const void *x0, *uGuess;
    void *xTraj, *uTraj, *J;
    uint8_t* converged;
    int32_t* iters;
    int32_t* alpha_log;  // (Bsz,maxIter) or null
    void* J_log;         // (Bsz,maxIter+1) or null
    void *Czz, *Vfxx;    // (Bsz,16,16), (Bsz,12,12)
};

// cooperative (12,4) Riccati recursion, fp64 (lqr_quad64.cuh): discreteFiniteHorizonLqr, and lqrMpc.solve when x0 is set
struct LqrQuadP {
    long long Bsz;
    int N, T;
    Arr A, B, Q, R, Qf;  // Qf.p == nullptr: terminal value Q[T-1] (lqrUtils.py:172)
    void* L;             // (Bsz,N,4,12) gains (the workspace for lqrMpc.solve)
    void* V0;            // (Bsz,12,12) or null
    const void* x0;      // null: gains only
    void *u0, *xTraj, *uTraj;
    int8_t* status;
    int32_t* iters;
    int cost_diagonal;   // the caller asserted diagonal Q, R (ZB_COST_DIAGONAL)
};

// fused fp64 closed-loop LQR-MPC of the quadcopter (lqr_quad64.cuh)
struct ClosedLoopQuadP {
    long long Bsz;
    int N, Tsim;
    double dt;
    double utrim[4];
    Arr Q, R, Qf;
    const void* x0;  // (Bsz,12)
    void* xSim;      // (Bsz,Tsim+1,12)
    void* uSim;      // (Bsz,Tsim,4)  applied deviation u_t (the control sent to the plant is u_trim + u_t)
    int cost_diagonal;
};

inline bool ilqr_fast_eligible(const Model& M, int second_order, bool cost_diagonal) {
    return M.kind == ZB_MODEL_QUADCOPTER && !M.has_wind && (!second_order || cost_diagonal);
}
inline bool fwd_quad_eligible(const Model& M) { return M.kind == ZB_MODEL_QUADCOPTER && !M.has_wind; }
// backward pass (ilqr_fast.cuh) and fused line search (ilqr_forward.cuh); defined in zb_ilqr_kernels.cu
int32_t ilqr_fast_launch(int32_t dtype, const IlqrFastP& P, cudaStream_t stream, bool cost_diagonal, bool second_order);
int32_t fwd_quad_launch(int32_t dtype, const FwdQuadP& P, cudaStream_t stream);
int32_t solve_setup_quad_launch(int32_t dtype, const SetupQuadP& P, cudaStream_t stream);
// perm_out[0..*count_out) <- problems with converged[b] == 0 (count_out zeroed on the stream first); any order
int32_t compact_active_launch(long long Bsz, const uint8_t* converged, int32_t* perm_out, int32_t* count_out, cudaStream_t stream);
int32_t riccati_quad_launch(int32_t dtype, const LqrQuadP& P, cudaStream_t stream);
int32_t mpc_closed_loop_quad64_launch(const ClosedLoopQuadP& P, cudaStream_t stream);

}  // namespace zb
