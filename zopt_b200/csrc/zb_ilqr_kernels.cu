// Translation unit of the cooperative kernels: quadcopter iLQR / DDP backward pass, fused 16-way line search, fp64 (12,4) Riccati.
#include "ilqr_fast.cuh"
#include "ilqr_forward.cuh"
#include "lqr_quad64.cuh"
