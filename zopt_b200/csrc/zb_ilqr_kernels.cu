// Translation unit of the quadcopter iLQR / DDP kernels: cooperative backward pass and fused 16-way line search.
#include "ilqr_fast.cuh"
#include "ilqr_forward.cuh"
