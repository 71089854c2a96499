// Translation unit of the small-batch kernels (nine lanes per problem, register-tiled): closed-loop LQR-MPC of BASELINE cfg 3
// sharded over 8 GPUs.  Kept apart from zb_api.cu so that a change here rebuilds in seconds.
#include "mpc_warp.cuh"
