// (n,m) = (12,4) cooperative Riccati kernel -- placeholder until the fast path lands.
#pragma once
#include "zb_common.cuh"

namespace zb {
inline bool lqr_fast_eligible(int32_t, const LqrP&) { return false; }
inline int32_t lqr_fast_launch(int32_t, const LqrP&, cudaStream_t) { return fail(-9, "fast path not built"); }
}  // namespace zb
