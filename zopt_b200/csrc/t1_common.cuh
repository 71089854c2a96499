// Helpers shared by the (12,4) fp32 kernels: packed-triangle indexing, packed FP32x2 FMA, closed-loop parameter block.
#pragma once
#include "lqr_fast.cuh"
#include "quad_model_gen.cuh"

namespace zb {
namespace t1 {

constexpr int X4 = 0;     // [A | B] rows: 12 rows x 4 float4 (chunk 3 = B row)
constexpr int W4 = 48;    // W rows: 12 x 3 float4
constexpr int Q4 = 84;    // Q, lower-triangle-covering chunks: row i has i/4+1 float4 (24 in all)
constexpr int R4 = 108;   // R lower packed (10 words) + 2 pad
constexpr int NF4 = 111;  // float4 slots per problem (dense cost)
constexpr int NF4_DIAG = 88;  // diagonal cost: slots Q4..Q4+2 = diag(Q), Q4+3 = diag(R)
constexpr int STG = 13;   // staging stride (float4) for the transposed gain store, odd -> conflict-free

__host__ __device__ constexpr int tri(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }
__host__ __device__ constexpr int qoff(int i) { return i < 4 ? i : (i < 8 ? 4 + 2 * (i - 4) : 12 + 3 * (i - 8)); }

#define ZB_F4(v, e) ((e) == 0 ? (v).x : (e) == 1 ? (v).y : (e) == 2 ? (v).z : (v).w)

// (d0, d1) += a * (b0, b1) as ONE packed FP32x2 FMA (Blackwell `fma.rn.f32x2`, SASS FFMA2): the step is bound by
// issue slots / dispatch with one warp per scheduler, and the packed form halves the FMA instruction count.
#ifndef ZB_USE_FFMA2
#define ZB_USE_FFMA2 1
#endif
#ifndef ZB_T1_PANEL_FFMA2
#define ZB_T1_PANEL_FFMA2 1  // packed FMAs in the V [A | B] panel product too
#endif
__device__ __forceinline__ void fma2(float& d0, float& d1, float a, float b0, float b1) {
#if ZB_USE_FFMA2
    const float2 r = __ffma2_rn(make_float2(a, a), make_float2(b0, b1), make_float2(d0, d1));
    d0 = r.x;
    d1 = r.y;
#else
    d0 = fmaf(a, b0, d0);
    d1 = fmaf(a, b1, d1);
#endif
}

struct ClosedLoopP {
    long long Bsz;
    int N, Tsim;
    float dt;
    float utrim[4];
    Arr Q, R, Qf;
    const float* x0;  // (Bsz,12)
    float* xSim;      // (Bsz,Tsim+1,12)
    float* uSim;      // (Bsz,Tsim,4)  applied deviation u_t (the control sent to the plant is u_trim + u_t)
};

// Model evaluations of the closed-loop kernels (thread-per-problem, lqr_t1.cuh, and nine-lane, mpc_warp.cuh): inlined, so
// the generated expressions are contracted into FMAs in the context of each kernel (as real calls they cost the
// thread-per-problem kernel 3 %: 34.5 against 33.5 ms on cfg 3); the two kernels then agree to rounding, not bit for bit.
static __device__ __forceinline__ void closed_loop_linearize(const float* __restrict__ x, const float* __restrict__ ut, float dt,
                                                          float* __restrict__ A) {  // A = I + dt dF/dx(x, ut), row-major 12x12
    float J[144];
    QuadTrig<float> tr = quad_trig(x);
    quad_jac_x(tr, x, ut, J);
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int j = 0; j < 12; ++j) A[i * 12 + j] = fmaf(dt, J[i * 12 + j], (i == j) ? 1.f : 0.f);
}
static __device__ __forceinline__ void closed_loop_plant(float* __restrict__ x, const float* __restrict__ ua, float dt) {  // x <- x + dt F(x, ua)
    float xd[12];
    QuadTrig<float> tr = quad_trig(x);
    quad_xdot(tr, x, ua, xd);
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = fmaf(dt, xd[i], x[i]);
}

}  // namespace t1
}  // namespace zb
