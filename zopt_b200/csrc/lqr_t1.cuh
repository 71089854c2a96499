// Thread-per-problem Riccati kernel for (n, m) = (12, 4), fp32, time-invariant A, B, Q, R  (round-1 v2).
//
// Why this shape.  ncu on the first cooperative kernel (lqr_fast.cuh, 4 threads per problem) showed the
// shared-memory pipe at 91 % and the FMA pipe at 45 %: a 128-bit shared load costs four LSU wavefronts per
// warp whatever is broadcast, so time is set by WORDS LOADED PER THREAD, and a 12x4 accumulator tile with
// both operands streamed gives only ~2.3 FMA per loaded word (the SM delivers 32 words/cycle against 128
// FMA/cycle, i.e. >= 4 are needed).  Here ONE thread owns one problem and keeps the whole symmetric value
// matrix V (78 words) in registers across the horizon:
//   1.  W = V A in three 12x4 panels: per k one 128-bit load of the panel row feeds 48 FMAs (12 per loaded
//       word); W -> smem.  VB = V B the same way with B's rows held in registers, G0 = B^T (V B).
//   2.  G = G0 + R, 4x4 Cholesky (its latency hides behind step 3's FMAs).
//   3.  One pass over the rows of A, W, B: V' = Q + A^T W (LOWER TRIANGLE ONLY: 78 accumulators = the new V,
//       symmetry exact by construction, a third of the update's FMAs gone) and M = B^T W; then L = G^-1 M for
//       the 12 columns and V' -= M^T L, all in registers.  (Algebraically the reference's Joseph form,
//       lqrUtils.py:169; fp32 parity gated at 1e-5.)
// ~4,580 FMA and ~950 loaded/stored words per problem-step (4.8 FMA/word) instead of 6,460 FMA-slots and
// ~2,800 words.  Operands live in shared memory interleaved by lane ([float4 slot][lane]) so every 128-bit
// access of a warp is one contiguous 512 B row: conflict-free, and no thread ever reads another thread's
// slots -- the main loop needs no barrier.  One warp per CTA, 4 CTAs per SM (111 float4 = 1,776 B per problem).
#pragma once
#include <stdlib.h>
#include <algorithm>
#include "t1_common.cuh"

namespace zb {
namespace t1 {

// load the lower triangle of a symmetric 12x12 (row-major, global) into v[78]
__device__ __forceinline__ void load_sym_lower(const float* g, float* v) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int c = 0; c <= i / 4; ++c) {
            const float4 q = __ldg(g4 + i * 3 + c);
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * c + e <= i) v[tri(i, 4 * c + e)] = ZB_F4(q, e);
        }
}

// One backward Riccati step for the problem owned by this thread: v (lower triangle of V) is updated in place and the
// gain L = (R + B'VB)^-1 B'VA is returned in registers.  S = this lane's column of the shared-memory slab.
// RS = row stride of the slab in float4 (32 lanes; 33 in the time-varying kernel, whose cooperative loads need the pad);
// RFULL = R stored as four full rows (slots R4..R4+3) instead of the packed lower triangle.
// SPLIT: [A | B] kept as two contiguous blocks (A rows at slots 0..35, B rows at 36..47 -- how bulk copies deliver them)
// instead of row-interleaved (row k = slots 4k..4k+3, chunk 3 = B row).
template <bool SPLIT>
__device__ __forceinline__ int xs(int k, int c) { return SPLIT ? (c < 3 ? k * 3 + c : 36 + k) : X4 + k * 4 + c; }

// Hook: called when operand slots fall free inside the step, so that a streaming caller can refill them with the NEXT step's
// operands while this step still computes (k_riccati_t1_tv): cost_free() once Q and R have been read (before pass 3a),
// row_free(kk) once row kk of [A | B] has been consumed by pass 3a.
struct NoHook {
    __device__ __forceinline__ void cost_free() const {}
    __device__ __forceinline__ void row_free(int) const {}
};

template <bool QDIAG, int RS = 32, bool RFULL = false, bool SPLIT = false, typename Hook = NoHook>
__device__ __forceinline__ void riccati_step(float4* S, float (&v)[78], float (&L)[4][12], const Hook hook = Hook()) {
    // ---- 1. [W | VB] = V [A | B] in four 12x4 panels.  ROLLED loop: one ~600-instruction body re-used four
    //         times keeps the step's code inside the instruction cache (the fully unrolled first version
    //         stalled 0.8 cycle/instruction on instruction fetch with one warp per scheduler).
    float G[10];
    float4 xfirst = S[xs<SPLIT>(0, 0) * RS];  // row 0 of the next panel, fetched before the previous panel's epilogue
#pragma unroll 1
    for (int p = 0; p < 4; ++p) {
        float acc[12][4];
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = 0.f;
#pragma unroll
        for (int kk = 0; kk < 12; ++kk) {
            const float4 x4 = (kk == 0) ? xfirst : S[xs<SPLIT>(kk, p) * RS];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const float vik = v[tri(i, kk)];
#if ZB_T1_PANEL_FFMA2
                fma2(acc[i][0], acc[i][1], vik, x4.x, x4.y);
                fma2(acc[i][2], acc[i][3], vik, x4.z, x4.w);
#else
                acc[i][0] = fmaf(vik, x4.x, acc[i][0]);
                acc[i][1] = fmaf(vik, x4.y, acc[i][1]);
                acc[i][2] = fmaf(vik, x4.z, acc[i][2]);
                acc[i][3] = fmaf(vik, x4.w, acc[i][3]);
#endif
            }
        }
        xfirst = S[xs<SPLIT>(0, (p < 3) ? p + 1 : 0) * RS];
        if (p < 3) {
#pragma unroll
            for (int i = 0; i < 12; ++i) S[(W4 + i * 3 + p) * RS] = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
        } else {  // acc = V B: G = R + B^T (V B), lower triangle
            if (QDIAG) {
                const float4 rd = S[(Q4 + 3) * RS];
                G[0] = rd.x; G[1] = 0.f; G[2] = rd.y; G[3] = 0.f; G[4] = 0.f; G[5] = rd.z; G[6] = 0.f; G[7] = 0.f; G[8] = 0.f; G[9] = rd.w;
            } else {
                if (RFULL) {  // R as four full rows
                    const float4 r0 = S[(R4 + 0) * RS], r1 = S[(R4 + 1) * RS], r2 = S[(R4 + 2) * RS], r3 = S[(R4 + 3) * RS];
                    G[0] = r0.x; G[1] = r1.x; G[2] = r1.y; G[3] = r2.x; G[4] = r2.y;
                    G[5] = r2.z; G[6] = r3.x; G[7] = r3.y; G[8] = r3.z; G[9] = r3.w;
                } else {  // packed lower triangle
                    const float4 r0 = S[(R4 + 0) * RS], r1 = S[(R4 + 1) * RS], r2 = S[(R4 + 2) * RS];
                    G[0] = r0.x; G[1] = r0.y; G[2] = r0.z; G[3] = r0.w; G[4] = r1.x;
                    G[5] = r1.y; G[6] = r1.z; G[7] = r1.w; G[8] = r2.x; G[9] = r2.y;
                }
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const float4 b4 = S[xs<SPLIT>(i, 3) * RS];
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c <= a; ++c) G[tri(a, c)] = fmaf(ZB_F4(b4, a), acc[i][c], G[tri(a, c)]);
            }
        }
    }
    // ---- 2. Cholesky G = C C^T (G already holds G0 + R) ------------------------------------------------
    const float d0 = rsqrtf(G[0]);
    const float c10 = G[1] * d0, c20 = G[3] * d0, c30 = G[6] * d0;
    const float d1 = rsqrtf(fmaf(-c10, c10, G[2]));
    const float c21 = fmaf(-c20, c10, G[4]) * d1, c31 = fmaf(-c30, c10, G[7]) * d1;
    const float d2 = rsqrtf(fmaf(-c21, c21, fmaf(-c20, c20, G[5])));
    const float c32 = fmaf(-c31, c21, fmaf(-c30, c20, G[8])) * d2;
    const float d3 = rsqrtf(fmaf(-c32, c32, fmaf(-c31, c31, fmaf(-c30, c30, G[9]))));
    // ---- 3a. V' (lower) = Q + A^T W and M = B^T W in one pass over the rows of A, W, B ------------------
    // (independent of the Cholesky chain above, so the scheduler can hide its latency behind these FMAs)
    if (QDIAG) {
        const float4 q0 = S[(Q4 + 0) * RS], q1 = S[(Q4 + 1) * RS], q2 = S[(Q4 + 2) * RS];
        const float qd[12] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w};
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) v[tri(i, j)] = (i == j) ? qd[i] : 0.f;
    } else {
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int c = 0; c <= i / 4; ++c) {
                const float4 q = S[(Q4 + qoff(i) + c) * RS];
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (4 * c + e <= i) v[tri(i, 4 * c + e)] = ZB_F4(q, e);
            }
    }
    hook.cost_free();
    float M[4][12];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int j = 0; j < 12; ++j) M[a][j] = 0.f;
    {
        float4 ra[3], rw[3], rb;  // rows kk of A, W, B; next rows are fetched while the current ones are consumed
#pragma unroll
        for (int c = 0; c < 3; ++c) { ra[c] = S[xs<SPLIT>(0, c) * RS]; rw[c] = S[(W4 + c) * RS]; }
        rb = S[xs<SPLIT>(0, 3) * RS];
#pragma unroll 1
        for (int kk = 0; kk < 12; ++kk) {
            const float a[12] = {ra[0].x, ra[0].y, ra[0].z, ra[0].w, ra[1].x, ra[1].y, ra[1].z, ra[1].w, ra[2].x, ra[2].y, ra[2].z, ra[2].w};
            const float w[12] = {rw[0].x, rw[0].y, rw[0].z, rw[0].w, rw[1].x, rw[1].y, rw[1].z, rw[1].w, rw[2].x, rw[2].y, rw[2].z, rw[2].w};
            const float4 b4 = rb;
            const int kn = (kk < 11) ? kk + 1 : 11;
#pragma unroll
            for (int c = 0; c < 3; ++c) { ra[c] = S[xs<SPLIT>(kn, c) * RS]; rw[c] = S[(W4 + kn * 3 + c) * RS]; }
            rb = S[xs<SPLIT>(kn, 3) * RS];
            hook.row_free(kk);  // row kk sits in registers (a, b4); row kn has been read just above
#pragma unroll
            for (int i = 0; i < 12; ++i)
#pragma unroll
                for (int j = 0; j <= i; j += 2) {
                    if (j + 1 <= i) fma2(v[tri(i, j)], v[tri(i, j + 1)], a[i], w[j], w[j + 1]);
                    else v[tri(i, j)] = fmaf(a[i], w[j], v[tri(i, j)]);
                }
#pragma unroll
            for (int j = 0; j < 12; j += 2) {
                fma2(M[0][j], M[0][j + 1], b4.x, w[j], w[j + 1]);
                fma2(M[1][j], M[1][j + 1], b4.y, w[j], w[j + 1]);
                fma2(M[2][j], M[2][j + 1], b4.z, w[j], w[j + 1]);
                fma2(M[3][j], M[3][j + 1], b4.w, w[j], w[j + 1]);
            }
        }
    }
    // ---- 2b. L = G^-1 M (12 right-hand sides) ------------------------------------------------------
#pragma unroll
    for (int j = 0; j < 12; ++j) {
        const float y0 = M[0][j] * d0;
        const float y1 = fmaf(-c10, y0, M[1][j]) * d1;
        const float y2 = fmaf(-c21, y1, fmaf(-c20, y0, M[2][j])) * d2;
        const float y3 = fmaf(-c32, y2, fmaf(-c31, y1, fmaf(-c30, y0, M[3][j]))) * d3;
        const float x3 = y3 * d3;
        const float x2 = fmaf(-c32, x3, y2) * d2;
        const float x1 = fmaf(-c31, x3, fmaf(-c21, x2, y1)) * d1;
        const float x0 = fmaf(-c30, x3, fmaf(-c20, x2, fmaf(-c10, x1, y0))) * d0;
        L[0][j] = x0; L[1][j] = x1; L[2][j] = x2; L[3][j] = x3;
    }
    // ---- 3b. V' -= M^T L (lower) --------------------------------------------------------------------
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int j = 0; j <= i; j += 2) {
                if (j + 1 <= i) fma2(v[tri(i, j)], v[tri(i, j + 1)], -M[a][i], L[a][j], L[a][j + 1]);
                else v[tri(i, j)] = fmaf(-M[a][i], L[a][j], v[tri(i, j)]);
            }
}

// QDIAG: the caller asserts Q, R, Qf diagonal (ZB_COST_DIAGONAL): 4 float4 of cost data instead of 27, so FIVE
// CTAs fit an SM (88 float4 = 1,408 B per problem) and 65,536 problems take 3 rounds of CTAs instead of 4.
// ROLL = false (with MPC): sweep only -- the gains stay in the workspace and the plan is rolled out by k_plan_rollout_q4,
// launched on a second stream so that it overlaps the next chunk's sweep (riccati_t1_mpc_split_launch).
template <bool MPC, bool QDIAG, bool ROLL = true>
__global__ void __launch_bounds__(32, QDIAG ? 5 : 4) k_riccati_t1(FastP P) {
    extern __shared__ float4 sm[];
    const int lane = threadIdx.x;
    const long long b_raw = (long long)blockIdx.x * 32 + lane;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    float4* S = sm + lane;  // slot s of this problem: S[s * 32]

    // ---- stage this problem's operands (each lane its own problem; L1 keeps the partially used lines) ----
    {
        const float4* gA = reinterpret_cast<const float4*>(P.A.at<float>(b));
        const float4* gB = reinterpret_cast<const float4*>(P.B.at<float>(b));
        const float4* gQ = reinterpret_cast<const float4*>(P.Q.at<float>(b, 0));
        const float4* gR = reinterpret_cast<const float4*>(P.R.at<float>(b));
#pragma unroll
        for (int k = 0; k < 12; ++k) {
#pragma unroll
            for (int c = 0; c < 3; ++c) S[(X4 + k * 4 + c) * 32] = __ldg(gA + k * 3 + c);
            S[(X4 + k * 4 + 3) * 32] = __ldg(gB + k);
        }
        if (QDIAG) {
            const float* q = P.Q.at<float>(b, 0);
            const float* r = P.R.at<float>(b);
#pragma unroll
            for (int c = 0; c < 3; ++c)
                S[(Q4 + c) * 32] = make_float4(__ldg(q + (4 * c) * 13), __ldg(q + (4 * c + 1) * 13), __ldg(q + (4 * c + 2) * 13), __ldg(q + (4 * c + 3) * 13));
            S[(Q4 + 3) * 32] = make_float4(__ldg(r), __ldg(r + 5), __ldg(r + 10), __ldg(r + 15));
        } else {
#pragma unroll
            for (int k = 0; k < 12; ++k)
#pragma unroll
                for (int c = 0; c <= k / 4; ++c) S[(Q4 + qoff(k) + c) * 32] = __ldg(gQ + k * 3 + c);
            const float4 r0 = __ldg(gR), r1 = __ldg(gR + 1), r2 = __ldg(gR + 2), r3 = __ldg(gR + 3);
            S[(R4 + 0) * 32] = make_float4(r0.x, r1.x, r1.y, r2.x);
            S[(R4 + 1) * 32] = make_float4(r2.y, r2.z, r3.x, r3.y);
            S[(R4 + 2) * 32] = make_float4(r3.z, r3.w, 0.f, 0.f);
        }
    }
    float v[78];
    if (QDIAG) {
        const float* qf = MPC ? P.Qf.at<float>(b) : P.Q.at<float>(b, P.T - 1);
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int j = 0; j <= i; ++j) v[tri(i, j)] = (i == j) ? __ldg(qf + i * 13) : 0.f;
    } else {
        load_sym_lower(MPC ? P.Qf.at<float>(b) : P.Q.at<float>(b, P.T - 1), v);
    }

    const long long wblk = blockIdx.x;  // 32 problems per CTA
    // MPC: gains in the workspace, layout [cta][k][12 float4][lane]; dfh: public (Bsz,N,4,12)
    float4* gws = reinterpret_cast<float4*>(P.gains) + wblk * (long long)P.N * 12 * 32 + lane;
    float* gpub = P.gains + (long long)blockIdx.x * 32 * (long long)P.N * 48;

    for (int k = P.N - 1; k >= 0; --k) {
        float L[4][12];
        riccati_step<QDIAG>(S, v, L);
        // ---- 4. gains out ------------------------------------------------------------------------------
        if (MPC) {
            float4* g = gws + (long long)k * 12 * 32;
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    g[(a * 3 + c) * 32] = make_float4(L[a][4 * c], L[a][4 * c + 1], L[a][4 * c + 2], L[a][4 * c + 3]);
        } else {
            // public layout (Bsz,N,4,12): transpose through smem (the W region is free now) for coalesced stores
            float4* stg = sm + W4 * 32;
            __syncwarp();  // every lane is done reading its W slots before the region is reused for staging
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    stg[lane * STG + a * 3 + c] = make_float4(L[a][4 * c], L[a][4 * c + 1], L[a][4 * c + 2], L[a][4 * c + 3]);
            __syncwarp();
#pragma unroll
            for (int r = 0; r < 12; ++r) {
                const int idx = r * 32 + lane, pr = idx / 12, j4 = idx - pr * 12;
                const float4 val = stg[pr * STG + j4];
                if ((long long)blockIdx.x * 32 + pr < P.Bsz)
                    *reinterpret_cast<float4*>(gpub + ((long long)pr * P.N + k) * 48 + j4 * 4) = val;
            }
            __syncwarp();
        }
    }

    if (P.V0 && active) {
        float* o = P.V0 + b * 144;
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int j = 0; j < 12; ++j) o[i * 12 + j] = v[tri(i, j)];
    }

    if (MPC && !ROLL) {  // sweep-only launch: publish "this group's gains are in memory" to the concurrent rollout kernel
        __threadfence();
        __syncwarp();
        if (lane == 0) *reinterpret_cast<volatile int*>(P.iters + blockIdx.x) = 1;  // (iters: the flag array in this launch)
    }
    if (MPC && ROLL) {
        // ---- plan rollout x+ = A x + B u, u = -L_k x (mpcUtils.py:55).  A, B rows from smem; the gains of the
        //      next step are prefetched (L2) while the current step's dependent chain runs ----
        float x[12];
        {
            const float4* gx = reinterpret_cast<const float4*>(P.x0 + b * 12);
            const float4 x0 = __ldg(gx), x1 = __ldg(gx + 1), x2 = __ldg(gx + 2);
            x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
            x[8] = x2.x; x[9] = x2.y; x[10] = x2.z; x[11] = x2.w;
        }
        float4* xT = reinterpret_cast<float4*>(P.xTraj + b * (long long)(P.N + 1) * 12);
        float4* uT = reinterpret_cast<float4*>(P.uTraj + b * (long long)P.N * 4);
        if (active) {
            xT[0] = make_float4(x[0], x[1], x[2], x[3]); xT[1] = make_float4(x[4], x[5], x[6], x[7]); xT[2] = make_float4(x[8], x[9], x[10], x[11]);
        }
        // gains stream back through a 3-stage cp.async ring in the (now free) W region: the copies of steps k+1, k+2
        // are in flight while step k's dependent chain runs; every lane reads only the slots it filled (no barrier).
        auto stage_ptr = [&](int st, int j) -> float4* { return &S[(W4 + st * 12 + j) * 32]; };
        auto issue = [&](int kk) {
            const int ks = (kk < P.N) ? kk : P.N - 1;
            const float4* src = gws + (long long)ks * 12 * 32;
            const int st = kk % 3;
#pragma unroll
            for (int j = 0; j < 12; ++j) {
                const unsigned dst = (unsigned)__cvta_generic_to_shared(stage_ptr(st, j));
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst), "l"(src + j * 32) : "memory");
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        };
        issue(0);
        issue(1);
#pragma unroll 1
        for (int k = 0; k < P.N; ++k) {
            issue(k + 2);
            asm volatile("cp.async.wait_group 2;\n" ::: "memory");
            const int st = k % 3;
            float u[4];
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float4 l0 = *stage_ptr(st, a * 3), l1 = *stage_ptr(st, a * 3 + 1), l2 = *stage_ptr(st, a * 3 + 2);
                float s0 = l0.x * x[0], s1 = l1.x * x[4], s2 = l2.x * x[8];
                s0 = fmaf(l0.y, x[1], s0); s1 = fmaf(l1.y, x[5], s1); s2 = fmaf(l2.y, x[9], s2);
                s0 = fmaf(l0.z, x[2], s0); s1 = fmaf(l1.z, x[6], s1); s2 = fmaf(l2.z, x[10], s2);
                s0 = fmaf(l0.w, x[3], s0); s1 = fmaf(l1.w, x[7], s1); s2 = fmaf(l2.w, x[11], s2);
                u[a] = -((s0 + s1) + s2);
            }
            float xn[12];
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const float4 a0 = S[(X4 + i * 4 + 0) * 32], a1 = S[(X4 + i * 4 + 1) * 32], a2 = S[(X4 + i * 4 + 2) * 32], b4 = S[(X4 + i * 4 + 3) * 32];
                float s0 = a0.x * x[0], s1 = a1.x * x[4], s2 = a2.x * x[8], s3f = b4.x * u[0];
                s0 = fmaf(a0.y, x[1], s0); s1 = fmaf(a1.y, x[5], s1); s2 = fmaf(a2.y, x[9], s2); s3f = fmaf(b4.y, u[1], s3f);
                s0 = fmaf(a0.z, x[2], s0); s1 = fmaf(a1.z, x[6], s1); s2 = fmaf(a2.z, x[10], s2); s3f = fmaf(b4.z, u[2], s3f);
                s0 = fmaf(a0.w, x[3], s0); s1 = fmaf(a1.w, x[7], s1); s2 = fmaf(a2.w, x[11], s2); s3f = fmaf(b4.w, u[3], s3f);
                xn[i] = (s0 + s1) + (s2 + s3f);
            }
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = xn[i];
            if (active) {
                uT[k] = make_float4(u[0], u[1], u[2], u[3]);
                float4* xr = xT + (long long)(k + 1) * 3;
                xr[0] = make_float4(x[0], x[1], x[2], x[3]); xr[1] = make_float4(x[4], x[5], x[6], x[7]); xr[2] = make_float4(x[8], x[9], x[10], x[11]);
                if (k == 0) *reinterpret_cast<float4*>(P.u0 + b * 4) = make_float4(u[0], u[1], u[2], u[3]);
            }
        }
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
        if (active) {
            P.status[b] = 0;
            if (P.iters) P.iters[b] = 0;
        }
    }
}


// -------------------------------------------------------------------------------------------------------------
// Time-varying discreteFiniteHorizonLqr (zopt/lqrUtils.py:144-173 with genuinely different A[k], B[k], Q[k], R[k]), fp32,
// (12,4).  Same per-thread step; the step's operands are STREAMED from HBM with cp.async, each lane copying ITS OWN
// problem's blocks (76 x 16 B: A 36, B 12, the lower-triangle chunks of Q 24, R 4 -- 1,216 B per problem-step) straight
// into its column of the lane-interleaved slab: consecutive lanes write consecutive 16 B (conflict-free without padding,
// so the slab is 112 x 32 float4 = 56 KB and FOUR one-warp CTAs fit an SM), source and destination offsets are
// instruction immediates (2 instructions per copy), and a lane's reads walk its problem's contiguous blocks, so every
// fetched 32-byte sector is used in full.  The copies of step k-1 are issued as soon as the lane's own step k is done (the
// slots are private), overlapping the gain stores.  (An L2 prefetch of the next step's blocks and cp.async.ca were both
// measured slower -- the kernel is bound by the load/store unit -- and are compiled out: ZB_TV_L2_PREFETCH, ZB_TV_CP_CA.)
// History (ncu: profiles/r1s3_lqr_tv_*), 65,536 solves x N=50: v1 issued coalesced copies with a per-copy transposing
// (problem, chunk) index into a padded slab (7,300 instructions per step, 3 CTAs per SM): 2.76 ms; v2 (this one,
// 4,900 instructions per step, 4 CTAs per SM): 1.88 ms with the L2 prefetch, 1.74 ms without; v3 made the copies coalesced again with a fixed per-lane
// schedule (four problems x eight consecutive chunks per instruction, 8-way bank conflicts on the shared-memory side):
// 2.17 ms, 2.32 ms with the L2 prefetch -- slower than the scattered but conflict-free copies, so it was dropped.
// Bound: the load/store unit (operand copies + the step's own 128-bit shared-memory reads: lg_throttle 21 %,
// short_scoreboard 30 % of the stall samples), not the FMA pipe and not yet HBM (2.5 TB/s of 6.5).
#ifndef ZB_TV_CP_CA
#define ZB_TV_CP_CA 0  // .ca (allocate in L1) measured 2.21 ms against 1.74 ms with .cg
#endif
#ifndef ZB_TV_L2_PREFETCH
#define ZB_TV_L2_PREFETCH 0  // prefetch.global.L2 of the next step's blocks measured 1.87 ms against 1.74 ms without: the kernel is LSU-bound
#endif
#ifndef ZB_TV_EARLY_ISSUE
#define ZB_TV_EARLY_ISSUE 1  // operand copies of step k-1 issued from inside step k as the slots fall free (TvHook)
#endif
constexpr int RS_TV = 32;
constexpr int NF4_TV = 112;  // X 48 + W 36 + Q 24 + R 4 (full rows)

__device__ __forceinline__ void cp_async16(float4* dst_smem, const float4* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
#if ZB_TV_CP_CA
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
#else
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
#endif
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];\n" ::"l"(p)); }

__device__ __forceinline__ void tv_load_step(float4* S, const FastP& P, long long b, int k) {
    const float4* gA = reinterpret_cast<const float4*>(P.A.at<float>(b, k));
    const float4* gB = reinterpret_cast<const float4*>(P.B.at<float>(b, k));
    const float4* gQ = reinterpret_cast<const float4*>(P.Q.at<float>(b, k));
    const float4* gR = reinterpret_cast<const float4*>(P.R.at<float>(b, k));
#pragma unroll
    for (int r = 0; r < 12; ++r) {
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async16(&S[(X4 + r * 4 + c) * RS_TV], gA + r * 3 + c);
        cp_async16(&S[(X4 + r * 4 + 3) * RS_TV], gB + r);
    }
#pragma unroll
    for (int r = 0; r < 12; ++r)
#pragma unroll
        for (int c = 0; c <= r / 4; ++c) cp_async16(&S[(Q4 + qoff(r) + c) * RS_TV], gQ + r * 3 + c);
#pragma unroll
    for (int j = 0; j < 4; ++j) cp_async16(&S[(R4 + j) * RS_TV], gR + j);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}
// The same copies issued from INSIDE the step (round 2): the cost blocks as soon as the step has read Q and R, row kk of
// [A | B] as soon as pass 3a has consumed it -- the transfers of step k-1 then overlap the second half of step k (with one warp
// per scheduler nothing else can cover them), and the 76 copies no longer reach the load/store unit as one burst.
// Measured 1.735 -> 1.45 ms per 65,536 x 50 (ncu: lg_throttle 21 % -> 4 % of the stall samples).  Waiting for rows 8..11 only
// where the first panel of the next step reaches them (two commit groups) changed nothing (1.47 ms) and is not kept.
struct TvHook {
    float4* S;
    const float4 *gA, *gB, *gQ, *gR;
    bool on;
    __device__ __forceinline__ void cost_free() const {
        if (!on) return;
#pragma unroll
        for (int r = 0; r < 12; ++r)
#pragma unroll
            for (int c = 0; c <= r / 4; ++c) cp_async16(&S[(Q4 + qoff(r) + c) * RS_TV], gQ + r * 3 + c);
#pragma unroll
        for (int j = 0; j < 4; ++j) cp_async16(&S[(R4 + j) * RS_TV], gR + j);
    }
    __device__ __forceinline__ void row_free(int kk) const {
        if (!on) return;
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async16(&S[(X4 + kk * 4 + c) * RS_TV], gA + kk * 3 + c);
        cp_async16(&S[(X4 + kk * 4 + 3) * RS_TV], gB + kk);
    }
};

// pull the blocks of step k into L2: one request per 128-byte line a 16-byte-aligned block can touch
__device__ __forceinline__ void tv_prefetch_step(const FastP& P, long long b, int k) {
    const char* gA = reinterpret_cast<const char*>(P.A.at<float>(b, k));
    const char* gB = reinterpret_cast<const char*>(P.B.at<float>(b, k));
    const char* gQ = reinterpret_cast<const char*>(P.Q.at<float>(b, k));
    const char* gR = reinterpret_cast<const char*>(P.R.at<float>(b, k));
#pragma unroll
    for (int i = 0; i < 6; ++i) { prefetch_l2(gA + (i < 5 ? i * 128 : 560)); prefetch_l2(gQ + (i < 5 ? i * 128 : 560)); }
    prefetch_l2(gB); prefetch_l2(gB + 128); prefetch_l2(gB + 176);
    prefetch_l2(gR); prefetch_l2(gR + 48);
}

// Work rotation for the thread-per-problem kernels (round 2).  One warp owns 32 problems for the whole horizon and a B200 has
// 592 warp schedulers: 65,536 problems are 2,048 warps = 3.46 rounds, run as FOUR (the last one 46 % full: 13 % of the launch).
// With QUEUE the grid is a fixed set of one-warp workers (one per scheduler) and the horizon is cut into chunks of CH steps;
// ready problem groups wait in a FIFO: a worker pops a group, advances it by one chunk and pushes it back (the last chunk also
// does the group's epilogue).  A group is in the queue only while nobody works on it, so no pass waits for another, and FIFO
// order makes all groups advance at the same pace: the launch takes groups / workers x (time of one warp alone).  Between
// chunks the value matrix (78 words per problem) goes through a scratch array, [group][word][lane] so that every access is a
// coalesced 128-byte row.  The step itself is untouched: results are bit-identical to the static mapping.
struct T1Queue {
    unsigned* head;  // pops so far
    unsigned* tail;  // pushes so far (entry index = number of groups + this)
    int* ring;       // (groups * chunks) entries, -1 = not pushed yet
    int* prog;       // (groups) chunks completed
    float* vst;      // (groups, 78, 32) value matrices between chunks
    int chunk;       // steps per chunk
};

// pop the next ready group (every lane polls: a lane-0 spin followed by __syncwarp() leaves the warp split and the pass runs at
// half speed, see mpc_warp.cuh); returns false when all passes have been handed out
__device__ __forceinline__ bool t1q_pop(const T1Queue& Wq, long long total, int lane, long long& group, int& my_chunk) {
    unsigned h = 0;
    if (lane == 0) h = atomicAdd(Wq.head, 1u);
    h = __shfl_sync(0xffffffffu, h, 0);
    if ((long long)h >= total) return false;
    const volatile int* slot = Wq.ring + h;
    int j;
    while ((j = *slot) < 0) __nanosleep(100);
    __threadfence();  // acquire: progress word and value matrix were written before the push
    group = j;
    my_chunk = *reinterpret_cast<const volatile int*>(Wq.prog + j);
    return true;
}
__device__ __forceinline__ void t1q_push(const T1Queue& Wq, long long group, int chunks_done, int nchunks, long long ngroups, int lane) {
    __threadfence();  // this pass's stores are visible before the group is handed on
    __syncwarp();
    *reinterpret_cast<volatile int*>(Wq.prog + group) = chunks_done;  // (every lane stores the same word)
    if (chunks_done < nchunks) {
        unsigned tpos = 0;
        if (lane == 0) tpos = (unsigned)ngroups + atomicAdd(Wq.tail, 1u);  // the first `ngroups` entries are the seeds
        tpos = __shfl_sync(0xffffffffu, tpos, 0);
        __threadfence();
        *reinterpret_cast<volatile int*>(Wq.ring + tpos) = (int)group;
    }
}
// The first `n` ring entries (every group is ready for its first chunk) are written by a kernel of its own BEFORE the workers: a
// worker-seeded ring deadlocks as soon as the grid exceeds what the device holds at once (see mpc_warp.cuh::k_queue_seed).
__global__ void k_t1q_seed(int* ring, long long n) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j < n) ring[j] = (int)j;
}

// scratch of a rotating launch, stream-ordered (concurrent calls do not share it): counters, progress words, the ring, the value
// matrices.  *scratch is released by the caller with cudaFreeAsync after the launch.
inline cudaError_t t1q_alloc(T1Queue& Wq, long long groups, int N, int chunk, void** scratch, cudaStream_t stream) {
    const long long nchunks = (N + chunk - 1) / chunk;
    const size_t n_ints = (size_t)64 + (size_t)groups + (size_t)(groups * nchunks);
    const size_t v_off = (n_ints * sizeof(int) + 255) & ~(size_t)255;
    const size_t bytes = v_off + (size_t)groups * 78 * 32 * sizeof(float);
    cudaError_t e = keep_pool_memory();
    if (e != cudaSuccess) return e;
    if ((e = cudaMallocAsync(scratch, bytes, stream)) != cudaSuccess) return e;
    int* p = reinterpret_cast<int*>(*scratch);
    if ((e = cudaMemsetAsync(p, 0, (64 + (size_t)groups) * sizeof(int), stream)) != cudaSuccess) return e;
    if ((e = cudaMemsetAsync(p + 64 + groups, 0xFF, (size_t)(groups * nchunks) * sizeof(int), stream)) != cudaSuccess) return e;  // ring: all -1
    Wq = T1Queue{reinterpret_cast<unsigned*>(p), reinterpret_cast<unsigned*>(p + 32), p + 64 + groups, p + 64,
                 reinterpret_cast<float*>(reinterpret_cast<char*>(*scratch) + v_off), chunk};
    return cudaSuccess;
}

template <bool QUEUE>
__global__ void __launch_bounds__(32, 4) k_riccati_t1_tv(FastP P, T1Queue Wq) {
    extern __shared__ float4 sm[];
    const int lane = threadIdx.x;
    const long long ngroups = (P.Bsz + 31) / 32;
    const int nchunks = QUEUE ? (P.N + Wq.chunk - 1) / Wq.chunk : 1;
    float4* S = sm + lane;
  for (;;) {  // QUEUE: one pass per popped group; otherwise a single pass
    long long group = blockIdx.x;
    int my_chunk = 0;
    if (QUEUE && !t1q_pop(Wq, ngroups * nchunks, lane, group, my_chunk)) return;
    const int k_hi = QUEUE ? P.N - 1 - my_chunk * Wq.chunk : P.N - 1;
    const int k_lo = QUEUE ? max(0, k_hi - Wq.chunk + 1) : 0;
    const long long b0 = group * 32;
    const long long b_raw = b0 + lane;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    float v[78];
    if (QUEUE && my_chunk > 0) {
        const float* vs = Wq.vst + group * (78 * 32) + lane;
#pragma unroll
        for (int e = 0; e < 78; ++e) v[e] = __ldcg(vs + e * 32);
    } else {
        load_sym_lower(P.Q.at<float>(b, P.T - 1), v);  // lqrUtils.py:172: terminal value is Q[-1]
    }
    float* gpub = P.gains + b0 * (long long)P.N * 48;
    __syncwarp();  // (QUEUE: the previous pass's staging reads are over before the slab is refilled)
    tv_load_step(S, P, b, k_hi);
    for (int k = k_hi; k >= k_lo; --k) {
#if ZB_TV_L2_PREFETCH
        if (k > 0) tv_prefetch_step(P, b, k - 1);  // HBM -> L2 while this step computes
#endif
        __syncwarp();  // everybody is done with the staging area (the W region) of the previous step
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");  // each lane reads only what it copied itself
        float L[4][12];
#if ZB_TV_EARLY_ISSUE
        {
            const int kn = k > 0 ? k - 1 : 0;
            const TvHook hook{S, reinterpret_cast<const float4*>(P.A.at<float>(b, kn)), reinterpret_cast<const float4*>(P.B.at<float>(b, kn)),
                              reinterpret_cast<const float4*>(P.Q.at<float>(b, kn)), reinterpret_cast<const float4*>(P.R.at<float>(b, kn)), k > k_lo};
            riccati_step<false, RS_TV, true, false, TvHook>(S, v, L, hook);
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
#else
        riccati_step<false, RS_TV, true>(S, v, L);
        // the operand slots are private to the lane and free from here on: the copies of step k-1 (L2 hits) overlap the
        // gain transposition and stores below
        if (k > k_lo) tv_load_step(S, P, b, k - 1);
#endif
        float4* stg = sm + W4 * RS_TV;  // W region is free now: transpose the gains for coalesced stores
        __syncwarp();
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c)
                stg[lane * STG + a * 3 + c] = make_float4(L[a][4 * c], L[a][4 * c + 1], L[a][4 * c + 2], L[a][4 * c + 3]);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 12; ++r) {
            const int idx = r * 32 + lane, pr = idx / 12, j4 = idx - pr * 12;
            const float4 val = stg[pr * STG + j4];
            if (b0 + pr < P.Bsz) *reinterpret_cast<float4*>(gpub + ((long long)pr * P.N + k) * 48 + j4 * 4) = val;
        }
    }
    if (k_lo == 0 && P.V0 && active) {
        float* o = P.V0 + b * 144;
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int j = 0; j < 12; ++j) o[i * 12 + j] = v[tri(i, j)];
    }
    if (!QUEUE) return;
    if (k_lo > 0) {
        float* vs = Wq.vst + group * (78 * 32) + lane;
#pragma unroll
        for (int e = 0; e < 78; ++e) __stcg(vs + e * 32, v[e]);
    }
    t1q_push(Wq, group, my_chunk + 1, nchunks, ngroups, lane);
  }
}

// -------------------------------------------------------------------------------------------------------------
// The same time-varying recursion with the operands delivered by the TMA engine (round 2): `cp.async.bulk` global -> shared
// with an mbarrier carrying the byte count.  ncu on the cp.async version above: the load/store unit binds (lg_throttle 21 %,
// short_scoreboard 30 %): every lane issues 76 sixteen-byte copies per step on top of the step's own ~830 shared-memory
// loads.  Here a lane issues FOUR bulk copies per step (A 576 B, B 192 B, R 64 B, Q rows 8..11 192 B), which run in the copy
// engine, not the LSU, plus twelve cp.async for the scattered lower-triangle chunks of Q rows 0..7 (the same 1,216 B).
// A bulk copy lands contiguously, so the slab is PROBLEM-major: problem p owns float4 slots [113 p, 113 p + 112); 113 is
// 1 mod 8, so the eight lanes of a quarter-warp hit eight different 16-byte bank groups on every 128-bit access
// (conflict-free without interleaving).  Two-warp CTAs of 64 problems: 64 x 113 x 16 B = 113 KB, two CTAs fill the SM's
// 228 KB exactly (four warps per SM, as before); the mbarriers live in pad slots.  [A | B] arrive as two blocks
// (riccati_step<.., SPLIT>).  One mbarrier per WARP: lane 0 arms it with the step's 32 x 1,216 bytes, every lane issues its
// own problem's copies, the warp waits on the phase; the two warps of a CTA never synchronise.  The slots are private to the lane, so the copies of step k-1 are
// issued as soon as the lane's step k is done (fence.proxy.async orders its reads before the engine's writes) and overlap
// the gain transposition and stores.
// MEASURED (65,536 x N=50, one B200): 1.77 ms against 1.72 ms for the cp.async kernel -- no gain, and the experiments say why.
// (i) With one warp per scheduler nothing covers a warp's wait: the step is 4.4 us of arithmetic + 4.5 us of waiting for the
// operands requested at the end of the previous step (the kernel with the arithmetic removed streams in 1.08 ms, the
// time-invariant kernel computes in 0.88 ms, together 1.96 ms: the two barely overlap).  (ii) De-phasing the warps of an SM
// changes nothing (each scheduler still idles while ITS warp waits); an L2 prefetch one step ahead makes it slower (1.93 ms),
// so the wait is not HBM latency but the copy path itself: the engine costs ~6 ns per request per SM (twelve requests per
// lane-step: 15.3 us per warp-step; four: 8.9 us).  (iii) A second operand buffer (76 slots) would leave two warps per SM.
// What would lift it: double-buffering A only (148 slots, three warps per SM, est. 1.2-1.3 ms).
constexpr int PS4_TVB = 113;

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned phase) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(phase)
        : "memory");
}

// FOUR bulk copies per problem-step (A, B, R, rows 8..11 of Q: 1,024 B) -- the copy engine takes tens of nanoseconds per
// request whatever its size, so the twelve-request version (Q rows 0..7 as eight 16/32-byte bulk copies) spent 10 us per
// warp-step in the engine against 4.4 us of arithmetic -- and twelve per-lane cp.async for the lower-triangle-covering
// chunks of Q rows 0..7 (192 B), which do not form a contiguous block.
__device__ __forceinline__ void tvb_issue(float4* S, const FastP& P, long long b, int k, unsigned long long* bar) {
    const float* gA = P.A.at<float>(b, k);
    const float* gB = P.B.at<float>(b, k);
    const float* gQ = P.Q.at<float>(b, k);
    const float* gR = P.R.at<float>(b, k);
    bulk_g2s(S + 0, gA, 576, bar);
    bulk_g2s(S + 36, gB, 192, bar);
    bulk_g2s(S + Q4 + 12, gQ + 96, 192, bar);  // rows 8..11, whole
    bulk_g2s(S + R4, gR, 64, bar);
    const float4* q4 = reinterpret_cast<const float4*>(gQ);
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c <= r / 4; ++c) cp_async16(S + Q4 + qoff(r) + c, q4 + r * 3 + c);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

__global__ void __launch_bounds__(64, 2) k_riccati_t1_tvb(FastP P) {
    extern __shared__ float4 sm_all[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* sm = sm_all + warp * 32 * PS4_TVB;  // this warp's 32 problem slabs
    unsigned long long& bar = *reinterpret_cast<unsigned long long*>(sm + 112);  // pad slot of the warp's first problem
    const long long b0 = ((long long)blockIdx.x * 2 + warp) * 32;
    const long long b_raw = b0 + lane;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;  // idle lanes shadow the last problem: the byte count per step is fixed
    float4* S = sm + lane * PS4_TVB;
    constexpr unsigned STEP_BYTES = 32u * 1024u;  // per warp-step through the copy engine (the other 192 B per problem: cp.async)
    if (lane == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(smem_u32(&bar)) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(STEP_BYTES) : "memory");
    }
    __syncwarp();
    tvb_issue(S, P, b, P.N - 1, &bar);
    float v[78];
    load_sym_lower(P.Q.at<float>(b, P.T - 1), v);  // lqrUtils.py:172: terminal value is Q[-1]
    float* gpub = P.gains + b0 * (long long)P.N * 48;
    unsigned phase = 0;
    for (int k = P.N - 1; k >= 0; --k) {
        asm volatile("cp.async.wait_group 0;\n" ::: "memory");  // this lane's own Q chunks
        mbar_wait(&bar, phase);
        phase ^= 1u;
        float L[4][12];
        riccati_step<false, 1, true, true>(S, v, L);
        if (k > 0) {  // this lane is done with its operand slots: hand them back to the copy engine for step k-1
            asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(&bar)), "r"(STEP_BYTES) : "memory");
            __syncwarp();
            tvb_issue(S, P, b, k - 1, &bar);
        }
        // gains: each lane parks its 12 float4 in its own W slots, then the warp stores them transposed (coalesced rows of the
        // public (Bsz,N,4,12) layout)
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 3; ++c) S[W4 + a * 3 + c] = make_float4(L[a][4 * c], L[a][4 * c + 1], L[a][4 * c + 2], L[a][4 * c + 3]);
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 12; ++r) {
            const int idx = r * 32 + lane, pr = idx / 12, j4 = idx - pr * 12;
            const float4 val = sm[pr * PS4_TVB + W4 + j4];
            if (b0 + pr < P.Bsz) *reinterpret_cast<float4*>(gpub + ((long long)pr * P.N + k) * 48 + j4 * 4) = val;
        }
        __syncwarp();  // the W slots are overwritten by the next step
    }
    if (P.V0 && active) {
        float* o = P.V0 + b * 144;
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int j = 0; j < 12; ++j) o[i * 12 + j] = v[tri(i, j)];
    }
}

// -------------------------------------------------------------------------------------------------------------
// Closed-loop LQR-MPC for the quadcopter (BASELINE cfg 3; the loop of demos/lqrMpc.py:42-47 with a nonlinear plant):
// for every simulation step t:  A_t = I + dt dF/dx(x_t, u_trim), B = dt dF/du  (linearised IN the kernel),
// full Riccati sweep of horizon N from Qf (nothing cached between steps), u_t = -L_0 x_t  (the first move of the
// unconstrained lqrMpc plan, mpcUtils.py:47-59 with infinite bounds), plant step x_{t+1} = x_t + dt F(x_t, u_trim + u_t).
// Only the simulated trajectory leaves the chip: no gains, no plans, no A/B in HBM.
template <bool QDIAG>
__global__ void __launch_bounds__(32, QDIAG ? 5 : 4) k_mpc_closed_loop_quad(ClosedLoopP P) {
    extern __shared__ float4 sm[];
    const int lane = threadIdx.x;
    const long long b_raw = (long long)blockIdx.x * 32 + lane;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    float4* S = sm + lane;
    const float dt = P.dt;
    // constant operands: B rows (chunk 3 of every X row), cost
#pragma unroll
    for (int k = 0; k < 12; ++k)
        S[(X4 + k * 4 + 3) * 32] = make_float4(k == 2 ? -dt : 0.f, k == 3 ? dt : 0.f, k == 4 ? dt : 0.f, k == 5 ? dt : 0.f);
    if (QDIAG) {
        const float* q = P.Q.at<float>(b);
        const float* r = P.R.at<float>(b);
#pragma unroll
        for (int c = 0; c < 3; ++c)
            S[(Q4 + c) * 32] = make_float4(__ldg(q + (4 * c) * 13), __ldg(q + (4 * c + 1) * 13), __ldg(q + (4 * c + 2) * 13), __ldg(q + (4 * c + 3) * 13));
        S[(Q4 + 3) * 32] = make_float4(__ldg(r), __ldg(r + 5), __ldg(r + 10), __ldg(r + 15));
    } else {
        const float4* gQ = reinterpret_cast<const float4*>(P.Q.at<float>(b));
        const float4* gR = reinterpret_cast<const float4*>(P.R.at<float>(b));
#pragma unroll
        for (int k = 0; k < 12; ++k)
#pragma unroll
            for (int c = 0; c <= k / 4; ++c) S[(Q4 + qoff(k) + c) * 32] = __ldg(gQ + k * 3 + c);
        const float4 r0 = __ldg(gR), r1 = __ldg(gR + 1), r2 = __ldg(gR + 2), r3 = __ldg(gR + 3);
        S[(R4 + 0) * 32] = make_float4(r0.x, r1.x, r1.y, r2.x);
        S[(R4 + 1) * 32] = make_float4(r2.y, r2.z, r3.x, r3.y);
        S[(R4 + 2) * 32] = make_float4(r3.z, r3.w, 0.f, 0.f);
    }
    float x[12];
    {
        const float4* gx = reinterpret_cast<const float4*>(P.x0 + b * 12);
        const float4 x0 = __ldg(gx), x1 = __ldg(gx + 1), x2 = __ldg(gx + 2);
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
        x[8] = x2.x; x[9] = x2.y; x[10] = x2.z; x[11] = x2.w;
    }
    float4* xS = reinterpret_cast<float4*>(P.xSim + b * (long long)(P.Tsim + 1) * 12);
    float4* uS = reinterpret_cast<float4*>(P.uSim + b * (long long)P.Tsim * 4);
    const float ut[4] = {P.utrim[0], P.utrim[1], P.utrim[2], P.utrim[3]};
#pragma unroll 1
    for (int t = 0; t < P.Tsim; ++t) {
        if (active) {
            xS[(long long)t * 3 + 0] = make_float4(x[0], x[1], x[2], x[3]);
            xS[(long long)t * 3 + 1] = make_float4(x[4], x[5], x[6], x[7]);
            xS[(long long)t * 3 + 2] = make_float4(x[8], x[9], x[10], x[11]);
        }
        // ---- linearise at (x_t, u_trim): A = I + dt * dF/dx  -> rows of X (chunks 0..2) ----
        {
            float Am[144];
            closed_loop_linearize(x, ut, dt, Am);
#pragma unroll
            for (int i = 0; i < 12; ++i)
#pragma unroll
                for (int c = 0; c < 3; ++c)
                    S[(X4 + i * 4 + c) * 32] = make_float4(Am[i * 12 + 4 * c + 0], Am[i * 12 + 4 * c + 1], Am[i * 12 + 4 * c + 2], Am[i * 12 + 4 * c + 3]);
        }
        // ---- Riccati sweep from Qf; only the last gain (k = 0) is needed ----
        float v[78];
        {
            const float* qf = P.Qf.at<float>(b);
            if (QDIAG) {
#pragma unroll
                for (int i = 0; i < 12; ++i)
#pragma unroll
                    for (int j = 0; j <= i; ++j) v[tri(i, j)] = (i == j) ? __ldg(qf + i * 13) : 0.f;
            } else {
                load_sym_lower(qf, v);
            }
        }
        float L[4][12];
#pragma unroll 1
        for (int k = P.N - 1; k >= 0; --k) riccati_step<QDIAG>(S, v, L);
        // ---- u_t = -L_0 x_t ; plant: x <- x + dt F(x, u_trim + u_t) ----
        float u[4], ua[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float s = 0.f;
#pragma unroll
            for (int j = 0; j < 12; ++j) s = fmaf(L[a][j], x[j], s);
            u[a] = -s;
            ua[a] = ut[a] + u[a];
        }
        if (active) uS[t] = make_float4(u[0], u[1], u[2], u[3]);
        closed_loop_plant(x, ua, dt);
    }
    if (active) {
        xS[(long long)P.Tsim * 3 + 0] = make_float4(x[0], x[1], x[2], x[3]);
        xS[(long long)P.Tsim * 3 + 1] = make_float4(x[4], x[5], x[6], x[7]);
        xS[(long long)P.Tsim * 3 + 2] = make_float4(x[8], x[9], x[10], x[11]);
    }
}

// -------------------------------------------------------------------------------------------------------------
// Plan rollout of lqrMpc.solve (x+ = A x + B u, u = -L_k x, zopt/mpcUtils.py:55) as its OWN kernel, four threads per problem,
// no shared memory and ~100 registers, so that its CTAs fit beside a full wave of sweep CTAs (which own all of the SM's
// shared memory but only 60 % of its registers).  In the fused kernel the rollout is 18 % of the launch: every warp of a
// wave reaches it at the same time, 227 MB of gains are re-read with nothing to overlap them.  Split, the rollout of chunk
// c runs on a second stream while chunk c+1 sweeps.  Thread t of a quad keeps rows 3t..3t+2 of [A | B] in registers, computes
// u_t = -L_k[t,:] x and its three rows of x+; x and u are all-gathered with shuffles.  The gains come from the workspace in the
// sweep's layout [32-problem group][k][12 float4][lane] (a quad reads 4 x 16 B per row, eight neighbouring problems 128 B),
// fetched one step ahead.  Operation order = the fused kernel's rollout, so both paths give the same bits.
constexpr int ROLL_D = 4;  // ring depth of the rollout kernel: 3 steps of gains in flight per thread (24 KB per CTA)
__global__ void __launch_bounds__(128) k_plan_rollout_q4(FastP P, const int* flags) {
    extern __shared__ float4 ring[];  // [ROLL_D][3][128]
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31, t = lane & 3, qbase = lane & ~3;
    // grid-stride over groups of 32 problems: beside a sweeping chunk the launcher starts only two CTAs per SM (more would take
    // the registers the sweep CTAs need and serialise the two kernels)
    for (long long grp = blockIdx.x; grp * 32 < P.Bsz; grp += gridDim.x) {
    if (flags) {  // the sweep CTA of this group signals when its gains are stored (it runs concurrently on the caller's stream)
        if (threadIdx.x == 0) {
            while (*reinterpret_cast<const volatile int*>(flags + grp) == 0) __nanosleep(200);
            __threadfence();
        }
        __syncthreads();
    }
    const long long b_raw = grp * 32 + (threadIdx.x >> 2);
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    // rows 3t..3t+2 of [A | B]
    float a[3][12], bb[3][4];
    {
        const float4* gA = reinterpret_cast<const float4*>(P.A.at<float>(b));
        const float4* gB = reinterpret_cast<const float4*>(P.B.at<float>(b));
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
            for (int c = 0; c < 3; ++c) {
                const float4 v = __ldg(gA + (3 * t + r) * 3 + c);
                a[r][4 * c] = v.x; a[r][4 * c + 1] = v.y; a[r][4 * c + 2] = v.z; a[r][4 * c + 3] = v.w;
            }
            const float4 v = __ldg(gB + 3 * t + r);
            bb[r][0] = v.x; bb[r][1] = v.y; bb[r][2] = v.z; bb[r][3] = v.w;
        }
    }
    float x[12];
    {
        const float4* gx = reinterpret_cast<const float4*>(P.x0 + b * 12);
        const float4 x0 = __ldg(gx), x1 = __ldg(gx + 1), x2 = __ldg(gx + 2);
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
        x[8] = x2.x; x[9] = x2.y; x[10] = x2.z; x[11] = x2.w;
    }
    float* xT = P.xTraj + b * (long long)(P.N + 1) * 12;
    float* uT = P.uTraj + b * (long long)P.N * 4;
    if (active && t < 3) *reinterpret_cast<float4*>(xT + 4 * t) = make_float4(x[4 * t], x[4 * t + 1], x[4 * t + 2], x[4 * t + 3]);
    // gain row t of step k: float4 slots t*3 .. t*3+2 of the problem's column in its group's slab, streamed through a
    // ROLL_D-stage cp.async ring whose slots are private to the thread (it reads what it copied: no barrier)
    const float4* g = reinterpret_cast<const float4*>(P.gains) + (b >> 5) * (long long)P.N * 12 * 32 + (t * 3) * 32 + (b & 31);
    float4* my = ring + threadIdx.x;  // stage st, chunk c: my[(st * 3 + c) * 128]
    auto issue = [&](int kk) {
        const float4* gn = g + (long long)(kk < P.N ? kk : P.N - 1) * 12 * 32;
        const int st = kk % ROLL_D;
#pragma unroll
        for (int c = 0; c < 3; ++c) cp_async16(my + (st * 3 + c) * 128, gn + c * 32);
        asm volatile("cp.async.commit_group;\n" ::: "memory");
    };
#pragma unroll
    for (int kk = 0; kk < ROLL_D - 1; ++kk) issue(kk);
#pragma unroll 1
    for (int k = 0; k < P.N; ++k) {
        issue(k + ROLL_D - 1);
        asm volatile("cp.async.wait_group %0;\n" ::"n"(ROLL_D - 1) : "memory");
        const int st = k % ROLL_D;
        const float4 c0 = my[(st * 3 + 0) * 128], c1 = my[(st * 3 + 1) * 128], c2 = my[(st * 3 + 2) * 128];
        float s0 = c0.x * x[0], s1 = c1.x * x[4], s2 = c2.x * x[8];
        s0 = fmaf(c0.y, x[1], s0); s1 = fmaf(c1.y, x[5], s1); s2 = fmaf(c2.y, x[9], s2);
        s0 = fmaf(c0.z, x[2], s0); s1 = fmaf(c1.z, x[6], s1); s2 = fmaf(c2.z, x[10], s2);
        s0 = fmaf(c0.w, x[3], s0); s1 = fmaf(c1.w, x[7], s1); s2 = fmaf(c2.w, x[11], s2);
        const float ut = -((s0 + s1) + s2);
        float u[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) u[q] = __shfl_sync(FULL, ut, qbase + q);
        float xn[3];
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            float p0 = a[r][0] * x[0], p1 = a[r][4] * x[4], p2 = a[r][8] * x[8], p3 = bb[r][0] * u[0];
            p0 = fmaf(a[r][1], x[1], p0); p1 = fmaf(a[r][5], x[5], p1); p2 = fmaf(a[r][9], x[9], p2); p3 = fmaf(bb[r][1], u[1], p3);
            p0 = fmaf(a[r][2], x[2], p0); p1 = fmaf(a[r][6], x[6], p1); p2 = fmaf(a[r][10], x[10], p2); p3 = fmaf(bb[r][2], u[2], p3);
            p0 = fmaf(a[r][3], x[3], p0); p1 = fmaf(a[r][7], x[7], p1); p2 = fmaf(a[r][11], x[11], p2); p3 = fmaf(bb[r][3], u[3], p3);
            xn[r] = (p0 + p1) + (p2 + p3);
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = __shfl_sync(FULL, xn[i % 3], qbase + i / 3);
        if (active) {
            if (t == 3) {
                *reinterpret_cast<float4*>(uT + (long long)k * 4) = make_float4(u[0], u[1], u[2], u[3]);
                if (k == 0) *reinterpret_cast<float4*>(P.u0 + b * 4) = make_float4(u[0], u[1], u[2], u[3]);
            } else {
                *reinterpret_cast<float4*>(xT + (long long)(k + 1) * 12 + 4 * t) = make_float4(x[4 * t], x[4 * t + 1], x[4 * t + 2], x[4 * t + 3]);
            }
        }
    }
    asm volatile("cp.async.wait_group 0;\n" ::: "memory");
    if (active && t == 0) {
        P.status[b] = 0;
        if (P.iters) P.iters[b] = 0;
    }
    }
}

}  // namespace t1

// eligibility on top of the 4-thread kernel's: Q time-invariant and symmetric handling (lower triangle is used)
template <bool MPC, bool QDIAG>
inline int32_t riccati_t1_launch_impl(const FastP& F, cudaStream_t stream) {
    const size_t smem = (size_t)(QDIAG ? t1::NF4_DIAG : t1::NF4) * 32 * sizeof(float4);
    ZB_CUDA(cudaFuncSetAttribute(t1::k_riccati_t1<MPC, QDIAG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((F.Bsz + 31) / 32);
    t1::k_riccati_t1<MPC, QDIAG><<<grid, 32, smem, stream>>>(F);
    ZB_CUDA(cudaGetLastError());
    return 0;
}

template <bool MPC>
inline int32_t riccati_t1_launch(const FastP& F, cudaStream_t stream, bool cost_diagonal = false) {
    return cost_diagonal ? riccati_t1_launch_impl<MPC, true>(F, stream) : riccati_t1_launch_impl<MPC, false>(F, stream);
}

// lqrMpc.solve for batches of more than one wave: ONE sweep launch over the whole batch on the caller's stream (CTAs scheduled
// dynamically, no wave barrier) and, concurrently on a side stream, a persistent rollout kernel of two CTAs per SM that
// follows it group by group: a sweep CTA raises its group's flag once its 32 problems' gains are stored, the rollout CTA
// that owns the group waits for the flag, then rolls the plans out.  The rollout kernel has no shared memory and ~100
// registers, so two of its CTAs fit beside five sweep CTAs on every SM.  It is launched AFTER the sweep so that tools that
// serialise kernels (ncu) still terminate; the caller's stream finally waits for it (fork / join with events, capturable).
// `flags`: (Bsz+31)/32 ints of the workspace, zeroed on the stream first.
// MEASURED (65,536 solves, one B200): 0.88 ms, the same as the fused kernel (0.873 ms) -- and so were a chunked version (one
// wave of sweeps per launch, the previous chunk's rollout beside it: 0.98-1.04 ms, each wave ends on the slowest scheduler)
// and a register-only rollout kernel at two CTAs per SM.  What the experiments show instead: the sweep alone (gains stored,
// no rollout) takes 0.727 ms = 4 rounds of 592 one-warp CTAs at ~182 us (65,536 problems are 3.46 rounds: the last one is
// half empty), and the rollout needs ~45 us per 32-problem group however it is fed, so ~5,000 problems must be in flight
// to keep up with the sweep -- more than the registers and shared memory left beside it hold.  Opt-in (ZB_MPC_SPLIT_ROLLOUT).
constexpr size_t ROLL_SMEM = (size_t)t1::ROLL_D * 3 * 128 * sizeof(float4);
template <bool QDIAG>
inline int32_t riccati_t1_sweep_only(const FastP& F, cudaStream_t stream) {
    // FOUR sweep CTAs per SM (one warp per scheduler: the same throughput as five, which only make two warps share a
    // scheduler) leave room for one rollout CTA and its gain ring: the request is padded so that exactly four fit beside it
    size_t smem = (size_t)(QDIAG ? t1::NF4_DIAG : t1::NF4) * 32 * sizeof(float4);
    const size_t pad4 = (233472 - (ROLL_SMEM + 1024)) / 4 - 1024;
    if (QDIAG && smem < pad4) smem = pad4 & ~(size_t)15;
    ZB_CUDA(cudaFuncSetAttribute(t1::k_riccati_t1<true, QDIAG, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    t1::k_riccati_t1<true, QDIAG, false><<<(unsigned)((F.Bsz + 31) / 32), 32, smem, stream>>>(F);
    ZB_CUDA(cudaGetLastError());
    return 0;
}
inline int32_t riccati_t1_mpc_split_launch(const FastP& F, cudaStream_t stream, bool cost_diagonal, int device, int sm_count, int* flags) {
    cudaStream_t side;
    int32_t rc = side_stream(device, &side);
    if (rc) return rc;
    // CTAs of two kernels share an SM only under the same shared-memory carve-out: the rollout kernel (no shared memory of
    // its own) is told to prefer the sweep kernel's maximum carve-out
    ZB_CUDA(cudaFuncSetAttribute(t1::k_plan_rollout_q4, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
    const long long groups = (F.Bsz + 31) / 32;
    ZB_CUDA(cudaMemsetAsync(flags, 0, sizeof(int) * (size_t)groups, stream));
    cudaEvent_t ev;
    ZB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ZB_CUDA(cudaEventRecord(ev, stream));  // fork: the side stream starts once the flags are cleared
    ZB_CUDA(cudaStreamWaitEvent(side, ev, 0));
    ZB_CUDA(cudaEventDestroy(ev));  // released once it has completed
    FastP S = F;
    S.iters = flags;  // the sweep-only kernel writes the group flags through this field
    rc = cost_diagonal ? riccati_t1_sweep_only<true>(S, stream) : riccati_t1_sweep_only<false>(S, stream);
    if (rc) return rc;
    const long long grid = groups < (long long)sm_count ? groups : (long long)sm_count;  // one persistent rollout CTA per SM
    ZB_CUDA(cudaFuncSetAttribute(t1::k_plan_rollout_q4, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ROLL_SMEM));
    t1::k_plan_rollout_q4<<<(unsigned)grid, 128, ROLL_SMEM, side>>>(F, flags);
    ZB_CUDA(cudaGetLastError());
    ZB_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    ZB_CUDA(cudaEventRecord(ev, side));  // join
    ZB_CUDA(cudaStreamWaitEvent(stream, ev, 0));
    ZB_CUDA(cudaEventDestroy(ev));
    return 0;
}


#ifndef ZB_TV_BULK
#define ZB_TV_BULK 0  // default kernel: 1 = operands by cp.async.bulk + mbarrier (k_riccati_t1_tvb), 0 = per-lane cp.async (k_riccati_t1_tv).
#endif                // Measured equal (1.77 vs 1.72 ms per 65,536 x 50): see the kernel's header; the flag ZB_TV_BULK_COPY selects it per call
inline int32_t riccati_t1_tv_launch(const FastP& F, cudaStream_t stream, bool bulk = ZB_TV_BULK != 0) {
    if (bulk) {
        const size_t smem = (size_t)t1::PS4_TVB * 64 * sizeof(float4);
        ZB_CUDA(cudaFuncSetAttribute(t1::k_riccati_t1_tvb, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const unsigned grid = (unsigned)((F.Bsz + 63) / 64);
        t1::k_riccati_t1_tvb<<<grid, 64, smem, stream>>>(F);
        ZB_CUDA(cudaGetLastError());
        return 0;
    }
    const size_t smem = (size_t)t1::NF4_TV * t1::RS_TV * sizeof(float4);
    const long long groups = (F.Bsz + 31) / 32;
    int dev = 0, sms = 148;
    ZB_CUDA(cudaGetDevice(&dev));
    ZB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long workers = 4LL * sms;  // one warp per scheduler = four of these 56 KB CTAs per SM
    // Work rotation (T1Queue) is OPT-IN here (ZB_T1_ROTATE=1): measured on this kernel it is bit-identical but not reliably faster.
    // 65,536 problems (3.46 rounds): 1.456 -> 1.44-1.47 ms -- the kernel is bound by the copy path, the 272 warps of its last round
    // already run faster than a full round's, so there is no idle tail to recover.  20,000 problems (1.06 rounds): 0.72 -> 0.54 ms in
    // one run, 0.71 ms in the next, 1.4 ms with two chunks per horizon: every pass starts with an exposed operand load.
    // (The same rotation on the time-invariant headline kernel k_riccati_t1, with the value matrix AND a re-staging of A, B, Q, R
    // per pass, was built and dropped: 0.827 -> 0.89 ms at two chunks per horizon, worse with more.  With five CTAs per SM the
    // static mapping already runs 3.46 rounds in 0.885 of the time of four, so at most 2 % was there to gain.)
    bool rotate = false;
    if (const char* e = getenv("ZB_T1_ROTATE")) rotate = atoi(e) != 0 && groups > 1 && F.N >= 2;
    if (!rotate) {
        ZB_CUDA(cudaFuncSetAttribute(t1::k_riccati_t1_tv<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        t1::k_riccati_t1_tv<false><<<(unsigned)groups, 32, smem, stream>>>(F, t1::T1Queue{nullptr, nullptr, nullptr, nullptr, nullptr, 0});
        ZB_CUDA(cudaGetLastError());
        return 0;
    }
    int chunk = (F.N + 3) / 4;  // four chunks per horizon
    if (const char* e = getenv("ZB_T1_CHUNK")) chunk = atoi(e) > 0 ? atoi(e) : chunk;
    t1::T1Queue Wq;
    void* scratch = nullptr;
    ZB_CUDA(t1::t1q_alloc(Wq, groups, F.N, chunk, &scratch, stream));
    ZB_CUDA(cudaFuncSetAttribute(t1::k_riccati_t1_tv<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    t1::k_t1q_seed<<<(unsigned)((groups + 255) / 256), 256, 0, stream>>>(Wq.ring, groups);
    ZB_CUDA(cudaGetLastError());
    t1::k_riccati_t1_tv<true><<<(unsigned)std::min<long long>(workers, groups), 32, smem, stream>>>(F, Wq);
    ZB_CUDA(cudaGetLastError());
    ZB_CUDA(cudaFreeAsync(scratch, stream));
    return 0;
}

inline int32_t mpc_closed_loop_launch(const t1::ClosedLoopP& P, cudaStream_t stream, bool cost_diagonal) {
    const unsigned grid = (unsigned)((P.Bsz + 31) / 32);
    if (cost_diagonal) {
        const size_t smem = (size_t)t1::NF4_DIAG * 32 * sizeof(float4);
        ZB_CUDA(cudaFuncSetAttribute(t1::k_mpc_closed_loop_quad<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        t1::k_mpc_closed_loop_quad<true><<<grid, 32, smem, stream>>>(P);
    } else {
        const size_t smem = (size_t)t1::NF4 * 32 * sizeof(float4);
        ZB_CUDA(cudaFuncSetAttribute(t1::k_mpc_closed_loop_quad<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        t1::k_mpc_closed_loop_quad<false><<<grid, 32, smem, stream>>>(P);
    }
    ZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace zb
