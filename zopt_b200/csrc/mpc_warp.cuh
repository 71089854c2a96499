// Register-tiled closed-loop LQR-MPC for SMALL per-GPU batches (BASELINE cfg 3 as stated: 16,384 problems sharded over
// 8 GPUs = 2,048 per GPU, i.e. ~14 problems per SM).  At that size the thread-per-problem kernel (lqr_t1.cuh) leaves most
// of the chip idle behind a 10,000-step sequential chain and the 4-threads-per-problem kernel (mpc_coop.cuh) still has < 2
// warps per SM, each step bound by its own instruction latency (1.7 us per Riccati step).  Here NINE lanes own a problem
// (three problems per warp) as a 3x3 grid of 4x4 tiles; lane (r,c) keeps in REGISTERS, for the whole horizon sweep,
//     Vr = V[4r..4r+3, :]   (the value matrix's rows it multiplies with)          48 words
//     Ac = A[:, 4c..4c+3],  Ar = A[:, 4r..4r+3]   (fixed during a sweep)           96 words
// so both big products of the step run without a single load:
//     W tile (r,c)  = Vr Ac                                  192 FMA
//     V' tile (r,c) = Q + Ar^T W[:, 4c..] - Mr^T Lc          192 + 64 FMA
// and only the redistribution goes through shared memory: the W tiles (each lane then reads the column block W[:, 4c..]
// and the four rows that form M) and the new V tiles (each lane reads its four rows back) -- 28 LDS.128 + 12 STS.128 per
// lane-step instead of ~700 loaded words per thread-step in the 4-thread kernel.
// The quadcopter's f_u = dt [e2 -> -1; e3, e4, e5 -> +1] (zopt/quadcopter.py:116-144) makes B'(.) a scaled row selection:
//     G = R + s s' o V[2..5, 2..5],  M = s o W[2..5, :]      (s = (-dt, dt, dt, dt))
// and the scaling can be taken out of the loop altogether: with S = diag(s), G = S (S^-1 R S^-1 + V[2..5,2..5]) S and
// M = S W[2..5,:], so L = S^-1 Lt with Lt = Gt^-1 Mt (Gt = Rt + V[2..5,2..5], Rt = S^-1 R S^-1 once per problem,
// Mt = W[2..5,:]) and M'L = Mt'Lt: no multiplication by s in the sweep, only the applied control is rescaled.
// V stays exactly symmetric ("lower triangle wins", the upper tiles are mirrored copies).  Same algebra as
// t1::riccati_step; the two kernels agree to rounding (tests: 1e-5 on 40 closed-loop steps, model evaluated by the same
// instruction sequence in both, t1_common.cuh).
// Reference loop: demos/lqrMpc.py:42-47 around zopt/mpcUtils.py:47-59 with inactive bounds; step: zopt/lqrUtils.py:167-170.
#pragma once
#include <algorithm>
#include "t1_common.cuh"
#include "quad_model_gen.cuh"

namespace zb {
namespace w9 {

constexpr int SV = 164;  // floats per problem in the V slab (12 rows of 12 + a 4-float pad after every 4 rows = 152; 164 = 4 mod 32)
constexpr int SW = 172;  // floats per problem in the W slab (144; 172 = 12 mod 32): see the bank notes at the loads
__host__ __device__ constexpr int vrow(int i) { return i * 12 + (i / 4) * 4; }  // offset of row i in the V slab
// 1/sqrt of a Cholesky pivot: the pivots are O(R + dt^2 V) > 0 and far from the subnormal range, so the bare MUFU.RSQ
// (2 ulp) replaces rsqrtf()'s range-checked sequence (5 instructions on the sweep's critical path)
__device__ __forceinline__ float rsq(float x) {
    float y;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// QUEUE (round 2): the batch does not fit a whole number of warps per scheduler.  Measured: 1,776 problems (592 warps on a B200's
// 592 schedulers) take 6.84 ms, 1,779 take 9.89 ms -- ONE scheduler with two warps sets the time for everybody, the other 591
// idle for the last third.  So the grid stays at `workers` one-warp CTAs (a multiple of the scheduler count) and the simulation
// is cut into chunks of CH steps; ready problem-triples wait in a FIFO: a worker pops a triple, advances it by one chunk and
// pushes it back.  A triple is in the queue only while nobody works on it, so no pass ever waits for another one, and the FIFO
// order makes all triples advance at the same pace: the run takes triples / workers x (time of one warp alone) instead of the
// time of a doubled-up scheduler.  The state between chunks is the trajectory row the kernel writes anyway; prog[triple] counts
// the chunks completed (release / acquire through __threadfence).  The ring has one entry per push ever made (triples x chunks),
// so it never wraps.
struct W9Queue {
    unsigned* head;  // pops so far
    unsigned* tail;  // pushes so far (entry index = number of triples + this: every triple is seeded once)
    int* ring;       // (triples * chunks) entries, -1 = not pushed yet; entry i is the i-th triple that became ready
    int* prog;       // (triples) chunks completed
    int chunk;       // simulation steps per chunk
};

// The first `n` ring entries: every triple is ready for its first chunk.  A kernel of its own, BEFORE the workers: were the workers to
// seed the ring themselves, a pop could wait for an entry owed by a CTA that is not resident yet -- with more workers than the device
// holds at once that CTA never starts (measured: a hang with three workers per scheduler).  Seeded up front, the workers form a plain
// work queue: every entry a pop waits for is owed by a RUNNING worker, whatever the grid size.
__global__ void k_queue_seed(int* ring, long long n) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j < n) ring[j] = (int)j;
}
// The whole scratch in one launch (instead of two memsets + the seed kernel: three host calls less per solve): counters and progress
// words 0, the first `n` ring entries seeded, the rest -1 (not pushed yet).
__global__ void k_queue_init(int* scratch, long long head_ints, long long n, long long total) {
    const long long j = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (j >= total) return;
    const long long r = j - head_ints;
    scratch[j] = (r < 0) ? 0 : (r < n ? (int)r : -1);
}

template <bool QUEUE>
__global__ void __launch_bounds__(32) k_mpc_closed_loop_quad_w9(t1::ClosedLoopP P, W9Queue Wq) {
    __shared__ __align__(16) float sV[3 * SV];
    __shared__ __align__(16) float sW[3 * SW];
    __shared__ __align__(16) float sS[3 * 3 * 48];  // scratch tiles of the redundant upper lanes
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const bool writer = lane < 27;           // lanes 27..31 shadow lane 26 and never store
    const int ql = writer ? lane : 26;
    const int g = ql / 9, q = ql - 9 * g;    // problem within the warp, lane within the problem
    const int r = q / 3, c = q - 3 * r;      // tile row / tile column
    const int base = 9 * g;                  // first lane of the problem
    const long long ntriples = (P.Bsz + 2) / 3;
    const int nchunks = QUEUE ? (P.Tsim + Wq.chunk - 1) / Wq.chunk : 1;
  for (;;) {  // QUEUE: one pass per popped triple; otherwise a single pass (the loop is left at its end)
    long long triple = blockIdx.x;
    int ts_begin = 0, ts_end = P.Tsim, my_chunk = 0;
    if (QUEUE) {
        unsigned h = 0;
        if (lane == 0) h = atomicAdd(Wq.head, 1u);
        h = __shfl_sync(FULL, h, 0);
        if ((long long)h >= ntriples * nchunks) return;
        // EVERY lane polls (uniform control flow): with `if (lane == 0) spin; __syncwarp();` the warp stayed split for the rest of
        // the pass and ran it at half speed (measured 17.3 against 8.4 ms at 2,048 problems)
        const volatile int* slot = Wq.ring + h;
        int j;
        while ((j = *slot) < 0) __nanosleep(100);
        __threadfence();  // acquire: the progress word and the trajectory row read below were written before the push
        triple = j;
        my_chunk = *reinterpret_cast<const volatile int*>(Wq.prog + j);
        ts_begin = my_chunk * Wq.chunk;
        ts_end = min(P.Tsim, ts_begin + Wq.chunk);
    }
    const long long b_raw = triple * 3 + g;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    float* Vs = sV + g * SV;
    float* Ws = sW + g * SW;
    // store targets, fixed for the kernel: own W tile; own V tile and its transpose, or a scratch tile for the redundant r < c lanes
    float* const Wdst = Ws + (4 * r) * 12 + 4 * c;
    const bool diag_tile = (r == c);
    float* const Vdst = (r >= c) ? Vs + vrow(4 * r) + 4 * c : sS + (g * 3 + (r + c - 1)) * 48;
    float* const Vdst_t = (r >= c) ? Vs + vrow(4 * c) + 4 * r : sS + (g * 3 + (r + c - 1)) * 48;
    const float dt = P.dt;
    const float is4[4] = {-1.f / dt, 1.f / dt, 1.f / dt, 1.f / dt};  // 1/s: f_u = dt dF/du has one entry in each of rows 2..5

    // ---- constant operands: this lane's tile of Q (lower triangle of Q wins, as t1::riccati_step reads it), R lower ----
    float Qt[4][4], Rl[10];
    {
        const float* gQ = P.Q.at<float>(b);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int gi = 4 * r + i, gj = 4 * c + j;
                Qt[i][j] = (gi >= gj) ? __ldg(gQ + gi * 12 + gj) : __ldg(gQ + gj * 12 + gi);
            }
        const float* gR = P.R.at<float>(b);
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int cc = 0; cc <= a; ++cc) Rl[t1::tri(a, cc)] = __ldg(gR + a * 4 + cc) * (is4[a] * is4[cc]);  // Rt = S^-1 R S^-1
    }
    float x[12];
    {
        // chunk 0: the initial state; later chunks: the trajectory row the previous chunk ended on (written by another worker:
        // read past L1).  Idle problem slots of the last triple keep re-reading the initial state of the problem they shadow.
        const float4* gx = (QUEUE && my_chunk > 0 && active) ? reinterpret_cast<const float4*>(P.xSim + (b * (long long)(P.Tsim + 1) + ts_begin) * 12)
                                                            : reinterpret_cast<const float4*>(P.x0 + b * 12);
        const float4 x0 = __ldcg(gx), x1 = __ldcg(gx + 1), x2 = __ldcg(gx + 2);
        x[0] = x0.x; x[1] = x0.y; x[2] = x0.z; x[3] = x0.w; x[4] = x1.x; x[5] = x1.y; x[6] = x1.z; x[7] = x1.w;
        x[8] = x2.x; x[9] = x2.y; x[10] = x2.z; x[11] = x2.w;
    }
    float4* xS = reinterpret_cast<float4*>(P.xSim + b * (long long)(P.Tsim + 1) * 12);
    float4* uS = reinterpret_cast<float4*>(P.uSim + b * (long long)P.Tsim * 4);
    const float ut[4] = {P.utrim[0], P.utrim[1], P.utrim[2], P.utrim[3]};
    const bool out_lane = active && writer && q == 0;

#pragma unroll 1
    for (int ts = ts_begin; ts < ts_end; ++ts) {
        if (out_lane) {
            xS[(long long)ts * 3 + 0] = make_float4(x[0], x[1], x[2], x[3]);
            xS[(long long)ts * 3 + 1] = make_float4(x[4], x[5], x[6], x[7]);
            xS[(long long)ts * 3 + 2] = make_float4(x[8], x[9], x[10], x[11]);
        }
        // ---- linearise at (x_t, u_trim): A = I + dt dF/dx.  Every lane of the problem evaluates the same J; the nine lanes
        //      write the same values to the (free) W slab, then each reads the two column blocks it multiplies with ----
        float Ac[12][4], Ar[12][4];
        {
            float Am[144];
            t1::closed_loop_linearize(x, ut, dt, Am);
#pragma unroll
            for (int i = 0; i < 12; ++i)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch)
                    sts4(Ws + i * 12 + 4 * ch, Am[i * 12 + 4 * ch + 0], Am[i * 12 + 4 * ch + 1], Am[i * 12 + 4 * ch + 2], Am[i * 12 + 4 * ch + 3]);
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < 12; ++k) {
            const float4 ac = lds4(Ws + k * 12 + 4 * c), ar = lds4(Ws + k * 12 + 4 * r);
            Ac[k][0] = ac.x; Ac[k][1] = ac.y; Ac[k][2] = ac.z; Ac[k][3] = ac.w;
            Ar[k][0] = ar.x; Ar[k][1] = ar.y; Ar[k][2] = ar.z; Ar[k][3] = ar.w;
        }
        __syncwarp();  // the W slab is overwritten by the first step below
        // ---- V <- Qf (lower triangle wins): rows 4r..4r+3 ----
        float Vr[4][12];
        {
            const float* gF = P.Qf.at<float>(b);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int k = 0; k < 12; ++k) {
                    const int gi = 4 * r + i;
                    Vr[i][k] = (gi >= k) ? __ldg(gF + gi * 12 + k) : __ldg(gF + k * 12 + gi);
                }
        }
        float Lc[4][4];  // gain tile Lt[:, 4c..4c+3] of the last step
        // Gt = Rt + V[2..5, 2..5] (lower triangle), carried across the steps: re-read with the lane's rows after every exchange
        float G[10];
        {
            const float* gF = P.Qf.at<float>(b);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int cc = 0; cc <= a; ++cc) G[t1::tri(a, cc)] = __ldg(gF + (2 + a) * 12 + 2 + cc) + Rl[t1::tri(a, cc)];
        }
        // Cholesky Gt = C C^T: its dependent chain (four MUFU.RSQ + ~10 FMAs) is formed as soon as Gt is known -- at the END of a
        // step, while the loads of the lane's new V rows are in flight -- instead of between the W product and the exchange
        float d0, d1, d2, d3, c10, c20, c30, c21, c31, c32;
        auto chol = [&]() {
            d0 = rsq(G[0]);
            c10 = G[1] * d0; c20 = G[3] * d0; c30 = G[6] * d0;
            d1 = rsq(fmaf(-c10, c10, G[2]));
            c21 = fmaf(-c20, c10, G[4]) * d1; c31 = fmaf(-c30, c10, G[7]) * d1;
            d2 = rsq(fmaf(-c21, c21, fmaf(-c20, c20, G[5])));
            c32 = fmaf(-c31, c21, fmaf(-c30, c20, G[8])) * d2;
            d3 = rsq(fmaf(-c32, c32, fmaf(-c31, c31, fmaf(-c30, c30, G[9]))));
        };
        chol();
#pragma unroll 1
        for (int k = P.N - 1; k >= 0; --k) {
            // ---- 1. W tile = Vr Ac: two independent accumulator sets (even / odd k) = 16 FFMA2 chains in flight ----
            float W0[4][4], W1[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) W0[i][j] = W1[i][j] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 12; kk += 2)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    t1::fma2(W0[i][0], W0[i][1], Vr[i][kk], Ac[kk][0], Ac[kk][1]);
                    t1::fma2(W0[i][2], W0[i][3], Vr[i][kk], Ac[kk][2], Ac[kk][3]);
                    t1::fma2(W1[i][0], W1[i][1], Vr[i][kk + 1], Ac[kk + 1][0], Ac[kk + 1][1]);
                    t1::fma2(W1[i][2], W1[i][3], Vr[i][kk + 1], Ac[kk + 1][2], Ac[kk + 1][3]);
                }
            // (lanes 27..31 shadow lane 26: same operands, same values, same address -- no predicate, no divergence)
#pragma unroll
            for (int i = 0; i < 4; ++i)
                sts4(Wdst + i * 12, W0[i][0] + W1[i][0], W0[i][1] + W1[i][1], W0[i][2] + W1[i][2], W0[i][3] + W1[i][3]);
            // ---- 2. (the Cholesky factor of Gt was formed at the end of the previous step, see 5.) ----
            __syncwarp();
            // ---- 3. V' tile = Q + Ar^T Wc - Mr^T Lc, again as two accumulator sets ----
            // (lanes of one problem read 3 distinct 16 B segments at 0/16/32 B; neighbouring problems are 12 banks apart)
            float V0[4][4], V1[4][4], Mc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) { V0[i][j] = Qt[i][j]; V1[i][j] = 0.f; }
#pragma unroll
            for (int kk = 0; kk < 12; kk += 2) {
                const float4 w4 = lds4(Ws + kk * 12 + 4 * c), w5 = lds4(Ws + (kk + 1) * 12 + 4 * c);
                if (kk >= 2 && kk < 6) {  // rows 2..5 of W: Mt[:, 4c..]
                    Mc[kk - 2][0] = w4.x; Mc[kk - 2][1] = w4.y; Mc[kk - 2][2] = w4.z; Mc[kk - 2][3] = w4.w;
                    Mc[kk - 1][0] = w5.x; Mc[kk - 1][1] = w5.y; Mc[kk - 1][2] = w5.z; Mc[kk - 1][3] = w5.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    t1::fma2(V0[i][0], V0[i][1], Ar[kk][i], w4.x, w4.y);
                    t1::fma2(V0[i][2], V0[i][3], Ar[kk][i], w4.z, w4.w);
                    t1::fma2(V1[i][0], V1[i][1], Ar[kk + 1][i], w5.x, w5.y);
                    t1::fma2(V1[i][2], V1[i][3], Ar[kk + 1][i], w5.z, w5.w);
                }
            }
            float Mr[4][4];  // Mt[:, 4r..4r+3]
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float4 w4 = lds4(Ws + (2 + a) * 12 + 4 * r);
                Mr[a][0] = w4.x; Mr[a][1] = w4.y; Mr[a][2] = w4.z; Mr[a][3] = w4.w;
            }
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float y0 = Mc[0][j] * d0;
                const float y1 = fmaf(-c10, y0, Mc[1][j]) * d1;
                const float y2 = fmaf(-c21, y1, fmaf(-c20, y0, Mc[2][j])) * d2;
                const float y3 = fmaf(-c32, y2, fmaf(-c31, y1, fmaf(-c30, y0, Mc[3][j]))) * d3;
                const float x3 = y3 * d3;
                const float x2 = fmaf(-c32, x3, y2) * d2;
                const float x1 = fmaf(-c31, x3, fmaf(-c21, x2, y1)) * d1;
                const float x0 = fmaf(-c30, x3, fmaf(-c20, x2, fmaf(-c10, x1, y0))) * d0;
                Lc[0][j] = x0; Lc[1][j] = x1; Lc[2][j] = x2; Lc[3][j] = x3;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                t1::fma2(V0[i][0], V0[i][1], -Mr[0][i], Lc[0][0], Lc[0][1]);
                t1::fma2(V0[i][2], V0[i][3], -Mr[0][i], Lc[0][2], Lc[0][3]);
                t1::fma2(V1[i][0], V1[i][1], -Mr[1][i], Lc[1][0], Lc[1][1]);
                t1::fma2(V1[i][2], V1[i][3], -Mr[1][i], Lc[1][2], Lc[1][3]);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                t1::fma2(V0[i][0], V0[i][1], -Mr[2][i], Lc[2][0], Lc[2][1]);
                t1::fma2(V0[i][2], V0[i][3], -Mr[2][i], Lc[2][2], Lc[2][3]);
                t1::fma2(V1[i][0], V1[i][1], -Mr[3][i], Lc[3][0], Lc[3][1]);
                t1::fma2(V1[i][2], V1[i][3], -Mr[3][i], Lc[3][2], Lc[3][3]);
            }
            float Vt[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) Vt[i][j] = V0[i][j] + V1[i][j];
            // ---- 4. write back, lower triangle wins, without divergence: tiles r > c store themselves and their transpose
            //      (= tile (c,r)); diagonal tiles store their symmetrised self twice (same values, same place); tiles r < c are
            //      redundant copies and go to a scratch slab ----
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float e[4], et[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    e[j] = (diag_tile && i < j) ? Vt[j][i] : Vt[i][j];
                    et[j] = (diag_tile && i > j) ? Vt[i][j] : Vt[j][i];
                }
                sts4(Vdst + i * 12, e[0], e[1], e[2], e[3]);
                sts4(Vdst_t + i * 12, et[0], et[1], et[2], et[3]);
            }
            __syncwarp();
            // ---- 5. this lane's rows of the new V (3 distinct row blocks per problem, 20 banks apart; problems 4 banks apart)
            //      and the block V[2..5, 2..5] of the next step's Gt (six 64-bit broadcast loads) ----
#pragma unroll
            for (int ch = 0; ch < 3; ++ch)  // chunk-major: the next step's product consumes columns 0..3 of all four rows first
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float4 v4 = lds4(Vs + vrow(4 * r + i) + 4 * ch);
                    Vr[i][4 * ch + 0] = v4.x; Vr[i][4 * ch + 1] = v4.y; Vr[i][4 * ch + 2] = v4.z; Vr[i][4 * ch + 3] = v4.w;
                }
            {
                const float2 g2 = *reinterpret_cast<const float2*>(Vs + vrow(2) + 2), g3 = *reinterpret_cast<const float2*>(Vs + vrow(3) + 2);
                const float2 g4a = *reinterpret_cast<const float2*>(Vs + vrow(4) + 2), g4b = *reinterpret_cast<const float2*>(Vs + vrow(4) + 4);
                const float2 g5a = *reinterpret_cast<const float2*>(Vs + vrow(5) + 2), g5b = *reinterpret_cast<const float2*>(Vs + vrow(5) + 4);
                G[0] = g2.x + Rl[0]; G[1] = g3.x + Rl[1]; G[2] = g3.y + Rl[2]; G[3] = g4a.x + Rl[3]; G[4] = g4a.y + Rl[4];
                G[5] = g4b.x + Rl[5]; G[6] = g5a.x + Rl[6]; G[7] = g5a.y + Rl[7]; G[8] = g5b.x + Rl[8]; G[9] = g5b.y + Rl[9];
            }
            chol();
        }
        // ---- u_t = -L_0 x_t: gather the three gain tiles from the problem's lanes (0,0), (0,1), (0,2) ----
        float u[4], ua[4];
        {
            float Lf[4][12];
#pragma unroll
            for (int cc = 0; cc < 3; ++cc)
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int j = 0; j < 4; ++j) Lf[a][4 * cc + j] = __shfl_sync(FULL, Lc[a][j], base + cc);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                float s = 0.f;
#pragma unroll
                for (int j = 0; j < 12; ++j) s = fmaf(Lf[a][j], x[j], s);
                u[a] = -s * is4[a];  // L = S^-1 Lt
                ua[a] = ut[a] + u[a];
            }
        }
        if (out_lane) uS[ts] = make_float4(u[0], u[1], u[2], u[3]);
        t1::closed_loop_plant(x, ua, dt);
    }
    if (out_lane) {  // the row after the last step of this pass: the final state, or the next chunk's starting point
        xS[(long long)ts_end * 3 + 0] = make_float4(x[0], x[1], x[2], x[3]);
        xS[(long long)ts_end * 3 + 1] = make_float4(x[4], x[5], x[6], x[7]);
        xS[(long long)ts_end * 3 + 2] = make_float4(x[8], x[9], x[10], x[11]);
    }
    if (!QUEUE) return;
    __threadfence();  // the rows above are visible before the triple is handed on
    __syncwarp();
    *reinterpret_cast<volatile int*>(Wq.prog + triple) = my_chunk + 1;  // (every lane stores the same word: no divergent tail)
    if (my_chunk + 1 < nchunks) {  // back into the queue for its next chunk
        unsigned tpos = 0;
        if (lane == 0) tpos = (unsigned)ntriples + atomicAdd(Wq.tail, 1u);  // the first `ntriples` entries are the seeds
        tpos = __shfl_sync(FULL, tpos, 0);
        __threadfence();
        *reinterpret_cast<volatile int*>(Wq.ring + tpos) = (int)triple;
    }
  }
}

}  // namespace w9

int32_t mpc_closed_loop_w9_launch(const t1::ClosedLoopP& P, cudaStream_t stream) {
    const long long triples = (P.Bsz + 2) / 3;
    int dev = 0, sms = 148;
    ZB_CUDA(cudaGetDevice(&dev));
    ZB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const long long sched = 4LL * sms;  // warp schedulers of the device
    // Work rotation (W9Queue) pays when the batch is just past a whole number m of warps per scheduler: the static mapping then runs
    // at the pace of m + 1 warps per scheduler although almost every scheduler holds m.  Measured on one B200 (200 x 50 steps):
    // 1,779 problems 10.1 -> 7.4 ms, 2,048 (cfg 3 on 8 GPUs) 10.0 -> 8.0 ms (median of 15 runs, 8.0-8.7), 2,400 10.5 -> 9.7 ms,
    // 4,096 (m = 2) 16.7 -> 12.9 ms;
    // at 3,000 (1.7 warps per scheduler) the static mapping wins (10.8 against 12.0 ms).  The run takes triples / workers x the time
    // of m warps per scheduler, + ~6 %.  ZB_W9_WORKERS_PER_SCHED overrides m (0: static mapping), ZB_W9_CHUNK the chunk length.
    // Beyond two warps per scheduler, 1,184 workers (what the device holds at once: eight 248-register warps per SM) rotating over
    // all triples still beat the hardware's own CTA dispatch: 6,000 problems 21.4 -> 18.7 ms, 8,192 26.7 -> 24.2 ms.  More workers
    // than are resident lose (8,192 with three per scheduler: 49 ms).
    int per = 0;
    if (triples > sched && 20 * triples <= 27 * sched) per = 1;
    if (triples > 2 * sched) per = 2;
    if (const char* e = getenv("ZB_W9_WORKERS_PER_SCHED")) per = atoi(e);
    if (per <= 0 || triples <= per * sched || P.Tsim < 2) {
        w9::k_mpc_closed_loop_quad_w9<false><<<(unsigned)triples, 32, 0, stream>>>(P, w9::W9Queue{nullptr, nullptr, nullptr, nullptr, 0});
        ZB_CUDA(cudaGetLastError());
        return 0;
    }
    int chunk = 5;
    if (P.Tsim > 5 * 4096) chunk = (P.Tsim + 4095) / 4096;  // the ring has one entry per (triple, chunk): keep it small for long simulations
    if (const char* e = getenv("ZB_W9_CHUNK")) chunk = atoi(e) > 0 ? atoi(e) : chunk;
    // scratch: head / tail counters, one progress word per triple, the ring; stream-ordered so that concurrent calls do not share it
    const long long nchunks = (P.Tsim + chunk - 1) / chunk;
    int* scratch = nullptr;
    const size_t n_ints = (size_t)64 + (size_t)triples + (size_t)(triples * nchunks);
    ZB_CUDA(keep_pool_memory());
    ZB_CUDA(cudaMallocAsync(reinterpret_cast<void**>(&scratch), n_ints * sizeof(int), stream));
    const w9::W9Queue Wq{reinterpret_cast<unsigned*>(scratch), reinterpret_cast<unsigned*>(scratch + 32), scratch + 64 + triples, scratch + 64, chunk};
    // `per` one-warp CTAs per scheduler; the block scheduler spreads them evenly over the SMs (measured: 592 CTAs run at the pace of
    // one warp per scheduler)
    w9::k_queue_init<<<(unsigned)((n_ints + 255) / 256), 256, 0, stream>>>(scratch, 64 + triples, triples, (long long)n_ints);
    ZB_CUDA(cudaGetLastError());
    w9::k_mpc_closed_loop_quad_w9<true><<<(unsigned)(per * sched), 32, 0, stream>>>(P, Wq);
    ZB_CUDA(cudaGetLastError());
    ZB_CUDA(cudaFreeAsync(scratch, stream));
    return 0;
}

}  // namespace zb
