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
// which is exactly what the thread-per-problem kernel's FMA chains evaluate (products with B's zeros are exact zeros), so
// the arithmetic -- operation order included -- is the one of t1::riccati_step and the two variants agree bit for bit;
// V stays exactly symmetric ("lower triangle wins", the upper tiles are mirrored copies).
// Reference loop: demos/lqrMpc.py:42-47 around zopt/mpcUtils.py:47-59 with inactive bounds; step: zopt/lqrUtils.py:167-170.
#pragma once
#include "t1_common.cuh"
#include "quad_model_gen.cuh"

namespace zb {
namespace w9 {

constexpr int SV = 164;  // floats per problem in the V slab (12 rows of 12 + a 4-float pad after every 4 rows = 152; 164 = 4 mod 32)
constexpr int SW = 172;  // floats per problem in the W slab (144; 172 = 12 mod 32): see the bank notes at the loads
__host__ __device__ constexpr int vrow(int i) { return i * 12 + (i / 4) * 4; }  // offset of row i in the V slab

__global__ void __launch_bounds__(32) k_mpc_closed_loop_quad_w9(t1::ClosedLoopP P) {
    __shared__ __align__(16) float sV[3 * SV];
    __shared__ __align__(16) float sW[3 * SW];
    const unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x;
    const bool writer = lane < 27;           // lanes 27..31 shadow lane 26 and never store
    const int ql = writer ? lane : 26;
    const int g = ql / 9, q = ql - 9 * g;    // problem within the warp, lane within the problem
    const int r = q / 3, c = q - 3 * r;      // tile row / tile column
    const int base = 9 * g;                  // first lane of the problem
    const long long b_raw = (long long)blockIdx.x * 3 + g;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    float* Vs = sV + g * SV;
    float* Ws = sW + g * SW;
    const float dt = P.dt;
    const float s4[4] = {-dt, dt, dt, dt};   // f_u = dt dF/du: rows 2..5, one entry each

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
            for (int cc = 0; cc <= a; ++cc) Rl[t1::tri(a, cc)] = __ldg(gR + a * 4 + cc);
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
    const bool out_lane = active && writer && q == 0;

#pragma unroll 1
    for (int ts = 0; ts < P.Tsim; ++ts) {
        if (out_lane) {
            xS[(long long)ts * 3 + 0] = make_float4(x[0], x[1], x[2], x[3]);
            xS[(long long)ts * 3 + 1] = make_float4(x[4], x[5], x[6], x[7]);
            xS[(long long)ts * 3 + 2] = make_float4(x[8], x[9], x[10], x[11]);
        }
        // ---- linearise at (x_t, u_trim): A = I + dt dF/dx.  Every lane of the problem evaluates the same J; the nine lanes
        //      write the same values to the (free) W slab, then each reads the two column blocks it multiplies with ----
        float Ac[12][4], Ar[12][4];
        {
            float J[144];
            QuadTrig<float> tr = quad_trig(x);
            quad_jac_x(tr, x, ut, J);
            if (writer) {
#pragma unroll
                for (int i = 0; i < 12; ++i)
#pragma unroll
                    for (int ch = 0; ch < 3; ++ch)
                        sts4(Ws + i * 12 + 4 * ch, fmaf(dt, J[i * 12 + 4 * ch + 0], (i == 4 * ch + 0) ? 1.f : 0.f),
                             fmaf(dt, J[i * 12 + 4 * ch + 1], (i == 4 * ch + 1) ? 1.f : 0.f),
                             fmaf(dt, J[i * 12 + 4 * ch + 2], (i == 4 * ch + 2) ? 1.f : 0.f),
                             fmaf(dt, J[i * 12 + 4 * ch + 3], (i == 4 * ch + 3) ? 1.f : 0.f));
            }
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
        float Lc[4][4];  // gain tile L[:, 4c..4c+3] of the last step
#pragma unroll 1
        for (int k = P.N - 1; k >= 0; --k) {
            // ---- G = R + B'VB = R + s s' o V[2..5, 2..5]: rows 2,3 live in the problem's lane (0,0), rows 4,5 in lane (1,0) ----
            float G[10];
            {
                const float v22 = __shfl_sync(FULL, Vr[2][2], base), v32 = __shfl_sync(FULL, Vr[3][2], base), v33 = __shfl_sync(FULL, Vr[3][3], base);
                const float v42 = __shfl_sync(FULL, Vr[0][2], base + 3), v43 = __shfl_sync(FULL, Vr[0][3], base + 3), v44 = __shfl_sync(FULL, Vr[0][4], base + 3);
                const float v52 = __shfl_sync(FULL, Vr[1][2], base + 3), v53 = __shfl_sync(FULL, Vr[1][3], base + 3), v54 = __shfl_sync(FULL, Vr[1][4], base + 3),
                            v55 = __shfl_sync(FULL, Vr[1][5], base + 3);
                const float vv[10] = {v22, v32, v33, v42, v43, v44, v52, v53, v54, v55};
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int cc = 0; cc <= a; ++cc)  // same rounding as the FMA chains of t1::riccati_step: (V s_c) first, then s_a (.) + R
                        G[t1::tri(a, cc)] = fmaf(s4[a], __fmul_rn(vv[t1::tri(a, cc)], s4[cc]), Rl[t1::tri(a, cc)]);
            }
            // ---- 1. W tile = Vr Ac ----
            float Wt[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) Wt[i][j] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 12; ++kk)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    t1::fma2(Wt[i][0], Wt[i][1], Vr[i][kk], Ac[kk][0], Ac[kk][1]);
                    t1::fma2(Wt[i][2], Wt[i][3], Vr[i][kk], Ac[kk][2], Ac[kk][3]);
                }
            if (writer) {
#pragma unroll
                for (int i = 0; i < 4; ++i) sts4(Ws + (4 * r + i) * 12 + 4 * c, Wt[i][0], Wt[i][1], Wt[i][2], Wt[i][3]);
            }
            // ---- 2. Cholesky G = C C^T (overlaps the exchange) ----
            const float d0 = rsqrtf(G[0]);
            const float c10 = G[1] * d0, c20 = G[3] * d0, c30 = G[6] * d0;
            const float d1 = rsqrtf(fmaf(-c10, c10, G[2]));
            const float c21 = fmaf(-c20, c10, G[4]) * d1, c31 = fmaf(-c30, c10, G[7]) * d1;
            const float d2 = rsqrtf(fmaf(-c21, c21, fmaf(-c20, c20, G[5])));
            const float c32 = fmaf(-c31, c21, fmaf(-c30, c20, G[8])) * d2;
            const float d3 = rsqrtf(fmaf(-c32, c32, fmaf(-c31, c31, fmaf(-c30, c30, G[9]))));
            __syncwarp();
            // ---- 3. V' tile = Q + Ar^T Wc - Mr^T Lc ----
            // (lanes of one problem read 3 distinct 16 B segments at 0/16/32 B; neighbouring problems are 12 banks apart)
            float Vt[4][4], Mc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) Vt[i][j] = Qt[i][j];
#pragma unroll
            for (int kk = 0; kk < 12; ++kk) {
                const float4 w4 = lds4(Ws + kk * 12 + 4 * c);
                if (kk >= 2 && kk < 6) {  // rows 2..5 of W: M[:, 4c..] = s o W[2..5, 4c..]
                    Mc[kk - 2][0] = __fmul_rn(s4[kk - 2], w4.x); Mc[kk - 2][1] = __fmul_rn(s4[kk - 2], w4.y);
                    Mc[kk - 2][2] = __fmul_rn(s4[kk - 2], w4.z); Mc[kk - 2][3] = __fmul_rn(s4[kk - 2], w4.w);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    t1::fma2(Vt[i][0], Vt[i][1], Ar[kk][i], w4.x, w4.y);
                    t1::fma2(Vt[i][2], Vt[i][3], Ar[kk][i], w4.z, w4.w);
                }
            }
            float Mr[4][4];  // M[:, 4r..4r+3]
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const float4 w4 = lds4(Ws + (2 + a) * 12 + 4 * r);
                Mr[a][0] = __fmul_rn(s4[a], w4.x); Mr[a][1] = __fmul_rn(s4[a], w4.y);
                Mr[a][2] = __fmul_rn(s4[a], w4.z); Mr[a][3] = __fmul_rn(s4[a], w4.w);
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
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    t1::fma2(Vt[i][0], Vt[i][1], -Mr[a][i], Lc[a][0], Lc[a][1]);
                    t1::fma2(Vt[i][2], Vt[i][3], -Mr[a][i], Lc[a][2], Lc[a][3]);
                }
            // ---- 4. write back, lower triangle wins: tile (r,c) with r > c also stores its transpose as tile (c,r) ----
            if (writer) {
                if (r > c) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        sts4(Vs + vrow(4 * r + i) + 4 * c, Vt[i][0], Vt[i][1], Vt[i][2], Vt[i][3]);
                        sts4(Vs + vrow(4 * c + i) + 4 * r, Vt[0][i], Vt[1][i], Vt[2][i], Vt[3][i]);
                    }
                } else if (r == c) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float e[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) e[j] = (i >= j) ? Vt[i][j] : Vt[j][i];
                        sts4(Vs + vrow(4 * r + i) + 4 * r, e[0], e[1], e[2], e[3]);
                    }
                }
            }
            __syncwarp();
            // ---- 5. this lane's rows of the new V (3 distinct row blocks per problem, 20 banks apart; problems 4 banks apart) ----
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) {
                    const float4 v4 = lds4(Vs + vrow(4 * r + i) + 4 * ch);
                    Vr[i][4 * ch + 0] = v4.x; Vr[i][4 * ch + 1] = v4.y; Vr[i][4 * ch + 2] = v4.z; Vr[i][4 * ch + 3] = v4.w;
                }
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
                u[a] = -s;
                ua[a] = ut[a] + u[a];
            }
        }
        if (out_lane) uS[ts] = make_float4(u[0], u[1], u[2], u[3]);
        {
            float xd[12];
            QuadTrig<float> tr = quad_trig(x);
            quad_xdot(tr, x, ua, xd);
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = fmaf(dt, xd[i], x[i]);
        }
    }
    if (out_lane) {
        xS[(long long)P.Tsim * 3 + 0] = make_float4(x[0], x[1], x[2], x[3]);
        xS[(long long)P.Tsim * 3 + 1] = make_float4(x[4], x[5], x[6], x[7]);
        xS[(long long)P.Tsim * 3 + 2] = make_float4(x[8], x[9], x[10], x[11]);
    }
}

}  // namespace w9

int32_t mpc_closed_loop_w9_launch(const t1::ClosedLoopP& P, cudaStream_t stream) {
    const unsigned grid = (unsigned)((P.Bsz + 2) / 3);
    w9::k_mpc_closed_loop_quad_w9<<<grid, 32, 0, stream>>>(P);
    ZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace zb
