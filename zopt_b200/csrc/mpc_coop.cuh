// Cooperative (4 threads per problem) closed-loop LQR-MPC for SMALL batches (BASELINE cfg 3 sharded over 8 GPUs leaves
// 2,048 problems per GPU: 64 one-warp CTAs of the thread-per-problem kernel, i.e. most SMs idle and a 10,000-step
// sequential chain per thread).  Splitting each problem over a quad shortens the per-step chain ~3x and uses 4x more
// warps.  The step is the one of lqr_fast.cuh (k_riccati_12x4) without the gain store; the linearisation is the
// select-and-store scheme of ilqr_fast.cuh.  Same arithmetic definition as k_mpc_closed_loop_quad (lqr_t1.cuh).
#pragma once
#include "lqr_t1.cuh"

namespace zb {

// one Riccati step on the quad's slab; returns this thread's 4x4 gain tile L[:, 4t..4t+3] (threads 0..2)
__device__ __forceinline__ void coop_riccati_step(float* Vs, const float* As, const float* Bs, const float* Qs, float* Ms, float* Gs,
                                                  const float* Rs, const float* Ct, int cstride, int tcol, int t, float (&L)[4][4]) {
    // ---- 1. [W | VB] tile = V * tile --------------------------------------------------
    float W[12][4];
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) W[i][c] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 12; ++kk) {
        const float4 c4 = lds4(Ct + kk * cstride);
        const float4 v0 = lds4(Vs + kk * 12), v1 = lds4(Vs + kk * 12 + 4), v2 = lds4(Vs + kk * 12 + 8);
        const float v[12] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w};
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            W[i][0] = fmaf(v[i], c4.x, W[i][0]);
            W[i][1] = fmaf(v[i], c4.y, W[i][1]);
            W[i][2] = fmaf(v[i], c4.z, W[i][2]);
            W[i][3] = fmaf(v[i], c4.w, W[i][3]);
        }
    }
    // ---- 2. [M | G0] tile = B^T * [W | VB] tile ---------------------------------------
    float M[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) M[a][c] = 0.f;
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        const float4 b4 = lds4(Bs + i * 4);
        const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) M[a][c] = fmaf(bb[a], W[i][c], M[a][c]);
    }
    // ---- 3. G = G0 + R ; share M and G through smem -----------------------------------
    if (t == 3) {
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float4 r4 = lds4(Rs + a * 4);
            sts4(Gs + a * 4, M[a][0] + r4.x, M[a][1] + r4.y, M[a][2] + r4.z, M[a][3] + r4.w);
        }
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) sts4(Ms + a * 12 + 4 * t, M[a][0], M[a][1], M[a][2], M[a][3]);
    }
    __syncwarp();
    // Cholesky G = C C^T (lower), d_i = 1/C_ii; every thread factors redundantly (no divergence)
    const float4 g0 = lds4(Gs), g1 = lds4(Gs + 4), g2 = lds4(Gs + 8), g3 = lds4(Gs + 12);
    const float d0 = rsqrtf(g0.x);
    const float c10 = g1.x * d0, c20 = g2.x * d0, c30 = g3.x * d0;
    const float d1 = rsqrtf(fmaf(-c10, c10, g1.y));
    const float c21 = fmaf(-c20, c10, g2.y) * d1, c31 = fmaf(-c30, c10, g3.y) * d1;
    const float d2 = rsqrtf(fmaf(-c21, c21, fmaf(-c20, c20, g2.z)));
    const float c32 = fmaf(-c31, c21, fmaf(-c30, c20, g3.z)) * d2;
    const float d3 = rsqrtf(fmaf(-c32, c32, fmaf(-c31, c31, fmaf(-c30, c30, g3.w))));
#pragma unroll
    for (int c = 0; c < 4; ++c) {
        const float y0 = M[0][c] * d0;
        const float y1 = fmaf(-c10, y0, M[1][c]) * d1;
        const float y2 = fmaf(-c21, y1, fmaf(-c20, y0, M[2][c])) * d2;
        const float y3 = fmaf(-c32, y2, fmaf(-c31, y1, fmaf(-c30, y0, M[3][c]))) * d3;
        const float x3 = y3 * d3;
        const float x2 = fmaf(-c32, x3, y2) * d2;
        const float x1 = fmaf(-c31, x3, fmaf(-c21, x2, y1)) * d1;
        const float x0 = fmaf(-c30, x3, fmaf(-c20, x2, fmaf(-c10, x1, y0))) * d0;
        L[0][c] = x0; L[1][c] = x1; L[2][c] = x2; L[3][c] = x3;
    }
    // ---- 5. V' tile = Q + A^T W - M^T L -------------------------------------------------
    float acc[12][4];
    {
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const float4 q4 = lds4(Qs + i * 12 + tcol);
            acc[i][0] = q4.x; acc[i][1] = q4.y; acc[i][2] = q4.z; acc[i][3] = q4.w;
        }
    }
#pragma unroll
    for (int kk = 0; kk < 12; ++kk) {
        const float4 a0 = lds4(As + kk * 12), a1 = lds4(As + kk * 12 + 4), a2 = lds4(As + kk * 12 + 8);
        const float a[12] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w, a2.x, a2.y, a2.z, a2.w};
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(a[i], W[kk][c], acc[i][c]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const float4 m0 = lds4(Ms + a * 12), m1 = lds4(Ms + a * 12 + 4), m2 = lds4(Ms + a * 12 + 8);
        const float mm_[12] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w, m2.x, m2.y, m2.z, m2.w};
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = fmaf(-mm_[i], L[a][c], acc[i][c]);
    }
    // write back, lower triangle wins (block (s,t) = rows 4s.., cols 4t..)
#pragma unroll
    for (int s = 0; s < 3; ++s) {
        if (s > t) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                sts4(Vs + (4 * s + r) * 12 + 4 * t, acc[4 * s + r][0], acc[4 * s + r][1], acc[4 * s + r][2], acc[4 * s + r][3]);
                sts4(Vs + (4 * t + r) * 12 + 4 * s, acc[4 * s][r], acc[4 * s + 1][r], acc[4 * s + 2][r], acc[4 * s + 3][r]);
            }
        } else if (s == t) {
            // diagonal block: entry (r,c) := lower value acc[4s+max(r,c)][min(r,c)]
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                float e[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) e[c] = (r >= c) ? acc[4 * s + r][c] : acc[4 * s + c][r];
                sts4(Vs + (4 * s + r) * 12 + 4 * s, e[0], e[1], e[2], e[3]);
            }
        }
    }
    __syncwarp();
}

__global__ void __launch_bounds__(FQ_THREADS, 3) k_mpc_closed_loop_quad_coop(t1::ClosedLoopP P) {
    extern __shared__ __align__(16) float smem[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3, quad = lane >> 2;
    const long long b_raw = ((long long)blockIdx.x * 4 + warp) * 8 + quad;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    float* S = smem + (warp * 8 + quad) * FQ_PS;
    float *Vs = S + FQ_V, *As = S + FQ_A, *Bs = S + FQ_B, *Qs = S + FQ_Q, *Ms = S + FQ_M, *Gs = S + FQ_G, *Rs = S + FQ_R;
    const float dt = P.dt;
    const unsigned FULL = 0xffffffffu;
    const int qbase = lane & ~3;
    // constant operands: Q, R, f_u
    {
        const float4* gR = reinterpret_cast<const float4*>(P.R.at<float>(b));
        const float4* gQ = reinterpret_cast<const float4*>(P.Q.at<float>(b));
#pragma unroll
        for (int i = 0; i < 9; ++i) reinterpret_cast<float4*>(Qs)[t + 4 * i] = __ldg(gQ + t + 4 * i);
        reinterpret_cast<float4*>(Rs)[t] = __ldg(gR + t);
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            const int k = t + 4 * i;  // row of f_u = dt * dF/du
            reinterpret_cast<float4*>(Bs)[k] = make_float4(k == 2 ? -dt : 0.f, k == 3 ? dt : 0.f, k == 4 ? dt : 0.f, k == 5 ? dt : 0.f);
        }
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
    const float* Ct = (t < 3) ? (As + 4 * t) : Bs;
    const int cstride = (t < 3) ? 12 : 4;
    const int tcol = (t < 3) ? 4 * t : 0;
    __syncwarp();
    for (int ts = 0; ts < P.Tsim; ++ts) {
        if (active) {
            if (t == 0) xS[(long long)ts * 3 + 0] = make_float4(x[0], x[1], x[2], x[3]);
            if (t == 1) xS[(long long)ts * 3 + 1] = make_float4(x[4], x[5], x[6], x[7]);
            if (t == 2) xS[(long long)ts * 3 + 2] = make_float4(x[8], x[9], x[10], x[11]);
        }
        // ---- linearise at (x_t, u_trim): every thread evaluates dF/dx (no divergence), stores rows 3t..3t+2 of A ----
        {
            float J[144];
            QuadTrig<float> tr = quad_trig(x);
            quad_jac_x(tr, x, ut, J);
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float row[12];
#pragma unroll
                for (int c = 0; c < 12; ++c) {
                    const float j0 = J[r * 12 + c], j1 = J[(3 + r) * 12 + c], j2 = J[(6 + r) * 12 + c], j3 = J[(9 + r) * 12 + c];
                    const float jv = (t == 0) ? j0 : (t == 1) ? j1 : (t == 2) ? j2 : j3;
                    row[c] = fmaf(dt, jv, (c == 3 * t + r) ? 1.f : 0.f);
                }
                float* dst = As + (3 * t + r) * 12;
                sts4(dst, row[0], row[1], row[2], row[3]);
                sts4(dst + 4, row[4], row[5], row[6], row[7]);
                sts4(dst + 8, row[8], row[9], row[10], row[11]);
            }
        }
        // ---- V <- Qf ----
        {
            const float4* gF = reinterpret_cast<const float4*>(P.Qf.at<float>(b));
#pragma unroll
            for (int i = 0; i < 9; ++i) reinterpret_cast<float4*>(Vs)[t + 4 * i] = __ldg(gF + t + 4 * i);
        }
        __syncwarp();
        float L[4][4];
        for (int k = P.N - 1; k >= 0; --k) coop_riccati_step(Vs, As, Bs, Qs, Ms, Gs, Rs, Ct, cstride, tcol, t, L);
        // ---- u_t = -L_0 x_t: thread t < 3 holds L[:, 4t..4t+3]; reduce the three partial products over the quad ----
        float u[4], ua[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            float s = 0.f;
            if (t < 3) {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const float xc = (t == 0) ? x[c] : (t == 1) ? x[4 + c] : x[8 + c];
                    s = fmaf(L[a][c], xc, s);
                }
            }
            // fixed summation order (tile 0 + tile 1) + tile 2, identical in every thread
            const float s0 = __shfl_sync(FULL, s, qbase + 0), s1 = __shfl_sync(FULL, s, qbase + 1), s2 = __shfl_sync(FULL, s, qbase + 2);
            u[a] = -((s0 + s1) + s2);
            ua[a] = ut[a] + u[a];
        }
        if (active && t == 3) uS[ts] = make_float4(u[0], u[1], u[2], u[3]);
        {
            float xd[12];
            QuadTrig<float> tr = quad_trig(x);
            quad_xdot(tr, x, ua, xd);
#pragma unroll
            for (int i = 0; i < 12; ++i) x[i] = fmaf(dt, xd[i], x[i]);
        }
    }
    if (active) {
        if (t == 0) xS[(long long)P.Tsim * 3 + 0] = make_float4(x[0], x[1], x[2], x[3]);
        if (t == 1) xS[(long long)P.Tsim * 3 + 1] = make_float4(x[4], x[5], x[6], x[7]);
        if (t == 2) xS[(long long)P.Tsim * 3 + 2] = make_float4(x[8], x[9], x[10], x[11]);
    }
}

inline int32_t mpc_closed_loop_coop_launch(const t1::ClosedLoopP& P, cudaStream_t stream) {
    const size_t smem = (size_t)FQ_PROBS * FQ_PS * sizeof(float);
    ZB_CUDA(cudaFuncSetAttribute(k_mpc_closed_loop_quad_coop, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const unsigned grid = (unsigned)((P.Bsz + FQ_PROBS - 1) / FQ_PROBS);
    k_mpc_closed_loop_quad_coop<<<grid, FQ_THREADS, smem, stream>>>(P);
    ZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace zb
