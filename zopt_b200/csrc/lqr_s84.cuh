// Thread-per-problem Riccati kernel for the reference demos' shape (n, m) = (8, 4), fp32 (round 2, VERDICT r1 task 10):
//   discreteFiniteHorizonLqr  (zopt/lqrUtils.py:144-173, demos/discreteFiniteHorizonLqr.py:29-35) and
//   bilinearAffineLqr         (zopt/lqrUtils.py:207-262, demos/bilinearLqrControl.py:21-43).
//
// The compile-time-size kernels of zb_steps.cuh (k_lqr_ct / k_bilinear_ct) evaluate the recursion as written on full 8x8
// matrices with scalar global loads: 255 registers + spills, ~176 scalar loads per step, 12 % of the FP32 peak by the dense count.
// This kernel follows the (12,4) design of lqr_t1.cuh at the smaller size, where everything fits in registers:
//   * the symmetric value matrix V is its lower triangle (36 registers) for the whole horizon;
//   * pass 1: [W | VB] = V [A | B], one row of [A | B] (three 128-bit shared loads) feeding 96 FMAs; W (64) and VB (32) stay
//     in registers, nothing goes back to shared memory;
//   * G = R + B'(VB) (lower triangle), 4x4 Cholesky by reciprocal square roots;
//   * pass 3: V' (lower, 36 accumulators) = Q + A'W and M = B'W (+ H) in one pass over the rows of [A | B];
//   * L = G^-1 M by two triangular solves per column, V' -= M'L (lower): algebraically the reference's Joseph form
//     (lqrUtils.py:169) resp. V = Q + A'VA - L'S_uu L (lqrUtils.py:259), symmetric by construction;
//   * bilinear: H, d, q, r of the step are loaded from global memory straight into registers at the top of the step (they are
//     used once, so shared memory would only add a hop; their latency hides behind pass 1: staged through the slab one after
//     the other they cost 25 % of the stalls as long_scoreboard and two resident warps per SM); v + V d follows pass 1,
//     S_u = r + B'(v + V d) rides the G loop, v' = q + A'(v + V d) - S_ux' l rides pass 3.
// ~1,700 FMA and ~270 loaded words per problem-step (6 FMA per word).  Operands live in shared memory interleaved by lane
// ([float4 slot][lane]): every 128-bit access of a warp is one contiguous 512 B row (conflict-free), each lane stages and
// reads ITS OWN problem only, so there is no barrier anywhere; operands with stride_t = 0 are staged once, the others
// again at every step.  Q, R are read as symmetric (lower triangle): callers with non-symmetric weights pass
// ZB_FORCE_GENERIC and get the as-written kernels.
#pragma once
#if defined(__CUDACC__)
#include "t1_common.cuh"
#endif  // else: tests/hostsim provides float4, tri, fma2, ZB_F4, Arr, aligned16 and runs the kernel body lane by lane on the host

namespace zb {
namespace s84 {

using t1::fma2;
using t1::tri;

constexpr int RS = 32;    // row stride of the slab in float4 (one warp per CTA)
constexpr int A4 = 0;     // A rows: 8 x 2 float4
constexpr int B4 = 16;    // B rows: 8 x 1
constexpr int Q4 = 24;    // Q rows: 8 x 2 (full rows; the lower-triangle-covering chunks are read)
constexpr int R4 = 40;    // R rows: 4 x 1
constexpr int NF4 = 44;   // float4 slots per problem (704 B); the bilinear operands H, d, q, r go from global memory to registers

struct S84P {
    long long Bsz;
    int N, T;
    Arr A, B, Q, R, H, d, q, r;
    float *L, *l, *V0;  // gains (Bsz,N,4,8); bilinear offsets (Bsz,N,4); value at step 0 (Bsz,8,8) or null
};

// Each lane stages ITS OWN problem's operand block into its column of the slab: fetch() global -> registers, put() -> slab.
template <int NSLOT>
__device__ __forceinline__ void fetch(float4 (&t)[NSLOT], const float* g) {
    const float4* g4 = reinterpret_cast<const float4*>(g);
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) t[j] = __ldg(g4 + j);
}
template <int NSLOT>
__device__ __forceinline__ void put(float4* S, int slot, const float4 (&t)[NSLOT]) {
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) S[(slot + j) * RS] = t[j];
}
template <int NSLOT>
__device__ __forceinline__ void stage(float4* S, int slot, const float* g) {
    float4 t[NSLOT];
    fetch<NSLOT>(t, g);
    put<NSLOT>(S, slot, t);
}

// lower triangle of the symmetric 8x8 staged at Q4 -> v[36]
__device__ __forceinline__ void read_q_lower(const float4* S, float (&v)[36]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c <= i / 4; ++c) {
            const float4 q = S[(Q4 + 2 * i + c) * RS];
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (4 * c + e <= i) v[tri(i, 4 * c + e)] = ZB_F4(q, e);
        }
}

// One backward step for the problem owned by this thread.  v: lower triangle of V (in/out); vv: the linear term (bilinear, in/out).
template <bool BILIN>
__device__ __forceinline__ void step(const float4* S, float (&v)[36], float (&vv)[8], float (&L)[4][8], float (&l)[4], const float* gH,
                                     const float* gd, const float* gq, const float* gr) {
    float W[8][8], VB[8][4], vVd[8], dd[8], M[4][8], Su[4], qn[8];
    if (BILIN) {  // this step's H, d, q, r: global -> registers, in flight during pass 1
        const float4* h4 = reinterpret_cast<const float4*>(gH);
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const float4 h0 = __ldg(h4 + 2 * a), h1 = __ldg(h4 + 2 * a + 1);
            M[a][0] = h0.x; M[a][1] = h0.y; M[a][2] = h0.z; M[a][3] = h0.w; M[a][4] = h1.x; M[a][5] = h1.y; M[a][6] = h1.z; M[a][7] = h1.w;
        }
        const float4 d0 = __ldg(reinterpret_cast<const float4*>(gd)), d1 = __ldg(reinterpret_cast<const float4*>(gd) + 1);
        dd[0] = d0.x; dd[1] = d0.y; dd[2] = d0.z; dd[3] = d0.w; dd[4] = d1.x; dd[5] = d1.y; dd[6] = d1.z; dd[7] = d1.w;
        const float4 q0 = __ldg(reinterpret_cast<const float4*>(gq)), q1 = __ldg(reinterpret_cast<const float4*>(gq) + 1);
        qn[0] = q0.x; qn[1] = q0.y; qn[2] = q0.z; qn[3] = q0.w; qn[4] = q1.x; qn[5] = q1.y; qn[6] = q1.z; qn[7] = q1.w;
        const float4 rv = __ldg(reinterpret_cast<const float4*>(gr));
        Su[0] = rv.x; Su[1] = rv.y; Su[2] = rv.z; Su[3] = rv.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
        for (int j = 0; j < 8; ++j) W[i][j] = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) VB[i][j] = 0.f;
        vVd[i] = BILIN ? vv[i] : 0.f;
    }
    // ---- 1. [W | VB] = V [A | B]  (+ v + V d) -------------------------------------------------------------------
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const float4 a0 = S[(A4 + 2 * kk) * RS], a1 = S[(A4 + 2 * kk + 1) * RS], b4 = S[(B4 + kk) * RS];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float vik = v[tri(i, kk)];
            fma2(W[i][0], W[i][1], vik, a0.x, a0.y);
            fma2(W[i][2], W[i][3], vik, a0.z, a0.w);
            fma2(W[i][4], W[i][5], vik, a1.x, a1.y);
            fma2(W[i][6], W[i][7], vik, a1.z, a1.w);
            fma2(VB[i][0], VB[i][1], vik, b4.x, b4.y);
            fma2(VB[i][2], VB[i][3], vik, b4.z, b4.w);
        }
    }
    if (BILIN) {  // v + V d
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
#pragma unroll
            for (int i = 0; i < 8; ++i) vVd[i] = fmaf(v[tri(i, kk)], dd[kk], vVd[i]);
    }
    // ---- 2. G = R + B'(VB) (lower), S_u = r + B'(v + V d); Cholesky G = C C' ------------------------------------
    float G[10];
    {
        const float4 r0 = S[(R4 + 0) * RS], r1 = S[(R4 + 1) * RS], r2 = S[(R4 + 2) * RS], r3 = S[(R4 + 3) * RS];
        G[0] = r0.x; G[1] = r1.x; G[2] = r1.y; G[3] = r2.x; G[4] = r2.y; G[5] = r2.z; G[6] = r3.x; G[7] = r3.y; G[8] = r3.z; G[9] = r3.w;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float4 b4 = S[(B4 + i) * RS];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int c = 0; c <= a; ++c) G[tri(a, c)] = fmaf(ZB_F4(b4, a), VB[i][c], G[tri(a, c)]);
            if (BILIN) Su[a] = fmaf(ZB_F4(b4, a), vVd[i], Su[a]);
        }
    }
    // the Cholesky G = C C' (four dependent rsqrt chains) is written inside pass 3, one column of C after each of its first four rows,
    // so that the scheduler can fill the chains' latency with the pass's independent FMAs
    float d0, d1, d2, d3, c10, c20, c30, c21, c31, c32;
    // ---- 3. V' (lower) = Q + A'W, M = (H +) B'W, v' = q + A'(v + V d) in one pass over the rows of [A | B] -------
    read_q_lower(S, v);  // V is dead from here on: its registers take the new value
    if (BILIN) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[i] = qn[i];
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) M[a][j] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        const float4 a0 = S[(A4 + 2 * kk) * RS], a1 = S[(A4 + 2 * kk + 1) * RS], b4 = S[(B4 + kk) * RS];
        const float ar[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j + 1 <= i; j += 2) fma2(v[tri(i, j)], v[tri(i, j + 1)], ar[i], W[kk][j], W[kk][j + 1]);
            if ((i & 1) == 0) v[tri(i, i)] = fmaf(ar[i], W[kk][i], v[tri(i, i)]);
            if (BILIN) vv[i] = fmaf(ar[i], vVd[kk], vv[i]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 8; j += 2) fma2(M[a][j], M[a][j + 1], ZB_F4(b4, a), W[kk][j], W[kk][j + 1]);
        if (kk == 0) {
            d0 = rsqrtf(G[0]);
            c10 = G[1] * d0; c20 = G[3] * d0; c30 = G[6] * d0;
        } else if (kk == 1) {
            d1 = rsqrtf(fmaf(-c10, c10, G[2]));
            c21 = fmaf(-c20, c10, G[4]) * d1; c31 = fmaf(-c30, c10, G[7]) * d1;
        } else if (kk == 2) {
            d2 = rsqrtf(fmaf(-c21, c21, fmaf(-c20, c20, G[5])));
            c32 = fmaf(-c31, c21, fmaf(-c30, c20, G[8])) * d2;
        } else if (kk == 3) {
            d3 = rsqrtf(fmaf(-c32, c32, fmaf(-c31, c31, fmaf(-c30, c30, G[9]))));
        }
    }
    // ---- 4. L = G^-1 M (and l = G^-1 S_u): C y = rhs, C' x = y ---------------------------------------------------
#pragma unroll
    for (int j = 0; j < 8; ++j) {
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
    if (BILIN) {
        const float y0 = Su[0] * d0;
        const float y1 = fmaf(-c10, y0, Su[1]) * d1;
        const float y2 = fmaf(-c21, y1, fmaf(-c20, y0, Su[2])) * d2;
        const float y3 = fmaf(-c32, y2, fmaf(-c31, y1, fmaf(-c30, y0, Su[3]))) * d3;
        l[3] = y3 * d3;
        l[2] = fmaf(-c32, l[3], y2) * d2;
        l[1] = fmaf(-c31, l[3], fmaf(-c21, l[2], y1)) * d1;
        l[0] = fmaf(-c30, l[3], fmaf(-c20, l[2], fmaf(-c10, l[1], y0))) * d0;
    }
    // ---- 5. V' -= M'L (lower), v' -= M'l --------------------------------------------------------------------------
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float nm = -M[a][i];
#pragma unroll
            for (int j = 0; j + 1 <= i; j += 2) fma2(v[tri(i, j)], v[tri(i, j + 1)], nm, L[a][j], L[a][j + 1]);
            if ((i & 1) == 0) v[tri(i, i)] = fmaf(nm, L[a][i], v[tri(i, i)]);
            if (BILIN) vv[i] = fmaf(nm, l[a], vv[i]);
        }
    }
}

template <bool BILIN>
__global__ void __launch_bounds__(32) k_riccati_s84(S84P P) {
    __shared__ float4 slab[NF4 * RS];
    const int lane = threadIdx.x;
    long long b = blockIdx.x * 32LL + lane;
    const bool live = b < P.Bsz;
    if (!live) b = P.Bsz - 1;  // tail lanes recompute the last problem and store nothing
    float4* S = slab + lane;
    float v[36], vv[8], L[4][8], l[4];
    // terminal carry (lqrUtils.py:172 / :261): Q[T-1] (and q[T-1])
    // The terminal Q with the first step's A, then B with R: two DRAM round trips at the start of the warp instead of four (all four
    // at once would raise the kernel to 255 registers and cost a resident warp per SM)
    // (every warp of a wave starts at the same time, so nothing hides them: the staging lines held 11 % of the samples, long_scoreboard)
    int qk = P.T - 1;
    const int k0 = P.N - 1;
    {
        float4 tq[16], ta[16];
        fetch<16>(tq, P.Q.at<float>(b, qk));
        fetch<16>(ta, P.A.at<float>(b, k0));
        put<16>(S, Q4, tq);
        put<16>(S, A4, ta);
    }
    {
        float4 tb[8], tr[4];
        fetch<8>(tb, P.B.at<float>(b, k0));
        fetch<4>(tr, P.R.at<float>(b, k0));
        put<8>(S, B4, tb);
        put<4>(S, R4, tr);
    }
    read_q_lower(S, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) vv[i] = 0.f;
    if (BILIN) {
        const float4* qf = reinterpret_cast<const float4*>(P.q.at<float>(b, P.T - 1));
        const float4 q0 = __ldg(qf), q1 = __ldg(qf + 1);
        vv[0] = q0.x; vv[1] = q0.y; vv[2] = q0.z; vv[3] = q0.w; vv[4] = q1.x; vv[5] = q1.y; vv[6] = q1.z; vv[7] = q1.w;
    }
    float* Lout = P.L + b * (long long)P.N * 32;
    float* lout = BILIN ? P.l + b * (long long)P.N * 4 : nullptr;
#pragma unroll 1
    for (int k = P.N - 1; k >= 0; --k) {
        // operands constant in time are staged once (above); the others again at every step, in two groups ([A | B] and [Q, R]) whose
        // loads are all in flight before the first store waits (one L2 latency per group; re-staging a constant member of a
        // group is harmless: the same block again)
        if (k != k0 && (P.A.st || P.B.st)) {
            float4 ta[16], tb[8];
            fetch<16>(ta, P.A.at<float>(b, k));
            fetch<8>(tb, P.B.at<float>(b, k));
            put<16>(S, A4, ta);
            put<8>(S, B4, tb);
        }
        if ((k != k0 && P.R.st) || (P.Q.st && k != qk)) {
            float4 tq[16], tr[4];
            fetch<16>(tq, P.Q.at<float>(b, P.Q.st ? k : qk));
            fetch<4>(tr, P.R.at<float>(b, k));
            put<16>(S, Q4, tq);
            put<4>(S, R4, tr);
            qk = P.Q.st ? k : qk;
        }
        if (BILIN) step<true>(S, v, vv, L, l, P.H.at<float>(b, k), P.d.at<float>(b, k), P.q.at<float>(b, k), P.r.at<float>(b, k));
        else step<false>(S, v, vv, L, l, nullptr, nullptr, nullptr, nullptr);
        if (live) {
            float4* o = reinterpret_cast<float4*>(Lout + (long long)k * 32);
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                o[2 * a] = make_float4(L[a][0], L[a][1], L[a][2], L[a][3]);
                o[2 * a + 1] = make_float4(L[a][4], L[a][5], L[a][6], L[a][7]);
            }
            if (BILIN) *reinterpret_cast<float4*>(lout + (long long)k * 4) = make_float4(l[0], l[1], l[2], l[3]);
        }
    }
    if (!BILIN && P.V0 && live) {
        float* V0 = P.V0 + b * 64;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            reinterpret_cast<float4*>(V0 + i * 8)[0] = make_float4(v[tri(i, 0)], v[tri(i, 1)], v[tri(i, 2)], v[tri(i, 3)]);
            reinterpret_cast<float4*>(V0 + i * 8)[1] = make_float4(v[tri(i, 4)], v[tri(i, 5)], v[tri(i, 6)], v[tri(i, 7)]);
        }
    }
}

inline bool s84_arr_ok(const Arr& a) { return aligned16(a.p) && (a.sb % 4 == 0) && (a.st % 4 == 0); }

}  // namespace s84
}  // namespace zb
