// The (8,4) thread-per-problem Riccati kernel of lqr_s84.cuh in fp64, the reference's own precision (jax x64):
//   discreteFiniteHorizonLqr (zopt/lqrUtils.py:144-173) and bilinearAffineLqr (zopt/lqrUtils.py:207-262) at the demos' shape.
//
// In fp64 the register-resident step of the fp32 kernel does not fit (V, W and VB alone are 264 registers), so the step is
// paneled the way lqr_t1.cuh panels (12,4): V (lower triangle, 36 doubles = 72 registers) stays in registers for the whole
// horizon, [W | VB] = V [A | B] is formed in three 8x4 panels (32 accumulators = 64 registers), the two W panels go to the
// thread's own column of the shared-memory slab and come back row by row in pass 3, where V' (lower) = Q + A'W and
// M = (H +) B'W accumulate in registers (72 + 64).  Same algebra as the fp32 kernel: G = R + B'VB, Cholesky, L = G^-1 M,
// V' -= M'L -- the reference's Joseph form (lqrUtils.py:169) resp. V = Q + A'VA - L'S_uu L (lqrUtils.py:259) up to rounding,
// symmetric by construction; gated at 1e-10 against the oracle and the reference's goldens.
// Slab: [double2 slot][lane], 106 slots = 1,696 B per problem (Q and R as lower-triangle-covering chunks), 53 KB per one-warp
// CTA, four CTAs per SM = one warp per scheduler.  The run-time-size kernel it replaces keeps everything in local memory
// (1.7 M solves/s at N=100).  Q, R are read as symmetric (lower triangle): ZB_FORCE_GENERIC selects the as-written kernels.
#pragma once
#if defined(__CUDACC__)
#include "lqr_s84.cuh"
#endif

namespace zb {
namespace s84d {

using t1::tri;

constexpr int RS = 32;
constexpr int A2 = 0;     // A rows: 8 x 4 double2
constexpr int B2 = 32;    // B rows: 8 x 2
constexpr int W2 = 48;    // W rows: 8 x 4
constexpr int Q2 = 80;    // Q: row i holds its first i/2+1 double2 (20 slots)
constexpr int R2 = 100;   // R: row a holds its first a/2+1 double2 (6 slots)
constexpr int NF2 = 106;

__host__ __device__ constexpr int qoff(int i) { return i < 2 ? i : (i < 4 ? 2 + 2 * (i - 2) : (i < 6 ? 6 + 3 * (i - 4) : 12 + 4 * (i - 6))); }
__host__ __device__ constexpr int roff(int a) { return a < 2 ? a : 2 + 2 * (a - 2); }

#define ZB_D2(v, e) ((e) == 0 ? (v).x : (v).y)

struct S84DP {
    long long Bsz;
    int N, T;
    Arr A, B, Q, R, H, d, q, r;
    double *L, *l, *V0;
};

// full ROWS x (2 C2) block: each lane stages ITS OWN problem into its column of the slab
template <int NSLOT>
__device__ __forceinline__ void stage(double2* S, int slot, const double* g) {
    const double2* g2 = reinterpret_cast<const double2*>(g);
    double2 t[NSLOT];
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) t[j] = __ldg(g2 + j);
#pragma unroll
    for (int j = 0; j < NSLOT; ++j) S[(slot + j) * RS] = t[j];
}
// the lower-triangle-covering chunks of a symmetric 8x8 (Q) / 4x4 (R)
__device__ __forceinline__ void stage_q(double2* S, const double* g) {
    const double2* g2 = reinterpret_cast<const double2*>(g);
    double2 t[20];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c <= i / 2; ++c) t[qoff(i) + c] = __ldg(g2 + i * 4 + c);
#pragma unroll
    for (int j = 0; j < 20; ++j) S[(Q2 + j) * RS] = t[j];
}
__device__ __forceinline__ void stage_r(double2* S, const double* g) {
    const double2* g2 = reinterpret_cast<const double2*>(g);
    double2 t[6];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c <= a / 2; ++c) t[roff(a) + c] = __ldg(g2 + a * 2 + c);
#pragma unroll
    for (int j = 0; j < 6; ++j) S[(R2 + j) * RS] = t[j];
}
__device__ __forceinline__ void read_q_lower(const double2* S, double (&v)[36]) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int c = 0; c <= i / 2; ++c) {
            const double2 q = S[(Q2 + qoff(i) + c) * RS];
#pragma unroll
            for (int e = 0; e < 2; ++e)
                if (2 * c + e <= i) v[tri(i, 2 * c + e)] = ZB_D2(q, e);
        }
}

template <bool BILIN>
__device__ __forceinline__ void step(double2* S, double (&v)[36], double (&vv)[8], double (&L)[4][8], double (&l)[4], const double* gH,
                                     const double* gd, const double* gq, const double* gr) {
    double vVd[8], dd[8], Su[4], G[10], M[4][8], qn[8];
    if (BILIN) {  // d and r of this step: global -> registers, in flight during the first panels
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double2 t = __ldg(reinterpret_cast<const double2*>(gd) + c);
            dd[2 * c] = t.x; dd[2 * c + 1] = t.y;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const double2 t = __ldg(reinterpret_cast<const double2*>(gr) + c);
            Su[2 * c] = t.x; Su[2 * c + 1] = t.y;
        }
    }
    // ---- 1. [W | VB] = V [A | B] in three 8x4 panels; the W panels go to the slab, the VB panel feeds G -------------------
#pragma unroll 1
    for (int p = 0; p < 3; ++p) {  // ROLLED: unrolled, the panels' accumulators overlap and spill (2.2 KB of stack)
        double acc[8][4];
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[i][c] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
            const int s0 = (p < 2) ? A2 + 4 * kk + 2 * p : B2 + 2 * kk;
            const double2 x0 = S[s0 * RS], x1 = S[(s0 + 1) * RS];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double vik = v[tri(i, kk)];
                acc[i][0] = fma(vik, x0.x, acc[i][0]);
                acc[i][1] = fma(vik, x0.y, acc[i][1]);
                acc[i][2] = fma(vik, x1.x, acc[i][2]);
                acc[i][3] = fma(vik, x1.y, acc[i][3]);
            }
        }
        if (p < 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                S[(W2 + 4 * i + 2 * p) * RS] = make_double2(acc[i][0], acc[i][1]);
                S[(W2 + 4 * i + 2 * p + 1) * RS] = make_double2(acc[i][2], acc[i][3]);
            }
        } else {
            if (BILIN) {  // v + V d
#pragma unroll
                for (int i = 0; i < 8; ++i) vVd[i] = vv[i];
#pragma unroll
                for (int kk = 0; kk < 8; ++kk)
#pragma unroll
                    for (int i = 0; i < 8; ++i) vVd[i] = fma(v[tri(i, kk)], dd[kk], vVd[i]);
            }
            // ---- 2. G = R + B'(VB) (lower), S_u = r + B'(v + V d) ---------------------------------------------------------
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c <= a / 2; ++c) {
                    const double2 r2 = S[(R2 + roff(a) + c) * RS];
#pragma unroll
                    for (int e = 0; e < 2; ++e)
                        if (2 * c + e <= a) G[tri(a, 2 * c + e)] = ZB_D2(r2, e);
                }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const double2 b0 = S[(B2 + 2 * i) * RS], b1 = S[(B2 + 2 * i + 1) * RS];
                const double ba[4] = {b0.x, b0.y, b1.x, b1.y};
#pragma unroll
                for (int a = 0; a < 4; ++a) {
#pragma unroll
                    for (int c = 0; c <= a; ++c) G[tri(a, c)] = fma(ba[a], acc[i][c], G[tri(a, c)]);
                    if (BILIN) Su[a] = fma(ba[a], vVd[i], Su[a]);
                }
            }
        }
    }
    // V is dead from here on.  H and q of this step: global -> registers, in flight during the Cholesky
    if (BILIN) {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const double2 t = __ldg(reinterpret_cast<const double2*>(gH) + a * 4 + c);
                M[a][2 * c] = t.x; M[a][2 * c + 1] = t.y;
            }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double2 t = __ldg(reinterpret_cast<const double2*>(gq) + c);
            qn[2 * c] = t.x; qn[2 * c + 1] = t.y;
        }
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) M[a][j] = 0.0;
    }
    // Cholesky G = C C' by reciprocal square roots: four dependent rsqrt chains, written INSIDE pass 3 (one column of C after each
    // of its first four rows) so that the scheduler can fill their latency with the pass's independent FMAs -- with one warp per
    // scheduler nothing else would
    // (the bilinear instantiation has no registers to spare for that -- it spilled 600 B -- and keeps the chains ahead of the pass)
    constexpr bool INTER = !BILIN;
    double d0, d1, d2, d3, c10, c20, c30, c21, c31, c32;
    if (!INTER) {
        d0 = rsqrt(G[0]);
        c10 = G[1] * d0; c20 = G[3] * d0; c30 = G[6] * d0;
        d1 = rsqrt(fma(-c10, c10, G[2]));
        c21 = fma(-c20, c10, G[4]) * d1; c31 = fma(-c30, c10, G[7]) * d1;
        d2 = rsqrt(fma(-c21, c21, fma(-c20, c20, G[5])));
        c32 = fma(-c31, c21, fma(-c30, c20, G[8])) * d2;
        d3 = rsqrt(fma(-c32, c32, fma(-c31, c31, fma(-c30, c30, G[9]))));
    }
    // ---- 3. V' (lower) = Q + A'W, M = (H +) B'W, v' = q + A'(v + V d) in one pass over the rows of A, W, B ---------------
    read_q_lower(S, v);
    if (BILIN) {
#pragma unroll
        for (int i = 0; i < 8; ++i) vv[i] = qn[i];
    }
#pragma unroll
    for (int kk = 0; kk < 8; ++kk) {
        double ar[8], wr[8], br[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double2 a2 = S[(A2 + 4 * kk + c) * RS], w2 = S[(W2 + 4 * kk + c) * RS];
            ar[2 * c] = a2.x; ar[2 * c + 1] = a2.y; wr[2 * c] = w2.x; wr[2 * c + 1] = w2.y;
        }
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const double2 b2 = S[(B2 + 2 * kk + c) * RS];
            br[2 * c] = b2.x; br[2 * c + 1] = b2.y;
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
#pragma unroll
            for (int j = 0; j <= i; ++j) v[tri(i, j)] = fma(ar[i], wr[j], v[tri(i, j)]);
            if (BILIN) vv[i] = fma(ar[i], vVd[kk], vv[i]);
        }
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int j = 0; j < 8; ++j) M[a][j] = fma(br[a], wr[j], M[a][j]);
        if (!INTER) {
        } else if (kk == 0) {
            d0 = rsqrt(G[0]);
            c10 = G[1] * d0; c20 = G[3] * d0; c30 = G[6] * d0;
        } else if (kk == 1) {
            d1 = rsqrt(fma(-c10, c10, G[2]));
            c21 = fma(-c20, c10, G[4]) * d1; c31 = fma(-c30, c10, G[7]) * d1;
        } else if (kk == 2) {
            d2 = rsqrt(fma(-c21, c21, fma(-c20, c20, G[5])));
            c32 = fma(-c31, c21, fma(-c30, c20, G[8])) * d2;
        } else if (kk == 3) {
            d3 = rsqrt(fma(-c32, c32, fma(-c31, c31, fma(-c30, c30, G[9]))));
        }
    }
    // ---- 4. L = G^-1 M (and l = G^-1 S_u): C y = rhs, C' x = y -------------------------------------------------------------
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const double y0 = M[0][j] * d0;
        const double y1 = fma(-c10, y0, M[1][j]) * d1;
        const double y2 = fma(-c21, y1, fma(-c20, y0, M[2][j])) * d2;
        const double y3 = fma(-c32, y2, fma(-c31, y1, fma(-c30, y0, M[3][j]))) * d3;
        const double x3 = y3 * d3;
        const double x2 = fma(-c32, x3, y2) * d2;
        const double x1 = fma(-c31, x3, fma(-c21, x2, y1)) * d1;
        const double x0 = fma(-c30, x3, fma(-c20, x2, fma(-c10, x1, y0))) * d0;
        L[0][j] = x0; L[1][j] = x1; L[2][j] = x2; L[3][j] = x3;
    }
    if (BILIN) {
        const double y0 = Su[0] * d0;
        const double y1 = fma(-c10, y0, Su[1]) * d1;
        const double y2 = fma(-c21, y1, fma(-c20, y0, Su[2])) * d2;
        const double y3 = fma(-c32, y2, fma(-c31, y1, fma(-c30, y0, Su[3]))) * d3;
        l[3] = y3 * d3;
        l[2] = fma(-c32, l[3], y2) * d2;
        l[1] = fma(-c31, l[3], fma(-c21, l[2], y1)) * d1;
        l[0] = fma(-c30, l[3], fma(-c20, l[2], fma(-c10, l[1], y0))) * d0;
    }
    // ---- 5. V' -= M'L (lower), v' -= M'l ------------------------------------------------------------------------------------
#pragma unroll
    for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const double nm = -M[a][i];
#pragma unroll
            for (int j = 0; j <= i; ++j) v[tri(i, j)] = fma(nm, L[a][j], v[tri(i, j)]);
            if (BILIN) vv[i] = fma(nm, l[a], vv[i]);
        }
    }
}

#if defined(__CUDACC__)
#define ZB_S84D_SLAB extern __shared__ double2 slab[]
#else
#define ZB_S84D_SLAB static double2 slab[NF2 * RS]
#endif

template <bool BILIN>
__global__ void __launch_bounds__(32) k_riccati_s84d(S84DP P) {
    ZB_S84D_SLAB;
    const int lane = threadIdx.x;
    long long b = blockIdx.x * 32LL + lane;
    const bool live = b < P.Bsz;
    if (!live) b = P.Bsz - 1;  // tail lanes recompute the last problem and store nothing
    double2* S = slab + lane;
    double v[36], vv[8], L[4][8], l[4];
    // terminal carry (lqrUtils.py:172 / :261): Q[T-1] (and q[T-1])
    int qk = P.T - 1;
    stage_q(S, P.Q.at<double>(b, qk));
    read_q_lower(S, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) vv[i] = 0.0;
    if (BILIN) {
        const double2* qf = reinterpret_cast<const double2*>(P.q.at<double>(b, P.T - 1));
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const double2 t = __ldg(qf + c);
            vv[2 * c] = t.x; vv[2 * c + 1] = t.y;
        }
    }
    double* Lout = P.L + b * (long long)P.N * 32;
    double* lout = BILIN ? P.l + b * (long long)P.N * 4 : nullptr;
#pragma unroll 1
    for (int k = P.N - 1; k >= 0; --k) {
        const bool first = (k == P.N - 1);
        if (first || P.A.st) stage<32>(S, A2, P.A.at<double>(b, k));
        if (first || P.B.st) stage<16>(S, B2, P.B.at<double>(b, k));
        if (first || P.R.st) stage_r(S, P.R.at<double>(b, k));
        if (P.Q.st && k != qk) { stage_q(S, P.Q.at<double>(b, k)); qk = k; }
        if (BILIN) step<true>(S, v, vv, L, l, P.H.at<double>(b, k), P.d.at<double>(b, k), P.q.at<double>(b, k), P.r.at<double>(b, k));
        else step<false>(S, v, vv, L, l, nullptr, nullptr, nullptr, nullptr);
        if (live) {
            double2* o = reinterpret_cast<double2*>(Lout + (long long)k * 32);
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) o[4 * a + c] = make_double2(L[a][2 * c], L[a][2 * c + 1]);
            if (BILIN) {
                double2* ol = reinterpret_cast<double2*>(lout + (long long)k * 4);
                ol[0] = make_double2(l[0], l[1]);
                ol[1] = make_double2(l[2], l[3]);
            }
        }
    }
    if (!BILIN && P.V0 && live) {
        double2* V0 = reinterpret_cast<double2*>(P.V0 + b * 64);
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) V0[i * 4 + c] = make_double2(v[tri(i, 2 * c)], v[tri(i, 2 * c + 1)]);
    }
}

inline bool arr_ok(const Arr& a) { return aligned16(a.p) && (a.sb % 2 == 0) && (a.st % 2 == 0); }

#if defined(__CUDACC__)
template <bool BILIN>
inline int32_t launch(const S84DP& F, cudaStream_t stream) {
    constexpr int smem = NF2 * RS * (int)sizeof(double2);
    // per-device attribute; cheap to repeat, so set on every call (multi-GPU callers loop devices)
    ZB_CUDA(cudaFuncSetAttribute(k_riccati_s84d<BILIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    k_riccati_s84d<BILIN><<<(unsigned)((F.Bsz + 31) / 32), 32, smem, stream>>>(F);
    ZB_CUDA(cudaGetLastError());
    return 0;
}
#endif

}  // namespace s84d
}  // namespace zb
