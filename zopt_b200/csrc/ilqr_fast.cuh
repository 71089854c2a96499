// Cooperative iLQR backward pass for the registered quadcopter model (n=12, m=4), fp32 and fp64.
//
// One launch = backwardPass_ilqr (zopt/ilqrUtils.py:176-181, step :153-173) for every active problem, with
// the expansion of ilqrUtils.py:308-313 done on the fly: f_x = I + dt*dF/dx is evaluated analytically at
// (x_k, u_k) inside the kernel (nothing of the linearisation is materialised in HBM), c_x = (Q+Q')x_k,
// c_u = (R+R')u_k, and the conditioned Hessians come from the per-problem block Czz / Vfxx prepared once
// per solve (quadratic costs have constant Hessians).
//
// Same mapping as lqr_fast.cuh: FOUR threads per problem, [f_x | f_u] split into four 12x4 column tiles,
// operands in a per-problem shared-memory slab read with 128-bit broadcast loads.  Per step:
//   1.  [W | VB] = v_xx [f_x | f_u] and, in the same loop, [Q_x | Q_u] = [f_x | f_u]' v_x  (+ c_x, c_u)
//   2.  [M | G0] = f_u' [W | VB]         ->  Q_ux = c_ux + M,  Q_uu = c_uu + G0
//   3.  Cholesky of Q_uu;  L = -Q_uu^-1 Q_ux (own tile),  l = -Q_uu^-1 Q_u
//   4.  l_k, L_k -> global
//   5.  v_xx' = c_xx + f_x' W + Q_ux' L  (== Q_xx - L'Q_uu L),  v_x' = Q_x + Q_ux' l  (== Q_x - L'Q_uu l)
//       written back "lower triangle wins" so v_xx stays exactly symmetric.
// The algebra equals the reference's as-written expressions; parity with the oracle is gated at 1e-10 (fp64).
#pragma once
#include <algorithm>
#include "ilqr_params.cuh"

namespace zb {


// slab layout in elements of T
constexpr int IQ_V = 0;        // v_xx 12x12
constexpr int IQ_A = 144;      // f_x  12x12
constexpr int IQ_B = 288;      // f_u  12x4
constexpr int IQ_CXX = 336;    // c_xx 12x12 (conditioned)
constexpr int IQ_QS = 480;     // Q+Q' 12x12
constexpr int IQ_M = 624;      // Q_ux 4x12
constexpr int IQ_CUX = 672;    // c_ux 4x12
constexpr int IQ_G = 720;      // Q_uu 4x4
constexpr int IQ_CUU = 736;    // c_uu 4x4
constexpr int IQ_RS = 752;     // R+R' 4x4
constexpr int IQ_VX = 768;     // v_x 12 (+4 pad)
constexpr int IQ_XK = 784;     // x_k 12, u_k 4
constexpr int IQ_QU = 800;     // Q_u 4
constexpr int IQ_XK2 = 804;    // second (x_k, u_k) buffer: the rows of step k-1 land here (cp.async) while step k computes
constexpr int IQ_PS_F32 = 820;  // 820/4 = 205 odd
constexpr int IQ_PS_F64 = 822;  // 822*8/16 = 411 odd (16-byte units)
// diagonal-cost variant (Q, R, Qf diagonal => conditioned Hessians diagonal, c_ux = 0) with f_u's known sparsity:
// no B, c_xx, Q+Q', c_ux, c_uu, R+R' matrices in the slab -> 420 words, twice as many problems (warps) per SM
constexpr int ID_V = 0, ID_A = 144, ID_M = 288, ID_G = 336, ID_CXX = 352, ID_QS = 364, ID_CUU = 376, ID_RS = 380;
constexpr int ID_VX = 384, ID_XK = 400, ID_QU = 416;
constexpr int ID_XK2 = 420;
constexpr int ID_PS_F32 = 436;  // 436/4 = 109 odd
constexpr int ID_PS_F64 = 438;  // 438*8/16 = 219 odd
// DDP (diagonal-cost variant only): + 81 words eigenvector exchange + 81 words clamped block (+ pad); no second (x_k, u_k)
// buffer: the trajectory rows are prefetched through registers.
constexpr int ID_WS = 420, ID_PC = 504;   // 84-word regions (81 used)
constexpr int IDD_PS_F32 = 588;  // 588/4 = 147 odd
// fp64 DDP ("PACK"): the slab is packed so that FOUR 16-problem CTAs fit an SM instead of three (the eigen-clamp is bound by
// the latency of its rotation chain, so warps per SM is what counts): the eigenvector exchange lives in the f_x region,
// which is dead during the eigen-solve (f_x is written after it), the clamped block is a packed lower triangle (45 words),
// (x_k, u_k) share the G region (read in steps 0-1, G is written in step 3), Q_u moves into the pad of v_x.
// (In fp32 the kernel is register-limited to two 4-warp CTAs either way, and the packed variant measured 5 % slower.)
constexpr int IDD_XK = ID_G;     // (x_k | u_k), 16 words
constexpr int IDD_QU = 396;      // Q_u 4
constexpr int IDD_PC = 400;      // packed lower triangle of the clamped 9x9 block (45 words); during the eigen-solve rows 5..8 of
                                 // Z = H V (warm start) -- the rotation exchange (48 words) lives in the dead f_x region, at word 84
constexpr int IDD_PS_F64 = 446;  // 446/2 = 223 odd

// Structure of the quadcopter's dF/dx as generated in quad_model_gen.cuh::quad_jac_x: '0' structurally zero, '1' state-
// independent constant, '2' depends on (x, u).  tests/test_abi_and_host_logic.py checks this table against the generated
// source.  Only 17 of the 36 four-column chunks of f_x = I + dt dF/dx change along a trajectory; the others are stored once.
__host__ __device__ constexpr int quad_jx_kind(int i, int j) {
    constexpr char K[12][13] = {"222022020000", "222202220000", "222220220000", "000100000000", "000010000000", "000001000000",
                                "000122220000", "000022200000", "000022220000", "222000222000", "222000222000", "222000220000"};
    return K[i][j] - '0';
}
__host__ __device__ constexpr bool quad_jx_chunk_varies(int i, int q) {
    return quad_jx_kind(i, 4 * q) == 2 || quad_jx_kind(i, 4 * q + 1) == 2 || quad_jx_kind(i, 4 * q + 2) == 2 || quad_jx_kind(i, 4 * q + 3) == 2;
}

template <typename T>
struct Vec4 {
    T v[4];
};
__device__ __forceinline__ Vec4<float> ldv4(const float* p) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    return Vec4<float>{{a.x, a.y, a.z, a.w}};
}
__device__ __forceinline__ Vec4<double> ldv4(const double* p) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    return Vec4<double>{{a.x, a.y, b.x, b.y}};
}
__device__ __forceinline__ void stv4(float* p, float a, float b, float c, float d) {
    *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d);
}
__device__ __forceinline__ void stv4(double* p, double a, double b, double c, double d) {
    *reinterpret_cast<double2*>(p) = make_double2(a, b);
    *reinterpret_cast<double2*>(p + 2) = make_double2(c, d);
}
__device__ __forceinline__ float rsq(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsq(double x) { return rsqrt(x); }  // MUFU.RSQ64H + Newton steps (~1 ulp) instead of DSQRT + DDIV

// ---- cooperative 9x9 symmetric eigen-clamp for the DDP step (ilqrUtils.py:217-219 on the v_x . f_xx block) ----------
// Cyclic Jacobi with COMPILE-TIME rotation indices: the packed lower triangle A[45] is replicated in the registers of the
// four threads of a quad (every thread applies the same rotations, so no exchange is needed for A), the eigenvector
// matrix is distributed by rows (thread t holds rows t, t+4, t+8; row 8 only in thread 0).
__host__ __device__ constexpr int tri9(int i, int j) { return i >= j ? i * (i + 1) / 2 + j : j * (j + 1) / 2 + i; }

// Rotation (c, s) that annihilates a_pq.  Only the ORTHOGONALITY of the rotation has to be exact to working precision
// (A stays similar to the input whatever the angle); the angle itself only steers convergence.  So tan(theta) is formed
// in fp32 from operands scaled by a power of two (no overflow / underflow whatever the magnitudes), with approximate
// sqrt / reciprocal, and c = rsqrt(1 + t^2), s = t c are then computed at full precision.  An fp64 rotation costs ~35
// instructions this way instead of the ~110 of three IEEE divisions and two square roots.
__device__ __forceinline__ float approx_sqrt(float x) { float r; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float approx_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }

__device__ __forceinline__ void jacobi_cs(double app, double aqq, double apq, double& c, double& s) {
    const double d = aqq - app, a2 = apq + apq;
    const double mx = fmax(fabs(d), fabs(a2));
    const int ex = __double2hiint(mx) & 0x7ff00000;
    const double sc = __hiloint2double(0x7fe00000 - ex, 0);  // 2^(1023 - exponent(mx)): the larger operand lands in [1, 2)
    const float df = (float)(d * sc), af = (float)(a2 * sc);
    float tf = af * approx_rcp(fabsf(df) + approx_sqrt(fmaf(df, df, af * af)));  // = sgn(tau) / (|tau| + sqrt(1 + tau^2)), tau = d / a2
    tf = __int_as_float(__float_as_int(tf) ^ (__double2hiint(d) & 0x80000000));  // * sgn(d), branch-free
    const double t = (apq != 0.0) ? (double)tf : 0.0;
    // c = 1 / sqrt(1 + t^2), 1 + t^2 in [1, 2]: fp32 MUFU.RSQ seed (23 bits) + two Newton steps in fp64 -- branch-free and on the
    // chain's critical path ~8 FP64 instructions instead of the library rsqrt()'s range checks (6 % "branch resolving" here)
    const double xx = fma(t, t, 1.0);
    float y0f;
    asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y0f) : "f"((float)xx));
    double y = (double)y0f;
    const double hx = 0.5 * xx;
    y = y * fma(-hx * y, y, 1.5);
    y = y * fma(-hx * y, y, 1.5);
    c = y;
    s = t * c;
}
__device__ __forceinline__ void jacobi_cs(float app, float aqq, float apq, float& c, float& s) {
    const float d = aqq - app, a2 = apq + apq;
    const float mx = fmaxf(fabsf(d), fabsf(a2));
    const int ex = __float_as_int(mx) & 0x7f800000;
    const float sc = __int_as_float(0x7e800000 - ex);  // 2^(126 - exponent(mx))
    const float df = d * sc, af = a2 * sc;
    float tf = af * approx_rcp(fabsf(df) + approx_sqrt(fmaf(df, df, af * af)));
    tf = __int_as_float(__float_as_int(tf) ^ (__float_as_int(d) & 0x80000000));  // * sgn(d), branch-free
    const float t = (apq != 0.0f) ? tf : 0.0f;
    c = 1.0f / sqrtf(fmaf(t, t, 1.0f));
    s = t * c;
}

// four-way select by the thread's index in the quad, as two levels of SEL on 32-bit words (the nested ?: on doubles was
// compiled to divergent branches: 6 % of the stall samples "branch resolving")
__device__ __forceinline__ float sel4(int t, float a, float b, float c, float d) {
    const float x = (t & 1) ? b : a, y = (t & 1) ? d : c;
    return (t & 2) ? y : x;
}
__device__ __forceinline__ double sel4(int t, double a, double b, double c, double d) {
    const int xl = (t & 1) ? __double2loint(b) : __double2loint(a), xh = (t & 1) ? __double2hiint(b) : __double2hiint(a);
    const int yl = (t & 1) ? __double2loint(d) : __double2loint(c), yh = (t & 1) ? __double2hiint(d) : __double2hiint(c);
    return __hiloint2double((t & 2) ? yh : xh, (t & 2) ? yl : xl);
}

// One rotation's share of the exchange buffer: (c, s) and the rotated 2x2 block, six words (the last one pads to 16 bytes).
template <typename T>
struct Rot {
    T c, s, npp, nqq, npq;
};
__device__ __forceinline__ void st_rot(double* p, const Rot<double>& r) {
    *reinterpret_cast<double2*>(p) = make_double2(r.c, r.s);
    *reinterpret_cast<double2*>(p + 2) = make_double2(r.npp, r.nqq);
    p[4] = r.npq;
}
__device__ __forceinline__ void st_rot(float* p, const Rot<float>& r) {  // 24-byte stride: 8-byte accesses
    *reinterpret_cast<float2*>(p) = make_float2(r.c, r.s);
    *reinterpret_cast<float2*>(p + 2) = make_float2(r.npp, r.nqq);
    p[4] = r.npq;
}
__device__ __forceinline__ Rot<double> ld_rot(const double* p) {
    const double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    return Rot<double>{a.x, a.y, b.x, b.y, p[4]};
}
__device__ __forceinline__ Rot<float> ld_rot(const float* p) {
    const float2 a = *reinterpret_cast<const float2*>(p), b = *reinterpret_cast<const float2*>(p + 2);
    return Rot<float>{a.x, a.y, b.x, b.y, p[4]};
}

// The off-diagonal part of rotation (PP,QQ) on the replicated matrix and on this thread's eigenvector rows, FOLLOWED BY THE
// SWAP of p and q (odd-even ordering, see jacobi_round): column P receives s p + c q, column Q receives c p - s q.  The 2x2
// block itself arrives ready-made (and already swapped) from the thread that computed the rotation.
template <typename T, int PP, int QQ>
__device__ __forceinline__ void jacobi_apply_off(T (&A)[45], T (&Wr)[3][9], const Rot<T>& r) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        if (k != PP && k != QQ) {
            const T akp = A[tri9(k, PP)], akq = A[tri9(k, QQ)];
            A[tri9(k, PP)] = r.s * akp + r.c * akq;
            A[tri9(k, QQ)] = r.c * akp - r.s * akq;
        }
    }
    A[tri9(PP, PP)] = r.npp;
    A[tri9(QQ, QQ)] = r.nqq;
    A[tri9(QQ, PP)] = r.npq;
#pragma unroll
    for (int w = 0; w < 3; ++w) {
        const T wp = Wr[w][PP], wq = Wr[w][QQ];
        Wr[w][PP] = r.s * wp + r.c * wq;
        Wr[w][QQ] = r.c * wp - r.s * wq;
    }
}

// One round of four disjoint rotations.  ODD-EVEN ORDERING (Brent-Luk's transposition scheme): even rounds pair the
// neighbours (0,1)(2,3)(4,5)(6,7), odd rounds (1,2)(3,4)(5,6)(7,8), and every rotation also SWAPS its two indices -- the swap is
// a relabelling of the outputs, so it costs nothing.  Nine such rounds are an odd-even transposition sort that reverses
// the order of the nine indices, so every pair meets exactly once: a cyclic sweep with only TWO code blocks and no data
// movement.  (Round 1 of this kernel used the tournament ring: three unrolled rounds + a 72-register permutation per loop
// trip, 48 moves per round.  Measured on cfg 5 trajectories both orderings need the same number of rounds: 25.2 vs 25.1.)
// V max(L, eps) V' does not depend on the order of the eigenpairs, so the accumulated permutation is never undone.
// Thread t of the quad computes rotation t of the round AND its rotated 2x2 block (general formulas: a_pq does not vanish
// exactly because tan was approximate -- it shrinks by ~1e-7 per visit, and quadratically once small, like the exact
// rotation; rounding noise of the annihilated entry is dropped once it is below one ulp of the diagonal, so a converged
// matrix has exact zeros and the loop terminates on its off-diagonal test).  The four (c, s, block) go through shared
// memory; every thread then applies the four rotations to the rest of its replica (round 1 computed the four 2x2 blocks in
// every thread: 68 of ~260 instructions per round, now 17).
template <typename T, int ODD>
__device__ __forceinline__ void jacobi_round(T (&A)[45], T (&Wr)[3][9], T* buf /* 24 words, this quad, this round parity */, int t, unsigned qmask) {
    constexpr int p1 = ODD, q1 = ODD + 1, p2 = ODD + 2, q2 = ODD + 3, p3 = ODD + 4, q3 = ODD + 5, p4 = ODD + 6, q4 = ODD + 7;
    const T app = sel4(t, A[tri9(p1, p1)], A[tri9(p2, p2)], A[tri9(p3, p3)], A[tri9(p4, p4)]);
    const T aqq = sel4(t, A[tri9(q1, q1)], A[tri9(q2, q2)], A[tri9(q3, q3)], A[tri9(q4, q4)]);
    const T apq = sel4(t, A[tri9(q1, p1)], A[tri9(q2, p2)], A[tri9(q3, p3)], A[tri9(q4, p4)]);
    Rot<T> r;
    jacobi_cs(app, aqq, apq, r.c, r.s);
    {
        const T cc = r.c * r.c, ss = r.s * r.s, cs = r.c * r.s;
        const T x2 = (cs + cs) * apq;
        r.nqq = fma(cc, app, fma(ss, aqq, -x2));  // the rotated a_pp lands in slot q (swap) ...
        r.npp = fma(ss, app, fma(cc, aqq, x2));   // ... and the rotated a_qq in slot p
        const T npq = fma(cs, app - aqq, (cc - ss) * apq);
        const T tiny = (sizeof(T) == 8 ? T(8.9e-16) : T(4.8e-7)) * (fabs(r.npp) + fabs(r.nqq));
        r.npq = fabs(npq) <= tiny ? T(0) : npq;
    }
    st_rot(buf + 6 * t, r);
    __syncwarp(qmask);
    const Rot<T> r1 = ld_rot(buf), r2 = ld_rot(buf + 6), r3 = ld_rot(buf + 12), r4 = ld_rot(buf + 18);
    jacobi_apply_off<T, p1, q1>(A, Wr, r1);
    jacobi_apply_off<T, p2, q2>(A, Wr, r2);
    jacobi_apply_off<T, p3, q3>(A, Wr, r3);
    jacobi_apply_off<T, p4, q4>(A, Wr, r4);
}

// Warm start of the eigen-solve (round 2).  Along a trajectory H_k changes little from one step to the next, so the
// eigenvectors V of step k+1 nearly diagonalise H_k: the sweep starts from A' = V' H V (off-diagonal norm ~1e-3 of the
// diagonal instead of ~1) and the rotations are accumulated onto V, which halves the number of rounds (measured on cfg 5
// trajectories: 48 -> 25 per step).  V is exactly orthogonal in working precision (a product of rotations), so A' is
// similar to H to rounding.  The product is DISTRIBUTED over the quad and exchanged through shared memory:
//   pass 1  Z = H V      thread t forms rows t, t+4 (t = 0: also row 8) of Z            3 x 81 FMA
//   pass 2  A' = V' Z    thread t forms rows t, t+4 (and 8) of A'                       3 x 81 FMA
// then the lower triangle of A' is gathered back into every thread's replica, and the thread's rows of V seed the
// eigenvector accumulator.  EV: V (9x9, row-major); Z0 (45 words) | Z1 (36 words): rows 0-4 | 5-8 of Z; Ag: 45 words.
template <typename T>
__device__ __forceinline__ void jacobi_warm_start(T (&A)[45], T (&Wr)[3][9], const T* EV, T* Z0, T* Z1, T* Ag, int t, unsigned qmask) {
    const int r0 = t, r1 = t + 4;
    T acc[3][9];
    {
        T h[3][9];
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            h[0][j] = sel4(t, A[tri9(0, j)], A[tri9(1, j)], A[tri9(2, j)], A[tri9(3, j)]);
            h[1][j] = sel4(t, A[tri9(4, j)], A[tri9(5, j)], A[tri9(6, j)], A[tri9(7, j)]);
            h[2][j] = A[tri9(8, j)];
        }
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int b = 0; b < 9; ++b) acc[r][b] = T(0);
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            T v[9];
#pragma unroll
            for (int b = 0; b < 9; ++b) v[b] = EV[j * 9 + b];
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int b = 0; b < 9; ++b) acc[r][b] = fma(h[r][j], v[b], acc[r][b]);
        }
    }
    {
        T* z0 = Z0 + 9 * r0;                              // rows 0..3
        T* z1 = (t == 0) ? Z0 + 36 : Z1 + 9 * (t - 1);    // rows 4..7
#pragma unroll
        for (int b = 0; b < 9; ++b) { z0[b] = acc[0][b]; z1[b] = acc[1][b]; }
        if (t == 0) {
#pragma unroll
            for (int b = 0; b < 9; ++b) Z1[27 + b] = acc[2][b];  // row 8
        }
    }
    __syncwarp(qmask);
#pragma unroll
    for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int b = 0; b < 9; ++b) acc[r][b] = T(0);
#pragma unroll
    for (int i = 0; i < 9; ++i) {
        const T* zr = (i < 5) ? Z0 + 9 * i : Z1 + 9 * (i - 5);
        T z[9];
#pragma unroll
        for (int b = 0; b < 9; ++b) z[b] = zr[b];
        const T v0 = EV[i * 9 + r0], v1 = EV[i * 9 + r1], v2 = EV[i * 9 + 8];
#pragma unroll
        for (int b = 0; b < 9; ++b) {
            acc[0][b] = fma(v0, z[b], acc[0][b]);
            acc[1][b] = fma(v1, z[b], acc[1][b]);
            acc[2][b] = fma(v2, z[b], acc[2][b]);
        }
    }
    {
        const int o0 = r0 * (r0 + 1) / 2, o1 = r1 * (r1 + 1) / 2;
#pragma unroll
        for (int b = 0; b < 9; ++b) {
            if (b <= r0) Ag[o0 + b] = acc[0][b];
            if (b <= r1) Ag[o1 + b] = acc[1][b];
            if (t == 0) Ag[36 + b] = acc[2][b];
        }
    }
    __syncwarp(qmask);
#pragma unroll
    for (int e = 0; e < 45; ++e) A[e] = Ag[e];
#pragma unroll
    for (int c = 0; c < 9; ++c) {
        Wr[0][c] = EV[r0 * 9 + c];
        Wr[1][c] = EV[r1 * 9 + c];
        Wr[2][c] = EV[72 + c];  // row 8 (kept by thread 0 only)
    }
    __syncwarp(qmask);  // Ag / EV may be overwritten from here on (rotation exchange, eigenvector exchange)
}

// f_x = I + dt dF/dx into the slab: row i is written by thread i / 3 of the quad (divergent on purpose: each entry of dF/dx is
// then evaluated once per quad by the thread that stores it, instead of by all four followed by a four-way select).
// ALL = every chunk (once, before the sweep: the state-independent entries), otherwise only the chunks that vary.
template <typename T, bool ALL>
__device__ __forceinline__ void store_fx(T* As, const T* J, T dt, int t) {
#pragma unroll
    for (int i = 0; i < 12; ++i) {
        if (t == i / 3) {
#pragma unroll
            for (int q = 0; q < 3; ++q) {
                if (ALL || quad_jx_chunk_varies(i, q)) {
                    T e[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) e[c] = dt * J[i * 12 + 4 * q + c] + ((4 * q + c == i) ? T(1) : T(0));
                    stv4(As + i * 12 + 4 * q, e[0], e[1], e[2], e[3]);
                }
            }
        }
    }
}

__device__ __forceinline__ void bw_cp16(void* dst_smem, const void* src) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(src) : "memory");
}
// this thread's quarter (4 words) of (x_k | u_k): x_k[4t..4t+3] for t < 3, u_k for t = 3
template <typename T>
__device__ __forceinline__ void stage_xu(T* dst16, const T* xT, const T* uT, int k, int t) {
    const T* src = (t < 3) ? xT + (long long)k * 12 + 4 * t : uT + (long long)k * 4;
    bw_cp16(dst16 + 4 * t, src);
    if (sizeof(T) == 8) bw_cp16(dst16 + 4 * t + 2, src + 2);
    asm volatile("cp.async.commit_group;\n" ::: "memory");
}

template <typename T, bool CDIAG, bool DDP>
__global__ void __launch_bounds__(256) k_ilqr_backward_quad(IlqrFastP P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    static_assert(!DDP || CDIAG, "the DDP fast path exists for diagonal costs only");
    constexpr int PS = DDP ? (sizeof(T) == 4 ? IDD_PS_F32 : IDD_PS_F64)
                           : CDIAG ? (sizeof(T) == 4 ? ID_PS_F32 : ID_PS_F64) : (sizeof(T) == 4 ? IQ_PS_F32 : IQ_PS_F64);
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3, quad = lane >> 2;
    const long long slot = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * 8 + quad;
    const bool in_range = slot < (P.act.perm ? (long long)*P.act.count : P.Bsz);
    const long long b = P.act.perm ? (in_range ? (long long)P.act.perm[slot] : 0) : (in_range ? slot : P.Bsz - 1);
    const bool active = in_range && !(P.done && P.done[b]);
    // a warp whose 8 problems are all frozen / out of range has nothing to do
    if (__ballot_sync(0xffffffffu, active) == 0u) {
        if (DDP && blockDim.x > 32)  // the DDP CTA meets at a barrier every step (see the loop): idle warps keep the count right
            for (int k = P.N - 1; k >= 0; --k) __syncthreads();
        return;
    }
    T* S = smem + (warp * 8 + quad) * PS;
    T *Vs = S + (CDIAG ? ID_V : IQ_V), *As = S + (CDIAG ? ID_A : IQ_A), *Bs = S + IQ_B, *Cxx = S + (CDIAG ? ID_CXX : IQ_CXX);
    T *Qs = S + (CDIAG ? ID_QS : IQ_QS), *Ms = S + (CDIAG ? ID_M : IQ_M), *Cux = S + IQ_CUX, *Gs = S + (CDIAG ? ID_G : IQ_G);
    T *Cuu = S + (CDIAG ? ID_CUU : IQ_CUU), *Rs = S + (CDIAG ? ID_RS : IQ_RS), *vx = S + (CDIAG ? ID_VX : IQ_VX);
    constexpr bool PACK = DDP && sizeof(T) == 8;  // packed DDP slab (see IDD_*)
    T *xk0 = S + (PACK ? IDD_XK : CDIAG ? ID_XK : IQ_XK), *Qu = S + (PACK ? IDD_QU : CDIAG ? ID_QU : IQ_QU);  // Bs, Cux unused when CDIAG
    constexpr bool STAGE = !DDP;  // (x_k, u_k) through a two-slot cp.async buffer; the DDP slab has no room for the second slot
    constexpr int XK2 = CDIAG ? (ID_XK2 - ID_XK) : (IQ_XK2 - IQ_XK);
    const int N = P.N;
    const T dt = T(P.dt);
    const T* xT = reinterpret_cast<const T*>(P.xTraj) + b * (long long)(N + 1) * 12;
    const T* uT = reinterpret_cast<const T*>(P.uTraj) + b * (long long)N * 4;
    T* lo = reinterpret_cast<T*>(P.l) + b * (long long)N * 4;
    T* Lo = reinterpret_cast<T*>(P.L) + b * (long long)N * 48;

    // ---- one-time staging: conditioned cost blocks, Q+Q', R+R', terminal value, constant f_u ----
    {
        const T* Czz = reinterpret_cast<const T*>(P.Czz) + b * 256;
        const T* Vf = reinterpret_cast<const T*>(P.Vfxx) + b * 144;
        const T* Q = P.C.Q.at<T>(b);
        const T* R = P.C.R.at<T>(b);
        const T* Qf = P.C.Qf.at<T>(b);
        if (CDIAG) {
            for (int e = t; e < 144; e += 4) Vs[e] = Vf[e];
            for (int i = t; i < 12; i += 4) { Cxx[i] = Czz[i * 17]; Qs[i] = T(2) * Q[i * 13]; }
            Cuu[t] = Czz[(12 + t) * 17];
            Rs[t] = T(2) * R[t * 5];
        } else {
            for (int e = t; e < 144; e += 4) {
                const int i = e / 12, j = e % 12;
                Cxx[e] = Czz[i * 16 + j];
                Qs[e] = Q[i * 12 + j] + Q[j * 12 + i];
                Vs[e] = Vf[e];
            }
            for (int e = t; e < 48; e += 4) {
                const int a = e / 12, j = e % 12;
                Cux[e] = Czz[(12 + a) * 16 + j];
                Bs[e] = T(0);
            }
            for (int e = t; e < 16; e += 4) {
                const int a = e / 4, c = e % 4;
                Cuu[e] = Czz[(12 + a) * 16 + 12 + c];
                Rs[e] = R[a * 4 + c] + R[c * 4 + a];
            }
            __syncwarp();
            if (t == 0) {  // f_u = dt * dF/du: [2,0] = -dt, [3,1] = [4,2] = [5,3] = +dt (quad_model_gen.cuh)
                Bs[2 * 4 + 0] = -dt; Bs[3 * 4 + 1] = dt; Bs[4 * 4 + 2] = dt; Bs[5 * 4 + 3] = dt;
            }
        }
        {   // the state-independent entries of f_x (zeros, 1 + dt * constants): dF/dx at the origin, folded at compile time
            const T z12[12] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)}, z4[4] = {T(0), T(0), T(0), T(0)};
            QuadTrig<T> tr0;
            tr0.sph = T(0); tr0.cph = T(1); tr0.sth = T(0); tr0.cth = T(1); tr0.sps = T(0); tr0.cps = T(1); tr0.tth = T(0); tr0.sec = T(1);
            T J0[144];
            quad_jac_x(tr0, z12, z4, J0);
            store_fx<T, true>(As, J0, dt, t);
        }
        // v_x(N) = (Qf + Qf') x_N
        const T* xN = xT + (long long)N * 12;
        for (int i = t; i < 12; i += 4) {
            T s = T(0);
            for (int j = 0; j < 12; ++j) s += (Qf[i * 12 + j] + Qf[j * 12 + i]) * xN[j];
            vx[i] = s;
        }
    }
    __syncwarp();

    const T* Ct = (t < 3) ? (As + 4 * t) : (CDIAG ? As : Bs);  // CDIAG: thread 3's rows of f_u are formed arithmetically
    const int cstride = (t < 3) ? 12 : (CDIAG ? 12 : 4);
    const int tcol = (t < 3) ? 4 * t : 0;

    // trajectory rows are fetched one step ahead: cp.async into the other (x_k, u_k) slot (iLQR), or through registers (DDP)
    Vec4<T> px0, px1, px2, pu0;
    if (STAGE) {
        stage_xu(xk0 + ((N - 1) & 1) * XK2, xT, uT, N - 1, t);
    } else {
        px0 = ldv4(xT + (long long)(N - 1) * 12); px1 = ldv4(xT + (long long)(N - 1) * 12 + 4); px2 = ldv4(xT + (long long)(N - 1) * 12 + 8);
        pu0 = ldv4(uT + (long long)(N - 1) * 4);
    }
    const int qbase = lane & 28;
    // DDP warm start: fp32 keeps the eigenvectors of the previous step in its slab; the packed fp64 slab sends them through
    // a per-problem global scratch (648 bytes, L2-resident) and copies them back into the dead f_x region at the top of a step
    T* evg = (DDP && P.ev) ? reinterpret_cast<T*>(P.ev) + b * 84 : nullptr;
    const bool warm_ok = DDP && !P.cold_start && (!PACK || P.ev != nullptr);
    for (int k = N - 1; k >= 0; --k) {
        T* xk = STAGE ? xk0 + (k & 1) * XK2 : xk0;
        // The DDP step is ~8,000 instructions (128 KB) of mostly straight-line code, far beyond the instruction cache: warps that
        // drift apart each stream it from L2 on their own (ncu: 50 % of the stall samples "no instruction").  Meeting once per
        // step keeps the warps of a CTA within one step of each other, so that they share the fetched lines.
        // (The iLQR instantiations are a third of the size and run 10-20 % SLOWER with the barrier: lockstep warps contend for the
        // same pipe at the same time.)
        // (Measured alternatives on cfg 5, 66.6 ms as is: a barrier every second / fourth step 68.8 / 70.2 ms; a split-phase barrier
        // -- two mbarriers, a warp enters step s once every warp has entered step s-1 -- 79 ms: one step of slack already spreads
        // the warps over the 128 KB body; a second barrier between the eigen-solve and the Riccati step: no change.)
        if (DDP && blockDim.x > 32) __syncthreads();
        // warm start of this step's eigen-solve -- except on every 64th step, which starts cold: the carried eigenvectors are a
        // product of thousands of rotations and their orthogonality drifts by ~1e-16 per rotation (1e-13 after 100 steps, measured);
        // the restart bounds that whatever the horizon, for one cold solve (46 rounds instead of 25) in 64 steps
        const bool warm_k = warm_ok && k < N - 1 && ((N - 1 - k) & 63) != 0;
        if (PACK && warm_k) {  // 81 words = 41 16-byte chunks (the scratch row is 84 words), 10-11 per thread
#pragma unroll
            for (int c = 0; c < 11; ++c) {
                const int ch = t + 4 * c;
                if (ch < 41) bw_cp16(As + 2 * ch, evg + 2 * ch);
            }
            asm volatile("cp.async.commit_group;\n" ::: "memory");
        }
        // ---- 0. linearise at (x_k, u_k) ----
        {
            T x[12], u[4], J[144];
            if (STAGE) {
                asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                __syncwarp();  // row k has landed for the whole quad; the other slot (read during step k+1) is free
                stage_xu(xk0 + ((k - 1) & 1) * XK2, xT, uT, k > 0 ? k - 1 : 0, t);
                const Vec4<T> x0 = ldv4(xk), x1 = ldv4(xk + 4), x2 = ldv4(xk + 8), u0 = ldv4(xk + 12);
#pragma unroll
                for (int i = 0; i < 4; ++i) { x[i] = x0.v[i]; x[4 + i] = x1.v[i]; x[8 + i] = x2.v[i]; u[i] = u0.v[i]; }
            } else {
                const Vec4<T> x0 = px0, x1 = px1, x2 = px2, u0 = pu0;
                {
                    const int kp = (k > 0) ? k - 1 : 0;
                    px0 = ldv4(xT + (long long)kp * 12); px1 = ldv4(xT + (long long)kp * 12 + 4); px2 = ldv4(xT + (long long)kp * 12 + 8);
                    pu0 = ldv4(uT + (long long)kp * 4);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) { x[i] = x0.v[i]; x[4 + i] = x1.v[i]; x[8 + i] = x2.v[i]; u[i] = u0.v[i]; }
                const Vec4<T> mine = (t == 0) ? x0 : (t == 1) ? x1 : (t == 2) ? x2 : u0;  // x_k quarter / u_k
                stv4(xk + 4 * t, mine.v[0], mine.v[1], mine.v[2], mine.v[3]);
            }
            // sin / cos of the three Euler angles: one angle per thread of the quad (thread 3 idles along with psi), exchanged
            // with shuffles -- the same sincos() result whichever thread evaluates it
            QuadTrig<T> tr;
            {
                const T ang = (t == 0) ? x[6] : (t == 1) ? x[7] : x[8];
                T sv, cv;
                sincos(ang, &sv, &cv);
                tr.sph = __shfl_sync(0xffffffffu, sv, qbase); tr.cph = __shfl_sync(0xffffffffu, cv, qbase);
                tr.sth = __shfl_sync(0xffffffffu, sv, qbase + 1); tr.cth = __shfl_sync(0xffffffffu, cv, qbase + 1);
                tr.sps = __shfl_sync(0xffffffffu, sv, qbase + 2); tr.cps = __shfl_sync(0xffffffffu, cv, qbase + 2);
                tr.sec = T(1) / tr.cth;
                tr.tth = tr.sth * tr.sec;
            }
            if (!PACK) {
                quad_jac_x(tr, x, u, J);
                store_fx<T, false>(As, J, dt, t);
            }
            if (DDP) {
                // conditionQuadraticDynamics (ilqrUtils.py:237-251): H = dt * sum_i v_x[i] d2F_i/dx2 touches states 0..8 only,
                // f_ux = f_uu = 0, so clampPD(blockdiag(H9, 0)) = blockdiag(clampPD(H9), eps I) exactly.
                T* Wsm = PACK ? As : S + ID_WS;  // PACK: the f_x region is dead until the eigen-solve is over
                T* Pc = S + (PACK ? IDD_PC : ID_PC);
                T lam[12];
#pragma unroll
                for (int i = 0; i < 12; ++i) lam[i] = vx[i];
                T A9[45], Wr[3][9];
                {
                    T h9[81];
                    quad_hess_contract(tr, x, u, lam, h9);  // lower triangle of the 9x9
#pragma unroll
                    for (int i = 0; i < 9; ++i)
#pragma unroll
                        for (int j = 0; j <= i; ++j) A9[tri9(i, j)] = dt * h9[i * 9 + j];
                }
                const unsigned qmask = 0xFu << (lane & 28);
                // scratch of the eigen-solve, all of it dead at this point of the step: the rotation exchange (48 words) and the
                // gathered A' share one region; Z = H V lives where Q_ux and the clamped block will be written later
                T* csbuf = PACK ? As + 84 : Ms;
                if (warm_k) {  // eigenvectors of step k+1 are in Wsm (PACK: just copied back from global)
                    if (PACK) {
                        asm volatile("cp.async.wait_group 0;\n" ::: "memory");
                        __syncwarp(qmask);
                    }
                    jacobi_warm_start<T>(A9, Wr, Wsm, PACK ? Ms : Pc, PACK ? Pc : Pc + 45, csbuf, t, qmask);
                } else {
#pragma unroll
                    for (int r = 0; r < 3; ++r)
#pragma unroll
                        for (int c = 0; c < 9; ++c) Wr[r][c] = (c == t + 4 * r) ? T(1) : T(0);  // rows t, t+4, t+8 of I
                }
                // convergence is tested before every pair of rounds (an even and an odd one; a sweep is nine rounds)
#pragma unroll 1
                for (int g = 0; g < 135; ++g) {
                    T o4[4] = {T(0), T(0), T(0), T(0)}, d2[2] = {T(0), T(0)};  // partial sums: the 36-term chain was 70 % "wait"
#pragma unroll
                    for (int i = 0; i < 9; ++i) {
                        d2[i & 1] = fma(A9[tri9(i, i)], A9[tri9(i, i)], d2[i & 1]);
#pragma unroll
                        for (int j = 0; j < i; ++j) o4[tri9(i, j) & 3] = fma(A9[tri9(i, j)], A9[tri9(i, j)], o4[tri9(i, j) & 3]);
                    }
                    const T off = (o4[0] + o4[1]) + (o4[2] + o4[3]), dg = d2[0] + d2[1];
                    const T thr = (sizeof(T) == 8) ? T(1e-29) : T(1e-15), tiny = (sizeof(T) == 8) ? T(1e-300) : T(1e-37);
                    if (!(off > thr * dg) || off < tiny) break;  // converged -- or NaN (a diverged problem must not hold its CTA for 270 rounds a step)
                    jacobi_round<T, 0>(A9, Wr, csbuf, t, qmask);
                    jacobi_round<T, 1>(A9, Wr, csbuf + 24, t, qmask);  // alternate buffers: a round's stores never race the previous round's loads
                }
                // exchange eigenvector rows, then P9 = W max(Lambda, eps) W^T, rows t, t+4, t+8 per thread
                const T eps = T(P.eps);
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int row = t + 4 * r;
                    if (row < 9) {
#pragma unroll
                        for (int c = 0; c < 9; ++c) Wsm[row * 9 + c] = Wr[r][c];
                        if (PACK && warm_ok && active && k > 0) {  // the packed slab has no room to keep them: via global (L2) to step k-1
#pragma unroll
                            for (int c = 0; c < 9; ++c) evg[row * 9 + c] = Wr[r][c];
                        }
                    }
                }
                __syncwarp();
                T lamc[9];
#pragma unroll
                for (int k2 = 0; k2 < 9; ++k2) lamc[k2] = A9[tri9(k2, k2)] > eps ? A9[tri9(k2, k2)] : eps;
#pragma unroll
                for (int r = 0; r < 3; ++r) {
                    const int row = t + 4 * r;
                    if (row < 9) {
                        T wl[9];
#pragma unroll
                        for (int k2 = 0; k2 < 9; ++k2) wl[k2] = Wr[r][k2] * lamc[k2];
#pragma unroll
                        for (int j = 0; j < 9; ++j) {  // lower triangle only (the block is symmetric); fixed trip count: nine independent chains
                            T acc9 = T(0);
#pragma unroll
                            for (int k2 = 0; k2 < 9; ++k2) acc9 = fma(wl[k2], Wsm[j * 9 + k2], acc9);
                            if (j <= row) Pc[PACK ? row * (row + 1) / 2 + j : row * 9 + j] = acc9;
                        }
                    }
                }
                if (PACK) {
                    __syncwarp();  // every thread is done reading the eigenvectors: the f_x region can be rewritten
                    // f_x = I + dt dF/dx at (x_k, u_k), every chunk (the exchange overwrote the state-independent ones); the state
                    // is re-read from the slab rather than kept in registers across the eigen-solve
                    const Vec4<T> x0 = ldv4(xk), x1 = ldv4(xk + 4), x2 = ldv4(xk + 8), u0 = ldv4(xk + 12);
#pragma unroll
                    for (int i = 0; i < 4; ++i) { x[i] = x0.v[i]; x[4 + i] = x1.v[i]; x[8 + i] = x2.v[i]; u[i] = u0.v[i]; }
                    quad_jac_x(tr, x, u, J);
                    store_fx<T, true>(As, J, dt, t);
                }
                // visibility of Pc and f_x for the steps below is ensured by the __syncwarp() that follows
            }
        }
        __syncwarp();
        // ---- 1. [W | VB] tile = v_xx * tile ;  [Q_x | Q_u] tile = tile' v_x + c_x / c_u -------------
        T W[12][4];
        T qv[4] = {T(0), T(0), T(0), T(0)};
#pragma unroll
        for (int i = 0; i < 12; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) W[i][c] = T(0);
#pragma unroll
        for (int kk = 0; kk < 12; ++kk) {
            Vec4<T> c4 = ldv4(Ct + kk * cstride);
            if (CDIAG && t == 3)  // row kk of f_u = dt*dF/du: (-dt at [2,0]; +dt at [3,1], [4,2], [5,3])
                c4 = Vec4<T>{{kk == 2 ? -dt : T(0), kk == 3 ? dt : T(0), kk == 4 ? dt : T(0), kk == 5 ? dt : T(0)}};
            const Vec4<T> v0 = ldv4(Vs + kk * 12), v1 = ldv4(Vs + kk * 12 + 4), v2 = ldv4(Vs + kk * 12 + 8);
            const T vxk = vx[kk];
#pragma unroll
            for (int c = 0; c < 4; ++c) qv[c] = fma(c4.v[c], vxk, qv[c]);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    W[i][c] = fma(v0.v[i], c4.v[c], W[i][c]);
                    W[4 + i][c] = fma(v1.v[i], c4.v[c], W[4 + i][c]);
                    W[8 + i][c] = fma(v2.v[i], c4.v[c], W[8 + i][c]);
                }
        }
        // c_x tile = (Q+Q')[4t..4t+3, :] x_k   /   c_u = (R+R') u_k  (thread 3)
        if (CDIAG) {
            const Vec4<T> xq = ldv4(xk + 4 * t);                 // x_k[4t..4t+3]  (thread 3: u_k)
            const Vec4<T> dq = (t < 3) ? ldv4(Qs + 4 * t) : ldv4(Rs);
#pragma unroll
            for (int c = 0; c < 4; ++c) qv[c] = fma(dq.v[c], xq.v[c], qv[c]);
        } else if (t < 3) {
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const T* qrow = Qs + (4 * t + c) * 12;
                const Vec4<T> q0 = ldv4(qrow), q1 = ldv4(qrow + 4), q2 = ldv4(qrow + 8);
                const Vec4<T> a0 = ldv4(xk), a1 = ldv4(xk + 4), a2 = ldv4(xk + 8);
                T s = T(0);
#pragma unroll
                for (int j = 0; j < 4; ++j) { s = fma(q0.v[j], a0.v[j], s); s = fma(q1.v[j], a1.v[j], s); s = fma(q2.v[j], a2.v[j], s); }
                qv[c] += s;
            }
        } else {
            const Vec4<T> uu = ldv4(xk + 12);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const Vec4<T> r4 = ldv4(Rs + c * 4);
                T s = T(0);
#pragma unroll
                for (int j = 0; j < 4; ++j) s = fma(r4.v[j], uu.v[j], s);
                qv[c] += s;
            }
        }
        // ---- 2. [M | G0] tile = f_u' * [W | VB] tile ------------------------------------------
        T M[4][4];
        if (CDIAG) {  // f_u' X = rows 2..5 of X scaled by (-dt, dt, dt, dt): no loads, no FMAs
#pragma unroll
            for (int c = 0; c < 4; ++c) { M[0][c] = -dt * W[2][c]; M[1][c] = dt * W[3][c]; M[2][c] = dt * W[4][c]; M[3][c] = dt * W[5][c]; }
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a)
#pragma unroll
                for (int c = 0; c < 4; ++c) M[a][c] = T(0);
#pragma unroll
            for (int i = 0; i < 12; ++i) {
                const Vec4<T> b4 = ldv4(Bs + i * 4);
#pragma unroll
                for (int a = 0; a < 4; ++a)
#pragma unroll
                    for (int c = 0; c < 4; ++c) M[a][c] = fma(b4.v[a], W[i][c], M[a][c]);
            }
        }
        // ---- 3. Q_uu = c_uu + G0 (thread 3), Q_ux tile = c_ux tile + M (threads 0..2); share through smem
        if (PACK) __syncwarp();  // (x_k, u_k) share the G region: every thread of the quad has read them (steps 0-1)
        if (t == 3) {
#pragma unroll
            const Vec4<T> cd = ldv4(Cuu);  // CDIAG: the diagonal of c_uu
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                Vec4<T> c4;
                if (CDIAG) c4 = Vec4<T>{{a == 0 ? cd.v[0] : T(0), a == 1 ? cd.v[1] : T(0), a == 2 ? cd.v[2] : T(0), a == 3 ? cd.v[3] : T(0)}};
                else c4 = ldv4(Cuu + a * 4);
                if (DDP) c4.v[a] += T(P.eps);  // vf_uu = eps I (ilqrUtils.py:197)
                stv4(Gs + a * 4, M[a][0] + c4.v[0], M[a][1] + c4.v[1], M[a][2] + c4.v[2], M[a][3] + c4.v[3]);
            }
            stv4(Qu, qv[0], qv[1], qv[2], qv[3]);
        } else {
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                if (!CDIAG) {
                    const Vec4<T> c4 = ldv4(Cux + a * 12 + 4 * t);
#pragma unroll
                    for (int c = 0; c < 4; ++c) M[a][c] += c4.v[c];
                }
                stv4(Ms + a * 12 + 4 * t, M[a][0], M[a][1], M[a][2], M[a][3]);
            }
        }
        __syncwarp();
        const Vec4<T> g0 = ldv4(Gs), g1 = ldv4(Gs + 4), g2 = ldv4(Gs + 8), g3 = ldv4(Gs + 12);
        const T d0 = rsq(g0.v[0]);
        const T c10 = g1.v[0] * d0, c20 = g2.v[0] * d0, c30 = g3.v[0] * d0;
        const T d1 = rsq(fma(-c10, c10, g1.v[1]));
        const T c21 = fma(-c20, c10, g2.v[1]) * d1, c31 = fma(-c30, c10, g3.v[1]) * d1;
        const T d2 = rsq(fma(-c21, c21, fma(-c20, c20, g2.v[2])));
        const T c32 = fma(-c31, c21, fma(-c30, c20, g3.v[2])) * d2;
        const T d3 = rsq(fma(-c32, c32, fma(-c31, c31, fma(-c30, c30, g3.v[3]))));
        T L[4][5];  // columns 0..3: -Q_uu^-1 Q_ux tile ; column 4: l = -Q_uu^-1 Q_u
        const Vec4<T> qu = ldv4(Qu);
#pragma unroll
        for (int c = 0; c < 5; ++c) {
            const T m0 = (c < 4) ? M[0][c] : qu.v[0], m1 = (c < 4) ? M[1][c] : qu.v[1];
            const T m2 = (c < 4) ? M[2][c] : qu.v[2], m3 = (c < 4) ? M[3][c] : qu.v[3];
            const T y0 = m0 * d0;
            const T y1 = fma(-c10, y0, m1) * d1;
            const T y2 = fma(-c21, y1, fma(-c20, y0, m2)) * d2;
            const T y3 = fma(-c32, y2, fma(-c31, y1, fma(-c30, y0, m3))) * d3;
            const T x3 = y3 * d3;
            const T x2 = fma(-c32, x3, y2) * d2;
            const T x1 = fma(-c31, x3, fma(-c21, x2, y1)) * d1;
            const T x0 = fma(-c30, x3, fma(-c20, x2, fma(-c10, x1, y0))) * d0;
            L[0][c] = -x0; L[1][c] = -x1; L[2][c] = -x2; L[3][c] = -x3;
        }
        // ---- 4. policy -> global --------------------------------------------------------------
        if (active) {
            if (t < 3) {
                T* g = Lo + (long long)k * 48 + 4 * t;
#pragma unroll
                for (int a = 0; a < 4; ++a) stv4(g + a * 12, L[a][0], L[a][1], L[a][2], L[a][3]);
            } else {
                stv4(lo + (long long)k * 4, L[0][4], L[1][4], L[2][4], L[3][4]);
            }
        }
        // ---- 5. v_x' tile = Q_x + Q_ux' l ; v_xx' tile = c_xx + f_x' W + Q_ux' L  (two half-tiles of rows) ----
        if (t < 3) {
            T nv[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) nv[c] = qv[c] + (M[0][c] * L[0][4] + M[1][c] * L[1][4] + M[2][c] * L[2][4] + M[3][c] * L[3][4]);
            stv4(vx + 4 * t, nv[0], nv[1], nv[2], nv[3]);
        }
        // v_xx' is symmetric: of the 9 4x4 blocks only 6 are distinct.  Thread t computes (t, t) and ((t+1)%3, t) -- for t = 2
        // that is the UPPER block (0, 2), stored together with its transpose (2, 0) -- so every thread does 2 blocks, not 3.
        // (Round 1's DDP instantiation kept a 3-block loop with a compile-time row block, 153 vs 197 ms then; with the eigen-solve
        // restructured in round 2 the two-block form wins there too: cfg 5 69.6 -> 67.2 ms.  ZB_DDP_NBLK=3 restores it.)
#ifndef ZB_DDP_NBLK
#define ZB_DDP_NBLK 2
#endif
        constexpr bool B3 = DDP && ZB_DDP_NBLK == 3;
        constexpr int NBLK = B3 ? 3 : 2;
#pragma unroll
        for (int bi = 0; bi < NBLK; ++bi) {
            const int tb = t < 3 ? t : 0;  // thread 3 (the f_u tile) owns no block: it shadows thread 0 and stores nothing
            const int sblk = B3 ? bi : ((bi == 0) ? tb : (tb == 2 ? 0 : tb + 1));  // rows 4*sblk .. 4*sblk+3 of the tile = block (sblk, t)
            T acc[4][4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                if (CDIAG) {  // c_xx diagonal: only the diagonal block (sblk == t) carries it
                    const Vec4<T> cd4 = ldv4(Cxx + tcol);
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = (sblk == t && i == c) ? cd4.v[c] : T(0);
                    if (DDP) {  // + vf_xx = blockdiag(clampPD(H9), eps I3), symmetric: use the lower value for both halves
                        const int gi = 4 * sblk + i;
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int gj = tcol + c;
                            const int hi_ = gi > gj ? gi : gj, lo_ = gi > gj ? gj : gi;
                            T add = T(0);
                            if (hi_ < 9) add = PACK ? (S + IDD_PC)[hi_ * (hi_ + 1) / 2 + lo_] : (S + ID_PC)[hi_ * 9 + lo_];
                            else if (gi == gj) add = T(P.eps);
                            acc[i][c] += add;
                        }
                    }
                } else {
                    const Vec4<T> q4 = ldv4(Cxx + (4 * sblk + i) * 12 + tcol);
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = q4.v[c];
                }
            }
#pragma unroll
            for (int kk = 0; kk < 12; ++kk) {
                const Vec4<T> a4 = ldv4(As + kk * 12 + 4 * sblk);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = fma(a4.v[i], W[kk][c], acc[i][c]);
            }
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const Vec4<T> m4 = ldv4(Ms + a * 12 + 4 * sblk);
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int c = 0; c < 4; ++c) acc[i][c] = fma(m4.v[i], L[a][c], acc[i][c]);
            }
            // write back, lower triangle wins: strictly-lower blocks are stored with their transpose,
            // diagonal blocks are mirrored, upper blocks are dropped (their transposes are authoritative)
            if (B3 ? (sblk > t) : (t < 3 && sblk != t)) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    stv4(Vs + (4 * sblk + r) * 12 + 4 * t, acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                    stv4(Vs + (4 * t + r) * 12 + 4 * sblk, acc[0][r], acc[1][r], acc[2][r], acc[3][r]);
                }
            } else if (B3 ? (sblk == t) : (t < 3)) {
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    T e[4];
#pragma unroll
                    for (int c = 0; c < 4; ++c) e[c] = (r >= c) ? acc[r][c] : acc[c][r];
                    stv4(Vs + (4 * sblk + r) * 12 + 4 * sblk, e[0], e[1], e[2], e[3]);
                }
            }
        }
        __syncwarp();
    }
}

// launch the cooperative backward pass
template <typename T, bool CDIAG, bool DDP = false>
inline int32_t ilqr_fast_launch_impl(const IlqrFastP& P, cudaStream_t stream) {
    constexpr int PS = DDP ? (sizeof(T) == 4 ? IDD_PS_F32 : IDD_PS_F64)
                           : CDIAG ? (sizeof(T) == 4 ? ID_PS_F32 : ID_PS_F64) : (sizeof(T) == 4 ? IQ_PS_F32 : IQ_PS_F64);
    // problems per CTA: fp32 32 (4 warps); fp64 16 (2 warps) so several CTAs share an SM's shared memory
    int warps = (sizeof(T) == 4) ? 4 : 2;
    // Wave balance.  A full SM holds maxw warps of this kernel (shared memory; 8 by registers).  When the batch needs more than
    // one wave of the machine, ONE CTA per SM is launched with the smallest warp count that keeps the number of waves, so that the
    // last wave is as full as the first (16,384 problems in fp64: 293 CTAs of 7 warps = 1.98 waves instead of 1.73 -> 2; cfg 4
    // 35.5 -> 33.1 ms, cfg 5 71.8 -> 70.4 ms).  A batch below one wave keeps small CTAs (fp32: 4 warps, fp64: 2) spread over the SMs.
    const long long nsm = 148;
    const int maxw = (int)std::min<size_t>(8, (size_t)232448 / ((size_t)8 * PS * sizeof(T)));
    const long long nprob = (P.act.perm && P.active_hint > 0 && P.active_hint < P.Bsz) ? P.active_hint : P.Bsz;  // slots in use
    auto waves = [&](int w) { const long long ctas = (nprob + 8 * w - 1) / (8 * w); return (ctas + nsm - 1) / nsm; };
    if (DDP) {
        // DDP: ONE CTA per SM whose warps meet at a barrier every step (see the kernel), also below one wave
        const int min_w = waves(maxw) == 1 ? 1 : (maxw + 1) / 2;
        warps = maxw;
        while (warps > min_w && waves(warps - 1) == waves(maxw)) --warps;
        const char* e = getenv("ZB_DDP_WARPS");
        if (e && atoi(e) >= 1 && atoi(e) <= maxw) warps = atoi(e);
    } else if (sizeof(T) == 8 && maxw >= 2 && waves(maxw) > 1) {
        warps = maxw;
        while (warps > (maxw + 1) / 2 && waves(warps - 1) == waves(maxw)) --warps;
    }
    size_t smem = (size_t)warps * 8 * PS * sizeof(T);
    // fp32 DDP: two 4-warp CTAs per SM run faster than three (measured 65 vs 79 ms on cfg 5: the ~1,100-instruction Jacobi loop
    // of three CTAs in different phases thrashes the instruction cache), so the request is padded past a third of the SM
    if (DDP && sizeof(T) == 4 && smem < 80 * 1024) smem = 80 * 1024;
    const unsigned grid = (unsigned)((nprob + warps * 8 - 1) / (warps * 8));
    ZB_CUDA(cudaFuncSetAttribute(k_ilqr_backward_quad<T, CDIAG, DDP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    IlqrFastP Pl = P;
    if (DDP && getenv("ZB_DDP_COLD_START")) Pl.cold_start = 1;
    k_ilqr_backward_quad<T, CDIAG, DDP><<<grid, warps * 32, smem, stream>>>(Pl);
    ZB_CUDA(cudaGetLastError());
    return 0;
}

int32_t ilqr_fast_launch(int32_t dtype, const IlqrFastP& P, cudaStream_t stream, bool cost_diagonal, bool second_order) {
    if (second_order)  // eligibility (diagonal costs) is checked by the caller
        return dtype == ZB_F32 ? ilqr_fast_launch_impl<float, true, true>(P, stream) : ilqr_fast_launch_impl<double, true, true>(P, stream);
    if (dtype == ZB_F32) return cost_diagonal ? ilqr_fast_launch_impl<float, true>(P, stream) : ilqr_fast_launch_impl<float, false>(P, stream);
    return cost_diagonal ? ilqr_fast_launch_impl<double, true>(P, stream) : ilqr_fast_launch_impl<double, false>(P, stream);
}

}  // namespace zb
