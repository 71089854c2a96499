// Cooperative (n,m) = (12,4) Riccati recursion in fp64 -- the reference itself computes in fp64; fp32 has the thread-per-
// problem kernels of lqr_t1.cuh (the templates below are written over T, only <double> is instantiated):
// discreteFiniteHorizonLqr (zopt/lqrUtils.py:144-173), the unconstrained lqrMpc.solve (zopt/mpcUtils.py:47-81) and the
// closed-loop LQR-MPC of BASELINE cfg 3 (demos/lqrMpc.py:42-47 on the nonlinear, re-linearised quadcopter).
//
// Same mapping as the iLQR backward kernel (ilqr_fast.cuh): FOUR threads per problem, [A | B] split into four 12x4 column
// tiles, V / A / B in a per-problem shared-memory slab read with 128-bit broadcast loads (an fp64 value matrix does not
// fit one thread's registers).  Per step, for thread t (tile t: columns 4t..4t+3 of A, t = 3: B):
//   1.  [W | VB] tile = V * tile
//   2.  [M | G0] tile = B' * [W | VB] tile           (M = B'VA, G0 = B'VB)
//   3.  G = R + G0 (thread 3) -> shared; Cholesky of G by every thread; L tile = G^-1 M tile
//   4.  L_k tile -> global
//   5.  V' = Q + A'W - M'L, two 4x4 blocks per thread, written back "lower triangle wins" (V stays exactly symmetric)
// V' = Q + A'VA - M'G^-1 M is the reference's Joseph form (lqrUtils.py:169) in exact arithmetic; parity with the fp64
// oracle is gated at 1e-10.  Q_k and R_k are read from global memory where they are needed (once per step, L1-resident
// when time-invariant; with ZB_COST_DIAGONAL their diagonals sit in the slab), so the slab is 418 words and 8 warps fit an SM in fp64.  A_k, B_k that really vary along the
// horizon are re-staged every step (synchronously: the streamed fp32 kernel of lqr_t1.cuh is the tuned path for that).
#pragma once
#include "ilqr_fast.cuh"

namespace zb {

constexpr int LQ_V = 0, LQ_A = 144, LQ_B = 288, LQ_M = 336, LQ_G = 384;
constexpr int LQ_QD = 400;      // diagonal-cost variants: diag(Q) 12 words, diag(R) 4 words
constexpr int LQ_PS_F32 = 420;  // 420/4 = 105 odd (16-byte units)
constexpr int LQ_PS_F64 = 418;  // 418/2 = 209 odd

// One backward Riccati step of the quad that owns the slab (Vs, As, Bs, Ms, Gs): V <- Q_k + A'VA - M'G^-1 M in place, and
// this thread's 4x4 tile of L_k = G^-1 B'VA (columns 4t..4t+3; thread 3's tile is not a gain).  Ends with a __syncwarp().
// QDIAG: the caller asserted diagonal Q, R (ZB_COST_DIAGONAL): Qk / Rk then point at their DIAGONALS (12 + 4 words, kept in
// the slab) -- reading dense Q_k blocks from global memory every step was 25 % of the stall samples (L1 is mostly carved
// out as shared memory, so they came from L2).
template <typename T, bool QDIAG = false>
__device__ __forceinline__ void riccati_quad_step(T* Vs, const T* As, const T* Bs, T* Ms, T* Gs, const T* Qk, const T* Rk, int t,
                                                  T (&L)[4][4]) {
    const T* Ct = (t < 3) ? (As + 4 * t) : Bs;
    const int cstride = (t < 3) ? 12 : 4;
    // ---- 1. [W | VB] tile = V * tile ----
    T W[12][4];
#pragma unroll
    for (int i = 0; i < 12; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) W[i][c] = T(0);
#pragma unroll
    for (int kk = 0; kk < 12; ++kk) {
        const Vec4<T> c4 = ldv4(Ct + kk * cstride);
        const Vec4<T> v0 = ldv4(Vs + kk * 12), v1 = ldv4(Vs + kk * 12 + 4), v2 = ldv4(Vs + kk * 12 + 8);
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                W[i][c] = fma(v0.v[i], c4.v[c], W[i][c]);
                W[4 + i][c] = fma(v1.v[i], c4.v[c], W[4 + i][c]);
                W[8 + i][c] = fma(v2.v[i], c4.v[c], W[8 + i][c]);
            }
    }
    // ---- 2. [M | G0] tile = B' * [W | VB] tile ----
    T M[4][4];
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
    // ---- 3. G = R_k + G0 (thread 3), M tile (threads 0..2) -> shared ----
    if (t == 3) {
#pragma unroll
        const Vec4<T> rd = QDIAG ? ldv4(Rk) : Vec4<T>{{T(0), T(0), T(0), T(0)}};
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            Vec4<T> r4;
            if (QDIAG) r4 = Vec4<T>{{a == 0 ? rd.v[0] : T(0), a == 1 ? rd.v[1] : T(0), a == 2 ? rd.v[2] : T(0), a == 3 ? rd.v[3] : T(0)}};
            else r4 = ldv4(Rk + a * 4);
            stv4(Gs + a * 4, M[a][0] + r4.v[0], M[a][1] + r4.v[1], M[a][2] + r4.v[2], M[a][3] + r4.v[3]);
        }
    } else {
#pragma unroll
        for (int a = 0; a < 4; ++a) stv4(Ms + a * 12 + 4 * t, M[a][0], M[a][1], M[a][2], M[a][3]);
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
#pragma unroll
    for (int c = 0; c < 4; ++c) {  // L tile = G^-1 M tile
        const T y0 = M[0][c] * d0;
        const T y1 = fma(-c10, y0, M[1][c]) * d1;
        const T y2 = fma(-c21, y1, fma(-c20, y0, M[2][c])) * d2;
        const T y3 = fma(-c32, y2, fma(-c31, y1, fma(-c30, y0, M[3][c]))) * d3;
        const T x3 = y3 * d3;
        const T x2 = fma(-c32, x3, y2) * d2;
        const T x1 = fma(-c31, x3, fma(-c21, x2, y1)) * d1;
        const T x0 = fma(-c30, x3, fma(-c20, x2, fma(-c10, x1, y0))) * d0;
        L[0][c] = x0; L[1][c] = x1; L[2][c] = x2; L[3][c] = x3;
    }
    // ---- 5. V' = Q_k + A'W - M'L: of the 9 4x4 blocks 6 are distinct; thread t computes (t,t) and ((t+1)%3, t) ----
#pragma unroll
    for (int bi = 0; bi < 2; ++bi) {
        const int tb = t < 3 ? t : 0;  // thread 3 (the B tile) owns no block: it shadows thread 0 and stores nothing
        const int sblk = (bi == 0) ? tb : (tb == 2 ? 0 : tb + 1);
        const int tcol = 4 * tb;
        T acc[4][4];
        if (QDIAG) {
            const Vec4<T> qd = ldv4(Qk + tcol);
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int c = 0; c < 4; ++c) acc[i][c] = (sblk == tb && i == c) ? qd.v[c] : T(0);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const Vec4<T> q4 = ldv4(Qk + (4 * sblk + i) * 12 + tcol);
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
                for (int c = 0; c < 4; ++c) acc[i][c] = fma(-m4.v[i], L[a][c], acc[i][c]);
        }
        // write back, lower triangle wins: off-diagonal blocks are stored with their transpose, diagonal blocks mirrored
        if (t < 3 && sblk != t) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                stv4(Vs + (4 * sblk + r) * 12 + 4 * t, acc[r][0], acc[r][1], acc[r][2], acc[r][3]);
                stv4(Vs + (4 * t + r) * 12 + 4 * sblk, acc[0][r], acc[1][r], acc[2][r], acc[3][r]);
            }
        } else if (t < 3) {
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

template <typename T, bool QDIAG>
__global__ void __launch_bounds__(256) k_riccati_quad(LqrQuadP P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PS = sizeof(T) == 4 ? LQ_PS_F32 : LQ_PS_F64;
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3, quad = lane >> 2;
    const long long b_raw = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * 8 + quad;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    T* S = smem + (warp * 8 + quad) * PS;
    T *Vs = S + LQ_V, *As = S + LQ_A, *Bs = S + LQ_B, *Ms = S + LQ_M, *Gs = S + LQ_G;
    const int N = P.N;
    const bool tvAB = (P.A.st != 0 || P.B.st != 0) && N > 1;
    // ---- staging: terminal value (Qf, or Q[T-1]: lqrUtils.py:172), A, B ----
    {
        const T* Vf = P.Qf.p ? P.Qf.at<T>(b) : P.Q.at<T>(b, P.T - 1);
        const T* A = P.A.at<T>(b, N - 1);
        const T* B = P.B.at<T>(b, N - 1);
        for (int e = t; e < 144; e += 4) { Vs[e] = Vf[e]; As[e] = A[e]; }
        for (int e = t; e < 48; e += 4) Bs[e] = B[e];
        if (QDIAG) {  // time-invariant diagonal costs (lqrMpc.solve): diagonals into the slab
            const T* Q = P.Q.at<T>(b);
            const T* R = P.R.at<T>(b);
            for (int i = t; i < 12; i += 4) S[LQ_QD + i] = Q[i * 13];
            S[LQ_QD + 12 + t] = R[t * 5];
        }
    }
    __syncwarp();
    T* Lo = reinterpret_cast<T*>(P.L) + b * (long long)N * 48;
    for (int k = N - 1; k >= 0; --k) {
        if (tvAB && k != N - 1) {  // genuinely time-varying dynamics: re-stage A_k, B_k
            const T* A = P.A.at<T>(b, k);
            const T* B = P.B.at<T>(b, k);
            for (int e = t; e < 144; e += 4) As[e] = A[e];
            for (int e = t; e < 48; e += 4) Bs[e] = B[e];
            __syncwarp();
        }
        T L[4][4];
        if (QDIAG) riccati_quad_step<T, true>(Vs, As, Bs, Ms, Gs, S + LQ_QD, S + LQ_QD + 12, t, L);
        else riccati_quad_step<T, false>(Vs, As, Bs, Ms, Gs, P.Q.at<T>(b, k), P.R.at<T>(b, k), t, L);
        if (active && t < 3) {  // gains -> global
            T* g = Lo + (long long)k * 48 + 4 * t;
#pragma unroll
            for (int a = 0; a < 4; ++a) stv4(g + a * 12, L[a][0], L[a][1], L[a][2], L[a][3]);
        }
    }
    if (P.V0 && active) {
        T* o = reinterpret_cast<T*>(P.V0) + b * 144;
        for (int e = t; e < 144; e += 4) o[e] = Vs[e];
    }
}

// Closed-loop LQR-MPC of the quadcopter in fp64 (BASELINE cfg 3 in the reference's own precision; the fp32 kernels are in
// lqr_t1.cuh / mpc_coop.cuh), fused: per simulation step the quad re-linearises the Euler quadcopter at (x_t, u_trim) in
// place (analytic dF/dx, one sincos per thread, only the 17 state-dependent chunks of f_x rewritten), runs the full
// N-step Riccati sweep from Qf keeping only L_0, applies u_t = -L_0 x_t and steps the nonlinear plant.  Only the simulated
// trajectory goes to HBM.  The state is replicated in the four threads of the quad (every thread steps the plant with the
// same inputs, hence the same bits).
template <typename T, bool QDIAG>
__global__ void __launch_bounds__(256) k_mpc_closed_loop_quad64(ClosedLoopQuadP P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int PS = sizeof(T) == 4 ? LQ_PS_F32 : LQ_PS_F64;
    T* smem = reinterpret_cast<T*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int t = lane & 3, quad = lane >> 2, qbase = lane & 28;
    const long long b_raw = ((long long)blockIdx.x * (blockDim.x >> 5) + warp) * 8 + quad;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    T* S = smem + (warp * 8 + quad) * PS;
    T *Vs = S + LQ_V, *As = S + LQ_A, *Bs = S + LQ_B, *Ms = S + LQ_M, *Gs = S + LQ_G;
    const T dt = T(P.dt);
    const T* Q = P.Q.at<T>(b);
    const T* R = P.R.at<T>(b);
    const T* Qf = P.Qf.at<T>(b);
    const T ut[4] = {T(P.utrim[0]), T(P.utrim[1]), T(P.utrim[2]), T(P.utrim[3])};
    T x[12];
    {
        const T* x0 = reinterpret_cast<const T*>(P.x0) + b * 12;
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = x0[i];
        // state-independent part of f_x, and f_u = dt dF/du (four entries: quad_model_gen.cuh)
        const T z12[12] = {T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0), T(0)}, z4[4] = {T(0), T(0), T(0), T(0)};
        QuadTrig<T> tr0;
        tr0.sph = T(0); tr0.cph = T(1); tr0.sth = T(0); tr0.cth = T(1); tr0.sps = T(0); tr0.cps = T(1); tr0.tth = T(0); tr0.sec = T(1);
        T J0[144];
        quad_jac_x(tr0, z12, z4, J0);
        store_fx<T, true>(As, J0, dt, t);
        for (int e = t; e < 48; e += 4) Bs[e] = T(0);
        __syncwarp();
        if (t == 0) { Bs[2 * 4 + 0] = -dt; Bs[3 * 4 + 1] = dt; Bs[4 * 4 + 2] = dt; Bs[5 * 4 + 3] = dt; }
        if (QDIAG) {
            for (int i = t; i < 12; i += 4) S[LQ_QD + i] = Q[i * 13];
            S[LQ_QD + 12 + t] = R[t * 5];
        }
    }
    T* xS = reinterpret_cast<T*>(P.xSim) + b * (long long)(P.Tsim + 1) * 12;
    T* uS = reinterpret_cast<T*>(P.uSim) + b * (long long)P.Tsim * 4;
    if (active && t == 0) {
#pragma unroll
        for (int q = 0; q < 3; ++q) stv4(xS + 4 * q, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
    }
    for (int s = 0; s < P.Tsim; ++s) {
        // ---- linearise at (x_t, u_trim) ----
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
            T J[144];
            quad_jac_x(tr, x, ut, J);
            store_fx<T, false>(As, J, dt, t);
        }
        for (int e = t; e < 144; e += 4) Vs[e] = Qf[e];
        __syncwarp();
        // ---- full Riccati sweep; only the first gain is used (nothing cached across simulation steps) ----
        T L[4][4];
        for (int k = P.N - 1; k >= 0; --k) {
            if (QDIAG) riccati_quad_step<T, true>(Vs, As, Bs, Ms, Gs, S + LQ_QD, S + LQ_QD + 12, t, L);
            else riccati_quad_step<T, false>(Vs, As, Bs, Ms, Gs, Q, R, t, L);
        }
        // ---- u_t = -L_0 x_t: partial products of the three A tiles, summed in a fixed order ----
        T part[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int c0 = (t < 3) ? 4 * t : 0;
            T sacc = T(0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const T xv = (c0 == 0) ? x[c] : (c0 == 4) ? x[4 + c] : x[8 + c];
                sacc = fma(L[a][c], xv, sacc);
            }
            part[a] = sacc;
        }
        T u[4], uapp[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const T p0 = __shfl_sync(0xffffffffu, part[a], qbase), p1 = __shfl_sync(0xffffffffu, part[a], qbase + 1),
                    p2 = __shfl_sync(0xffffffffu, part[a], qbase + 2);
            u[a] = -((p0 + p1) + p2);
            uapp[a] = ut[a] + u[a];
        }
        // ---- plant: x_{t+1} = x_t + dt F(x_t, u_trim + u_t) (zopt/quadcopter.py:116-144) ----
        T xd[12];
        quad_xdot(tr, x, uapp, xd);
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = x[i] + dt * xd[i];
        if (active && t == 0) {
#pragma unroll
            for (int q = 0; q < 3; ++q) stv4(xS + (long long)(s + 1) * 12 + 4 * q, x[4 * q], x[4 * q + 1], x[4 * q + 2], x[4 * q + 3]);
            stv4(uS + (long long)s * 4, u[0], u[1], u[2], u[3]);
        }
    }
}

// Plan rollout of the unconstrained lqrMpc.solve (mpcUtils.py:47-59 with no active bound): u_k = -L_k x_k,
// x_{k+1} = A x_k + B u_k.  Four threads per problem: thread t keeps rows 3t..3t+2 of [A | B] in registers and computes
// u_k[t] from row t of L_k; u and x are all-gathered inside the quad with shuffles.
template <typename T>
__global__ void __launch_bounds__(128) k_plan_rollout_quad(LqrQuadP P) {
    const long long tid = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31, t = lane & 3, qbase = lane & 28;
    const long long b_raw = tid >> 2;
    const bool active = b_raw < P.Bsz;
    const long long b = active ? b_raw : P.Bsz - 1;
    const int N = P.N;
    const T* A = P.A.at<T>(b);
    const T* B = P.B.at<T>(b);
    T Ar[3][12], Br[3][4];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const Vec4<T> a4 = ldv4(A + (3 * t + r) * 12 + 4 * q);
#pragma unroll
            for (int c = 0; c < 4; ++c) Ar[r][4 * q + c] = a4.v[c];
        }
        const Vec4<T> b4 = ldv4(B + (3 * t + r) * 4);
#pragma unroll
        for (int c = 0; c < 4; ++c) Br[r][c] = b4.v[c];
    }
    const T* Lg = reinterpret_cast<const T*>(P.L) + b * (long long)N * 48;
    const T* x0 = reinterpret_cast<const T*>(P.x0) + b * 12;
    T* xT = reinterpret_cast<T*>(P.xTraj) + b * (long long)(N + 1) * 12;
    T* uT = reinterpret_cast<T*>(P.uTraj) + b * (long long)N * 4;
    T x[12], own[3];  // own = x[3t..3t+2], the rows this thread produces
#pragma unroll
    for (int i = 0; i < 12; ++i) x[i] = x0[i];
#pragma unroll
    for (int r = 0; r < 3; ++r) own[r] = (t == 0) ? x[r] : (t == 1) ? x[3 + r] : (t == 2) ? x[6 + r] : x[9 + r];
    for (int k = 0; k < N; ++k) {
        const T* Lk = Lg + (long long)k * 48 + t * 12;
        T s = T(0);
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const Vec4<T> l4 = ldv4(Lk + 4 * q);
#pragma unroll
            for (int c = 0; c < 4; ++c) s = fma(l4.v[c], x[4 * q + c], s);
        }
        const T ut = -s;
        T u[4];
#pragma unroll
        for (int a = 0; a < 4; ++a) u[a] = __shfl_sync(0xffffffffu, ut, qbase + a);
        if (active) {
#pragma unroll
            for (int r = 0; r < 3; ++r) xT[(long long)k * 12 + 3 * t + r] = own[r];
            uT[(long long)k * 4 + t] = ut;
            if (k == 0) reinterpret_cast<T*>(P.u0)[b * 4 + t] = ut;
        }
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            T acc = T(0);
#pragma unroll
            for (int j = 0; j < 12; ++j) acc = fma(Ar[r][j], x[j], acc);
#pragma unroll
            for (int j = 0; j < 4; ++j) acc = fma(Br[r][j], u[j], acc);
            own[r] = acc;
        }
#pragma unroll
        for (int i = 0; i < 12; ++i) x[i] = __shfl_sync(0xffffffffu, own[i % 3], qbase + i / 3);
    }
    if (active) {
#pragma unroll
        for (int r = 0; r < 3; ++r) xT[(long long)N * 12 + 3 * t + r] = own[r];
        if (t == 0) {
            P.status[b] = 0;
            if (P.iters) P.iters[b] = 0;
        }
    }
}

// Warps per CTA.  An SM holds eight warps (64 problems) of these kernels.  A batch of more than one wave of the machine is
// launched as ONE CTA per SM with the smallest warp count that keeps the number of waves, so that the last wave is as full as
// the first (16,384 problems: 293 CTAs of 7 warps = 1.98 waves instead of 1.73 -> 2); smaller batches keep two-warp CTAs
// spread over the SMs (four per SM).
inline int quad64_warps(long long Bsz) {
    const long long nsm = 148;
    auto waves = [&](int w) { const long long ctas = (Bsz + 8 * w - 1) / (8 * w); return (ctas + nsm - 1) / nsm; };
    if (waves(8) == 1) return 2;
    int w = 8;
    while (w > 4 && waves(w - 1) == waves(8)) --w;
    return w;
}

int32_t riccati_quad_launch(int32_t dtype, const LqrQuadP& P, cudaStream_t stream) {
    ZB_ARG(dtype == ZB_F64, "the cooperative (12,4) Riccati kernel is instantiated for fp64 (fp32 has its own kernels, lqr_t1.cuh)");
    const int warps = quad64_warps(P.Bsz);
    const size_t smem = (size_t)warps * 8 * LQ_PS_F64 * sizeof(double);
    const unsigned grid = (unsigned)((P.Bsz + warps * 8 - 1) / (warps * 8));
    if (P.cost_diagonal && P.Q.st == 0 && P.R.st == 0) {
        ZB_CUDA(cudaFuncSetAttribute(k_riccati_quad<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_riccati_quad<double, true><<<grid, warps * 32, smem, stream>>>(P);
    } else {
        ZB_CUDA(cudaFuncSetAttribute(k_riccati_quad<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_riccati_quad<double, false><<<grid, warps * 32, smem, stream>>>(P);
    }
    ZB_CUDA(cudaGetLastError());
    if (P.x0) {  // lqrMpc.solve: plan rollout against the gains just written
        const unsigned g2 = (unsigned)((P.Bsz * 4 + 127) / 128);
        k_plan_rollout_quad<double><<<g2, 128, 0, stream>>>(P);
        ZB_CUDA(cudaGetLastError());
    }
    return 0;
}

int32_t mpc_closed_loop_quad64_launch(const ClosedLoopQuadP& P, cudaStream_t stream) {
    const int warps = quad64_warps(P.Bsz);
    const size_t smem = (size_t)warps * 8 * LQ_PS_F64 * sizeof(double);
    const unsigned grid = (unsigned)((P.Bsz + warps * 8 - 1) / (warps * 8));
    if (P.cost_diagonal) {
        ZB_CUDA(cudaFuncSetAttribute(k_mpc_closed_loop_quad64<double, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_mpc_closed_loop_quad64<double, true><<<grid, warps * 32, smem, stream>>>(P);
    } else {
        ZB_CUDA(cudaFuncSetAttribute(k_mpc_closed_loop_quad64<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_mpc_closed_loop_quad64<double, false><<<grid, warps * 32, smem, stream>>>(P);
    }
    ZB_CUDA(cudaGetLastError());
    return 0;
}

}  // namespace zb
