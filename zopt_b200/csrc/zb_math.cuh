// Small dense linear algebra on per-thread arrays with run-time dimensions.
// Shared by the generic (one thread per problem) kernels of every entry point.
// Everything is row-major.  ZB_HD lets tests/hostsim compile the same arithmetic
// for the host as a pre-GPU logic check (test infrastructure only; the product
// never runs it on the CPU).
#pragma once

#if defined(__CUDACC__)
#define ZB_HD __host__ __device__ __forceinline__
#else
#define ZB_HD inline
#endif

#include <math.h>

#ifndef ZB_PD_MAX
#define ZB_PD_MAX 24  // largest symmetric block handled by pd_clamp (ZB_MAX_N + ZB_MAX_M)
#endif

namespace zb {

// C(p x r) = A(p x q) * B(q x r)
template <typename T, typename TA, typename TB>
ZB_HD void mm(T* C, const TA* A, const TB* B, int p, int q, int r) {
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < r; ++j) {
            T s = T(0);
            for (int k = 0; k < q; ++k) s += T(A[i * q + k]) * T(B[k * r + j]);
            C[i * r + j] = s;
        }
}

// C(p x r) = A^T * B with A stored (q x p), B (q x r)
template <typename T, typename TA, typename TB>
ZB_HD void mm_tn(T* C, const TA* A, const TB* B, int p, int q, int r) {
    for (int i = 0; i < p; ++i)
        for (int j = 0; j < r; ++j) {
            T s = T(0);
            for (int k = 0; k < q; ++k) s += T(A[k * p + i]) * T(B[k * r + j]);
            C[i * r + j] = s;
        }
}

// y(p) = A(p x q) x
template <typename T, typename TA, typename TX>
ZB_HD void mv(T* y, const TA* A, const TX* x, int p, int q) {
    for (int i = 0; i < p; ++i) {
        T s = T(0);
        for (int k = 0; k < q; ++k) s += T(A[i * q + k]) * T(x[k]);
        y[i] = s;
    }
}

// y(p) = A^T x with A stored (q x p)
template <typename T, typename TA, typename TX>
ZB_HD void mv_t(T* y, const TA* A, const TX* x, int p, int q) {
    for (int i = 0; i < p; ++i) {
        T s = T(0);
        for (int k = 0; k < q; ++k) s += T(A[k * p + i]) * T(x[k]);
        y[i] = s;
    }
}

template <typename T, typename TX, typename TY>
ZB_HD T dot(const TX* x, const TY* y, int p) {
    T s = T(0);
    for (int k = 0; k < p; ++k) s += T(x[k]) * T(y[k]);
    return s;
}

// In-place LU with partial pivoting of G (m x m), then solve G X = RHS for nrhs columns
// (RHS is m x nrhs row-major, overwritten with X).  Same algorithm as LAPACK getrf/getrs,
// which jnp.linalg.solve lowers to on CPU (zopt/lqrUtils.py:168, ilqrUtils.py:167-168).
template <typename T>
ZB_HD void lu_solve(T* G, int m, T* RHS, int nrhs) {
    for (int c = 0; c < m; ++c) {
        int piv = c;
        T best = fabs(G[c * m + c]);
        for (int i = c + 1; i < m; ++i) {
            T a = fabs(G[i * m + c]);
            if (a > best) { best = a; piv = i; }
        }
        if (piv != c) {
            for (int j = 0; j < m; ++j) { T t = G[c * m + j]; G[c * m + j] = G[piv * m + j]; G[piv * m + j] = t; }
            for (int j = 0; j < nrhs; ++j) { T t = RHS[c * nrhs + j]; RHS[c * nrhs + j] = RHS[piv * nrhs + j]; RHS[piv * nrhs + j] = t; }
        }
        T inv = T(1) / G[c * m + c];
        for (int i = c + 1; i < m; ++i) {
            T f = G[i * m + c] * inv;
            G[i * m + c] = f;
            for (int j = c + 1; j < m; ++j) G[i * m + j] -= f * G[c * m + j];
            for (int j = 0; j < nrhs; ++j) RHS[i * nrhs + j] -= f * RHS[c * nrhs + j];
        }
    }
    for (int i = m - 1; i >= 0; --i) {
        T inv = T(1) / G[i * m + i];
        for (int j = 0; j < nrhs; ++j) {
            T s = RHS[i * nrhs + j];
            for (int k = i + 1; k < m; ++k) s -= G[i * m + k] * RHS[k * nrhs + j];
            RHS[i * nrhs + j] = s * inv;
        }
    }
}

// ensurePositiveDefinite (zopt/ilqrUtils.py:217-219): S <- V max(Lambda, eps) V^T for symmetric S (p x p).
// Cyclic Jacobi eigen-decomposition; the result is a spectral function of S, hence independent of the
// eigenvector basis LAPACK syevd would have picked.  W is p x p scratch (eigenvectors).
// In a user-model plug-in p is a compile-time constant at every call site (n or n + m of that model): for small blocks the
// loops are unrolled after inlining, so S and W are indexed statically and live in registers instead of local memory.
#if defined(ZB_USER_MODEL_HEADER) && ZB_PD_MAX <= 8  // (the header name comes on the plug-in's nvcc command line)
#define ZB_PD_UNROLL _Pragma("unroll")
#else
#define ZB_PD_UNROLL
#endif
template <typename T>
ZB_HD void pd_clamp(T* S, T* W, int p, T eps) {
    ZB_PD_UNROLL
    for (int i = 0; i < p; ++i)
        ZB_PD_UNROLL
        for (int j = 0; j < p; ++j) W[i * p + j] = (i == j) ? T(1) : T(0);
    // jnp.linalg.eigh symmetrises its input by default (symmetrize_input=True): (S + S^T)/2
    ZB_PD_UNROLL
    for (int i = 0; i < p; ++i)
        ZB_PD_UNROLL
        for (int j = i + 1; j < p; ++j) {
            T a = T(0.5) * (S[i * p + j] + S[j * p + i]);
            S[i * p + j] = a;
            S[j * p + i] = a;
        }
    const T tiny = (sizeof(T) == 8) ? T(1e-300) : T(1e-37);
    for (int sweep = 0; sweep < 30; ++sweep) {
        T off = T(0), diag = T(0);
        ZB_PD_UNROLL
        for (int i = 0; i < p; ++i) {
            diag += S[i * p + i] * S[i * p + i];
            ZB_PD_UNROLL
            for (int j = 0; j < i; ++j) off += S[i * p + j] * S[i * p + j];
        }
        const T thr = (sizeof(T) == 8) ? T(1e-32) : T(1e-15);
        if (off <= thr * diag || off < tiny) break;
        ZB_PD_UNROLL
        for (int a = 0; a < p - 1; ++a)
            ZB_PD_UNROLL
            for (int b = a + 1; b < p; ++b) {
                T apq = S[a * p + b];
                if (apq == T(0)) continue;
                T app = S[a * p + a], aqq = S[b * p + b];
                T tau = (aqq - app) / (T(2) * apq);
                T t = (tau >= T(0) ? T(1) : T(-1)) / (fabs(tau) + sqrt(T(1) + tau * tau));
                T c = T(1) / sqrt(T(1) + t * t), s = t * c;
                ZB_PD_UNROLL
                for (int k = 0; k < p; ++k) {  // columns a,b of S
                    T ska = S[k * p + a], skb = S[k * p + b];
                    S[k * p + a] = c * ska - s * skb;
                    S[k * p + b] = s * ska + c * skb;
                }
                ZB_PD_UNROLL
                for (int k = 0; k < p; ++k) {  // rows a,b of S
                    T sak = S[a * p + k], sbk = S[b * p + k];
                    S[a * p + k] = c * sak - s * sbk;
                    S[b * p + k] = s * sak + c * sbk;
                }
                ZB_PD_UNROLL
                for (int k = 0; k < p; ++k) {  // accumulate eigenvectors (columns of W)
                    T wka = W[k * p + a], wkb = W[k * p + b];
                    W[k * p + a] = c * wka - s * wkb;
                    W[k * p + b] = s * wka + c * wkb;
                }
            }
    }
    // eigenvalues on the diagonal of S; rebuild
    T lam[ZB_PD_MAX];
    ZB_PD_UNROLL
    for (int i = 0; i < p; ++i) lam[i] = S[i * p + i] > eps ? S[i * p + i] : eps;
    ZB_PD_UNROLL
    for (int i = 0; i < p; ++i)
        ZB_PD_UNROLL
        for (int j = 0; j <= i; ++j) {
            T s = T(0);
            ZB_PD_UNROLL
            for (int k = 0; k < p; ++k) s += W[i * p + k] * lam[k] * W[j * p + k];
            S[i * p + j] = s;
            S[j * p + i] = s;
        }
}

}  // namespace zb
