// Host-side plumbing of the C ABI: error reporting, device/stream guards, dtype dispatch.
#pragma once

#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include "../../include/zopt_b200.h"
#include "zb_problems.cuh"

namespace zb {

inline char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

inline int32_t fail(int32_t code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

#define ZB_ARG(cond, ...)                                      \
    do {                                                       \
        if (!(cond)) return zb::fail(-1, __VA_ARGS__);         \
    } while (0)

#define ZB_CUDA(expr)                                                                                        \
    do {                                                                                                     \
        cudaError_t e__ = (expr);                                                                            \
        if (e__ != cudaSuccess)                                                                              \
            return zb::fail((int32_t)e__, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, \
                            __LINE__);                                                                       \
    } while (0)

// select the device for the duration of a call, restore on exit
struct DeviceGuard {
    int prev = -1;
    cudaError_t err = cudaSuccess;
    explicit DeviceGuard(int dev) {
        err = cudaGetDevice(&prev);
        if (err == cudaSuccess && prev != dev) err = cudaSetDevice(dev);
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
};

inline Arr to_arr(const zb_arr* a) {
    Arr r;
    if (a) {
        r.p = a->ptr;
        r.sb = a->stride_b;
        r.st = a->stride_t;
    } else {
        r.p = nullptr;
        r.sb = r.st = 0;
    }
    return r;
}

inline int32_t to_model(const zb_model* m, Model& M) {
    ZB_ARG(m != nullptr, "model is NULL");
    M.kind = m->kind;
    M.n = m->n;
    M.m = m->m;
    M.has_wind = m->has_wind;
    M.dt = m->dt;
    for (int i = 0; i < 3; ++i) M.wind[i] = m->wind[i];
    M.A = to_arr(&m->A);
    M.B = to_arr(&m->B);
    if (m->kind == ZB_MODEL_QUADCOPTER) {
        ZB_ARG(m->n == 12 && m->m == 4, "quadcopter model has n=12, m=4 (got %d, %d)", m->n, m->m);
    } else if (m->kind == ZB_MODEL_LINEAR) {
        ZB_ARG(m->n >= 1 && m->n <= ZB_MAX_N && m->m >= 1 && m->m <= ZB_MAX_M, "linear model: n<=%d, m<=%d (got %d, %d)",
               ZB_MAX_N, ZB_MAX_M, m->n, m->m);
        ZB_ARG(m->A.ptr && m->B.ptr, "linear model needs A and B");
    } else {
        return fail(-2, "unsupported model kind %d (no CPU fallback for arbitrary callables)", m->kind);
    }
    return 0;
}

inline Cost to_cost(const zb_cost* c) {
    Cost C;
    C.Q = to_arr(c ? &c->Q : nullptr);
    C.R = to_arr(c ? &c->R : nullptr);
    C.Qf = to_arr(c ? &c->Qf : nullptr);
    return C;
}

inline int32_t check_dims(int32_t dtype, int64_t Bsz, int32_t n, int32_t m) {
    ZB_ARG(dtype == ZB_F32 || dtype == ZB_F64, "dtype must be ZB_F32 or ZB_F64 (got %d)", dtype);
    ZB_ARG(Bsz >= 0, "negative batch size");
    ZB_ARG(n >= 1 && n <= ZB_MAX_N, "n must be in [1,%d] (got %d)", ZB_MAX_N, n);
    ZB_ARG(m >= 1 && m <= ZB_MAX_M, "m must be in [1,%d] (got %d)", ZB_MAX_M, m);
    return 0;
}

// One non-blocking side stream per (host thread, device), created on first use and kept: the fork/join overlap of
// lqrMpc.solve's plan rollout with the next chunk's sweep (lqr_t1.cuh) runs its rollouts there.
inline int32_t side_stream(int device, cudaStream_t* out) {
    static thread_local cudaStream_t streams[64] = {nullptr};
    ZB_ARG(device >= 0 && device < 64, "device index %d out of range", device);
    if (!streams[device]) {
        // greatest priority: the block scheduler keeps dispatching the CTAs of the kernel it started first and turns to another
        // equal-priority kernel only when that grid is exhausted; a higher-priority kernel gets its (few, small) CTAs placed
        // as soon as they fit
        int lo = 0, hi = 0;
        ZB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));  // hi = greatest priority (numerically lowest)
        ZB_CUDA(cudaStreamCreateWithPriority(&streams[device], cudaStreamNonBlocking, hi));
    }
    *out = streams[device];
    return 0;
}

constexpr int GEN_THREADS = 64;  // generic kernels: threads per block (local-memory heavy)
// Stream-ordered scratch (cudaMallocAsync / cudaFreeAsync) comes from the device's default memory pool, whose release threshold is
// 0 by default: every synchronisation hands the freed blocks back to the OS and the next call allocates afresh -- measured on the
// nine-lane closed-loop launch: 1.5 ms of HOST time per call, with outliers of 30-57 ms that land between the caller's events
// (cfg 3's shard of eight: 8.2 ms median but 11-65 ms in one launch of ten).  Keeping the pool's memory makes the allocation a
// sub-microsecond pool operation.  Once per device.
inline cudaError_t keep_pool_memory() {
    static unsigned long long done = 0;  // bit per device; a benign race only repeats the call
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && ((done >> dev) & 1ull)) return cudaSuccess;
    cudaMemPool_t pool;
    if ((e = cudaDeviceGetDefaultMemPool(&pool, dev)) != cudaSuccess) return e;
    unsigned long long keep = ~0ull;
    if ((e = cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep)) != cudaSuccess) return e;
    if (dev < 64) done |= 1ull << dev;
    return cudaSuccess;
}

inline unsigned gen_grid(long long work) { return (unsigned)((work + GEN_THREADS - 1) / GEN_THREADS); }

}  // namespace zb
