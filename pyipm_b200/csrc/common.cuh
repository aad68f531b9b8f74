// common.cuh -- error handling, launch accounting and warp/block reductions shared by all kernels.
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <atomic>

namespace b200 {

extern thread_local std::string g_last_error;
extern std::atomic<long long> g_launches;

inline int fail(const char* what, const char* file, int line, cudaError_t e) {
    char buf[512];
    snprintf(buf, sizeof(buf), "%s failed at %s:%d: %s", what, file, line, cudaGetErrorString(e));
    g_last_error = buf;
    return -1;
}
inline int fail_msg(const char* msg) {
    g_last_error = msg;
    return -2;
}

#define CU(call)                                                        \
    do {                                                                \
        cudaError_t e__ = (call);                                       \
        if (e__ != cudaSuccess) return ::b200::fail(#call, __FILE__, __LINE__, e__); \
    } while (0)

// count a kernel launch and surface launch-configuration errors immediately
#define LAUNCHED()                                                      \
    do {                                                                \
        ::b200::g_launches.fetch_add(1, std::memory_order_relaxed);     \
        cudaError_t e__ = cudaGetLastError();                           \
        if (e__ != cudaSuccess) return ::b200::fail("kernel launch", __FILE__, __LINE__, e__); \
    } while (0)

#define RET(call)                        \
    do {                                 \
        int r__ = (call);                \
        if (r__ != 0) return r__;        \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t rup(size_t a, size_t b) { return (a + b - 1) / b * b; }

// ------------------------------------------------------------------------------------ device-side timeline trace
// Profiling aid (b200ipm_trace_start / b200ipm_trace_dump): when enabled, selected CTAs of the factorisation kernels
// stamp %globaltimer at entry and exit.  Disabled (null pointer) it costs one global load per kernel.
struct TraceRec { int id, blk; unsigned long long t0, t1; unsigned long long tag; };
constexpr int TRACE_CAP = 1 << 16;
__device__ TraceRec* g_trace = nullptr;
__device__ int g_trace_n = 0;
__device__ __forceinline__ unsigned long long trace_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ int trace_begin(int id, const void* tag = nullptr) {
    TraceRec* t = g_trace;
    if (!t) return -1;
    const int i = atomicAdd(&g_trace_n, 1);
    if (i >= TRACE_CAP) return -1;
    t[i].id = id;
    t[i].blk = (int)(blockIdx.x + blockIdx.y * gridDim.x);
    t[i].t0 = trace_now();
    t[i].t1 = 0;
    t[i].tag = (unsigned long long)tag;     // e.g. the control block of the factorisation the kernel belongs to
    return i;
}
__device__ __forceinline__ void trace_end(int i) {
    if (i >= 0) g_trace[i].t1 = trace_now();
}
enum { TR_TILE = 1, TR_PANEL = 2, TR_MINI = 3, TR_SUB64 = 4, TR_DMMA = 5, TR_OZ = 6, TR_SLICE = 7, TR_BINV = 8 };

// ------------------------------------------------------------------------------------ device reductions
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// NaN-propagating sum over a block of up to 1024 threads; result valid in every thread.
// `sh` must hold 33 doubles.  Deterministic (fixed tree).
__device__ __forceinline__ double block_sum(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? sh[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}
__device__ __forceinline__ double block_min(double v, double* sh) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_min(v);
    __syncthreads();
    if (lane == 0) sh[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? sh[lane] : INFINITY;
        t = warp_min(t);
        if (lane == 0) sh[32] = t;
    }
    __syncthreads();
    return sh[32];
}

}  // namespace b200
