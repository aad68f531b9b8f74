// vec.cuh -- HBM-bound building blocks: GEMV (both orientations), transposes, small reductions.
// All matrices are row-major float64.  Kernels are written for coalesced, vectorised (16-byte) HBM loads;
// the residual GEMV (a1 in SURVEY.md section 8) is the headline bandwidth-bound kernel.
#pragma once
#include "common.cuh"

namespace b200 {

// ---------------------------------------------------------------------------------------------- gemv_n
// y[r] = ys * y0[r] + alpha * sum_c A[r, c] * v[c]      (one warp per row, double2 loads when aligned)
// Optionally accumulates sum_r y[r]^2 per block into sq_part[blockIdx.x] (KKT norm fused into the GEMV).
template <bool VEC>
__global__ void __launch_bounds__(256) gemv_n_kernel(const double* __restrict__ A, int lda, int rows, int cols,
                                                     const double* __restrict__ v, const double* __restrict__ y0,
                                                     double ys, double alpha, double* __restrict__ y,
                                                     double* __restrict__ sq_part) {
    __shared__ double sh[8];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int r = blockIdx.x * 8 + w;
    double acc = 0.0;
    if (r < rows) {
        const double* a = A + (size_t)r * lda;
        if (VEC) {
            const double2* a2 = reinterpret_cast<const double2*>(a);
            const double2* v2 = reinterpret_cast<const double2*>(v);
            const int n2 = cols >> 1;
            int c = lane;
            // 4 independent 16-byte loads in flight per lane
            for (; c + 96 < n2; c += 128) {
                double2 p0 = __ldg(a2 + c), p1 = __ldg(a2 + c + 32), p2 = __ldg(a2 + c + 64), p3 = __ldg(a2 + c + 96);
                double2 q0 = __ldg(v2 + c), q1 = __ldg(v2 + c + 32), q2 = __ldg(v2 + c + 64), q3 = __ldg(v2 + c + 96);
                acc += p0.x * q0.x + p0.y * q0.y;
                acc += p1.x * q1.x + p1.y * q1.y;
                acc += p2.x * q2.x + p2.y * q2.y;
                acc += p3.x * q3.x + p3.y * q3.y;
            }
            for (; c < n2; c += 32) {
                double2 p = __ldg(a2 + c), q = __ldg(v2 + c);
                acc += p.x * q.x + p.y * q.y;
            }
            if ((cols & 1) && lane == 0) acc += a[cols - 1] * v[cols - 1];
        } else {
            for (int c = lane; c < cols; c += 32) acc += a[c] * v[c];
        }
    }
    acc = warp_sum(acc);
    double out = 0.0;
    if (r < rows) {
        out = alpha * acc + (y0 ? ys * y0[r] : 0.0);
        if (lane == 0) y[r] = out;
    }
    if (sq_part) {
        if (lane == 0) sh[w] = out * out;
        __syncthreads();
        if (threadIdx.x == 0) {
            double t = 0.0;
#pragma unroll
            for (int i = 0; i < 8; i++) t += sh[i];
            sq_part[blockIdx.x] = t;
        }
    }
}

inline int gemv_n(cudaStream_t st, const double* A, int lda, int rows, int cols, const double* v, const double* y0,
                  double ys, double alpha, double* y, double* sq_part = nullptr) {
    if (rows <= 0) return 0;
    const bool vec = ((lda & 1) == 0) && ((reinterpret_cast<uintptr_t>(A) & 15) == 0) &&
                     ((reinterpret_cast<uintptr_t>(v) & 15) == 0);
    const int grid = cdiv(rows, 8);
    if (vec)
        gemv_n_kernel<true><<<grid, 256, 0, st>>>(A, lda, rows, cols, v, y0, ys, alpha, y, sq_part);
    else
        gemv_n_kernel<false><<<grid, 256, 0, st>>>(A, lda, rows, cols, v, y0, ys, alpha, y, sq_part);
    LAUNCHED();
    return 0;
}
inline int gemv_n_blocks(int rows) { return cdiv(rows, 8); }

// ---------------------------------------------------------------------------------------------- gemv_t
// out[c] = ys * y0[c] + alpha * sum_r A[r, c] * v[r]   (two deterministic phases: row-chunk partials, then a
// column reduction; threads run along the contiguous dimension so every load is coalesced)
#define GEMVT_RC 64
__global__ void __launch_bounds__(256) gemv_t_partial_kernel(const double* __restrict__ A, int lda, int rows, int cols,
                                                             const double* __restrict__ v, double* __restrict__ part) {
    __shared__ double vs[GEMVT_RC];
    const int c = blockIdx.x * 256 + threadIdx.x;
    const int r0 = blockIdx.y * GEMVT_RC;
    const int nr = min(GEMVT_RC, rows - r0);
    if (threadIdx.x < GEMVT_RC) vs[threadIdx.x] = (threadIdx.x < nr) ? v[r0 + threadIdx.x] : 0.0;
    __syncthreads();
    if (c >= cols) return;
    const double* a = A + (size_t)r0 * lda + c;
    double acc = 0.0;
    int r = 0;
    for (; r + 8 <= nr; r += 8) {
        double t[8];
#pragma unroll
        for (int i = 0; i < 8; i++) t[i] = __ldg(a + (size_t)(r + i) * lda);
#pragma unroll
        for (int i = 0; i < 8; i++) acc += t[i] * vs[r + i];
    }
    for (; r < nr; r++) acc += __ldg(a + (size_t)r * lda) * vs[r];
    part[(size_t)blockIdx.y * cols + c] = acc;
}
__global__ void gemv_t_reduce_kernel(const double* __restrict__ part, int nchunk, int cols, const double* __restrict__ y0,
                                     double ys, double alpha, double* __restrict__ out) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= cols) return;
    double acc = 0.0;
    for (int k = 0; k < nchunk; k++) acc += part[(size_t)k * cols + c];
    out[c] = alpha * acc + (y0 ? ys * y0[c] : 0.0);
}
inline size_t gemv_t_scratch_doubles(int rows, int cols) { return (size_t)cdiv(rows, GEMVT_RC) * cols; }
inline int gemv_t(cudaStream_t st, const double* A, int lda, int rows, int cols, const double* v, const double* y0,
                  double ys, double alpha, double* out, double* scratch) {
    if (cols <= 0) return 0;
    const int nchunk = cdiv(rows, GEMVT_RC);
    dim3 grid(cdiv(cols, 256), nchunk);
    gemv_t_partial_kernel<<<grid, 256, 0, st>>>(A, lda, rows, cols, v, scratch);
    LAUNCHED();
    gemv_t_reduce_kernel<<<cdiv(cols, 256), 256, 0, st>>>(scratch, nchunk, cols, y0, ys, alpha, out);
    LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------- transpose
// out (cols x rows, ldo) = in (rows x cols, ldi)^T, 32x32 smem tiles, padded against bank conflicts
__global__ void transpose_kernel(const double* __restrict__ in, int ldi, int rows, int cols, double* __restrict__ out,
                                 int ldo) {
    __shared__ double t[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = by + i, c = bx + threadIdx.x;
        if (r < rows && c < cols) t[i][threadIdx.x] = in[(size_t)r * ldi + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = bx + i, r = by + threadIdx.x;
        if (r < rows && c < cols) out[(size_t)c * ldo + r] = t[threadIdx.x][i];
    }
}
inline int transpose(cudaStream_t st, const double* in, int ldi, int rows, int cols, double* out, int ldo) {
    if (rows <= 0 || cols <= 0) return 0;
    dim3 grid(cdiv(cols, 32), cdiv(rows, 32)), blk(32, 8);
    transpose_kernel<<<grid, blk, 0, st>>>(in, ldi, rows, cols, out, ldo);
    LAUNCHED();
    return 0;
}

// sum over a short array of per-block partials -> out[0] (deterministic), optional sqrt
__global__ void sum_partials_kernel(const double* __restrict__ p, int n, double* out, int do_sqrt) {
    __shared__ double sh[33];
    double a = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) a += p[i];
    a = block_sum(a, sh);
    if (threadIdx.x == 0) out[0] = do_sqrt ? sqrt(a) : a;
}

}  // namespace b200
