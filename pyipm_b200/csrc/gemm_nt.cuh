// gemm_nt.cuh -- fp64 "A * diag(w) * B^T" contraction on the tensor cores (DMMA.8x8x4 on sm_100a).
//
//   C[i, j] = beta * Cin[i, j] + [i == j] * (dadd[i] + shift) + sum_t alpha_t * sum_k A_t[i, k] * w_t[k] * B_t[j, k]
//
// Both operands are row-major with the contraction index contiguous (K-major), which is exactly the layout of
// the reference's transposed Jacobians (dci is D x N, pyipm.py:138): the Lagrangian-Hessian terms
// Ut diag(lda_e) Ut', Vt diag(lda_i) Vt', the condensation dci diag(sigma) dci' and the LDL^T trailing update
// W L' all map onto this one kernel.
//
// Modes
//   GEMM_UPPER_MIRROR: C is n x n symmetric.  Only tiles on/above the diagonal are computed, only elements
//                      i <= j are authoritative, each is mirrored to (j, i) => bitwise symmetric output, and
//                      Cin is read from its UPPER triangle only (the reference symmetrises triu(d2L),
//                      pyipm.py:785,827,843).
//   GEMM_LOWER_ONLY:   tiles on/below the diagonal, no mirror (LDL^T trailing update).
//   GEMM_FULL:         all tiles of an n x m result.
//
// Tiling: 128 x 128 x 16 CTA tile, 8 warps (2 x 4), 64 x 32 warp tile = 8 x 4 DMMA fragments, 4-stage cp.async
// pipeline (16-byte LDGSTS, zero-filled edges), smem row stride 20 doubles => every fragment LDS.64 is the
// minimum two wavefronts.  Algorithmic FLOPs per launch: 2 * (#computed tile elements) * sum_t K_t.
#pragma once
#include <cuda.h>
#include <cstdint>
#include "common.cuh"

namespace b200 {

enum { GEMM_UPPER_MIRROR = 0, GEMM_LOWER_ONLY = 1, GEMM_FULL = 2, GEMM_BC_LOWER = 3 };

struct GemmTerm {
    const double* A;
    const double* B;
    const double* w;   // may be null (all ones)
    int lda, ldb, K;
    double alpha;
};
struct GemmArgs {
    double* C;
    const double* Cin;    // may be null
    const double* dadd;   // may be null
    int ldc, ldcin;
    int n, m;             // rows, cols of C
    double beta, shift;
    int mode, nterms;
    GemmTerm t[3];
    // GEMM_BC_LOWER: C is a local piece of a 2-D block-cyclic matrix (block bc_b, grid bc_P x bc_Q, this rank at
    // (bc_p, bc_q), local block offsets bc_li0 / bc_lj0 of C's origin): a tile is computed iff its global block row >=
    // its global block column (tiles never straddle blocks: bc_b is a multiple of the tile size).
    int bc_b, bc_P, bc_Q, bc_p, bc_q, bc_li0, bc_lj0;
    const int* ctrl;      // optional LDL^T control block: ctrl[4] != 0 => the factorisation was abandoned, do nothing
    int max_ctas;         // > 0 (GEMM_LOWER_ONLY, square): persistent launch with at most this many CTAs, so that the
                          // rest of the SMs stay free for the latency-critical kernels of a concurrent stream
    int persistent_tiles; // set by gemm_nt()
};

constexpr int G_BM = 128, G_BN = 128, G_BK = 32, G_LDS = 36, G_STAGES = 3;
constexpr int G_SMEM = G_STAGES * ((G_BM + G_BN) * G_LDS + G_BK) * 8;

__device__ __forceinline__ void cp_async16(double* smem_dst, const double* gsrc, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async8(double* smem_dst, const double* gsrc, int bytes) {
    unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;\n" ::"r"(s), "l"(gsrc), "r"(bytes));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}
__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// one operand tile: 128 rows x 16 doubles, 1024 16-byte chunks spread over the CTA; rows/k beyond the edge zero-fill
template <int THREADS>
__device__ __forceinline__ void g_load_tile(double* sdst, const double* __restrict__ G, int ld, int row0, int nrows,
                                            int k0, int K, int tid) {
    constexpr int CPR = G_BK / 2;                      // 16-byte chunks per row
#pragma unroll
    for (int i = 0; i < G_BM * CPR / THREADS; i++) {
        const int c = tid + THREADS * i;
        const int row = c / CPR, kc = (c % CPR) * 2;
        const int gr = row0 + row, gk = k0 + kc;
        int bytes = 0;
        const double* src = G;
        if (gr < nrows && gk < K) {
            bytes = min(16, (K - gk) * 8);
            src = G + (size_t)gr * ld + gk;
        }
        cp_async16(sdst + row * G_LDS + kc, src, bytes);
    }
}
// the diagonal weights of one k-tile (16 doubles) ride the same cp.async pipeline as the operands, so the main
// loop never waits on a global load (8-byte copies: w only needs natural alignment)
__device__ __forceinline__ void g_load_w(double* sdst, const double* __restrict__ w, int k0, int K, int tid) {
    if (tid < G_BK) {
        const int k = k0 + tid;
        const bool ok = k < K;
        cp_async8(sdst + tid, ok ? w + k : w, ok ? 8 : 0);
    }
}

// Warp layout NWM x NWN over the 128 x 128 CTA tile (warp tile = (128/NWM) x (128/NWN)).  Measured on B200 at config 3:
// 2 x 4 (256 threads, 64 x 32 warp tiles) 26.0 TF/s, 4 x 4 (512 threads, 32 x 32 warp tiles) 24.3 TF/s.
constexpr int G_NWM = 2, G_NWN = 4, G_THREADS = 32 * G_NWM * G_NWN;
constexpr int G_MT = G_BM / (8 * G_NWM), G_NT = G_BN / (8 * G_NWN);
__device__ __forceinline__ void gemm_nt_tile(const GemmArgs& a, const int ti, const int tj, double* g_smem) {
    if (a.mode == GEMM_UPPER_MIRROR && ti > tj) return;
    if (a.mode == GEMM_LOWER_ONLY && ti < tj) return;
    if (a.mode == GEMM_BC_LOWER) {
        const int I = (a.bc_li0 + (ti * G_BM) / a.bc_b) * a.bc_P + a.bc_p;
        const int J = (a.bc_lj0 + (tj * G_BN) / a.bc_b) * a.bc_Q + a.bc_q;
        if (I < J) return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp / G_NWN, wn = warp % G_NWN;
    constexpr int WTM = G_BM / G_NWM, WTN = G_BN / G_NWN;
    const int g = lane >> 2, tg = lane & 3;
    const int row0 = ti * G_BM, col0 = tj * G_BN;
    double* As = g_smem;
    double* Bs = g_smem + G_STAGES * G_BM * G_LDS;
    double* Wsm = g_smem + G_STAGES * (G_BM + G_BN) * G_LDS;

    double acc[G_MT][G_NT][2];
#pragma unroll
    for (int i = 0; i < G_MT; i++)
#pragma unroll
        for (int j = 0; j < G_NT; j++) acc[i][j][0] = acc[i][j][1] = 0.0;

    for (int t = 0; t < a.nterms; t++) {
        const GemmTerm T = a.t[t];
        const int nk = (T.K + G_BK - 1) / G_BK;
        // prologue
#pragma unroll
        for (int s = 0; s < G_STAGES - 1; s++) {
            if (s < nk) {
                g_load_tile<G_THREADS>(As + s * G_BM * G_LDS, T.A, T.lda, row0, a.n, s * G_BK, T.K, tid);
                g_load_tile<G_THREADS>(Bs + s * G_BN * G_LDS, T.B, T.ldb, col0, a.m, s * G_BK, T.K, tid);
                if (T.w) g_load_w(Wsm + s * G_BK, T.w, s * G_BK, T.K, tid);
            }
            cp_async_commit();
        }
        for (int kt = 0; kt < nk; kt++) {
            cp_async_wait<G_STAGES - 2>();
            __syncthreads();
            {   // refill the stage consumed at iteration kt-1
                const int kn = kt + G_STAGES - 1;
                if (kn < nk) {
                    const int s = kn % G_STAGES;
                    g_load_tile<G_THREADS>(As + s * G_BM * G_LDS, T.A, T.lda, row0, a.n, kn * G_BK, T.K, tid);
                    g_load_tile<G_THREADS>(Bs + s * G_BN * G_LDS, T.B, T.ldb, col0, a.m, kn * G_BK, T.K, tid);
                    if (T.w) g_load_w(Wsm + s * G_BK, T.w, kn * G_BK, T.K, tid);
                }
                cp_async_commit();
            }
            const int s = kt % G_STAGES;
            const double* ws = Wsm + s * G_BK + tg;
            const double* as = As + s * G_BM * G_LDS + (wm * WTM + g) * G_LDS + tg;
            const double* bs = Bs + s * G_BN * G_LDS + (wn * WTN + g) * G_LDS + tg;
#pragma unroll
            for (int kk = 0; kk < G_BK / 4; kk++) {
                double af[G_MT], bf[G_NT];
#pragma unroll
                for (int mt = 0; mt < G_MT; mt++) af[mt] = as[mt * 8 * G_LDS + kk * 4];
                const double wk = T.w ? ws[kk * 4] * T.alpha : T.alpha;
#pragma unroll
                for (int nt = 0; nt < G_NT; nt++) bf[nt] = bs[nt * 8 * G_LDS + kk * 4] * wk;
#pragma unroll
                for (int mt = 0; mt < G_MT; mt++)
#pragma unroll
                    for (int nt = 0; nt < G_NT; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
            }
        }
        cp_async_wait<0>();
        __syncthreads();
    }

    // epilogue.  Per 8-row fragment group the Cin values are fetched first (4 independent 16-byte loads in flight
    // per thread), then combined and stored: the short-K launches of the LDL^T are epilogue-bound otherwise.
    const bool diag_tile = (ti == tj) && (a.mode == GEMM_UPPER_MIRROR || a.mode == GEMM_LOWER_ONLY);
    const bool vec_ok = ((a.ldc & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.C) & 15) == 0) &&
                        (a.Cin == nullptr || (((a.ldcin & 1) == 0) && ((reinterpret_cast<uintptr_t>(a.Cin) & 15) == 0)));
    const bool interior = vec_ok && !diag_tile && (row0 + G_BM <= a.n) && (col0 + G_BN <= a.m);
    if (interior) {
#pragma unroll
        for (int mt = 0; mt < G_MT; mt++) {
            const int i = row0 + wm * WTM + mt * 8 + g;
            const int jb = col0 + wn * WTN + tg * 2;
            double2 cin[G_NT];
            if (a.Cin) {
#pragma unroll
                for (int nt = 0; nt < G_NT; nt++)
                    cin[nt] = *reinterpret_cast<const double2*>(a.Cin + (size_t)i * a.ldcin + jb + nt * 8);
            }
#pragma unroll
            for (int nt = 0; nt < G_NT; nt++) {
                double2 v = make_double2(acc[mt][nt][0], acc[mt][nt][1]);
                if (a.Cin) { v.x += a.beta * cin[nt].x; v.y += a.beta * cin[nt].y; }
                const int j = jb + nt * 8;
                *reinterpret_cast<double2*>(a.C + (size_t)i * a.ldc + j) = v;
                if (a.mode == GEMM_UPPER_MIRROR) {
                    a.C[(size_t)j * a.ldc + i] = v.x;
                    a.C[(size_t)(j + 1) * a.ldc + i] = v.y;
                }
            }
        }
        return;
    }
#pragma unroll
    for (int mt = 0; mt < G_MT; mt++) {
        const int i = row0 + wm * WTM + mt * 8 + g;
        if (i >= a.n) continue;
        double cin[G_NT][2];
        if (a.Cin) {
#pragma unroll
            for (int nt = 0; nt < G_NT; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = col0 + wn * WTN + nt * 8 + tg * 2 + e;
                    const bool need = (j < a.m) && !(a.mode == GEMM_UPPER_MIRROR && diag_tile && i > j);
                    cin[nt][e] = need ? a.Cin[(size_t)i * a.ldcin + j] : 0.0;
                }
        }
#pragma unroll
        for (int nt = 0; nt < G_NT; nt++) {
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = col0 + wn * WTN + nt * 8 + tg * 2 + e;
                if (j >= a.m) continue;
                if (a.mode == GEMM_UPPER_MIRROR && diag_tile && i > j) continue;
                double v = acc[mt][nt][e];
                if (a.Cin) v += a.beta * cin[nt][e];
                if (i == j) v += a.shift + (a.dadd ? a.dadd[i] : 0.0);
                a.C[(size_t)i * a.ldc + j] = v;
                if (a.mode == GEMM_UPPER_MIRROR && i != j) a.C[(size_t)j * a.ldc + i] = v;
            }
        }
    }
}

__global__ void __launch_bounds__(G_THREADS, 1) gemm_nt_dmma_kernel(const GemmArgs a) {
    extern __shared__ __align__(16) double g_smem[];
    if (a.ctrl) {   // one read per CTA: the flag can be raised while this kernel runs (look-ahead streams)
        __shared__ int s_abort;
        if (threadIdx.x == 0) s_abort = *reinterpret_cast<const volatile int*>(a.ctrl + 4);
        __syncthreads();
        if (s_abort) return;
    }
    const int trc = (threadIdx.x == 0 && ((blockIdx.x == 0 && blockIdx.y == 0) || (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1)))
                        ? trace_begin(TR_DMMA, a.ctrl) : -1;
    if (a.persistent_tiles > 0) {
        // lower-triangular tile t -> (r, c) with c <= r, row by row
        for (int t = blockIdx.x; t < a.persistent_tiles; t += gridDim.x) {
            int r = (int)((sqrt(8.0 * (double)t + 1.0) - 1.0) * 0.5);
            while ((r + 1) * (r + 2) / 2 <= t) r++;
            while (r * (r + 1) / 2 > t) r--;
            gemm_nt_tile(a, r, t - r * (r + 1) / 2, g_smem);
            __syncthreads();
        }
    } else {
        gemm_nt_tile(a, blockIdx.y, blockIdx.x, g_smem);
    }
    trace_end(trc);
}

// scalar reference kernel: any size / alignment (tiny problems, odd leading dimensions, test cross-check)
__global__ void gemm_nt_simple_kernel(const GemmArgs a) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i >= a.n || j >= a.m) return;
    if (a.mode == GEMM_UPPER_MIRROR && i > j) return;
    if (a.mode == GEMM_LOWER_ONLY && i < j) return;
    if (a.mode == GEMM_BC_LOWER) {
        const int I = (a.bc_li0 + i / a.bc_b) * a.bc_P + a.bc_p, J = (a.bc_lj0 + j / a.bc_b) * a.bc_Q + a.bc_q;
        if (I < J) return;
    }
    double v = 0.0;
    for (int t = 0; t < a.nterms; t++) {
        const GemmTerm T = a.t[t];
        const double* pa = T.A + (size_t)i * T.lda;
        const double* pb = T.B + (size_t)j * T.ldb;
        double s = 0.0;
        if (T.w)
            for (int k = 0; k < T.K; k++) s += pa[k] * (pb[k] * T.w[k]);
        else
            for (int k = 0; k < T.K; k++) s += pa[k] * pb[k];
        v += T.alpha * s;
    }
    if (a.Cin) v += a.beta * a.Cin[(size_t)i * a.ldcin + j];
    if (i == j) v += a.shift + (a.dadd ? a.dadd[i] : 0.0);
    a.C[(size_t)i * a.ldc + j] = v;
    if (a.mode == GEMM_UPPER_MIRROR && i != j) a.C[(size_t)j * a.ldc + i] = v;
}

inline bool gemm_nt_sub_uses_tiles(int rows, const double* A, int lda, const double* B, int ldb) {
    return !(lda & 1) && !(ldb & 1) && !(reinterpret_cast<uintptr_t>(A) & 15) && !(reinterpret_cast<uintptr_t>(B) & 15) && rows >= 48;
}
inline bool gemm_nt_can_dmma(const GemmArgs& a) {
    for (int t = 0; t < a.nterms; t++) {
        const GemmTerm& T = a.t[t];
        if ((T.lda & 1) || (T.ldb & 1)) return false;
        if ((reinterpret_cast<uintptr_t>(T.A) & 15) || (reinterpret_cast<uintptr_t>(T.B) & 15)) return false;
    }
    return true;
}

inline int gemm_nt(cudaStream_t st, const GemmArgs& a, bool force_simple = false) {
    if (a.n <= 0 || a.m <= 0) return 0;
    const bool big = (a.n >= 48 && a.m >= 48);
    if (!force_simple && big && gemm_nt_can_dmma(a)) {
        dim3 grid(cdiv(a.m, G_BN), cdiv(a.n, G_BM));
        GemmArgs b = a;
        b.persistent_tiles = 0;
        if (a.max_ctas > 0 && a.mode == GEMM_LOWER_ONLY && a.n == a.m) {
            const int nt = (int)grid.y * ((int)grid.y + 1) / 2;
            if (nt > a.max_ctas) {
                b.persistent_tiles = nt;
                grid = dim3(a.max_ctas, 1);
            }
        }
        gemm_nt_dmma_kernel<<<grid, G_THREADS, G_SMEM, st>>>(b);
        LAUNCHED();
    } else {
        dim3 blk(32, 8), grid(cdiv(a.m, 32), cdiv(a.n, 8));
        gemm_nt_simple_kernel<<<grid, blk, 0, st>>>(a);
        LAUNCHED();
    }
    return 0;
}

// ---------------------------------------------------------------------------------------------------------
// Small-tile variant for the skinny updates on the LDL^T critical path:  C -= A * B^T  in place, C is rows x cols
// with cols <= a few hundred (the remaining columns of an outer panel, K = 64, or the next panel's 256 columns,
// K = 256).  With 128 x 128 tiles those launches have fewer CTAs than the GPU has SMs and every CTA is long;
// 64 x 64 tiles (128 threads, 4 warps x (16 x 64), 3-stage cp.async over K-chunks of 32, two CTAs per SM) spread
// the same work over all SMs and cut the per-launch latency ~3x.
constexpr int S_BM = 64, S_BN = 64, S_BK = 32, S_LDS = 36, S_STAGES = 3;
constexpr int S_SMEM = S_STAGES * (S_BM + S_BN) * S_LDS * 8;

__device__ __forceinline__ void s_load_tile(double* sdst, const double* __restrict__ G, int ld, int row0, int nrows,
                                            int k0, int K, int tid) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = tid + 128 * i;
        const int row = c >> 4, kc = (c & 15) * 2;
        const int gr = row0 + row, gk = k0 + kc;
        int bytes = 0;
        const double* src = G;
        if (gr < nrows && gk < K) {
            bytes = min(16, (K - gk) * 8);
            src = G + (size_t)gr * ld + gk;
        }
        cp_async16(sdst + row * S_LDS + kc, src, bytes);
    }
}

__global__ void __launch_bounds__(128, 2) gemm_nt_sub64_kernel(double* __restrict__ C, int ldc, int rows, int cols,
                                                               const double* __restrict__ A, int lda,
                                                               const double* __restrict__ B, int ldb, int K,
                                                               const int* __restrict__ ctrl, int skip00) {
    extern __shared__ __align__(16) double s_smem[];
    if (skip00 && blockIdx.x == 0 && blockIdx.y == 0) return;   // that tile was already updated on the chain stream
    const int trc = (threadIdx.x == 0 && ((blockIdx.x == 1 && blockIdx.y == 0) || (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1)))
                        ? trace_begin(TR_SUB64, ctrl) : -1;
    if (ctrl) {
        __shared__ int s_abort;
        if (threadIdx.x == 0) s_abort = *reinterpret_cast<const volatile int*>(ctrl + 4);
        __syncthreads();
        if (s_abort) return;
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int row0 = blockIdx.y * S_BM, col0 = blockIdx.x * S_BN;
    double* As = s_smem;
    double* Bs = s_smem + S_STAGES * S_BM * S_LDS;
    double acc[2][8][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    const int nk = (K + S_BK - 1) / S_BK;
#pragma unroll
    for (int s = 0; s < S_STAGES - 1; s++) {
        if (s < nk) {
            s_load_tile(As + s * S_BM * S_LDS, A, lda, row0, rows, s * S_BK, K, tid);
            s_load_tile(Bs + s * S_BN * S_LDS, B, ldb, col0, cols, s * S_BK, K, tid);
        }
        cp_async_commit();
    }
    for (int kt = 0; kt < nk; kt++) {
        cp_async_wait<S_STAGES - 2>();
        __syncthreads();
        {
            const int kn = kt + S_STAGES - 1;
            if (kn < nk) {
                const int s = kn % S_STAGES;
                s_load_tile(As + s * S_BM * S_LDS, A, lda, row0, rows, kn * S_BK, K, tid);
                s_load_tile(Bs + s * S_BN * S_LDS, B, ldb, col0, cols, kn * S_BK, K, tid);
            }
            cp_async_commit();
        }
        const int s = kt % S_STAGES;
        const double* as = As + s * S_BM * S_LDS + (warp * 16 + g) * S_LDS + tg;
        const double* bs = Bs + s * S_BN * S_LDS + g * S_LDS + tg;
#pragma unroll
        for (int kk = 0; kk < S_BK / 4; kk++) {
            double af[2], bf[8];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) af[mt] = as[mt * 8 * S_LDS + kk * 4];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) bf[nt] = bs[nt * 8 * S_LDS + kk * 4];
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    cp_async_wait<0>();
    const bool interior = (row0 + S_BM <= rows) && (col0 + S_BN <= cols) && ((ldc & 1) == 0) &&
                          ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        const int i = row0 + warp * 16 + mt * 8 + g;
        if (interior) {
            double2 cin[8];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) cin[nt] = *reinterpret_cast<const double2*>(C + (size_t)i * ldc + col0 + nt * 8 + tg * 2);
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                cin[nt].x -= acc[mt][nt][0];
                cin[nt].y -= acc[mt][nt][1];
                *reinterpret_cast<double2*>(C + (size_t)i * ldc + col0 + nt * 8 + tg * 2) = cin[nt];
            }
        } else if (i < rows) {
#pragma unroll
            for (int nt = 0; nt < 8; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = col0 + nt * 8 + tg * 2 + e;
                    if (j < cols) C[(size_t)i * ldc + j] -= acc[mt][nt][e];
                }
        }
    }
    trace_end(trc);
}


// ------------------------------------------------------------------------------------------- TMA-staged 64 x 64 update
// Same contract as gemm_nt_sub64_kernel (C -= A B', 64 x 64 tiles, 128 threads, fp64 DMMA), but the operand tiles are
// staged by the TENSOR MEMORY ACCELERATOR: one elected thread issues cp.async.bulk.tensor.2d (SASS: UTMALDG) against two
// cuTensorMap descriptors that describe the WHOLE operand matrices (the L panels inside the KKT matrix, the W = L D
// scratch panels), boxes of 64 rows x 16 doubles (128 bytes) with the 128-byte swizzle, completion on per-stage
// mbarriers.  The swizzle (16-byte chunk index XOR row mod 8) makes the DMMA fragment reads -- 8 rows x 4 consecutive
// doubles per quarter-warp -- hit every bank exactly twice, the minimum for 64-bit accesses, without padding the rows.
// No per-thread address arithmetic or LDGSTS issue for the loads: the four warps only compute.
constexpr int T_STAGES = 3;
constexpr int T_BOX_BYTES = 64 * 128;                     // one box: 64 rows x 16 doubles
constexpr int T_STAGE_BYTES = 4 * T_BOX_BYTES;            // A: two k-halves, B: two k-halves (K step = 32)
constexpr int T_SMEM = T_STAGES * T_STAGE_BYTES + 1024;   // + alignment slack (boxes must be 1024-byte aligned)

struct TmaMat {              // host-side handle of one operand matrix
    CUtensorMap map;         // 2-D tensor [rows][ld] of doubles, box 64 x 16, SWIZZLE_128B
    const double* base = nullptr;
    int ld = 0;
    bool ok = false;
};

__device__ __forceinline__ uint32_t t_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void t_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void t_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool t_mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void t_tma_load_2d(uint32_t dst, const CUtensorMap* map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n" ::"r"(dst),
        "l"(map), "r"(c0), "r"(c1), "r"(bar)
        : "memory");
}
// element (row, k) of a swizzled 64 x 16 box
__device__ __forceinline__ double t_box_ld(const unsigned char* box, int row, int k) {
    return *reinterpret_cast<const double*>(box + row * 128 + ((((k >> 1) ^ (row & 7)) << 4) | ((k & 1) << 3)));
}

__global__ void __launch_bounds__(128, 2) gemm_nt_sub64_tma_kernel(double* __restrict__ C, int ldc, int rows, int cols,
                                                                   const __grid_constant__ CUtensorMap mapA, int a_row0,
                                                                   int a_col0, const __grid_constant__ CUtensorMap mapB,
                                                                   int b_row0, int b_col0, int K,
                                                                   const int* __restrict__ ctrl, int skip00) {
    extern __shared__ __align__(1024) unsigned char t_smem_raw[];
    __shared__ __align__(8) uint64_t bar_full[T_STAGES];
    if (skip00 && blockIdx.x == 0 && blockIdx.y == 0) return;   // that tile was already updated on the chain stream
    if (ctrl) {
        __shared__ int s_abort;
        if (threadIdx.x == 0) s_abort = *reinterpret_cast<const volatile int*>(ctrl + 4);
        __syncthreads();
        if (s_abort) return;
    }
    const int trc = (threadIdx.x == 0 && ((blockIdx.x == 1 && blockIdx.y == 0) || (blockIdx.x == gridDim.x - 1 && blockIdx.y == gridDim.y - 1)))
                        ? trace_begin(TR_SUB64, ctrl) : -1;
    unsigned char* sm = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(t_smem_raw) + 1023) & ~(uintptr_t)1023);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int row0 = blockIdx.y * S_BM, col0 = blockIdx.x * S_BN;
    const int nk = K / 32;
    if (tid == 0) {
        for (int s = 0; s < T_STAGES; s++) t_mbar_init(t_smem_u32(&bar_full[s]), 1);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int kt) {
        const int s = kt % T_STAGES;
        const uint32_t bar = t_smem_u32(&bar_full[s]);
        const uint32_t dst = t_smem_u32(sm + s * T_STAGE_BYTES);
        t_mbar_expect_tx(bar, T_STAGE_BYTES);
        const int k0 = kt * 32;
        t_tma_load_2d(dst, &mapA, a_col0 + k0, a_row0 + row0, bar);
        t_tma_load_2d(dst + T_BOX_BYTES, &mapA, a_col0 + k0 + 16, a_row0 + row0, bar);
        t_tma_load_2d(dst + 2 * T_BOX_BYTES, &mapB, b_col0 + k0, b_row0 + col0, bar);
        t_tma_load_2d(dst + 3 * T_BOX_BYTES, &mapB, b_col0 + k0 + 16, b_row0 + col0, bar);
    };
    if (tid == 0)
        for (int s = 0; s < T_STAGES - 1 && s < nk; s++) issue(s);
    double acc[2][8][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    for (int kt = 0; kt < nk; kt++) {
        const int s = kt % T_STAGES;
        const uint32_t bar = t_smem_u32(&bar_full[s]);
        const uint32_t ph = (uint32_t)(kt / T_STAGES) & 1u;
        {
            const long long t0 = clock64();
            while (!t_mbar_try(bar, ph)) {
                if (clock64() - t0 > 4000000000LL) break;        // bounded: never hang the device
            }
        }
        __syncthreads();         // everybody is past stage (kt - 1): its buffer may be refilled
        if (tid == 0 && kt + T_STAGES - 1 < nk) issue(kt + T_STAGES - 1);
        const unsigned char* st = sm + s * T_STAGE_BYTES;
#pragma unroll
        for (int kk = 0; kk < 8; kk++) {
            const unsigned char* ab = st + (kk >> 2) * T_BOX_BYTES;
            const unsigned char* bb = st + (2 + (kk >> 2)) * T_BOX_BYTES;
            const int k = (kk & 3) * 4 + tg;
            double af[2], bf[8];
#pragma unroll
            for (int mt = 0; mt < 2; mt++) af[mt] = t_box_ld(ab, warp * 16 + mt * 8 + g, k);
#pragma unroll
            for (int nt = 0; nt < 8; nt++) bf[nt] = t_box_ld(bb, nt * 8 + g, k);
#pragma unroll
            for (int mt = 0; mt < 2; mt++)
#pragma unroll
                for (int nt = 0; nt < 8; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
        }
    }
    const bool interior = (row0 + S_BM <= rows) && (col0 + S_BN <= cols) && ((ldc & 1) == 0) &&
                          ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
#pragma unroll
    for (int mt = 0; mt < 2; mt++) {
        const int i = row0 + warp * 16 + mt * 8 + g;
        if (interior) {
            double2 cin[8];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) cin[nt] = *reinterpret_cast<const double2*>(C + (size_t)i * ldc + col0 + nt * 8 + tg * 2);
#pragma unroll
            for (int nt = 0; nt < 8; nt++) {
                cin[nt].x -= acc[mt][nt][0];
                cin[nt].y -= acc[mt][nt][1];
                *reinterpret_cast<double2*>(C + (size_t)i * ldc + col0 + nt * 8 + tg * 2) = cin[nt];
            }
        } else if (i < rows) {
#pragma unroll
            for (int nt = 0; nt < 8; nt++)
#pragma unroll
                for (int e = 0; e < 2; e++) {
                    const int j = col0 + nt * 8 + tg * 2 + e;
                    if (j < cols) C[(size_t)i * ldc + j] -= acc[mt][nt][e];
                }
        }
    }
    trace_end(trc);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no -lcuda at link time)
typedef CUresult (*t_encode_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline t_encode_fn tma_encode_fn() {
    static t_encode_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<t_encode_fn>(p);
        cudaGetLastError();
    }
    return fn;
}
// describe a row-major matrix of doubles (rows x ld) for 64 x 16 boxes with the 128-byte swizzle
inline int tma_describe(TmaMat& m, const double* base, int rows, int ld) {
    m.ok = false;
    m.base = base;
    m.ld = ld;
    t_encode_fn fn = tma_encode_fn();
    if (!fn || (reinterpret_cast<uintptr_t>(base) & 15) || (ld & 1)) return 0;      // silently: the cp.async kernel is used
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 8};
    const cuuint32_t box[2] = {16, 64};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(&m.map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    m.ok = (r == CUDA_SUCCESS);
    return 0;
}

// C (rows x cols, in place) -= A (rows x K) * B (cols x K)^T
inline int gemm_nt_sub(cudaStream_t st, double* C, int ldc, int rows, int cols, const double* A, int lda, const double* B,
                       int ldb, int K, const int* ctrl = nullptr, int skip00 = 0, const TmaMat* ta = nullptr,
                       const TmaMat* tb = nullptr) {
    if (rows <= 0 || cols <= 0) return 0;
    const bool ok = !(lda & 1) && !(ldb & 1) && !(reinterpret_cast<uintptr_t>(A) & 15) && !(reinterpret_cast<uintptr_t>(B) & 15);
    if (ok && rows >= 48) {
        dim3 grid(cdiv(cols, S_BN), cdiv(rows, S_BM));
        if (ta && tb && ta->ok && tb->ok && (K % 32) == 0 && ta->ld == lda && tb->ld == ldb) {
            // operands through the tensor memory accelerator: coordinates of the tile origin inside the described matrices
            const size_t offa = (size_t)(A - ta->base), offb = (size_t)(B - tb->base);
            const int a_row0 = (int)(offa / (size_t)lda), a_col0 = (int)(offa % (size_t)lda);
            const int b_row0 = (int)(offb / (size_t)ldb), b_col0 = (int)(offb % (size_t)ldb);
            if ((a_col0 & 1) == 0 && (b_col0 & 1) == 0) {
                gemm_nt_sub64_tma_kernel<<<grid, 128, T_SMEM, st>>>(C, ldc, rows, cols, ta->map, a_row0, a_col0, tb->map, b_row0,
                                                                    b_col0, K, ctrl, skip00);
                LAUNCHED();
                return 0;
            }
        }
        gemm_nt_sub64_kernel<<<grid, 128, S_SMEM, st>>>(C, ldc, rows, cols, A, lda, B, ldb, K, ctrl, skip00);
        LAUNCHED();
        return 0;
    }
    if (skip00) return fail_msg("gemm_nt_sub: skip00 needs the 64 x 64 tile kernel");
    GemmArgs u{};
    u.C = C; u.ldc = ldc; u.Cin = C; u.ldcin = ldc; u.n = rows; u.m = cols; u.beta = 1.0; u.mode = GEMM_FULL; u.nterms = 1;
    u.t[0] = GemmTerm{A, B, nullptr, lda, ldb, K, -1.0};
    u.ctrl = ctrl;
    return gemm_nt(st, u);
}

// computed tile elements x 2 x sum K  (what one launch is credited with in the roofline)
inline double gemm_nt_flops(const GemmArgs& a) {
    double ksum = 0;
    for (int t = 0; t < a.nterms; t++) ksum += a.t[t].K;
    double elems = (a.mode == GEMM_FULL) ? (double)a.n * a.m : 0.5 * (double)a.n * ((double)a.n + 1.0);
    return 2.0 * elems * ksum;
}

}  // namespace b200
