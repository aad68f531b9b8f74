// ozaki_i8.cuh -- fp64-accurate "A * diag(w) * A^T" (SYRK) on the 5th-generation tensor cores.
//
// tcgen05.mma has no f64 kind, so the two dense contractions of a Newton step -- the Lagrangian-Hessian terms
// Ut diag(lda_e) Ut' + Vt diag(lda_i) Vt' (pyipm.py:817-821) and the condensation dci diag(sigma) dci'
// (pyipm.py:498, 824-844 eliminated) -- are computed by an error-free transformation
// (Ozaki scheme): every row of the scaled operand  L[i, k] = A[i, k] * sqrt|alpha w_k|  is written as
//
//     L[i, k] = 2^e_i * sum_{p = 0..6}  t_p[i, k] * 2^(-7 - 8 p),      t_p in [-128, 127]  (int8)
//
// with BALANCED base-256 digits (|L| 2^-e_i < 1/2, so t_0 in [-64, 65]; 6 + 6*8 = 54 bits + rounding = 55 bits below
// the row's largest entry; balanced digits are zero-mean, so the truncated slice pairs carry no systematic bias).
// The slice products  sum_k t_p[i, k] u_q[j, k]  are EXACT in the int32 accumulators of tcgen05.mma.kind::i8
// (|t u| <= 2^14, 7 pairs per accumulator, K <= 18432), all pairs with p + q = d share one TMEM accumulator, and the
// seven diagonals d = 0..6 (28 slice pairs) are recombined in fp64:
//
//     C[i, j] = beta Cin[i, j] + [i == j] (dadd_i + shift) + 2^(e_i + e_j - 14) * sum_d acc_d[i, j] * 2^(-8 d).
//
// Negative weights (lda_e) are carried by a second operand R = L with the sign of alpha*w_k applied per column.
//
// Data movement: the slicing kernel writes the int8 slices PRE-TILED in the shared-memory image the tensor core
// reads (K-major, no swizzle: 8-row x 16-byte core matrices, [128-row block][32-byte k-block][slice][row group]
// [k chunk][row]), so a pipeline stage is filled by plain 1-D bulk copies (cp.async.bulk -> UBLKCP, completion on
// an mbarrier) without tensor maps.  Kernel roles: warp 0 = copy producer (+ TMEM allocation), warp 1 = one
// elected thread issuing tcgen05.mma, warps 2-5 = epilogue (tcgen05.ld -> fp64 Horner recombination -> global,
// mirrored to the lower triangle so the result is bitwise symmetric like the DMMA kernel's).
#pragma once
#include <cstdint>
#include <vector>
#include <algorithm>
#include "common.cuh"
#include "gemm_nt.cuh"

namespace b200 {

constexpr int OZ_NS = 7;                     // slices per operand (balanced base-256 digits)
constexpr int OZ_NPAIRS = OZ_NS * (OZ_NS + 1) / 2;   // slice pairs with p + q <= OZ_NS - 1
constexpr int OZ_KMAX = 18432;               // OZ_NS * K * 2^14 < 2^31
constexpr int OZ_KB = 32;                    // int8 elements (bytes) of K per pipeline stage = one MMA K step
constexpr int OZ_BM = 128;                   // rows per block (UMMA M)
constexpr int OZ_CHUNK = OZ_BM * OZ_KB;      // bytes of one slice of one row block for one k-block
constexpr int OZ_THREADS = 320;   // producer warp, MMA warp, 8 epilogue warps (two per TMEM lane quarter)
constexpr unsigned OZ_SMEM_BUDGET = 200 * 1024;

struct OzTerm { const double* A; const double* w; int lda, K, koff, sgn, vec; double alpha; };   // koff: multiple of OZ_KB; vec: 16-byte loads ok
struct OzSliceArgs {
    OzTerm t[3];
    int nterms, n, nkb;
    int8_t* L;
    int8_t* R;          // sign-carrying copy, written only for the k-blocks of terms with sgn != 0
    const double* sw;   // sw[koff + k] = copysign(sqrt|alpha w_k|, alpha w_k)   (oz_weight_kernel)
    int* rexp;
    int* err;           // err[0] |= 1: non-finite operand, |= 2: negative weight in a term announced as non-negative
    const int* ctrl;    // optional LDL^T control block: ctrl[4] != 0 => the factorisation was abandoned, do nothing
};

// sw = signed square roots of the column weights, once per call (instead of once per row in the slicing kernel)
__global__ void oz_weight_kernel(const OzSliceArgs a, double* sw) {
    if (a.ctrl && a.ctrl[4]) return;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= a.nkb * OZ_KB) return;
    double v = 0.0;
    for (int t = 0; t < a.nterms; t++) {
        const int k = idx - a.t[t].koff;
        if (k >= 0 && k < a.t[t].K) {
            const double wk = a.t[t].w ? a.t[t].w[k] * a.t[t].alpha : a.t[t].alpha;
            v = copysign(sqrt(fabs(wk)), wk);
            if (wk < 0.0 && !a.t[t].sgn) atomicOr(a.err, 2);
            if (!(fabs(wk) <= 1.7e308)) atomicOr(a.err, 1);
        }
    }
    sw[idx] = v;
}

// ------------------------------------------------------------------------------------------- slicing kernels
__device__ __forceinline__ double oz_scaled(const OzTerm& T, const double* __restrict__ sw, int row, int k, bool& neg) {
    const double wk = sw[T.koff + k];
    neg = wk < 0.0;
    return T.A[(size_t)row * T.lda + k] * fabs(wk);
}
// pass 1: one warp per row -> exponent e_i with |L[i, :]| * 2^-e_i < 1
__global__ void __launch_bounds__(128) oz_rowmax_kernel(const OzSliceArgs a) {
    if (a.ctrl && a.ctrl[4]) return;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int row = blockIdx.x * 4 + warp;
    if (row >= a.n) return;
    double m = 0.0;
    bool bad = false;
    for (int t = 0; t < a.nterms; t++) {
        const OzTerm T = a.t[t];
        const double* __restrict__ ar = T.A + (size_t)row * T.lda;
        const double* __restrict__ sw = a.sw + T.koff;
        for (int k = lane; k < T.K; k += 32) {
            const double v = fabs(ar[k] * sw[k]);
            if (!(v <= 1.7e308)) bad = true;      // inf / nan
            m = fmax(m, v);
        }
    }
    m = warp_max(m);
    bad = __any_sync(0xffffffffu, bad);
    if (lane == 0) {
        a.rexp[row] = (m > 0.0 && !bad) ? ilogb(m) + 2 : 0;     // |L| * 2^-e < 1/2
        if (bad) atomicOr(a.err, 1);
    }
}
// pass 2: CTA = (8-row group, k-range); thread = (row r of the group, 16-element k-chunk); 8 consecutive threads write
// one 128-byte core matrix per slice
__global__ void __launch_bounds__(256) oz_slice_kernel(const OzSliceArgs a) {
    if (a.ctrl && a.ctrl[4]) return;
    const int tid = threadIdx.x;
    const int row0 = blockIdx.x * 8;
    const int r = tid & 7;
    const int row = row0 + r;
    const int e = (row < a.n) ? a.rexp[row] : 0;
    const int rb = row0 / OZ_BM, g = (row0 % OZ_BM) / 8;
    const int nchunks = a.nkb * 2;
    const int per = (nchunks + gridDim.y - 1) / gridDim.y;
    const int c_lo = blockIdx.y * per, c_hi = min(nchunks, c_lo + per);
    for (int cc = c_lo + (tid >> 3); cc < c_hi; cc += 32) {
        const int k0 = cc * 16;
        int t = -1;
        for (int u = 0; u < a.nterms; u++)
            if (k0 >= a.t[u].koff && k0 < a.t[u].koff + ((a.t[u].K + OZ_KB - 1) & ~(OZ_KB - 1))) t = u;
        double v[16];
        unsigned negmask = 0;
#pragma unroll
        for (int b = 0; b < 16; b++) v[b] = 0.0;
        if (t >= 0 && row < a.n) {
            const OzTerm T = a.t[t];
            const int kl = k0 - T.koff;
            const double* __restrict__ ar = T.A + (size_t)row * T.lda + kl;
            const double* __restrict__ sw = a.sw + k0;
            if (T.vec && kl + 16 <= T.K) {
#pragma unroll
                for (int b = 0; b < 16; b += 2) {
                    const double2 x = *reinterpret_cast<const double2*>(ar + b);
                    const double2 w = *reinterpret_cast<const double2*>(sw + b);
                    v[b] = x.x * fabs(w.x);
                    v[b + 1] = x.y * fabs(w.y);
                    if (w.x < 0.0) negmask |= 1u << b;
                    if (w.y < 0.0) negmask |= 2u << b;
                }
            } else {
#pragma unroll
                for (int b = 0; b < 16; b++) {
                    if (kl + b < T.K) {
                        const double w = sw[b];
                        v[b] = ar[b] * fabs(w);
                        if (w < 0.0) negmask |= 1u << b;
                    }
                }
            }
        }
        const bool write_r = (t >= 0) && a.t[t].sgn;
        // digits of L (and, for signed terms, of the column-sign-flipped copy R: negating a digit -128 would overflow,
        // so R gets the digits of -v)
        const size_t base = ((size_t)(rb * a.nkb + (cc >> 1)) * OZ_NS) * OZ_CHUNK + g * 256 + (cc & 1) * 128 + r * 16;
        unsigned wl[OZ_NS][4], wr[OZ_NS][4];
        // fixed point: X = rint(v * 2^(55 - e)), |X| < 2^54, then balanced base-256 digits d_i in [-128, 127] (d_6 in
        // [-64, 64]) with X = sum_i d_i 256^i.  Adding the bias 128 * 256^i at every position turns them into the ordinary
        // base-256 digits of Y = X + 0x00808080 80808080, so ALL SEVEN digits come from one 64-bit add, and the byte that is
        // stored (the two's complement of d_i) is byte_i(Y) ^ 0x80; digit i is slice p = 6 - i.  The bytes of four
        // consecutive elements are gathered into one 32-bit word per slice with three PRMTs.  (2^(55-e) is applied as two
        // exact power-of-two factors: a single one would overflow for rows whose largest entry is below 2^-968.)
        const int sh = 55 - e, sh_a = sh / 2, sh_b = sh - sh_a;
        const double sc_a = __longlong_as_double((long long)(1023 + sh_a) << 52);
        const double sc_b = __longlong_as_double((long long)(1023 + sh_b) << 52);
        const unsigned long long BIAS = 0x0080808080808080ull;
#pragma unroll
        for (int q = 0; q < 4; q++) {
            unsigned lo[2][4], hi[2][4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const int b = 4 * q + k;
                const long long X0 = __double2ll_rn((v[b] * sc_a) * sc_b);
                const unsigned long long Y0 = ((unsigned long long)X0 + BIAS) ^ BIAS;
                lo[0][k] = (unsigned)Y0;
                hi[0][k] = (unsigned)(Y0 >> 32);
                if (write_r) {
                    const long long X1 = ((negmask >> b) & 1u) ? -X0 : X0;
                    const unsigned long long Y1 = ((unsigned long long)X1 + BIAS) ^ BIAS;
                    lo[1][k] = (unsigned)Y1;
                    hi[1][k] = (unsigned)(Y1 >> 32);
                }
            }
#pragma unroll
            for (int p = 0; p < OZ_NS; p++) {
                const int i = OZ_NS - 1 - p;                          // byte index of slice p
                const unsigned sel = (unsigned)((i & 3) | (((i & 3) + 4) << 4));
#pragma unroll
                for (int side = 0; side < 2; side++) {
                    if (side == 1 && !write_r) continue;
                    const unsigned* src = (i < 4) ? lo[side] : hi[side];
                    const unsigned t01 = __byte_perm(src[0], src[1], sel);
                    const unsigned t23 = __byte_perm(src[2], src[3], sel);
                    const unsigned word = __byte_perm(t01, t23, 0x5410);
                    if (side == 0) wl[p][q] = word; else wr[p][q] = word;
                }
            }
        }
#pragma unroll
        for (int p = 0; p < OZ_NS; p++) {
            if (a.L) *reinterpret_cast<uint4*>(a.L + base + (size_t)p * OZ_CHUNK) = make_uint4(wl[p][0], wl[p][1], wl[p][2], wl[p][3]);
            if (write_r)
                *reinterpret_cast<uint4*>(a.R + base + (size_t)p * OZ_CHUNK) = make_uint4(wr[p][0], wr[p][1], wr[p][2], wr[p][3]);
        }
    }
}

// ------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t oz_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void oz_mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void oz_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void oz_mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool oz_mbar_try(uint32_t bar, uint32_t parity) {
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
// Bounded wait: a protocol error must never hang the GPU.  After ~2 s the CTA-wide sticky flag is raised, every later
// wait falls through, the kernel tears down normally (TMEM is released) and the host sees err |= 4.
__device__ __forceinline__ void oz_mbar_wait(uint32_t bar, uint32_t parity, volatile int* dead, int* err) {
    if (*dead) return;
    if (oz_mbar_try(bar, parity)) return;
    const long long t0 = clock64();
    while (!oz_mbar_try(bar, parity)) {
        if (*dead) return;
        if (clock64() - t0 > 4000000000LL) {
            *dead = 1;
            atomicOr(err, 4);
            return;
        }
    }
}
__device__ __forceinline__ void oz_bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ bool oz_elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void oz_tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void oz_tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void oz_tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, int8 x int8 -> int32, M = 128, N from the instruction descriptor, K = 32
__device__ __forceinline__ void oz_mma_i8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void oz_tmem_ld8(uint32_t taddr, int (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void oz_tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): K-major, no swizzle
//   [0,14) start >> 4 | [16,30) leading-dimension byte offset >> 4 (between the two 16-byte K chunks of one MMA)
//   [32,46) stride byte offset >> 4 (between 8-row groups) | [46,48) version = 1 | [61,64) layout type (0 = none)
__device__ __forceinline__ uint64_t oz_desc(uint32_t saddr, uint64_t hi_bits) {
    return hi_bits | (uint64_t)((saddr >> 4) & 0x3FFFu);
}

struct OzGemmArgs {
    const int8_t* L;
    const int8_t* R;
    const int* rexp;
    const int2* tiles;
    int* err;
    double* C;
    const double* Cin;
    const double* dadd;
    int ldc, ldcin, n, nkb;
    double beta, shift;
    uint64_t desc_hi;      // descriptor bits above the start address (LBO, SBO, version, layout)
    uint32_t idesc;
    int kb_end[3];         // k-blocks [kb_end[t-1], kb_end[t]) belong to term t
    int sgn[3];            // the B side of a signed term's k-blocks comes from R (sign applied), otherwise from L
    // general product C (+)= A B' with DISTINCT operands (LDL^T trailing update C -= W L'): the A side streams from L
    // (slices of W, row exponents rexp), the B side from R (slices of -Lpanel, row exponents rexp_col); `lower`: only
    // elements i >= j are authoritative, read from / written to the lower triangle in place, no mirror
    const int* rexp_col;   // null: same as rexp
    int lower;             // 0: upper triangle authoritative + mirror (SYRK); 1: lower triangle only, in place
    int ncols;             // > 0: only columns j < ncols exist (block-column update); 0: square (n)
    const int* ctrl;       // optional LDL^T control block (ctrl[4] != 0: abandoned, return at once)
};

// BN = tile width (UMMA N), ND = number of slice-pair diagonals kept (p + q < ND, p, q < OZ_NS):
//   ND = 8: 34 pairs, error ~ fp64 rounding of the product;  ND = 7: 28 pairs, ~1e-15 relative to the row scales;
//   ND = 6: 21 pairs, ~1e-12 (enough for a matrix that is only factored and refined against).
// At most MAXD = 512 / BN diagonals fit in the 512 TMEM columns, so they are done in NPASS passes from the top: pass pi
// covers diagonals [d_lo, d_hi), d_hi = ND - pi * MAXD, and needs the slices 0 .. min(d_hi, OZ_NS) - 1 of both operands.
template <int BN, int ND>
struct OzShape {
    static constexpr int MAXD = 512 / BN;
    static constexpr int NPASS = (ND + MAXD - 1) / MAXD;
    static constexpr int A_BYTES = OZ_NS * OZ_CHUNK;           // all slices of the A side for one k-block
    static constexpr int B_SLICE = BN * OZ_KB;
    static constexpr int B_BYTES = OZ_NS * B_SLICE;
    static constexpr int STAGE = A_BYTES + B_BYTES;
    static constexpr int STAGES = (int)(OZ_SMEM_BUDGET / STAGE) > 4 ? 4 : (int)(OZ_SMEM_BUDGET / STAGE);
    static constexpr int SMEM = STAGES * STAGE + 1024;
    __host__ __device__ static constexpr int d_hi(int pi) { return ND - pi * MAXD; }
    __host__ __device__ static constexpr int d_lo(int pi) { return d_hi(pi) - MAXD > 0 ? d_hi(pi) - MAXD : 0; }
};

template <int BN, int ND>
__global__ void __launch_bounds__(OZ_THREADS, 1) oz_syrk_kernel(const OzGemmArgs a) {
    using S = OzShape<BN, ND>;
    constexpr int MAXD = S::MAXD, NPASS = S::NPASS, A_BYTES = S::A_BYTES, B_SLICE = S::B_SLICE, STAGE = S::STAGE,
                  STAGES = S::STAGES;
    static_assert(STAGES >= 2, "pipeline needs two stages");
    extern __shared__ __align__(1024) uint8_t oz_smem[];
    __shared__ __align__(8) uint64_t bar_full[STAGES], bar_empty[STAGES], bar_tfull, bar_tempty;
    __shared__ uint32_t s_tmem;
    __shared__ int s_dead;
    __shared__ double s_cscale[BN];   // 2^(e_j - 7) of the tile's columns
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (a.ctrl && *reinterpret_cast<const volatile int*>(a.ctrl + 4)) return;      // before any barrier / TMEM allocation
    const int trc = (a.ctrl && tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) ? trace_begin(TR_OZ, a.ctrl) : -1;
    const int2 tile = a.tiles[blockIdx.x];
    const int ti = tile.x, tj = tile.y;
    const int row0 = ti * OZ_BM, col0 = tj * BN;
    const int* __restrict__ rexp_c = a.rexp_col ? a.rexp_col : a.rexp;
    const int ncol = a.ncols > 0 ? a.ncols : a.n;

    if (tid == 0) {
        s_dead = 0;
        for (int s = 0; s < STAGES; s++) {
            oz_mbar_init(oz_smem_u32(&bar_full[s]), 1);
            oz_mbar_init(oz_smem_u32(&bar_empty[s]), 1);
        }
        oz_mbar_init(oz_smem_u32(&bar_tfull), 1);
        oz_mbar_init(oz_smem_u32(&bar_tempty), 8);
        asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
    }
    for (int c = tid; c < BN; c += OZ_THREADS) s_cscale[c] = scalbn(1.0, ((col0 + c < ncol) ? rexp_c[col0 + c] : 0) - 7);
    if (warp == 0) {
        const uint32_t ncols = 512;
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(oz_smem_u32(&s_tmem)), "r"(ncols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n" ::: "memory");
    }
    oz_tc_fence_before();
    __syncthreads();
    oz_tc_fence_after();
    const uint32_t tmem = s_tmem;
    volatile int* dead = &s_dead;
    const uint32_t smem0 = oz_smem_u32(oz_smem);

    if (warp == 0) {
        // ===== producer: one thread streams the pre-tiled slices, stage by stage
        if (lane == 0) {
            int it = 0;
            for (int pi = 0; pi < NPASS; pi++) {
                const int nsl = S::d_hi(pi) < OZ_NS ? S::d_hi(pi) : OZ_NS;   // slices 0 .. nsl-1 are needed for diagonals < d_hi
                for (int kb = 0; kb < a.nkb; kb++, it++) {
                    const int s = it % STAGES;
                    const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                    oz_mbar_wait(oz_smem_u32(&bar_empty[s]), ph ^ 1u, dead, a.err);
                    const uint32_t full = oz_smem_u32(&bar_full[s]);
                    oz_mbar_expect_tx(full, (uint32_t)(nsl * (OZ_CHUNK + B_SLICE)));
                    const uint32_t sa = smem0 + s * STAGE, sb = sa + A_BYTES;
                    const int8_t* srcA = a.L + ((size_t)(ti * a.nkb + kb) * OZ_NS) * OZ_CHUNK;
                    oz_bulk_g2s(sa, srcA, (uint32_t)(nsl * OZ_CHUNK), full);
                    const int term = (kb < a.kb_end[0]) ? 0 : ((kb < a.kb_end[1]) ? 1 : 2);
                    const int8_t* Bsrc = a.sgn[term] ? a.R : a.L;
                    if (BN <= OZ_BM) {
                        const int rbB = col0 / OZ_BM, subB = (col0 % OZ_BM) * OZ_KB;   // byte offset of the BN rows in a chunk
                        const int8_t* srcB = Bsrc + ((size_t)(rbB * a.nkb + kb) * OZ_NS) * OZ_CHUNK + subB;
                        if (BN == OZ_BM) {
                            oz_bulk_g2s(sb, srcB, (uint32_t)(nsl * OZ_CHUNK), full);
                        } else {
                            for (int q = 0; q < nsl; q++)
                                oz_bulk_g2s(sb + q * B_SLICE, srcB + (size_t)q * OZ_CHUNK, (uint32_t)B_SLICE, full);
                        }
                    } else {
                        // BN = 256: a slice tile is two 128-row chunks from consecutive row blocks (rows beyond n: the
                        // slicing kernel zero-fills whole row blocks, and nrb is padded so the second block exists)
                        const int rbB = col0 / OZ_BM;
                        for (int q = 0; q < nsl; q++)
                            for (int hh = 0; hh < BN / OZ_BM; hh++)
                                oz_bulk_g2s(sb + q * B_SLICE + hh * OZ_CHUNK,
                                            Bsrc + ((size_t)((rbB + hh) * a.nkb + kb) * OZ_NS + q) * OZ_CHUNK, (uint32_t)OZ_CHUNK, full);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer.  The WHOLE warp runs the loop with warp-uniform values (descriptors stay in uniform
        // registers); only the tcgen05 instructions themselves are issued by one elected lane.  With a single-lane
        // branch around the loop the compiler has to re-uniformise every descriptor (ELECT / R2UR loops, ~30
        // instructions per MMA) and the issue rate, not the tensor pipe, bounds the kernel.
        int it = 0;
#pragma unroll
        for (int pi = 0; pi < NPASS; pi++) {
            const int d_lo = S::d_lo(pi), d_hi = S::d_hi(pi);
            if (pi > 0) {
                oz_mbar_wait(oz_smem_u32(&bar_tempty), (uint32_t)(pi - 1) & 1u, dead, a.err);
                oz_tc_fence_after();
            }
            for (int kb = 0; kb < a.nkb; kb++, it++) {
                const int s = it % STAGES;
                const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
                oz_mbar_wait(oz_smem_u32(&bar_full[s]), ph, dead, a.err);
                oz_tc_fence_after();
                const uint32_t sa = smem0 + s * STAGE, sb = sa + A_BYTES;
                const uint64_t da0 = oz_desc(sa, a.desc_hi), db0 = oz_desc(sb, a.desc_hi);
                const uint32_t acc_first = (kb > 0) ? 1u : 0u;
                if (oz_elect_one()) {
#pragma unroll
                    for (int d = d_lo; d < d_hi; d++) {
                        const int p0 = (d - (OZ_NS - 1) > 0) ? d - (OZ_NS - 1) : 0;      // p, q <= OZ_NS - 1
#pragma unroll
                        for (int p = p0; p <= d && p < OZ_NS; p++) {
                            const int q = d - p;
                            oz_mma_i8(tmem + (uint32_t)((d - d_lo) * BN), da0 + (uint64_t)(p * (OZ_CHUNK >> 4)),
                                      db0 + (uint64_t)(q * (B_SLICE >> 4)), a.idesc, (p > p0) ? 1u : acc_first);
                        }
                    }
                    oz_tc_commit(oz_smem_u32(&bar_empty[s]));      // frees the stage when these MMAs have read it
                }
                __syncwarp();
            }
            if (oz_elect_one()) oz_tc_commit(oz_smem_u32(&bar_tfull));   // accumulators of this pass are complete
            __syncwarp();
        }
    } else {
        // ===== epilogue: warps 2..9; warp (warp & 3) owns TMEM lanes 32*(warp & 3) .. +31 = tile rows, and the two warps of
        // a lane quarter take alternate 8-column chunks.  Per chunk: the global operands (partial sum of the higher
        // diagonals, Cin) are prefetched ahead -- ALL of them before the accumulators are complete when the kernel has a
        // single pass (the in-place LDL^T updates: their latency hides behind the MMA phase) --, the int32 accumulators of
        // the pass are read with tcgen05.ld, recombined by Horner in fp64 and scaled by exact powers of 2.
        const int q4 = warp & 3;
        const int half = (warp - 2) >> 2;
        const int rl = q4 * 32 + lane;
        const int i = row0 + rl;
        const bool rowok = i < a.n;
        const int ei = rowok ? a.rexp[i] : 0;
        const double dii = rowok ? (a.shift + (a.dadd ? a.dadd[i] : 0.0)) : 0.0;
        constexpr int NCH = BN / 16;              // chunks per epilogue warp
        constexpr bool PRE_ALL = (NPASS == 1) && (NCH <= 4);
#pragma unroll
        for (int pi = 0; pi < NPASS; pi++) {
            const int d_lo = S::d_lo(pi), d_hi = S::d_hi(pi);
            const int dn = d_hi - d_lo;
            const bool last = (pi == NPASS - 1);
            const bool need_part = (pi > 0), need_cin = last && (a.Cin != nullptr);
            const double si = scalbn(1.0, ei - 7 - 8 * d_lo);
            double pre_p[8], pre_c[8];
            double pre_all[PRE_ALL ? NCH * 8 : 1];
            auto prefetch = [&](int cb) {
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    const int j = col0 + cb * 8 + c;
                    const bool ok = rowok && j < ncol && (a.lower ? i >= j : i <= j);
                    pre_p[c] = (ok && need_part) ? a.C[(size_t)i * a.ldc + j] : 0.0;
                    pre_c[c] = (ok && need_cin) ? a.Cin[(size_t)i * a.ldcin + j] : 0.0;
                }
            };
            if (PRE_ALL) {
#pragma unroll
                for (int k = 0; k < (PRE_ALL ? NCH : 0); k++)
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const int j = col0 + (half + 2 * k) * 8 + c;
                        const bool ok = rowok && j < ncol && (a.lower ? i >= j : i <= j);
                        pre_all[k * 8 + c] = (ok && need_cin) ? a.Cin[(size_t)i * a.ldcin + j] : 0.0;
                    }
            } else {
                prefetch(half);
            }
            oz_mbar_wait(oz_smem_u32(&bar_tfull), (uint32_t)pi & 1u, dead, a.err);
            oz_tc_fence_after();
            auto chunk = [&](int cb, const double* cur_p, const double* cur_c, int (&acc)[MAXD][8]) {
                if (rowok) {
#pragma unroll
                    for (int c = 0; c < 8; c++) {
                        const int j = col0 + cb * 8 + c;
                        if (j >= ncol || (a.lower ? i < j : i > j)) continue;   // only one triangle is authoritative
                        double h = 0.0;
#pragma unroll
                        for (int dd = MAXD - 1; dd >= 0; dd--)
                            if (dd < dn) h = fma(h, 0.00390625, (double)acc[dd][c]);
                        double v = (h * si) * s_cscale[cb * 8 + c] + cur_p[c];
                        double* cp = a.C + (size_t)i * a.ldc + j;
                        if (last) {
                            v = fma(a.beta, cur_c[c], v);
                            if (i == j) v += dii;
                            *cp = v;
                            if (i != j && !a.lower) a.C[(size_t)j * a.ldc + i] = v;
                        } else {
                            *cp = v;
                        }
                    }
                }
            };
            if (PRE_ALL) {
                const double zero8[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int k = 0; k < (PRE_ALL ? NCH : 0); k++) {
                    const int cb = half + 2 * k;
                    int acc[MAXD][8];
#pragma unroll
                    for (int dd = 0; dd < MAXD; dd++)
                        if (dd < dn) oz_tmem_ld8(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(dd * BN + cb * 8), acc[dd]);
                    oz_tmem_wait_ld();
                    chunk(cb, zero8, pre_all + k * 8, acc);
                }
            } else {
                for (int cb = half; cb < BN / 8; cb += 2) {
                    int acc[MAXD][8];
#pragma unroll
                    for (int dd = 0; dd < MAXD; dd++)
                        if (dd < dn) oz_tmem_ld8(tmem + ((uint32_t)(q4 * 32) << 16) + (uint32_t)(dd * BN + cb * 8), acc[dd]);
                    double cur_p[8], cur_c[8];
#pragma unroll
                    for (int c = 0; c < 8; c++) { cur_p[c] = pre_p[c]; cur_c[c] = pre_c[c]; }
                    if (cb + 2 < BN / 8) prefetch(cb + 2);
                    oz_tmem_wait_ld();
                    chunk(cb, cur_p, cur_c, acc);
                }
            }
            oz_tc_fence_before();
            __syncwarp();
            if (lane == 0) oz_mbar_arrive(oz_smem_u32(&bar_tempty));
        }
    }
    oz_tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        oz_tc_fence_after();
        __syncwarp();
        const uint32_t ncols = 512;
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(tmem), "r"(ncols) : "memory");
    }
    trace_end(trc);
}

// ------------------------------------------------------------------------------------------- host side
struct OzWs {
    int8_t* L = nullptr;
    int8_t* R = nullptr;
    size_t cap = 0;           // bytes of each of L / R
    int* rexp = nullptr;
    int rexp_cap = 0;
    double* sw = nullptr;     // signed square roots of the column weights
    int sw_cap = 0;
    int* err = nullptr;       // device flag word
    int2* tiles = nullptr;
    int tiles_n = 0, tiles_bn = 0, ntiles = 0;
    // descriptor strides of the pre-tiled K-major / no-swizzle image: 128 B between the two 16-byte K chunks of an MMA
    // (leading-dimension byte offset), 256 B between 8-row groups (stride byte offset) -- pinned on hardware by
    // tools/oz_probe.py (the swapped assignment gives garbage)
    int lbo = 128, sbo = 256;
    int variant = 1;          // 0: 128x64 tiles, 1 pass (7 diagonals);  1: 128x128 tiles, 2 passes;  2: 128x256, 4 passes
    int ndiag = 8;            // 128x128 tiles: slice-pair diagonals kept (8: 34 pairs, 7: 28, 6: 21), see oz_syrk_kernel
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // optional: [0] start, [1] after slicing, [2] end
    bool reuse_slices = false; // profiling: skip the slicing kernels (the slices of the previous identical call are kept)
};
inline void oz_free(OzWs& w) {
    cudaFree(w.L); cudaFree(w.R); cudaFree(w.rexp); cudaFree(w.err); cudaFree(w.tiles); cudaFree(w.sw);
    for (int i = 0; i < 3; i++) if (w.ev[i]) cudaEventDestroy(w.ev[i]);
    w = OzWs();
}
inline int oz_variant_bn(int variant) { return variant == 0 ? 64 : (variant == 2 ? 256 : 128); }
inline uint64_t oz_desc_hi(int lbo, int sbo) {
    return ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) | ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}
inline uint32_t oz_idesc(int bn) {
    // cute::UMMA::InstrDescriptor: c_format [4,6) = 2 (S32), a_format [7,10) = 1 (signed 8-bit), b_format [10,13) = 1,
    // a/b major = K (0), n_dim [17,23) = N >> 3, m_dim [24,29) = M >> 4
    return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(OZ_BM >> 4) << 24);
}
template <int BN, int ND>
inline int oz_launch(cudaStream_t st, const OzGemmArgs& g, int ntiles) {
    constexpr int SMEM = OzShape<BN, ND>::SMEM;
    static bool attr_done = false;
    if (!attr_done) {
        CU(cudaFuncSetAttribute(oz_syrk_kernel<BN, ND>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
        attr_done = true;
    }
    oz_syrk_kernel<BN, ND><<<ntiles, OZ_THREADS, SMEM, st>>>(g);
    LAUNCHED();
    return 0;
}

// C = beta Cin + diag + sum_t alpha_t A_t diag(w_t) A_t'   (GEMM_UPPER_MIRROR semantics of gemm_nt).
// `may_be_negative`: bit t set if alpha_t * w_t can be negative (those k-blocks get a second, sign-carrying operand).
inline int oz_syrk(cudaStream_t st, const GemmArgs& a, OzWs& w, unsigned may_be_negative) {
    if (a.mode != GEMM_UPPER_MIRROR || a.n != a.m) return fail_msg("oz_syrk: symmetric (upper-mirror) products only");
    if (a.C == a.Cin && w.variant != 0) return fail_msg("oz_syrk: in-place accumulation needs the one-pass variant");
    if (a.nterms < 1 || a.nterms > 3) return fail_msg("oz_syrk: 1..3 terms");
    OzSliceArgs s{};
    OzGemmArgs g{};
    int koff = 0;
    bool anysigned = false;
    for (int t = 0; t < 3; t++) { g.kb_end[t] = 1 << 30; g.sgn[t] = 0; }
    for (int t = 0; t < a.nterms; t++) {
        if (a.t[t].A != a.t[t].B || a.t[t].lda != a.t[t].ldb) return fail_msg("oz_syrk: terms must be A diag(w) A'");
        const int sg = (may_be_negative >> t) & 1u;
        const int vec = (!(a.t[t].lda & 1) && !(reinterpret_cast<uintptr_t>(a.t[t].A) & 15)) ? 1 : 0;
        s.t[t] = OzTerm{a.t[t].A, a.t[t].w, a.t[t].lda, a.t[t].K, koff, sg, vec, a.t[t].alpha};
        koff += (int)rup((size_t)a.t[t].K, OZ_KB);
        g.kb_end[t] = koff / OZ_KB;
        g.sgn[t] = sg;
        anysigned = anysigned || sg;
    }
    if (koff > OZ_KMAX) return fail_msg("oz_syrk: contraction too long for exact int32 accumulation");
    const int nkb = std::max(koff, OZ_KB) / OZ_KB;
    const int bn = oz_variant_bn(w.variant);
    const int nrb = (int)rup((size_t)cdiv(a.n, OZ_BM), (size_t)std::max(1, bn / OZ_BM));   // whole tiles of row blocks
    const size_t need = (size_t)nrb * nkb * OZ_NS * OZ_CHUNK;
    if (need > w.cap) {
        cudaFree(w.L); cudaFree(w.R);
        w.L = w.R = nullptr;
        CU(cudaMalloc(&w.L, need));
        w.cap = need;
    }
    if (anysigned && !w.R) CU(cudaMalloc(&w.R, w.cap));
    if (nrb * OZ_BM > w.rexp_cap) {
        cudaFree(w.rexp);
        CU(cudaMalloc(&w.rexp, sizeof(int) * nrb * OZ_BM));
        w.rexp_cap = nrb * OZ_BM;
    }
    if (nkb * OZ_KB > w.sw_cap) {
        cudaFree(w.sw);
        CU(cudaMalloc(&w.sw, sizeof(double) * nkb * OZ_KB));
        w.sw_cap = nkb * OZ_KB;
    }
    if (!w.err) {
        CU(cudaMalloc(&w.err, sizeof(int)));
        CU(cudaMemsetAsync(w.err, 0, sizeof(int), st));
    }
    if (w.tiles_n != a.n || w.tiles_bn != bn) {
        std::vector<int2> tl;
        for (int tj = 0; tj * bn < a.n; tj++)
            for (int ti = 0; ti * OZ_BM < a.n && ti * OZ_BM <= tj * bn + bn - 1; ti++) tl.push_back(make_int2(ti, tj));
        cudaFree(w.tiles);
        CU(cudaMalloc(&w.tiles, sizeof(int2) * tl.size()));
        CU(cudaMemcpyAsync(w.tiles, tl.data(), sizeof(int2) * tl.size(), cudaMemcpyHostToDevice, st));
        CU(cudaStreamSynchronize(st));
        w.tiles_n = a.n; w.tiles_bn = bn; w.ntiles = (int)tl.size();
    }
    s.nterms = a.nterms; s.n = a.n; s.nkb = nkb;
    s.L = w.L; s.R = w.R; s.sw = w.sw; s.rexp = w.rexp; s.err = w.err;
    if (w.ev[0]) CU(cudaEventRecord(w.ev[0], st));
    if (!w.reuse_slices) {
        oz_weight_kernel<<<cdiv(nkb * OZ_KB, 256), 256, 0, st>>>(s, w.sw);
        LAUNCHED();
        oz_rowmax_kernel<<<cdiv(a.n, 4), 128, 0, st>>>(s);
        LAUNCHED();
        oz_slice_kernel<<<dim3(nrb * (OZ_BM / 8), std::max(1, std::min(8, nkb / 16))), 256, 0, st>>>(s);
        LAUNCHED();
    }
    if (w.ev[1]) CU(cudaEventRecord(w.ev[1], st));
    g.L = w.L; g.R = w.R; g.rexp = w.rexp; g.tiles = w.tiles; g.err = w.err;
    g.C = a.C; g.Cin = a.Cin; g.dadd = a.dadd; g.ldc = a.ldc; g.ldcin = a.ldcin; g.n = a.n; g.nkb = nkb;
    g.beta = a.beta; g.shift = a.shift;
    g.desc_hi = oz_desc_hi(w.lbo, w.sbo);
    g.idesc = oz_idesc(bn);
    int rc;
    if (w.variant == 0) rc = oz_launch<64, 7>(st, g, w.ntiles);             // test variants: 7 diagonals
    else if (w.variant == 2) rc = oz_launch<256, 7>(st, g, w.ntiles);
    else if (w.ndiag <= 6) rc = oz_launch<128, 6>(st, g, w.ntiles);
    else if (w.ndiag == 7) rc = oz_launch<128, 7>(st, g, w.ntiles);
    else rc = oz_launch<128, 8>(st, g, w.ntiles);
    if (rc == 0 && w.ev[2]) CU(cudaEventRecord(w.ev[2], st));
    return rc;
}


// ------------------------------------------------------------------------------------------- LDL^T trailing update
// C (n x n, lower triangle, in place) -= W * Lp'   with W, Lp n x K row-major (K <= 256): the panel-update contraction
// inside the factorisation on tcgen05 (BASELINE north_star), through the same error-free int8 split -- W's rows and
// Lp's rows are sliced separately (own row exponents), the B side streams the digits of -Lp, so 2x2 pivots (W != Lp D
// column-wise) need no special case.  128 x 64 tiles, six slice-pair diagonals (21 pairs, ~4e-12 of the row-scale
// products: the factor is a preconditioner + inertia test, every solve is refined against the unreduced system) in ONE
// pass, which is what allows the in-place accumulation.  Everything is preallocated (oz_upd_alloc) so that the calls can
// be captured into the factorisation's CUDA graph.
constexpr int OZ_UPD_NBUF = 4;     // panels in flight (same rotation as the W scratch panels of the factorisation)
struct OzUpdWs {
    int nmax = 0, kmax = 0;
    size_t stride = 0;                                     // bytes of the digits of one 128-row block (all k-blocks, all slices)
    int8_t *LA[OZ_UPD_NBUF] = {}, *RB[OZ_UPD_NBUF] = {};   // digits of the W rows / of the -L rows of a panel
    int *rexpA[OZ_UPD_NBUF] = {}, *rexpB[OZ_UPD_NBUF] = {};
    double* sw = nullptr;                                  // +1 / -1 column weights (two halves)
    std::vector<int2*> tiles;      // per (trailing order / 128, columns / 64)
    std::vector<int> ntiles;
    int ncolmax = 0;
};
inline void oz_upd_free(OzUpdWs& w) {
    for (int i = 0; i < OZ_UPD_NBUF; i++) { cudaFree(w.LA[i]); cudaFree(w.RB[i]); cudaFree(w.rexpA[i]); cudaFree(w.rexpB[i]); }
    cudaFree(w.sw);
    for (int2* t : w.tiles) cudaFree(t);
    w = OzUpdWs();
}
inline int oz_upd_alloc(OzUpdWs& w, int nmax, int kmax) {
    w.nmax = nmax; w.kmax = kmax;
    const int nrb = cdiv(nmax, OZ_BM), nkb = cdiv(kmax, OZ_KB);
    w.stride = (size_t)nkb * OZ_NS * OZ_CHUNK;
    for (int i = 0; i < OZ_UPD_NBUF; i++) {
        CU(cudaMalloc(&w.LA[i], (size_t)nrb * w.stride)); CU(cudaMalloc(&w.RB[i], (size_t)nrb * w.stride));
        CU(cudaMalloc(&w.rexpA[i], sizeof(int) * nrb * OZ_BM)); CU(cudaMalloc(&w.rexpB[i], sizeof(int) * nrb * OZ_BM));
    }
    CU(cudaMalloc(&w.sw, sizeof(double) * 2 * nkb * OZ_KB));
    {   // the column weights never change: +1 for the W side, -1 for the L side (C -= W L')
        std::vector<double> h((size_t)2 * nkb * OZ_KB);
        for (int i = 0; i < nkb * OZ_KB; i++) { h[i] = 1.0; h[(size_t)nkb * OZ_KB + i] = -1.0; }
        CU(cudaMemcpy(w.sw, h.data(), sizeof(double) * h.size(), cudaMemcpyHostToDevice));
    }
    w.ncolmax = nrb * 2 + 1;
    w.tiles.assign((size_t)(nrb + 1) * w.ncolmax, nullptr);
    w.ntiles.assign((size_t)(nrb + 1) * w.ncolmax, 0);
    return 0;
}
// tile list of a lower-trapezoidal piece: rows 0 .. n-1, columns 0 .. ncols-1 (ncols <= n), origin on the diagonal
// (cached; must be called OUTSIDE stream capture the first time)
inline int oz_upd_tiles(OzUpdWs& w, int n, int ncols) {
    const int kr = cdiv(n, OZ_BM), kc = cdiv(ncols, 64);
    if (kr >= (int)(w.tiles.size() / w.ncolmax) || kc >= w.ncolmax) return fail_msg("oz_upd_tiles: order exceeds the workspace");
    const size_t key = (size_t)kr * w.ncolmax + kc;
    if (w.tiles[key]) return 0;
    std::vector<int2> tl;
    for (int ti = 0; ti * OZ_BM < n; ti++)
        for (int tj = 0; tj * 64 < ncols && tj * 64 <= ti * OZ_BM + OZ_BM - 1; tj++) tl.push_back(make_int2(ti, tj));
    CU(cudaMalloc(&w.tiles[key], sizeof(int2) * tl.size()));
    CU(cudaMemcpy(w.tiles[key], tl.data(), sizeof(int2) * tl.size(), cudaMemcpyHostToDevice));
    w.ntiles[key] = (int)tl.size();
    return 0;
}
// digits of the rows of one outer panel: W (rows x K, leading dimension ldw) and -Lp (rows x K, ldl) -> buffer set `buf`
inline int oz_panel_slice(cudaStream_t st, int rows, const double* W, int ldw, const double* Lp, int ldl, int K, OzUpdWs& w,
                          int buf, int* err, const int* ctrl) {
    if (rows > w.nmax || K > w.kmax) return fail_msg("oz_panel_slice: workspace too small");
    const int nkb = cdiv(K, OZ_KB), nrb = cdiv(rows, OZ_BM);
    for (int side = 0; side < 2; side++) {
        OzSliceArgs s{};
        const double* A = side ? Lp : W;
        const int lda = side ? ldl : ldw;
        const int vec = (!(lda & 1) && !(reinterpret_cast<uintptr_t>(A) & 15)) ? 1 : 0;
        s.t[0] = OzTerm{A, nullptr, lda, K, 0, side, vec, side ? -1.0 : 1.0};
        s.nterms = 1; s.n = rows; s.nkb = nkb;
        s.L = side ? nullptr : w.LA[buf]; s.R = side ? w.RB[buf] : nullptr;
        s.sw = w.sw + (size_t)side * (w.kmax / OZ_KB) * OZ_KB; s.rexp = side ? w.rexpB[buf] : w.rexpA[buf];
        s.err = err; s.ctrl = ctrl;
        oz_rowmax_kernel<<<cdiv(rows, 4), 128, 0, st>>>(s);
        LAUNCHED();
        oz_slice_kernel<<<dim3(nrb * (OZ_BM / 8), 1), 256, 0, st>>>(s);
        LAUNCHED();
    }
    return 0;
}
// C (lower trapezoid n x ncols, origin on the diagonal, in place) -= W L'  from the digits in buffer set `buf`; the piece
// starts rb_off 128-row blocks below the first sliced row of the panel.  128 x 64 tiles, six slice-pair diagonals in ONE
// pass (21 pairs, ~4e-12 of the row-scale products) -- the single pass is what allows the in-place accumulation.
inline int oz_panel_update(cudaStream_t st, double* C, int ldc, int n, int ncols, int rb_off, int K, OzUpdWs& w, int buf,
                           int* err, const int* ctrl, int max_ctas) {
    const int kr = cdiv(n, OZ_BM), kc = cdiv(ncols, 64);
    const size_t key = (size_t)kr * w.ncolmax + kc;
    if (key >= w.tiles.size() || !w.tiles[key]) return fail_msg("oz_panel_update: tile list not prepared (oz_upd_tiles)");
    const int nkb = cdiv(K, OZ_KB);
    OzGemmArgs g{};
    g.L = w.LA[buf] + (size_t)rb_off * w.stride; g.R = w.RB[buf] + (size_t)rb_off * w.stride;
    g.rexp = w.rexpA[buf] + rb_off * OZ_BM; g.rexp_col = w.rexpB[buf] + rb_off * OZ_BM;
    g.err = err;
    g.C = C; g.Cin = C; g.dadd = nullptr; g.ldc = ldc; g.ldcin = ldc; g.n = n; g.ncols = ncols; g.nkb = nkb;
    g.beta = 1.0; g.shift = 0.0; g.lower = 1; g.ctrl = ctrl;
    g.desc_hi = oz_desc_hi(128, 256);
    g.idesc = oz_idesc(64);
    for (int t = 0; t < 3; t++) { g.kb_end[t] = 1 << 30; g.sgn[t] = 0; }
    g.kb_end[0] = nkb; g.sgn[0] = 1;
    // one CTA = one tile, ~200 KB of shared memory: a full-width launch would occupy every SM and the latency-critical
    // chain kernels of the factorisation would wait for a slot, so the tile list goes out in waves of at most max_ctas
    const int wave = max_ctas > 0 ? max_ctas : w.ntiles[key];
    for (int off = 0; off < w.ntiles[key]; off += wave) {
        g.tiles = w.tiles[key] + off;
        RET((oz_launch<64, 6>(st, g, std::min(wave, w.ntiles[key] - off))));
    }
    return 0;
}

// int8 multiply-add operations one call issues on the tensor cores (slice pairs x computed tiles x K)
inline int oz_npairs(int nd) {
    int np = 0;
    for (int d = 0; d < nd; d++)
        for (int p = 0; p <= d; p++) np += (p < OZ_NS && d - p < OZ_NS) ? 1 : 0;
    return np;
}
inline double oz_syrk_int8_ops(const GemmArgs& a, int bn, int nd) {
    double ksum = 0;
    for (int t = 0; t < a.nterms; t++) ksum += (double)rup((size_t)a.t[t].K, OZ_KB);
    double tiles = 0;
    for (int tj = 0; tj * bn < a.n; tj++)
        for (int ti = 0; ti * OZ_BM < a.n && ti * OZ_BM <= tj * bn + bn - 1; ti++) tiles += 1;
    return 2.0 * tiles * OZ_BM * bn * ksum * (double)oz_npairs(nd);
}

}  // namespace b200
