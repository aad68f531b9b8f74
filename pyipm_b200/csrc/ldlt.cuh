// ldlt.cuh -- dense symmetric-INDEFINITE factorisation  A = Lhat * D * Lhat^T  with inertia, on the device.
//
// Replaces the reference's  eigvalsh(K, I)  inertia test (pyipm.py:906-909, 1378-1403; >= 90 % of its Newton
// step) and its general-LU solve (pyipm.py:18-20, 911-914, 1720).
//
// Algorithm: right-looking blocked LDL^T, block size NB = 64, Bunch-Kaufman (1x1 / 2x2) pivoting restricted to
// the diagonal tile ("tile pivoting": no pivot search touches HBM outside the 64 x 64 tile, so the only
// serial part of a panel step is one CTA working out of shared memory).  Per panel k (rows/cols k0..k1):
//   1. ldlt_tile_kernel   P T P^T = L D L^T  of the diagonal tile in smem, fused Gauss-Jordan accumulation of
//                         X = L^-1; emits LinvP = X * P (so nobody downstream needs the permutation),
//                         D^-1 blocks, inertia counts and |D| extremes.
//   2. gemm_nt (DMMA)     W = B * LinvP^T            (W = Lpanel * D, rows below the tile)
//   3. ldlt_scale_kernel  Lpanel = W * D^-1          (overwrites B in place)
//   4. gemm_nt (DMMA)     A22 -= W * Lpanel^T        (lower tiles only)  -- Kc^3/3 FLOPs total, the tensor work
// Sylvester's law of inertia makes #negative eigenvalues = #negative 1x1 pivots + #2x2 blocks with det < 0
// (+ 2 for negative-definite 2x2 blocks), which is what pyipm.py:1381,1399 tests.
//
// Solves are ONE launch per direction.  Default ("block-256 solves", end of this file): a link of the serial chain is a
// 256-row block handled by one thread-block cluster of eight CTAs, which polls the values published by earlier links,
// exchanges partial results through distributed shared memory and applies the explicit inverse of the block's 256 x 256
// diagonal part (built by ldlt_blockinv_kernel at the end of the factorisation).  B200IPM_SOLVE256=0 keeps the original
// chain: CTA i owns 64-row block-row i, spins on the values published by the CTAs of earlier block rows (tickets
// guarantee forward progress), and applies LinvP_i as a 64 x 64 GEMV.
#pragma once
#include <algorithm>
#include <cooperative_groups.h>
#include "common.cuh"
#include "gemm_nt.cuh"
#include "ozaki_i8.cuh"
#include "vec.cuh"

namespace b200 {
namespace cg = cooperative_groups;

constexpr int NB = 64;
constexpr int NBP = NB + 1;   // padded smem row
constexpr int TILE_SMEM = 2 * NB * NBP * 8;

struct LdltWs {
    int n = 0, ld = 0, nblk = 0;
    double* A = nullptr;       // n x ld, factored in place (strictly-lower part holds Lhat panels)
    double* Wp = nullptr;      // n x NBO  outer-panel scratch (L * D)
    double* LinvP = nullptr;   // nblk x NB x NB
    double* dinfo = nullptr;   // 4 x n: [dinv_a | dinv_b | d_a | d_b]
    int* kind = nullptr;       // n: 0 = 1x1, 1 = first of 2x2, 2 = second of 2x2
    int* counts = nullptr;     // [neg, zero, pos, -, abandoned flag, negative-pivot limit, -, -]  (control block)
    double* dstat = nullptr;   // [min |eig(D)|, max |eig(D)|]
    unsigned* flags = nullptr; // 2 x nblk epoch flags (forward, backward)
    unsigned* ticket = nullptr;
    double *yv = nullptr, *zv = nullptr, *xv = nullptr;
    unsigned epoch = 0;
    cudaStream_t st = nullptr;
    double* Wp2 = nullptr;     // second / third outer-panel scratch (look-ahead depth 2: triple buffering)
    double* Wp3 = nullptr;
    double* Wp4 = nullptr;
    cudaStream_t urg = nullptr;    // urgent pieces of the trailing update (block columns of the next two panels)
    cudaEvent_t ev_urg[2] = {nullptr, nullptr}, ev_slice[2] = {nullptr, nullptr};
    cudaStream_t side = nullptr;   // trailing updates beyond the next panel run here, overlapped with the next panel
    cudaEvent_t ev_panel[2] = {nullptr, nullptr}, ev_upd[2] = {nullptr, nullptr};
    cudaStream_t upd = nullptr;    // in-panel updates off the chain (everything of a tile step but the next diagonal tile)
    cudaEvent_t ev_tile = nullptr, ev_mini = nullptr, ev_urest = nullptr, ev_a2 = nullptr, ev_corner = nullptr;
    int split_a = 1;               // B200IPM_LDLT_SPLITA=0: the whole boundary update on the chain stream
    int* sig = nullptr;            // device word set to 1 when tile step sig_tile starts (baked into the graph)
    int sig_tile = -1;
    int tile_blocked = 1;          // blocked fast attempt inside the tile kernel (B200IPM_TILE_BLOCKED=0: per-pivot barrier version)
    int use_mini = 1;              // B200IPM_LDLT_MINI=0 restores the tile -> panel -> update chain
    cudaStream_t cap = nullptr;    // internal capture-origin stream (the caller's stream may be the legacy default one)
    cudaGraphExec_t gexec = nullptr;
    double graph_u = -1.0;         // pivot_u the captured graph was built with
    int graph_sig_tile = -1;       // sig_tile the captured graph was built with
    int graph_nodes = 0;
    int graph_state = 0;           // 0 = not tried, 1 = usable, -1 = capture failed: direct launches
    int neg_limit = 0x7fffffff; // value last written to counts[5]
    int side_ctas = 0;         // > 0: the bulk trailing updates run as persistent kernels of at most this many CTAs
    int* serr = nullptr;       // device error word of the solve kernels (bit 8: a poll timed out)
    int tc_update = 0;         // bulk trailing updates on tcgen05 (error-free int8 split), B200IPM_LDLT_TC=0 restores DMMA
    int tc_ctas = 96;          // CTAs per wave of the tcgen05 update (B200IPM_LDLT_TC_CTAS)
    OzUpdWs tcu;
    TmaMat tmA, tmW[4];        // tensor maps of the KKT matrix and the four W scratch panels (TMA-staged updates)
    int use_tma = 1;           // B200IPM_LDLT_TMA=0: cp.async staging (gemm_nt_sub64_kernel)
    double *Xinv = nullptr, *XinvT = nullptr;   // explicit inverses of the 256-row diagonal blocks (block-256 solves) + transposes
    int nb256 = 0;
    int solve256 = 1;          // B200IPM_SOLVE256=0: the 64-row chain (ldlt_fwd_kernel / ldlt_bwd_kernel)
    int binv_mode = 1;         // where the block inverses are built: 1 one launch at the end (default), 0 per outer panel on the bulk stream
    cudaEvent_t col_rest_event = nullptr;   // column-block mode: rows below the diagonal block are ready once this event fires
    double pivot_u = 0.01;     // threshold of the fast (unpivoted) tile attempt: accept step j iff |d_j| >= u * max_i |T[i][j]|
};

inline int ldlt_alloc(LdltWs& w, int n, cudaStream_t st, bool background = false) {
    w.n = n;
    w.ld = (int)rup(n, 16);
    w.nblk = cdiv(n, NB);
    w.st = st;
    w.epoch = 0;
    const size_t npad = (size_t)w.nblk * NB;
    CU(cudaMalloc(&w.A, sizeof(double) * (size_t)npad * w.ld));
    CU(cudaMalloc(&w.Wp, sizeof(double) * npad * 256));
    CU(cudaMalloc(&w.Wp2, sizeof(double) * npad * 256));
    CU(cudaMalloc(&w.Wp3, sizeof(double) * npad * 256));
    CU(cudaMalloc(&w.Wp4, sizeof(double) * npad * 256));
    // the serial chain (tile -> panel -> in-panel update) is captured on a HIGH-priority stream, the bulk trailing
    // updates on a LOW-priority one: a chain kernel never queues behind a full wave of long update CTAs (captured
    // kernel nodes inherit the priority of the stream they were captured on)
    int prio_lo = 0, prio_hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    // (numerically smaller = more urgent).  A BACKGROUND workspace (the speculative delta = 0 inertia test whose
    // verdict is only needed later) sits entirely below a foreground one.
    const int span = prio_lo - prio_hi;
    const int p_chain = background ? prio_hi + std::min(span, 2) : prio_hi;
    const int p_side = background ? prio_lo : prio_hi + std::min(span, 1);
    {
        // SMs left to the serial chain while a bulk update runs (B200IPM_SIDE_FG / B200IPM_SIDE_BG override, 0 = all)
        const char* e = getenv(background ? "B200IPM_SIDE_BG" : "B200IPM_SIDE_FG");
        w.side_ctas = e ? atoi(e) : (background ? 48 : 96);   // measured at config 3 (tools/sweep_side.py)
    }
    CU(cudaStreamCreateWithPriority(&w.side, cudaStreamNonBlocking, p_side));
    CU(cudaStreamCreateWithPriority(&w.cap, cudaStreamNonBlocking, p_chain));
    CU(cudaStreamCreateWithPriority(&w.upd, cudaStreamNonBlocking, p_chain));
    CU(cudaStreamCreateWithPriority(&w.urg, cudaStreamNonBlocking, std::min(p_chain + 1, p_side)));
    CU(cudaEventCreateWithFlags(&w.ev_tile, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&w.ev_mini, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&w.ev_urest, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&w.ev_a2, cudaEventDisableTiming));
    CU(cudaEventCreateWithFlags(&w.ev_corner, cudaEventDisableTiming));
    { const char* e = getenv("B200IPM_LDLT_SPLITA"); if (e) w.split_a = atoi(e); }
    { const char* e = getenv("B200IPM_LDLT_MINI"); if (e) w.use_mini = atoi(e); }
    { const char* e = getenv("B200IPM_TILE_BLOCKED"); if (e) w.tile_blocked = atoi(e); }
    for (int i = 0; i < 2; i++) {
        CU(cudaEventCreateWithFlags(&w.ev_panel[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&w.ev_upd[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&w.ev_urg[i], cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&w.ev_slice[i], cudaEventDisableTiming));
    }
    CU(cudaMalloc(&w.LinvP, sizeof(double) * (size_t)w.nblk * NB * NB));
    CU(cudaMalloc(&w.dinfo, sizeof(double) * 4 * npad));
    CU(cudaMalloc(&w.kind, sizeof(int) * npad));
    CU(cudaMalloc(&w.counts, sizeof(int) * 8));
    {
        const int init[8] = {0, 0, 0, 0, 0, 0x7fffffff, 0, 0};
        CU(cudaMemcpy(w.counts, init, sizeof(init), cudaMemcpyHostToDevice));
    }
    CU(cudaMalloc(&w.dstat, sizeof(double) * 2));
    CU(cudaMalloc(&w.serr, sizeof(int)));
    CU(cudaMemset(w.serr, 0, sizeof(int)));
    {
        const char* e = getenv("B200IPM_LDLT_TMA");
        if (e) w.use_tma = atoi(e);
        if (w.use_tma) {
            RET(tma_describe(w.tmA, w.A, (int)npad, w.ld));
            double* wb[4] = {w.Wp, w.Wp2, w.Wp3, w.Wp4};
            for (int i = 0; i < 4; i++) RET(tma_describe(w.tmW[i], wb[i], (int)npad, 256));
        }
    }
    {
        // tcgen05 trailing updates pay off once the bulk piece (everything beyond the next three outer panels) exists
        const char* e = getenv("B200IPM_LDLT_TC");
        w.tc_update = (e ? atoi(e) : 1) && n >= 2048;
        const char* c = getenv("B200IPM_LDLT_TC_CTAS");
        // config-3 size: waves of 96 CTAs leave SMs to the chain kernels (the factorisation is chain-bound); from order 8192
        // on the updates are the bound and every tile goes out in one launch (measured at order 16384: 63 -> 47 ms)
        w.tc_ctas = (n >= 8192) ? 0 : 96;
        if (c) w.tc_ctas = atoi(c);
        if (w.tc_update) {
            RET(oz_upd_alloc(w.tcu, n, 256));
            // tile lists of every piece the factorisation will launch (they must exist before the graph is captured)
            for (int c0 = 0; c0 < n; c0 += 256) {
                const int c1 = std::min(c0 + 256, n), rows2 = n - c1;
                const int na = std::min(256, rows2), rows3 = rows2 - na;
                if (rows3 <= 0) continue;
                const int na2 = std::min(256, rows3), rows4 = rows3 - na2;
                RET(oz_upd_tiles(w.tcu, rows3, na2));
                if (rows4 <= 0) continue;
                const int na3 = std::min(256, rows4), rows5 = rows4 - na3;
                RET(oz_upd_tiles(w.tcu, rows4, na3));
                if (rows5 > 0) RET(oz_upd_tiles(w.tcu, rows5, rows5));
            }
        }
    }
    CU(cudaMalloc(&w.flags, sizeof(unsigned) * 2 * w.nblk));
    CU(cudaMalloc(&w.ticket, sizeof(unsigned) * 2));
    w.nb256 = cdiv(w.nblk, 4);
    { const char* e = getenv("B200IPM_SOLVE256"); if (e) w.solve256 = atoi(e); }
    { const char* e = getenv("B200IPM_BINV_MODE"); if (e) w.binv_mode = atoi(e); }
    const size_t npad256 = (size_t)w.nb256 * 256;
    CU(cudaMalloc(&w.yv, sizeof(double) * npad256));
    CU(cudaMalloc(&w.zv, sizeof(double) * npad256));
    CU(cudaMalloc(&w.xv, sizeof(double) * npad256));
    CU(cudaMalloc(&w.Xinv, sizeof(double) * npad256 * 256));
    CU(cudaMalloc(&w.XinvT, sizeof(double) * npad256 * 256));
    CU(cudaMemsetAsync(w.Xinv, 0, sizeof(double) * npad256 * 256, st));
    CU(cudaMemsetAsync(w.XinvT, 0, sizeof(double) * npad256 * 256, st));
    CU(cudaMemsetAsync(w.flags, 0, sizeof(unsigned) * 2 * w.nblk, st));
    CU(cudaMemsetAsync(w.A, 0, sizeof(double) * (size_t)npad * w.ld, st));
    return 0;
}
inline void ldlt_free(LdltWs& w) {
    cudaFree(w.A); cudaFree(w.Wp); cudaFree(w.Wp2); cudaFree(w.Wp3); cudaFree(w.Wp4);
    oz_upd_free(w.tcu);
    if (w.urg) cudaStreamDestroy(w.urg);
    for (int i = 0; i < 2; i++) { if (w.ev_urg[i]) cudaEventDestroy(w.ev_urg[i]); if (w.ev_slice[i]) cudaEventDestroy(w.ev_slice[i]); }
    if (w.gexec) cudaGraphExecDestroy(w.gexec);
    if (w.side) cudaStreamDestroy(w.side);
    if (w.upd) cudaStreamDestroy(w.upd);
    if (w.ev_tile) cudaEventDestroy(w.ev_tile);
    if (w.ev_mini) cudaEventDestroy(w.ev_mini);
    if (w.ev_urest) cudaEventDestroy(w.ev_urest);
    if (w.ev_a2) cudaEventDestroy(w.ev_a2);
    if (w.ev_corner) cudaEventDestroy(w.ev_corner);
    if (w.cap) cudaStreamDestroy(w.cap);
    for (int i = 0; i < 2; i++) { if (w.ev_panel[i]) cudaEventDestroy(w.ev_panel[i]); if (w.ev_upd[i]) cudaEventDestroy(w.ev_upd[i]); }
    cudaFree(w.LinvP); cudaFree(w.dinfo); cudaFree(w.kind); cudaFree(w.counts);
    cudaFree(w.dstat); cudaFree(w.serr); cudaFree(w.flags); cudaFree(w.ticket); cudaFree(w.yv); cudaFree(w.zv); cudaFree(w.xv);
    cudaFree(w.Xinv); cudaFree(w.XinvT);
    w = LdltWs();
}

// ------------------------------------------------------------------------------------------- tile kernel
// One CTA, 256 threads: P T P^T = L D L^T of one 64 x 64 diagonal tile with Bunch-Kaufman pivoting, fused with the
// Gauss-Jordan accumulation of X = L^-1.  The 64 pivot steps are a serial chain (measured: the chain, not the
// arithmetic, is the cost), so the kernel is organised for latency:
//   * warp w owns rows w, w+8, ..., w+56, lanes own columns lane and lane+32: every thread keeps its 8 x 2 slab of
//     T and of X in REGISTERS for the whole factorisation (statically indexed, fully unrolled);
//   * for column m a step touches EITHER T (m beyond the pivot) OR X (m up to the pivot), never both;
//   * the only shared-memory traffic of a step is the pivot row (T and X parts), published one step ahead into a
//     double-buffered row by the warp that owns it;
//   * software pipeline: in the unrolled row loop the next pivot row is the first ACTIVE row of its owner, so that
//     warp updates it, runs the first-level Bunch-Kaufman test on the values it holds in registers (ONE vote.any in
//     the common case; redux.sync arg-max only when the test fails) and publishes the decision together with d and
//     1/d (rcp.approx + two Newton steps) while everybody else is still updating => one barrier per step;
//   * anything else (second-level test, interchange, 2x2 pivot -- rare on barrier-regularised KKT matrices) takes a
//     slow step: registers are spilled to the smem copy of T/X, the step runs there, registers are reloaded.
constexpr int TILE_THREADS = 256;
constexpr int TILE_RPW = 8;   // rows per warp
struct TileDec { int code; int r; double d, dinv, absakk, colmax; };   // code = need2 | kstep << 1 | kp << 8

__device__ __forceinline__ double warp_absmax_bits(double v) {   // max of non-negative doubles via redux.sync
    const unsigned long long b = (unsigned long long)__double_as_longlong(v);
    const unsigned h = (unsigned)(b >> 32), l = (unsigned)b;
    const unsigned mh = __reduce_max_sync(0xffffffffu, h);
    const unsigned ml = __reduce_max_sync(0xffffffffu, (h == mh) ? l : 0u);
    return __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
}
__device__ __forceinline__ double fast_rcp(double x) {
    const double ax = fabs(x);
    if (!(ax > 1e-290 && ax < 1e290)) return 1.0 / x;
    double r;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
    double e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    e = fma(-x, r, 1.0);
    r = fma(r, e, r);
    return r;
}
// First-level Bunch-Kaufman test for pivot row jn, executed by ONE warp.  v0/v1 are T[jn][lane], T[jn][lane+32].
__device__ __forceinline__ void tile_search(int jn, int nb, int lane, double v0, double v1, TileDec* out) {
    const double BK_ALPHA = 0.6403882032022076;   // (1 + sqrt(17)) / 8
    const double djj = __shfl_sync(0xffffffffu, (jn < 32) ? v0 : v1, jn & 31);
    const double absakk = fabs(djj);
    const int c0 = lane, c1 = lane + 32;
    const bool in0 = (c0 > jn && c0 < nb), in1 = (c1 > jn && c1 < nb);
    // need2  <=>  absakk < alpha * colmax  <=>  some |v_m| * alpha exceeds absakk   (both-zero case: false)
    const bool viol = (in0 && BK_ALPHA * fabs(v0) > absakk) || (in1 && BK_ALPHA * fabs(v1) > absakk);
    const double rinv = (djj != 0.0) ? fast_rcp(djj) : 0.0;   // independent of the vote: overlaps its latency
    if (!__any_sync(0xffffffffu, viol)) {
        if (lane == 0) {
            out->code = (1 << 1) | (jn << 8);
            out->d = djj;
            out->dinv = rinv;
        }
        return;
    }
    double cm = -1.0;
    int r = jn;
    if (in0) { cm = fabs(v0); r = c0; }
    if (in1) { const double v = fabs(v1); if (v > cm) { cm = v; r = c1; } }
    const bool have = (cm >= 0.0);
    const unsigned long long bits = have ? (unsigned long long)__double_as_longlong(cm) : 0ull;
    const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
    const unsigned mh = __reduce_max_sync(0xffffffffu, hi);
    const unsigned ml = __reduce_max_sync(0xffffffffu, (hi == mh) ? lo : 0u);
    const unsigned cand = (have && hi == mh && lo == ml) ? (unsigned)r : 0xffffu;
    const unsigned rsel = __reduce_min_sync(0xffffffffu, cand);
    if (lane == 0) {
        out->code = 1 | (1 << 1) | (jn << 8);
        out->r = (rsel == 0xffffu) ? jn : (int)rsel;
        out->d = djj;
        out->dinv = 0.0;
        out->absakk = absakk;
        out->colmax = __longlong_as_double((long long)(((unsigned long long)mh << 32) | ml));
    }
}

// Eight consecutive unpivoted elimination steps j = 8*JB .. 8*JB+7 of the fast attempt (see ldlt_tile_kernel).
// Ownership here is TRANSPOSED with respect to the pivoting loop: lane l owns rows l and l+32, warp w owns columns
// w, w+8, ..., w+56 (register b[k][h] = element (l + 32h, w + 8k)).  A step then is, per warp: 11 LDS (pivot-row
// entries), 2 row masks, 16 FMAs; the column that switches from T-type to X-type is a warp-uniform reset of
// b[JB][*] (static index), its L entries are two stores per lane, and the next pivot row is published by the one
// lane per warp that holds it.  The kernel is bound by the number of warp-instructions per step (each SMSP pipe
// needs ~2 cycles per instruction), hence the effort to keep the step at ~65 instructions per warp.
template <int JB>
__device__ __forceinline__ void tile_fast_block(double (&b)[TILE_RPW][2], int nb, int lane, int warp, int tid,
                                                double* __restrict__ Tf, double* __restrict__ prow0,
                                                double* __restrict__ pdiag, double* __restrict__ sda,
                                                int& viol, bool& allpos, bool& allneg, const double pivot_u) {
    if (JB * 8 >= nb) return;
    const int r0 = lane, r1 = lane + 32;
    for (int w8 = 0; w8 < 8; w8++) {
        const int j = JB * 8 + w8;
        if (j >= nb) break;
        const int par = j & 1;
        const double* pr = prow0 + par * NB;
        const double d = pdiag[par * 2], dinv = pdiag[par * 2 + 1];
        const double u0 = pr[r0], u1 = pr[r1];
        double q[TILE_RPW];
#pragma unroll
        for (int k = 0; k < TILE_RPW; k++) q[k] = pr[warp + 8 * k];
        const bool act0 = (r0 > j) && (r0 < nb), act1 = (r1 > j) && (r1 < nb);
        const double g0 = act0 ? u0 * dinv : 0.0;      // L[r0][j]; zero for finished / padded rows => unchanged
        const double g1 = act1 ? u1 * dinv : 0.0;
        if (warp == 0) {   // sticky first-level Bunch-Kaufman check on the pivot column (= pivot row by symmetry)
            const double thr = fabs(d);
            viol |= ((act0 && pivot_u * fabs(u0) > thr) || (act1 && pivot_u * fabs(u1) > thr)) ? 1 : 0;
        }
        allpos = allpos && (d > 0.0);
        allneg = allneg && (d < 0.0);
        if (warp == w8) {  // column j switches to X-type: L entries out, registers restart from e_j
            if (act0) Tf[r0 * NBP + j] = g0;
            if (act1) Tf[r1 * NBP + j] = g1;
            b[JB][0] = (r0 == j) ? 1.0 : 0.0;
            b[JB][1] = (r1 == j) ? 1.0 : 0.0;
            if (lane == 0) sda[j] = d;
        }
#pragma unroll
        for (int k = 0; k < TILE_RPW; k++) {
            b[k][0] = fma(-g0, q[k], b[k][0]);
            b[k][1] = fma(-g1, q[k], b[k][1]);
        }
        // ---- publish the next pivot row: in every warp, the lane that holds row jn writes its 8 columns
        const int jn = j + 1;
        if (jn < nb && lane == (jn & 31)) {
            double* pw = prow0 + (par ^ 1) * NB;
            const bool hi = (JB > 3) || (JB == 3 && w8 == 7);     // row jn lives in the upper half (rows >= 32)
#pragma unroll
            for (int k = 0; k < TILE_RPW; k++) pw[warp + 8 * k] = hi ? b[k][1] : b[k][0];
            if (warp == (jn & 7)) {   // this lane also holds the diagonal T[jn][jn]
                constexpr int KN = (JB + 1 < TILE_RPW) ? JB + 1 : JB;
                const double dn = (w8 < 7) ? (hi ? b[JB][1] : b[JB][0]) : (hi ? b[KN][1] : b[KN][0]);
                pw[jn] = 1.0;
                pdiag[(par ^ 1) * 2] = dn;
                pdiag[(par ^ 1) * 2 + 1] = (dn != 0.0) ? fast_rcp(dn) : 0.0;
            }
        }
        __syncthreads();
    }
}

// Blocked variant of the fast attempt (same acceptance rule, same outputs: L^-1 in Xf, pivots in sda).
// The 64 x 64 tile is eliminated in eight 8-column blocks:
//   (1) the 8 x 8 diagonal block is factored by ONE thread entirely in registers (the serial chain of a pivot step is
//       rcp -> mul -> fma, no shuffle and no barrier); the rows below it then follow independently, one thread per row,
//       by forward substitution against U = D L8'.  The threshold test |d_j| >= u |T[i][j]| is applied in the
//       equivalent form |l_ij| <= 1/u.  (A first version ran the whole 64 x 8 panel on one warp with shuffles: 483
//       cycles per pivot, no faster than the per-pivot barrier kernel.)
//   (2) the trailing part of the tile gets its rank-8 update  T -= (L D) L'  from all eight warps by DMMA
//       (lower 8 x 8 sub-tiles only: the elimination never reads the upper triangle);
//   (3) after the last block  X = L^-1  is built by block forward substitution, one block column per warp:
//       X_jj = inv(L_jj),  X_ij = -X_ii * sum_{k=j}^{i-1} L_ik X_kj.
// Two CTA barriers per block instead of one per pivot step.
__device__ __forceinline__ void tile_fast_blocked(double* __restrict__ Tf, double* __restrict__ Xf, double* __restrict__ sda,
                                                  double* __restrict__ xscr, double* __restrict__ pinv, int nb, int lane, int warp,
                                                  double pivot_u, int& viol, bool& allpos, bool& allneg) {
    const int g = lane >> 2, tg = lane & 3;
#ifdef TILE_PROF
    long long cy_x = 0, cy_t = clock64();
#endif
    // (1a) an 8 x 8 diagonal block, by ONE thread entirely in registers: no shuffle, no barrier on the serial chain
    // (~2200 cycles per block, TILE_PROF; a warp-cooperative version -- 36 elements over the lanes, three shuffle rounds per
    // pivot -- was measured at 31 us per tile against 19: shuffle latency on the dependent chain costs more than it saves)
    auto diag_block = [&](const int c0) {
        double a[8][8];
#pragma unroll
        for (int r = 0; r < 8; r++)
#pragma unroll
            for (int c = 0; c <= r; c++) a[r][c] = Tf[(c0 + r) * NBP + c0 + c];
        // (the threshold test of the block's own multipliers and the sign bookkeeping of its pivots are NOT on this
        // thread's instruction stream: 28 otherwise idle threads test the stored multipliers in phase (1b), the signs are
        // read off sda after the last block)
#pragma unroll
        for (int j = 0; j < 8; j++) {
            const double d = a[j][j];
            const double dinv = (d != 0.0) ? fast_rcp(d) : 0.0;
            double l[8];
#pragma unroll
            for (int r = j + 1; r < 8; r++) l[r] = a[r][j] * dinv;
#pragma unroll
            for (int r = j + 1; r < 8; r++)
#pragma unroll
                for (int c = j + 1; c <= r; c++) a[r][c] = fma(-l[r], a[c][j], a[r][c]);
#pragma unroll
            for (int r = j + 1; r < 8; r++) a[r][j] = l[r];
            sda[c0 + j] = d;
            pinv[j] = dinv;
        }
#pragma unroll
        for (int r = 1; r < 8; r++)
#pragma unroll
            for (int c = 0; c < r; c++) Tf[(c0 + r) * NBP + c0 + c] = a[r][c];
    };
    auto update_tile = [&](const int kb, const int t) {      // t-th lower 8 x 8 tile of the part behind block kb
        const int c0 = 8 * kb;
        int bi = 0, rem = t;
        while (rem > bi) { rem -= bi + 1; bi++; }
        const int BI = kb + 1 + bi, BM = kb + 1 + rem;
        double* cp = Tf + (8 * BI + g) * NBP + 8 * BM + 2 * tg;
        double acc0 = cp[0], acc1 = cp[1];
#pragma unroll
        for (int kk = 0; kk < 8; kk += 4) {
            const int col = c0 + kk + tg;
            const double av = -Tf[(8 * BI + g) * NBP + col] * sda[col];
            const double bv = Tf[(8 * BM + g) * NBP + col];
            dmma884(acc0, acc1, av, bv);
        }
        cp[0] = acc0;
        cp[1] = acc1;
    };
    // X = L^-1 one block ROW at a time, by warps 6 and 7 (block columns of even / odd index), concurrently with the elimination of the following block columns
    // (row block i of L is final once diagonal block i is factored, i.e. at the barrier that opens iteration i):
    //   X_ii = inv(L_ii),   X_ij = -X_ii * sum_{k=j}^{i-1} L_ik X_kj   (j < i; the X_kj are earlier block rows)
    auto inverse_block_row = [&](const int i) {
        double* scr = xscr + warp * 64;
        if (lane < 8) {
            const int b0 = 8 * i, c = lane;
            double x[8];
#pragma unroll
            for (int r = 0; r < 8; r++) x[r] = (r == c) ? 1.0 : 0.0;
#pragma unroll
            for (int r = 1; r < 8; r++) {
                double sacc = 0.0;
#pragma unroll
                for (int k = 0; k < r; k++) sacc = fma(Tf[(b0 + r) * NBP + b0 + k], x[k], sacc);   // x[k] = 0 for k < c
                if (r > c) x[r] = -sacc;
            }
#pragma unroll
            for (int r = 0; r < 8; r++)
                if (r >= c) Xf[(b0 + r) * NBP + b0 + c] = x[r];
        }
        __syncwarp();
        for (int j = (warp & 1); j < i; j += 2) {      // both inverting warps hold X_ii (identical values); they split the block columns
            double s0 = 0.0, s1 = 0.0;
            for (int k = j; k < i; k++) {
#pragma unroll
                for (int kk = 0; kk < 8; kk += 4) {
                    const double av = Tf[(8 * i + g) * NBP + 8 * k + kk + tg];           // L_ik[g][kk + tg]
                    const double bv = Xf[(8 * k + kk + tg) * NBP + 8 * j + g];           // X_kj[kk + tg][g]
                    dmma884(s0, s1, av, bv);
                }
            }
            scr[g * 8 + 2 * tg] = s0;
            scr[g * 8 + 2 * tg + 1] = s1;
            __syncwarp();
            double x0 = 0.0, x1 = 0.0;
#pragma unroll
            for (int kk = 0; kk < 8; kk += 4) {
                const double av = -Xf[(8 * i + g) * NBP + 8 * i + kk + tg];               // -X_ii[g][kk + tg]
                const double bv = scr[(kk + tg) * 8 + g];                                 // S[kk + tg][g]
                dmma884(x0, x1, av, bv);
            }
            Xf[(8 * i + g) * NBP + 8 * j + 2 * tg] = x0;
            Xf[(8 * i + g) * NBP + 8 * j + 2 * tg + 1] = x1;
            __syncwarp();
        }
    };
    constexpr int NUPD = TILE_THREADS / 32 - 2;        // warps 0 .. 5 eliminate, warps 6 and 7 invert
    if (warp == 0 && lane == 0) diag_block(0);
#pragma unroll 1
    for (int kb = 0; kb < NB / 8; kb++) {
        const int c0 = 8 * kb;
#ifdef TILE_PROF
        const long long pp0 = clock64();
#endif
        __syncthreads();                   // diagonal block kb is factored; the update behind block kb-1 is complete
#ifdef TILE_PROF
        const long long pp1 = clock64();
#endif
        if (warp >= NUPD) {
            inverse_block_row(kb);         // reads only FINAL entries of T (columns < c0 + 8 of rows c0 .. c0 + 7) and writes X
            continue;                      // next: the barrier that opens iteration kb + 1
        }
        // (1b) the rows below the block, one thread per row:  l_i U = t_i  with  U = D L8' (forward substitution along the
        // eight columns; rows are independent)
        {
            const int row = c0 + 8 + (int)threadIdx.x;
            const int e0 = (int)threadIdx.x - 64;      // threads 64 .. 91: one multiplier of the diagonal block each
            if (e0 >= 0 && e0 < 28) {
                int r = 1, e = e0;
                while (e >= r) { e -= r; r++; }
                viol |= (fabs(Tf[(c0 + r) * NBP + c0 + e]) > 1.0 / pivot_u) ? 1 : 0;   // |d_j| >= u |T[r][j]|  <=>  |l_rj| <= 1/u
            }
            if (row < NB) {
                double t[8], l[8];
#pragma unroll
                for (int c = 0; c < 8; c++) t[c] = Tf[row * NBP + c0 + c];
                const double lim = 1.0 / pivot_u;
                int v = 0;
#pragma unroll
                for (int c = 0; c < 8; c++) {
                    double acc = t[c];
#pragma unroll
                    for (int j = 0; j < c; j++) acc = fma(-l[j] * sda[c0 + j], Tf[(c0 + c) * NBP + c0 + j], acc);   // U[j][c] = d_j l_cj
                    l[c] = acc * pinv[c];
                    v |= (fabs(l[c]) > lim) ? 1 : 0;
                }
                viol |= v;
#pragma unroll
                for (int c = 0; c < 8; c++) Tf[row * NBP + c0 + c] = l[c];
            }
        }
        asm volatile("bar.sync 1, %0;\n" ::"n"(NUPD * 32) : "memory");      // the eliminating warps only
#ifdef TILE_PROF
        const long long pp2 = clock64();
#endif
        // (2) rank-8 update of the lower tiles behind the block.  Warp 0 takes the next diagonal tile first and thread 0
        // goes straight on to factor it (pinv and the next sda entries are free again: (1b) is over), overlapped with
        // the other six warps updating the remaining tiles.
        if (kb < NB / 8 - 1) {
            const int nbt = NB / 8 - 1 - kb;
            const int ntiles = nbt * (nbt + 1) / 2;
            if (warp == 0) {
                update_tile(kb, 0);
                __syncwarp();
#ifdef TILE_PROF
                const long long pp3 = clock64();
#endif
                if (lane == 0) diag_block(c0 + 8);
#ifdef TILE_PROF
                if (lane == 0) printf("TILE_PROF kb=%d wait_top=%lld rows_below=%lld upd0=%lld diag=%lld\n", kb, pp1 - pp0, pp2 - pp1, pp3 - pp2, clock64() - pp3);
#endif
            } else {
                for (int t = warp; t < ntiles; t += NUPD - 1) update_tile(kb, t);
#ifdef TILE_PROF
                if (warp == 1 && lane == 0) printf("TILE_PROF kb=%d warp1 upd=%lld (tiles %d)\n", kb, clock64() - pp2, ntiles);
#endif
            }
        }
    }
    __syncthreads();
#ifdef TILE_PROF
    cy_x = clock64() - cy_t;
    if (warp == 0 && lane == 0) printf("TILE_PROF blocked: total=%lld cycles\n", cy_x);
#endif
}

__global__ void __launch_bounds__(TILE_THREADS) ldlt_tile_kernel(double* __restrict__ A, int ld, int nb,
                                                                 double* __restrict__ LinvP, double* __restrict__ dinv_a,
                                                                 double* __restrict__ dinv_b, double* __restrict__ d_a,
                                                                 double* __restrict__ d_b, int* __restrict__ kind,
                                                                 int* __restrict__ perm_out, int* __restrict__ counts,
                                                                 double* __restrict__ dstat, const double pivot_u,
                                                                 int* __restrict__ sig, const int blocked) {
    extern __shared__ __align__(16) double tsm[];
    if (sig != nullptr && threadIdx.x == 0) atomicExch(sig, 1);   // "this far" marker for a delayed background factorisation
    double* Tf = tsm;                 // T[i][m] = Tf[i * NBP + m]   (authoritative only inside slow steps / at the ends)
    double* Xf = tsm + NB * NBP;      // X[i][m] = Xf[i * NBP + m]
    __shared__ double sda[NB], sdb[NB];
    __shared__ double prowT[2][NB];   // unified pivot row (T part right of the pivot, X part left of it), double buffered
    __shared__ int sperm[NB], skind[NB];
    __shared__ TileDec sdec[2];
    __shared__ int s_kp2, s_kstep2;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const double BK_ALPHA = 0.6403882032022076;
    // an inertia test that has already failed (more negative pivots than the limit in counts[5]) abandons the rest
    // of the factorisation: every remaining kernel of the captured graph returns at once
    if (tid == 0) s_kp2 = counts[4];
    __syncthreads();
    if (s_kp2) return;
    __syncthreads();
    const int trc = (tid == 0) ? trace_begin(TR_TILE, counts) : -1;
#ifdef TILE_PROF
    long long tp0 = clock64(), tp1 = 0, tp2 = 0, tp3 = 0;
    int nslow = 0;
#endif

    // load the tile from its LOWER triangle (coalesced rows, all 16 loads of a thread in flight at once), mirror
    // through shared memory, identity-pad to NB
    {
        double v[NB * NB / TILE_THREADS];
#pragma unroll
        for (int it = 0; it < NB * NB / TILE_THREADS; it++) {
            const int idx = tid + it * TILE_THREADS;
            const int i = idx / NB, j = idx % NB;
            v[it] = (i < nb && j <= i) ? A[(size_t)i * ld + j] : ((i == j) ? 1.0 : 0.0);
        }
#pragma unroll
        for (int it = 0; it < NB * NB / TILE_THREADS; it++) {
            const int idx = tid + it * TILE_THREADS;
            const int i = idx / NB, j = idx % NB;
            if (j <= i) {
                Tf[i * NBP + j] = v[it];
                Tf[j * NBP + i] = v[it];
            }
            Xf[i * NBP + j] = (i == j) ? 1.0 : 0.0;
        }
    }
    if (tid < NB) {
        sperm[tid] = tid;
        skind[tid] = 0;
        sda[tid] = 1.0;
        sdb[tid] = 0.0;
    }
    __syncthreads();
    const int m0 = lane, m1 = lane + 32;
    // ONE register per (row, column): column m is "T-type" (trailing matrix) until pivot step m and "X-type" (L^-1)
    // afterwards.  At its switch step the L entries of the column are written to Tf once and the registers restart
    // from the X initial values; Xf keeps the identity for columns that have not switched yet.
    double a[TILE_RPW][2];
#pragma unroll
    for (int k = 0; k < TILE_RPW; k++) {
        const int row = warp + 8 * k;
        a[k][0] = Tf[row * NBP + m0]; a[k][1] = Tf[row * NBP + m1];
    }
    double* prow0 = &prowT[0][0];     // unified pivot row, double buffered: prow[par][m]
    __shared__ double pdiag[4];   // [parity][d, 1/d] of the published pivot

    // =====================================================================================================
    // FAST ATTEMPT: no pivot search on the critical path.  The tile is eliminated in natural order with THRESHOLD
    // pivoting: step j is acceptable iff |d_j| >= u * max_i |T[i][j]| (u = pivot_u, default 0.01 as in the
    // multifrontal codes interior-point solvers use; u = 0.64 would be Bunch-Kaufman's first-level test).  Every
    // lane checks the test on the pivot-column entries it reads anyway and keeps a sticky flag.  The attempt is
    // accepted if no test failed, or if all pivots came out with the same sign (the tile was definite, so the
    // unpivoted elimination is unconditionally stable).  Otherwise the tile is reloaded and the full
    // Bunch-Kaufman loop below runs.  Element growth is bounded by (1 + 1/u) per accepted step; accuracy is
    // restored by the iterative refinement against the unreduced KKT system, and the engine re-factors with
    // u = 0.64 if the refined residual is ever poor.
    // Per step the serial chain is: barrier -> LDS pivot row -> rcp -> 2 FMA on the next pivot row -> STS.
    // =====================================================================================================
    int fast_ok = 0;
    if (blocked) {
        __shared__ double xscr[(TILE_THREADS / 32) * 64];
        __shared__ int s_sign[2];
        __shared__ double s_pinv[8];
        int viol = 0;
        bool allpos = true, allneg = true;
        tile_fast_blocked(Tf, Xf, sda, xscr, s_pinv, nb, lane, warp, pivot_u, viol, allpos, allneg);
        if (warp == 0) {   // signs of the pivots, read off sda (the routine ends with a CTA barrier)
            for (int p2 = lane; p2 < nb; p2 += 32) { allpos = allpos && (sda[p2] > 0.0); allneg = allneg && (sda[p2] < 0.0); }
            allpos = __all_sync(0xffffffffu, allpos);
            allneg = __all_sync(0xffffffffu, allneg);
            if (lane == 0) { s_sign[0] = allpos ? 1 : 0; s_sign[1] = allneg ? 1 : 0; }
        }
        const int anyviol = __syncthreads_or(viol);
        const bool ap = s_sign[0] != 0, an = s_sign[1] != 0;
        fast_ok = ((!anyviol) || ap || an) ? 1 : 0;
        if (!(ap || an)) {
            bool haszero = false;
            for (int p2 = 0; p2 < nb; p2++) haszero = haszero || (sda[p2] == 0.0);
            if (haszero) fast_ok = 0;
        }
        if (fast_ok && tid < NB) { sdb[tid] = 0.0; skind[tid] = 0; }
    } else {
        double b[TILE_RPW][2];     // transposed ownership: rows lane / lane+32, columns warp + 8k
#pragma unroll
        for (int k = 0; k < TILE_RPW; k++) {
            b[k][0] = Tf[lane * NBP + warp + 8 * k];
            b[k][1] = Tf[(lane + 32) * NBP + warp + 8 * k];
        }
        if (lane == 0) {           // row 0 is the first pivot row
#pragma unroll
            for (int k = 0; k < TILE_RPW; k++) prow0[warp + 8 * k] = b[k][0];
            if (warp == 0) {
                prow0[0] = 1.0;
                pdiag[0] = b[0][0];
                pdiag[1] = (b[0][0] != 0.0) ? fast_rcp(b[0][0]) : 0.0;
            }
        }
        __syncthreads();
        int viol = 0;
        bool allpos = true, allneg = true;
        tile_fast_block<0>(b, nb, lane, warp, tid, Tf, prow0, pdiag, sda, viol, allpos, allneg, pivot_u);
        tile_fast_block<1>(b, nb, lane, warp, tid, Tf, prow0, pdiag, sda, viol, allpos, allneg, pivot_u);
        tile_fast_block<2>(b, nb, lane, warp, tid, Tf, prow0, pdiag, sda, viol, allpos, allneg, pivot_u);
        tile_fast_block<3>(b, nb, lane, warp, tid, Tf, prow0, pdiag, sda, viol, allpos, allneg, pivot_u);
        tile_fast_block<4>(b, nb, lane, warp, tid, Tf, prow0, pdiag, sda, viol, allpos, allneg, pivot_u);
        tile_fast_block<5>(b, nb, lane, warp, tid, Tf, prow0, pdiag, sda, viol, allpos, allneg, pivot_u);
        tile_fast_block<6>(b, nb, lane, warp, tid, Tf, prow0, pdiag, sda, viol, allpos, allneg, pivot_u);
        tile_fast_block<7>(b, nb, lane, warp, tid, Tf, prow0, pdiag, sda, viol, allpos, allneg, pivot_u);
        const int anyviol = __syncthreads_or(viol);
        // accept: no first-level violation anywhere, or a definite tile (all pivots of one sign, none zero)
        fast_ok = ((!anyviol) || allpos || allneg) ? 1 : 0;
        if (!(allpos || allneg)) {   // mixed signs are fine without violations, zero pivots are not
            bool haszero = false;
            for (int p2 = 0; p2 < nb; p2++) haszero = haszero || (sda[p2] == 0.0);
            if (haszero) fast_ok = 0;
        }
        if (fast_ok) {
            // all columns < nb have switched: the registers hold X; hand over to the common output stage
#pragma unroll
            for (int k = 0; k < TILE_RPW; k++) {
                const int c = warp + 8 * k;
                if (c < nb) {
                    Xf[lane * NBP + c] = b[k][0];
                    Xf[(lane + 32) * NBP + c] = b[k][1];
                }
            }
            if (tid < NB) { sdb[tid] = 0.0; skind[tid] = 0; }
        }
    }
    if (!fast_ok) {
        // reload the tile and the bookkeeping, then fall through to the pivoting loop
        __syncthreads();
        for (int idx = tid; idx < NB * NB; idx += TILE_THREADS) {
            const int i = idx / NB, jj = idx % NB;
            if (jj <= i) {
                const double v = (i < nb) ? A[(size_t)i * ld + jj] : ((i == jj) ? 1.0 : 0.0);
                Tf[i * NBP + jj] = v;
                Tf[jj * NBP + i] = v;
            }
            Xf[i * NBP + jj] = (i == jj) ? 1.0 : 0.0;
        }
        if (tid < NB) { sperm[tid] = tid; skind[tid] = 0; sda[tid] = 1.0; sdb[tid] = 0.0; }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < TILE_RPW; k++) {
            const int row = warp + 8 * k;
            a[k][0] = Tf[row * NBP + m0]; a[k][1] = Tf[row * NBP + m1];
        }
        if (warp == 0) {
            prow0[m0] = (m0 == 0) ? 1.0 : a[0][0]; prow0[m1] = a[0][1];
            tile_search(0, nb, lane, a[0][0], a[0][1], &sdec[0]);
        }
        __syncthreads();
    }
#ifdef TILE_PROF
    tp1 = clock64();
#endif

    int j = fast_ok ? nb : 0, par = 0;
    while (j < nb) {
        const int code = sdec[par].code;
        int jn;
        if (code == ((1 << 1) | (j << 8))) {
            // ================= fast step: 1x1 pivot on the diagonal, no interchange =================
            const double* pr = prow0 + par * NB;
            const double d = sdec[par].d, dinv = sdec[par].dinv;
            const double q0 = pr[m0], q1 = pr[m1];
            double ua[TILE_RPW];
#pragma unroll
            for (int k = 0; k < TILE_RPW; k++) ua[k] = pr[warp + 8 * k];      // all smem loads of the step up front
            if (tid == 0) { sda[j] = d; sdb[j] = 0.0; skind[j] = 0; }
            const bool sw0 = (m0 == j), sw1 = (m1 == j);
            jn = j + 1;
            // ---- phase A (owner of the next pivot row only): update THAT row, publish it, first-level test
            const bool ownerA = (jn < nb) && ((jn & 7) == warp);
            if (ownerA) {
#pragma unroll
                for (int k = 0; k < TILE_RPW; k++) {
                    if ((jn >> 3) == k) {
                        const double uu = ua[k];
                        a[k][0] = fma(-(uu * q0), dinv, a[k][0]);
                        a[k][1] = fma(-(uu * q1), dinv, a[k][1]);
                        if (sw0 || sw1) {
                            const double lij = uu * dinv;
                            Tf[jn * NBP + j] = lij;
                            if (sw0) a[k][0] = -lij; else a[k][1] = -lij;
                        }
                        double* pw = prow0 + (par ^ 1) * NB;
                        pw[m0] = (m0 == jn) ? 1.0 : a[k][0];
                        pw[m1] = (m1 == jn) ? 1.0 : a[k][1];
                        tile_search(jn, nb, lane, a[k][0], a[k][1], &sdec[par ^ 1]);
                    }
                }
            }
            // ---- phase B (everybody): the remaining active rows; inactive rows get a zero multiplier
#pragma unroll
            for (int k = 0; k < TILE_RPW; k++) {
                const int row = warp + 8 * k;
                const bool act = (row > j) && (row < nb) && !(ownerA && row == jn);
                const double uu = act ? ua[k] : 0.0;
                a[k][0] = fma(-(uu * q0), dinv, a[k][0]);
                a[k][1] = fma(-(uu * q1), dinv, a[k][1]);
            }
            // ---- the column that switches from T-type to X-type at this step (one lane per warp)
            if (sw0 || sw1) {
#pragma unroll
                for (int k = 0; k < TILE_RPW; k++) {
                    const int row = warp + 8 * k;
                    if (!(ownerA && row == jn)) {
                        const bool act = (row > j) && (row < nb);
                        const double lij = act ? ua[k] * dinv : 0.0;
                        if (act) Tf[row * NBP + j] = lij;
                        const double xi = act ? -lij : ((row == j) ? 1.0 : 0.0);   // X[row][j] after this step
                        if (sw0) a[k][0] = xi; else a[k][1] = xi;
                    }
                }
            }
        } else {
            // ================= slow step: second-level test / interchange / 2x2 pivot, in shared memory ==========
#ifdef TILE_PROF
            nslow++;
#endif
#pragma unroll
            for (int k = 0; k < TILE_RPW; k++) {
                const int row = warp + 8 * k;
                if (m0 >= j) Tf[row * NBP + m0] = a[k][0]; else Xf[row * NBP + m0] = a[k][0];
                if (m1 >= j) Tf[row * NBP + m1] = a[k][1]; else Xf[row * NBP + m1] = a[k][1];
            }
            __syncthreads();
            if (warp == 0) {
                // second-level test of LAPACK dsytf2 (needs row r)
                const TileDec dec = sdec[par];
                int kp = j, kstep = 1;
                if (dec.code & 1) {
                    const int r = dec.r;
                    const int a0 = j + lane, a1 = a0 + 32;
                    double rm = 0.0;
                    if (a0 < nb && a0 != r) rm = fabs(Tf[r * NBP + a0]);
                    if (a1 < nb && a1 != r) rm = fmax(rm, fabs(Tf[r * NBP + a1]));
                    rm = warp_absmax_bits(rm);
                    if (dec.absakk * rm >= BK_ALPHA * dec.colmax * dec.colmax) { kp = j; kstep = 1; }
                    else if (fabs(Tf[r * NBP + r]) >= BK_ALPHA * rm) { kp = r; kstep = 1; }
                    else { kp = r; kstep = 2; }
                }
                if (lane == 0) { s_kp2 = kp; s_kstep2 = kstep; }
            }
            __syncthreads();
            const int kp = s_kp2, kstep = s_kstep2;
            const int kk = j + kstep - 1;
            if (kp != kk) {   // symmetric interchange kk <-> kp on T and X (rows, then columns)
                if (tid < NB) {
                    double t0 = Tf[kk * NBP + tid]; Tf[kk * NBP + tid] = Tf[kp * NBP + tid]; Tf[kp * NBP + tid] = t0;
                    double x0 = Xf[kk * NBP + tid]; Xf[kk * NBP + tid] = Xf[kp * NBP + tid]; Xf[kp * NBP + tid] = x0;
                }
                __syncthreads();
                if (tid < NB) {
                    double t0 = Tf[tid * NBP + kk]; Tf[tid * NBP + kk] = Tf[tid * NBP + kp]; Tf[tid * NBP + kp] = t0;
                    double x0 = Xf[tid * NBP + kk]; Xf[tid * NBP + kk] = Xf[tid * NBP + kp]; Xf[tid * NBP + kp] = x0;
                }
                if (tid == 0) { int p = sperm[kk]; sperm[kk] = sperm[kp]; sperm[kp] = p; }
                __syncthreads();
            }
            double* P0 = (m0 > kk) ? Tf : Xf;
            double* P1 = (m1 > kk) ? Tf : Xf;
            if (kstep == 1) {
                const double d = Tf[j * NBP + j];
                const double dinv = (d != 0.0) ? 1.0 / d : 0.0;
                if (tid == 0) { sda[j] = d; sdb[j] = 0.0; skind[j] = 0; }
                const double pj0 = P0[j * NBP + m0], pj1 = P1[j * NBP + m1];
#pragma unroll
                for (int k = 0; k < TILE_RPW; k++) {
                    const int row = warp + 8 * k;
                    if (row > j && row < nb && d != 0.0) {
                        const double ur = Tf[j * NBP + row];
                        P0[row * NBP + m0] = fma(-(ur * pj0), dinv, P0[row * NBP + m0]);
                        P1[row * NBP + m1] = fma(-(ur * pj1), dinv, P1[row * NBP + m1]);
                        if (m0 == j || m1 == j) Tf[row * NBP + j] = ur * dinv;
                    }
                }
            } else {
                const double a11 = Tf[j * NBP + j], a21 = Tf[(j + 1) * NBP + j], a22 = Tf[(j + 1) * NBP + j + 1];
                if (tid == 0) {
                    sda[j] = a11; sdb[j] = a21; sda[j + 1] = a22; sdb[j + 1] = 0.0;
                    skind[j] = 1; skind[j + 1] = 2;
                }
                // LAPACK's scaled 2x2 inverse (dsytf2): robust against overflow of the determinant
                const double d11 = a22 / a21, d22 = a11 / a21;
                const double tt = 1.0 / (d11 * d22 - 1.0);
                const double d21i = tt / a21;
                const double pj0 = P0[j * NBP + m0], pj1 = P1[j * NBP + m1];
                const double pk0 = P0[(j + 1) * NBP + m0], pk1 = P1[(j + 1) * NBP + m1];
#pragma unroll
                for (int k = 0; k < TILE_RPW; k++) {
                    const int row = warp + 8 * k;
                    if (row > kk && row < nb) {
                        const double u = Tf[j * NBP + row], v = Tf[(j + 1) * NBP + row];
                        P0[row * NBP + m0] -= d21i * ((d11 * (u * pj0) + d22 * (v * pk0)) - (v * pj0 + u * pk0));
                        P1[row * NBP + m1] -= d21i * ((d11 * (u * pj1) + d22 * (v * pk1)) - (v * pj1 + u * pk1));
                        if (m0 == j || m1 == j) Tf[row * NBP + j] = d21i * (d11 * u - v);
                        if (m0 == j + 1 || m1 == j + 1) Tf[row * NBP + j + 1] = d21i * (d22 * v - u);
                    }
                }
                // (T[j+1][j] keeps a21 until the output stage)
            }
            __syncthreads();
            jn = j + kstep;
#pragma unroll
            for (int k = 0; k < TILE_RPW; k++) {
                const int row = warp + 8 * k;
                a[k][0] = (m0 >= jn) ? Tf[row * NBP + m0] : Xf[row * NBP + m0];
                a[k][1] = (m1 >= jn) ? Tf[row * NBP + m1] : Xf[row * NBP + m1];
                if (row == jn && jn < nb) {
                    double* pw = prow0 + (par ^ 1) * NB;
                    pw[m0] = (m0 == jn) ? 1.0 : a[k][0];
                    pw[m1] = (m1 == jn) ? 1.0 : a[k][1];
                    tile_search(jn, nb, lane, a[k][0], a[k][1], &sdec[par ^ 1]);
                }
            }
        }
        __syncthreads();
        j = jn;
        par ^= 1;
    }
    // pivoting loop: all columns < nb have switched, the registers hold X
    if (!fast_ok) {
#pragma unroll
        for (int k = 0; k < TILE_RPW; k++) {
            const int row = warp + 8 * k;
            if (m0 < nb) Xf[row * NBP + m0] = a[k][0];
            if (m1 < nb) Xf[row * NBP + m1] = a[k][1];
        }
    }
    __syncthreads();
#ifdef TILE_PROF
    tp2 = clock64();
#endif

    // ---- outputs
    // LinvP[r][perm[m]] = X[r][m]   (column scatter folds the permutation into the inverse)
    for (int idx = tid; idx < NB * NB; idx += TILE_THREADS) {
        const int r = idx / NB, m = idx % NB;
        LinvP[r * NB + sperm[m]] = (m <= r) ? Xf[r * NBP + m] : 0.0;
    }
    // (the tile's own L is not written back: every consumer of a diagonal tile uses LinvP and the D^-1 blocks)
    if (tid < NB) {
        const int p = tid;
        double ia = 0.0, ib = 0.0;
        if (p < nb) {
            if (skind[p] == 0) {
                ia = (sda[p] != 0.0) ? 1.0 / sda[p] : 0.0;
            } else {
                const int p0 = (skind[p] == 1) ? p : p - 1;
                const double a11 = sda[p0], a21 = sdb[p0], a22 = sda[p0 + 1];
                const double d11 = a22 / a21, d22 = a11 / a21;
                const double tt = 1.0 / (d11 * d22 - 1.0);
                const double d21i = tt / a21;
                // D^-1 = d21i * [[d11, -1], [-1, d22]]
                ia = (skind[p] == 1) ? d21i * d11 : d21i * d22;
                ib = (skind[p] == 1) ? -d21i : 0.0;
            }
        } else {
            ia = 1.0;
        }
        dinv_a[p] = ia;
        dinv_b[p] = ib;
        d_a[p] = sda[p];
        d_b[p] = sdb[p];
        kind[p] = (p < nb) ? skind[p] : 0;
        if (perm_out) perm_out[p] = sperm[p];
    }
    // inertia and |eig(D)| extremes: one warp, two positions per lane
    if (warp == 0) {
        int neg = 0, zero = 0, pos = 0;
        double mn = INFINITY, mx = 0.0;
        for (int p = lane; p < nb; p += 32) {
            if (skind[p] == 0) {
                const double d = sda[p];
                if (d > 0.0) pos++; else if (d < 0.0) neg++; else zero++;
                mn = fmin(mn, fabs(d)); mx = fmax(mx, fabs(d));
            } else if (skind[p] == 1) {
                const double a = sda[p], b = sdb[p], c = sda[p + 1];
                const double hm = 0.5 * (a + c), hd = 0.5 * (a - c);
                const double rad = sqrt(hd * hd + b * b);
                const double e1 = hm + rad, e2 = hm - rad;
                if (e1 > 0.0) pos++; else if (e1 < 0.0) neg++; else zero++;
                if (e2 > 0.0) pos++; else if (e2 < 0.0) neg++; else zero++;
                mn = fmin(mn, fmin(fabs(e1), fabs(e2))); mx = fmax(mx, fmax(fabs(e1), fabs(e2)));
            }
        }
        neg = __reduce_add_sync(0xffffffffu, neg);
        zero = __reduce_add_sync(0xffffffffu, zero);
        pos = __reduce_add_sync(0xffffffffu, pos);
        mn = warp_min(mn);
        mx = warp_max(mx);
        if (lane == 0) {
            const int negs = counts[0] + neg;
            counts[0] = negs; counts[1] += zero; counts[2] += pos;
            if (negs > counts[5]) counts[4] = 1;
            dstat[0] = fmin(dstat[0], mn); dstat[1] = fmax(dstat[1], mx);
        }
    }
    __syncthreads();
    trace_end(trc);
#ifdef TILE_PROF
    __syncthreads();
    tp3 = clock64();
    if (tid == 0) printf("TILE_PROF nb=%d load=%lld loop=%lld out=%lld slow_steps=%d fast_ok=%d\n", nb, tp1 - tp0, tp2 - tp1, tp3 - tp2, nslow, fast_ok);
#endif
}

// Fused panel step for 64 rows per CTA (4 warps, DMMA):  W = B * LinvP^T  (W = Lpanel * D) and
// Lpanel = W * D^-1 (block diagonal, 1x1 / 2x2).  B (rows x 64, leading dimension ld) is overwritten with Lpanel,
// W goes to Wout (leading dimension ldw).  Both operands are staged whole in smem (K = 64), row stride 68 doubles
// keeps the fragment loads at the minimum two wavefronts.
constexpr int P_LDS = 68;
constexpr int PANEL_SMEM = 2 * NB * P_LDS * 8;
__global__ void __launch_bounds__(128) ldlt_panel_kernel(double* __restrict__ B, int ld, int rows,
                                                         const double* __restrict__ LinvP,
                                                         const double* __restrict__ dinv_a,
                                                         const double* __restrict__ dinv_b, const int* __restrict__ kind,
                                                         double* __restrict__ Wout, int ldw,
                                                         const int* __restrict__ ctrl) {
    extern __shared__ __align__(16) double psm[];
    if (ctrl) {
        __shared__ int s_abort;
        if (threadIdx.x == 0) s_abort = *reinterpret_cast<const volatile int*>(ctrl + 4);
        __syncthreads();
        if (s_abort) return;
    }
    double* As = psm;
    double* Bs = psm + NB * P_LDS;
    __shared__ double sia[NB], sib[NB];
    __shared__ int skd[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int r0 = blockIdx.x * NB;
    const int trc = (tid == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1)) ? trace_begin(TR_PANEL, ctrl) : -1;
    if (tid < NB) { sia[tid] = dinv_a[tid]; sib[tid] = dinv_b[tid]; skd[tid] = kind[tid]; }
#pragma unroll
    for (int i = 0; i < 16; i++) {
        const int c = tid + 128 * i;
        const int row = c >> 5, kc = (c & 31) * 2;
        const bool ok = (r0 + row) < rows;
        cp_async16(As + row * P_LDS + kc, ok ? B + (size_t)(r0 + row) * ld + kc : B, ok ? 16 : 0);
        cp_async16(Bs + row * P_LDS + kc, LinvP + row * NB + kc, 16);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    double acc[2][8][2];
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 8; b++) acc[a][b][0] = acc[a][b][1] = 0.0;
    const double* as = As + (warp * 16 + g) * P_LDS + tg;
    const double* bs = Bs + g * P_LDS + tg;
#pragma unroll
    for (int kk = 0; kk < 16; kk++) {
        double af[2], bf[8];
#pragma unroll
        for (int mt = 0; mt < 2; mt++) af[mt] = as[mt * 8 * P_LDS + kk * 4];
#pragma unroll
        for (int nt = 0; nt < 8; nt++) bf[nt] = bs[nt * 8 * P_LDS + kk * 4];
#pragma unroll
        for (int mt = 0; mt < 2; mt++)
#pragma unroll
            for (int nt = 0; nt < 8; nt++) dmma884(acc[mt][nt][0], acc[mt][nt][1], af[mt], bf[nt]);
    }
    __syncthreads();   // everybody is done reading As: reuse it for the W tile
#pragma unroll
    for (int mt = 0; mt < 2; mt++)
#pragma unroll
        for (int nt = 0; nt < 8; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) As[(warp * 16 + mt * 8 + g) * P_LDS + nt * 8 + tg * 2 + e] = acc[mt][nt][e];
    __syncthreads();
    const int col = tid & 63;
    const int k = skd[col];
    const double ia = sia[col];
    const double ibn = (k == 1) ? sib[col] : ((k == 2) ? sib[col - 1] : 0.0);
    const int nbr = (k == 1) ? col + 1 : ((k == 2) ? col - 1 : col);
#pragma unroll 4
    for (int i = 0; i < 32; i++) {
        const int r = (tid >> 6) + 2 * i;
        const int gr = r0 + r;
        if (gr < rows) {
            const double wv = As[r * P_LDS + col];
            Wout[(size_t)gr * ldw + col] = wv;
            B[(size_t)gr * ld + col] = wv * ia + As[r * P_LDS + nbr] * ibn;
        }
    }
    trace_end(trc);
}

// Chain shortcut ("mini" step).  MINI_CTAS CTAs of 8 warps (one SM's DMMA pipe needs 2.1 us per 64^3 product: the second
// product, T -= W L', is split by rows over the CTAs; each computes W = B LinvP' itself because it needs all of L): the panel computation of ldlt_panel_kernel for the 64 rows right
// below tile k (W = B LinvP', L = W D^-1, stored exactly as the panel kernel stores them), followed by the update of the
// NEXT diagonal tile  T -= W L'.  Tile k+1 can then be factored while the rest of panel step k (all other rows, all
// other tiles of the in-panel update) runs on the update stream: the serial chain per tile step is tile + mini instead
// of tile + panel + update.  Latency-oriented: B, LinvP and T are fetched together with cp.async at entry, every warp
// owns one 8-row fragment strip (half the DMMA chain of a 4-warp layout).
constexpr int MINI_THREADS = 256;
constexpr int MINI_CTAS = 4;      // the update of the next diagonal tile is split over four CTAs (the first product is redundant)
constexpr int MINI_SMEM = 3 * NB * P_LDS * 8;
// The four CTAs read ALL 64 rows of B at entry and each later overwrites its own 16 rows with L in place: a CTA that starts
// late (no free SM while the trailing updates saturate the GPU) would read rows its siblings have already overwritten.
// The CTAs therefore form one thread-block cluster (co-scheduled) and pass a cluster barrier between the loads and the
// first store; the barrier's arrive / wait are split around the first product, which hides its latency.
__global__ void __cluster_dims__(MINI_CTAS, 1, 1) __launch_bounds__(MINI_THREADS) ldlt_mini_kernel(double* __restrict__ B, int ld, int rows,
                                                                 const double* __restrict__ LinvP,
                                                                 const double* __restrict__ dinv_a,
                                                                 const double* __restrict__ dinv_b, const int* __restrict__ kind,
                                                                 double* __restrict__ Wout, int ldw, double* __restrict__ T,
                                                                 const int* __restrict__ ctrl) {
    extern __shared__ __align__(16) double psm[];
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = *reinterpret_cast<const volatile int*>(ctrl + 4);
    __syncthreads();
    if (s_abort) return;
    double* As = psm;
    double* Bs = psm + NB * P_LDS;
    double* Ts = psm + 2 * NB * P_LDS;
    __shared__ double sia[NB], sib[NB];
    __shared__ int skd[NB];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tg = lane & 3;
    const int nn = min(NB, rows);
    const int trc = (tid == 0 && blockIdx.x == 0) ? trace_begin(TR_MINI, ctrl) : -1;
    if (tid < NB) { sia[tid] = dinv_a[tid]; sib[tid] = dinv_b[tid]; skd[tid] = kind[tid]; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int c = tid + MINI_THREADS * i;
        const int row = c >> 5, kc = (c & 31) * 2;
        const bool ok = row < nn;
        cp_async16(As + row * P_LDS + kc, ok ? B + (size_t)row * ld + kc : B, ok ? 16 : 0);
        cp_async16(Bs + row * P_LDS + kc, LinvP + row * NB + kc, 16);
        const bool okt = ok && (kc < nn);                              // nn is even unless the matrix order is odd
        cp_async16(Ts + row * P_LDS + kc, okt ? T + (size_t)row * ld + kc : T, okt ? min(16, (nn - kc) * 8) : 0);
    }
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");     // this CTA holds its copy of B
    double acc[8][2];
    const double* as = As + (warp * 8 + g) * P_LDS + tg;
    const double* bs = Bs + g * P_LDS + tg;
    auto product = [&]() {
#pragma unroll
        for (int b = 0; b < 8; b++) acc[b][0] = acc[b][1] = 0.0;
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            double bf[8];
            const double af = as[kk * 4];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) bf[nt] = bs[nt * 8 * P_LDS + kk * 4];
#pragma unroll
            for (int nt = 0; nt < 8; nt++) dmma884(acc[nt][0], acc[nt][1], af, bf[nt]);
        }
    };
    product();                       // W = B * LinvP'   (every CTA: the second product needs all of L)
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");       // every sibling holds its copy of B
    __syncthreads();                 // everybody is done reading As / Bs
#pragma unroll
    for (int nt = 0; nt < 8; nt++)
#pragma unroll
        for (int e = 0; e < 2; e++) As[(warp * 8 + g) * P_LDS + nt * 8 + tg * 2 + e] = acc[nt][e];
    __syncthreads();
    const int q = blockIdx.x;        // this CTA's quarter: rows 16 q .. 16 q + 15 of W / L / T
    {
        const int col = tid & 63;
        const int k = skd[col];
        const double ia = sia[col];
        const double ibn = (k == 1) ? sib[col] : ((k == 2) ? sib[col - 1] : 0.0);
        const int nbr = (k == 1) ? col + 1 : ((k == 2) ? col - 1 : col);
#pragma unroll 4
        for (int i = 0; i < 16; i++) {
            const int r = (tid >> 6) + 4 * i;
            const double wv = As[r * P_LDS + col];
            const double lv = wv * ia + As[r * P_LDS + nbr] * ibn;
            Bs[r * P_LDS + col] = lv;                 // rows >= nn are exact zeros (zero-filled loads)
            if (r < nn && (r >> 4) == q) {
                Wout[(size_t)r * ldw + col] = wv;
                B[(size_t)r * ld + col] = lv;
            }
        }
    }
    __syncthreads();
    // T[16 q .. 16 q + 15, :] -= W L' : warp w -> 8-row fragment (w & 1), column blocks 2 (w >> 1), 2 (w >> 1) + 1
    {
        double a2[2][2] = {{0.0, 0.0}, {0.0, 0.0}};
        const int rr = 16 * q + 8 * (warp & 1) + g;
        const int cb = 2 * (warp >> 1);
        const double* as2 = As + rr * P_LDS + tg;
        const double* bs2 = Bs + (cb * 8 + g) * P_LDS + tg;
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            const double af = as2[kk * 4];
            const double b0 = bs2[kk * 4], b1 = bs2[8 * P_LDS + kk * 4];
            dmma884(a2[0][0], a2[0][1], af, b0);
            dmma884(a2[1][0], a2[1][1], af, b1);
        }
#pragma unroll
        for (int nt = 0; nt < 2; nt++)
#pragma unroll
            for (int e = 0; e < 2; e++) {
                const int j = (cb + nt) * 8 + tg * 2 + e;
                if (rr < nn && j < nn) T[(size_t)rr * ld + j] = Ts[rr * P_LDS + j] - a2[nt][e];
            }
    }
    __syncthreads();
    trace_end(trc);
}

__global__ void ldlt_reset_kernel(int* counts, double* dstat, unsigned* ticket) {
    counts[0] = counts[1] = counts[2] = counts[3] = counts[4] = 0;
    counts[6] = 0;      // error word of the tcgen05 trailing updates (1: non-finite operand, 4: pipeline timeout)
    dstat[0] = INFINITY;
    dstat[1] = 0.0;
    ticket[0] = ticket[1] = 0;
}

// negative-pivot limit of the NEXT factorisation(s) of this workspace (0x7fffffff = never abandon); a plain stream-ordered
// store outside the captured graph
__global__ void ldlt_limit_kernel(int* counts, int limit) { counts[5] = limit; }
inline int ldlt_set_neg_limit(LdltWs& w, int limit) {
    if (w.neg_limit == limit) return 0;
    ldlt_limit_kernel<<<1, 1, 0, w.st>>>(w.counts, limit);
    LAUNCHED();
    w.neg_limit = limit;
    return 0;
}

// Bounded device-side wait for the marker above (background factorisation: start only when the foreground one is past
// its update-heavy first third).  Gives up after ~20 ms so that a missing marker can never hang the stream.
__global__ void ldlt_wait_sig_kernel(int* sig) {
    const unsigned long long t0 = trace_now();
    while (atomicCAS(sig, 1, 0) != 1) {
        __nanosleep(500);
        if (trace_now() - t0 > 20000000ull) break;
    }
}

inline int ldlt_init_attrs() {
    CU(cudaFuncSetAttribute(ldlt_tile_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CU(cudaFuncSetAttribute(ldlt_panel_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PANEL_SMEM));
    CU(cudaFuncSetAttribute(ldlt_mini_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MINI_SMEM));
    CU(cudaFuncSetAttribute(gemm_nt_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G_SMEM));
    CU(cudaFuncSetAttribute(gemm_nt_sub64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, S_SMEM));
    CU(cudaFuncSetAttribute(gemm_nt_sub64_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T_SMEM));
    CU(cudaFuncSetAttribute(oz_syrk_kernel<64, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, OzShape<64, 6>::SMEM));
    return 0;
}

// Factor w.A in place (two-level blocking with look-ahead).  Outer panels of NBO = 256 columns: inside a panel the
// 64-wide tile steps update only the panel's own remaining columns (K = 64, small); the trailing matrix is updated
// ONCE per outer panel with K = 256 -- 4x less read-modify-write traffic on A22 and a contraction long enough to keep
// the DMMA pipeline full.  The serial chain is the 72 tile steps, so the trailing update is split: the next
// panel's 256 columns are updated on the main stream (the chain needs them), everything to the right of them on a
// side stream, overlapped with the next panel's tile steps (W is double buffered).  Results stay on the device
// (counts/dstat) until the caller needs the inertia decision.
constexpr int NBO = 256;
struct LdltWs;
inline int ldlt_blockinv_launch(LdltWs& w, cudaStream_t st, int blk0, int nblocks);   // defined with the solve kernels
// Wb0 != nullptr: COLUMN-BLOCK mode (block-column-cyclic multi-GPU driver): w.A is the top-left corner of an n x 256 block
// column (leading dimension w.ld); only the first outer panel is processed -- the four tile steps with their mini / panel /
// in-panel kernels, L in place, W = L D to Wb0 (n x 256, row index = row of the block column) -- and nothing is updated
// to the right of it.
inline int ldlt_factor_launch(LdltWs& w, cudaStream_t st, double* Wb0 = nullptr) {
    cudaStream_t sd = w.side;
    const int n = w.n, ld = w.ld;
    const size_t npad = (size_t)w.nblk * NB;
    double *ia = w.dinfo, *ib = w.dinfo + npad, *da = w.dinfo + 2 * npad, *db = w.dinfo + 3 * npad;
    ldlt_reset_kernel<<<1, 1, 0, st>>>(w.counts, w.dstat, w.ticket);
    LAUNCHED();
    int p = 0;
    bool side_used = false, upd_pending = false, u1_used = false, urg_used = false, r_prev = false, a2_pending = false;
    for (int c0 = 0; c0 < n; c0 += NBO, p++) {
        const int c1 = min(c0 + NBO, n);
        double* Wb = Wb0 ? Wb0 : ((p % 4 == 0) ? w.Wp : ((p % 4 == 1) ? w.Wp2 : ((p % 4 == 2) ? w.Wp3 : w.Wp4)));
        const TmaMat* tW = (w.use_tma && !Wb0) ? &w.tmW[p % 4] : nullptr;
        const TmaMat* tA = (w.use_tma && !Wb0) ? &w.tmA : nullptr;
        bool corner_pre = false;
        for (int k0 = c0; k0 < c1; k0 += NB) {
            const int k = k0 / NB, nb = min(NB, n - k0), k1 = k0 + nb;
            double* Akk = w.A + (size_t)k0 * ld + k0;
            double* Lk = w.LinvP + (size_t)k * NB * NB;
            ldlt_tile_kernel<<<1, TILE_THREADS, TILE_SMEM, st>>>(Akk, ld, nb, Lk, ia + k0, ib + k0, da + k0, db + k0, w.kind + k0,
                                                        nullptr, w.counts, w.dstat, w.pivot_u,
                                                        (k == w.sig_tile) ? w.sig : nullptr, w.tile_blocked);
            LAUNCHED();
            const int rows = n - k1;
            if (rows <= 0) break;
            double* B = w.A + (size_t)k1 * ld + k0;                      // rows below the tile
            double* Wt = Wb + (size_t)k1 * NBO + (k0 - c0);             // W = L * D for this tile step
            const int mcols = c1 - k1;                                   // remaining columns of this outer panel
            double* C11 = w.A + (size_t)k1 * ld + k1;
            const bool mini = w.use_mini && mcols > 0 && rows >= 2 * NB && gemm_nt_sub_uses_tiles(rows, Wt, NBO, B, ld);
            if (upd_pending) {   // the chain consumes what the update stream produced for the previous tile step
                CU(cudaStreamWaitEvent(st, w.ev_urest, 0));
                upd_pending = false;
            }
            if (!mini) {
                if (a2_pending) {    // the whole block column is read here: the deferred part of the boundary update must be in
                    CU(cudaStreamWaitEvent(st, w.ev_a2, 0));
                    a2_pending = false;
                }
                if (Wb0 && w.col_rest_event) CU(cudaStreamWaitEvent(st, w.col_rest_event, 0));
                ldlt_panel_kernel<<<cdiv(rows, NB), 128, PANEL_SMEM, st>>>(B, ld, rows, Lk, ia + k0, ib + k0, w.kind + k0, Wt, NBO, w.counts);
                LAUNCHED();
                if (mcols > 0) RET(gemm_nt_sub(st, C11, ld, rows, mcols, Wt, NBO, B, ld, NB, w.counts, 0, tW, tA));
            } else {
                CU(cudaEventRecord(w.ev_tile, st));
                ldlt_mini_kernel<<<MINI_CTAS, MINI_THREADS, MINI_SMEM, st>>>(B, ld, rows, Lk, ia + k0, ib + k0, w.kind + k0, Wt, NBO, C11, w.counts);
                LAUNCHED();
                CU(cudaEventRecord(w.ev_mini, st));
                CU(cudaStreamWaitEvent(w.upd, w.ev_tile, 0));
                // column-block mode: tile and mini steps stay inside the 256 x 256 diagonal block; everything below it
                // may still be receiving the look-ahead update of the previous panel
                if (Wb0 && w.col_rest_event && k0 == c0) CU(cudaStreamWaitEvent(w.upd, w.col_rest_event, 0));
                ldlt_panel_kernel<<<cdiv(rows - NB, NB), 128, PANEL_SMEM, w.upd>>>(B + (size_t)NB * ld, ld, rows - NB, Lk, ia + k0, ib + k0,
                                                                                   w.kind + k0, Wt + (size_t)NB * NBO, NBO, w.counts);
                LAUNCHED();
                CU(cudaStreamWaitEvent(w.upd, w.ev_mini, 0));
                RET(gemm_nt_sub(w.upd, C11, ld, rows, mcols, Wt, NBO, B, ld, NB, w.counts, /*skip00=*/1, tW, tA));
                CU(cudaEventRecord(w.ev_urest, w.upd));
                upd_pending = true;
                a2_pending = false;      // ev_urest is recorded after (a2) on the same stream: waiting for it covers (a2)
                if (w.split_a && !Wb0 && k1 == c1 - NB && c1 - c0 == NBO && n - c1 > 2 * NB) {
                    // the NEXT tile is the last of this outer panel.  The 128 x 128 corner of the next panel -- all the chain
                    // reads after the boundary -- already gets the contributions of the first three tile steps here, on the
                    // update stream, while the last tile is being factored; the boundary itself then only adds the last
                    // step's rank-64 term (4 CTAs, K = 64 instead of K = 256).
                    if (p >= 1 && u1_used) CU(cudaStreamWaitEvent(w.upd, w.ev_urg[(p - 1) & 1], 0));
                    RET(gemm_nt_sub(w.upd, w.A + (size_t)c1 * ld + c1, ld, 2 * NB, 2 * NB, Wb + (size_t)c1 * NBO, NBO,
                                    w.A + (size_t)c1 * ld + c0, ld, 3 * NB, w.counts, 0, tW, tA));
                    CU(cudaEventRecord(w.ev_corner, w.upd));
                    corner_pre = true;
                }
            }
        }
        if (upd_pending) {
            CU(cudaStreamWaitEvent(st, w.ev_urest, 0));
            upd_pending = false;
        }
        if (Wb0) return 0;
        const int rows2 = n - c1;
        if (rows2 <= 0) break;
        const int kw = c1 - c0;
        const double* Wpan = Wb + (size_t)c1 * NBO;                      // W rows c1..n of this outer panel
        const double* Lpan = w.A + (size_t)c1 * ld + c0;                 // L rows c1..n
        const int na = min(NBO, rows2);                                  // width of the next outer panel
        CU(cudaEventRecord(w.ev_panel[p & 1], st));                      // W / L of panel p are complete
        // (a) next panel's columns, chain stream: A[c1:, c1:c1+na] -= W L[c1:c1+na]^T.  The same columns were
        // read-modify-written by U1 of the previous panel, which must be complete.
        if (p >= 1 && u1_used) CU(cudaStreamWaitEvent(st, w.ev_urg[(p - 1) & 1], 0));
        const int ncrit = 2 * NB;
        if (w.split_a && rows2 > ncrit && na > NB) {
            // (a1) chain stream: only the 128 x 128 corner -- the next diagonal tile, the 64 rows below it and the diagonal
            // tile after that, i.e. everything the next tile + mini step read.  (a2) the rest of the block column goes to the
            // update stream, ahead of the next panel's own panel / in-panel kernels (stream order), so a panel boundary costs
            // the chain one 4-CTA launch instead of a 250-CTA one that queues behind the bulk update.
            if (corner_pre) {
                CU(cudaStreamWaitEvent(st, w.ev_corner, 0));
                RET(gemm_nt_sub(st, w.A + (size_t)c1 * ld + c1, ld, ncrit, ncrit, Wpan + 3 * NB, NBO, Lpan + 3 * NB, ld, NB, w.counts, 0, tW, tA));
            } else {
                RET(gemm_nt_sub(st, w.A + (size_t)c1 * ld + c1, ld, ncrit, ncrit, Wpan, NBO, Lpan, ld, kw, w.counts, 0, tW, tA));
            }
            CU(cudaStreamWaitEvent(w.upd, w.ev_panel[p & 1], 0));
            if (p >= 1 && u1_used) CU(cudaStreamWaitEvent(w.upd, w.ev_urg[(p - 1) & 1], 0));
            RET(gemm_nt_sub(w.upd, w.A + (size_t)(c1 + ncrit) * ld + c1, ld, rows2 - ncrit, na, Wpan + (size_t)ncrit * NBO, NBO, Lpan, ld,
                            kw, w.counts, 0, tW, tA));
            CU(cudaEventRecord(w.ev_a2, w.upd));
            a2_pending = true;
        } else {
            RET(gemm_nt_sub(st, w.A + (size_t)c1 * ld + c1, ld, rows2, na, Wpan, NBO, Lpan, ld, kw, w.counts, 0, tW, tA));
        }
        // (b) everything to the right of the next panel, in three pieces on two more streams (look-ahead depth 3, W in
        // four buffers): U1 = the block column of panel p+2 and U2 = that of panel p+3 on the URGENT stream, R = the
        // rest (lower tiles only, persistent SM-budgeted kernel) on the bulk stream.  Per block column q the updates are
        // ordered  R_(q-4) -> U2_(q-3) -> U1_(q-2) -> (a)_(q-1): the urgent stream is in order, U2 waits for the
        // previous panel's R, (a) waits for U1.  An R therefore has two panel periods before anything waits for it,
        // and the short urgent pieces never queue behind it.
        u1_used = false;
        const int rows3 = rows2 - na;
        if (rows3 > 0) {
            CU(cudaStreamWaitEvent(w.urg, w.ev_panel[p & 1], 0));
            const int na2 = min(NBO, rows3);
            const int o2 = c1 + na;
            // tcgen05 path: the digits of W / -L rows o2 .. n of this panel are produced ONCE (urgent stream) and shared by the
            // three pieces U1, U2 and R, which address them through 128-row block offsets
            const bool tc = w.tc_update && kw == NBO && na == NBO && (rows3 % OZ_BM) == 0 && rows3 >= 2 * NBO;
            if (tc) {
                RET(oz_panel_slice(w.urg, rows3, Wpan + (size_t)na * NBO, NBO, Lpan + (size_t)na * ld, ld, kw, w.tcu, p % OZ_UPD_NBUF,
                                   w.counts + 6, w.counts));
                CU(cudaEventRecord(w.ev_slice[p & 1], w.urg));
                RET(oz_panel_update(w.urg, w.A + (size_t)o2 * ld + o2, ld, rows3, na2, 0, kw, w.tcu, p % OZ_UPD_NBUF, w.counts + 6,
                                    w.counts, 0));
            } else {
                RET(gemm_nt_sub(w.urg, w.A + (size_t)o2 * ld + o2, ld, rows3, na2, Wpan + (size_t)na * NBO, NBO, Lpan + (size_t)na * ld, ld,
                                kw, w.counts, 0, tW, tA));
            }
            CU(cudaEventRecord(w.ev_urg[p & 1], w.urg));
            u1_used = true;
            urg_used = true;
            const int rows4 = rows3 - na2;
            bool r_now = false;
            if (rows4 > 0) {
                const int na3 = min(NBO, rows4);
                const int o3 = o2 + na2;
                if (r_prev) CU(cudaStreamWaitEvent(w.urg, w.ev_upd[(p - 1) & 1], 0));
                if (tc && na2 == NBO) {
                    RET(oz_panel_update(w.urg, w.A + (size_t)o3 * ld + o3, ld, rows4, na3, na2 / OZ_BM, kw, w.tcu, p % OZ_UPD_NBUF,
                                        w.counts + 6, w.counts, 0));
                } else {
                    RET(gemm_nt_sub(w.urg, w.A + (size_t)o3 * ld + o3, ld, rows4, na3, Wpan + (size_t)(na + na2) * NBO, NBO,
                                    Lpan + (size_t)(na + na2) * ld, ld, kw, w.counts, 0, tW, tA));
                }
                const int rows5 = rows4 - na3;
                if (rows5 > 0) {
                    const int o4 = o3 + na3;
                    CU(cudaStreamWaitEvent(sd, w.ev_panel[p & 1], 0));
                    if (tc && na2 == NBO && na3 == NBO) {
                        // the bulk of the panel-update contraction on tcgen05 (int8 error-free split, 21 slice pairs, in place)
                        CU(cudaStreamWaitEvent(sd, w.ev_slice[p & 1], 0));
                        RET(oz_panel_update(sd, w.A + (size_t)o4 * ld + o4, ld, rows5, rows5, (na2 + na3) / OZ_BM, kw, w.tcu,
                                            p % OZ_UPD_NBUF, w.counts + 6, w.counts, w.tc_ctas));
                    } else {
                        GemmArgs u{};
                        u.C = w.A + (size_t)o4 * ld + o4; u.ldc = ld; u.Cin = u.C; u.ldcin = ld; u.n = rows5; u.m = rows5;
                        u.beta = 1.0; u.mode = GEMM_LOWER_ONLY; u.nterms = 1; u.ctrl = w.counts; u.max_ctas = w.side_ctas;
                        u.t[0] = GemmTerm{Wpan + (size_t)(na + na2 + na3) * NBO, Lpan + (size_t)(na + na2 + na3) * ld, nullptr, NBO, ld, kw, -1.0};
                        RET(gemm_nt(sd, u));
                    }
                    CU(cudaEventRecord(w.ev_upd[p & 1], sd));
                    r_now = true;
                    side_used = true;
                }
            }
            r_prev = r_now;
        } else {
            r_prev = false;
        }
        if (w.solve256 && w.binv_mode == 0) {
            // explicit inverse of this panel's 256 x 256 diagonal block (block-256 solves): off the chain, behind the bulk
            // update of the same panel -- nothing waits for it before the end of the factorisation
            CU(cudaStreamWaitEvent(sd, w.ev_panel[p & 1], 0));
            RET(ldlt_blockinv_launch(w, sd, p, 1));
            side_used = true;
        }
    }
    if (w.solve256 && w.binv_mode == 0) RET(ldlt_blockinv_launch(w, st, p, 1));     // the last panel: on the chain stream itself
    // join the forked streams
    if (a2_pending) CU(cudaStreamWaitEvent(st, w.ev_a2, 0));
    if (urg_used) {
        CU(cudaEventRecord(w.ev_urg[0], w.urg));
        CU(cudaStreamWaitEvent(st, w.ev_urg[0], 0));
    }
    if (side_used) {
        CU(cudaEventRecord(w.ev_upd[0], sd));
        CU(cudaStreamWaitEvent(st, w.ev_upd[0], 0));
    }
    if (w.solve256 && w.binv_mode == 1) RET(ldlt_blockinv_launch(w, st, 0, w.nb256));
    return 0;
}
// back to fp64 DMMA trailing updates (the captured graph contains the tcgen05 launches: drop it)
inline void ldlt_disable_tc(LdltWs& w) {
    w.tc_update = 0;
    if (w.gexec) { cudaGraphExecDestroy(w.gexec); w.gexec = nullptr; }
    w.graph_state = 0;
}
// The ~230 short, mutually dependent launches of one factorisation are captured ONCE per workspace into a CUDA graph
// (both streams; fork/join through the events) and replayed: the matrix lives in the same buffers every time, so
// only the launch overhead changes.  Falls back to direct launches if capture is not possible.
inline int ldlt_factor(LdltWs& w) {
    // A different pivot threshold (the rare strict re-factorisation) is launched directly and leaves the cached graph
    // alone: re-capturing twice (there and back) costs ~20 ms, 230 direct launches less than 1 ms.
    if (w.graph_state == 1 && w.graph_u != w.pivot_u) return ldlt_factor_launch(w, w.st);
    if (w.graph_state == 1 && w.graph_sig_tile != w.sig_tile) {   // baked parameter changed: rebuild
        cudaGraphExecDestroy(w.gexec);
        w.gexec = nullptr;
        w.graph_state = 0;
    }
    if (w.graph_state == 0) {
        w.graph_state = -1;
        cudaGraph_t graph = nullptr;
        if (cudaStreamBeginCapture(w.cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
            const long long before = g_launches.load();
            const int rc = ldlt_factor_launch(w, w.cap);
            g_launches.store(before);               // captured launches are counted when the graph is replayed
            const cudaError_t e = cudaStreamEndCapture(w.cap, &graph);
            if (rc == 0 && e == cudaSuccess && graph != nullptr &&
                cudaGraphInstantiate(&w.gexec, graph, 0) == cudaSuccess) {
                w.graph_state = 1;
                w.graph_u = w.pivot_u;
                w.graph_sig_tile = w.sig_tile;
                size_t nn = 0;
                cudaGraphGetNodes(graph, nullptr, &nn);
                w.graph_nodes = (int)nn;
            }
            if (graph) cudaGraphDestroy(graph);
        }
        cudaGetLastError();   // clear any capture error; direct launches still work
    }
    if (w.graph_state == 1) {
        CU(cudaGraphLaunch(w.gexec, w.st));
        g_launches.fetch_add(w.graph_nodes, std::memory_order_relaxed);
        return 0;
    }
    return ldlt_factor_launch(w, w.st);
}

// ------------------------------------------------------------------------------------------- solve kernels
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}

// Block results are published by VALUE: yv / xv are pre-filled with a NaN sentinel and a consumer polls the entries it
// needs until they differ from it -- no flag, no fence, one L2 round trip less per step of the serial chain.
constexpr unsigned long long SOLVE_SENTINEL = 0x7FF8B200DEADBEEFull;
__global__ void ldlt_fill_sentinel_kernel(double* a, double* b, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        a[i] = __longlong_as_double((long long)SOLVE_SENTINEL);
        b[i] = __longlong_as_double((long long)SOLVE_SENTINEL);
    }
}

// A poll that times out raises bit 8 of *err (the caller must then discard the solve) instead of silently passing the
// sentinel on as data; publication / polling use relaxed gpu-scope accesses (morally strong: no tearing, no caching in L1).
__device__ __forceinline__ double ld_poll_f64_err(const double* p, int* err) {
    const long long t0 = clock64();
    unsigned long long b;
    do {
        asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];\n" : "=l"(b) : "l"(p) : "memory");
        if (b != SOLVE_SENTINEL) return __longlong_as_double((long long)b);
    } while (clock64() - t0 < 400000000LL);
    if (err) atomicOr(err, 8);
    return 0.0;
}
__device__ __forceinline__ void st_publish_f64(double* p, double v) {
    asm volatile("st.relaxed.gpu.global.f64 [%0], %1;\n" ::"l"(p), "d"(v) : "memory");
}

// forward:  y_i = LinvP_i * (b_i - sum_{j<i} L_ij y_j),  z_i = D_i^-1 y_i
__global__ void __launch_bounds__(256) ldlt_fwd_kernel(const double* __restrict__ A, int ld, int n, int nblk,
                                                       const double* __restrict__ LinvP, const double* __restrict__ dinv_a,
                                                       const double* __restrict__ dinv_b, const int* __restrict__ kind,
                                                       const double* __restrict__ b, double* yv, double* __restrict__ zv,
                                                       int* err, unsigned* ticket) {
    __shared__ double Ls[NB][NB + 4];   // LinvP_i; row stride 68 + interleaved columns: conflict-free 64-bit reads
    __shared__ double accv[NB], ys[NB], ybuf[2][NB];
    __shared__ int s_i;
    const int tid = threadIdx.x;
    if (tid == 0) s_i = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int i = s_i;
    const int r0 = i * NB;
    const int r = tid >> 2, q = tid & 3;
    const bool rowok = (r0 + r) < n;
    {   // stage LinvP_i while earlier block rows are still being solved
        const double* src = LinvP + (size_t)i * NB * NB;
        for (int idx = tid; idx < NB * NB; idx += 256) Ls[idx / NB][idx % NB] = src[idx];
    }
    double part = 0.0;
    const double* arow = A + (size_t)(r0 + r) * ld + q * 16;
    for (int j = 0; j < i; j++) {
        double lv[16];
        if (rowok) {
            const double2* p2 = reinterpret_cast<const double2*>(arow + (size_t)j * NB);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const double2 t = __ldg(p2 + c);
                lv[2 * c] = t.x;
                lv[2 * c + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 16; c++) lv[c] = 0.0;
        }
        double* yb = ybuf[j & 1];
        if (tid < NB) yb[tid] = ld_poll_f64_err(yv + (size_t)j * NB + tid, err);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; c++) part += lv[c] * yb[q * 16 + c];
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    if (q == 0) accv[r] = rowok ? (b[r0 + r] - part) : 0.0;
    __syncthreads();
    double yp = 0.0;
#pragma unroll
    for (int c = 0; c < 16; c++) yp += Ls[r][4 * c + q] * accv[4 * c + q];
    yp += __shfl_xor_sync(0xffffffffu, yp, 1);
    yp += __shfl_xor_sync(0xffffffffu, yp, 2);
    if (q == 0) {
        ys[r] = yp;
        st_publish_f64(yv + r0 + r, rowok ? yp : 0.0);     // publication (padded rows too: consumers poll whole blocks)
    }
    __syncthreads();
    if (q == 0 && rowok) {
        const int gidx = r0 + r;
        const int k = kind[gidx];
        double z = dinv_a[gidx] * ys[r];
        if (k == 1) z += dinv_b[gidx] * ys[r + 1];
        else if (k == 2) z += dinv_b[gidx - 1] * ys[r - 1];
        zv[gidx] = z;
    }
}

// backward:  x_i = LinvP_i^T * (z_i - sum_{j>i} L_ji^T x_j)
__global__ void __launch_bounds__(256) ldlt_bwd_kernel(const double* __restrict__ A, int ld, int n, int nblk,
                                                       const double* __restrict__ LinvP, const double* __restrict__ zv,
                                                       double* xv, int* err, unsigned* ticket) {
    extern __shared__ __align__(16) double bsm[];
    double(*P)[NBP] = reinterpret_cast<double(*)[NBP]>(bsm);          // partial sums [row r][col c]
    double(*Ls)[NBP] = reinterpret_cast<double(*)[NBP]>(bsm + NB * NBP);
    __shared__ double tv[NB];
    __shared__ int s_i;
    const int tid = threadIdx.x;
    if (tid == 0) s_i = nblk - 1 - (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int i = s_i;
    const int c0 = i * NB;
    const int r = tid >> 2, q = tid & 3;
    {
        const double* src = LinvP + (size_t)i * NB * NB;
        for (int idx = tid; idx < NB * NB; idx += 256) Ls[idx / NB][idx % NB] = src[idx];
    }
    double pacc[16];
#pragma unroll
    for (int c = 0; c < 16; c++) pacc[c] = 0.0;
    for (int j = nblk - 1; j > i; j--) {
        const int gr = j * NB + r;
        const bool rowok = gr < n;
        double lv[16];
        if (rowok) {
            const double2* p2 = reinterpret_cast<const double2*>(A + (size_t)gr * ld + c0 + q * 16);
#pragma unroll
            for (int c = 0; c < 8; c++) {
                const double2 t = __ldg(p2 + c);
                lv[2 * c] = t.x;
                lv[2 * c + 1] = t.y;
            }
        } else {
#pragma unroll
            for (int c = 0; c < 16; c++) lv[c] = 0.0;
        }
        const double xr = ld_poll_f64_err(xv + gr, err);   // published by the CTA of block row j (zeros in padded rows)
#pragma unroll
        for (int c = 0; c < 16; c++) pacc[c] += lv[c] * xr;
    }
#pragma unroll
    for (int c = 0; c < 16; c++) P[r][q * 16 + c] = pacc[c];
    __syncthreads();
    {   // column sums and x = LinvP^T tv with all 256 threads: column cc, four interleaved row chunks, shuffle-reduced
        const int cc = tid >> 2, qq = tid & 3;
        double s = 0.0;
#pragma unroll
        for (int k = 0; k < 16; k++) s += P[4 * k + qq][cc];
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        if (qq == 0) tv[cc] = ((c0 + cc) < n ? zv[c0 + cc] : 0.0) - s;
        __syncthreads();
        double xx = 0.0;
#pragma unroll
        for (int k = 0; k < 16; k++) xx = fma(Ls[4 * k + qq][cc], tv[4 * k + qq], xx);
        xx += __shfl_xor_sync(0xffffffffu, xx, 1);
        xx += __shfl_xor_sync(0xffffffffu, xx, 2);
        if (qq == 0) st_publish_f64(xv + c0 + cc, (c0 + cc < n) ? xx : 0.0);     // publication
    }
}


// ------------------------------------------------------------------------------------------- block-256 solves
// The 64-row chain above costs one L2 round trip per tile (72 links of ~2.5 us at config 3 = 0.07 of the HBM roofline).
// Here a link of the chain is a 256-row block I: the factorisation also emits the explicit inverse  X_I = T_I^-1  of the
// block lower-triangular matrix  T_I = [Lt_ij], Lt_ii = (LinvP_i)^-1, Lt_ij = the stored L tile (i > j)  of the four
// tiles of the block (ldlt_blockinv_kernel: X_ii = LinvP_i, X_ij = -LinvP_i sum_{j<=k<i} Lt_ik X_kj), and a solve is
//   forward   y_I = X_I   (b_I - sum_{J<I} L_IJ y_J),   z = D^-1 y
//   backward  x_I = X_I^T (z_I - sum_{J>I} L_JI^T x_J)
// run by one thread-block CLUSTER of eight CTAs per link: CTA c owns 32 rows / columns of the block, polls the 256
// published values of every earlier link once, and the eight CTAs exchange their 32 partial results through distributed
// shared memory (one cluster barrier) before each applies its rows of X_I from registers.  A link costs one L2 round
// trip + one cluster barrier instead of four L2 round trips.
constexpr int SB = 256;
constexpr int SBT = SB / NB;
static_assert(SBT == 4, "the solve clusters are written for four tiles per link");
constexpr int BINV_SLAB = 8;           // columns of X_I per CTA of the inverse builder
constexpr int BINV_NSL = NB / BINV_SLAB;   // slabs per block column
constexpr int BINV_LDA = NB + 2;       // row stride of the staged left operand: 16-byte aligned rows, conflict-free reads
constexpr int BINV_SMEM = (2 * NB * BINV_LDA + 3 * NB * BINV_SLAB + NB * BINV_SLAB) * 8;

// grid (4 * BINV_NSL, blocks): blockIdx.x = BINV_NSL * j + slab -> 8 columns of block column j of X_I, I = blk0 +
// blockIdx.y.  Columns of a triangular inverse are independent, so no CTA waits for another.  The left operands of the
// successive 64 x 64 x 8 products (Lt_ik for k = j .. i-1, then LinvP_i, for i = j+1 ..) are all INPUTS, so they are
// streamed through a two-deep cp.async ring one product ahead.  X (row-major 256 x 256 per block) and XT (its
// transpose) are zero-initialised once; entries of padded rows / columns (index >= n) are never set.
__global__ void __launch_bounds__(256) ldlt_blockinv_kernel(const double* __restrict__ A, int ld, int n, int nblk,
                                                            const double* __restrict__ LinvP, double* __restrict__ X,
                                                            double* __restrict__ XT, int blk0, const int* __restrict__ ctrl) {
    extern __shared__ __align__(16) double bsm[];
    __shared__ int s_abort;
    if (threadIdx.x == 0) s_abort = ctrl ? *reinterpret_cast<const volatile int*>(ctrl + 4) : 0;
    __syncthreads();
    if (s_abort) return;
    double* As = bsm;                               // [2][64][66]  left operand ring
    double* Xs = As + 2 * NB * BINV_LDA;            // [3][64][8]   X_kj slabs, k = j .. j + 2
    double* Ss = Xs + 3 * NB * BINV_SLAB;           // [64][8]      S = sum_k Lt_ik X_kj
    const int tid = threadIdx.x;
    const int I = blk0 + blockIdx.y, j = blockIdx.x / BINV_NSL, c0 = (blockIdx.x % BINV_NSL) * BINV_SLAB;
    const int t0 = I * SBT, ntl = min(SBT, nblk - t0);
    if (j >= ntl) return;
    double* Xo = X + (size_t)I * SB * SB;
    double* XTo = XT + (size_t)I * SB * SB;
    const int r = tid >> 2, cq = (tid & 3) * 2;          // this thread's outputs: row r, slab columns cq, cq + 1
    const int gcol0 = (t0 + j) * NB + c0;                // global index of slab column 0
    // the sequence of left operands: for i = j+1 .. ntl-1: Lt_ij, ..., Lt_i,i-1, LinvP_i
    const int nrounds = (ntl - 1 - j) * (ntl - j) / 2 + (ntl - 1 - j);
    auto round_ik = [&](int rd, int& i, int& k) {        // k == i: the LinvP_i round
        i = j + 1;
        while (rd >= i - j + 1) { rd -= i - j + 1; i++; }
        k = j + rd;
    };
    auto prefetch = [&](int rd) {
        if (rd < nrounds) {
            int i, k;
            round_ik(rd, i, k);
            const double* src = (k == i) ? LinvP + (size_t)(t0 + i) * NB * NB : A + (size_t)(t0 + i) * NB * ld + (size_t)(t0 + k) * NB;
            const int lds = (k == i) ? NB : ld;
            double* dst = As + (rd & 1) * NB * BINV_LDA;
#pragma unroll
            for (int it = 0; it < NB * NB / 2 / 256; it++) {
                const int idx = tid + 256 * it;
                const int row = idx >> 5, kc = (idx & 31) * 2;
                const bool ok = (t0 + i) * NB + row < n;   // rows of L beyond the matrix: zeros (so are padded rows of S)
                cp_async16(dst + row * BINV_LDA + kc, ok ? src + (size_t)row * lds + kc : src, ok ? 16 : 0);
            }
        }
        cp_async_commit();
    };
    auto product = [&](const double* Al, const double* Bs, double (&acc)[2]) {      // acc += Al[r][:] * Bs[:][cq, cq + 1]
        double a0 = 0.0, a1 = 0.0, b0 = 0.0, b1 = 0.0;
#pragma unroll 8
        for (int k = 0; k < NB; k += 2) {
            const double2 a = *reinterpret_cast<const double2*>(Al + r * BINV_LDA + k);
            const double2 x0 = *reinterpret_cast<const double2*>(Bs + k * BINV_SLAB + cq);
            const double2 x1 = *reinterpret_cast<const double2*>(Bs + (k + 1) * BINV_SLAB + cq);
            a0 = fma(a.x, x0.x, a0); a1 = fma(a.x, x0.y, a1);
            b0 = fma(a.y, x1.x, b0); b1 = fma(a.y, x1.y, b1);
        }
        acc[0] += a0 + b0;
        acc[1] += a1 + b1;
    };
    auto emit = [&](int i, const double (&v)[2], double* slab) {   // tile (i, j) of X, this thread's two entries
        const int gi = (t0 + i) * NB + r;
#pragma unroll
        for (int e = 0; e < 2; e++) {
            const double x = (gi < n && gcol0 + cq + e < n) ? v[e] : 0.0;
            if (slab) slab[r * BINV_SLAB + cq + e] = x;
            Xo[(size_t)(i * NB + r) * SB + j * NB + c0 + cq + e] = x;
            XTo[(size_t)(j * NB + c0 + cq + e) * SB + i * NB + r] = x;
        }
    };
    prefetch(0);
    prefetch(1);
    {   // X_jj = LinvP_j
        const double* Lj = LinvP + (size_t)(t0 + j) * NB * NB;
        double v[2];
#pragma unroll
        for (int e = 0; e < 2; e++) v[e] = Lj[r * NB + c0 + cq + e];
        emit(j, v, Xs);
    }
    int rd = 0;
    for (int i = j + 1; i < ntl; i++) {
        double sacc[2] = {0.0, 0.0};
        for (int k = j; k <= i; k++, rd++) {
            cp_async_wait<1>();                  // round rd has landed (one younger group may still be in flight)
            if (k == i) {
#pragma unroll
                for (int e = 0; e < 2; e++) Ss[r * BINV_SLAB + cq + e] = sacc[e];
            }
            __syncthreads();                     // ring slot rd & 1, the X slabs and (k == i) S are visible
            const double* Al = As + (rd & 1) * NB * BINV_LDA;
            if (k < i) {
                product(Al, Xs + (k - j) * NB * BINV_SLAB, sacc);
            } else {
                double x[2] = {0.0, 0.0};
                product(Al, Ss, x);
                x[0] = -x[0]; x[1] = -x[1];
                emit(i, x, (i - j < 3) ? Xs + (i - j) * NB * BINV_SLAB : nullptr);
            }
            __syncthreads();                     // everybody is done with ring slot rd & 1
            prefetch(rd + 2);
        }
    }
    cp_async_wait<0>();
}

// forward link: one cluster of SBC = 8 CTAs, CTA c = rows 32 c .. 32 c + 31 of block I (ticket order = scheduling order);
// 8 lanes per row, so a thread keeps 32 entries of its row of X_I and 32 entries of the current L block in registers
constexpr int SBC = 8;             // CTAs per cluster
constexpr int SBR = SB / SBC;      // rows (forward) / columns (backward) per CTA
__global__ void __cluster_dims__(SBC, 1, 1) __launch_bounds__(256)
ldlt_fwd256_kernel(const double* __restrict__ A, int ld, int n, int nblk, const double* __restrict__ X,
                   const double* __restrict__ b, double* yv, int* err, unsigned* ticket) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ __align__(16) double accv[SB];        // b_I - sum: the 32 values of CTA c' arrive from that CTA
    __shared__ __align__(16) double ybuf[2][SB];
    __shared__ int s_I;
    const int tid = threadIdx.x;
    const int cr = (int)cluster.block_rank();
    if (cr == 0 && tid == 0) s_I = (int)atomicAdd(ticket, 1u);
    cluster.sync();
    const int I = *cluster.map_shared_rank(&s_I, 0);
    const int r0 = I * SB + cr * SBR;
    const int r = tid >> 3, q = tid & 7;
    const bool rowok = (r0 + r) < n;
    const int ncol = NB * (cr / 2 + 1);              // X_I is block lower triangular in 64 x 64 tiles
    double xr[32];
    {
        const double* xrow = X + ((size_t)I * SB + cr * SBR + r) * SB;
#pragma unroll
        for (int c = 0; c < 32; c++) xr[c] = (8 * c < ncol) ? __ldg(xrow + 8 * c + q) : 0.0;
    }
    const double bval = rowok ? b[r0 + r] : 0.0;
    double part[4] = {0.0, 0.0, 0.0, 0.0};
    // lane q of a row reads the 16-byte pieces 8 c' + q of the 256 columns: the eight lanes cover 128 contiguous bytes
    const double2* arow = reinterpret_cast<const double2*>(A + (size_t)(rowok ? r0 + r : 0) * ld) + q;
    for (int J = 0; J < I; J++) {
        double2 lv[16];
#pragma unroll
        for (int c = 0; c < 16; c++) lv[c] = rowok ? __ldg(arow + (size_t)J * (SB / 2) + 8 * c) : make_double2(0.0, 0.0);
        double* yb = ybuf[J & 1];
        yb[tid] = ld_poll_f64_err(yv + (size_t)J * SB + tid, err);
        __syncthreads();
#pragma unroll
        for (int c = 0; c < 16; c++) {
            const double2 yy = *reinterpret_cast<const double2*>(yb + 16 * c + 2 * q);
            part[c & 3] = fma(lv[c].x, yy.x, part[c & 3]);
            part[c & 3] = fma(lv[c].y, yy.y, part[c & 3]);
        }
    }
    double p = (part[0] + part[1]) + (part[2] + part[3]);
    p += __shfl_xor_sync(0xffffffffu, p, 1);
    p += __shfl_xor_sync(0xffffffffu, p, 2);
    p += __shfl_xor_sync(0xffffffffu, p, 4);
    const double acc = rowok ? (bval - p) : 0.0;
    if (q >= (cr & ~1)) cluster.map_shared_rank(accv, q)[cr * SBR + r] = acc;     // lane q delivers the row's value to CTA q
    cluster.sync();
    double yp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 32; c++)
        if (8 * c < ncol) yp[c & 3] = fma(xr[c], accv[8 * c + q], yp[c & 3]);
    double y = (yp[0] + yp[1]) + (yp[2] + yp[3]);
    y += __shfl_xor_sync(0xffffffffu, y, 1);
    y += __shfl_xor_sync(0xffffffffu, y, 2);
    y += __shfl_xor_sync(0xffffffffu, y, 4);
    if (q == 0) st_publish_f64(yv + r0 + r, rowok ? y : 0.0);     // publication (padded rows too)
}

// backward link: CTA c owns columns 32 c .. 32 c + 31 of block I; z = D^-1 y is formed here from the finished y
__global__ void __cluster_dims__(SBC, 1, 1) __launch_bounds__(256)
ldlt_bwd256_kernel(const double* __restrict__ A, int ld, int n, int nblk, int nb256, const double* __restrict__ XT,
                   const double* __restrict__ dinv_a, const double* __restrict__ dinv_b, const int* __restrict__ kind,
                   const double* __restrict__ yv, double* xv, int* err, unsigned* ticket) {
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ double P[SBR][SBR + 1];               // partial sums [row r][column]
    __shared__ __align__(16) double tvv[SB];         // z_I - sum: the 32 values of CTA c' arrive from that CTA
    __shared__ __align__(16) double xbuf[2][SB];
    __shared__ int s_I;
    const int tid = threadIdx.x;
    const int cr = (int)cluster.block_rank();
    if (cr == 0 && tid == 0) s_I = nb256 - 1 - (int)atomicAdd(ticket, 1u);
    cluster.sync();
    const int I = *cluster.map_shared_rank(&s_I, 0);
    const int c0 = I * SB + cr * SBR;
    const int r = tid >> 3, q = tid & 7;
    const int kmin = NB * (cr / 2);                  // X_I^T is block UPPER triangular: row 32 cr + cc needs tv[kmin ..]
    double xt[32];
    {
        const double* xrow = XT + ((size_t)I * SB + cr * SBR + r) * SB;
#pragma unroll
        for (int c = 0; c < 32; c++) xt[c] = (8 * c >= kmin) ? __ldg(xrow + 8 * c + q) : 0.0;
    }
    double zval = 0.0;                               // z of column c0 + r (r doubles as the column index cc below)
    if (c0 + r < n) {
        const int g = c0 + r;
        const int k = kind[g];
        zval = dinv_a[g] * yv[g];
        if (k == 1) zval += dinv_b[g] * yv[g + 1];
        else if (k == 2) zval += dinv_b[g - 1] * yv[g - 1];
    }
    double pacc[4] = {0.0, 0.0, 0.0, 0.0};
    const bool colsok = c0 < n;
    for (int J = nb256 - 1; J > I; J--) {
        double2 lv[16];                              // rows J*256 + 32 u + r (u = 0..7), columns c0 + 4 q .. + 3
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const int gr = J * SB + u * SBR + r;
            const bool ok = colsok && gr < n;
            const double2* p2 = reinterpret_cast<const double2*>(A + (size_t)(ok ? gr : 0) * ld + c0 + 4 * q);
            lv[2 * u] = ok ? __ldg(p2) : make_double2(0.0, 0.0);
            lv[2 * u + 1] = ok ? __ldg(p2 + 1) : make_double2(0.0, 0.0);
        }
        double* xb = xbuf[J & 1];
        xb[tid] = ld_poll_f64_err(xv + (size_t)J * SB + tid, err);     // zeros in padded rows
        __syncthreads();
#pragma unroll
        for (int u = 0; u < 8; u++) {
            const double xr = xb[u * SBR + r];
            pacc[0] = fma(lv[2 * u].x, xr, pacc[0]);
            pacc[1] = fma(lv[2 * u].y, xr, pacc[1]);
            pacc[2] = fma(lv[2 * u + 1].x, xr, pacc[2]);
            pacc[3] = fma(lv[2 * u + 1].y, xr, pacc[3]);
        }
    }
#pragma unroll
    for (int e = 0; e < 4; e++) P[r][4 * q + e] = pacc[e];
    __syncthreads();
    const int cc = r, qq = q;
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < 4; k++) s += P[8 * k + qq][cc];
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    const bool colok = (c0 + cc) < n;
    const double tv = colok ? (zval - s) : 0.0;
    if (qq <= (cr | 1)) cluster.map_shared_rank(tvv, qq)[cr * SBR + cc] = tv;     // lane qq delivers to CTA qq
    cluster.sync();
    double xp[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
    for (int c = 0; c < 32; c++)
        if (8 * c >= kmin) xp[c & 3] = fma(xt[c], tvv[8 * c + qq], xp[c & 3]);
    double xx = (xp[0] + xp[1]) + (xp[2] + xp[3]);
    xx += __shfl_xor_sync(0xffffffffu, xx, 1);
    xx += __shfl_xor_sync(0xffffffffu, xx, 2);
    xx += __shfl_xor_sync(0xffffffffu, xx, 4);
    if (qq == 0) st_publish_f64(xv + c0 + cc, colok ? xx : 0.0);     // publication
}

inline int ldlt_init_solve_attrs() {
    CU(cudaFuncSetAttribute(ldlt_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TILE_SMEM));
    CU(cudaFuncSetAttribute(ldlt_blockinv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BINV_SMEM));
    return 0;
}
inline int ldlt_blockinv_launch(LdltWs& w, cudaStream_t st, int blk0, int nblocks) {
    ldlt_blockinv_kernel<<<dim3(4 * BINV_NSL, nblocks), 256, BINV_SMEM, st>>>(w.A, w.ld, w.n, w.nblk, w.LinvP, w.Xinv, w.XinvT, blk0, w.counts);
    LAUNCHED();
    return 0;
}

// dst <- the factorisation held by src (same order): factor, per-tile inverses, D blocks, kinds, inertia counters.
// Lets the engine adopt a factorisation computed in another workspace without touching streams or captured graphs.
inline int ldlt_copy_factor(LdltWs& dst, const LdltWs& src, cudaStream_t st) {
    if (dst.n != src.n || dst.ld != src.ld) return fail_msg("ldlt_copy_factor: workspaces differ");
    const size_t npad = (size_t)src.nblk * NB;
    CU(cudaMemcpyAsync(dst.A, src.A, sizeof(double) * npad * src.ld, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(dst.LinvP, src.LinvP, sizeof(double) * (size_t)src.nblk * NB * NB, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(dst.Xinv, src.Xinv, sizeof(double) * (size_t)src.nb256 * SB * SB, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(dst.XinvT, src.XinvT, sizeof(double) * (size_t)src.nb256 * SB * SB, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(dst.dinfo, src.dinfo, sizeof(double) * 4 * npad, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(dst.kind, src.kind, sizeof(int) * npad, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(dst.counts, src.counts, sizeof(int) * 4, cudaMemcpyDeviceToDevice, st));
    CU(cudaMemcpyAsync(dst.dstat, src.dstat, sizeof(double) * 2, cudaMemcpyDeviceToDevice, st));
    return 0;
}

// Work vectors of one triangular solve; a second set lets two solves with the SAME factorisation run concurrently on
// two streams (the factor is only read).
struct LdltSolveBuf { double *yv = nullptr, *zv = nullptr, *xv = nullptr; unsigned* ticket = nullptr; };
inline int ldlt_solvebuf_alloc(LdltSolveBuf& sb, int nblk) {
    const size_t npad = (size_t)cdiv(nblk, 4) * 256;
    CU(cudaMalloc(&sb.yv, sizeof(double) * npad));
    CU(cudaMalloc(&sb.zv, sizeof(double) * npad));
    CU(cudaMalloc(&sb.xv, sizeof(double) * npad));
    CU(cudaMalloc(&sb.ticket, sizeof(unsigned) * 2));
    return 0;
}
inline void ldlt_solvebuf_free(LdltSolveBuf& sb) {
    cudaFree(sb.yv); cudaFree(sb.zv); cudaFree(sb.xv); cudaFree(sb.ticket);
    sb = LdltSolveBuf();
}
// x = A^-1 b using the factorisation in w; b and x are device vectors of length n (may alias)
inline int ldlt_solve_on(LdltWs& w, cudaStream_t st, const LdltSolveBuf& sb, const double* b, double* x) {
    const size_t npad = (size_t)w.nblk * NB;
    double *ia = w.dinfo, *ib = w.dinfo + npad;
    CU(cudaMemsetAsync(sb.ticket, 0, sizeof(unsigned) * 2, st));
    if (w.solve256) {
        const int np256 = w.nb256 * SB;
        ldlt_fill_sentinel_kernel<<<cdiv(np256, 256), 256, 0, st>>>(sb.yv, sb.xv, np256);
        LAUNCHED();
        ldlt_fwd256_kernel<<<w.nb256 * SBC, 256, 0, st>>>(w.A, w.ld, w.n, w.nblk, w.Xinv, b, sb.yv, w.serr, sb.ticket);
        LAUNCHED();
        ldlt_bwd256_kernel<<<w.nb256 * SBC, 256, 0, st>>>(w.A, w.ld, w.n, w.nblk, w.nb256, w.XinvT, ia, ib, w.kind, sb.yv, sb.xv,
                                                          w.serr, sb.ticket + 1);
        LAUNCHED();
        CU(cudaMemcpyAsync(x, sb.xv, sizeof(double) * w.n, cudaMemcpyDeviceToDevice, st));
        return 0;
    }
    ldlt_fill_sentinel_kernel<<<cdiv((int)npad, 256), 256, 0, st>>>(sb.yv, sb.xv, (int)npad);
    LAUNCHED();
    ldlt_fwd_kernel<<<w.nblk, 256, 0, st>>>(w.A, w.ld, w.n, w.nblk, w.LinvP, ia, ib, w.kind, b, sb.yv, sb.zv, w.serr, sb.ticket);
    LAUNCHED();
    ldlt_bwd_kernel<<<w.nblk, 256, TILE_SMEM, st>>>(w.A, w.ld, w.n, w.nblk, w.LinvP, sb.zv, sb.xv, w.serr, sb.ticket + 1);
    LAUNCHED();
    CU(cudaMemcpyAsync(x, sb.xv, sizeof(double) * w.n, cudaMemcpyDeviceToDevice, st));
    return 0;
}
inline int ldlt_solve(LdltWs& w, const double* b, double* x) {
    LdltSolveBuf sb;
    sb.yv = w.yv; sb.zv = w.zv; sb.xv = w.xv; sb.ticket = w.ticket;
    return ldlt_solve_on(w, w.st, sb, b, x);
}

}  // namespace b200
