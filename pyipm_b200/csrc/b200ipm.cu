// b200ipm.cu -- engine + C ABI (include/b200ipm.h).  One translation unit; kernels live in the .cuh files.
//
// One inner iteration of the primal-dual interior-point method (pyipm.py:1714-1754) on the device:
//   evaluate f/ce/ci derivatives (lowered problem forms) -> residual g and KKT norms (one pass over J)
//   -> Lagrangian Hessian W (DMMA SYRK) -> condensed KKT matrix  [W + delta I + dci S dci', dce; dce', -reg I]
//   -> tile-pivoted LDL^T with inertia, delta/reg retry loop of reghess (pyipm.py:1373-1406)
//   -> solve + iterative refinement against the UNREDUCED KKT system -> nu rule, fraction-to-the-boundary,
//   Armijo backtracking with speculative batched trials (+ second-order correction) -> state update.
// Host code only sequences kernels and takes the scalar branch decisions the reference takes in Python.
#include "../../include/b200ipm.h"
#include "common.cuh"
#include "vec.cuh"
#include "gemm_nt.cuh"
#include "ozaki_i8.cuh"
#include "ldlt.cuh"
#include "engine_kernels.cuh"
#include "batch.cuh"

#include <vector>
#include <algorithm>

namespace b200 {
thread_local std::string g_last_error;
std::atomic<long long> g_launches{0};
}  // namespace b200
using namespace b200;

enum { KIND_NONE = 0, KIND_QUAD = 1, KIND_POLY = 2, KIND_CALLABLE = 3 };
enum { EV_START = 0, EV_EVAL, EV_ASSEMBLE, EV_FACTOR, EV_SOLVE, EV_SEARCH, EV_HESS0, EV_HESS1, EV_COND0, EV_COND1, EV_N };

struct b200ipm_engine {
    int D = 0, M = 0, N = 0, K = 0, Kc = 0, C = 0, ldJ = 0, ldW = 0;
    int device = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    b200ipm_params p{};
    double mu = 0, nu = 0, delta = 0, mu_host = 0;
    int kind = KIND_NONE;
    // state
    double *x = nullptr, *s = nullptr, *lam = nullptr;
    // derivatives at the state
    double *fval = nullptr, *df = nullptr, *ce = nullptr, *ci = nullptr, *J = nullptr, *W = nullptr;
    bool eval_valid = false, hess_valid = false, resid_valid = false;
    // residual / direction
    double *g = nullptr, *sigma = nullptr, *bvec = nullptr, *tvec = nullptr, *rhs = nullptr, *sol = nullptr;
    double *ycur = nullptr, *ycor = nullptr, *rho = nullptr, *dz = nullptr, *wx = nullptr, *jt = nullptr;
    double *Hb = nullptr;   // W + dci S dci'  (D x ldW), without delta
    double reg_cur = 0, delta_eff = 0;
    bool strict_retry = false;
    int n_strict = 0;        // number of strict re-factorisations triggered by a poor residual
    int n_fast_ok = 0, n_fast_miss = 0;   // one-sync steps taken / abandoned for the sequential path
    int* h_fi = nullptr;     // pinned: control block + error word of the one-sync step
    int n_phys = 0;          // factorisations physically executed in the current step (info->n_factor_phys)
    LdltWs F;               // condensed KKT factorisation (order Kc)
    OzWs oz;                // tcgen05 int8 slices (B200IPM_FLAG_TCGEN05_SYRK)
    OzWs oz_soc;            // same for the (M+N)-order normal equations of the second-order correction (lazy)
    bool oz_used = false;   // the current W / Hb came from the tcgen05 path
    bool oz_off = false;    // this step fell back to DMMA
    LdltWs Fb;              // speculative second attempt of reghess (lazy), factored concurrently on stB
    bool Fb_ready = false;
    cudaStream_t stB = nullptr;
    cudaEvent_t ev_fork = nullptr;
    // negative-curvature certificate (replaces the background delta = 0 factorisation in the common case)
    LdltSolveBuf csb;        // second set of solve vectors: the certificate solve runs beside the main solve
    double *c_rhs = nullptr, *c_sol = nullptr, *c_hv = nullptr, *c_u = nullptr, *c_red = nullptr;
    double* h_cert = nullptr;   // pinned: 2 x (v'Hv, v'v, Hv'Hv, v'u)
    bool cert_ready = false, cert_pending = false;
    bool first_failed_last = true;   // did the delta = 0 test of the previous step fail?  (speculate only then)
    int n_cert_ok = 0, n_cert_miss = 0;
    int* d_sig = nullptr;    // marker word: foreground factorisation -> delayed start of the background one
    bool pendingA = false;   // the background delta = 0 test has not been collected yet
    double pend_delta_in = 0, pend_rcondB = 0;
    int *h_cntB = nullptr;   // pinned: control block (8 ints) of the background attempt
    double *h_dsB = nullptr; // pinned: dstat (2 doubles) of the speculative attempt
    LdltWs F2;              // pseudo-inverse / second-order-correction systems (lazy)
    bool F2_ready = false;
    int F2_n = 0;
    LdltWs F2alt;           // the previously used order of F2: init_lambda (order min(D, M+N)) and the second-order correction
    bool F2alt_ready = false;   // (order M+N) alternate within one solve, and a workspace costs ~100 ms to create
    int F2alt_n = 0;
    double *Jt = nullptr;   // (M+N) x D transposed Jacobian for the SOC normal equations (lazy)
    // scratch
    double *scr = nullptr, *part = nullptr, *red = nullptr, *trial = nullptr, *xt = nullptr, *st_ = nullptr;
    double *pvec = nullptr, *cnew = nullptr, *uvec = nullptr;
    double *h_red = nullptr;   // pinned
    int max_batch = 512;
    // QUAD
    double *Q = nullptr, *qc = nullptr, *At = nullptr, *Ut = nullptr, *qb = nullptr, *Gt = nullptr, *Vt = nullptr,
           *qr = nullptr, *xdiag = nullptr;
    double q4 = 0;
    double *qx = nullptr, *ax = nullptr, *ux = nullptr, *gx = nullptr, *vx = nullptr;
    double *qd = nullptr, *ad = nullptr, *ud = nullptr, *gd = nullptr, *vd = nullptr;
    // POLY
    int *p_rowptr = nullptr, *p_ptr = nullptr, *p_fvar = nullptr, *p_fpow = nullptr;
    double* p_coeff = nullptr;
    PolyData poly{};
    cudaEvent_t ev[EV_N];
    double last_red[8];     // host copy of the residual reductions at the current state
    // L-BFGS mode (pyipm.py:993-1371): storage S, Y on the device (row k = pair k, oldest first), the small m x m arrays
    // SS, L, D of the compact representation on the host
    int lb_max = 0, lb_m = 0, lb_fail = 0, lb_eq_reg = 0;
    double lb_zeta = 1.0, lb_zeta0 = 1.0, lb_rcond = 0.0;
    double *lb_S = nullptr, *lb_Y = nullptr, *lb_W = nullptr, *lb_xold = nullptr, *lb_gold = nullptr, *lb_dx = nullptr,
           *lb_dg = nullptr, *lb_X00 = nullptr, *lb_X01 = nullptr, *lb_zg = nullptr, *lb_t = nullptr, *lb_q = nullptr,
           *lb_dots = nullptr, *lb_coef = nullptr;
    std::vector<double> lb_SS, lb_L, lb_D;
    // device-resident snapshot
    double *sv_x = nullptr, *sv_s = nullptr, *sv_lam = nullptr;
    double sv_mu = 0, sv_nu = 0, sv_delta = 0, sv_mu_host = 0;
    bool sv_valid = false;
};
typedef b200ipm_engine Eng;

struct b200ipm_ldlt {
    LdltWs F;
    double *A0 = nullptr, *b = nullptr, *x = nullptr, *r = nullptr, *c = nullptr;
    int device = 0;
    cudaStream_t st = nullptr;
    bool own_stream = false;
    bool factored = false;
    bool tile_counts_live = false;
    OzUpdWs ozd;               // digits of the current panel for the tcgen05 trailing updates of the block-column-cyclic driver
    int* ozd_err = nullptr;    // their device error word
};

template <typename T>
static int dalloc(T** p, size_t n) {
    CU(cudaMalloc(p, sizeof(T) * std::max<size_t>(n, 1)));
    return 0;
}
static int up(Eng* h, double* dst, const double* src, size_t n, int on_device) {
    if (n == 0 || src == nullptr) return 0;
    CU(cudaMemcpyAsync(dst, src, sizeof(double) * n, on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->st));
    return 0;
}
static int down(Eng* h, double* dst, const double* src, size_t n) {
    if (n == 0 || dst == nullptr) return 0;
    CU(cudaMemcpyAsync(dst, src, sizeof(double) * n, cudaMemcpyDeviceToHost, h->st));
    return 0;
}
static int fetch_red(Eng* h, const double* dsrc, int n) {
    CU(cudaMemcpyAsync(h->h_red, dsrc, sizeof(double) * n, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

// ------------------------------------------------------------------------------------------ evaluation
static QuadData quad_data(Eng* h) {
    return QuadData{h->Q, h->qc, h->At, h->Ut, h->qb, h->Gt, h->Vt, h->qr, h->q4};
}
// linear images of a vector v (length D):  Q v, A v, U v, G v, V v
static int quad_images(Eng* h, const double* v, double* qv, double* av, double* uv, double* gv, double* vv) {
    const int D = h->D, M = h->M, N = h->N;
    RET(gemv_n(h->st, h->Q, D, D, D, v, nullptr, 0.0, 1.0, qv));
    if (M) {
        RET(gemv_t(h->st, h->At, M, D, M, v, nullptr, 0.0, 1.0, av, h->scr));
        if (h->Ut) RET(gemv_t(h->st, h->Ut, M, D, M, v, nullptr, 0.0, 1.0, uv, h->scr));
    }
    if (N) {
        RET(gemv_t(h->st, h->Gt, N, D, N, v, nullptr, 0.0, 1.0, gv, h->scr));
        if (h->Vt) RET(gemv_t(h->st, h->Vt, N, D, N, v, nullptr, 0.0, 1.0, vv, h->scr));
    }
    return 0;
}

// f, df, ce, ci, J at the current x (first-order quantities: all the KKT conditions need)
static int eval_derivs(Eng* h) {
    if (h->eval_valid) return 0;
    const int D = h->D, M = h->M, N = h->N;
    if (h->kind == KIND_QUAD) {
        RET(quad_images(h, h->x, h->qx, h->ax, h->ux, h->gx, h->vx));
        quad_point_kernel<<<1, 1024, 0, h->st>>>(D, M, N, quad_data(h), h->x, h->qx, h->ax, h->ux, h->gx, h->vx, h->df,
                                                 h->xdiag, h->ce, h->ci, h->fval);
        LAUNCHED();
        if (h->C) {
            quad_jac_kernel<<<std::min(cdiv(D * h->C, 256), 148 * 16), 256, 0, h->st>>>(D, M, N, quad_data(h), h->ux, h->vx,
                                                                                         h->J, h->ldJ);
            LAUNCHED();
        }
    } else if (h->kind == KIND_POLY) {
        poly_eval_kernel<<<1, 256, 0, h->st>>>(D, M, N, h->poly, h->x, h->fval, h->df, h->ce, h->ci, h->J, h->ldJ);
        LAUNCHED();
    } else if (h->kind == KIND_CALLABLE) {
        return fail_msg("callable mode: b200ipm_set_derivs must be called after every state change");
    } else {
        return fail_msg("no problem bound (b200ipm_bind_quad / b200ipm_bind_poly / b200ipm_set_derivs)");
    }
    h->eval_valid = true;
    h->hess_valid = false;
    h->resid_valid = false;
    return 0;
}
static bool use_tc(Eng* h) {
    return (h->p.flags & B200IPM_FLAG_TCGEN05_SYRK) && !h->oz_off && h->D >= 256 &&
           (int)(rup((size_t)h->M, OZ_KB) + rup((size_t)h->N, OZ_KB)) <= OZ_KMAX;   // exact int32 accumulation
}
static void oz_configure(Eng* h) { h->oz.variant = (h->p.flags >> 2) & 3; if (h->oz.variant == 3) h->oz.variant = 1; }
// W = d2L at the current (x, lda): only needed when a search direction is computed (a3)
static int eval_hessian(Eng* h) {
    oz_configure(h);
    RET(eval_derivs(h));
    if (h->hess_valid) return 0;
    const int D = h->D, M = h->M, N = h->N;
    CU(cudaEventRecord(h->ev[EV_HESS0], h->st));
    if (h->kind == KIND_QUAD) {
        // W = Q + 3 q4 diag(x^2) - Ut diag(lda_e) Ut' + Vt diag(lda_i) Vt'      (DMMA SYRK)
        GemmArgs a{};
        a.C = h->W; a.ldc = h->ldW; a.Cin = h->Q; a.ldcin = D; a.dadd = h->xdiag; a.n = D; a.m = D; a.beta = 1.0;
        a.shift = 0.0; a.mode = GEMM_UPPER_MIRROR; a.nterms = 0;
        if (M && h->Ut) a.t[a.nterms++] = GemmTerm{h->Ut, h->Ut, h->lam, M, M, M, -1.0};
        if (N && h->Vt) a.t[a.nterms++] = GemmTerm{h->Vt, h->Vt, h->lam + M, N, N, N, 1.0};
        if (use_tc(h) && a.nterms) {
            // d2L enters the refinement residual: 28 pairs = 1.4e-15 max|C| (the DMMA kernel: 6.2e-15); FULLCOND: all 34
            h->oz.ndiag = (h->p.flags & B200IPM_FLAG_TCGEN05_FULLCOND) ? 8 : 7;
            RET(oz_syrk(h->st, a, h->oz, (M && h->Ut) ? 1u : 0u));   // lda_e has either sign, lda_i >= 0
            h->oz_used = true;
        } else {
            RET(gemm_nt(h->st, a));
        }
    } else if (h->kind == KIND_POLY) {
        poly_hess_kernel<<<cdiv(D * D, 256), 256, 0, h->st>>>(D, M, N, h->poly, h->x, h->lam, h->W, h->ldW);
        LAUNCHED();
    }   // KIND_CALLABLE: W was uploaded by set_derivs
    CU(cudaEventRecord(h->ev[EV_HESS1], h->st));
    h->hess_valid = true;
    return 0;
}

// g, sigma, KKT norms, ||con||_1 at the current state  (a1, a10)
static int residual(Eng* h) {
    RET(eval_derivs(h));
    if (h->resid_valid) return 0;
    const int D = h->D, M = h->M, N = h->N;
    int npart = 0;
    if (h->C) {
        // g_x = df - J * lda   (the headline HBM-bound kernel: one pass over J), squared norm fused
        RET(gemv_n(h->st, h->J, h->ldJ, D, h->C, h->lam, h->df, 1.0, -1.0, h->g, h->part));
        npart = gemv_n_blocks(D);
    } else {
        npart = cdiv(D, 256);
        copy_sq_kernel<<<npart, 256, 0, h->st>>>(D, h->df, h->g, h->part);   // unconstrained: g_x = df
        LAUNCHED();
    }
    residual_tail_kernel<<<1, 1024, 0, h->st>>>(D, M, N, h->s, h->lam, h->ce, h->ci, h->mu, h->p.eps, h->g, h->sigma,
                                                h->part, npart, h->fval, h->red);
    LAUNCHED();
    RET(fetch_red(h, h->red, 6));
    for (int i = 0; i < 6; i++) h->last_red[i] = h->h_red[i];
    h->resid_valid = true;
    return 0;
}

// ------------------------------------------------------------------------------------------ KKT formation
// Hb = W + dci diag(sigma) dci'      (condensation, a3/a4; DMMA SYRK)
static int condense(Eng* h) {
    const int D = h->D, M = h->M, N = h->N;
    GemmArgs a{};
    a.C = h->Hb; a.ldc = h->ldW; a.Cin = h->W; a.ldcin = h->ldW; a.dadd = nullptr; a.n = D; a.m = D; a.beta = 1.0;
    a.shift = 0.0; a.mode = GEMM_UPPER_MIRROR; a.nterms = 0;
    if (N) a.t[a.nterms++] = GemmTerm{h->J + M, h->J + M, h->sigma, h->ldJ, h->ldJ, N, 1.0};
    CU(cudaEventRecord(h->ev[EV_COND0], h->st));
    if (use_tc(h) && a.nterms) {
        // Hb is only factored (inertia + preconditioner of the refinement against the unreduced system, which uses W, J
        // and Sigma themselves): B200IPM_FLAG_TCGEN05_FULLCOND keeps all 34 slice pairs, the default keeps 21 (~1e-12)
        h->oz.ndiag = (h->p.flags & B200IPM_FLAG_TCGEN05_FULLCOND) ? 8 : 6;
        RET(oz_syrk(h->st, a, h->oz, 0u));                           // sigma = lda_i / (s + eps) >= 0 in the interior
        h->oz_used = true;
    } else {
        RET(gemm_nt(h->st, a));
    }
    CU(cudaEventRecord(h->ev[EV_COND1], h->st));
    return 0;
}
// Kc = [[Hb + delta I, .], [dce', -reg I]]  (lower triangle is what the factorisation reads)
static int build_kc_into(Eng* h, LdltWs& F, cudaStream_t st, double delta, double reg) {
    const int D = h->D, M = h->M;
    kc_xx_kernel<<<std::min(cdiv(D * D, 256), 148 * 32), 256, 0, st>>>(h->Hb, h->ldW, D, delta, F.A, F.ld);
    LAUNCHED();
    if (M) {
        RET(transpose(st, h->J, h->ldJ, D, M, F.A + (size_t)D * F.ld, F.ld));
        kc_ee_kernel<<<cdiv(M * M, 256), 256, 0, st>>>(F.A, F.ld, D, M, reg);
        LAUNCHED();
    }
    return 0;
}
static int build_kc(Eng* h, double delta, double reg) {
    RET(build_kc_into(h, h->F, h->st, delta, reg));
    h->delta_eff = delta;
    h->reg_cur = reg;
    return 0;
}
// Speculation (reghess, pyipm.py:1373-1406): once a shift was needed, the delta = 0 test almost always fails again
// and the next candidate max(delta/2, delta0) is known in advance.  The delta = 0 test ("A") is therefore started in
// the BACKGROUND -- own workspace, low-priority streams -- while the candidate ("B") is factored in the foreground
// and the solve proceeds with it; A's verdict is collected afterwards (resolve_pending).  In the rare cases where A
// passes, or asks for the eq-block regularisation (rcond <= eps), the step is redone with the sequential loop, so
// the decisions are always exactly the reference's.
static int spec_launch_background(Eng* h, int neg_limit) {
    if (!h->Fb_ready) {
        int prio_lo = 0, prio_hi = 0;
        CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        if (!h->stB) {
            CU(cudaStreamCreateWithPriority(&h->stB, cudaStreamNonBlocking, prio_lo));
            CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        }
        CU(cudaMallocHost(&h->h_cntB, sizeof(int) * 8));
        CU(cudaMallocHost(&h->h_dsB, sizeof(double) * 2));
        RET(ldlt_alloc(h->Fb, h->Kc, h->stB, /*background=*/true));
        h->Fb.pivot_u = h->F.pivot_u;
        h->Fb_ready = true;
    }
    // optional: the background test starts when the foreground factorisation (launched next, on the main stream)
    // reaches tile step nblk/3 (~70 % of its trailing-update work done).  Measured at config 3: the foreground gets
    // faster (4.8 -> 4.2 ms) but the verdict arrives later by more than that, so it is off unless asked for.
    const bool delay = h->F.nblk >= 24 && (h->p.flags & B200IPM_FLAG_DELAY_BG);
    if (!h->d_sig) { RET(dalloc(&h->d_sig, 1)); }
    h->F.sig = h->d_sig;
    h->F.sig_tile = delay ? h->F.nblk / 3 : -1;
    CU(cudaMemsetAsync(h->d_sig, 0, sizeof(int), h->st));
    CU(cudaEventRecord(h->ev_fork, h->st));          // Hb and J are complete on the main stream
    CU(cudaStreamWaitEvent(h->stB, h->ev_fork, 0));
    RET(ldlt_set_neg_limit(h->Fb, neg_limit));
    RET(build_kc_into(h, h->Fb, h->stB, 0.0, 0.0));
    h->n_phys++;
    if (delay) {
        ldlt_wait_sig_kernel<<<1, 1, 0, h->stB>>>(h->d_sig);
        LAUNCHED();
    }
    RET(ldlt_factor(h->Fb));
    CU(cudaMemcpyAsync(h->h_cntB, h->Fb.counts, sizeof(int) * 8, cudaMemcpyDeviceToHost, h->stB));
    CU(cudaMemcpyAsync(h->h_dsB, h->Fb.dstat, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->stB));
    return 0;
}
// Certificate that the delta = 0 inertia test FAILS, without factoring the unshifted matrix.
// The condensed matrix [[Hb, dce], [dce', 0]] has inertia (D, M, 0) iff dce has full column rank and Hb is positive
// definite on null(dce').  With the accepted factorisation of B = [[Hb + delta1 I, dce], [dce', 0]] at hand, one
// solve  B [v; y] = [u; 0]  gives a vector v in null(dce') rich in the lowest modes of the reduced Hessian (inverse
// iteration, warm-started from the previous Newton step); if  v' Hb v < 0  by a safe margin the reduced Hessian is
// not positive definite, so the reference's first test (pyipm.py:1381) fails -- a proof, not a heuristic.  Two
// iterations run on the background stream beside the main solve.  No certificate (e.g. the problem has become
// convex) => the caller falls back to the plain sequential loop.  The rcond <= eps branch of the failed test cannot
// be seen this way; it needs a numerically singular unshifted matrix while the shifted one is accepted with full-rank
// dce, and is treated as not taken.
static int cert_launch(Eng* h) {
    const int D = h->D, Kc = h->Kc;
    if (!h->cert_ready) {
        if (!h->stB) {
            int prio_lo = 0, prio_hi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
            CU(cudaStreamCreateWithPriority(&h->stB, cudaStreamNonBlocking, prio_lo));
            CU(cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
        }
        RET(ldlt_solvebuf_alloc(h->csb, h->F.nblk));
        RET(dalloc(&h->c_rhs, Kc)); RET(dalloc(&h->c_sol, Kc)); RET(dalloc(&h->c_hv, D)); RET(dalloc(&h->c_u, D));
        RET(dalloc(&h->c_red, 8));
        CU(cudaMemset(h->c_red, 0, sizeof(double) * 8));
        CU(cudaMallocHost(&h->h_cert, sizeof(double) * 8));
        std::vector<double> u0(D);
        unsigned long long sd = 0x9E3779B97F4A7C15ull;
        for (int i = 0; i < D; i++) {   // fixed pseudo-random start vector (splitmix64)
            sd += 0x9E3779B97F4A7C15ull;
            unsigned long long z = sd;
            z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
            z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
            z ^= z >> 31;
            u0[i] = (double)(z >> 11) * (1.0 / 9007199254740992.0) - 0.5;
        }
        CU(cudaMemcpy(h->c_u, u0.data(), sizeof(double) * D, cudaMemcpyHostToDevice));
        h->cert_ready = true;
    }
    CU(cudaEventRecord(h->ev_fork, h->st));          // the factorisation and Hb are complete on the main stream
    CU(cudaStreamWaitEvent(h->stB, h->ev_fork, 0));
    for (int it = 0; it < 2; it++) {
        cert_rhs_kernel<<<cdiv(Kc, 256), 256, 0, h->stB>>>(D, Kc, h->c_u, h->c_rhs);
        LAUNCHED();
        RET(ldlt_solve_on(h->F, h->stB, h->csb, h->c_rhs, h->c_sol));
        RET(gemv_n(h->stB, h->Hb, h->ldW, D, D, h->c_sol, nullptr, 0.0, 1.0, h->c_hv));
        cert_stats_kernel<<<1, 1024, 0, h->stB>>>(D, h->c_sol, h->c_hv, h->c_u, h->c_red + 4 * it);
        LAUNCHED();
    }
    CU(cudaMemcpyAsync(h->h_cert, h->c_red, sizeof(double) * 8, cudaMemcpyDeviceToHost, h->stB));
    return 0;
}
// One inertia test = one factorisation.  `neg_limit`: the test can only pass with exactly M negative pivots, so a
// factorisation that has already produced more is abandoned on the device (ldlt.cuh control block) -- *abandoned then
// reports it, and n_neg / rcond describe the part that was factored.
static int factor_once(Eng* h, double delta, double reg, int neg_limit, int* n_neg, int* n_zero, double* rcond,
                       int* abandoned) {
    RET(ldlt_set_neg_limit(h->F, neg_limit));
    int cnt[8];
    double ds[2];
    for (int attempt = 0; attempt < 2; attempt++) {
        RET(build_kc(h, delta, reg));
        RET(ldlt_factor(h->F));
        h->n_phys++;
        CU(cudaMemcpyAsync(cnt, h->F.counts, sizeof(int) * 8, cudaMemcpyDeviceToHost, h->st));
        CU(cudaMemcpyAsync(ds, h->F.dstat, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        if (!cnt[6] || !h->F.tc_update) break;
        // the tcgen05 trailing updates met an operand they cannot represent (non-finite entry) or a pipeline timeout:
        // this workspace goes back to the fp64 DMMA updates for good and the factorisation is redone
        ldlt_disable_tc(h->F);
    }
    *n_neg = cnt[0];
    *n_zero = cnt[1];
    *rcond = (cnt[1] > 0 || !(ds[1] > 0.0)) ? 0.0 : ds[0] / ds[1];
    if (abandoned) *abandoned = cnt[4];
    return 0;
}
// verdict of the background delta = 0 test (call after stB has been synchronised)
static void background_verdict(Eng* h, int* n_neg, int* n_zero, double* rcond, int* abandoned) {
    *n_neg = h->h_cntB[0];
    *n_zero = h->h_cntB[1];
    *abandoned = h->h_cntB[4];
    *rcond = (*n_zero > 0 || !(h->h_dsB[1] > 0.0)) ? 0.0 : h->h_dsB[0] / h->h_dsB[1];
}
// reghess (pyipm.py:1373-1406) on the condensed matrix: full-K inertia (D+N, M+N, 0)  <=>  condensed has
// exactly M negative pivots (Haynsworth).
// Abandoning a failed test early hides the pivots that were never computed, and with them the rcond <= eps test of
// pyipm.py:1381 that decides the eq-block regularisation.  The usual case (rcond far above eps) is unaffected; if any
// LATER, completed attempt of the same step does look singular the whole sequence is redone without abandoning, so
// the decisions are always those of the reference's loop.
static int factor_regularised(Eng* h, b200ipm_step_info* info, bool allow_abandon = true, bool allow_spec = true) {
    const int M = h->M;
    const int limit = (allow_abandon && !(h->p.flags & B200IPM_FLAG_NO_ABANDON)) ? M : 0x7fffffff;
    const double delta_in = h->delta;
    int n_neg = 0, n_zero = 0, nfac = 0, ab_first = 0, ab = 0;
    double rcond = 0.0, rcond0 = 0.0;
    const double delta1 = (h->delta == 0.0) ? h->p.reg_coef : std::max(h->delta / 2.0, h->p.reg_coef);
    const bool spec = allow_spec && (h->delta > 0.0) && !(h->p.flags & B200IPM_FLAG_NO_SPECULATION) && !h->strict_retry &&
                      h->first_failed_last;
    int eq_reg = 0;
    bool redo = false, first_failed = false;
    double reg = 0.0;
    h->pendingA = false;
    const bool use_cert = spec && !(h->p.flags & B200IPM_FLAG_NO_CERT);
    if (spec) {
        if (!use_cert) RET(spec_launch_background(h, limit));                    // A: delta = 0, background
        RET(factor_once(h, delta1, 0.0, limit, &n_neg, &n_zero, &rcond, &ab));  // B: the expected candidate
        nfac = 2;
        if (n_neg == M && !(use_cert && rcond <= h->p.eps)) {
            // proceed with B; A's verdict (or the proof that it fails) is collected by resolve_pending() after the solve
            h->delta = delta1;
            h->pendingA = true;
            h->cert_pending = use_cert;
            if (use_cert) RET(cert_launch(h));
            h->pend_delta_in = delta_in;
            h->pend_rcondB = rcond;
            if (info) {
                info->n_neg = n_neg; info->n_zero = n_zero; info->n_factor = nfac; info->eq_reg = 0; info->delta = h->delta;
                info->n_spec = 1; info->spec_used = 1;
            }
            return 0;
        }
        if (use_cert) {   // B failed and A was never started: the plain loop from the incoming delta
            h->delta = delta_in;
            return factor_regularised(h, info, allow_abandon, false);
        }
        // B failed too: A's verdict is needed now to continue the reference's sequence
        int nA = 0, zA = 0;
        CU(cudaStreamSynchronize(h->stB));
        background_verdict(h, &nA, &zA, &rcond0, &ab_first);
        if (info) { info->n_neg_first = nA; info->n_zero_first = zA; }
        if (!(rcond0 <= h->p.eps || nA != M)) {
            // the unshifted matrix passes although the shifted one did not (only possible through rounding or a
            // rank-deficient Jacobian): take the plain sequential loop from the incoming delta
            h->delta = delta_in;
            return factor_regularised(h, info, allow_abandon, false);
        }
        first_failed = true;
        if (rcond0 <= h->p.eps && M) {
            reg = h->p.reg_coef * h->p.eta * pow(h->mu_host, h->p.beta);
            eq_reg = 1;
        }
        h->delta = delta1;
        if (reg != 0.0) RET(factor_once(h, h->delta, reg, limit, &n_neg, &n_zero, &rcond, &ab));   // B had assumed reg = 0
    } else {
        RET(factor_once(h, 0.0, 0.0, limit, &n_neg, &n_zero, &rcond, &ab_first));
        nfac = 1;
        rcond0 = rcond;
        h->first_failed_last = (rcond <= h->p.eps || n_neg != M);
        if (info) { info->n_neg_first = n_neg; info->n_zero_first = n_zero; }
        if (rcond <= h->p.eps || n_neg != M) {
            first_failed = true;
            if (rcond <= h->p.eps && M) {
                reg = h->p.reg_coef * h->p.eta * pow(h->mu_host, h->p.beta);
                eq_reg = 1;
            }
            h->delta = delta1;
            RET(factor_once(h, h->delta, reg, limit, &n_neg, &n_zero, &rcond, &ab));
            nfac++;
        }
    }
    if (first_failed) {
        if (ab_first && !ab && rcond <= h->p.eps) redo = true;
        int guard = 0;
        while (n_neg != M && !redo) {
            if (++guard > h->p.max_reg_retries) {
                if (ab_first) { redo = true; break; }
                return fail_msg("reghess: inertia correction did not converge");
            }
            h->delta *= 10.0;
            RET(factor_once(h, h->delta, reg, limit, &n_neg, &n_zero, &rcond, &ab));
            nfac++;
            if (ab_first && !ab && rcond <= h->p.eps) redo = true;
        }
    }
    if (redo) {
        h->delta = delta_in;
        return factor_regularised(h, info, false, false);
    }
    if (info) {
        info->n_neg = n_neg; info->n_zero = n_zero; info->n_factor = nfac; info->rcond = rcond0; info->eq_reg = eq_reg;
        info->delta = h->delta; info->n_spec = spec ? 1 : 0; info->spec_used = 0;
        info->abandoned_first = ab_first;
    }
    return 0;
}
// Collect the verdict of the background delta = 0 test.  *redo = true: the tentative choice of B was not what the
// reference's loop would have done (A passes, A asks for the eq-block regularisation, or A was abandoned while B
// looks singular) -- the caller restarts the step's factorisation sequentially from the incoming delta.
static int resolve_pending(Eng* h, b200ipm_step_info* info, bool* redo, bool* resolve_only) {
    *redo = false;
    *resolve_only = false;
    if (!h->pendingA) return 0;
    h->pendingA = false;
    CU(cudaStreamSynchronize(h->stB));
    if (h->cert_pending) {
        h->cert_pending = false;
        bool proven = false;
        for (int it = 0; it < 2 && !proven; it++) {
            const double q = h->h_cert[4 * it], nv = h->h_cert[4 * it + 1], hn = h->h_cert[4 * it + 2], vu = h->h_cert[4 * it + 3];
            if (!(nv > 0.0) || !(hn >= 0.0) || !(fabs(q) <= 1.7e308) || !(fabs(vu) <= 1.7e308)) continue;
            const double rho = q / nv;                              // Rayleigh quotient of Hb on null(dce'), explicit
            // the same quantity from the solve itself: (Hb + delta1 I) v + dce y = u and dce' v = 0  =>  v'(Hb + delta1 I) v = v'u.
            // The two agree only if the solve was accurate AND v is (numerically) in null(dce'): a consistency check of
            // exactly the assumptions the proof rests on.
            const double rho2 = vu / nv - h->delta_eff;
            const double scale = std::max(sqrt(hn / nv), h->delta_eff);
            if (rho < -1e-7 * scale && rho2 < -1e-7 * scale && fabs(rho - rho2) <= 0.05 * fabs(rho)) proven = true;
        }
        if (info) { info->n_neg_first = -1; info->n_zero_first = -1; info->rcond = -1.0; info->abandoned_first = 0; info->cert_used = proven ? 1 : 0; }
        if (proven) {
            h->n_cert_ok++;
            h->first_failed_last = true;
            return 0;
        }
        // no proof (cold start vector, shift much larger than the negative curvature, or the problem has become
        // convex): run the delta = 0 test itself now, keeping the candidate's factorisation, and judge it below
        h->n_cert_miss++;
        const int limit = (h->p.flags & B200IPM_FLAG_NO_ABANDON) ? 0x7fffffff : h->M;
        RET(spec_launch_background(h, limit));
        CU(cudaStreamSynchronize(h->stB));
    }
    int nA = 0, zA = 0, abA = 0;
    double rcA = 0.0;
    background_verdict(h, &nA, &zA, &rcA, &abA);
    if (info) { info->n_neg_first = nA; info->n_zero_first = zA; info->rcond = rcA; info->abandoned_first = abA; }
    const bool a_fails = (rcA <= h->p.eps || nA != h->M);
    const bool a_eqreg = (rcA <= h->p.eps && h->M);
    const bool hidden_singular = abA && (h->pend_rcondB <= h->p.eps);
    h->first_failed_last = a_fails;
    if (a_fails && !a_eqreg && !hidden_singular) return 0;
    h->delta = h->pend_delta_in;
    if (!a_fails && !abA) {
        // the unshifted matrix passes (pyipm.py:1381): its factorisation already exists in the background workspace --
        // adopt it (device-to-device copy of the factor data, ~0.1 ms) and only redo the solve
        RET(ldlt_copy_factor(h->F, h->Fb, h->st));
        h->delta_eff = 0.0;
        h->reg_cur = 0.0;
        if (info) {
            info->n_neg = nA; info->n_zero = zA; info->n_factor = 1; info->eq_reg = 0; info->delta = h->delta;
            info->n_spec = 1; info->spec_used = 0;
        }
        *redo = true;
        *resolve_only = true;
        return 0;
    }
    *redo = true;
    return 0;
}

// ------------------------------------------------------------------------------------------ solve
// K y for the unreduced system -> rho = b - K y, returns ||rho||_inf in h->red[0] (device)
// jt_valid: h->jt already holds J' y_x -- true right after condensed_solve() produced this very y (its expansion step needs
// the same product: one pass over J saved, bitwise the same numbers)
static int kkt_residual_vec(Eng* h, const double* b, const double* y, double* rho, bool jt_valid = false) {
    const int D = h->D, M = h->M, N = h->N, K = h->K;
    if (h->C) {
        RET(gemv_n(h->st, h->J, h->ldJ, D, h->C, y + D + N, nullptr, 0.0, 1.0, h->wx));           // J [y_e; y_i]
        RET(gemv_n(h->st, h->W, h->ldW, D, D, y, h->wx, 1.0, 1.0, h->wx));                        // + W dx
        if (!(jt_valid && N)) RET(gemv_t(h->st, h->J, h->ldJ, D, h->C, y, nullptr, 0.0, 1.0, h->jt, h->scr));   // J' dx
    } else {
        RET(gemv_n(h->st, h->W, h->ldW, D, D, y, nullptr, 0.0, 1.0, h->wx));
    }
    const int nb = cdiv(K, 256);
    kkt_resid_kernel<<<nb, 256, 0, h->st>>>(D, M, N, h->sigma, h->delta_eff, h->reg_cur, b, y, h->wx, h->jt, rho, h->part);
    LAUNCHED();
    max_partials_kernel<<<1, 256, 0, h->st>>>(h->part, nb, h->red + 8);
    LAUNCHED();
    return 0;
}
// generic condensed solve of  K y = b  (b, y in reference ordering, y with the internal multiplier sign)
static int condensed_solve(Eng* h, const double* b, double* y) {
    const int D = h->D, M = h->M, N = h->N;
    if (N) {
        cond_t_kernel<<<cdiv(N, 256), 256, 0, h->st>>>(D, M, N, h->sigma, b, h->tvec);
        LAUNCHED();
        RET(gemv_n(h->st, h->J + M, h->ldJ, D, N, h->tvec, b, 1.0, 1.0, h->rhs));   // b_x + dci t
    } else {
        CU(cudaMemcpyAsync(h->rhs, b, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
    }
    if (M) CU(cudaMemcpyAsync(h->rhs + D, b + D + N, sizeof(double) * M, cudaMemcpyDeviceToDevice, h->st));
    RET(ldlt_solve(h->F, h->rhs, h->sol));
    if (N) RET(gemv_t(h->st, h->J, h->ldJ, D, h->C, h->sol, nullptr, 0.0, 1.0, h->jt, h->scr));
    expand_kernel<<<cdiv(std::max(D, std::max(M, N)), 256), 256, 0, h->st>>>(D, M, N, h->sigma, b, h->sol, h->jt, y);
    LAUNCHED();
    return 0;
}
static int solve_direction(Eng* h, b200ipm_step_info* info) {
    const int K = h->K;
    axpby_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(K, -1.0, h->g, 0.0, nullptr, h->bvec);   // b = -g (pyipm.py:1717)
    LAUNCHED();
    RET(condensed_solve(h, h->bvec, h->ycur));
    // iterative refinement against the UNREDUCED system; stop as soon as the residual is at rounding level
    // relative to the right-hand side (typically after one sweep), at most nrefine sweeps
    double bnorm = 0.0;
    for (int i = 0; i < 4; i++) bnorm = std::max(bnorm, h->last_red[i]);
    const double tol = 1e-14 * std::max(1.0, bnorm);
    for (int it = 0;; it++) {
        RET(kkt_residual_vec(h, h->bvec, h->ycur, h->rho, it == 0));
        if (it >= h->p.nrefine) break;
        RET(fetch_red(h, h->red + 8, 1));
        if (h->h_red[0] <= tol) break;
        RET(condensed_solve(h, h->rho, h->ycor));
        axpby_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(K, 1.0, h->ycur, 1.0, h->ycor, h->ycur);
        LAUNCHED();
    }
    // safety net for the relaxed pivot threshold: a poor refined residual triggers ONE strict (Bunch-Kaufman
    // threshold) re-factorisation of the same matrix and a fresh solve
    CU(cudaMemcpyAsync(h->h_fi + 9, h->F.serr, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    RET(fetch_red(h, h->red + 8, 1));
    if (h->h_fi[9]) {
        CU(cudaMemsetAsync(h->F.serr, 0, sizeof(int), h->st));
        return fail_msg("triangular solve: a block result never arrived (device poll timed out); the direction is not usable");
    }
    if (!(h->h_red[0] <= 1e-7 * std::max(1.0, bnorm)) && h->F.pivot_u < 0.64 && !h->strict_retry) {
        if (h->pendingA && h->stB) CU(cudaStreamSynchronize(h->stB));   // the certificate solve reads this factorisation
        const double u_save = h->F.pivot_u;
        h->F.pivot_u = 0.6403882032022076;
        h->strict_retry = true;
        RET(build_kc(h, h->delta_eff, h->reg_cur));
        RET(ldlt_factor(h->F));
        h->n_phys++;
        int r = solve_direction(h, info);
        h->strict_retry = false;
        h->F.pivot_u = u_save;
        h->n_strict++;
        return r;
    }
    flip_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(h->D, h->N, K, h->ycur, h->dz);
    LAUNCHED();
    (void)info;
    return 0;
}

// ------------------------------------------------------------------------------------------ merit at a point
// (f, ||con||_1, sum log s) at an explicit point (xt, st) -> h->trial[0..2] (device)
static int merit_pieces_at(Eng* h, const double* xt, const double* st) {
    const int D = h->D, M = h->M, N = h->N;
    if (h->kind == KIND_POLY) {
        poly_point_merit_kernel<<<1, 256, 0, h->st>>>(D, M, N, h->poly, xt, st, h->trial);
        LAUNCHED();
    } else if (h->kind == KIND_QUAD) {
        // images of xt into the direction slots, then a k=0 "trial" with alpha = 0 around (xt, st)
        RET(quad_images(h, xt, h->qd, h->ad, h->ud, h->gd, h->vd));
        QuadImages im{h->qd, h->ad, h->ud, h->gd, h->vd, h->qd, h->ad, h->ud, h->gd, h->vd};
        quad_trial_kernel<<<1, 256, 0, h->st>>>(D, M, N, quad_data(h), im, xt, st, xt, st, 0.0, 1.0, 0, h->trial);
        LAUNCHED();
    } else {
        return fail_msg("merit evaluation needs a lowered problem (quad / poly)");
    }
    return 0;
}
// trials alpha0 * tau^(k0 + k), k = 0..nb-1 along the current dz -> h->trial (3 per trial), copied to host
static int merit_trials_launch(Eng* h, double alpha0, int k0, int nb, const double* alpha_dev) {
    const int D = h->D, M = h->M, N = h->N;
    if (h->kind == KIND_POLY) {
        if (sizeof(double) * (D + N) > 40 * 1024) return fail_msg("polynomial lowering supports D + N <= 5120");
        poly_trial_kernel<<<nb, 256, sizeof(double) * (D + N), h->st>>>(D, M, N, h->poly, h->x, h->s, h->dz, h->dz + D,
                                                                         alpha0, h->p.tau, k0, h->trial, alpha_dev);
        LAUNCHED();
    } else if (h->kind == KIND_QUAD) {
        QuadImages im{h->qx, h->ax, h->ux, h->gx, h->vx, h->qd, h->ad, h->ud, h->gd, h->vd};
        quad_trial_kernel<<<nb, 256, 0, h->st>>>(D, M, N, quad_data(h), im, h->x, h->s, h->dz, h->dz + D, alpha0, h->p.tau,
                                                 k0, h->trial, alpha_dev);
        LAUNCHED();
    } else {
        return fail_msg("line search needs a lowered problem (quad / poly)");
    }
    return 0;
}
static int merit_trials(Eng* h, double alpha0, int k0, int nb, std::vector<double>& out) {
    RET(merit_trials_launch(h, alpha0, k0, nb, nullptr));
    out.resize((size_t)3 * nb);
    CU(cudaMemcpyAsync(out.data(), h->trial, sizeof(double) * 3 * nb, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

// ------------------------------------------------------------------------------------------ F2 helpers
static int ensure_F2(Eng* h, int n) {
    if (h->F2_ready && h->F2_n == n) return 0;
    if (h->F2alt_ready && h->F2alt_n == n) {      // back to the other order: swap instead of re-creating
        std::swap(h->F2, h->F2alt);
        std::swap(h->F2_ready, h->F2alt_ready);
        std::swap(h->F2_n, h->F2alt_n);
        return 0;
    }
    if (h->F2_ready) {                             // keep the current one as the alternate
        if (h->F2alt_ready) ldlt_free(h->F2alt);
        std::swap(h->F2, h->F2alt);
        h->F2alt_ready = true;
        h->F2alt_n = h->F2_n;
        h->F2_ready = false;
        h->F2 = LdltWs();
    }
    RET(ldlt_alloc(h->F2, n, h->st));
    h->F2_ready = true;
    h->F2_n = n;
    return 0;
}
static int max_row_sqnorm(Eng* h, const double* Mx, int ld, int rows, int cols, double* out) {
    const int nb = std::min(rows, 296);
    row_sqnorm_max_kernel<<<nb, 256, 0, h->st>>>(Mx, ld, rows, cols, h->part);
    LAUNCHED();
    max_partials_kernel<<<1, 256, 0, h->st>>>(h->part, nb, h->red + 9);
    LAUNCHED();
    RET(fetch_red(h, h->red + 9, 1));
    *out = h->h_red[0];
    return 0;
}

// ------------------------------------------------------------------------------------------ second-order correction
// dz_p = -lstsq(A, c_new) with A = jaco(x0)' (pyipm.py:1468-1477, 1520-1529): minimum-norm solution through the
// normal equations  z = A' (A A' + eps I)^-1 c  with iterated-Tikhonov sweeps,  A A' = J'J + diag(0_M, I_N).
static int soc_direction(Eng* h, const double* cnew /* device, M+N */, double* pz /* device, D+N */) {
    const int D = h->D, M = h->M, N = h->N, C = h->C;
    if (!h->Jt) RET(dalloc(&h->Jt, (size_t)C * rup(D, 16)));
    const int ldt = (int)rup(D, 16);
    RET(transpose(h->st, h->J, h->ldJ, D, C, h->Jt, ldt));
    RET(ensure_F2(h, C));
    double scale = 0.0;
    RET(max_row_sqnorm(h, h->Jt, ldt, C, D, &scale));
    scale += 1.0;
    const double tik = 1e-7 * scale;
    // G = Jt Jt' + diag(0, I) + tik I : build SYRK into F2.A then add the slack identity on the diagonal
    GemmArgs a{};
    a.C = h->F2.A; a.ldc = h->F2.ld; a.Cin = nullptr; a.n = C; a.m = C; a.beta = 0.0; a.shift = tik; a.dadd = nullptr;
    a.mode = GEMM_UPPER_MIRROR; a.nterms = 1;
    a.t[0] = GemmTerm{h->Jt, h->Jt, nullptr, ldt, ldt, D, 1.0};
    // dadd: ones on the inequality rows
    CU(cudaMemsetAsync(h->uvec, 0, sizeof(double) * C, h->st));
    if (N) {
        fill_kernel<<<cdiv(N, 256), 256, 0, h->st>>>(N, 1.0, h->uvec + M);
        LAUNCHED();
    }
    a.dadd = h->uvec;
    if (use_tc(h) && C >= 256 && (int)rup((size_t)D, OZ_KB) <= OZ_KMAX) {
        h->oz_soc.variant = 1;
        h->oz_soc.ndiag = 7;
        if (!h->oz_soc.err) h->oz_soc.err = h->oz.err;   // one error word for the engine (owned by h->oz)
        RET(oz_syrk(h->st, a, h->oz_soc, 0u));   // the normal-equations product on tcgen05 as well (weights are all 1)
    } else {
        RET(gemm_nt(h->st, a));
    }
    RET(ldlt_factor(h->F2));
    // iterated Tikhonov: z_{k+1} = z_k + A'(G)^-1 (c - A z_k);   A z = [J' z_x]_e, [J' z_x]_i - z_s
    CU(cudaMemsetAsync(pz, 0, sizeof(double) * (D + N), h->st));
    for (int it = 0; it < 6; it++) {
        // r = c - A z  -> h->cnew-sized scratch h->tvec? use h->ycor (length K >= C)
        if (it == 0) {
            CU(cudaMemcpyAsync(h->ycor, cnew, sizeof(double) * C, cudaMemcpyDeviceToDevice, h->st));
        } else {
            RET(gemv_t(h->st, h->J, h->ldJ, D, C, pz, nullptr, 0.0, 1.0, h->jt, h->scr));   // J' z_x
            axpby_kernel<<<cdiv(C, 256), 256, 0, h->st>>>(C, 1.0, cnew, -1.0, h->jt, h->ycor);
            LAUNCHED();
            if (N) {
                axpby_kernel<<<cdiv(N, 256), 256, 0, h->st>>>(N, 1.0, h->ycor + M, 1.0, pz + D, h->ycor + M);
                LAUNCHED();
            }
            // a consistent system (full-rank Jacobian) converges in one or two sweeps: stop when the residual is at
            // rounding level; an inconsistent / rank-deficient one never gets there and takes all six (the pinv limit)
            absmax2_kernel<<<1, 1024, 0, h->st>>>(C, h->ycor, cnew, h->red + 12);
            LAUNCHED();
            RET(fetch_red(h, h->red + 12, 2));
            if (h->h_red[0] <= 1e-13 * h->h_red[1]) break;
        }
        RET(ldlt_solve(h->F2, h->ycor, h->rho));                                            // u = G^-1 r
        // z_x += J u ; z_s += -u_i
        RET(gemv_n(h->st, h->J, h->ldJ, D, C, h->rho, pz, 1.0, 1.0, pz));
        if (N) {
            axpby_kernel<<<cdiv(N, 256), 256, 0, h->st>>>(N, 1.0, pz + D, -1.0, h->rho + M, pz + D);
            LAUNCHED();
        }
    }
    // dz_p = -z
    axpby_kernel<<<cdiv(D + N, 256), 256, 0, h->st>>>(D + N, -1.0, pz, 0.0, nullptr, pz);
    LAUNCHED();
    return 0;
}

// ------------------------------------------------------------------------------------------ line search
// search() (pyipm.py:1438-1565).  Scalars and branch decisions on the host, every vector operation on the device.
static int line_search(Eng* h, b200ipm_step_info* info, const double* stats /* host dir stats */,
                       const double* first_trial = nullptr /* host: the trial at alpha_smax, already evaluated */) {
    const int D = h->D, M = h->M, N = h->N;
    const double eta = h->p.eta, tau = h->p.tau, eps = h->p.eps;
    const bool con = (M + N) > 0;
    const double f0 = h->last_red[5], c1_old = h->last_red[4];
    const double logs0 = stats[3];
    double phi0 = f0;
    if (con) phi0 += h->nu * c1_old;
    if (N) phi0 -= h->mu * logs0;
    double dphi0 = stats[1];
    if (con) dphi0 -= h->nu * c1_old;
    if (N) dphi0 -= stats[2];
    const double ndx = sqrt(stats[4]), nds = sqrt(stats[5]);
    double a_s = N ? stats[6] : 1.0, a_l = N ? stats[7] : 1.0;
    if (!con) a_l = 0.0;
    info->alpha_smax = a_s;
    info->alpha_lmax = a_l;
    info->phi0 = phi0;
    info->dphi0 = dphi0;
    info->n_backtracks = 0;
    info->soc_tried = info->soc_accepted = 0;
    info->alpha_corr = 0.0;
    info->signal = 0;

    if (h->kind == KIND_QUAD && !first_trial) RET(quad_images(h, h->dz, h->qd, h->ad, h->ud, h->gd, h->vd));
    auto phi_of = [&](const double* t) {
        double v = t[0];
        if (con) v += h->nu * t[1];
        if (N) v -= h->mu * t[2];
        return v;
    };
    std::vector<double> tr;
    if (first_trial) tr.assign(first_trial, first_trial + 3);
    else RET(merit_trials(h, a_s, 0, 1, tr));
    bool correction = false;
    double alpha_corr = 0.0;
    if (phi_of(tr.data()) > phi0 + a_s * eta * dphi0) {
        const double c1_new = tr[1];
        if (con && c1_new > c1_old) {
            // second-order correction (pyipm.py:1464-1489 / 1516-1536)
            info->soc_tried = 1;
            // c_new = con(x0 + a_s dx, s0 + a_s ds): evaluate through the explicit-point path
            trial_point_kernel<<<cdiv(std::max(D, N), 256), 256, 0, h->st>>>(D, N, h->x, h->s, h->dz, a_s, nullptr, 0.0, 1.0,
                                                                            h->xt, h->st_);
            LAUNCHED();
            // con at the trial point: reuse the evaluation kernels on (xt, st)
            if (h->kind == KIND_POLY) {
                poly_eval_kernel<<<1, 256, 0, h->st>>>(D, M, N, h->poly, h->xt, h->trial, nullptr, h->cnew, h->cnew + M,
                                                       nullptr, 0);
                LAUNCHED();
            } else {
                RET(quad_images(h, h->xt, h->qd, h->ad, h->ud, h->gd, h->vd));
                quad_point_kernel<<<1, 1024, 0, h->st>>>(D, M, N, quad_data(h), h->xt, h->qd, h->ad, h->ud, h->gd, h->vd,
                                                         nullptr, nullptr, h->cnew, h->cnew + M, h->trial);
                LAUNCHED();
            }
            if (N) {
                axpby_kernel<<<cdiv(N, 256), 256, 0, h->st>>>(N, 1.0, h->cnew + M, -1.0, h->st_, h->cnew + M);
                LAUNCHED();
            }
            RET(soc_direction(h, h->cnew, h->pvec));
            // phi(x0 + a_s dx + p_x, s0 + a_s ds + p_s)
            trial_point_kernel<<<cdiv(std::max(D, N), 256), 256, 0, h->st>>>(D, N, h->x, h->s, h->dz, a_s, h->pvec, 1.0, 1.0,
                                                                            h->xt, h->st_);
            LAUNCHED();
            RET(merit_pieces_at(h, h->xt, h->st_));
            RET(fetch_red(h, h->trial, 3));
            double t3[3] = {h->h_red[0], h->h_red[1], h->h_red[2]};
            if (phi_of(t3) <= phi0 + a_s * eta * dphi0) {
                if (N) {
                    // alpha_corr = step(s0, a_s ds + p_s)
                    axpby_kernel<<<cdiv(N, 256), 256, 0, h->st>>>(N, a_s, h->dz + D, 1.0, h->pvec + D, h->uvec);
                    LAUNCHED();
                    ftb_kernel<<<1, 1024, 0, h->st>>>(N, h->s, h->uvec, tau, h->red + 10);
                    LAUNCHED();
                    RET(fetch_red(h, h->red + 10, 1));
                    alpha_corr = h->h_red[0];
                    trial_point_kernel<<<cdiv(std::max(D, N), 256), 256, 0, h->st>>>(D, N, h->x, h->s, h->dz, a_s, h->pvec,
                                                                                    1.0, alpha_corr, h->xt, h->st_);
                    LAUNCHED();
                    RET(merit_pieces_at(h, h->xt, h->st_));
                    RET(fetch_red(h, h->trial, 3));
                    double t4[3] = {h->h_red[0], h->h_red[1], h->h_red[2]};
                    if (phi_of(t4) <= phi0 + a_s * eta * dphi0) correction = true;
                } else {
                    alpha_corr = 1.0;
                    correction = true;
                }
            }
            if (h->kind == KIND_QUAD && !correction)   // direction images were clobbered by the explicit-point path
                RET(quad_images(h, h->dz, h->qd, h->ad, h->ud, h->gd, h->vd));
        }
        if (!correction) {
            // backtracking (pyipm.py:1490-1505 / 1537-1551): alpha <- tau * alpha until Armijo holds
            const double a0_s = a_s, a0_l = a_l;
            int k = 1;   // trial index: alpha = a0 * tau^k
            a_s *= tau;
            a_l *= tau;
            info->n_backtracks = 1;
            int batch = std::max(1, h->p.ls_batch);
            bool done = false;
            while (!done) {
                const int nb = std::min(batch, h->max_batch);
                RET(merit_trials(h, a0_s, k, nb, tr));
                for (int q = 0; q < nb; q++) {
                    // a_s, a_l already correspond to trial k+q
                    if (!(phi_of(&tr[3 * q]) > phi0 + a_s * eta * dphi0)) { done = true; break; }
                    const double nrm = N ? sqrt((a_s * ndx) * (a_s * ndx) + (a_l * nds) * (a_l * nds)) : a_s * ndx;
                    if (nrm < eps) {
                        info->signal = -2;   // pyipm.py:1502 / 1548: state left untouched
                        info->alpha_s = info->alpha_l = 0.0;
                        return 0;
                    }
                    a_s *= tau;
                    a_l *= tau;
                    info->n_backtracks++;
                }
                k += nb;
                batch *= 2;
            }
            (void)a0_l;
        }
    }
    info->soc_accepted = correction ? 1 : 0;
    info->alpha_corr = alpha_corr;
    info->alpha_s = a_s;
    info->alpha_l = a_l;
    // state update
    if (correction) {
        // x = x0 + alpha_corr (a_s dx + p_x), s likewise, lda = lda0 + a_l dl
        trial_point_kernel<<<cdiv(std::max(D, N), 256), 256, 0, h->st>>>(D, N, h->x, h->s, h->dz, a_s, h->pvec, 1.0, alpha_corr,
                                                                        h->xt, h->st_);
        LAUNCHED();
        CU(cudaMemcpyAsync(h->x, h->xt, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
        if (N) CU(cudaMemcpyAsync(h->s, h->st_, sizeof(double) * N, cudaMemcpyDeviceToDevice, h->st));
        if (con) {
            axpby_kernel<<<cdiv(M + N, 256), 256, 0, h->st>>>(M + N, 1.0, h->lam, a_l, h->dz + D + N, h->lam);
            LAUNCHED();
        }
    } else {
        update_state_kernel<<<cdiv(std::max(D, M + N), 256), 256, 0, h->st>>>(D, M, N, a_s, a_l, h->dz, h->x, h->s, h->lam);
        LAUNCHED();
    }
    h->eval_valid = false;
    h->hess_valid = false;
    h->resid_valid = false;
    return 0;
}

// ------------------------------------------------------------------------------------------ direction + step
static int compute_direction(Eng* h, b200ipm_step_info* info) {
    h->n_phys = 0;
    RET(residual(h));
    h->oz_off = false;
    h->oz_used = false;
    const double delta_in = h->delta;
    RET(eval_hessian(h));
    CU(cudaEventRecord(h->ev[EV_EVAL], h->st));
    RET(condense(h));
    CU(cudaEventRecord(h->ev[EV_ASSEMBLE], h->st));
    RET(factor_regularised(h, info));
    if (h->oz_used) {
        // the tcgen05 path reports inputs it cannot represent (non-finite entries, a negative weight where none was
        // announced, a pipeline timeout) through a device word: redo the step's contractions in fp64 DMMA then
        int ew = 0;
        CU(cudaMemcpyAsync(&ew, h->oz.err, sizeof(int), cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        if (ew) {
            CU(cudaMemsetAsync(h->oz.err, 0, sizeof(int), h->st));
            h->oz_off = true;
            h->oz_used = false;
            h->hess_valid = false;
            h->delta = delta_in;
            RET(eval_hessian(h));
            RET(condense(h));
            RET(factor_regularised(h, info));
        }
    }
    if (info) info->tc_syrk = h->oz_used ? 1 : 0;
    CU(cudaEventRecord(h->ev[EV_FACTOR], h->st));
    RET(solve_direction(h, info));
    {
        bool redo = false, resolve_only = false;
        RET(resolve_pending(h, info, &redo, &resolve_only));
        if (redo) {
            if (!resolve_only)
                RET(factor_regularised(h, info, /*allow_abandon=*/false, /*allow_spec=*/false));   // every pivot of every test is seen
            RET(solve_direction(h, info));
        }
    }
    CU(cudaEventRecord(h->ev[EV_SOLVE], h->st));
    if (info) info->n_factor_phys = h->n_phys;
    return 0;
}
static int dir_stats(Eng* h, double* stats) {
    dir_stats_kernel<<<1, 1024, 0, h->st>>>(h->D, h->M, h->N, h->df, h->s, h->lam, h->dz, h->mu, h->p.eps, h->p.tau, h->red);
    LAUNCHED();
    RET(fetch_red(h, h->red, 9));
    for (int i = 0; i < 9; i++) stats[i] = h->h_red[i];
    return 0;
}
static void fill_times(Eng* h, b200ipm_step_info* info) {
    float t;
    auto el = [&](int a, int b) { t = 0.f; cudaEventElapsedTime(&t, h->ev[a], h->ev[b]); return t; };
    info->ms_eval = el(EV_START, EV_EVAL);
    info->ms_assemble = el(EV_EVAL, EV_ASSEMBLE);
    info->ms_factor = el(EV_ASSEMBLE, EV_FACTOR);
    info->ms_solve = el(EV_FACTOR, EV_SOLVE);
    info->ms_search = el(EV_SOLVE, EV_SEARCH);
    info->ms_total = el(EV_START, EV_SEARCH);
    info->ms_hess_kernel = el(EV_HESS0, EV_HESS1);
    info->ms_condense_kernel = el(EV_COND0, EV_COND1);
}


// ------------------------------------------------------------------------------------------ L-BFGS mode
// pyipm.py:993-1371 behind the same boundary.  The O(D m), O(D (M+N)) and O((M+N)^3) work (storage products, J / J'
// images, the Schur complement B' A^-1 B and its factorisation) runs on the device; the 2m x 2m systems of the compact
// representation are solved on the host (general LU with partial pivoting, as sym_solve does, pyipm.py:18-20).
static int lu_solve_host(int n, std::vector<double>& A, std::vector<double>& b) {
    for (int k = 0; k < n; k++) {
        int piv = k;
        double best = fabs(A[(size_t)k * n + k]);
        for (int i = k + 1; i < n; i++) {
            const double v = fabs(A[(size_t)i * n + k]);
            if (v > best) { best = v; piv = i; }
        }
        if (best == 0.0) return fail_msg("L-BFGS: singular compact-representation system");
        if (piv != k) {
            for (int j = 0; j < n; j++) std::swap(A[(size_t)k * n + j], A[(size_t)piv * n + j]);
            std::swap(b[k], b[piv]);
        }
        const double d = A[(size_t)k * n + k];
        for (int i = k + 1; i < n; i++) {
            const double l = A[(size_t)i * n + k] / d;
            if (l == 0.0) continue;
            for (int j = k + 1; j < n; j++) A[(size_t)i * n + j] -= l * A[(size_t)k * n + j];
            b[i] -= l * b[k];
        }
    }
    for (int i = n - 1; i >= 0; i--) {
        double acc = b[i];
        for (int j = i + 1; j < n; j++) acc -= A[(size_t)i * n + j] * b[j];
        b[i] = acc / A[(size_t)i * n + i];
    }
    return 0;
}
static void lb_reset(Eng* h) {   // lbfgs_init, pyipm.py:993-1005
    h->lb_zeta = h->lb_zeta0;
    h->lb_m = 0;
    h->lb_fail = 0;
    std::fill(h->lb_SS.begin(), h->lb_SS.end(), 0.0);
    std::fill(h->lb_L.begin(), h->lb_L.end(), 0.0);
    std::fill(h->lb_D.begin(), h->lb_D.end(), 0.0);
}
static int lb_alloc(Eng* h, int m) {
    if (h->lb_S && h->lb_max == m) return 0;
    double* old[] = {h->lb_S, h->lb_Y, h->lb_W, h->lb_xold, h->lb_gold, h->lb_dx, h->lb_dg, h->lb_X00, h->lb_X01, h->lb_zg,
                     h->lb_t, h->lb_q, h->lb_dots, h->lb_coef};
    for (double* b : old) cudaFree(b);
    const size_t cap = (size_t)m + 1, D = h->D, C = std::max(h->C, 1), P = h->D + h->N;
    RET(dalloc(&h->lb_S, cap * D)); RET(dalloc(&h->lb_Y, cap * D)); RET(dalloc(&h->lb_W, 2 * cap * D));
    RET(dalloc(&h->lb_xold, D)); RET(dalloc(&h->lb_gold, D)); RET(dalloc(&h->lb_dx, D)); RET(dalloc(&h->lb_dg, D));
    RET(dalloc(&h->lb_X00, 2 * cap * C)); RET(dalloc(&h->lb_X01, 2 * cap * P)); RET(dalloc(&h->lb_zg, (size_t)h->K));
    RET(dalloc(&h->lb_t, C)); RET(dalloc(&h->lb_q, C)); RET(dalloc(&h->lb_dots, 4 * cap + 8)); RET(dalloc(&h->lb_coef, 2 * cap));
    h->lb_max = m;
    h->lb_SS.assign(cap * cap, 0.0);
    h->lb_L.assign(cap * cap, 0.0);
    h->lb_D.assign(cap * cap, 0.0);
    return 0;
}
// lbfgs_update, pyipm.py:1282-1371.  gradx_old: dL/dx at (x_old; current s, lda) on the host, or NULL for a lowered problem
// (then it is evaluated here; the reference recomputes it the same way, pyipm.py:1706).
static int lb_update(Eng* h, const double* gradx_old) {
    const int D = h->D, C = h->C, cap = h->lb_max + 1;
    const double eps = h->p.eps, rt = sqrt(eps);
    if (gradx_old) {
        RET(up(h, h->lb_gold, gradx_old, D, 0));
    } else {
        if (h->kind != KIND_QUAD && h->kind != KIND_POLY) return fail_msg("lbfgs_update: gradx_old is required in callable mode");
        CU(cudaMemcpyAsync(h->xt, h->x, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
        CU(cudaMemcpyAsync(h->x, h->lb_xold, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
        h->eval_valid = false;
        RET(eval_derivs(h));
        if (C) RET(gemv_n(h->st, h->J, h->ldJ, D, C, h->lam, h->df, 1.0, -1.0, h->lb_gold));
        else CU(cudaMemcpyAsync(h->lb_gold, h->df, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
        CU(cudaMemcpyAsync(h->x, h->xt, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
        h->eval_valid = false;
    }
    RET(residual(h));     // g at the current point (also the g of the direction, pyipm.py:1711)
    axpby_kernel<<<cdiv(D, 256), 256, 0, h->st>>>(D, 1.0, h->g, -1.0, h->lb_gold, h->lb_dg);       // dg = g_old - g_new (g = -grad)
    LAUNCHED();
    axpby_kernel<<<cdiv(D, 256), 256, 0, h->st>>>(D, 1.0, h->x, -1.0, h->lb_xold, h->lb_dx);
    LAUNCHED();
    dots_kernel<<<1, 256, 0, h->st>>>(D, h->lb_dg, h->lb_dx, D, h->lb_dots);
    LAUNCHED();
    dots_kernel<<<1, 256, 0, h->st>>>(D, h->lb_dx, h->lb_dx, D, h->lb_dots + 1);
    LAUNCHED();
    dots_kernel<<<1, 256, 0, h->st>>>(D, h->lb_dg, h->lb_dg, D, h->lb_dots + 2);
    LAUNCHED();
    RET(fetch_red(h, h->lb_dots, 3));
    const double dgdx = h->h_red[0], dxdx = h->h_red[1], dgdg = h->h_red[2];
    const double zeta_new = C ? dgdx / (dxdx + eps) : dgdx / (dgdg + eps);
    if (dgdx > rt && zeta_new > rt) {
        h->lb_zeta = zeta_new;
        int m = h->lb_m;
        if (m > h->lb_max) {
            for (int k = 1; k < m; k++) {
                CU(cudaMemcpyAsync(h->lb_S + (size_t)(k - 1) * D, h->lb_S + (size_t)k * D, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
                CU(cudaMemcpyAsync(h->lb_Y + (size_t)(k - 1) * D, h->lb_Y + (size_t)k * D, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
            }
            for (int i = 0; i + 1 < m; i++)
                for (int j = 0; j + 1 < m; j++) {
                    h->lb_SS[(size_t)i * cap + j] = h->lb_SS[(size_t)(i + 1) * cap + j + 1];
                    h->lb_L[(size_t)i * cap + j] = h->lb_L[(size_t)(i + 1) * cap + j + 1];
                    h->lb_D[(size_t)i * cap + j] = h->lb_D[(size_t)(i + 1) * cap + j + 1];
                }
        } else {
            m += 1;
            for (int i = 0; i < m; i++) {
                h->lb_SS[(size_t)i * cap + m - 1] = h->lb_SS[(size_t)(m - 1) * cap + i] = 0.0;
                h->lb_L[(size_t)i * cap + m - 1] = h->lb_L[(size_t)(m - 1) * cap + i] = 0.0;
                h->lb_D[(size_t)i * cap + m - 1] = h->lb_D[(size_t)(m - 1) * cap + i] = 0.0;
            }
            h->lb_m = m;
        }
        CU(cudaMemcpyAsync(h->lb_S + (size_t)(m - 1) * D, h->lb_dx, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
        CU(cudaMemcpyAsync(h->lb_Y + (size_t)(m - 1) * D, h->lb_dg, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
        // constrained: SS_update = S' dx, L_update = dx' Y;  unconstrained: "SS" = Y' dg (YY), "L" = S' dg (R)
        dots_kernel<<<m, 256, 0, h->st>>>(D, C ? h->lb_dx : h->lb_dg, C ? h->lb_S : h->lb_Y, D, h->lb_dots);
        LAUNCHED();
        dots_kernel<<<m, 256, 0, h->st>>>(D, C ? h->lb_dx : h->lb_dg, C ? h->lb_Y : h->lb_S, D, h->lb_dots + cap);
        LAUNCHED();
        RET(fetch_red(h, h->lb_dots, 2 * cap));
        for (int i = 0; i < m; i++) {
            h->lb_SS[(size_t)i * cap + m - 1] = h->h_red[i];
            h->lb_SS[(size_t)(m - 1) * cap + i] = h->h_red[i];
        }
        if (C) {
            for (int j = 0; j < m; j++) h->lb_L[(size_t)(m - 1) * cap + j] = h->h_red[cap + j];
            h->lb_L[(size_t)(m - 1) * cap + m - 1] = 0.0;
        } else {
            for (int i = 0; i < m; i++) h->lb_L[(size_t)i * cap + m - 1] = h->h_red[cap + i];
        }
        h->lb_D[(size_t)(m - 1) * cap + m - 1] = dgdx;
        h->lb_fail = 0;
    } else {
        h->lb_fail++;
    }
    if (h->lb_fail > h->lb_max && h->lb_m > 0) lb_reset(h);
    CU(cudaMemcpyAsync(h->lb_xold, h->x, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
    return 0;
}
// lbfgs_dir, pyipm.py:1184-1246 with the graph of lbfgs_builder (1007-1182); result in h->dz (reference sign convention)
static int lb_direction(Eng* h, b200ipm_step_info* info) {
    const int D = h->D, M = h->M, N = h->N, C = h->C, K = h->K, m = h->lb_m, P = D + N;
    const double zeta = h->lb_zeta;
    RET(residual(h));
    axpby_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(K, -1.0, h->g, 0.0, nullptr, h->bvec);   // g of the reference = -grad
    LAUNCHED();
    CU(cudaMemsetAsync(h->red + 8, 0, sizeof(double), h->st));
    h->lb_eq_reg = 0;
    if (C) {
        // G = B' A^-1 B = J'J / zeta + diag(0_M, 1/Sigma)      (pyipm.py:1101-1104)
        const int ldt = (int)rup(D, 16);
        if (!h->Jt) RET(dalloc(&h->Jt, (size_t)C * ldt));
        RET(transpose(h->st, h->J, h->ldJ, D, C, h->Jt, ldt));
        RET(ensure_F2(h, C));
        lb_gdiag_kernel<<<cdiv(std::max(M, N), 256), 256, 0, h->st>>>(M, N, h->sigma, h->uvec);
        LAUNCHED();
        for (int attempt = 0; attempt < 2; attempt++) {
            GemmArgs a{};
            a.C = h->F2.A; a.ldc = h->F2.ld; a.Cin = nullptr; a.n = C; a.m = C; a.beta = 0.0; a.shift = 0.0; a.dadd = h->uvec;
            a.mode = GEMM_UPPER_MIRROR; a.nterms = 1;
            a.t[0] = GemmTerm{h->Jt, h->Jt, nullptr, ldt, ldt, D, 1.0 / zeta};
            RET(gemm_nt(h->st, a));
            if (attempt == 1) {
                // eq block regularised (pyipm.py:1110-1112): + sqrt(eps) eta mu^beta I on G[:M, :M]
                fill_kernel<<<cdiv(M, 256), 256, 0, h->st>>>(M, h->p.reg_coef * h->p.eta * pow(h->mu, h->p.beta), h->lb_q);
                LAUNCHED();
                kc_diag_add_kernel<<<cdiv(M, 256), 256, 0, h->st>>>(h->F2.A, h->F2.ld, M, h->lb_q);
                LAUNCHED();
            }
            RET(ldlt_factor(h->F2));
            if (attempt == 1 || !M) break;
            // rcond of the eq block G[:M, :M] (pyipm.py:1107-1109, eigh there): estimated from the pivots of its LDL^T,
            // which are the first M pivots of G's (positive definite: no interchanges)
            std::vector<double> piv((size_t)M);
            const size_t npad = (size_t)h->F2.nblk * NB;
            CU(cudaMemcpyAsync(piv.data(), h->F2.dinfo + 2 * npad, sizeof(double) * M, cudaMemcpyDeviceToHost, h->st));
            CU(cudaStreamSynchronize(h->st));
            double mn = INFINITY, mx = 0.0;
            for (int i = 0; i < M; i++) { mn = std::min(mn, fabs(piv[i])); mx = std::max(mx, fabs(piv[i])); }
            h->lb_rcond = (mx > 0.0) ? mn / mx : 0.0;
            if (!(h->lb_rcond <= h->p.eps)) break;
            h->lb_eq_reg = 1;
        }
        // t = G^-1 (B' A^-1 g_p - g_d);   Zg = [g_p / Adiag - A^-1 B t ; t]        (pyipm.py:1115-1124, merged by linearity)
        const double* gp = h->bvec;
        const double* gd = h->bvec + P;
        RET(gemv_t(h->st, h->J, h->ldJ, D, C, gp, nullptr, 0.0, 1.0, h->jt, h->scr));
        lb_btainv_kernel<<<cdiv(C, 256), 256, 0, h->st>>>(M, N, zeta, h->jt, N ? gp + D : nullptr, h->sigma, gd, h->lb_q);
        LAUNCHED();
        RET(ldlt_solve(h->F2, h->lb_q, h->lb_t));
        RET(gemv_n(h->st, h->J, h->ldJ, D, C, h->lb_t, nullptr, 0.0, 1.0, h->wx));
        lb_ainvb_kernel<<<cdiv(P, 256), 256, 0, h->st>>>(D, M, N, zeta, gp, 1.0, h->wx, h->lb_t, -1.0, h->sigma, h->lb_zg);
        LAUNCHED();
        CU(cudaMemcpyAsync(h->lb_zg + P, h->lb_t, sizeof(double) * C, cudaMemcpyDeviceToDevice, h->st));
        if (m > 0) {
            const int m2 = 2 * m;
            dim3 gw(cdiv(D, 256), m2);
            lb_build_w_kernel<<<gw, 256, 0, h->st>>>(D, m, D, h->lb_S, h->lb_Y, zeta, 1.0, h->lb_W);   // W = [zeta S, Y]
            LAUNCHED();
            for (int c = 0; c < m2; c++) {
                const double* Wc = h->lb_W + (size_t)c * D;
                double* x00 = h->lb_X00 + (size_t)c * C;
                double* x01 = h->lb_X01 + (size_t)c * P;
                // X00_c = -G^-1 (B' W_c / zeta);  X01_c = W_c / zeta + A^-1 B X00_c      (pyipm.py:1134-1136)
                RET(gemv_t(h->st, h->J, h->ldJ, D, C, Wc, nullptr, 0.0, 1.0, h->jt, h->scr));
                lb_btainv_kernel<<<cdiv(C, 256), 256, 0, h->st>>>(M, N, zeta, h->jt, nullptr, h->sigma, nullptr, h->lb_q);
                LAUNCHED();
                RET(ldlt_solve(h->F2, h->lb_q, x00));
                axpby_kernel<<<cdiv(C, 256), 256, 0, h->st>>>(C, -1.0, x00, 0.0, nullptr, x00);
                LAUNCHED();
                RET(gemv_n(h->st, h->J, h->ldJ, D, C, x00, nullptr, 0.0, 1.0, h->wx));
                CU(cudaMemsetAsync(x01 + D, 0, sizeof(double) * N, h->st));
                CU(cudaMemcpyAsync(x01, Wc, sizeof(double) * D, cudaMemcpyDeviceToDevice, h->st));
                lb_ainvb_kernel<<<cdiv(P, 256), 256, 0, h->st>>>(D, M, N, zeta, x01, 1.0, h->wx, x00, 1.0, h->sigma, x01);
                LAUNCHED();
            }
            // X02 = W' X01 (2m x 2m), v10 = W' Zg_p   ->  host
            std::vector<double> X02((size_t)m2 * m2), v10((size_t)m2), tmp((size_t)m2);
            double* dd = h->lb_dots;   // 2*cap + ... entries: reuse per column
            for (int c = 0; c <= m2; c++) {
                const double* vec = (c < m2) ? h->lb_X01 + (size_t)c * P : h->lb_zg;
                dots_kernel<<<m2, 256, 0, h->st>>>(D, vec, h->lb_W, D, dd);
                LAUNCHED();
                CU(cudaMemcpyAsync(tmp.data(), dd, sizeof(double) * m2, cudaMemcpyDeviceToHost, h->st));
                CU(cudaStreamSynchronize(h->st));
                for (int a2 = 0; a2 < m2; a2++) {
                    if (c < m2) X02[(size_t)a2 * m2 + c] = tmp[a2];
                    else v10[a2] = tmp[a2];
                }
            }
            // (X02 - Minv) v11 = v10,  Minv = [[zeta SS, L], [L', -D]]        (pyipm.py:1138-1144)
            const int cap = h->lb_max + 1;
            for (int i = 0; i < m; i++)
                for (int j = 0; j < m; j++) {
                    X02[(size_t)i * m2 + j] -= zeta * h->lb_SS[(size_t)i * cap + j];
                    X02[(size_t)i * m2 + m + j] -= h->lb_L[(size_t)i * cap + j];
                    X02[(size_t)(m + i) * m2 + j] -= h->lb_L[(size_t)j * cap + i];
                    X02[(size_t)(m + i) * m2 + m + j] += h->lb_D[(size_t)i * cap + j];
                }
            RET(lu_solve_host(m2, X02, v10));
            CU(cudaMemcpyAsync(h->lb_coef, v10.data(), sizeof(double) * m2, cudaMemcpyHostToDevice, h->st));
            // dz = Zg - [X01; -X00] v11
            lb_combine_kernel<<<cdiv(P, 256), 256, 0, h->st>>>(P, m2, P, h->lb_X01, h->lb_coef, -1.0, h->lb_zg);
            LAUNCHED();
            lb_combine_kernel<<<cdiv(C, 256), 256, 0, h->st>>>(C, m2, C, h->lb_X00, h->lb_coef, 1.0, h->lb_zg + P);
            LAUNCHED();
            CU(cudaStreamSynchronize(h->st));    // v10 lives on this stack frame
        }
    } else {
        // unconstrained: dz = zeta g + W [A; B],  W = [S, zeta Y]   (pyipm.py:1149-1175; R = L, Y'Y = SS)
        axpby_kernel<<<cdiv(D, 256), 256, 0, h->st>>>(D, zeta, h->bvec, 0.0, nullptr, h->lb_zg);
        LAUNCHED();
        if (m > 0) {
            const int m2 = 2 * m, cap = h->lb_max + 1;
            dim3 gw(cdiv(D, 256), m2);
            lb_build_w_kernel<<<gw, 256, 0, h->st>>>(D, m, D, h->lb_S, h->lb_Y, 1.0, zeta, h->lb_W);
            LAUNCHED();
            dots_kernel<<<m2, 256, 0, h->st>>>(D, h->bvec, h->lb_W, D, h->lb_dots);
            LAUNCHED();
            std::vector<double> wtg((size_t)m2);
            CU(cudaMemcpyAsync(wtg.data(), h->lb_dots, sizeof(double) * m2, cudaMemcpyDeviceToHost, h->st));
            CU(cudaStreamSynchronize(h->st));
            std::vector<double> Lm((size_t)m * m), Lt((size_t)m * m), bv(wtg.begin(), wtg.begin() + m), rhs((size_t)m), t2(wtg.begin() + m, wtg.end());
            for (int i = 0; i < m; i++)
                for (int j = 0; j < m; j++) { Lm[(size_t)i * m + j] = h->lb_L[(size_t)i * cap + j]; Lt[(size_t)j * m + i] = h->lb_L[(size_t)i * cap + j]; }
            std::vector<double> A1 = Lm;
            RET(lu_solve_host(m, A1, bv));                       // B = -L^-1 WT_g[:m]
            for (int i = 0; i < m; i++) bv[i] = -bv[i];
            for (int i = 0; i < m; i++) {
                double acc = 0.0;
                for (int j = 0; j < m; j++) acc += (h->lb_D[(size_t)i * cap + j] + zeta * h->lb_SS[(size_t)i * cap + j]) * bv[j];
                rhs[i] = acc;
            }
            std::vector<double> A2 = Lt, A3 = Lt;
            RET(lu_solve_host(m, A2, rhs));                      // L^-T (D + zeta SS) B
            RET(lu_solve_host(m, A3, t2));                       // L^-T WT_g[m:]
            std::vector<double> coef((size_t)m2);
            for (int i = 0; i < m; i++) { coef[i] = -rhs[i] - t2[i]; coef[m + i] = bv[i]; }
            CU(cudaMemcpyAsync(h->lb_coef, coef.data(), sizeof(double) * m2, cudaMemcpyHostToDevice, h->st));
            lb_combine_kernel<<<cdiv(D, 256), 256, 0, h->st>>>(D, m2, D, h->lb_W, h->lb_coef, 1.0, h->lb_zg);
            LAUNCHED();
            CU(cudaStreamSynchronize(h->st));
        }
    }
    flip_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(D, N, K, h->lb_zg, h->dz);     // pyipm.py:1723-1725
    LAUNCHED();
    if (info) { info->eq_reg = h->lb_eq_reg; info->rcond = h->lb_rcond; info->n_factor = 0; info->delta = h->delta; }
    return 0;
}

// =========================================================================================== C ABI
extern "C" {

int b200ipm_version(void) { return B200IPM_VERSION; }
const char* b200ipm_last_error(void) { return g_last_error.c_str(); }
long long b200ipm_launch_count(void) { return g_launches.load(); }
int b200ipm_struct_size(int which) {
    return which == 0 ? (int)sizeof(b200ipm_params) : (which == 1 ? (int)sizeof(b200ipm_step_info) : -1);
}

int b200ipm_create(int D, int M, int N, const b200ipm_params* p, int device, void* stream, b200ipm_handle* out) {
    if (!out || !p || D <= 0 || M < 0 || N < 0) return fail_msg("b200ipm_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail_msg("b200ipm_create: no CUDA device available (this library has no CPU fallback)");
    CU(cudaSetDevice(device));
    Eng* h = new Eng();
    h->D = D; h->M = M; h->N = N; h->C = M + N; h->K = D + 2 * N + M; h->Kc = D + M;
    h->ldJ = (int)rup(std::max(h->C, 1), 16);
    h->ldW = (int)rup(D, 16);
    h->device = device;
    h->p = *p;
    h->mu = p->mu; h->nu = p->nu; h->delta = 0.0; h->mu_host = p->mu;
    if (stream) { h->st = (cudaStream_t)stream; h->own_stream = false; }
    else { CU(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)); h->own_stream = true; }
    for (int i = 0; i < EV_N; i++) CU(cudaEventCreate(&h->ev[i]));
    RET(ldlt_init_attrs());
    RET(ldlt_init_solve_attrs());
    const int K = h->K, C = h->C;
    RET(dalloc(&h->x, D)); RET(dalloc(&h->s, N)); RET(dalloc(&h->lam, C));
    RET(dalloc(&h->fval, 8)); RET(dalloc(&h->df, D)); RET(dalloc(&h->ce, M)); RET(dalloc(&h->ci, N));
    RET(dalloc(&h->J, (size_t)D * h->ldJ)); RET(dalloc(&h->W, (size_t)D * h->ldW)); RET(dalloc(&h->Hb, (size_t)D * h->ldW));
    CU(cudaMemsetAsync(h->J, 0, sizeof(double) * (size_t)D * h->ldJ, h->st));
    RET(dalloc(&h->g, K)); RET(dalloc(&h->sigma, N)); RET(dalloc(&h->bvec, K)); RET(dalloc(&h->tvec, N));
    RET(dalloc(&h->rhs, h->Kc)); RET(dalloc(&h->sol, h->Kc)); RET(dalloc(&h->ycur, K)); RET(dalloc(&h->ycor, K));
    RET(dalloc(&h->rho, K)); RET(dalloc(&h->dz, K)); RET(dalloc(&h->wx, D)); RET(dalloc(&h->jt, std::max(C, 1)));
    CU(cudaMemsetAsync(h->jt, 0, sizeof(double) * std::max(C, 1), h->st));
    const size_t scr = std::max(gemv_t_scratch_doubles(D, std::max(C, 1)), gemv_t_scratch_doubles(D, D));
    RET(dalloc(&h->scr, scr));
    RET(dalloc(&h->part, std::max<size_t>(4096, (size_t)cdiv(K, 8) + 64)));
    RET(dalloc(&h->red, 32));
    RET(dalloc(&h->trial, (size_t)3 * h->max_batch));
    RET(dalloc(&h->xt, D)); RET(dalloc(&h->st_, N)); RET(dalloc(&h->pvec, D + N)); RET(dalloc(&h->cnew, C));
    RET(dalloc(&h->uvec, std::max(C, D)));
    RET(dalloc(&h->xdiag, D));
    RET(dalloc(&h->qx, D)); RET(dalloc(&h->ax, M)); RET(dalloc(&h->ux, M)); RET(dalloc(&h->gx, N)); RET(dalloc(&h->vx, N));
    RET(dalloc(&h->qd, D)); RET(dalloc(&h->ad, M)); RET(dalloc(&h->ud, M)); RET(dalloc(&h->gd, N)); RET(dalloc(&h->vd, N));
    CU(cudaMallocHost(&h->h_red, sizeof(double) * 64));
    CU(cudaMallocHost(&h->h_fi, sizeof(int) * 16));
    memset(h->h_fi, 0, sizeof(int) * 16);
    RET(ldlt_alloc(h->F, h->Kc, h->st));
    CU(cudaStreamSynchronize(h->st));
    *out = h;
    return 0;
}

int b200ipm_destroy(b200ipm_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    double* bufs[] = {h->x, h->s, h->lam, h->fval, h->df, h->ce, h->ci, h->J, h->W, h->Hb, h->g, h->sigma, h->bvec, h->tvec,
                      h->rhs, h->sol, h->ycur, h->ycor, h->rho, h->dz, h->wx, h->jt, h->scr, h->part, h->red, h->trial,
                      h->xt, h->st_, h->pvec, h->cnew, h->uvec, h->xdiag, h->qx, h->ax, h->ux, h->gx, h->vx, h->qd, h->ad,
                      h->ud, h->gd, h->vd, h->Q, h->qc, h->At, h->Ut, h->qb, h->Gt, h->Vt, h->qr, h->Jt, h->p_coeff, h->sv_x,
                      h->sv_s, h->sv_lam};
    for (double* b : bufs) cudaFree(b);
    for (double* b : {h->lb_S, h->lb_Y, h->lb_W, h->lb_xold, h->lb_gold, h->lb_dx, h->lb_dg, h->lb_X00, h->lb_X01, h->lb_zg, h->lb_t,
                      h->lb_q, h->lb_dots, h->lb_coef}) cudaFree(b);
    cudaFree(h->p_rowptr); cudaFree(h->p_ptr); cudaFree(h->p_fvar); cudaFree(h->p_fpow); cudaFree(h->d_sig);
    cudaFreeHost(h->h_red);
    if (h->h_fi) cudaFreeHost(h->h_fi);
    if (h->stB) cudaStreamSynchronize(h->stB);
    if (h->cert_ready) {
        ldlt_solvebuf_free(h->csb);
        cudaFree(h->c_rhs); cudaFree(h->c_sol); cudaFree(h->c_hv); cudaFree(h->c_u); cudaFree(h->c_red);
        cudaFreeHost(h->h_cert);
    }
    ldlt_free(h->F);
    if (h->oz_soc.err == h->oz.err) h->oz_soc.err = nullptr;
    oz_free(h->oz_soc);
    oz_free(h->oz);
    if (h->F2_ready) ldlt_free(h->F2);
    if (h->F2alt_ready) ldlt_free(h->F2alt);
    if (h->Fb_ready) {
        ldlt_free(h->Fb);
        cudaFreeHost(h->h_cntB);
        cudaFreeHost(h->h_dsB);
    }
    if (h->stB) {
        cudaStreamDestroy(h->stB);
        cudaEventDestroy(h->ev_fork);
    }
    for (int i = 0; i < EV_N; i++) cudaEventDestroy(h->ev[i]);
    if (h->own_stream) cudaStreamDestroy(h->st);
    delete h;
    return 0;
}

int b200ipm_set_params(b200ipm_handle h, const b200ipm_params* p) {
    if (!h || !p) return fail_msg("null argument");
    h->p = *p;
    return 0;
}
int b200ipm_sync(b200ipm_handle h) {
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

int b200ipm_bind_quad(b200ipm_handle h, const double* Q, const double* c, double q4, const double* At, const double* Ut,
                      const double* b, const double* Gt, const double* Vt, const double* r, int on_device) {
    if (!h || !Q || !c) return fail_msg("bind_quad: Q and c are required");
    CU(cudaSetDevice(h->device));
    const size_t D = h->D, M = h->M, N = h->N;
    if (M && (!At || !b)) return fail_msg("bind_quad: At and b are required when M > 0");
    if (N && (!Gt || !r)) return fail_msg("bind_quad: Gt and r are required when N > 0");
    if (!h->Q) { RET(dalloc(&h->Q, D * D)); RET(dalloc(&h->qc, D)); }
    RET(up(h, h->Q, Q, D * D, on_device)); RET(up(h, h->qc, c, D, on_device));
    h->q4 = q4;
    if (M) {
        if (!h->At) { RET(dalloc(&h->At, D * M)); RET(dalloc(&h->qb, M)); }
        RET(up(h, h->At, At, D * M, on_device)); RET(up(h, h->qb, b, M, on_device));
        if (Ut) { if (!h->Ut) RET(dalloc(&h->Ut, D * M)); RET(up(h, h->Ut, Ut, D * M, on_device)); }
        else if (h->Ut) { cudaFree(h->Ut); h->Ut = nullptr; }
    }
    if (N) {
        if (!h->Gt) { RET(dalloc(&h->Gt, D * N)); RET(dalloc(&h->qr, N)); }
        RET(up(h, h->Gt, Gt, D * N, on_device)); RET(up(h, h->qr, r, N, on_device));
        if (Vt) { if (!h->Vt) RET(dalloc(&h->Vt, D * N)); RET(up(h, h->Vt, Vt, D * N, on_device)); }
        else if (h->Vt) { cudaFree(h->Vt); h->Vt = nullptr; }
    }
    CU(cudaStreamSynchronize(h->st));
    h->kind = KIND_QUAD;
    h->eval_valid = h->hess_valid = h->resid_valid = false;
    return 0;
}

int b200ipm_bind_poly(b200ipm_handle h, int nterms, const int* term_row, const double* term_coeff, const int* term_ptr,
                      const int* fac_var, const int* fac_pow, double xlogx_coeff, double xlogx_shift) {
    if (!h || nterms < 0) return fail_msg("bind_poly: bad arguments");
    CU(cudaSetDevice(h->device));
    const int R = 1 + h->M + h->N;
    std::vector<int> rowptr(R + 1, 0);
    for (int t = 0; t < nterms; t++) {
        if (term_row[t] < 0 || term_row[t] >= R) return fail_msg("bind_poly: term_row out of range");
        if (t > 0 && term_row[t] < term_row[t - 1]) return fail_msg("bind_poly: terms must be sorted by row");
        rowptr[term_row[t] + 1]++;
    }
    for (int r = 0; r < R; r++) rowptr[r + 1] += rowptr[r];
    const int nfac = nterms ? term_ptr[nterms] : 0;
    for (int a = 0; a < nfac; a++)
        if (fac_var[a] < 0 || fac_var[a] >= h->D || fac_pow[a] < 1) return fail_msg("bind_poly: bad factor");
    cudaFree(h->p_rowptr); cudaFree(h->p_ptr); cudaFree(h->p_fvar); cudaFree(h->p_fpow); cudaFree(h->p_coeff);
    RET(dalloc(&h->p_rowptr, R + 1)); RET(dalloc(&h->p_ptr, nterms + 1)); RET(dalloc(&h->p_fvar, nfac));
    RET(dalloc(&h->p_fpow, nfac)); RET(dalloc(&h->p_coeff, nterms));
    CU(cudaMemcpyAsync(h->p_rowptr, rowptr.data(), sizeof(int) * (R + 1), cudaMemcpyHostToDevice, h->st));
    std::vector<int> tp(nterms + 1, 0);
    for (int t = 0; t <= nterms; t++) tp[t] = nterms ? term_ptr[t] : 0;
    CU(cudaMemcpyAsync(h->p_ptr, tp.data(), sizeof(int) * (nterms + 1), cudaMemcpyHostToDevice, h->st));
    if (nfac) {
        CU(cudaMemcpyAsync(h->p_fvar, fac_var, sizeof(int) * nfac, cudaMemcpyHostToDevice, h->st));
        CU(cudaMemcpyAsync(h->p_fpow, fac_pow, sizeof(int) * nfac, cudaMemcpyHostToDevice, h->st));
    }
    if (nterms) CU(cudaMemcpyAsync(h->p_coeff, term_coeff, sizeof(double) * nterms, cudaMemcpyHostToDevice, h->st));
    CU(cudaStreamSynchronize(h->st));
    h->poly = PolyData{nterms, R, h->p_rowptr, h->p_coeff, h->p_ptr, h->p_fvar, h->p_fpow, xlogx_coeff, xlogx_shift};
    h->kind = KIND_POLY;
    h->eval_valid = h->hess_valid = h->resid_valid = false;
    return 0;
}

int b200ipm_set_derivs(b200ipm_handle h, double fval, const double* df, const double* ce, const double* ci, const double* J,
                       const double* d2L, int on_device) {
    if (!h || !df) return fail_msg("set_derivs: df is required");   // d2L may be NULL in L-BFGS mode (never used there)
    CU(cudaSetDevice(h->device));
    const int D = h->D, M = h->M, N = h->N, C = h->C;
    if (C && !J) return fail_msg("set_derivs: J is required when there are constraints");
    const cudaMemcpyKind kd = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    CU(cudaMemcpyAsync(h->fval, &fval, sizeof(double), cudaMemcpyHostToDevice, h->st));
    RET(up(h, h->df, df, D, on_device));
    if (M) RET(up(h, h->ce, ce, M, on_device));
    if (N) RET(up(h, h->ci, ci, N, on_device));
    if (C) CU(cudaMemcpy2DAsync(h->J, sizeof(double) * h->ldJ, J, sizeof(double) * C, sizeof(double) * C, D, kd, h->st));
    if (d2L) {
        CU(cudaMemcpy2DAsync(h->W, sizeof(double) * h->ldW, d2L, sizeof(double) * D, sizeof(double) * D, D, kd, h->st));
        dim3 blk(32, 8), grid(cdiv(D, 32), cdiv(D, 8));
        sym_from_upper_kernel<<<grid, blk, 0, h->st>>>(h->W, h->ldW, D);
        LAUNCHED();
    }
    CU(cudaStreamSynchronize(h->st));   // fval lives on the caller's stack
    if (h->kind == KIND_NONE) h->kind = KIND_CALLABLE;
    h->eval_valid = true;
    h->hess_valid = (d2L != nullptr);
    h->resid_valid = false;
    return 0;
}

int b200ipm_set_state(b200ipm_handle h, const double* x, const double* s, const double* lda, double mu, double nu,
                      double delta) {
    if (!h) return fail_msg("null handle");
    CU(cudaSetDevice(h->device));
    if (x) RET(up(h, h->x, x, h->D, 0));
    if (s) RET(up(h, h->s, s, h->N, 0));
    if (lda) RET(up(h, h->lam, lda, h->C, 0));
    CU(cudaStreamSynchronize(h->st));
    h->mu = mu; h->nu = nu; h->delta = delta;
    if (x) h->eval_valid = false;            // f, df, ce, ci, J depend on x only
    if (x || lda) h->hess_valid = false;     // W depends on (x, lda)
    h->resid_valid = false;
    return 0;
}
int b200ipm_get_state(b200ipm_handle h, double* x, double* s, double* lda, double* mu, double* nu, double* delta) {
    if (!h) return fail_msg("null handle");
    CU(cudaSetDevice(h->device));
    RET(down(h, x, h->x, h->D)); RET(down(h, s, h->s, h->N)); RET(down(h, lda, h->lam, h->C));
    CU(cudaStreamSynchronize(h->st));
    if (mu) *mu = h->mu;
    if (nu) *nu = h->nu;
    if (delta) *delta = h->delta;
    return 0;
}
int b200ipm_set_mu_host(b200ipm_handle h, double mu_host) {
    h->mu_host = mu_host;
    return 0;
}

int b200ipm_state_save(b200ipm_handle h) {
    if (!h) return fail_msg("null handle");
    CU(cudaSetDevice(h->device));
    if (!h->sv_x) { RET(dalloc(&h->sv_x, h->D)); RET(dalloc(&h->sv_s, h->N)); RET(dalloc(&h->sv_lam, h->C)); }
    CU(cudaMemcpyAsync(h->sv_x, h->x, sizeof(double) * h->D, cudaMemcpyDeviceToDevice, h->st));
    if (h->N) CU(cudaMemcpyAsync(h->sv_s, h->s, sizeof(double) * h->N, cudaMemcpyDeviceToDevice, h->st));
    if (h->C) CU(cudaMemcpyAsync(h->sv_lam, h->lam, sizeof(double) * h->C, cudaMemcpyDeviceToDevice, h->st));
    h->sv_mu = h->mu; h->sv_nu = h->nu; h->sv_delta = h->delta; h->sv_mu_host = h->mu_host;
    h->sv_valid = true;
    return 0;
}
int b200ipm_state_restore(b200ipm_handle h) {
    if (!h || !h->sv_valid) return fail_msg("state_restore: nothing saved");
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(h->x, h->sv_x, sizeof(double) * h->D, cudaMemcpyDeviceToDevice, h->st));
    if (h->N) CU(cudaMemcpyAsync(h->s, h->sv_s, sizeof(double) * h->N, cudaMemcpyDeviceToDevice, h->st));
    if (h->C) CU(cudaMemcpyAsync(h->lam, h->sv_lam, sizeof(double) * h->C, cudaMemcpyDeviceToDevice, h->st));
    h->mu = h->sv_mu; h->nu = h->sv_nu; h->delta = h->sv_delta; h->mu_host = h->sv_mu_host;
    h->eval_valid = false;
    h->hess_valid = false;
    h->resid_valid = false;
    return 0;
}

int b200ipm_profile_kernel(b200ipm_handle h, int which, int reps, float* ms_per_launch, double* work) {
    if (!h || reps <= 0) return fail_msg("profile_kernel: bad arguments");
    CU(cudaSetDevice(h->device));
    RET(residual(h));
    RET(eval_hessian(h));
    const int D = h->D, M = h->M, N = h->N, C = h->C;
    double wk = 0.0;
    if (which >= 2) RET(condense(h));
    if (which >= 4 && which <= 5) { RET(build_kc(h, h->delta, 0.0)); RET(ldlt_factor(h->F)); }
    CU(cudaStreamSynchronize(h->st));
    CU(cudaEventRecord(h->ev[EV_START], h->st));
    for (int r = 0; r < reps; r++) {
        switch (which) {
            case 0:
                RET(gemv_n(h->st, h->J, h->ldJ, D, C, h->lam, h->df, 1.0, -1.0, h->g, h->part));
                wk = 8.0 * ((double)D * C + 2.0 * D + C);
                break;
            case 1: {
                GemmArgs a{};
                a.C = h->W; a.ldc = h->ldW; a.Cin = h->Q; a.ldcin = D; a.dadd = h->xdiag; a.n = D; a.m = D; a.beta = 1.0;
                a.mode = GEMM_UPPER_MIRROR; a.nterms = 0;
                if (h->kind != KIND_QUAD) return fail_msg("profile_kernel(1) needs a bound quad problem");
                if (M && h->Ut) a.t[a.nterms++] = GemmTerm{h->Ut, h->Ut, h->lam, M, M, M, -1.0};
                if (N && h->Vt) a.t[a.nterms++] = GemmTerm{h->Vt, h->Vt, h->lam + M, N, N, N, 1.0};
                RET(gemm_nt(h->st, a));
                wk = gemm_nt_flops(a);
                break;
            }
            case 2: {
                GemmArgs a{};
                a.C = h->Hb; a.ldc = h->ldW; a.Cin = h->W; a.ldcin = h->ldW; a.n = D; a.m = D; a.beta = 1.0;
                a.mode = GEMM_UPPER_MIRROR; a.nterms = 0;
                if (N) a.t[a.nterms++] = GemmTerm{h->J + M, h->J + M, h->sigma, h->ldJ, h->ldJ, N, 1.0};
                RET(gemm_nt(h->st, a));
                wk = gemm_nt_flops(a);
                break;
            }
            case 6: {
                GemmArgs a{};
                a.C = h->W; a.ldc = h->ldW; a.Cin = h->Q; a.ldcin = D; a.dadd = h->xdiag; a.n = D; a.m = D; a.beta = 1.0;
                a.mode = GEMM_UPPER_MIRROR; a.nterms = 0;
                if (h->kind != KIND_QUAD) return fail_msg("profile_kernel(6) needs a bound quad problem");
                if (M && h->Ut) a.t[a.nterms++] = GemmTerm{h->Ut, h->Ut, h->lam, M, M, M, -1.0};
                if (N && h->Vt) a.t[a.nterms++] = GemmTerm{h->Vt, h->Vt, h->lam + M, N, N, N, 1.0};
                oz_configure(h);
                h->oz.ndiag = (h->p.flags & B200IPM_FLAG_TCGEN05_FULLCOND) ? 8 : 7;
                RET(oz_syrk(h->st, a, h->oz, (M && h->Ut) ? 1u : 0u));
                wk = gemm_nt_flops(a);
                break;
            }
            case 8: {   // the tcgen05 kernel of case 6 alone (slices kept from the first repetition); work = int8 operations
                GemmArgs a{};
                a.C = h->W; a.ldc = h->ldW; a.Cin = h->Q; a.ldcin = D; a.dadd = h->xdiag; a.n = D; a.m = D; a.beta = 1.0;
                a.mode = GEMM_UPPER_MIRROR; a.nterms = 0;
                if (h->kind != KIND_QUAD) return fail_msg("profile_kernel(8) needs a bound quad problem");
                if (M && h->Ut) a.t[a.nterms++] = GemmTerm{h->Ut, h->Ut, h->lam, M, M, M, -1.0};
                if (N && h->Vt) a.t[a.nterms++] = GemmTerm{h->Vt, h->Vt, h->lam + M, N, N, N, 1.0};
                oz_configure(h);
                h->oz.ndiag = (h->p.flags & B200IPM_FLAG_TCGEN05_FULLCOND) ? 8 : 7;
                if (r == 0) {
                    RET(oz_syrk(h->st, a, h->oz, (M && h->Ut) ? 1u : 0u));
                    CU(cudaEventRecord(h->ev[EV_START], h->st));   // restart the clock after the slicing pass
                }
                h->oz.reuse_slices = true;
                const int rc8 = oz_syrk(h->st, a, h->oz, (M && h->Ut) ? 1u : 0u);
                h->oz.reuse_slices = false;
                RET(rc8);
                wk = oz_syrk_int8_ops(a, oz_variant_bn(h->oz.variant), h->oz.variant == 1 ? h->oz.ndiag : 7);
                break;
            }
            case 7: {
                GemmArgs a{};
                a.C = h->Hb; a.ldc = h->ldW; a.Cin = h->W; a.ldcin = h->ldW; a.n = D; a.m = D; a.beta = 1.0;
                a.mode = GEMM_UPPER_MIRROR; a.nterms = 0;
                if (N) a.t[a.nterms++] = GemmTerm{h->J + M, h->J + M, h->sigma, h->ldJ, h->ldJ, N, 1.0};
                oz_configure(h);
                h->oz.ndiag = (h->p.flags & B200IPM_FLAG_TCGEN05_FULLCOND) ? 8 : 6;
                RET(oz_syrk(h->st, a, h->oz, 0u));
                wk = gemm_nt_flops(a);
                break;
            }
            case 3:
                RET(build_kc(h, h->delta, 0.0));
                RET(ldlt_factor(h->F));
                wk = (double)h->Kc * h->Kc * h->Kc / 3.0;
                break;
            case 4:
                RET(ldlt_solve(h->F, h->rhs, h->sol));
                wk = 8.0 * (double)h->Kc * h->Kc;   // bytes: both triangular sweeps read the factor once each (half matrix x2)
                break;
            case 5:
                RET(gemv_t(h->st, h->J, h->ldJ, D, C, h->x, nullptr, 0.0, 1.0, h->jt, h->scr));
                wk = 8.0 * ((double)D * C + D + C);
                break;
            default:
                return fail_msg("profile_kernel: unknown kernel id");
        }
    }
    CU(cudaEventRecord(h->ev[EV_SEARCH], h->st));
    CU(cudaEventSynchronize(h->ev[EV_SEARCH]));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, h->ev[EV_START], h->ev[EV_SEARCH]));
    if (ms_per_launch) *ms_per_launch = ms / reps;
    if (work) *work = wk;
    h->eval_valid = false;   // W / Hb / g were rewritten with the same values; keep the cache honest anyway
    h->hess_valid = false;
    h->resid_valid = false;
    return 0;
}

int b200ipm_cost(b200ipm_handle h, double* fval) {
    CU(cudaSetDevice(h->device));
    RET(residual(h));
    *fval = h->last_red[5];
    return 0;
}
int b200ipm_residual(b200ipm_handle h, double* g, double kkt_norm[4]) {
    CU(cudaSetDevice(h->device));
    RET(residual(h));
    if (g) { RET(down(h, g, h->g, h->K)); CU(cudaStreamSynchronize(h->st)); }
    if (kkt_norm) for (int i = 0; i < 4; i++) kkt_norm[i] = h->last_red[i];
    return 0;
}
int b200ipm_kkt(b200ipm_handle h, double* kkt1, double* kkt2, double* kkt3, double* kkt4) {
    CU(cudaSetDevice(h->device));
    RET(residual(h));
    const int D = h->D, M = h->M, N = h->N;
    RET(down(h, kkt1, h->g, D));
    if (N && kkt2) {
        // kkt2 = g_s * s (pyipm.py:972)
        mul_kernel<<<cdiv(N, 256), 256, 0, h->st>>>(N, h->g + D, h->s, h->tvec);
        LAUNCHED();
        RET(down(h, kkt2, h->tvec, N));
    }
    if (M) RET(down(h, kkt3, h->g + D + N, M));
    if (N) RET(down(h, kkt4, h->g + D + N + M, N));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}
int b200ipm_con_jac(b200ipm_handle h, double* con, double* J) {
    CU(cudaSetDevice(h->device));
    RET(residual(h));
    const int D = h->D, N = h->N, C = h->C;
    if (con && C) RET(down(h, con, h->g + D + N, C));
    if (J && C)
        CU(cudaMemcpy2DAsync(J, sizeof(double) * C, h->J, sizeof(double) * h->ldJ, sizeof(double) * C, D,
                             cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}
int b200ipm_hess_full(b200ipm_handle h, double* H) {
    CU(cudaSetDevice(h->device));
    RET(residual(h));
    RET(eval_hessian(h));
    const size_t K = h->K;
    double* dH = nullptr;
    RET(dalloc(&dH, K * K));
    full_kkt_kernel<<<(int)std::min<size_t>((K * K + 255) / 256, 148 * 32), 256, 0, h->st>>>(h->D, h->M, h->N, h->W, h->ldW,
                                                                                             h->J, h->ldJ, h->sigma, dH);
    LAUNCHED();
    CU(cudaMemcpyAsync(H, dH, sizeof(double) * K * K, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    cudaFree(dH);
    return 0;
}
int b200ipm_d2L(b200ipm_handle h, double* W) {
    CU(cudaSetDevice(h->device));
    RET(eval_hessian(h));
    CU(cudaMemcpy2DAsync(W, sizeof(double) * h->D, h->W, sizeof(double) * h->ldW, sizeof(double) * h->D, h->D,
                         cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}
int b200ipm_merit(b200ipm_handle h, double* phi, double* dphi) {
    CU(cudaSetDevice(h->device));
    RET(residual(h));
    double stats[9];
    RET(dir_stats(h, stats));
    const bool con = h->C > 0;
    double p0 = h->last_red[5];
    if (con) p0 += h->nu * h->last_red[4];
    if (h->N) p0 -= h->mu * stats[3];
    double d0 = stats[1];
    if (con) d0 -= h->nu * h->last_red[4];
    if (h->N) d0 -= stats[2];
    if (phi) *phi = p0;
    if (dphi) *dphi = d0;
    return 0;
}
int b200ipm_init_slack(b200ipm_handle h) {
    CU(cudaSetDevice(h->device));
    if (!h->N) return 0;
    RET(eval_derivs(h));
    init_slack_kernel<<<cdiv(h->N, 256), 256, 0, h->st>>>(h->N, h->ci, h->p.Ktol, h->s);
    LAUNCHED();
    h->resid_valid = false;
    return 0;
}
// lda0 = pinv(J) df (pyipm.py:723-730): minimum-norm least squares through (J J' + eps I), iterated Tikhonov
int b200ipm_init_lambda(b200ipm_handle h) {
    // lda0 = pinv(J) df (pyipm.py:729-730): the minimum-norm least-squares solution of J lda = df, J = [dce | dci] (D x C),
    // by NONSTATIONARY iterated Tikhonov on the SMALLER Gram matrix (J'J when C < D, J J' otherwise -- nonsingular whenever
    // J has full rank, so the right-hand side of every sweep goes to zero with the residual and no null-space component is
    // amplified by 1/t):   lda += (J'J + t I)^-1 J' (df - J lda)   resp.   lda += J' (J J' + t I)^-1 (df - J lda).
    // A sweep contracts the component along a singular value sigma by t / (sigma^2 + t): with t = 1e-7 sigma_max^2 a
    // well-conditioned Jacobian converges in two or three sweeps (the test below stops it); if the update is still large
    // after six sweeps the Jacobian is ill conditioned and t drops to 1e-10, then 1e-13 (cond(J) up to ~1e6).  An update
    // that has merely reached the rounding floor of its level (exactly dependent constraints) does NOT escalate: a smaller
    // t would only amplify that floor.
    CU(cudaSetDevice(h->device));
    const int D = h->D, M = h->M, N = h->N, C = h->C;
    if (!C) return 0;
    RET(eval_derivs(h));
    const bool small = C < D;
    const int ng = small ? C : D;
    const double* Gop = h->J;
    int ldg = h->ldJ, kg = C;
    if (small) {
        if (!h->Jt) RET(dalloc(&h->Jt, (size_t)C * rup(D, 16)));
        ldg = (int)rup(D, 16);
        RET(transpose(h->st, h->J, h->ldJ, D, C, h->Jt, ldg));
        Gop = h->Jt;
        kg = D;
    }
    RET(ensure_F2(h, ng));
    double scale = 0.0;
    RET(max_row_sqnorm(h, Gop, ldg, ng, kg, &scale));
    if (!(scale > 0.0)) scale = 1.0;
    CU(cudaMemsetAsync(h->lam, 0, sizeof(double) * C, h->st));
    static const double levels[3] = {1e-7, 1e-10, 1e-13};
    bool done = false;
    for (int lv = 0; lv < 3 && !done; lv++) {
        GemmArgs a{};
        a.C = h->F2.A; a.ldc = h->F2.ld; a.Cin = nullptr; a.dadd = nullptr; a.n = ng; a.m = ng; a.beta = 0.0;
        a.shift = levels[lv] * scale; a.mode = GEMM_UPPER_MIRROR; a.nterms = 1;
        a.t[0] = GemmTerm{Gop, Gop, nullptr, ldg, ldg, kg, 1.0};
        RET(gemm_nt(h->st, a));
        RET(ldlt_factor(h->F2));
        const int nsweep = (lv == 2) ? 12 : 6;
        double last = INFINITY;
        for (int it = 0; it < nsweep; it++) {
            RET(gemv_n(h->st, h->J, h->ldJ, D, C, h->lam, h->df, 1.0, -1.0, h->wx));              // r = df - J lda
            double* upd = nullptr;
            if (small) {
                RET(gemv_t(h->st, h->J, h->ldJ, D, C, h->wx, nullptr, 0.0, 1.0, h->rho, h->scr));   // w = J' r
                RET(ldlt_solve(h->F2, h->rho, h->ycor));                                            // u = (J'J + tI)^-1 w
                upd = h->ycor;
            } else {
                RET(ldlt_solve(h->F2, h->wx, h->xt));                                               // u = (JJ' + tI)^-1 r
                RET(gemv_t(h->st, h->J, h->ldJ, D, C, h->xt, nullptr, 0.0, 1.0, h->rho, h->scr));   // J' u
                upd = h->rho;
            }
            axpby_kernel<<<cdiv(C, 256), 256, 0, h->st>>>(C, 1.0, h->lam, 1.0, upd, h->lam);
            LAUNCHED();
            absmax2_kernel<<<1, 1024, 0, h->st>>>(C, upd, h->lam, h->red + 12);
            LAUNCHED();
            RET(fetch_red(h, h->red + 12, 2));
            last = h->h_red[0] / std::max(h->h_red[1], 1e-300);
            if (!(last > 1e-13)) { done = true; break; }
        }
        if (!(last > 1e-8)) done = true;      // at the rounding floor of this level: not slow convergence, do not escalate
    }
    if (N) {
        fix_lambda_kernel<<<cdiv(N, 256), 256, 0, h->st>>>(M, N, h->p.Ktol, h->lam);
        LAUNCHED();
    }
    h->hess_valid = false;   // W depends on lda
    h->resid_valid = false;
    return 0;
}
int b200ipm_update_mu(b200ipm_handle h, double* mu_new) {
    CU(cudaSetDevice(h->device));
    const int N = h->N;
    if (!N) { *mu_new = h->mu; return 0; }
    mu_stats_kernel<<<1, 1024, 0, h->st>>>(h->M, N, h->s, h->lam, h->red + 12);
    LAUNCHED();
    RET(fetch_red(h, h->red + 12, 2));
    const double mn = h->h_red[0], dot = h->h_red[1], eps = h->p.eps;
    const double xi = N * mn / (dot + eps);                         // pyipm.py:1806-1808
    const double t = std::min(0.05 * (1.0 - xi) / (xi + eps), 2.0);
    double mu = 0.1 * (t * t * t) * dot / N;                        // pyipm.py:1809-1810
    if (mu < 0.0) mu = 0.0;
    *mu_new = mu;
    return 0;
}

int b200ipm_direction(b200ipm_handle h, double* dz, b200ipm_step_info* info) {
    if (!h) return fail_msg("null handle");
    CU(cudaSetDevice(h->device));
    b200ipm_step_info tmp{};
    if (!info) info = &tmp;
    memset(info, 0, sizeof(*info));
    CU(cudaEventRecord(h->ev[EV_START], h->st));
    RET(compute_direction(h, info));
    RET(fetch_red(h, h->red + 8, 1));
    info->resid = h->h_red[0];
    if (dz) { RET(down(h, dz, h->dz, h->K)); CU(cudaStreamSynchronize(h->st)); }
    info->mu = h->mu; info->nu = h->nu;
    return 0;
}
int b200ipm_step_max(b200ipm_handle h, double* alpha_smax, double* alpha_lmax) {
    CU(cudaSetDevice(h->device));
    double stats[9];
    RET(dir_stats(h, stats));
    if (alpha_smax) *alpha_smax = h->N ? stats[6] : 1.0;
    if (alpha_lmax) *alpha_lmax = h->N ? stats[7] : 1.0;
    return 0;
}

// ------------------------------------------------------------------------------------------ one-sync step
// The common step of a nonconvex solve (delta > 0 on entry, the previous delta = 0 test failed, certificate enabled) takes
// the same decisions every time: the candidate delta / 2 passes, one refinement sweep brings the unreduced residual to
// rounding level, the failure of the delta = 0 test is proven.  fast_step() therefore ISSUES that whole sequence --
// factorisation, certificate solves, solve, one refinement sweep, direction statistics, the merit trial at the step
// limit -- without waiting for any of its own decisions, reads every scalar back ONCE, and only then checks them.  Any
// check that fails (wrong inertia, singular candidate, poor residual, no proof, tcgen05 error word) falls back to the
// sequential code path from the untouched state: the decisions are those of compute_direction() in every case.
// returns 1 = done (info filled, stats / trial valid), 0 = not applicable or a check failed (caller runs the slow path)
static int fast_step(Eng* h, b200ipm_step_info* info, double* stats, double* trial, int* done) {
    *done = 0;
    const int M = h->M, K = h->K;
    const bool lowered = (h->kind == KIND_QUAD || h->kind == KIND_POLY);
    const bool spec = (h->delta > 0.0) && !(h->p.flags & (B200IPM_FLAG_NO_SPECULATION | B200IPM_FLAG_NO_CERT | B200IPM_FLAG_SLOW_STEP)) &&
                      !h->strict_retry && h->first_failed_last && lowered && h->p.nrefine >= 1;
    if (!spec) return 0;
    h->n_phys = 0;
    RET(residual(h));
    h->oz_off = false;
    h->oz_used = false;
    const double delta_in = h->delta;
    RET(eval_hessian(h));
    CU(cudaEventRecord(h->ev[EV_EVAL], h->st));
    RET(condense(h));
    CU(cudaEventRecord(h->ev[EV_ASSEMBLE], h->st));
    const int limit = (h->p.flags & B200IPM_FLAG_NO_ABANDON) ? 0x7fffffff : M;
    const double delta1 = std::max(h->delta / 2.0, h->p.reg_coef);
    RET(ldlt_set_neg_limit(h->F, limit));
    RET(build_kc(h, delta1, 0.0));
    RET(ldlt_factor(h->F));
    h->n_phys++;
    CU(cudaMemcpyAsync(h->h_fi, h->F.counts, sizeof(int) * 8, cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(h->h_red + 32, h->F.dstat, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->st));
    h->h_fi[8] = 0;
    if (h->oz_used) CU(cudaMemcpyAsync(h->h_fi + 8, h->oz.err, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    h->delta = delta1;
    h->pendingA = true;
    h->cert_pending = true;
    h->pend_delta_in = delta_in;
    RET(cert_launch(h));
    CU(cudaEventRecord(h->ev[EV_FACTOR], h->st));
    // solve + exactly one refinement sweep against the unreduced system
    axpby_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(K, -1.0, h->g, 0.0, nullptr, h->bvec);
    LAUNCHED();
    RET(condensed_solve(h, h->bvec, h->ycur));
    RET(kkt_residual_vec(h, h->bvec, h->ycur, h->rho, true));
    CU(cudaMemcpyAsync(h->red + 16, h->red + 8, sizeof(double), cudaMemcpyDeviceToDevice, h->st));
    RET(condensed_solve(h, h->rho, h->ycor));
    axpby_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(K, 1.0, h->ycur, 1.0, h->ycor, h->ycur);
    LAUNCHED();
    RET(kkt_residual_vec(h, h->bvec, h->ycur, h->rho));
    flip_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(h->D, h->N, K, h->ycur, h->dz);
    LAUNCHED();
    CU(cudaEventRecord(h->ev[EV_SOLVE], h->st));
    dir_stats_kernel<<<1, 1024, 0, h->st>>>(h->D, h->M, h->N, h->df, h->s, h->lam, h->dz, h->mu, h->p.eps, h->p.tau, h->red);
    LAUNCHED();
    if (h->kind == KIND_QUAD) RET(quad_images(h, h->dz, h->qd, h->ad, h->ud, h->gd, h->vd));
    RET(merit_trials_launch(h, 1.0, 0, 1, h->N ? h->red + 6 : nullptr));
    CU(cudaMemcpyAsync(h->h_red, h->red, sizeof(double) * 17, cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(h->h_red + 20, h->trial, sizeof(double) * 3, cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(h->h_fi + 9, h->F.serr, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));                                            // <-- the one read-back
    if (h->h_fi[9]) {
        CU(cudaMemsetAsync(h->F.serr, 0, sizeof(int), h->st));
        return fail_msg("triangular solve: a block result never arrived (device poll timed out); the direction is not usable");
    }
    // ---- checks, in the order the sequential path would have met them
    auto bail = [&]() -> int {
        if (h->stB) CU(cudaStreamSynchronize(h->stB));
        h->pendingA = false;
        h->cert_pending = false;
        h->delta = delta_in;
        h->n_fast_miss++;
        return 0;
    };
    const int n_neg = h->h_fi[0], n_zero = h->h_fi[1];
    const double dmin = h->h_red[32], dmax = h->h_red[33];
    const double rcond = (n_zero > 0 || !(dmax > 0.0)) ? 0.0 : dmin / dmax;
    if (h->h_fi[8]) {   // tcgen05 error word: the slow path redoes the contractions in fp64 DMMA
        CU(cudaMemsetAsync(h->oz.err, 0, sizeof(int), h->st));
        h->oz_off = true;
        h->oz_used = false;
        h->hess_valid = false;
        RET(bail());
        RET(eval_hessian(h));
        return 0;
    }
    if (!(n_neg == M && !(rcond <= h->p.eps))) return bail();
    double bnorm = 0.0;
    for (int i = 0; i < 4; i++) bnorm = std::max(bnorm, h->last_red[i]);
    const double tol = 1e-14 * std::max(1.0, bnorm);
    double res = h->h_red[8];
    bool extra = false;
    for (int it = 1; !(res <= tol) && it < h->p.nrefine; it++) {      // rare: further sweeps, each read back
        RET(condensed_solve(h, h->rho, h->ycor));
        axpby_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(K, 1.0, h->ycur, 1.0, h->ycor, h->ycur);
        LAUNCHED();
        RET(kkt_residual_vec(h, h->bvec, h->ycur, h->rho));
        RET(fetch_red(h, h->red + 8, 1));
        res = h->h_red[0];
        extra = true;
    }
    if (!(res <= 1e-7 * std::max(1.0, bnorm))) return bail();         // the strict re-factorisation is the slow path's job
    if (extra) {
        flip_kernel<<<cdiv(K, 256), 256, 0, h->st>>>(h->D, h->N, K, h->ycur, h->dz);
        LAUNCHED();
        dir_stats_kernel<<<1, 1024, 0, h->st>>>(h->D, h->M, h->N, h->df, h->s, h->lam, h->dz, h->mu, h->p.eps, h->p.tau, h->red);
        LAUNCHED();
        if (h->kind == KIND_QUAD) RET(quad_images(h, h->dz, h->qd, h->ad, h->ud, h->gd, h->vd));
        RET(merit_trials_launch(h, 1.0, 0, 1, h->N ? h->red + 6 : nullptr));
        CU(cudaMemcpyAsync(h->h_red, h->red, sizeof(double) * 9, cudaMemcpyDeviceToHost, h->st));
        CU(cudaMemcpyAsync(h->h_red + 20, h->trial, sizeof(double) * 3, cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
    }
    h->pend_rcondB = rcond;
    info->n_neg = n_neg; info->n_zero = n_zero; info->n_factor = 2; info->eq_reg = 0; info->delta = h->delta;
    info->n_spec = 1; info->spec_used = 1;
    info->tc_syrk = h->oz_used ? 1 : 0;
    bool redo = false, resolve_only = false;
    RET(resolve_pending(h, info, &redo, &resolve_only));
    if (redo || !info->cert_used) {
        // no proof: resolve_pending has already run the delta = 0 test itself; if it confirmed the tentative choice nothing
        // has to be redone, otherwise the sequential path decides
        if (redo) { h->delta = delta_in; h->n_fast_miss++; return 0; }
    }
    info->n_factor_phys = h->n_phys;
    for (int i = 0; i < 9; i++) stats[i] = h->h_red[i];
    for (int i = 0; i < 3; i++) trial[i] = h->h_red[20 + i];
    *done = 1;
    h->n_fast_ok++;
    return 0;
}

// nu rule (pyipm.py:1727-1735), step rules + line search (1737-1749), KKT at the new point (1754): everything of an inner
// iteration after the search direction h->dz is known
static int finish_step(Eng* h, b200ipm_step_info* info, const double* stats_in = nullptr, const double* first_trial = nullptr) {
    double stats[9];
    if (stats_in) { for (int i = 0; i < 9; i++) stats[i] = stats_in[i]; }
    else RET(dir_stats(h, stats));
    info->resid = stats_in ? stats_in[8] : h->h_red[8];   // ||b - K dz||_inf of the last refinement sweep (red[8])
    info->con_l1 = h->last_red[4];
    if (h->C) {
        // merit parameter update (pyipm.py:1727-1735); IEEE semantics for ||con||_1 == 0 are the reference's
        const double nu_thres = stats[0] / (1.0 - h->p.rho) / h->last_red[4];
        if (h->nu < nu_thres) h->nu = nu_thres;
    }
    RET(line_search(h, info, stats, first_trial));
    CU(cudaEventRecord(h->ev[EV_SEARCH], h->st));
    // KKT conditions at the new point (pyipm.py:1754); doubles as the residual of the next step
    if (h->kind != KIND_CALLABLE) {
        RET(residual(h));
        for (int i = 0; i < 4; i++) info->kkt_norm[i] = h->last_red[i];
        info->fval = h->last_red[5];
    }
    info->mu = h->mu; info->nu = h->nu; info->delta = h->delta;
    CU(cudaStreamSynchronize(h->st));
    fill_times(h, info);
    return 0;
}

int b200ipm_newton_step(b200ipm_handle h, b200ipm_step_info* info) {
    if (!h || !info) return fail_msg("null argument");
    CU(cudaSetDevice(h->device));
    memset(info, 0, sizeof(*info));
    CU(cudaEventRecord(h->ev[EV_START], h->st));
    {
        double stats[9], trial[3];
        int done = 0;
        RET(fast_step(h, info, stats, trial, &done));
        if (done) return finish_step(h, info, stats, trial);
    }
    RET(compute_direction(h, info));
    return finish_step(h, info);
}

// second-order correction direction (pyipm.py:1468-1477, 1520-1529) at the CURRENT state: dz_p = -lstsq(jaco(x)', c_new).
// Exposed for callable mode, where the host evaluates c_new = con(x0 + a dx, s0 + a ds) with the user's functions.
int b200ipm_soc_direction(b200ipm_handle h, const double* cnew, double* pz) {
    if (!h || !cnew || !pz) return fail_msg("null argument");
    if (!h->C) return fail_msg("soc_direction: the problem has no constraints");
    CU(cudaSetDevice(h->device));
    RET(eval_derivs(h));
    RET(up(h, h->cnew, cnew, h->C, 0));
    RET(soc_direction(h, h->cnew, h->pvec));
    RET(down(h, pz, h->pvec, h->D + h->N));
    CU(cudaStreamSynchronize(h->st));
    return 0;
}

// ------------------------------------------------------------------------------------------ L-BFGS ABI
int b200ipm_lbfgs_init(b200ipm_handle h, int m, double zeta) {
    if (!h || m <= 0 || !(zeta > 0.0)) return fail_msg("lbfgs_init: bad arguments");
    CU(cudaSetDevice(h->device));
    RET(lb_alloc(h, m));
    h->lb_zeta0 = zeta;
    lb_reset(h);
    CU(cudaMemcpyAsync(h->lb_xold, h->x, sizeof(double) * h->D, cudaMemcpyDeviceToDevice, h->st));
    return 0;
}
int b200ipm_lbfgs_update(b200ipm_handle h, const double* gradx_old) {
    if (!h || !h->lb_S) return fail_msg("lbfgs_update: call b200ipm_lbfgs_init first");
    CU(cudaSetDevice(h->device));
    return lb_update(h, gradx_old);
}
int b200ipm_lbfgs_direction(b200ipm_handle h, double* dz, b200ipm_step_info* info) {
    if (!h || !h->lb_S) return fail_msg("lbfgs_direction: call b200ipm_lbfgs_init first");
    CU(cudaSetDevice(h->device));
    b200ipm_step_info tmp{};
    if (!info) info = &tmp;
    memset(info, 0, sizeof(*info));
    CU(cudaEventRecord(h->ev[EV_START], h->st));
    RET(lb_direction(h, info));
    if (dz) { RET(down(h, dz, h->dz, h->K)); }
    CU(cudaStreamSynchronize(h->st));
    info->mu = h->mu; info->nu = h->nu;
    return 0;
}
int b200ipm_lbfgs_step(b200ipm_handle h, int do_update, b200ipm_step_info* info) {
    if (!h || !info) return fail_msg("null argument");
    if (!h->lb_S) return fail_msg("lbfgs_step: call b200ipm_lbfgs_init first");
    if (h->kind != KIND_QUAD && h->kind != KIND_POLY) return fail_msg("lbfgs_step needs a lowered problem (callable mode: update + direction)");
    CU(cudaSetDevice(h->device));
    memset(info, 0, sizeof(*info));
    CU(cudaEventRecord(h->ev[EV_START], h->st));
    if (do_update) RET(lb_update(h, nullptr));       // pyipm.py:1705-1710
    for (int e : {EV_EVAL, EV_ASSEMBLE, EV_HESS0, EV_HESS1, EV_COND0, EV_COND1}) CU(cudaEventRecord(h->ev[e], h->st));
    RET(lb_direction(h, info));                      // pyipm.py:1711-1713
    CU(cudaEventRecord(h->ev[EV_FACTOR], h->st));
    CU(cudaEventRecord(h->ev[EV_SOLVE], h->st));
    return finish_step(h, info);
}
int b200ipm_lbfgs_state(b200ipm_handle h, int* m, double* zeta, int* fail) {
    if (!h) return fail_msg("null handle");
    if (m) *m = h->lb_m;
    if (zeta) *zeta = h->lb_zeta;
    if (fail) *fail = h->lb_fail;
    return 0;
}

// ------------------------------------------------------------------------------------------ batched multi-start
int b200ipm_batch_solve_poly(int D, int M, int N, int nterms, const int* term_row, const double* term_coeff,
                             const int* term_ptr, const int* fac_var, const int* fac_pow, double xlogx_coeff,
                             double xlogx_shift, const b200ipm_params* p, int niter, int miter, int use_ftol, double Ftol,
                             int batch, const double* x0, int device, double* x, double* s, double* lda, double* fval,
                             double* kkt_norm, int* signal, int* iters, float* ms) {
    if (!p || !x0 || !x || batch <= 0 || D <= 0 || M < 0 || N < 0) return fail_msg("batch_solve_poly: bad arguments");
    const int K = D + 2 * N + M, C = M + N;
    if (K > BK_MAX) return fail_msg("batch_solve_poly: K = D + 2N + M must be <= 32 (one warp per instance)");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail_msg("b200ipm_batch_solve_poly: no CUDA device available (this library has no CPU fallback)");
    CU(cudaSetDevice(device));
    const int R = 1 + M + N;
    std::vector<int> rowptr(R + 1, 0);
    for (int t = 0; t < nterms; t++) {
        if (term_row[t] < 0 || term_row[t] >= R) return fail_msg("batch_solve_poly: term_row out of range");
        if (t > 0 && term_row[t] < term_row[t - 1]) return fail_msg("batch_solve_poly: terms must be sorted by row");
        rowptr[term_row[t] + 1]++;
    }
    for (int r = 0; r < R; r++) rowptr[r + 1] += rowptr[r];
    const int nfac = nterms ? term_ptr[nterms] : 0;
    int *d_rowptr = nullptr, *d_ptr = nullptr, *d_fvar = nullptr, *d_fpow = nullptr, *d_sig = nullptr, *d_it = nullptr;
    double *d_coeff = nullptr, *d_x0 = nullptr, *d_x = nullptr, *d_s = nullptr, *d_l = nullptr, *d_f = nullptr, *d_k = nullptr;
    RET(dalloc(&d_rowptr, R + 1)); RET(dalloc(&d_ptr, nterms + 1)); RET(dalloc(&d_fvar, nfac)); RET(dalloc(&d_fpow, nfac));
    RET(dalloc(&d_coeff, nterms));
    RET(dalloc(&d_x0, (size_t)batch * D)); RET(dalloc(&d_x, (size_t)batch * D)); RET(dalloc(&d_s, (size_t)batch * N));
    RET(dalloc(&d_l, (size_t)batch * C)); RET(dalloc(&d_f, batch)); RET(dalloc(&d_k, (size_t)batch * 4));
    RET(dalloc(&d_sig, batch)); RET(dalloc(&d_it, batch));
    std::vector<int> tp(nterms + 1, 0);
    for (int t = 0; t <= nterms; t++) tp[t] = nterms ? term_ptr[t] : 0;
    CU(cudaMemcpy(d_rowptr, rowptr.data(), sizeof(int) * (R + 1), cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_ptr, tp.data(), sizeof(int) * (nterms + 1), cudaMemcpyHostToDevice));
    if (nfac) {
        CU(cudaMemcpy(d_fvar, fac_var, sizeof(int) * nfac, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(d_fpow, fac_pow, sizeof(int) * nfac, cudaMemcpyHostToDevice));
    }
    if (nterms) CU(cudaMemcpy(d_coeff, term_coeff, sizeof(double) * nterms, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_x0, x0, sizeof(double) * (size_t)batch * D, cudaMemcpyHostToDevice));
    PolyData P{nterms, R, d_rowptr, d_coeff, d_ptr, d_fvar, d_fpow, xlogx_coeff, xlogx_shift};
    BatchParams bp{p->mu, p->nu, p->rho, p->tau, p->eta, p->beta, p->Ktol, Ftol, p->eps, p->reg_coef, niter, miter, use_ftol};
    CU(cudaFuncSetAttribute(batch_solve_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(BatchWs)));
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    CU(cudaEventRecord(e0, nullptr));
    batch_solve_kernel<<<batch, 32, sizeof(BatchWs), nullptr>>>(D, M, N, P, bp, batch, d_x0, d_x, d_s, d_l, d_f, d_k, d_sig, d_it);
    LAUNCHED();
    CU(cudaEventRecord(e1, nullptr));
    CU(cudaEventSynchronize(e1));
    if (ms) CU(cudaEventElapsedTime(ms, e0, e1));
    CU(cudaMemcpy(x, d_x, sizeof(double) * (size_t)batch * D, cudaMemcpyDeviceToHost));
    if (s && N) CU(cudaMemcpy(s, d_s, sizeof(double) * (size_t)batch * N, cudaMemcpyDeviceToHost));
    if (lda && C) CU(cudaMemcpy(lda, d_l, sizeof(double) * (size_t)batch * C, cudaMemcpyDeviceToHost));
    if (fval) CU(cudaMemcpy(fval, d_f, sizeof(double) * batch, cudaMemcpyDeviceToHost));
    if (kkt_norm) CU(cudaMemcpy(kkt_norm, d_k, sizeof(double) * (size_t)batch * 4, cudaMemcpyDeviceToHost));
    if (signal) CU(cudaMemcpy(signal, d_sig, sizeof(int) * batch, cudaMemcpyDeviceToHost));
    if (iters) CU(cudaMemcpy(iters, d_it, sizeof(int) * batch, cudaMemcpyDeviceToHost));
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaFree(d_rowptr); cudaFree(d_ptr); cudaFree(d_fvar); cudaFree(d_fpow); cudaFree(d_coeff); cudaFree(d_x0); cudaFree(d_x);
    cudaFree(d_s); cudaFree(d_l); cudaFree(d_f); cudaFree(d_k); cudaFree(d_sig); cudaFree(d_it);
    return 0;
}

// ------------------------------------------------------------------------------------------ generic LDL^T
int b200ipm_ldlt_create(int n, int device, void* stream, b200ipm_ldlt_handle* out) {
    if (!out || n <= 0) return fail_msg("ldlt_create: bad arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail_msg("b200ipm_ldlt_create: no CUDA device available (this library has no CPU fallback)");
    CU(cudaSetDevice(device));
    b200ipm_ldlt* h = new b200ipm_ldlt();
    h->device = device;
    if (stream) { h->st = (cudaStream_t)stream; } else { CU(cudaStreamCreateWithFlags(&h->st, cudaStreamNonBlocking)); h->own_stream = true; }
    RET(ldlt_init_attrs());
    RET(ldlt_init_solve_attrs());
    RET(ldlt_alloc(h->F, n, h->st));
    RET(dalloc(&h->A0, (size_t)n * h->F.ld));
    RET(dalloc(&h->b, n)); RET(dalloc(&h->x, n)); RET(dalloc(&h->r, n)); RET(dalloc(&h->c, n));
    *out = h;
    return 0;
}
int b200ipm_ldlt_destroy(b200ipm_ldlt_handle h) {
    if (!h) return 0;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->st);
    ldlt_free(h->F);
    oz_upd_free(h->ozd);
    cudaFree(h->ozd_err);
    cudaFree(h->A0); cudaFree(h->b); cudaFree(h->x); cudaFree(h->r); cudaFree(h->c);
    if (h->own_stream) cudaStreamDestroy(h->st);
    delete h;
    return 0;
}
int b200ipm_ldlt_factor(b200ipm_ldlt_handle h, const double* A, int lda, int on_device, int inertia[3], double* rcond_est) {
    if (!h || !A) return fail_msg("null argument");
    CU(cudaSetDevice(h->device));
    const int n = h->F.n, ld = h->F.ld;
    CU(cudaMemcpy2DAsync(h->A0, sizeof(double) * ld, A, sizeof(double) * lda, sizeof(double) * n, n,
                         on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, h->st));
    dim3 blk(32, 8), grid(cdiv(n, 32), cdiv(n, 8));
    sym_from_lower_kernel<<<grid, blk, 0, h->st>>>(h->A0, ld, n);
    LAUNCHED();
    CU(cudaMemcpyAsync(h->F.A, h->A0, sizeof(double) * (size_t)n * ld, cudaMemcpyDeviceToDevice, h->st));
    RET(ldlt_factor(h->F));
    int cnt[4];
    double ds[2];
    CU(cudaMemcpyAsync(cnt, h->F.counts, sizeof(int) * 4, cudaMemcpyDeviceToHost, h->st));
    CU(cudaMemcpyAsync(ds, h->F.dstat, sizeof(double) * 2, cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    if (inertia) { inertia[0] = cnt[2]; inertia[1] = cnt[0]; inertia[2] = cnt[1]; }
    if (rcond_est) *rcond_est = (cnt[1] > 0 || !(ds[1] > 0.0)) ? 0.0 : ds[0] / ds[1];
    h->factored = true;
    return 0;
}
int b200ipm_ldlt_solve(b200ipm_ldlt_handle h, double* B, int nrhs, int nrefine, int on_device) {
    if (!h || !B) return fail_msg("null argument");
    if (!h->factored) return fail_msg("ldlt_solve: factor first");
    CU(cudaSetDevice(h->device));
    const int n = h->F.n, ld = h->F.ld;
    const cudaMemcpyKind kin = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    const cudaMemcpyKind kout = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost;
    for (int r = 0; r < nrhs; r++) {
        CU(cudaMemcpyAsync(h->b, B + (size_t)r * n, sizeof(double) * n, kin, h->st));
        RET(ldlt_solve(h->F, h->b, h->x));
        for (int it = 0; it < nrefine; it++) {
            RET(gemv_n(h->st, h->A0, ld, n, n, h->x, h->b, 1.0, -1.0, h->r));   // r = b - A x
            RET(ldlt_solve(h->F, h->r, h->c));
            axpby_kernel<<<cdiv(n, 256), 256, 0, h->st>>>(n, 1.0, h->x, 1.0, h->c, h->x);
            LAUNCHED();
        }
        CU(cudaMemcpyAsync(B + (size_t)r * n, h->x, sizeof(double) * n, kout, h->st));
    }
    CU(cudaStreamSynchronize(h->st));
    return 0;
}
int b200ipm_ldlt_tile_factor(b200ipm_ldlt_handle h, double* tile_dev, int ld, int nb, double* linv_dev, double* dblk_dev,
                             int* perm_dev, int counts[3]) {
    if (!h || !tile_dev || !linv_dev || !dblk_dev || nb <= 0 || nb > NB) return fail_msg("tile_factor: bad arguments");
    CU(cudaSetDevice(h->device));
    // dblk_dev layout: [dinv_a (NB) | dinv_b (NB) | d_a (NB) | d_b (NB)] followed by NB ints of `kind`.
    // The inertia counts ACCUMULATE in the handle (device side, no synchronisation) until a call passes a non-NULL
    // `counts`, which returns the totals since the previous such call and resets them.
    int* kind = reinterpret_cast<int*>(dblk_dev + 4 * NB);
    if (!h->tile_counts_live) {
        ldlt_reset_kernel<<<1, 1, 0, h->st>>>(h->F.counts, h->F.dstat, h->F.ticket);
        LAUNCHED();
        h->tile_counts_live = true;
    }
    ldlt_tile_kernel<<<1, TILE_THREADS, TILE_SMEM, h->st>>>(tile_dev, ld, nb, linv_dev, dblk_dev, dblk_dev + NB, dblk_dev + 2 * NB,
                                                   dblk_dev + 3 * NB, kind, perm_dev, h->F.counts, h->F.dstat, h->F.pivot_u, nullptr, h->F.tile_blocked);
    LAUNCHED();
    if (counts) {
        int cnt[4];
        CU(cudaMemcpyAsync(cnt, h->F.counts, sizeof(int) * 4, cudaMemcpyDeviceToHost, h->st));
        CU(cudaStreamSynchronize(h->st));
        counts[0] = cnt[2]; counts[1] = cnt[0]; counts[2] = cnt[1];
        h->tile_counts_live = false;
    }
    return 0;
}
int b200ipm_ldlt_panel(b200ipm_ldlt_handle h, double* panel_dev, int ld, int rows, const double* linv_dev,
                       const double* dblk_dev, const int* perm_dev, double* w_dev, int ldw) {
    (void)perm_dev;
    if (!h || !panel_dev || !w_dev) return fail_msg("panel: bad arguments");
    if (rows <= 0) return 0;
    CU(cudaSetDevice(h->device));
    const int* kind = reinterpret_cast<const int*>(dblk_dev + 4 * NB);
    ldlt_panel_kernel<<<cdiv(rows, NB), 128, PANEL_SMEM, h->st>>>(panel_dev, ld, rows, linv_dev, dblk_dev, dblk_dev + NB,
                                                                   kind, w_dev, ldw, nullptr);
    LAUNCHED();
    return 0;
}
int b200ipm_ldlt_import(b200ipm_ldlt_handle h, const double* A_dev, int lda, const double* linvp_dev,
                        const double* dinfo_dev, const int* kind_dev) {
    if (!h || !A_dev || !linvp_dev || !dinfo_dev || !kind_dev) return fail_msg("ldlt_import: null argument");
    CU(cudaSetDevice(h->device));
    const int n = h->F.n, ld = h->F.ld;
    const size_t npad = (size_t)h->F.nblk * NB;
    CU(cudaMemcpy2DAsync(h->F.A, sizeof(double) * ld, A_dev, sizeof(double) * lda, sizeof(double) * n, n,
                         cudaMemcpyDeviceToDevice, h->st));
    CU(cudaMemcpyAsync(h->F.LinvP, linvp_dev, sizeof(double) * (size_t)h->F.nblk * NB * NB, cudaMemcpyDeviceToDevice, h->st));
    CU(cudaMemcpyAsync(h->F.dinfo, dinfo_dev, sizeof(double) * 4 * npad, cudaMemcpyDeviceToDevice, h->st));
    CU(cudaMemcpyAsync(h->F.kind, kind_dev, sizeof(int) * npad, cudaMemcpyDeviceToDevice, h->st));
    if (h->F.solve256) RET(ldlt_blockinv_launch(h->F, h->st, 0, h->F.nb256));    // inverses of the 256-row diagonal blocks
    CU(cudaStreamSynchronize(h->st));
    h->factored = true;
    return 0;
}
int b200ipm_gemm_nt_update(b200ipm_ldlt_handle h, double* C_dev, int ldc, int rows, int cols, const double* A_dev, int lda,
                           const double* B_dev, int ldb, int k, int lower_only) {
    if (!h || !C_dev || !A_dev || !B_dev) return fail_msg("gemm_nt_update: bad arguments");
    if (rows <= 0 || cols <= 0) return 0;
    CU(cudaSetDevice(h->device));
    if (!lower_only && cols <= 512) return gemm_nt_sub(h->st, C_dev, ldc, rows, cols, A_dev, lda, B_dev, ldb, k);
    GemmArgs u{};
    u.C = C_dev; u.ldc = ldc; u.Cin = C_dev; u.ldcin = ldc; u.n = rows; u.m = cols; u.beta = 1.0;
    u.mode = lower_only ? GEMM_LOWER_ONLY : GEMM_FULL; u.nterms = 1;
    u.t[0] = GemmTerm{A_dev, B_dev, nullptr, lda, ldb, k, -1.0};
    RET(gemm_nt(h->st, u));
    return 0;
}

// Composite calls of the block-cyclic driver: one C call per block column instead of a dozen (the host-side launch
// sequencing of a column was the critical path of the multi-GPU factorisation).  diag_dev layout (CudaTileOps): [b*b copy
// of the factored block | per 64-tile: LinvP (64*64) | dinv_a | dinv_b | d_a | d_b (64 each) | kind (64 ints = 32 doubles)].
int b200ipm_ldlt_block_factor(b200ipm_ldlt_handle h, double* A_dev, int ld, int b, double* diag_dev, double* wdiag_dev) {
    if (!h || !A_dev || !diag_dev || !wdiag_dev || b <= 0 || (b % NB) != 0) return fail_msg("block_factor: bad arguments");
    CU(cudaSetDevice(h->device));
    const int nt = b / NB, tile_doubles = NB * NB + 4 * NB + NB / 2;
    if (!h->tile_counts_live) {
        ldlt_reset_kernel<<<1, 1, 0, h->st>>>(h->F.counts, h->F.dstat, h->F.ticket);
        LAUNCHED();
        h->tile_counts_live = true;
    }
    for (int t = 0; t < nt; t++) {
        const int k0 = t * NB;
        double* linv = diag_dev + (size_t)b * b + (size_t)t * tile_doubles;
        double* dblk = linv + NB * NB;
        int* kind = reinterpret_cast<int*>(dblk + 4 * NB);
        ldlt_tile_kernel<<<1, TILE_THREADS, TILE_SMEM, h->st>>>(A_dev + (size_t)k0 * ld + k0, ld, NB, linv, dblk, dblk + NB, dblk + 2 * NB,
                                                       dblk + 3 * NB, kind, nullptr, h->F.counts, h->F.dstat, h->F.pivot_u, nullptr,
                                                       h->F.tile_blocked);
        LAUNCHED();
        const int rows = b - k0 - NB;
        if (rows > 0) {
            double* pp = A_dev + (size_t)(k0 + NB) * ld + k0;
            double* wp = wdiag_dev + (size_t)(k0 + NB) * b + k0;
            ldlt_panel_kernel<<<cdiv(rows, NB), 128, PANEL_SMEM, h->st>>>(pp, ld, rows, linv, dblk, dblk + NB, kind, wp, b, nullptr);
            LAUNCHED();
            GemmArgs u{};
            u.C = A_dev + (size_t)(k0 + NB) * ld + (k0 + NB); u.ldc = ld; u.Cin = u.C; u.ldcin = ld; u.n = rows; u.m = rows; u.beta = 1.0;
            u.mode = GEMM_LOWER_ONLY; u.nterms = 1;
            u.t[0] = GemmTerm{wp, pp, nullptr, b, ld, NB, -1.0};
            RET(gemm_nt(h->st, u));
        }
    }
    CU(cudaMemcpy2DAsync(diag_dev, sizeof(double) * b, A_dev, sizeof(double) * ld, sizeof(double) * b, b, cudaMemcpyDeviceToDevice, h->st));
    return 0;
}
// B_dev (rows x b, leading dimension ld) <- L of the panel, W_dev (rows x b, contiguous) <- W = L D, given the factor
// data of the diagonal block
int b200ipm_ldlt_block_panel(b200ipm_ldlt_handle h, double* B_dev, int ld, int rows, int b, const double* diag_dev, double* W_dev) {
    if (!h || !B_dev || !diag_dev || !W_dev || b <= 0 || (b % NB) != 0) return fail_msg("block_panel: bad arguments");
    if (rows <= 0) return 0;
    CU(cudaSetDevice(h->device));
    const int nt = b / NB, tile_doubles = NB * NB + 4 * NB + NB / 2;
    for (int t = 0; t < nt; t++) {
        const int k0 = t * NB;
        const double* linv = diag_dev + (size_t)b * b + (size_t)t * tile_doubles;
        const double* dblk = linv + NB * NB;
        const int* kind = reinterpret_cast<const int*>(dblk + 4 * NB);
        ldlt_panel_kernel<<<cdiv(rows, NB), 128, PANEL_SMEM, h->st>>>(B_dev + k0, ld, rows, linv, dblk, dblk + NB, kind, W_dev + k0, b, nullptr);
        LAUNCHED();
        const int cols = b - k0 - NB;
        if (cols > 0)   // remaining columns of this block column: B[:, k0+64:] -= W_t * L_kk[k0+64:, k0:k0+64]^T
            RET(gemm_nt_sub(h->st, B_dev + k0 + NB, ld, rows, cols, W_dev + k0, b, diag_dev + (size_t)(k0 + NB) * b + k0, b, NB));
    }
    return 0;
}

// One block column of the block-column-cyclic driver in ONE call: A_dev = top-left corner of the (rows_total x b) block
// column (b = 256), factored with the single-GPU factorisation's own panel schedule (tile -> 4-CTA mini step on the chain,
// panel rows + in-panel updates on the workspace's update stream), L in place, W = L D to Wb_dev (rows_total x b), factor
// data of the diagonal block packed into diag_dev (layout of block_factor).  The handle must have order >= b.
__global__ void diag_pack_kernel(const double* __restrict__ A, int ld, int b, const double* __restrict__ LinvP,
                                 const double* __restrict__ dinfo, int npad, const int* __restrict__ kind, double* __restrict__ diag) {
    const int t = blockIdx.x, tid = threadIdx.x;
    const int tile_doubles = NB * NB + 4 * NB + NB / 2;
    double* linv = diag + (size_t)b * b + (size_t)t * tile_doubles;
    double* dblk = linv + NB * NB;
    int* kd = reinterpret_cast<int*>(dblk + 4 * NB);
    for (int i = tid; i < NB * NB; i += blockDim.x) linv[i] = LinvP[(size_t)t * NB * NB + i];
    for (int i = tid; i < 4 * NB; i += blockDim.x) dblk[i] = dinfo[(size_t)(i / NB) * npad + t * NB + (i % NB)];
    for (int i = tid; i < NB; i += blockDim.x) kd[i] = kind[t * NB + i];
    for (int i = tid; i < NB * b; i += blockDim.x) {       // rows t*64 .. t*64+63 of the factored diagonal block
        const int r = t * NB + i / b, c = i % b;
        diag[(size_t)r * b + c] = A[(size_t)r * ld + c];
    }
}
int b200ipm_ldlt_colblock_factor(b200ipm_ldlt_handle h, double* A_dev, int ld, int rows_total, int b, double* diag_dev,
                                 double* Wb_dev, void* rest_ready_event) {
    if (!h || !A_dev || !diag_dev || !Wb_dev || b != NBO || rows_total < b || (rows_total % NB) != 0)
        return fail_msg("colblock_factor: bad arguments");
    if (h->F.n < b) return fail_msg("colblock_factor: the handle's order must be at least the block size");
    if ((ld & 1) || (reinterpret_cast<uintptr_t>(A_dev) & 15) || (reinterpret_cast<uintptr_t>(Wb_dev) & 15))
        return fail_msg("colblock_factor: operands must be 16-byte aligned");
    CU(cudaSetDevice(h->device));
    LdltWs& F = h->F;
    double* A_save = F.A;
    const int ld_save = F.ld, n_save = F.n;
    F.A = A_dev; F.ld = ld; F.n = rows_total;
    F.col_rest_event = reinterpret_cast<cudaEvent_t>(rest_ready_event);
    const int rc = ldlt_factor_launch(F, h->st, Wb_dev);
    F.A = A_save; F.ld = ld_save; F.n = n_save;
    F.col_rest_event = nullptr;
    RET(rc);
    const int npad = F.nblk * NB;
    diag_pack_kernel<<<b / NB, 256, 0, h->st>>>(A_dev, ld, b, F.LinvP, F.dinfo, npad, F.kind, diag_dev);
    LAUNCHED();
    return 0;
}

// tcgen05 trailing updates for the block-column-cyclic driver (the same int8 error-free split as inside the single-GPU
// factorisation: 21 slice pairs, in place on the lower trapezoid).  panel_slice: digits of W (rows x 256) and of -L (rows x
// 256) of the current panel, ONCE per panel; block_update: C (n x ncols lower trapezoid with its origin on the diagonal, the
// piece starting row_off rows below the first sliced row) -= W L' from those digits.  max_rows sizes the workspace at the
// first call.  A non-finite operand raises the handle's error word (b200ipm_oz_status).
int b200ipm_oz_panel_slice(b200ipm_ldlt_handle h, int rows, const double* W_dev, int ldw, const double* L_dev, int ldl, int max_rows) {
    if (!h || !W_dev || !L_dev || rows <= 0) return fail_msg("oz_panel_slice: bad arguments");
    CU(cudaSetDevice(h->device));
    if (h->ozd.nmax < rows) {
        CU(cudaStreamSynchronize(h->st));
        oz_upd_free(h->ozd);
        RET(oz_upd_alloc(h->ozd, std::max(rows, max_rows), NBO));
        if (!h->ozd_err) { RET(dalloc(&h->ozd_err, 1)); CU(cudaMemset(h->ozd_err, 0, sizeof(int))); }
    }
    return oz_panel_slice(h->st, rows, W_dev, ldw, L_dev, ldl, NBO, h->ozd, 0, h->ozd_err, nullptr);
}
int b200ipm_oz_block_update(b200ipm_ldlt_handle h, double* C_dev, int ldc, int n, int ncols, int row_off) {
    if (!h || !C_dev || n <= 0 || ncols <= 0 || (row_off % OZ_BM) != 0 || (n % OZ_BM) != 0) return fail_msg("oz_block_update: bad arguments");
    if (h->ozd.nmax < n + row_off) return fail_msg("oz_block_update: slice the panel first");
    CU(cudaSetDevice(h->device));
    RET(oz_upd_tiles(h->ozd, n, ncols));
    return oz_panel_update(h->st, C_dev, ldc, n, ncols, row_off / OZ_BM, NBO, h->ozd, 0, h->ozd_err, nullptr, 0);
}
int b200ipm_oz_status(b200ipm_ldlt_handle h, int* err_word) {
    if (!h || !err_word) return fail_msg("oz_status: bad arguments");
    *err_word = 0;
    if (!h->ozd_err) return 0;
    CU(cudaSetDevice(h->device));
    CU(cudaMemcpyAsync(err_word, h->ozd_err, sizeof(int), cudaMemcpyDeviceToHost, h->st));
    CU(cudaStreamSynchronize(h->st));
    if (*err_word) CU(cudaMemsetAsync(h->ozd_err, 0, sizeof(int), h->st));
    return 0;
}

// y (rows) = A (rows x cols, row-major, leading dimension lda) * v (cols): the HBM-bound GEMV kernel of the residual,
// for the distributed refinement mat-vec of the block-cyclic driver (device pointers, asynchronous on the handle's stream)
int b200ipm_ldlt_gemv(b200ipm_ldlt_handle h, const double* A_dev, int lda, int rows, int cols, const double* v_dev,
                      double* y_dev) {
    if (!h || !A_dev || !v_dev || !y_dev) return fail_msg("ldlt_gemv: bad arguments");
    if (rows <= 0 || cols <= 0) return 0;
    CU(cudaSetDevice(h->device));
    return gemv_n(h->st, A_dev, lda, rows, cols, v_dev, nullptr, 0.0, 1.0, y_dev);
}

int b200ipm_gemm_nt_update_bc(b200ipm_ldlt_handle h, double* C_dev, int ldc, int rows, int cols, const double* A_dev,
                              int lda, const double* B_dev, int ldb, int k, int block, int P, int Q, int p, int q, int li0,
                              int lj0) {
    if (!h || !C_dev || !A_dev || !B_dev || block <= 0 || (block % G_BM) != 0) return fail_msg("gemm_nt_update_bc: bad arguments");
    if (rows <= 0 || cols <= 0) return 0;
    CU(cudaSetDevice(h->device));
    GemmArgs u{};
    u.C = C_dev; u.ldc = ldc; u.Cin = C_dev; u.ldcin = ldc; u.n = rows; u.m = cols; u.beta = 1.0;
    u.mode = GEMM_BC_LOWER; u.nterms = 1;
    u.t[0] = GemmTerm{A_dev, B_dev, nullptr, lda, ldb, k, -1.0};
    u.bc_b = block; u.bc_P = P; u.bc_Q = Q; u.bc_p = p; u.bc_q = q; u.bc_li0 = li0; u.bc_lj0 = lj0;
    RET(gemm_nt(h->st, u));
    return 0;
}

// ------------------------------------------------------------------------------------------ test hooks
int b200ipm_test_syrk(int n, const double* Cin, double beta, const double* dadd, double shift, int nterms,
                      const double* const* A, const double* const* w, const int* K, const double* alpha, double* C,
                      int force_simple, float* ms) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("no CUDA device available");
    if (nterms < 0 || nterms > 3) return fail_msg("test_syrk: 0..3 terms");
    RET(ldlt_init_attrs());
    cudaStream_t st = nullptr;
    double *dC = nullptr, *dCin = nullptr, *dd = nullptr, *dA[3] = {nullptr, nullptr, nullptr}, *dw[3] = {nullptr, nullptr, nullptr};
    const int ld = (int)rup(n, 16);
    RET(dalloc(&dC, (size_t)n * ld));
    GemmArgs a{};
    a.C = dC; a.ldc = ld; a.n = n; a.m = n; a.beta = beta; a.shift = shift; a.mode = GEMM_UPPER_MIRROR; a.nterms = nterms;
    if (Cin) {
        RET(dalloc(&dCin, (size_t)n * ld));
        CU(cudaMemcpy2D(dCin, sizeof(double) * ld, Cin, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice));
        a.Cin = dCin; a.ldcin = ld;
    }
    if (dadd) { RET(dalloc(&dd, n)); CU(cudaMemcpy(dd, dadd, sizeof(double) * n, cudaMemcpyHostToDevice)); a.dadd = dd; }
    for (int t = 0; t < nterms; t++) {
        const int ldk = (int)rup(K[t], 16);
        RET(dalloc(&dA[t], (size_t)n * ldk));
        CU(cudaMemcpy2D(dA[t], sizeof(double) * ldk, A[t], sizeof(double) * K[t], sizeof(double) * K[t], n, cudaMemcpyHostToDevice));
        if (w && w[t]) { RET(dalloc(&dw[t], K[t])); CU(cudaMemcpy(dw[t], w[t], sizeof(double) * K[t], cudaMemcpyHostToDevice)); }
        a.t[t] = GemmTerm{dA[t], dA[t], dw[t], ldk, ldk, K[t], alpha[t]};
    }
    cudaEvent_t e0, e1;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    RET(gemm_nt(st, a, force_simple != 0));   // warm-up
    CU(cudaEventRecord(e0, st));
    RET(gemm_nt(st, a, force_simple != 0));
    CU(cudaEventRecord(e1, st));
    CU(cudaEventSynchronize(e1));
    if (ms) CU(cudaEventElapsedTime(ms, e0, e1));
    CU(cudaMemcpy2D(C, sizeof(double) * n, dC, sizeof(double) * ld, sizeof(double) * n, n, cudaMemcpyDeviceToHost));
    cudaFree(dC); cudaFree(dCin); cudaFree(dd);
    for (int t = 0; t < 3; t++) { cudaFree(dA[t]); cudaFree(dw[t]); }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    return 0;
}
static TraceRec* g_trace_host_ptr = nullptr;
int b200ipm_trace_start(void) {
    if (!g_trace_host_ptr) CU(cudaMalloc(&g_trace_host_ptr, sizeof(TraceRec) * TRACE_CAP));
    CU(cudaMemset(g_trace_host_ptr, 0, sizeof(TraceRec) * TRACE_CAP));
    const int zero = 0;
    CU(cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(int)));
    CU(cudaMemcpyToSymbol(g_trace, &g_trace_host_ptr, sizeof(TraceRec*)));
    return 0;
}
int b200ipm_trace_dump(int* id, int* blk, unsigned long long* t0, unsigned long long* t1, unsigned long long* tag, int max,
                       int* n) {
    CU(cudaDeviceSynchronize());
    TraceRec* nullp = nullptr;
    CU(cudaMemcpyToSymbol(g_trace, &nullp, sizeof(TraceRec*)));
    int cnt = 0;
    CU(cudaMemcpyFromSymbol(&cnt, g_trace_n, sizeof(int)));
    cnt = std::min(cnt, std::min(max, TRACE_CAP));
    std::vector<TraceRec> h(std::max(cnt, 1));
    if (cnt && g_trace_host_ptr) CU(cudaMemcpy(h.data(), g_trace_host_ptr, sizeof(TraceRec) * cnt, cudaMemcpyDeviceToHost));
    for (int i = 0; i < cnt; i++) {
        id[i] = h[i].id; blk[i] = h[i].blk; t0[i] = h[i].t0; t1[i] = h[i].t1;
        if (tag) tag[i] = h[i].tag;
    }
    if (n) *n = cnt;
    return 0;
}
// Same contract as b200ipm_test_syrk, computed by the tcgen05 int8 Ozaki path (ozaki_i8.cuh).  variant: 0 = 128x64 tiles,
// one pass; 1 = 128x128 tiles, two passes.  lbo/sbo <= 0 keep the default descriptor strides.  ms[0] = slicing kernel,
// ms[1] = tensor-core kernel; *err = device error word (1 non-finite, 2 negative weight w/o sign operand, 4 timeout).
int b200ipm_test_syrk_i8(int n, const double* Cin, double beta, const double* dadd, double shift, int nterms,
                         const double* const* A, const double* const* w, const int* K, const double* alpha, double* C,
                         unsigned signed_mask, int variant, int lbo, int sbo, float* ms, int* err) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("no CUDA device available");
    if (nterms < 0 || nterms > 3) return fail_msg("test_syrk_i8: 0..3 terms");
    cudaStream_t st = nullptr;
    double *dC = nullptr, *dCin = nullptr, *dd = nullptr, *dA[3] = {nullptr, nullptr, nullptr}, *dw[3] = {nullptr, nullptr, nullptr};
    const int ld = (int)rup(n, 16);
    RET(dalloc(&dC, (size_t)n * ld));
    CU(cudaMemset(dC, 0xff, sizeof(double) * (size_t)n * ld));   // NaN pattern: untouched elements are visible
    GemmArgs a{};
    a.C = dC; a.ldc = ld; a.n = n; a.m = n; a.beta = beta; a.shift = shift; a.mode = GEMM_UPPER_MIRROR; a.nterms = nterms;
    if (Cin) {
        RET(dalloc(&dCin, (size_t)n * ld));
        CU(cudaMemcpy2D(dCin, sizeof(double) * ld, Cin, sizeof(double) * n, sizeof(double) * n, n, cudaMemcpyHostToDevice));
        a.Cin = dCin; a.ldcin = ld;
    }
    if (dadd) { RET(dalloc(&dd, n)); CU(cudaMemcpy(dd, dadd, sizeof(double) * n, cudaMemcpyHostToDevice)); a.dadd = dd; }
    for (int t = 0; t < nterms; t++) {
        const int ldk = (int)rup(K[t], 16);
        RET(dalloc(&dA[t], (size_t)n * ldk));
        CU(cudaMemcpy2D(dA[t], sizeof(double) * ldk, A[t], sizeof(double) * K[t], sizeof(double) * K[t], n, cudaMemcpyHostToDevice));
        if (w && w[t]) { RET(dalloc(&dw[t], K[t])); CU(cudaMemcpy(dw[t], w[t], sizeof(double) * K[t], cudaMemcpyHostToDevice)); }
        a.t[t] = GemmTerm{dA[t], dA[t], dw[t], ldk, ldk, K[t], alpha[t]};
    }
    OzWs ws;
    ws.variant = variant & 15;
    if (variant >> 4) ws.ndiag = variant >> 4;     // 128x128 tiles: slice-pair diagonals kept (6, 7, 8)
    if (lbo > 0) ws.lbo = lbo;
    if (sbo > 0) ws.sbo = sbo;
    for (int i = 0; i < 3; i++) CU(cudaEventCreate(&ws.ev[i]));
    cudaEvent_t e0, e1, e2;
    CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1)); CU(cudaEventCreate(&e2));
    int rc = oz_syrk(st, a, ws, signed_mask);   // warm-up (allocations, attributes)
    if (rc == 0 && cudaStreamSynchronize(st) != cudaSuccess) rc = fail_msg("test_syrk_i8: kernel failed");
    if (rc == 0) {
        // timed: the two kernels separately
        CU(cudaEventRecord(e0, st));
        rc = oz_syrk(st, a, ws, signed_mask);
        CU(cudaEventRecord(e2, st));
        CU(cudaEventSynchronize(e2));
        float t = 0.f, tsl = 0.f;
        CU(cudaEventElapsedTime(&t, e0, e2));
        if (rc == 0) CU(cudaEventElapsedTime(&tsl, ws.ev[0], ws.ev[1]));
        if (ms) { ms[0] = tsl; ms[1] = t; }
    }
    if (rc == 0) {
        if (err) CU(cudaMemcpy(err, ws.err, sizeof(int), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy2D(C, sizeof(double) * n, dC, sizeof(double) * ld, sizeof(double) * n, n, cudaMemcpyDeviceToHost));
    }
    oz_free(ws);
    cudaFree(dC); cudaFree(dCin); cudaFree(dd);
    for (int t = 0; t < 3; t++) { cudaFree(dA[t]); cudaFree(dw[t]); }
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return rc;
}
int b200ipm_test_gemv(int rows, int cols, const double* A, const double* v, double* y, int transpose_) {
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail_msg("no CUDA device available");
    double *dA = nullptr, *dv = nullptr, *dy = nullptr, *scr = nullptr;
    const int nin = transpose_ ? rows : cols, nout = transpose_ ? cols : rows;
    RET(dalloc(&dA, (size_t)rows * cols)); RET(dalloc(&dv, nin)); RET(dalloc(&dy, nout));
    RET(dalloc(&scr, gemv_t_scratch_doubles(rows, cols)));
    CU(cudaMemcpy(dA, A, sizeof(double) * rows * cols, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(dv, v, sizeof(double) * nin, cudaMemcpyHostToDevice));
    if (transpose_) RET(gemv_t(nullptr, dA, cols, rows, cols, dv, nullptr, 0.0, 1.0, dy, scr));
    else RET(gemv_n(nullptr, dA, cols, rows, cols, dv, nullptr, 0.0, 1.0, dy));
    CU(cudaMemcpy(y, dy, sizeof(double) * nout, cudaMemcpyDeviceToHost));
    cudaFree(dA); cudaFree(dv); cudaFree(dy); cudaFree(scr);
    return 0;
}

}  // extern "C"
