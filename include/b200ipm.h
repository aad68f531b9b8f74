/*
 * b200ipm.h -- C ABI of libb200ipm.so: the B200-native (sm_100a) Newton-step engine that replaces the
 * per-iteration hot path of jkaardal/pyipm (reference: /root/reference/pyipm.py @ ccc74da).
 *
 * The reference has no FFI: its "operator interface" is the set of compiled-callable slots that
 * IPM.compile() assigns (pyipm.py:855-954) and IPM.solve() consumes (pyipm.py:1658-1814).  Each entry point
 * below names the slot / code region it replaces.  INTEGRATION.md shows the ctypes binding a maintainer of
 * the reference would add (pyipm_b200/_lib.py is that binding, in full).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error (b200ipm_last_error() has the text); no exceptions
 *     and no C++/torch types cross this boundary;
 *   - a handle owns ALL its device memory and one CUDA stream (the one given at create time, or its own);
 *     a handle is not thread-safe, distinct handles are independent;
 *   - all host arrays are caller-owned, C-order float64, copied in/out;  `on_device != 0` on the bind/set
 *     calls means the pointers are device pointers on the handle's device (copied device-to-device);
 *   - vector ordering is the reference's: z = [x (D) | s (N) | lda_e (M) | lda_i (N)]  (pyipm.py:655-668),
 *     Jacobians are D x M / D x N ("transposed", pyipm.py:117,138,223-225), lda = [lda_e | lda_i];
 *   - the product never falls back to the CPU: if no CUDA device is present create() fails.
 */
#ifndef B200IPM_H
#define B200IPM_H

#ifdef __cplusplus
extern "C" {
#endif

#define B200IPM_VERSION 103

typedef struct b200ipm_engine* b200ipm_handle;
typedef struct b200ipm_ldlt*   b200ipm_ldlt_handle;

/* Solver constants: IPM.__init__ kwargs (pyipm.py:311-314, defaults :161-212) + engine knobs. */
typedef struct b200ipm_params {
    double mu, nu, rho, tau, eta, beta;   /* pyipm.py:338-343 */
    double Xtol, Ktol;                    /* pyipm.py:346-350 */
    double eps;                           /* np.finfo(float64).eps, pyipm.py:336 */
    double reg_coef;                      /* sqrt(eps) = delta0, pyipm.py:353,372 */
    int    nrefine;                       /* iterative-refinement sweeps against the UNREDUCED KKT residual */
    int    ls_batch;                      /* speculative line-search trials evaluated per launch */
    int    max_reg_retries;               /* cap on the delta*=10 loop (pyipm.py:1399-1403 has none) */
    int    flags;                         /* B200IPM_FLAG_* bits */
} b200ipm_params;
#define B200IPM_FLAG_NO_SPECULATION 1
#define B200IPM_FLAG_DELAY_BG       32  /* start the background inertia test only when the foreground factorisation is past
                                           its first third (measured: no net gain at config 3, off by default) */
#define B200IPM_FLAG_TCGEN05_FULLCOND 128 /* keep all 34 slice pairs in both tcgen05 products (default: 28 for d2L = 1.4e-15,
                                            21 for the condensation, which is only factored) */
#define B200IPM_FLAG_NO_CERT        64  /* never replace the delta = 0 inertia test by a negative-curvature certificate
                                           (then the test itself runs in the background) */
#define B200IPM_FLAG_NO_ABANDON     16  /* always complete a failed inertia test (n_neg_first is then the full count) */
#define B200IPM_FLAG_SLOW_STEP      256 /* b200ipm_newton_step never takes the one-sync path (every decision read back when it
                                           is taken, as b200ipm_direction does): A/B knob, identical results */
#define B200IPM_FLAG_TCGEN05_SYRK   2   /* d2L and condensation contractions on tcgen05 (int8 error-free split) */
#define B200IPM_FLAG_TCGEN05_TILE(v) ((v) << 2)   /* with TCGEN05_SYRK: 0 = 128x64 tiles / 1 pass, 1 = 128x128 / 2, 2 = 128x256 / 4 */

/* Everything one inner iteration (pyipm.py:1714-1754) reports back. */
typedef struct b200ipm_step_info {
    double kkt_norm[4];      /* 2-norms of (kkt1, kkt2, kkt3, kkt4) at the NEW point (pyipm.py:1754, 958-991) */
    double fval;             /* f at the new x (pyipm.py:1758,1778) */
    double delta;            /* diagonal shift after reghess (pyipm.py:1390-1402) */
    double mu, nu;           /* barrier / merit parameters after the step (nu: pyipm.py:1727-1735) */
    double alpha_smax, alpha_lmax;   /* fraction-to-the-boundary limits (pyipm.py:1739-1742) */
    double alpha_s, alpha_l;         /* accepted step lengths (pyipm.py:1507-1510,1553-1562) */
    double alpha_corr;               /* second-order-correction scaling, 0 if none */
    double phi0, dphi0;              /* merit value / directional derivative at the old point */
    double rcond;                    /* reciprocal-condition estimate used for the pyipm.py:1381 test (-1: the test was
                                        replaced by a certificate, see cert_used) */
    double resid;                    /* || b - K dz ||_inf of the unreduced system after refinement */
    double con_l1;                   /* ||con||_1 at the old point (pyipm.py:1732) */
    int    n_neg, n_zero;            /* inertia of the accepted condensed KKT matrix */
    int    n_factor;                 /* factorisations this step (= eigvalsh calls in the reference) */
    int    n_backtracks;             /* tau-multiplications in search() (pyipm.py:1492-1505) */
    int    soc_tried, soc_accepted;  /* pyipm.py:1464-1489 / 1516-1536 */
    int    signal;                   /* 0 ok, -2 bad direction (pyipm.py:1502,1548) */
    int    eq_reg;                   /* 1 if the eq-multiplier block was regularised (pyipm.py:1383-1389) */
    int    n_neg_first, n_zero_first;/* inertia of the FIRST (delta = 0) attempt, i.e. what pyipm.py:1381 tests */
    float  ms_eval, ms_assemble, ms_factor, ms_solve, ms_search, ms_total;   /* CUDA-event phase times */
    float  ms_hess_kernel, ms_condense_kernel;   /* the two SYRK-shaped fp64 contractions, per launch */
    int    n_spec;                   /* 1 if the delta = 0 and delta = max(delta/2, delta0) attempts of reghess ran
                                        concurrently on two streams (same decisions, n_factor counts both) */
    int    spec_used;                /* 1 if the speculative attempt was the accepted factorisation */
    int    tc_syrk;                  /* 1 if the two contractions of this step ran on tcgen05 (0: fp64 DMMA, also after
                                        a fallback because of non-finite inputs / unexpected negative weights) */
    int    abandoned_first;          /* 1 if the delta = 0 inertia test was abandoned on the device as soon as it had
                                        more than M negative pivots (n_neg_first is then a partial count) */
    int    cert_used;                /* 1 if the failure of the delta = 0 test was PROVEN by a negative-curvature vector in
                                        null(dce') instead of being computed (n_neg_first = -1) */
    int    n_factor_phys;            /* factorisations PHYSICALLY executed this step (n_factor counts the reference's
                                        eigvalsh-equivalents: a proven delta = 0 failure is counted there but costs none here) */
} b200ipm_step_info;

int         b200ipm_version(void);
const char* b200ipm_last_error(void);
/* number of kernels this library has launched in this process (bench.py's "gpu_launches") */
long long   b200ipm_launch_count(void);
/* sizeof(b200ipm_params) (which = 0) / sizeof(b200ipm_step_info) (which = 1) as compiled: lets a binding verify its
 * struct mirror before the first call */
int         b200ipm_struct_size(int which);

/* ---- lifecycle ----------------------------------------------------------------------------------- */
/* Replaces IPM.__init__/compile() workspace setup (pyipm.py:311-376, 410-467).  `stream` is a cudaStream_t
 * (NULL => the handle creates its own non-blocking stream). */
int b200ipm_create(int D, int M, int N, const b200ipm_params* p, int device, void* stream, b200ipm_handle* out);
int b200ipm_destroy(b200ipm_handle h);
int b200ipm_set_params(b200ipm_handle h, const b200ipm_params* p);
int b200ipm_sync(b200ipm_handle h);

/* ---- problem binding: replaces symbolic autodiff of f/ce/ci (pyipm.py:473-509) ------------------- */
/* Dense synthetic family (BASELINE.json configs 2/3/5): f = 1/2 x'Qx + c'x + q4/4 sum x^4,
 * ce = A x + 1/2 (U x)^2 - b, ci = G x - 1/2 (V x)^2 + r.  At/Ut are D x M, Gt/Vt are D x N (row-major);
 * Ut, Vt may be NULL. */
int b200ipm_bind_quad(b200ipm_handle h, const double* Q, const double* c, double q4,
                      const double* At, const double* Ut, const double* b,
                      const double* Gt, const double* Vt, const double* r, int on_device);
/* Sparse-polynomial lowering (the ten example problems, pyipm.py:1920-2131): CSR monomial table, row 0 = f,
 * rows 1..M = ce, rows M+1..M+N = ci; optional xlogx_coeff * sum x*log(x + xlogx_shift) added to f. */
int b200ipm_bind_poly(b200ipm_handle h, int nterms, const int* term_row, const double* term_coeff,
                      const int* term_ptr, const int* fac_var, const int* fac_pow,
                      double xlogx_coeff, double xlogx_shift);
/* User-callable mode: the caller evaluated f, df (D), ce (M), ci (N), J = [dce | dci] (D x (M+N)) and
 * d2L = d2f - d2ce - d2ci (D x D; only its upper triangle is used, pyipm.py:785,827) at the current state. */
int b200ipm_set_derivs(b200ipm_handle h, double fval, const double* df, const double* ce, const double* ci,
                       const double* J, const double* d2L, int on_device);

/* ---- state: (x, s, lda) and the hidden shared scalars mu_dev / nu_dev / delta (pyipm.py:363-364,1628) */
int b200ipm_set_state(b200ipm_handle h, const double* x, const double* s, const double* lda,
                      double mu, double nu, double delta);
int b200ipm_get_state(b200ipm_handle h, double* x, double* s, double* lda,
                      double* mu, double* nu, double* delta);
int b200ipm_set_mu_host(b200ipm_handle h, double mu_host);   /* pyipm.py:1603/1606 (reghess uses mu_host) */
/* Device-resident snapshot of (x, s, lda, mu, nu, delta, mu_host): warm start / teacher forcing without any
 * host<->device traffic (solve(x0, s0, lda0) warm start, pyipm.py:1567-1578). */
int b200ipm_state_save(b200ipm_handle h);
int b200ipm_state_restore(b200ipm_handle h);

/* ---- operator slots -------------------------------------------------------------------------------- */
/* cost(x), pyipm.py:855-857 */
int b200ipm_cost(b200ipm_handle h, double* fval);
/* grad(x,s,lda), pyipm.py:610-668,865-869: g (K = D+2N+M, may be NULL) and the four KKT 2-norms with
 * kkt2 = g_s * s (pyipm.py:958-991).  One pass over J. */
int b200ipm_residual(b200ipm_handle h, double* g, double kkt_norm[4]);
/* KKT(x,s,lda), pyipm.py:958-991: the four condition vectors themselves (D, N, M, N). */
int b200ipm_kkt(b200ipm_handle h, double* kkt1, double* kkt2, double* kkt3, double* kkt4);
/* con(x,s) pyipm.py:564-579 (M+N) and jaco(x)[:D,:] pyipm.py:581-607 (D x (M+N)); either may be NULL. */
int b200ipm_con_jac(b200ipm_handle h, double* con, double* J);
/* hess(x,s,lda), pyipm.py:768-844: the reference's FULL K x K symmetric KKT matrix (drop-in slot and parity
 * check; the hot path never forms it).  H is K*K doubles on the host. */
int b200ipm_hess_full(b200ipm_handle h, double* H);
/* Lagrangian Hessian d2L (D x D, symmetrised from its upper triangle). */
int b200ipm_d2L(b200ipm_handle h, double* W);
/* phi(x,s) / dphi(x,s,dz[:D+N]), pyipm.py:670-721,890-904 at the current state (dz from the last solve). */
int b200ipm_merit(b200ipm_handle h, double* phi, double* dphi);
/* init_slack / init_lambda, pyipm.py:723-744,1596-1621: sets s and/or lda in the state. */
int b200ipm_init_slack(b200ipm_handle h);
int b200ipm_init_lambda(b200ipm_handle h);
/* barrier update, pyipm.py:1804-1814: returns the new mu (caller stores it via set_state/set_mu_host). */
int b200ipm_update_mu(b200ipm_handle h, double* mu_new);

/* Second-order feasibility correction, pyipm.py:1464-1489 / 1516-1536: dz_p = -lstsq(jaco(x)', c_new) at the current state
 * (minimum-norm least squares on the device).  cnew: M+N doubles (host), pz: D+N doubles (host).  The lowered-problem
 * line search calls the same routine internally; this entry point serves callable mode, where c_new comes from the
 * user's functions. */
int b200ipm_soc_direction(b200ipm_handle h, const double* cnew, double* pz);
/* reghess + sym_solve_cmp, pyipm.py:1373-1406,1718-1725: condensed KKT formation, inertia-corrected LDL^T,
 * solve + refinement, multiplier sign flip.  dz (K, may be NULL) is in the reference ordering. */
int b200ipm_direction(b200ipm_handle h, double* dz, b200ipm_step_info* info);
/* step(), pyipm.py:1408-1436 (closed-form ratio test; equals the golden-section bracket to ~1e-15). */
int b200ipm_step_max(b200ipm_handle h, double* alpha_smax, double* alpha_lmax);
/* One full inner iteration, pyipm.py:1714-1754: grad, hess, reghess, solve, nu update, step rules, line
 * search (+ second-order correction), state update, KKT at the new point. */
int b200ipm_newton_step(b200ipm_handle h, b200ipm_step_info* info);

/* ---- L-BFGS mode (pyipm.py:993-1371): compact representation + Woodbury on the device ---------- */
/* lbfgs_init (pyipm.py:993-1005, 1633-1637): m = IPM's `lbfgs` (up to m + 1 pairs are kept, pyipm.py:1300), zeta =
 * `lbfgs_zeta`; empties the storage and snapshots the current x as x_old. */
int b200ipm_lbfgs_init(b200ipm_handle h, int m, double zeta);
/* lbfgs_update (pyipm.py:1282-1371) with dx = x - x_old, dg = dL/dx(x) - dL/dx(x_old) at the CURRENT (s, lda)
 * (pyipm.py:1706-1707).  gradx_old (D doubles, host) = dL/dx at x_old; NULL: evaluated here (lowered problems). */
int b200ipm_lbfgs_update(b200ipm_handle h, const double* gradx_old);
/* lbfgs_dir (pyipm.py:1184-1246, graph 1007-1182) + sign flip (1723-1725): dz (K, may be NULL). */
int b200ipm_lbfgs_direction(b200ipm_handle h, double* dz, b200ipm_step_info* info);
/* One inner iteration in L-BFGS mode (pyipm.py:1702-1713, 1723-1754): [update] + direction + nu rule + step rules +
 * line search + KKT at the new point.  do_update = (inner > 0 or outer > 0), pyipm.py:1705. */
int b200ipm_lbfgs_step(b200ipm_handle h, int do_update, b200ipm_step_info* info);
int b200ipm_lbfgs_state(b200ipm_handle h, int* m, double* zeta, int* fail);

/* ---- batched multi-start solves of SMALL problems (the reference's examples, pyipm.py:1920-2131) -----
 * The whole IPM.solve() loop (pyipm.py:1567-1863, exact-Hessian mode) for `batch` starting points of ONE polynomial
 * problem (same descriptor as b200ipm_bind_poly), one warp per instance, one launch for the whole batch; K = D + 2N + M
 * <= 32.  At this size the reference is followed literally: full K x K KKT matrix, reghess on its eigenvalues (cyclic
 * Jacobi), LU with partial pivoting.  x0 is batch x D (host); outputs (host, any of s/lda/fval/kkt_norm/signal/iters/ms may
 * be NULL): x batch x D, s batch x N, lda batch x (M+N), fval batch, kkt_norm batch x 4, signal batch (pyipm.py:1656
 * codes: 1 Ktol, 2 Ftol, -1 max iterations, -2 bad direction), iters batch (total Newton steps), ms = kernel time. */
int b200ipm_batch_solve_poly(int D, int M, int N, int nterms, const int* term_row, const double* term_coeff,
                             const int* term_ptr, const int* fac_var, const int* fac_pow, double xlogx_coeff,
                             double xlogx_shift, const b200ipm_params* p, int niter, int miter, int use_ftol, double Ftol,
                             int batch, const double* x0, int device, double* x, double* s, double* lda, double* fval,
                             double* kkt_norm, int* signal, int* iters, float* ms);

/* ---- generic dense symmetric-indefinite factor/solve (sym_solve_cmp slot, pyipm.py:911-914; config 4) -- */
/* A is n x n (leading dimension lda) symmetric, lower triangle referenced; factored in a private copy. */
int b200ipm_ldlt_create(int n, int device, void* stream, b200ipm_ldlt_handle* out);
int b200ipm_ldlt_destroy(b200ipm_ldlt_handle h);
int b200ipm_ldlt_factor(b200ipm_ldlt_handle h, const double* A, int lda, int on_device,
                        int inertia[3] /* pos, neg, zero */, double* rcond_est);
/* B is n x nrhs column-major-by-rhs (rhs r occupies B[r*n .. r*n+n)); solved in place; nrefine sweeps of
 * iterative refinement against the original A. */
int b200ipm_ldlt_solve(b200ipm_ldlt_handle h, double* B, int nrhs, int nrefine, int on_device);
/* Tile-level building blocks used by the multi-GPU 2-D block-cyclic driver (pyipm_b200/dist_ldlt.py).  All
 * pointers are device pointers; calls are asynchronous on the handle's stream.  tile_factor accumulates the
 * inertia counts (pos, neg, zero) on the device; a non-NULL `counts` returns the totals since the previous such call
 * (one synchronisation) and resets them. */
int b200ipm_ldlt_tile_factor(b200ipm_ldlt_handle h, double* tile_dev, int ld, int nb,
                             double* linv_dev, double* dblk_dev, int* perm_dev, int counts[3]);
/* panel_dev (rows x 64, leading dimension ld) <- L = W * D^-1, w_dev (leading dimension ldw) <- W = B * LinvP^T */
int b200ipm_ldlt_panel(b200ipm_ldlt_handle h, double* panel_dev, int ld, int rows, const double* linv_dev,
                       const double* dblk_dev, const int* perm_dev, double* w_dev, int ldw);
/* Adopt a factorisation assembled elsewhere (the replicated factor of the multi-GPU driver): A_dev holds the L
 * panels in its strictly-lower part (n x n, leading dimension lda), linvp_dev the per-tile inverses
 * (ceil(n/64) x 64 x 64), dinfo_dev the four D arrays [dinv_a | dinv_b | d_a | d_b] each of length
 * 64*ceil(n/64), kind_dev the pivot kinds.  All device pointers.  Afterwards b200ipm_ldlt_solve(nrefine = 0)
 * works on this handle. */
int b200ipm_ldlt_import(b200ipm_ldlt_handle h, const double* A_dev, int lda, const double* linvp_dev,
                        const double* dinfo_dev, const int* kind_dev);
int b200ipm_gemm_nt_update(b200ipm_ldlt_handle h, double* C_dev, int ldc, int rows, int cols,
                           const double* A_dev, int lda, const double* B_dev, int ldb, int k, int lower_only);
/* Composite calls of the block-cyclic driver (one call per block column): factor the b x b diagonal block at A_dev in
 * place (b a multiple of 64; tile steps + in-block panel / update) and pack its factor data into diag_dev -- [b*b copy of
 * the factored block | per 64-tile: LinvP 64*64, dinv_a, dinv_b, d_a, d_b (64 each), kind (64 ints)]; wdiag_dev is b*b
 * doubles of scratch.  block_panel: B_dev (rows x b, leading dimension ld) <- L, W_dev (rows x b contiguous) <- W = L D. */
int b200ipm_ldlt_block_factor(b200ipm_ldlt_handle h, double* A_dev, int ld, int b, double* diag_dev, double* wdiag_dev);
int b200ipm_ldlt_block_panel(b200ipm_ldlt_handle h, double* B_dev, int ld, int rows, int b, const double* diag_dev,
                             double* W_dev);
/* Block-column-cyclic driver, one call per block column: A_dev = top-left corner of the rows_total x b block column
 * (b = 256, rows_total a multiple of 64, >= b), factored with the single-GPU factorisation's panel schedule; L in place,
 * W = L D to Wb_dev (rows_total x b contiguous, row index = row of the block column), factor data of the diagonal block
 * packed into diag_dev (layout of block_factor).  The handle must have been created with order >= b.  rest_ready_event
 * (a cudaEvent_t, or NULL): the rows below the b x b diagonal block are not touched before this event has fired -- the
 * serial tile steps of the diagonal block overlap the tail of the previous panel's look-ahead update.  Replaces
 * block_factor + block_panel on a 1 x Q grid (sym_solve_cmp slot at config-4 size, pyipm.py:18-20). */
int b200ipm_ldlt_colblock_factor(b200ipm_ldlt_handle h, double* A_dev, int ld, int rows_total, int b, double* diag_dev,
                                 double* Wb_dev, void* rest_ready_event);
/* tcgen05 trailing updates of the block-column-cyclic driver (int8 error-free split, 21 slice pairs, in place):
 * oz_panel_slice -- digits of W (rows x 256) and -L (rows x 256) of the current panel, once per panel (max_rows sizes
 * the workspace at the first call); oz_block_update -- C (n x ncols lower trapezoid, origin on the diagonal, starting
 * row_off rows below the first sliced row; n, row_off multiples of 128) -= W L'; oz_status -- reads and clears the
 * error word (1: non-finite operand, 4: pipeline timeout).  All asynchronous on the handle's stream but oz_status. */
int b200ipm_oz_panel_slice(b200ipm_ldlt_handle h, int rows, const double* W_dev, int ldw, const double* L_dev, int ldl,
                           int max_rows);
int b200ipm_oz_block_update(b200ipm_ldlt_handle h, double* C_dev, int ldc, int n, int ncols, int row_off);
int b200ipm_oz_status(b200ipm_ldlt_handle h, int* err_word);
/* y (rows) = A (rows x cols, row-major, leading dimension lda) * v: the residual's HBM-bound GEMV kernel on device pointers
 * (distributed refinement mat-vec of the block-cyclic driver); asynchronous on the handle's stream. */
int b200ipm_ldlt_gemv(b200ipm_ldlt_handle h, const double* A_dev, int lda, int rows, int cols, const double* v_dev,
                      double* y_dev);
/* Trailing update of a rank's local piece of a 2-D block-cyclic matrix in ONE launch: C (rows x cols, local
 * storage, origin at local block (li0, lj0)) -= A B^T on every 128 x 128 tile whose global block row
 * (li*P + p) >= its global block column (lj*Q + q); block must be a multiple of 128. */
int b200ipm_gemm_nt_update_bc(b200ipm_ldlt_handle h, double* C_dev, int ldc, int rows, int cols,
                              const double* A_dev, int lda, const double* B_dev, int ldb, int k,
                              int block, int P, int Q, int p, int q, int li0, int lj0);

/* ---- measurement hook (bench.py roofline legs, ncu captures) -------------------------------------- */
/* Re-launches ONE hot kernel `reps` times on the handle's stream between two CUDA events at the current state
 * and returns the mean milliseconds per launch plus its algorithmic work (FLOPs or bytes):
 *   which = 0 residual GEMV g_x = df - J*lda (HBM)      1 Lagrangian-Hessian SYRK (fp64 tensor)
 *           2 condensation SYRK dci*S*dci' (fp64 tensor) 3 one LDL^T factorisation of the condensed KKT matrix
 *           4 one forward+backward triangular solve      5 J'*dx GEMV (HBM)
 *           6 / 7 = 1 / 2 on the tcgen05 int8 path (work is still the fp64-equivalent FLOP count)
 *           8 = the tcgen05 kernel of 6 alone, slices reused (work = int8 operations issued) */
int b200ipm_profile_kernel(b200ipm_handle h, int which, int reps, float* ms_per_launch, double* work);

/* ---- test hooks: individual kernels against the oracle (tests/test_gpu_kernels.py) ---------------- */
/* C (n x n, symmetric, mirrored) = beta*sym(triu(Cin)) + diag(dadd) + shift*I + sum_t alpha_t A_t diag(w_t) A_t'.
 * Host arrays; A_t is n x K_t row-major.  force_simple != 0 runs the scalar reference kernel. */
int b200ipm_test_syrk(int n, const double* Cin, double beta, const double* dadd, double shift,
                      int nterms, const double* const* A, const double* const* w, const int* K,
                      const double* alpha, double* C, int force_simple, float* ms);
int b200ipm_test_gemv(int rows, int cols, const double* A, const double* v, double* y, int transpose);
/* Device-side timeline of the factorisation kernels (profiling aid): after trace_start, selected CTAs of the LDL^T
 * kernels stamp %globaltimer (ns) at entry/exit; trace_dump copies up to `max` records (kernel id 1 tile, 2 panel,
 * 3 mini, 4 in-panel update, 5 DMMA trailing update; CTA index; t0; t1) to the host and switches the trace off. */
int b200ipm_trace_start(void);
int b200ipm_trace_dump(int* id, int* blk, unsigned long long* t0, unsigned long long* t1, unsigned long long* tag,
                       int max, int* n);   /* tag = control block of the factorisation a record belongs to */
/* Same product as b200ipm_test_syrk, computed on the tcgen05 tensor cores by the int8 error-free (Ozaki) path:
 * signed_mask bit t = alpha_t*w_t may be negative; variant & 15: 0 = 128x64 tiles / one pass, 1 = 128x128 tiles / two
 * passes, 2 = 128x256 / four; variant >> 4 (128x128 tiles only): slice-pair diagonals kept, 6 / 7 / 8 (default 8); lbo, sbo <= 0 keep the default shared-memory descriptor strides; ms[2] = {slicing ms,
 * total ms of one call};
 * *err = device error word (1 non-finite input, 2 negative weight without sign operand, 4 pipeline timeout). */
int b200ipm_test_syrk_i8(int n, const double* Cin, double beta, const double* dadd, double shift,
                         int nterms, const double* const* A, const double* const* w, const int* K,
                         const double* alpha, double* C, unsigned signed_mask, int variant, int lbo, int sbo,
                         float* ms, int* err);

#ifdef __cplusplus
}
#endif
#endif /* B200IPM_H */
