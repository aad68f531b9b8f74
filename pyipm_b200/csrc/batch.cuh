// batch.cuh -- batched multi-start solver for SMALL problems: ONE WARP PER PROBLEM INSTANCE.
//
// The reference's own problems (pyipm.py:1920-2131, unit_tests.py:96-237) have D <= 6 and K = D + 2N + M <= 20: a single
// solve is launch-latency bound on any GPU.  Here every instance of a batch -- the same polynomial problem, different
// starting points (multi-start) -- is solved START TO FINISH by one warp out of shared memory, one launch for the whole
// batch, thousands of instances in flight.  At this size the reference's algorithm is followed LITERALLY, no condensation:
//   * the full K x K KKT matrix of pyipm.py:768-844;
//   * reghess (pyipm.py:1373-1406) on its EIGENVALUES (cyclic Jacobi in shared memory instead of LAPACK dsygvd): inertia
//     count w < -eps, rcond = min|w| / max|w| <= eps, eq-block regularisation, delta loop;
//   * the direction by LU with partial pivoting (pyipm.py:18-20: assume_a = 'gen');
//   * nu rule, fraction-to-the-boundary (closed form of pyipm.py:1408-1436), Armijo backtracking with second-order
//     correction (pyipm.py:1438-1565), KKT tests, Ftol logic, barrier update (pyipm.py:1658-1814), signals.
// Lanes parallelise the O(K) / O(K^2) inner loops (one matrix row per lane); every scalar decision is warp-uniform.
#pragma once
#include "common.cuh"
#include "engine_kernels.cuh"

namespace b200 {

constexpr int BK_MAX = 32;            // K = D + 2N + M <= 32 (one matrix row per lane)
constexpr int BLD = BK_MAX + 1;       // padded row stride

struct BatchParams {
    double mu, nu, rho, tau, eta, beta, Ktol, Ftol, eps, reg_coef;
    int niter, miter, use_ftol;
};

// per-warp shared-memory workspace
struct BatchWs {
    double H[BK_MAX * BLD];           // KKT matrix (regularised in place by reghess)
    double Hw[BK_MAX * BLD];          // working copy (eigenvalues / LU / normal equations)
    double J[BK_MAX * BLD];           // [dce | dci], D x C
    double x[BK_MAX], s[BK_MAX], lam[BK_MAX];
    double xt[BK_MAX], st[BK_MAX];    // trial point
    double g[BK_MAX], dz[BK_MAX], pz[BK_MAX], rhs[BK_MAX], tmp[BK_MAX], tmp2[BK_MAX];
    double df[BK_MAX], ce[BK_MAX], ci[BK_MAX], cnew[BK_MAX];
    int piv[BK_MAX];
};

__device__ __forceinline__ double bw_sum(double v) { return warp_sum(v); }
__device__ __forceinline__ double bw_max(double v) { return warp_max(v); }
__device__ __forceinline__ double bw_min(double v) { return warp_min(v); }

// f, df, ce, ci, J at x
__device__ void bt_eval(const PolyData& P, int D, int M, int N, const double* x, double* fval, double* df, double* ce,
                        double* ci, double* J, int lane) {
    const int R = 1 + M + N;
    double f = 0.0;
    for (int r = lane; r < R; r += 32) {
        const double v = poly_row_value(P, r, x);
        if (r == 0) f = v;
        else if (r <= M) ce[r - 1] = v;
        else ci[r - 1 - M] = v;
    }
    if (P.xl_c != 0.0)
        for (int i = lane; i < D; i += 32) f += P.xl_c * x[i] * log(x[i] + P.xl_s);
    f = bw_sum(f);
    if (fval) *fval = f;
    if (df) {
        for (int idx = lane; idx < R * D; idx += 32) {
            const int r = idx / D, v = idx % D;
            double gv = poly_row_grad(P, r, v, x);
            if (r == 0) {
                if (P.xl_c != 0.0) gv += P.xl_c * (log(x[v] + P.xl_s) + x[v] / (x[v] + P.xl_s));
                df[v] = gv;
            } else {
                J[v * BLD + (r - 1)] = gv;
            }
        }
    }
    __syncwarp();
}
// merit pieces at (xt, st): f, ||ce||_1 + ||ci - st||_1, sum log st
__device__ void bt_merit(const PolyData& P, int D, int M, int N, const double* xt, const double* st, double* f, double* c1,
                         double* ls, int lane) {
    const int R = 1 + M + N;
    double fa = 0.0, ca = 0.0, la = 0.0;
    for (int r = lane; r < R; r += 32) {
        const double v = poly_row_value(P, r, xt);
        if (r == 0) fa += v;
        else if (r <= M) ca += fabs(v);
        else ca += fabs(v - st[r - 1 - M]);
    }
    if (P.xl_c != 0.0)
        for (int i = lane; i < D; i += 32) fa += P.xl_c * xt[i] * log(xt[i] + P.xl_s);
    for (int j = lane; j < N; j += 32) la += log(st[j]);
    *f = bw_sum(fa);
    *c1 = bw_sum(ca);
    *ls = bw_sum(la);
}
// grad(x, s, lda) (pyipm.py:610-668) -> g (K); needs df, ce, ci, J at x
__device__ void bt_grad(int D, int M, int N, const BatchWs& w, double mu, double eps, double* g, int lane) {
    const int C = M + N;
    for (int i = lane; i < D; i += 32) {
        double acc = w.df[i];
        for (int c = 0; c < C; c++) acc -= w.J[i * BLD + c] * w.lam[c];
        g[i] = acc;
    }
    for (int j = lane; j < N; j += 32) {
        g[D + j] = w.lam[M + j] - mu / (w.s[j] + eps);
        g[D + N + M + j] = w.ci[j] - w.s[j];
    }
    for (int j = lane; j < M; j += 32) g[D + N + j] = w.ce[j];
    __syncwarp();
}
// hess(x, s, lda) (pyipm.py:768-844): full symmetric K x K matrix
__device__ void bt_hess(const PolyData& P, int D, int M, int N, BatchWs& w, double eps, int lane) {
    const int K = D + 2 * N + M, R = 1 + M + N;
    for (int idx = lane; idx < K * K; idx += 32) w.H[(idx / K) * BLD + (idx % K)] = 0.0;
    __syncwarp();
    for (int idx = lane; idx < D * D; idx += 32) {
        const int i = idx / D, j = idx % D;
        if (i > j) continue;
        double acc = poly_row_hess(P, 0, i, j, w.x);
        if (i == j && P.xl_c != 0.0) {
            const double t = w.x[i] + P.xl_s;
            acc += P.xl_c * (1.0 / t + P.xl_s / (t * t));
        }
        for (int r = 1; r < R; r++) {
            const double hh = poly_row_hess(P, r, i, j, w.x);
            if (hh != 0.0) acc -= w.lam[r - 1] * hh;
        }
        w.H[i * BLD + j] = acc;
        w.H[j * BLD + i] = acc;
    }
    for (int idx = lane; idx < D * (M + N); idx += 32) {
        const int i = idx / (M + N), c = idx % (M + N);
        const double v = w.J[i * BLD + c];
        w.H[i * BLD + D + N + c] = v;
        w.H[(D + N + c) * BLD + i] = v;
    }
    for (int j = lane; j < N; j += 32) {
        w.H[(D + j) * BLD + D + j] = w.lam[M + j] / (w.s[j] + eps);
        w.H[(D + j) * BLD + D + N + M + j] = -1.0;
        w.H[(D + N + M + j) * BLD + D + j] = -1.0;
    }
    __syncwarp();
}
// eigenvalues of the symmetric n x n matrix A (destroyed; eigenvalues end up on its diagonal): cyclic Jacobi
__device__ void bt_eig(double* A, int n, int lane) {
    for (int sweep = 0; sweep < 30; sweep++) {
        double off = 0.0, dia = 0.0;
        if (lane < n) {
            for (int j = 0; j < n; j++) {
                const double v = A[lane * BLD + j];
                if (j == lane) dia += v * v; else off += v * v;
            }
        }
        off = bw_sum(off);
        dia = bw_sum(dia);
        if (!(off > 1e-34 * (dia + off)) || !(off == off)) break;      // (NaN: stop, the caller sees NaN eigenvalues)
        for (int p = 0; p < n - 1; p++) {
            for (int q = p + 1; q < n; q++) {
                const double apq = A[p * BLD + q];
                if (apq == 0.0) continue;                             // warp-uniform (smem broadcast)
                const double app = A[p * BLD + p], aqq = A[q * BLD + q];
                const double theta = (aqq - app) / (2.0 * apq);
                const double t = ((theta >= 0.0) ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
                const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
                __syncwarp();
                if (lane < n && lane != p && lane != q) {
                    const double akp = A[lane * BLD + p], akq = A[lane * BLD + q];
                    const double np_ = c * akp - sn * akq, nq_ = sn * akp + c * akq;
                    A[lane * BLD + p] = np_; A[p * BLD + lane] = np_;
                    A[lane * BLD + q] = nq_; A[q * BLD + lane] = nq_;
                }
                if (lane == 0) {
                    A[p * BLD + p] = app - t * apq;
                    A[q * BLD + q] = aqq + t * apq;
                    A[p * BLD + q] = 0.0;
                    A[q * BLD + p] = 0.0;
                }
                __syncwarp();
            }
        }
    }
    __syncwarp();
}
// general LU with partial pivoting on the n x n matrix A (destroyed), b -> solution (in place).  returns 0 if singular
__device__ int bt_lu_solve(double* A, int n, double* b, int lane) {
    for (int k = 0; k < n; k++) {
        // pivot search in column k, rows >= k
        double best = (lane >= k && lane < n) ? fabs(A[lane * BLD + k]) : -1.0;
        int bi = lane;
        for (int o = 16; o > 0; o >>= 1) {
            const double ob = __shfl_xor_sync(0xffffffffu, best, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        if (!(best > 0.0)) return 0;
        if (bi != k) {
            for (int j = lane; j < n; j += 32) {
                const double t = A[k * BLD + j]; A[k * BLD + j] = A[bi * BLD + j]; A[bi * BLD + j] = t;
            }
            if (lane == 0) { const double t = b[k]; b[k] = b[bi]; b[bi] = t; }
        }
        __syncwarp();
        const double d = A[k * BLD + k];
        if (lane > k && lane < n) {
            const double l = A[lane * BLD + k] / d;
            if (l != 0.0) {
                for (int j = k + 1; j < n; j++) A[lane * BLD + j] -= l * A[k * BLD + j];
                b[lane] -= l * b[k];
            }
        }
        __syncwarp();
    }
    for (int i = n - 1; i >= 0; i--) {
        double acc = 0.0;
        for (int j = i + 1 + lane; j < n; j += 32) acc += A[i * BLD + j] * b[j];
        acc = bw_sum(acc);
        if (lane == 0) b[i] = (b[i] - acc) / A[i * BLD + i];
        __syncwarp();
    }
    return 1;
}
// inertia statistics of H's eigenvalues: #(w < -eps), min|w|, max|w|
__device__ void bt_inertia(BatchWs& w, int K, double eps, int* nneg, double* rcond, int lane) {
    for (int idx = lane; idx < K * K; idx += 32) w.Hw[(idx / K) * BLD + (idx % K)] = w.H[(idx / K) * BLD + (idx % K)];
    __syncwarp();
    bt_eig(w.Hw, K, lane);
    double mn = INFINITY, mx = 0.0;
    int neg = 0;
    if (lane < K) {
        const double e = w.Hw[lane * BLD + lane];
        neg = (e < -eps) ? 1 : 0;
        mn = fabs(e); mx = fabs(e);
        if (!(e == e)) { mn = 0.0; mx = INFINITY; }
    }
    neg = __reduce_add_sync(0xffffffffu, neg);
    mn = bw_min(mn);
    mx = bw_max(mx);
    *nneg = neg;
    *rcond = mn / mx;
}
// fraction-to-the-boundary (closed form of pyipm.py:1408-1436; the exact test at alpha = 1 first)
__device__ double bt_step(const double* v, const double* dv, int n, double tau, int lane) {
    const double omt = 1.0 - tau;
    double a = INFINITY;
    bool ok = true;
    for (int j = lane; j < n; j += 32) ftb_accum(v[j], dv[j], omt, ok, a);
    a = bw_min(a);
    const int nok = __reduce_add_sync(0xffffffffu, ok ? 0 : 1);
    return nok == 0 ? 1.0 : fmin(a, 1.0);
}
// minimum-norm least squares  z = A' (A A' + t I)^-1 c  with iterated Tikhonov (as the large engine, b200ipm.cu:
// soc_direction / init_lambda).  A is given through the callback-free form used here: rows of A = columns of Jfull, where
// Jfull = [[J],[0, -I]] ((D+N) x C).  which = 0: init_lambda (A = J' restricted: solve J lam = df in the LS sense);
// which = 1: SOC (A = jaco(x0)' , C x (D+N)).
__device__ void bt_soc(BatchWs& w, int D, int M, int N, const double* cnew, double* pz, int lane) {
    const int C = M + N, P = D + N;
    // G = J'J + diag(0_M, I_N) + tik I   (C x C)
    double scale = 0.0;
    for (int c = lane; c < C; c += 32) {
        double acc = 0.0;
        for (int i = 0; i < D; i++) acc += w.J[i * BLD + c] * w.J[i * BLD + c];
        scale = fmax(scale, acc);
    }
    scale = bw_max(scale) + 1.0;
    const double tik = 1e-7 * scale;
    for (int i = lane; i < P; i += 32) pz[i] = 0.0;
    __syncwarp();
    for (int it = 0; it < 6; it++) {
        // r = c - A z ;  (A z)_c = sum_i J[i][c] z_x[i]  - [c >= M] z_s[c - M]
        double rmax = 0.0, cmax = 0.0;
        for (int c = lane; c < C; c += 32) {
            double acc = cnew[c];
            for (int i = 0; i < D; i++) acc -= w.J[i * BLD + c] * pz[i];
            if (c >= M) acc += pz[D + c - M];
            w.rhs[c] = acc;
            rmax = fmax(rmax, fabs(acc));
            cmax = fmax(cmax, fabs(cnew[c]));
        }
        rmax = bw_max(rmax);
        cmax = bw_max(cmax);
        __syncwarp();
        if (it > 0 && rmax <= 1e-13 * cmax) break;
        for (int idx = lane; idx < C * C; idx += 32) {
            const int a = idx / C, b = idx % C;
            double acc = (a == b) ? tik + ((a >= M) ? 1.0 : 0.0) : 0.0;
            for (int i = 0; i < D; i++) acc += w.J[i * BLD + a] * w.J[i * BLD + b];
            w.Hw[a * BLD + b] = acc;
        }
        __syncwarp();
        bt_lu_solve(w.Hw, C, w.rhs, lane);
        // z += A' u : z_x += J u ; z_s -= u_i
        for (int i = lane; i < D; i += 32) {
            double acc = 0.0;
            for (int c = 0; c < C; c++) acc += w.J[i * BLD + c] * w.rhs[c];
            pz[i] += acc;
        }
        for (int j = lane; j < N; j += 32) pz[D + j] -= w.rhs[M + j];
        __syncwarp();
    }
    for (int i = lane; i < P; i += 32) pz[i] = -pz[i];
    __syncwarp();
}
// lda0 = pinv(J) df (pyipm.py:723-730): min-norm LS through (J J' + t I), iterated Tikhonov; negative lda_i -> Ktol
__device__ void bt_init_lambda(BatchWs& w, int D, int M, int N, double Ktol, int lane) {
    const int C = M + N;
    double scale = 0.0;
    for (int i = lane; i < D; i += 32) {
        double acc = 0.0;
        for (int c = 0; c < C; c++) acc += w.J[i * BLD + c] * w.J[i * BLD + c];
        scale = fmax(scale, acc);
    }
    scale = bw_max(scale);
    const double tik = 1e-7 * (scale > 0.0 ? scale : 1.0);
    for (int c = lane; c < C; c += 32) w.lam[c] = 0.0;
    __syncwarp();
    for (int it = 0; it < 6; it++) {
        for (int i = lane; i < D; i += 32) {
            double acc = w.df[i];
            for (int c = 0; c < C; c++) acc -= w.J[i * BLD + c] * w.lam[c];
            w.rhs[i] = acc;
        }
        for (int idx = lane; idx < D * D; idx += 32) {
            const int a = idx / D, b = idx % D;
            double acc = (a == b) ? tik : 0.0;
            for (int c = 0; c < C; c++) acc += w.J[a * BLD + c] * w.J[b * BLD + c];
            w.Hw[a * BLD + b] = acc;
        }
        __syncwarp();
        bt_lu_solve(w.Hw, D, w.rhs, lane);
        for (int c = lane; c < C; c += 32) {
            double acc = 0.0;
            for (int i = 0; i < D; i++) acc += w.J[i * BLD + c] * w.rhs[i];
            w.lam[c] += acc;
        }
        __syncwarp();
    }
    for (int j = lane; j < N; j += 32)
        if (w.lam[M + j] < 0.0) w.lam[M + j] = Ktol;
    __syncwarp();
}

// One warp = one instance.  out_x (batch x D), out_s (batch x N), out_lam (batch x C), out_f, out_kkt (batch x 4),
// out_sig / out_it (batch).
__global__ void __launch_bounds__(32) batch_solve_kernel(int D, int M, int N, PolyData P, BatchParams prm, int batch,
                                                         const double* __restrict__ X0, double* __restrict__ out_x,
                                                         double* __restrict__ out_s, double* __restrict__ out_lam,
                                                         double* __restrict__ out_f, double* __restrict__ out_kkt,
                                                         int* __restrict__ out_sig, int* __restrict__ out_it) {
    extern __shared__ __align__(16) unsigned char bsm_raw[];
    BatchWs& w = *reinterpret_cast<BatchWs*>(bsm_raw);
    const int lane = threadIdx.x;
    const int b = blockIdx.x;
    if (b >= batch) return;
    const int C = M + N, K = D + 2 * N + M, Pn = D + N;
    const double eps = prm.eps, tau = prm.tau, eta = prm.eta;
    for (int i = lane; i < D; i += 32) w.x[i] = X0[(size_t)b * D + i];
    __syncwarp();
    double mu_dev = prm.mu, mu_host, nu = prm.nu, delta = 0.0, fval = 0.0;
    // ---- initialisation (pyipm.py:1596-1628)
    bt_eval(P, D, M, N, w.x, &fval, w.df, w.ce, w.ci, w.J, lane);
    if (N) {
        for (int j = lane; j < N; j += 32) w.s[j] = fmax(w.ci[j], prm.Ktol);
        mu_host = prm.mu;
    } else {
        mu_host = prm.Ktol;
        mu_dev = mu_host;
    }
    __syncwarp();
    if (C) bt_init_lambda(w, D, M, N, prm.Ktol, lane);
    bt_grad(D, M, N, w, mu_dev, eps, w.g, lane);
    auto kkt_norms = [&](double* n4) {
        double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
        for (int i = lane; i < D; i += 32) a0 += w.g[i] * w.g[i];
        for (int j = lane; j < N; j += 32) {
            const double v = w.g[D + j] * w.s[j];               // kkt2 = g_s * s (pyipm.py:972)
            a1 += v * v;
            a3 += w.g[D + N + M + j] * w.g[D + N + M + j];
        }
        for (int j = lane; j < M; j += 32) a2 += w.g[D + N + j] * w.g[D + N + j];
        n4[0] = sqrt(bw_sum(a0)); n4[1] = sqrt(bw_sum(a1)); n4[2] = sqrt(bw_sum(a2)); n4[3] = sqrt(bw_sum(a3));
    };
    double nrm[4];
    kkt_norms(nrm);
    int iter_count = 0, signal = 0;
    double f_past = fval;
    bool ftol_conv = false;
    const double reg_coef = prm.reg_coef;
    for (int outer = 0; outer < prm.niter; outer++) {
        if (nrm[0] <= prm.Ktol && nrm[1] <= prm.Ktol && nrm[2] <= prm.Ktol && nrm[3] <= prm.Ktol) { signal = 1; break; }
        for (int inner = 0; inner < prm.miter; inner++) {
            const double muTol = fmax(prm.Ktol, mu_host);
            if (nrm[0] <= muTol && nrm[1] <= muTol && nrm[2] <= muTol && nrm[3] <= muTol) {
                if (!C) signal = 1;
                break;
            }
            // ---- one Newton step (pyipm.py:1714-1754); df/ce/ci/J/g are valid at (x, s, lam)
            bt_hess(P, D, M, N, w, eps, lane);
            // reghess (pyipm.py:1373-1406)
            {
                int nneg; double rcond;
                bt_inertia(w, K, eps, &nneg, &rcond, lane);
                if (rcond <= eps || nneg != C) {
                    if (rcond <= eps && M) {
                        const double reg = reg_coef * eta * pow(mu_host, prm.beta);
                        for (int j = lane; j < M; j += 32) w.H[(D + N + j) * BLD + D + N + j] -= reg;
                    }
                    delta = (delta == 0.0) ? reg_coef : fmax(delta / 2.0, reg_coef);
                    for (int i = lane; i < D; i += 32) w.H[i * BLD + i] += delta;
                    __syncwarp();
                    bt_inertia(w, K, eps, &nneg, &rcond, lane);
                    int guard = 0;
                    while (nneg != C && guard++ < 400) {
                        for (int i = lane; i < D; i += 32) w.H[i * BLD + i] -= delta;
                        delta *= 10.0;
                        for (int i = lane; i < D; i += 32) w.H[i * BLD + i] += delta;
                        __syncwarp();
                        bt_inertia(w, K, eps, &nneg, &rcond, lane);
                    }
                }
            }
            // dz = Hc^-1 (-g), multiplier sign flip (pyipm.py:1717-1725)
            for (int i = lane; i < K; i += 32) w.dz[i] = -w.g[i];
            for (int idx = lane; idx < K * K; idx += 32) w.Hw[(idx / K) * BLD + (idx % K)] = w.H[(idx / K) * BLD + (idx % K)];
            __syncwarp();
            bt_lu_solve(w.Hw, K, w.dz, lane);
            for (int i = Pn + lane; i < K; i += 32) w.dz[i] = -w.dz[i];
            __syncwarp();
            // nu rule (pyipm.py:1727-1735)
            double c1_old = 0.0;
            for (int j = lane; j < M; j += 32) c1_old += fabs(w.ce[j]);
            for (int j = lane; j < N; j += 32) c1_old += fabs(w.ci[j] - w.s[j]);
            c1_old = bw_sum(c1_old);
            double dfdx = 0.0, bards = 0.0, logs0 = 0.0;
            for (int i = lane; i < D; i += 32) dfdx += w.df[i] * w.dz[i];
            for (int j = lane; j < N; j += 32) {
                bards += mu_dev / (w.s[j] + eps) * w.dz[D + j];
                logs0 += log(w.s[j]);
            }
            dfdx = bw_sum(dfdx); bards = bw_sum(bards); logs0 = bw_sum(logs0);
            if (C) {
                const double nu_thres = (dfdx - bards) / (1.0 - prm.rho) / c1_old;
                if (nu < nu_thres) nu = nu_thres;
            }
            // step rules (pyipm.py:1737-1742)
            double a_s = 1.0, a_l = 1.0;
            if (N) {
                a_s = bt_step(w.s, w.dz + D, N, tau, lane);
                a_l = bt_step(w.lam + M, w.dz + D + N + M, N, tau, lane);
            }
            if (!C) a_l = 0.0;
            // search (pyipm.py:1438-1565)
            double phi0 = fval;
            if (C) phi0 += nu * c1_old;
            if (N) phi0 -= mu_dev * logs0;
            double dphi0 = dfdx;
            if (C) dphi0 -= nu * c1_old;
            if (N) dphi0 -= bards;
            auto phi_at = [&](double a, const double* p, double cs, double scale, double* c1out) {
                for (int i = lane; i < D; i += 32) w.xt[i] = w.x[i] + scale * (a * w.dz[i] + (p ? cs * p[i] : 0.0));
                for (int j = lane; j < N; j += 32) w.st[j] = w.s[j] + scale * (a * w.dz[D + j] + (p ? cs * p[D + j] : 0.0));
                __syncwarp();
                double f, c1, ls;
                bt_merit(P, D, M, N, w.xt, w.st, &f, &c1, &ls, lane);
                if (c1out) *c1out = c1;
                double v = f;
                if (C) v += nu * c1;
                if (N) v -= mu_dev * ls;
                return v;
            };
            bool correction = false, bad = false;
            double alpha_corr = 0.0, c1_new = 0.0;
            if (phi_at(a_s, nullptr, 0.0, 1.0, &c1_new) > phi0 + a_s * eta * dphi0) {
                if (C && c1_new > c1_old) {
                    // second-order correction: c_new = con(x0 + a_s dx, s0 + a_s ds)
                    const int R = 1 + M + N;
                    for (int r = 1 + lane; r < R; r += 32) {
                        const double v = poly_row_value(P, r, w.xt);
                        if (r <= M) w.cnew[r - 1] = v; else w.cnew[r - 1] = v - w.st[r - 1 - M];
                    }
                    __syncwarp();
                    bt_soc(w, D, M, N, w.cnew, w.pz, lane);
                    if (phi_at(a_s, w.pz, 1.0, 1.0, nullptr) <= phi0 + a_s * eta * dphi0) {
                        if (N) {
                            for (int j = lane; j < N; j += 32) w.tmp[j] = a_s * w.dz[D + j] + w.pz[D + j];
                            __syncwarp();
                            alpha_corr = bt_step(w.s, w.tmp, N, tau, lane);
                            if (phi_at(a_s, w.pz, 1.0, alpha_corr, nullptr) <= phi0 + a_s * eta * dphi0) correction = true;
                        } else {
                            alpha_corr = 1.0;
                            correction = true;
                        }
                    }
                }
                if (!correction) {
                    double ndx = 0.0, nds = 0.0;
                    for (int i = lane; i < D; i += 32) ndx += w.dz[i] * w.dz[i];
                    for (int j = lane; j < N; j += 32) nds += w.dz[D + j] * w.dz[D + j];
                    ndx = sqrt(bw_sum(ndx)); nds = sqrt(bw_sum(nds));
                    a_s *= tau;
                    a_l *= tau;
                    int guard = 0;
                    while (phi_at(a_s, nullptr, 0.0, 1.0, nullptr) > phi0 + a_s * eta * dphi0) {
                        const double nrmstep = N ? sqrt((a_s * ndx) * (a_s * ndx) + (a_l * nds) * (a_l * nds)) : a_s * ndx;
                        if (nrmstep < eps || ++guard > 20000) { bad = true; break; }
                        a_s *= tau;
                        a_l *= tau;
                    }
                }
            }
            if (bad) {
                signal = -2;                     // pyipm.py:1502 / 1548: the iterate is left untouched
            } else {
                if (correction) {
                    for (int i = lane; i < D; i += 32) w.x[i] += alpha_corr * (a_s * w.dz[i] + w.pz[i]);
                    for (int j = lane; j < N; j += 32) w.s[j] += alpha_corr * (a_s * w.dz[D + j] + w.pz[D + j]);
                } else {
                    for (int i = lane; i < D; i += 32) w.x[i] += a_s * w.dz[i];
                    for (int j = lane; j < N; j += 32) w.s[j] += a_s * w.dz[D + j];
                }
                for (int c = lane; c < C; c += 32) w.lam[c] += a_l * w.dz[Pn + c];
                __syncwarp();
            }
            iter_count++;
            // KKT at the new point (pyipm.py:1754)
            bt_eval(P, D, M, N, w.x, &fval, w.df, w.ce, w.ci, w.J, lane);
            bt_grad(D, M, N, w, mu_dev, eps, w.g, lane);
            kkt_norms(nrm);
            if (prm.use_ftol && !N && signal != -2) {
                if (fabs(f_past - fval) <= fabs(prm.Ftol)) { signal = 2; ftol_conv = true; break; }
                f_past = fval;
            }
            if (signal == -2) break;
        }
        if (prm.use_ftol && N && signal != -2) {
            if (fabs(f_past - fval) <= fabs(prm.Ftol)) { signal = 2; ftol_conv = true; }
            else f_past = fval;
        }
        if (prm.use_ftol && ftol_conv) break;
        if (signal == -2) break;
        if (signal == 1 && !C) break;            // unconstrained: converged inside the inner loop (pyipm.py:1689-1692)
        if (outer >= prm.niter - 1) { signal = -1; break; }
        if (N) {
            // barrier update (pyipm.py:1804-1814)
            double mn = INFINITY, dot = 0.0;
            for (int j = lane; j < N; j += 32) {
                const double v = w.s[j] * w.lam[M + j];
                mn = fmin(mn, v);
                dot += v;
            }
            mn = bw_min(mn);
            dot = bw_sum(dot);
            const double xi = N * mn / (dot + eps);
            const double t = fmin(0.05 * (1.0 - xi) / (xi + eps), 2.0);
            mu_host = 0.1 * (t * t * t) * dot / N;
            if (mu_host < 0.0) mu_host = 0.0;
            mu_dev = mu_host;
            bt_grad(D, M, N, w, mu_dev, eps, w.g, lane);     // g_s depends on mu
            kkt_norms(nrm);
        }
    }
    for (int i = lane; i < D; i += 32) out_x[(size_t)b * D + i] = w.x[i];
    for (int j = lane; j < N; j += 32) out_s[(size_t)b * N + j] = w.s[j];
    for (int c = lane; c < C; c += 32) out_lam[(size_t)b * C + c] = w.lam[c];
    if (lane == 0) {
        out_f[b] = fval;
        for (int k = 0; k < 4; k++) out_kkt[(size_t)b * 4 + k] = nrm[k];
        out_sig[b] = signal;
        out_it[b] = iter_count;
    }
}

}  // namespace b200
