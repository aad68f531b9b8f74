// engine_kernels.cuh -- problem evaluation (polynomial / quadratic-form lowerings), residual tail, KKT block
// assembly, direction statistics (nu rule, fraction-to-the-boundary, merit pieces) and speculative
// line-search trials.  All of these are O(D + M + N) or one pass over a matrix: HBM/latency bound.
#pragma once
#include "common.cuh"

namespace b200 {

// ================================================================================= QUAD lowering (C2/C3)
struct QuadData {
    const double *Q, *c, *At, *Ut, *b, *Gt, *Vt, *r;
    double q4;
};

// After the linear images qx = Qx, ax = A x, ux = U x, gx = G x, vx = V x are known (GEMVs):
//   df = qx + c + q4 x^3,  xdiag = 3 q4 x^2,  ce = ax + ux^2/2 - b,  ci = gx - vx^2/2 + r,
//   f = x.qx/2 + c.x + q4/4 sum x^4                                    (single CTA: sizes are O(D+M+N))
__global__ void __launch_bounds__(1024) quad_point_kernel(int D, int M, int N, QuadData q, const double* __restrict__ x,
                                                          const double* __restrict__ qx, const double* __restrict__ ax,
                                                          const double* __restrict__ ux, const double* __restrict__ gx,
                                                          const double* __restrict__ vx, double* __restrict__ df,
                                                          double* __restrict__ xdiag, double* __restrict__ ce,
                                                          double* __restrict__ ci, double* __restrict__ fval) {
    __shared__ double sh[33];
    double fa = 0.0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const double xi = x[i], x2 = xi * xi;
        if (df) df[i] = qx[i] + q.c[i] + q.q4 * x2 * xi;
        if (xdiag) xdiag[i] = 3.0 * q.q4 * x2;
        fa += 0.5 * xi * qx[i] + q.c[i] * xi + 0.25 * q.q4 * x2 * x2;
    }
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        double v = ax[j] - q.b[j];
        if (q.Ut) v += 0.5 * ux[j] * ux[j];
        ce[j] = v;
    }
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double v = gx[j] + q.r[j];
        if (q.Vt) v -= 0.5 * vx[j] * vx[j];
        ci[j] = v;
    }
    fa = block_sum(fa, sh);
    if (threadIdx.x == 0) fval[0] = fa;
}

// J[d, j<M] = At[d,j] + Ut[d,j]*ux[j];   J[d, M+j] = Gt[d,j] - Vt[d,j]*vx[j]
__global__ void quad_jac_kernel(int D, int M, int N, QuadData q, const double* __restrict__ ux,
                                const double* __restrict__ vx, double* __restrict__ J, int ldJ) {
    const int C = M + N;
    const size_t total = (size_t)D * C;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int d = (int)(idx / C), j = (int)(idx % C);
        double v;
        if (j < M) {
            v = q.At[(size_t)d * M + j];
            if (q.Ut) v += q.Ut[(size_t)d * M + j] * ux[j];
        } else {
            const int jj = j - M;
            v = q.Gt[(size_t)d * N + jj];
            if (q.Vt) v -= q.Vt[(size_t)d * N + jj] * vx[jj];
        }
        J[(size_t)d * ldJ + j] = v;
    }
}

// Speculative trials  alpha_k = alpha0 * tau^(k0 + k)  (k-th block; repeated multiplication exactly as
// pyipm.py:1492-1505 does on the host).  With the linear images of x and of the direction cached, a trial costs
// O(D+M+N): out[3k..3k+2] = (f_t, ||ce_t||_1 + ||ci_t - s_t||_1, sum log s_t).
struct QuadImages {
    const double *qx, *ax, *ux, *gx, *vx;   // images of x
    const double *qd, *ad, *ud, *gd, *vd;   // images of the direction dx
};
__global__ void __launch_bounds__(256) quad_trial_kernel(int D, int M, int N, QuadData q, QuadImages im,
                                                         const double* __restrict__ x, const double* __restrict__ s,
                                                         const double* __restrict__ dx, const double* __restrict__ ds,
                                                         double alpha0, double tau, int k0, double* __restrict__ out,
                                                         const double* __restrict__ alpha_dev = nullptr) {
    __shared__ double sh[33];
    double alpha = alpha_dev ? *alpha_dev : alpha0;      // alpha_dev: the step limit a kernel of the same stream just wrote
    for (int k = 0; k < k0 + (int)blockIdx.x; k++) alpha *= tau;
    double fa = 0.0, c1 = 0.0, ls = 0.0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const double xt = x[i] + alpha * dx[i];
        const double qt = im.qx[i] + alpha * im.qd[i];
        const double x2 = xt * xt;
        fa += 0.5 * xt * qt + q.c[i] * xt + 0.25 * q.q4 * x2 * x2;
    }
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        double v = (im.ax[j] + alpha * im.ad[j]) - q.b[j];
        if (q.Ut) {
            const double u = im.ux[j] + alpha * im.ud[j];
            v += 0.5 * u * u;
        }
        c1 += fabs(v);
    }
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        double v = (im.gx[j] + alpha * im.gd[j]) + q.r[j];
        if (q.Vt) {
            const double u = im.vx[j] + alpha * im.vd[j];
            v -= 0.5 * u * u;
        }
        const double st = s[j] + alpha * ds[j];
        c1 += fabs(v - st);
        ls += log(st);
    }
    fa = block_sum(fa, sh);
    c1 = block_sum(c1, sh);
    ls = block_sum(ls, sh);
    if (threadIdx.x == 0) {
        out[3 * blockIdx.x + 0] = fa;
        out[3 * blockIdx.x + 1] = c1;
        out[3 * blockIdx.x + 2] = ls;
    }
}

// ================================================================================= POLY lowering (examples)
struct PolyData {
    int nterms, nrows;          // nrows = 1 + M + N
    const int* row_ptr;         // nrows + 1 (terms are sorted by row)
    const double* coeff;
    const int* ptr;             // nterms + 1 into fvar / fpow
    const int* fvar;
    const int* fpow;
    double xl_c, xl_s;          // xlogx term
};
__device__ __forceinline__ double ipow(double x, int p) {
    double r = 1.0;
    for (int k = 0; k < p; k++) r *= x;
    return r;
}
__device__ __forceinline__ double poly_row_value(const PolyData& P, int r, const double* x) {
    double acc = 0.0;
    for (int t = P.row_ptr[r]; t < P.row_ptr[r + 1]; t++) {
        double m = P.coeff[t];
        for (int a = P.ptr[t]; a < P.ptr[t + 1]; a++) m *= ipow(x[P.fvar[a]], P.fpow[a]);
        acc += m;
    }
    return acc;
}
// d(row r)/d x_v
__device__ __forceinline__ double poly_row_grad(const PolyData& P, int r, int v, const double* x) {
    double acc = 0.0;
    for (int t = P.row_ptr[r]; t < P.row_ptr[r + 1]; t++) {
        int av = -1;
        for (int a = P.ptr[t]; a < P.ptr[t + 1]; a++)
            if (P.fvar[a] == v) av = a;
        if (av < 0) continue;
        double m = P.coeff[t] * P.fpow[av] * ipow(x[v], P.fpow[av] - 1);
        for (int a = P.ptr[t]; a < P.ptr[t + 1]; a++)
            if (a != av) m *= ipow(x[P.fvar[a]], P.fpow[a]);
        acc += m;
    }
    return acc;
}
// d2(row r)/d x_i d x_j
__device__ __forceinline__ double poly_row_hess(const PolyData& P, int r, int i, int j, const double* x) {
    double acc = 0.0;
    for (int t = P.row_ptr[r]; t < P.row_ptr[r + 1]; t++) {
        int ai = -1, aj = -1;
        for (int a = P.ptr[t]; a < P.ptr[t + 1]; a++) {
            if (P.fvar[a] == i) ai = a;
            if (P.fvar[a] == j) aj = a;
        }
        if (ai < 0 || aj < 0) continue;
        double m;
        if (i == j) {
            const int p = P.fpow[ai];
            if (p < 2) continue;
            m = P.coeff[t] * p * (p - 1) * ipow(x[i], p - 2);
        } else {
            m = P.coeff[t] * P.fpow[ai] * ipow(x[i], P.fpow[ai] - 1) * P.fpow[aj] * ipow(x[j], P.fpow[aj] - 1);
        }
        for (int a = P.ptr[t]; a < P.ptr[t + 1]; a++)
            if (a != ai && a != aj) m *= ipow(x[P.fvar[a]], P.fpow[a]);
        acc += m;
    }
    return acc;
}

// f, df, ce, ci, J at x   (grid-stride over (row, var) pairs; tiny problems)
__global__ void poly_eval_kernel(int D, int M, int N, PolyData P, const double* __restrict__ x, double* __restrict__ fval,
                                 double* __restrict__ df, double* __restrict__ ce, double* __restrict__ ci,
                                 double* __restrict__ J, int ldJ) {
    const int R = 1 + M + N;
    const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
    for (int r = tid; r < R; r += nth) {
        double v = poly_row_value(P, r, x);
        if (r == 0) {
            if (P.xl_c != 0.0)
                for (int i = 0; i < D; i++) v += P.xl_c * x[i] * log(x[i] + P.xl_s);
            fval[0] = v;
        } else if (r <= M) {
            ce[r - 1] = v;
        } else {
            ci[r - 1 - M] = v;
        }
    }
    if (df == nullptr) return;
    for (int idx = tid; idx < R * D; idx += nth) {
        const int r = idx / D, v = idx % D;
        double gval = poly_row_grad(P, r, v, x);
        if (r == 0) {
            if (P.xl_c != 0.0) gval += P.xl_c * (log(x[v] + P.xl_s) + x[v] / (x[v] + P.xl_s));
            df[v] = gval;
        } else {
            J[(size_t)v * ldJ + (r - 1)] = gval;
        }
    }
}
// W = d2f - sum_j lda_e[j] d2ce_j - sum_j lda_i[j] d2ci_j   (upper triangle computed, mirrored)
__global__ void poly_hess_kernel(int D, int M, int N, PolyData P, const double* __restrict__ x,
                                 const double* __restrict__ lam, double* __restrict__ W, int ldW) {
    const int R = 1 + M + N;
    for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < D * D; idx += gridDim.x * blockDim.x) {
        const int i = idx / D, j = idx % D;
        if (i > j) continue;
        double acc = poly_row_hess(P, 0, i, j, x);
        if (i == j && P.xl_c != 0.0) {
            const double t = x[i] + P.xl_s;
            acc += P.xl_c * (1.0 / t + P.xl_s / (t * t));
        }
        for (int r = 1; r < R; r++) {
            const double h = poly_row_hess(P, r, i, j, x);
            if (h != 0.0) acc -= lam[r - 1] * h;
        }
        W[(size_t)i * ldW + j] = acc;
        W[(size_t)j * ldW + i] = acc;
    }
}
// merit pieces at an explicit point (xt, st): out = (f, ||ce||_1 + ||ci - st||_1, sum log st)   [one CTA]
__global__ void __launch_bounds__(256) poly_point_merit_kernel(int D, int M, int N, PolyData P,
                                                               const double* __restrict__ xt, const double* __restrict__ st,
                                                               double* __restrict__ out) {
    __shared__ double sh[33];
    const int R = 1 + M + N;
    double fa = 0.0, c1 = 0.0, ls = 0.0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const double v = poly_row_value(P, r, xt);
        if (r == 0) fa += v;
        else if (r <= M) c1 += fabs(v);
        else c1 += fabs(v - st[r - 1 - M]);
    }
    if (P.xl_c != 0.0)
        for (int i = threadIdx.x; i < D; i += blockDim.x) fa += P.xl_c * xt[i] * log(xt[i] + P.xl_s);
    for (int j = threadIdx.x; j < N; j += blockDim.x) ls += log(st[j]);
    fa = block_sum(fa, sh);
    c1 = block_sum(c1, sh);
    ls = block_sum(ls, sh);
    if (threadIdx.x == 0) { out[0] = fa; out[1] = c1; out[2] = ls; }
}
// speculative trials for the polynomial lowering: block k evaluates alpha0*tau^(k0+k); xt/st staged in smem
__global__ void __launch_bounds__(256) poly_trial_kernel(int D, int M, int N, PolyData P, const double* __restrict__ x,
                                                         const double* __restrict__ s, const double* __restrict__ dx,
                                                         const double* __restrict__ ds, double alpha0, double tau, int k0,
                                                         double* __restrict__ out, const double* __restrict__ alpha_dev = nullptr) {
    extern __shared__ double psm[];
    __shared__ double sh[33];
    double* xt = psm;
    double* st = psm + D;
    double alpha = alpha_dev ? *alpha_dev : alpha0;
    for (int k = 0; k < k0 + (int)blockIdx.x; k++) alpha *= tau;
    for (int i = threadIdx.x; i < D; i += blockDim.x) xt[i] = x[i] + alpha * dx[i];
    for (int j = threadIdx.x; j < N; j += blockDim.x) st[j] = s[j] + alpha * ds[j];
    __syncthreads();
    const int R = 1 + M + N;
    double fa = 0.0, c1 = 0.0, ls = 0.0;
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        const double v = poly_row_value(P, r, xt);
        if (r == 0) fa += v;
        else if (r <= M) c1 += fabs(v);
        else c1 += fabs(v - st[r - 1 - M]);
    }
    if (P.xl_c != 0.0)
        for (int i = threadIdx.x; i < D; i += blockDim.x) fa += P.xl_c * xt[i] * log(xt[i] + P.xl_s);
    for (int j = threadIdx.x; j < N; j += blockDim.x) ls += log(st[j]);
    fa = block_sum(fa, sh);
    c1 = block_sum(c1, sh);
    ls = block_sum(ls, sh);
    if (threadIdx.x == 0) {
        out[3 * blockIdx.x + 0] = fa;
        out[3 * blockIdx.x + 1] = c1;
        out[3 * blockIdx.x + 2] = ls;
    }
}

// ================================================================================= engine core kernels
// in-place symmetrisation from the upper triangle (user-supplied d2L; quirk ii, pyipm.py:785,843)
__global__ void sym_from_upper_kernel(double* W, int ld, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < n && j < n && i > j) W[(size_t)i * ld + j] = W[(size_t)j * ld + i];
}
__global__ void sym_from_lower_kernel(double* W, int ld, int n) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, i = blockIdx.y * blockDim.y + threadIdx.y;
    if (i < n && j < n && i < j) W[(size_t)i * ld + j] = W[(size_t)j * ld + i];
}

// Residual tail (single CTA).  g = [g_x | g_s | g_e | g_i] with g_x already written by the GEMV:
//   g_s = lda_i - mu/(s+eps)  (pyipm.py:624-627, 665-666),  g_e = ce,  g_i = ci - s,  sigma = lda_i/(s+eps)
//   red[0..3] = ||g_x||, ||g_s*s||, ||ce||, ||ci-s||  (pyipm.py:958-991);  red[4] = ||con||_1;  red[5] = f
__global__ void __launch_bounds__(1024) residual_tail_kernel(int D, int M, int N, const double* __restrict__ s,
                                                             const double* __restrict__ lam, const double* __restrict__ ce,
                                                             const double* __restrict__ ci, double mu, double eps,
                                                             double* __restrict__ g, double* __restrict__ sigma,
                                                             const double* __restrict__ sq_part, int npart,
                                                             const double* __restrict__ fval, double* __restrict__ red) {
    __shared__ double sh[33];
    double n1 = 0.0, n2 = 0.0, n3 = 0.0, n4 = 0.0, l1 = 0.0;
    for (int i = threadIdx.x; i < npart; i += blockDim.x) n1 += sq_part[i];
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const double sj = s[j], lj = lam[M + j];
        const double gs = lj - mu / (sj + eps);
        g[D + j] = gs;
        sigma[j] = lj / (sj + eps);
        const double k2 = gs * sj;
        n2 += k2 * k2;
        const double gi = ci[j] - sj;
        g[D + N + M + j] = gi;
        n4 += gi * gi;
        l1 += fabs(gi);
    }
    for (int j = threadIdx.x; j < M; j += blockDim.x) {
        const double v = ce[j];
        g[D + N + j] = v;
        n3 += v * v;
        l1 += fabs(v);
    }
    n1 = block_sum(n1, sh);
    n2 = block_sum(n2, sh);
    n3 = block_sum(n3, sh);
    n4 = block_sum(n4, sh);
    l1 = block_sum(l1, sh);
    if (threadIdx.x == 0) {
        red[0] = sqrt(n1); red[1] = sqrt(n2); red[2] = sqrt(n3); red[3] = sqrt(n4); red[4] = l1; red[5] = fval[0];
    }
}

// Kc[0:D,0:D] = Hb + delta*I  (full block copy; the factorisation reads the lower triangle)
__global__ void kc_xx_kernel(const double* __restrict__ Hb, int ldh, int D, double delta, double* __restrict__ Kc, int ld) {
    const size_t total = (size_t)D * D;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int i = (int)(idx / D), j = (int)(idx % D);
        double v = Hb[(size_t)i * ldh + j];
        if (i == j) v += delta;
        Kc[(size_t)i * ld + j] = v;
    }
}
// Kc[D:, D:] = -reg * I
__global__ void kc_ee_kernel(double* __restrict__ Kc, int ld, int D, int M, double reg) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * M) return;
    const int i = idx / M, j = idx % M;
    Kc[(size_t)(D + i) * ld + D + j] = (i == j) ? -reg : 0.0;
}

// The reference's full K x K matrix (pyipm.py:824-844), for the hess() slot / parity tests.
__global__ void full_kkt_kernel(int D, int M, int N, const double* __restrict__ W, int ldW, const double* __restrict__ J,
                                int ldJ, const double* __restrict__ sigma, double* __restrict__ H) {
    const int K = D + 2 * N + M;
    const size_t total = (size_t)K * K;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        int i = (int)(idx / K), j = (int)(idx % K);
        const int ii = min(i, j), jj = max(i, j);   // evaluate the upper-triangle representative
        double v = 0.0;
        if (jj < D) {
            v = W[(size_t)ii * ldW + jj];
        } else if (ii < D) {
            if (jj >= D + N) v = J[(size_t)ii * ldJ + (jj - D - N)];     // [dce | dci] columns
        } else if (ii < D + N) {
            const int a = ii - D;
            if (jj == ii) v = sigma[a];
            else if (jj == D + N + M + a) v = -1.0;
        }
        H[idx] = v;
    }
}

// dst = src and per-block partial sums of src^2 (unconstrained problems: g_x = df)
__global__ void __launch_bounds__(256) copy_sq_kernel(int n, const double* __restrict__ src, double* __restrict__ dst,
                                                      double* __restrict__ part) {
    __shared__ double sh[33];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (i < n) { v = src[i]; dst[i] = v; }
    v = block_sum(v * v, sh);
    if (threadIdx.x == 0) part[blockIdx.x] = v;
}
__global__ void fill_kernel(int n, double val, double* __restrict__ p) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = val;
}
__global__ void mul_kernel(int n, const double* __restrict__ a, const double* __restrict__ b, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a[i] * b[i];
}
// b = -g ; or generic  out = a*x + b*y
__global__ void axpby_kernel(int n, double a, const double* __restrict__ x, double b, const double* __restrict__ y,
                             double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = a * x[i] + (y ? b * y[i] : 0.0);
}

// condensed right-hand side, stage 1:  t = sigma*b_i + b_s        (b in reference ordering [x|s|e|i])
__global__ void cond_t_kernel(int D, int M, int N, const double* __restrict__ sigma, const double* __restrict__ b,
                              double* __restrict__ t) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N) t[j] = sigma[j] * b[D + N + M + j] + b[D + j];
}
// expansion:  ds = (dci' dx) - b_i ;  y_i = sigma*ds - b_s ;  y_e copied from the condensed solution
//   jt = J' dx  (length M+N, only its dci part is used here)
__global__ void expand_kernel(int D, int M, int N, const double* __restrict__ sigma, const double* __restrict__ b,
                              const double* __restrict__ sol, const double* __restrict__ jt, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D) y[i] = sol[i];
    if (i < M) y[D + N + i] = sol[D + i];
    if (i < N) {
        const double ds = jt[M + i] - b[D + N + M + i];
        y[D + i] = ds;
        y[D + N + M + i] = sigma[i] * ds - b[D + i];
    }
}
// unreduced residual  rho = b - K y  given  wx = W dx + J*[y_e; y_i]  and  jt = J' dx:
//   rho_x = b_x - (wx + delta dx), rho_s = b_s - (sigma ds - y_i), rho_e = b_e - (jt_e - reg y_e),
//   rho_i = b_i - (jt_i - ds);   block partial max |rho| -> part[blockIdx.x]
__global__ void kkt_resid_kernel(int D, int M, int N, const double* __restrict__ sigma, double delta, double reg,
                                 const double* __restrict__ b, const double* __restrict__ y, const double* __restrict__ wx,
                                 const double* __restrict__ jt, double* __restrict__ rho, double* __restrict__ part) {
    __shared__ double sh[33];
    const int K = D + 2 * N + M;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    double v = 0.0;
    if (i < K) {
        if (i < D) {
            v = b[i] - (wx[i] + delta * y[i]);
        } else if (i < D + N) {
            const int a = i - D;
            v = b[i] - (sigma[a] * y[i] - y[D + N + M + a]);
        } else if (i < D + N + M) {
            const int a = i - D - N;
            v = b[i] - (jt[a] - reg * y[i]);
        } else {
            const int a = i - D - N - M;
            v = b[i] - (jt[M + a] - y[D + a]);
        }
        rho[i] = v;
    }
    double m = fabs(v);
    if (!(m == m)) m = INFINITY;   // NaN -> inf so that it survives the max
    m = warp_max(m);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) sh[w] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t = fmax(t, sh[k]);
        part[blockIdx.x] = t;
    }
}
__global__ void max_partials_kernel(const double* __restrict__ p, int n, double* out) {
    double m = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) m = fmax(m, p[i]);
    __shared__ double sh[32];
    m = warp_max(m);
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = m;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) t = fmax(t, sh[k]);
        out[0] = t;
    }
}
// dz (reference sign convention, pyipm.py:1723-1725) = [dx | ds | -y_e | -y_i]
__global__ void flip_kernel(int D, int N, int K, const double* __restrict__ y, double* __restrict__ dz) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < K) dz[i] = (i < D + N) ? y[i] : -y[i];
}

// Direction statistics (single CTA), everything the host needs for the nu rule, the step rules and the merit:
//  red[0] = grad(f + barrier) . dz[:D+N]   (pyipm.py:1729-1731)      red[1] = df . dx
//  red[2] = sum mu/(s+eps) * ds            (pyipm.py:699-702)        red[3] = sum log s   (pyipm.py:673)
//  red[4] = ||dx||^2   red[5] = ||ds||^2
//  red[6] = alpha_smax,  red[7] = alpha_lmax   fraction-to-the-boundary (pyipm.py:1408-1436, closed form)
__device__ __forceinline__ void ftb_accum(double v, double dv, double omt, bool& ok, double& amin) {
    const double thr = omt * v;
    if (!(v + dv >= thr)) {
        ok = false;
        if (dv < 0.0) amin = fmin(amin, (v - thr) / (-dv));
        else amin = fmin(amin, 0.0);   // v itself violates (non-positive v): no admissible step
    }
}
__global__ void __launch_bounds__(1024) dir_stats_kernel(int D, int M, int N, const double* __restrict__ df,
                                                         const double* __restrict__ s, const double* __restrict__ lam,
                                                         const double* __restrict__ dz, double mu, double eps, double tau,
                                                         double* __restrict__ red) {
    __shared__ double sh[33];
    double dfdx = 0.0, bards = 0.0, logs = 0.0, ndx = 0.0, nds = 0.0;
    double as = INFINITY, al = INFINITY;
    bool oks = true, okl = true;
    const double omt = 1.0 - tau;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const double d = dz[i];
        dfdx += df[i] * d;
        ndx += d * d;
    }
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const double sj = s[j], d = dz[D + j];
        bards += mu / (sj + eps) * d;
        logs += log(sj);
        nds += d * d;
        ftb_accum(sj, d, omt, oks, as);
        ftb_accum(lam[M + j], dz[D + N + M + j], omt, okl, al);
    }
    dfdx = block_sum(dfdx, sh);
    bards = block_sum(bards, sh);
    logs = block_sum(logs, sh);
    ndx = block_sum(ndx, sh);
    nds = block_sum(nds, sh);
    as = block_min(as, sh);
    al = block_min(al, sh);
    const double nok_s = block_sum(oks ? 0.0 : 1.0, sh), nok_l = block_sum(okl ? 0.0 : 1.0, sh);
    if (threadIdx.x == 0) {
        red[0] = dfdx - bards;
        red[1] = dfdx;
        red[2] = bards;
        red[3] = logs;
        red[4] = ndx;
        red[5] = nds;
        red[6] = (nok_s == 0.0) ? 1.0 : fmin(as, 1.0);
        red[7] = (nok_l == 0.0) ? 1.0 : fmin(al, 1.0);
    }
}
// fraction-to-the-boundary for an explicit (v, dv) pair (second-order correction, pyipm.py:1480-1481)
__global__ void __launch_bounds__(1024) ftb_kernel(int n, const double* __restrict__ v, const double* __restrict__ dv,
                                                   double tau, double* __restrict__ out) {
    __shared__ double sh[33];
    double a = INFINITY;
    bool ok = true;
    const double omt = 1.0 - tau;
    for (int j = threadIdx.x; j < n; j += blockDim.x) ftb_accum(v[j], dv[j], omt, ok, a);
    a = block_min(a, sh);
    const double nok = block_sum(ok ? 0.0 : 1.0, sh);
    if (threadIdx.x == 0) out[0] = (nok == 0.0) ? 1.0 : fmin(a, 1.0);
}

// state update (pyipm.py:1507-1510, 1553-1562):  x += a_s dx, s += a_s ds, lda += a_l dlda
__global__ void update_state_kernel(int D, int M, int N, double a_s, double a_l, const double* __restrict__ dz,
                                    double* __restrict__ x, double* __restrict__ s, double* __restrict__ lam) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D) x[i] += a_s * dz[i];
    if (i < N) s[i] += a_s * dz[D + i];
    if (i < M + N) lam[i] += a_l * dz[D + N + i];
}
// xt = x + a*dx (+ c*p_x), st = s + a*ds (+ c*p_s)  -- explicit trial points for the second-order correction
__global__ void trial_point_kernel(int D, int N, const double* __restrict__ x, const double* __restrict__ s,
                                   const double* __restrict__ dz, double a, const double* __restrict__ p, double c,
                                   double scale, double* __restrict__ xt, double* __restrict__ st) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D) xt[i] = x[i] + scale * (a * dz[i] + (p ? c * p[i] : 0.0));
    if (i < N) st[i] = s[i] + scale * (a * dz[D + i] + (p ? c * p[D + i] : 0.0));
}

// init_slack (pyipm.py:732-744): s = max(ci(x), Ktol)
__global__ void init_slack_kernel(int N, const double* __restrict__ ci, double Ktol, double* __restrict__ s) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N) s[j] = fmax(ci[j], Ktol);
}
// post-processing of init_lambda (pyipm.py:1614-1621): negative inequality multipliers -> Ktol
__global__ void fix_lambda_kernel(int M, int N, double Ktol, double* __restrict__ lam) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < N && lam[M + j] < 0.0) lam[M + j] = Ktol;
}
// barrier update pieces (pyipm.py:1804-1810): red[0] = min(s*lda_i), red[1] = s . lda_i
__global__ void __launch_bounds__(1024) mu_stats_kernel(int M, int N, const double* __restrict__ s,
                                                        const double* __restrict__ lam, double* __restrict__ red) {
    __shared__ double sh[33];
    double mn = INFINITY, dot = 0.0;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const double p = s[j] * lam[M + j];
        mn = fmin(mn, p);
        dot += p;
    }
    mn = block_min(mn, sh);
    dot = block_sum(dot, sh);
    if (threadIdx.x == 0) { red[0] = mn; red[1] = dot; }
}
// max_i sum_c J[i,c]^2  (scale for the Tikhonov parameter of the pseudo-inverse)
__global__ void __launch_bounds__(256) row_sqnorm_max_kernel(const double* __restrict__ J, int ldJ, int rows, int cols,
                                                             double* __restrict__ part) {
    __shared__ double sh[33];
    double m = 0.0;
    for (int r = blockIdx.x; r < rows; r += gridDim.x) {
        double a = 0.0;
        for (int c = threadIdx.x; c < cols; c += blockDim.x) {
            const double v = J[(size_t)r * ldJ + c];
            a += v * v;
        }
        a = block_sum(a, sh);
        m = fmax(m, a);
    }
    if (threadIdx.x == 0) part[blockIdx.x] = m;
}

// out[0] = max |a|, out[1] = max |b|   (one CTA; NaNs propagate as +inf so that callers never stop early on them)
__global__ void __launch_bounds__(1024) absmax2_kernel(int n, const double* __restrict__ a, const double* __restrict__ b,
                                                       double* __restrict__ out) {
    double ma = 0.0, mb = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) {
        const double x = fabs(a[i]), y = fabs(b[i]);
        ma = (x <= 1.7e308) ? fmax(ma, x) : INFINITY;
        mb = (y <= 1.7e308) ? fmax(mb, y) : INFINITY;
    }
    __shared__ double sh[2][32];
    ma = warp_max(ma);
    mb = warp_max(mb);
    if ((threadIdx.x & 31) == 0) { sh[0][threadIdx.x >> 5] = ma; sh[1][threadIdx.x >> 5] = mb; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double ta = 0.0, tb = 0.0;
        for (int k = 0; k < (int)(blockDim.x >> 5); k++) { ta = fmax(ta, sh[0][k]); tb = fmax(tb, sh[1][k]); }
        out[0] = ta;
        out[1] = tb;
    }
}

// ------------------------------------------------------------------------------------------- curvature certificate
// rhs = [u; 0]  (length Kc = D + M)
__global__ void cert_rhs_kernel(int D, int Kc, const double* __restrict__ u, double* __restrict__ rhs) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Kc) rhs[i] = (i < D) ? u[i] : 0.0;
}
// red[0..3] = (v' Hv, v' v, Hv' Hv, v' u);  then u <- v / ||v||   (one CTA, deterministic tree)
__global__ void __launch_bounds__(1024) cert_stats_kernel(int D, const double* __restrict__ v, const double* __restrict__ hv,
                                                          double* __restrict__ u, double* __restrict__ red) {
    __shared__ double sh[33];
    double q = 0.0, nv = 0.0, hn = 0.0, vu = 0.0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) {
        const double a = v[i], b = hv[i];
        q += a * b; nv += a * a; hn += b * b; vu += a * u[i];
    }
    q = block_sum(q, sh);
    nv = block_sum(nv, sh);
    hn = block_sum(hn, sh);
    vu = block_sum(vu, sh);
    const double sc = (nv > 0.0 && nv < 1.7e308) ? rsqrt(nv) : 0.0;
    for (int i = threadIdx.x; i < D; i += blockDim.x) u[i] = v[i] * sc;
    if (threadIdx.x == 0) { red[0] = q; red[1] = nv; red[2] = hn; red[3] = vu; }
}

// ------------------------------------------------------------------------------------------- L-BFGS (pyipm.py:993-1371)
// out[j] = a . B[j, :]   for j < k   (B row-major k x n, leading dimension ldb; one CTA per row, deterministic tree)
__global__ void __launch_bounds__(256) dots_kernel(int n, const double* __restrict__ a, const double* __restrict__ B, int ldb,
                                                   double* __restrict__ out) {
    __shared__ double sh[33];
    const double* b = B + (size_t)blockIdx.x * ldb;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) acc += a[i] * b[i];
    acc = block_sum(acc, sh);
    if (threadIdx.x == 0) out[blockIdx.x] = acc;
}
// W (2m x D): rows 0..m-1 = cs * S_k, rows m..2m-1 = cy * Y_k   (constrained: cs = zeta, cy = 1; unconstrained: 1, zeta)
__global__ void lb_build_w_kernel(int D, int m, int ld, const double* __restrict__ S, const double* __restrict__ Y, double cs,
                                  double cy, double* __restrict__ W) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (i < D) W[(size_t)r * ld + i] = (r < m) ? cs * S[(size_t)r * ld + i] : cy * Y[(size_t)(r - m) * ld + i];
}
// u (C) <- [0_M ; 1 / sigma]   (the slack part of B' A^-1 B, pyipm.py:1099-1104)
__global__ void lb_gdiag_kernel(int M, int N, const double* __restrict__ sigma, double* __restrict__ u) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M) u[i] = 0.0;
    if (i < N) u[M + i] = 1.0 / sigma[i];
}
// q (C) <- jt / zeta + [0 ; -v_s / sigma] - sub      (B' A^-1 v for v = [v_x; v_s], jt = J' v_x; sub optional)
__global__ void lb_btainv_kernel(int M, int N, double zeta, const double* __restrict__ jt, const double* __restrict__ vs,
                                 const double* __restrict__ sigma, const double* __restrict__ sub, double* __restrict__ q) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < M + N) {
        double v = jt[i] / zeta;
        if (i >= M && vs) v -= vs[i - M] / sigma[i - M];
        if (sub) v -= sub[i];
        q[i] = v;
    }
}
// out (D+N) <- a_scale * base / Adiag + sgn * A^-1 B u :  x rows: (a_scale * base_x + sgn * ju) / zeta,
// s rows: (a_scale * base_s - sgn * u_i) / sigma         (ju = J u; base optional)
__global__ void lb_ainvb_kernel(int D, int M, int N, double zeta, const double* __restrict__ base, double a_scale,
                                const double* __restrict__ ju, const double* __restrict__ u, double sgn,
                                const double* __restrict__ sigma, double* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < D) out[i] = ((base ? a_scale * base[i] : 0.0) + sgn * ju[i]) / zeta;
    else if (i < D + N) out[i] = ((base ? a_scale * base[i] : 0.0) - sgn * u[M + i - D]) / sigma[i - D];
}
// A[i, i] += d[i]  for i < n
__global__ void kc_diag_add_kernel(double* __restrict__ A, int ld, int n, const double* __restrict__ d) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) A[(size_t)i * ld + i] += d[i];
}
// y <- y - sum_c coef[c] * X[c, :]    (X row-major ncol x n)
__global__ void lb_combine_kernel(int n, int ncol, int ldx, const double* __restrict__ X, const double* __restrict__ coef,
                                  double sgn, double* __restrict__ y) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        double acc = 0.0;
        for (int c = 0; c < ncol; c++) acc += coef[c] * X[(size_t)c * ldx + i];
        y[i] += sgn * acc;
    }
}

}  // namespace b200
