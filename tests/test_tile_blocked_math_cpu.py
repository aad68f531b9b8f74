"""CPU restatement of the blocked fast attempt of the LDL^T tile kernel (pyipm_b200/csrc/ldlt.cuh, tile_fast_blocked):
8 x 8 diagonal blocks factored serially, panel rows by per-row forward substitution against U = D L8', rank-8 updates
of the lower tiles only, L^-1 by block forward substitution; the threshold test in its |l_ij| <= 1/u form.  Checks the
identities the CUDA code relies on against dense NumPy algebra."""
import numpy as np
import pytest

NB = 64


def blocked_ldlt(T0, u=0.01):
    n = T0.shape[0]
    T = np.tril(T0).copy() + np.triu(np.full((n, n), np.nan), 1)     # the kernel never reads the upper triangle
    d = np.zeros(n)
    viol = False
    for kb in range(n // 8):
        c0 = 8 * kb
        a = np.tril(T[c0:c0 + 8, c0:c0 + 8])
        pinv = np.zeros(8)
        for j in range(8):                                           # (1a) one thread, registers only
            d[c0 + j] = a[j, j]
            pinv[j] = 1.0 / a[j, j]
            l = a[:, j] * pinv[j]
            viol |= bool(np.any(np.abs(l[j + 1:]) > 1.0 / u))
            for r in range(j + 1, 8):
                for c in range(j + 1, r + 1):
                    a[r, c] -= l[r] * a[c, j]
            a[j + 1:, j] = l[j + 1:]
        for r in range(1, 8):
            T[c0 + r, c0:c0 + r] = a[r, :r]
        for row in range(c0 + 8, n):                                 # (1b) one thread per row
            t = T[row, c0:c0 + 8].copy()
            l = np.zeros(8)
            for c in range(8):
                acc = t[c]
                for j in range(c):
                    acc -= l[j] * d[c0 + j] * T[c0 + c, c0 + j]      # U[j][c] = d_j l_cj
                l[c] = acc * pinv[c]
            viol |= bool(np.any(np.abs(l) > 1.0 / u))
            T[row, c0:c0 + 8] = l
        for BI in range(kb + 1, n // 8):                             # (2) rank-8 update, lower tiles only
            for BM in range(kb + 1, BI + 1):
                blk = T[8 * BI:8 * BI + 8, 8 * BM:8 * BM + 8]
                upd = (-T[8 * BI:8 * BI + 8, c0:c0 + 8] * d[c0:c0 + 8]) @ T[8 * BM:8 * BM + 8, c0:c0 + 8].T
                T[8 * BI:8 * BI + 8, 8 * BM:8 * BM + 8] = np.where(np.isnan(blk), np.nan, blk) + np.where(np.isnan(blk), 0.0, upd)
    L = np.tril(np.nan_to_num(T), -1) + np.eye(n)
    return L, d, viol


def block_inverse(L):
    n = L.shape[0]
    X = np.eye(n)
    for b in range(n // 8):                                          # diagonal blocks: forward substitution per column
        b0 = 8 * b
        X[b0:b0 + 8, b0:b0 + 8] = np.linalg.inv(L[b0:b0 + 8, b0:b0 + 8])
    for j in range(n // 8):                                          # one block column per warp
        for i in range(j + 1, n // 8):
            S = sum(L[8 * i:8 * i + 8, 8 * k:8 * k + 8] @ X[8 * k:8 * k + 8, 8 * j:8 * j + 8] for k in range(j, i))
            X[8 * i:8 * i + 8, 8 * j:8 * j + 8] = -X[8 * i:8 * i + 8, 8 * i:8 * i + 8] @ S
    return X


@pytest.mark.parametrize('seed', [0, 1, 2])
def test_blocked_elimination_reproduces_ldlt_and_inverse(seed):
    rng = np.random.default_rng(seed)
    B = rng.standard_normal((NB, NB))
    T0 = B @ B.T / NB + 0.5 * np.eye(NB)
    T0[40:, 40:] -= 2.0 * np.eye(NB - 40)                            # indefinite: negative pivots in the last block rows
    T0 = (T0 + T0.T) / 2
    L, d, viol = blocked_ldlt(T0)
    assert np.max(np.abs(L @ np.diag(d) @ L.T - T0)) < 1e-12 * np.max(np.abs(T0))
    assert int(np.sum(d < 0)) == int(np.sum(np.linalg.eigvalsh(T0) < 0))          # Sylvester: inertia from D
    X = block_inverse(L)
    assert np.max(np.abs(X @ L - np.eye(NB))) < 1e-11
    # the threshold test of the unblocked kernel, |d_j| >= u max_i |T^(j)[i][j]|, is the same statement as |l_ij| <= 1/u
    assert viol == bool(np.any(np.abs(np.tril(L, -1)) > 100.0))


def test_threshold_violation_is_detected():
    T0 = np.eye(NB)
    T0[0, 0] = 1e-6                                                  # tiny pivot under a unit column entry
    T0[5, 0] = T0[0, 5] = 1.0
    _, _, viol = blocked_ldlt(T0)
    assert viol
