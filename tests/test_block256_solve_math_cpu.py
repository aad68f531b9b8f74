"""CPU restatement of the block-256 triangular solves (pyipm_b200/csrc/ldlt.cuh: ldlt_blockinv_kernel, ldlt_fwd256_kernel,
ldlt_bwd256_kernel), scaled down (tile 8, link = 4 tiles = 32 rows, 8 "CTAs" of 4 rows per link): the recurrence that
builds the explicit inverse of a link's diagonal block from the per-tile inverses, the two sweeps with one exchange of
partial results per link, and the tile-granular triangular structure each CTA relies on when it truncates its row of X."""
import numpy as np
import pytest

NB, SBT, SBC = 8, 4, 8          # tile, tiles per link, CTAs per link
SB = NB * SBT
SBR = SB // SBC


def make_factor(n, rng):
    """unit-lower-triangular-with-tile-permutations factor in the layout of the device code: per-tile LinvP_i = (P_i' L_ii)^-1,
    strictly-lower tiles L_ij stored as they are, D with 1x1 and 2x2 blocks inside tiles"""
    nblk = -(-n // NB)
    npad = -(-nblk // SBT) * SB                         # the restatement pads to whole links (the device guards instead)
    A = np.zeros((npad, npad))
    linvp = np.zeros((nblk, NB, NB))
    for i in range(nblk):
        nbl = min(NB, n - i * NB)                       # a partial last tile: the device factors its leading nbl x nbl part,
        L = np.eye(NB)                                  # the padded part stays the identity
        L[:nbl, :nbl] += np.tril(rng.standard_normal((nbl, nbl)), -1) * 0.5
        P = np.eye(NB)
        P[:nbl, :nbl] = np.eye(nbl)[rng.permutation(nbl)]
        linvp[i] = np.linalg.inv(P.T @ L)
        for j in range(i):
            A[i * NB:(i + 1) * NB, j * NB:(j + 1) * NB] = rng.standard_normal((NB, NB)) * 0.3
    A[n:, :] = 0.0
    dinv = np.zeros((npad, npad))
    k = 0
    while k < n:                                        # D^-1: 2x2 blocks never straddle a tile
        if k % NB < NB - 1 and rng.random() < 0.3 and k + 1 < n:
            B = rng.standard_normal((2, 2))
            B = B + B.T + np.array([[0.0, 3.0], [3.0, 0.0]])
            dinv[k:k + 2, k:k + 2] = np.linalg.inv(B)
            k += 2
        else:
            dinv[k, k] = 1.0 / (rng.standard_normal() + (3.0 if rng.random() < 0.5 else -3.0))
            k += 1
    return A, linvp, dinv, nblk


def block_inverses(A, linvp, n, nblk):
    """ldlt_blockinv_kernel: X_ii = LinvP_i, X_ij = -LinvP_i sum_{j<=k<i} L_ik X_kj; rows / columns >= n stay zero"""
    nb256 = -(-nblk // SBT)
    X = np.zeros((nb256, SB, SB))
    for I in range(nb256):
        t0 = I * SBT
        ntl = min(SBT, nblk - t0)
        for j in range(ntl):
            X[I, j * NB:(j + 1) * NB, j * NB:(j + 1) * NB] = linvp[t0 + j]
            for i in range(j + 1, ntl):
                S = np.zeros((NB, NB))
                for k in range(j, i):
                    S += A[(t0 + i) * NB:(t0 + i + 1) * NB, (t0 + k) * NB:(t0 + k + 1) * NB] @ X[I, k * NB:(k + 1) * NB, j * NB:(j + 1) * NB]
                X[I, i * NB:(i + 1) * NB, j * NB:(j + 1) * NB] = -linvp[t0 + i] @ S
        g = np.arange(SB) + I * SB
        X[I][g >= n, :] = 0.0
        X[I][:, g >= n] = 0.0
    return X


def solve_links(A, X, dinv, n, b):
    """forward / backward sweeps, one link = SBC CTAs of SBR rows; CTA c only reads the columns of X_I that the
    tile-granular triangular structure can make non-zero (ncol = NB * (c // 2 + 1) forward, k >= NB * (c // 2) backward)"""
    nb256 = X.shape[0]
    npad = nb256 * SB
    bb = np.zeros(npad)
    bb[:n] = b
    y = np.zeros(npad)
    for I in range(nb256):
        acc = np.zeros(SB)
        for c in range(SBC):                              # each CTA: its rows of b_I - sum_{J<I} L_IJ y_J
            r = slice(I * SB + c * SBR, I * SB + (c + 1) * SBR)
            acc[c * SBR:(c + 1) * SBR] = bb[r] - A[r, :I * SB] @ y[:I * SB]
        for c in range(SBC):                              # ... exchange, then its rows of X_I
            ncol = NB * (c // 2 + 1)
            rows = slice(c * SBR, (c + 1) * SBR)
            y[I * SB + c * SBR:I * SB + (c + 1) * SBR] = X[I][rows, :ncol] @ acc[:ncol]
    y[n:] = 0.0
    z = dinv @ y
    x = np.zeros(npad)
    for I in range(nb256 - 1, -1, -1):
        tv = np.zeros(SB)
        for c in range(SBC):
            cols = slice(I * SB + c * SBR, I * SB + (c + 1) * SBR)
            tv[c * SBR:(c + 1) * SBR] = z[cols] - A[(I + 1) * SB:, cols].T @ x[(I + 1) * SB:]
        tv[(np.arange(SB) + I * SB) >= n] = 0.0
        for c in range(SBC):
            kmin = NB * (c // 2)
            rows = slice(c * SBR, (c + 1) * SBR)
            x[I * SB + c * SBR:I * SB + (c + 1) * SBR] = X[I].T[rows, kmin:] @ tv[kmin:]
    return x[:n]


@pytest.mark.parametrize('n', [5, 8, 31, 32, 33, 70, 96, 131])
def test_block_links_reproduce_the_tile_chain(n):
    rng = np.random.default_rng(n)
    A, linvp, dinv, nblk = make_factor(n, rng)
    npad = A.shape[0]
    # dense restatement of the factor: Lt = block lower triangular with diagonal tiles (LinvP_i)^-1
    Lt = A.copy()
    for i in range(nblk):
        Lt[i * NB:(i + 1) * NB, i * NB:(i + 1) * NB] = np.linalg.inv(linvp[i])
    keep = np.arange(npad) < n
    Ltn = Lt[np.ix_(keep, keep)]
    b = rng.standard_normal(n)
    X = block_inverses(A, linvp, n, nblk)
    for I in range(X.shape[0]):                          # X_I inverts the diagonal block of the link (inside the matrix)
        g = np.arange(SB) + I * SB
        m = g < n
        T = Lt[np.ix_(g[m], g[m])] if m.any() else np.zeros((0, 0))
        np.testing.assert_allclose(X[I][np.ix_(m, m)] @ T, np.eye(m.sum()), atol=1e-9)
        for c in range(SBC):                             # the structure the CTAs rely on
            assert np.all(X[I][c * SBR:(c + 1) * SBR, NB * (c // 2 + 1):] == 0.0)
    x = solve_links(A, X, dinv, n, b)
    xref = np.linalg.solve(Ltn.T, dinv[np.ix_(keep, keep)] @ np.linalg.solve(Ltn, b))
    np.testing.assert_allclose(x, xref, rtol=1e-8, atol=1e-8 * np.max(np.abs(xref)))
