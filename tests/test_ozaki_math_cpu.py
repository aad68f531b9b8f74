"""CPU restatement of the arithmetic behind the tcgen05 contraction path (pyipm_b200/csrc/ozaki_i8.cuh): the balanced
base-256 digit split done with 64-bit fixed-point integers, exact int32 slice products, fp64 recombination.  Checks the
invariants the CUDA kernels rely on (digit range, exact reconstruction, no int32 overflow at the documented K limit)
and the error of the 8 / 7 / 6-diagonal truncations against an exact rational product."""
from fractions import Fraction

import numpy as np
import pytest

NS = 7


def split(L):
    """rows of L -> (exponents e, digits t[p] in [-128, 127]) with L = 2^e * sum_p t[p] 2^(-7 - 8p) (+ rounding)."""
    m = np.abs(L).max(axis=1)
    e = np.where(m > 0, np.frexp(m)[1] + 1, 0)                  # ilogb(m) + 2:  |L| 2^-e < 1/2
    X = np.rint(np.ldexp(L, (55 - e)[:, None])).astype(np.int64)
    assert np.abs(X).max() < 2 ** 54
    dg = [None] * NS
    for p in range(NS - 1, -1, -1):
        t = (X & 255).astype(np.uint8).astype(np.int8).astype(np.int64)   # sign-extended low byte
        dg[p] = t
        X = (X - t) >> 8
    assert not X.any()
    return e, dg


def product(e, dg_l, dg_r, nd):
    n = dg_l[0].shape[0]
    h = np.zeros((n, dg_r[0].shape[0]))
    for d in range(nd - 1, -1, -1):
        acc = sum(dg_l[p] @ dg_r[d - p].T for p in range(max(0, d - NS + 1), min(d, NS - 1) + 1))
        assert np.abs(acc).max() < 2 ** 31                       # exact in the int32 TMEM accumulators
        h = h * 2.0 ** -8 + acc
    return np.ldexp(h, e[0][:, None] + e[1][None, :] - 14)


def exact(L, R, rows):
    return np.array([[float(sum(Fraction(a) * Fraction(b) for a, b in zip(L[i], R[j]))) for j in range(rows)]
                     for i in range(rows)])


def test_digits_are_int8_and_reconstruct_to_55_bits():
    rng = np.random.default_rng(0)
    L = rng.standard_normal((64, 500)) * 10.0 ** rng.uniform(-9, 3, (1, 500))
    e, dg = split(L)
    assert all(d.min() >= -128 and d.max() <= 127 for d in dg)
    assert -64 <= dg[0].min() and dg[0].max() <= 64              # |L| 2^-e < 1/2 leaves room for the carries
    rec = sum(dg[p] * 2.0 ** (-7 - 8 * p) for p in range(NS)) * np.ldexp(1.0, e)[:, None]
    assert np.max(np.abs(rec - L) / np.abs(L).max(axis=1, keepdims=True)) <= 2.0 ** -54


@pytest.mark.parametrize('nd,tol', [(8, 1e-15), (7, 5e-15), (6, 2e-11)])
def test_truncated_slice_products_against_exact_rational_arithmetic(nd, tol):
    rng = np.random.default_rng(nd)
    n, K = 48, 1500
    A = rng.standard_normal((n, K))
    w = 10.0 ** rng.uniform(-6, 6, K) * np.sign(rng.standard_normal(K))
    L = A * np.sqrt(np.abs(w))
    R = L * np.sign(w)                                           # sign-carrying operand (lda_e may be negative)
    eL, dL = split(L)
    eR, dR = split(R)
    assert np.array_equal(eL, eR)
    C = product((eL, eR), dL, dR, nd)
    ref = exact(L[:6], R[:6], 6)
    assert np.max(np.abs(C[:6, :6] - ref)) <= tol * np.max(np.abs(ref))


def test_int32_accumulators_cannot_overflow_at_the_documented_k_limit():
    # worst case: every digit at -128, 7 pairs on the longest diagonal
    assert NS * 18432 * 128 * 128 < 2 ** 31


def test_biased_add_yields_the_same_digits_as_the_recurrence():
    """oz_slice_kernel: the seven balanced base-256 digits of X (|X| < 2^54) used to come from seven rounds of
    `t = sign-extended low byte; X = (X - t) >> 8`; now they are the bytes of Y = X + 0x0080808080808080 with the top bit of
    each byte flipped (stored byte = two's complement of the digit).  Same digits, bit for bit -- including the carries
    through runs of 0x80 / 0x7f bytes and both signs."""
    rng = np.random.default_rng(7)
    BIAS = 0x0080808080808080
    xs = [0, 1, -1, 127, 128, -128, -129, 255, 256, 0x7f7f7f7f7f7f, -0x7f7f7f7f7f7f, 0x808080808080, -0x808080808080,
          (1 << 54) - 1, -(1 << 54) + 1, 0x3f80807f80ff7f, -0x3f80807f80ff7f]
    xs += [int(v) for v in rng.integers(-(1 << 54) + 1, (1 << 54) - 1, size=20000)]
    for X0 in xs:
        X, ref = X0, []
        for _ in range(7):                       # the recurrence (lowest digit first)
            byte = X & 0xff
            t = byte - 256 if byte >= 128 else byte
            ref.append(byte)
            X = (X - t) >> 8
        assert X == 0, X0                        # seven digits suffice below 2^54 (the last one is in [-64, 64])
        Y = ((X0 + BIAS) & 0xffffffffffffffff) ^ BIAS
        got = [(Y >> (8 * i)) & 0xff for i in range(7)]
        assert got == ref, (X0, got, ref)
        assert (Y >> 56) == 0
        digits = [b - 256 if b >= 128 else b for b in got]
        assert sum(d << (8 * i) for i, d in enumerate(digits)) == X0
