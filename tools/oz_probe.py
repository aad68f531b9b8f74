"""Hardware probe for the tcgen05 int8 (Ozaki) SYRK: python tools/oz_probe.py <case> [variant] [lbo] [sbo]
Prints one line per run; dumps inputs/outputs of the small case so a wrong descriptor can be diagnosed offline."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pyipm_b200 import _lib  # noqa: E402


def ref_syrk(n, Cin, beta, dadd, shift, terms):
    C = np.zeros((n, n))
    for A, w, al in terms:
        C += al * (A * (w if w is not None else 1.0)) @ A.T
    if Cin is not None:
        U = np.triu(Cin)
        C += beta * (U + U.T - np.diag(np.diag(Cin)))
    if dadd is not None:
        C += np.diag(dadd)
    C += shift * np.eye(n)
    return C


def main():
    case = sys.argv[1]
    variant = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    lbo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
    sbo = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    rng = np.random.default_rng(7)
    if case == 'tiny':          # one tile, one k-block, exact small integers: every slice but the first is zero
        n, K = 128, 32
        A = rng.integers(-31, 32, size=(n, K)).astype(np.float64)
        A[:, 0] = 32.0          # row max 32 -> e = 6: first slice holds the integers exactly
        terms = [(A, None, 1.0)]
        Cin = dadd = None
        beta = shift = 0.0
        mask = 0
    elif case == 'small':
        n, K = 128, 32
        A = rng.standard_normal((n, K))
        terms = [(A, None, 1.0)]
        Cin = dadd = None
        beta = shift = 0.0
        mask = 0
    elif case == 'mid':
        n = 300
        A0 = rng.standard_normal((n, 70))
        A1 = rng.standard_normal((n, 203))
        w0 = rng.standard_normal(70)
        w1 = 10.0 ** rng.uniform(-6, 6, size=203)
        terms = [(A0, w0, -1.0), (A1, w1, 1.0)]
        Cin = rng.standard_normal((n, n))
        dadd = rng.standard_normal(n)
        beta, shift, mask = 1.0, 0.25, 1
    elif case == 'big':
        n = 4096
        A0 = rng.standard_normal((n, 512)) / 64.0
        A1 = rng.standard_normal((n, 4096)) / 64.0
        w0 = rng.standard_normal(512)
        w1 = 10.0 ** rng.uniform(-3, 3, size=4096)
        terms = [(A0, w0, -1.0), (A1, w1, 1.0)]
        Cin = rng.standard_normal((n, n))
        dadd = None
        beta, shift, mask = 1.0, 0.0, 1
    else:
        raise SystemExit('unknown case')
    t0 = time.time()
    C, ms, err = _lib.test_syrk_i8(n, Cin, beta, dadd, shift, terms, signed_mask=mask, variant=variant, lbo=lbo, sbo=sbo)
    wall = time.time() - t0
    ref = ref_syrk(n, Cin, beta, dadd, shift, terms)
    scale = np.max(np.abs(ref))
    bad = ~np.isfinite(C)
    diff = np.where(bad, np.inf, np.abs(C - ref))
    relerr = float(np.max(diff) / scale)
    sym = bool(np.array_equal(C, C.T))
    print('case=%s variant=%d (tile %d, ndiag %d) lbo=%d sbo=%d  err_word=%d  max|C-ref|/max|ref|=%.3e  nonfinite=%d  symmetric=%s  ms(slice,total)=%.3f,%.3f  wall=%.1fs'
          % (case, variant, variant & 15, variant >> 4, lbo, sbo, err, relerr, int(bad.sum()), sym, ms[0], ms[1], wall), flush=True)
    if case in ('tiny', 'small') and relerr > 1e-12:
        out = os.path.join('gpurun_out', 'oz_%s_v%d_%d_%d.npz' % (case, variant, lbo, sbo))
        np.savez_compressed(out, A=terms[0][0], C=C, ref=ref)
        print('  dumped', out, flush=True)
    if case == 'big':
        Cd, msd = _lib.test_syrk(n, Cin, beta, dadd, shift, terms)
        print('  DMMA kernel: ms=%.3f  max|C_dmma-ref|/max|ref|=%.3e   |C_i8 - C_dmma|/max=%.3e'
              % (msd, float(np.max(np.abs(Cd - ref)) / scale), float(np.max(np.abs(np.where(bad, 0, C) - Cd)) / scale)), flush=True)


if __name__ == '__main__':
    main()
