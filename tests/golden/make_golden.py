#!/usr/bin/env python
"""Generate golden fixtures by running the UNMODIFIED reference (/root/reference/pyipm.py).

Run in the authoring container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

The reference imports `aesara`, which is not installable here; `oracle/aesara_shim` (a lazy expression
evaluator forwarding to the same SciPy/NumPy calls Aesara forwards to) stands in for it.  The reference is
driven in its documented *precompiled function* input mode (pyipm.py:216-231): every derivative is a NumPy
callable (from pyipm_b200.problems) presented as a compiled `Function`.  Its source is not touched: per-step
traces are captured by wrapping the instance's callable slots (`hess`, `reghess`, `eigh`, `sym_solve_cmp`,
`step`, `search`, `phi`) from outside after `compile()`.

Output: tests/golden/ref_<name>.npz (+ this script = the provenance the oracle is pinned with).
"""
from __future__ import print_function

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'oracle', 'aesara_shim'))
sys.path.insert(0, '/root/reference')

import aesara  # noqa: E402  (the stand-in)
import aesara.tensor as T  # noqa: E402
import pyipm as ref  # noqa: E402  (the unmodified reference)

from pyipm_b200 import problems  # noqa: E402

F = lambda fn: aesara.Function(pyfunc=fn)  # noqa: E731


def run_reference(prob, x0, n_solves=1, **kw):
    cal = {k: F(v) for k, v in prob.callables().items()}
    p = ref.IPM(x0=np.array(x0, dtype=np.float64), x_dev=T.vector('x_dev'), lambda_dev=T.vector('lda_dev'),
                verbosity=-1, **cal, **kw)
    p.compile(nvar=prob.nvar)
    D, M, N = p.nvar, p.neq, p.nineq
    K = D + 2 * N + M
    steps = []
    cur = {}

    o_hess, o_reghess, o_eigh, o_solve, o_step, o_search, o_phi = (p.hess, p.reghess, p.eigh, p.sym_solve_cmp,
                                                                   p.step, p.search, p.phi)

    def w_hess(x, s, lda):
        cur.clear()
        cur.update(x=np.copy(x), s=np.copy(s), lda=np.copy(lda), mu=float(p.mu_dev.get_value()),
                   mu_host=float(p.mu_host), nu_before=float(p.nu_dev.get_value()),
                   delta_before=float(p.delta), n_eig=0, nneg=[], alphas=[], n_phi=0)
        return o_hess(x, s, lda)

    def w_eigh(Mx):
        w = o_eigh(Mx)
        cur['n_eig'] += 1
        cur['nneg'].append(int(np.sum(w < -p.eps)))
        if cur['n_eig'] == 1:
            cur['rcond'] = float(np.min(np.abs(w)) / np.max(np.abs(w)))
        return w

    def w_reghess(Hc):
        cur['Hfull'] = np.copy(Hc)
        out = o_reghess(Hc)
        cur['delta'] = float(p.delta)
        cur['Hreg'] = np.copy(out)
        return out

    def w_solve(Mx, b):
        out = o_solve(Mx, b)
        if 'dz_raw' not in cur and b.shape[0] == K:
            cur['g'] = np.copy(b).reshape(-1)
            cur['dz_raw'] = np.copy(out).reshape(-1)
        return out

    def w_step(v, dv):
        a = o_step(v, dv)
        cur['alphas'].append(float(a))
        return a

    def w_phi(x, s):
        cur['n_phi'] += 1
        return o_phi(x, s)

    def w_search(x0_, s0_, lda0_, dz, a_s, a_l):
        cur['dz'] = np.copy(dz)
        cur['nu'] = float(p.nu_dev.get_value())
        cur['alpha_smax'] = float(a_s)
        cur['alpha_lmax'] = float(a_l)
        n_phi0 = cur['n_phi']
        x, s, lda = o_search(x0_, s0_, lda0_, dz, a_s, a_l)
        cur['n_phi_search'] = cur['n_phi'] - n_phi0
        cur['x_new'], cur['s_new'], cur['lda_new'] = np.copy(x), np.copy(s), np.copy(lda)
        cur['signal'] = int(p.signal)
        steps.append(dict(cur))
        return x, s, lda

    p.hess, p.reghess, p.eigh, p.sym_solve_cmp, p.step, p.search, p.phi = (w_hess, w_reghess, w_eigh, w_solve,
                                                                            w_step, w_search, w_phi)
    results = []
    for _ in range(n_solves):
        x, s, lda, fval, kkt = p.solve()
        results.append(dict(x=np.copy(x), s=np.copy(s), lda=np.copy(lda), fval=float(fval),
                            kkt=[np.atleast_1d(np.asarray(k, dtype=np.float64)) for k in kkt],
                            signal=int(p.signal), nsteps=len(steps)))
    return p, steps, results


def run_reference_lbfgs(prob, x0, lbfgs=4, **kw):
    """The reference's L-BFGS mode (pyipm.py:993-1371, hooks 1702-1713), lbfgs=4 as in unit_tests.py:49; second
    derivatives are not passed (the mode never uses them)."""
    cal = {k: F(v) for k, v in prob.callables().items() if not k.startswith('d2')}
    p = ref.IPM(x0=np.array(x0, dtype=np.float64), x_dev=T.vector('x_dev'), lambda_dev=T.vector('lda_dev'),
                verbosity=-1, lbfgs=lbfgs, **cal, **kw)
    p.compile(nvar=prob.nvar)
    steps = []
    cur = {}
    o_dir, o_upd, o_search = p.lbfgs_dir, p.lbfgs_update, p.search

    def w_upd(x_old, x_new, g_old, g_new, zeta, S, Y, SS, L, D, fail):
        out = o_upd(x_old, x_new, g_old, g_new, zeta, S, Y, SS, L, D, fail)
        cur['upd'] = True
        return out

    def w_dir(x, s, lda, g, zeta, S, Y, SS, L, D):
        upd = cur.get('upd', False)
        cur.clear()
        cur.update(x=np.copy(x), s=np.copy(s), lda=np.copy(lda), g=np.copy(g), zeta=float(zeta), m=int(S.shape[1]),
                   mu=float(p.mu_dev.get_value()), mu_host=float(p.mu_host), nu_before=float(p.nu_dev.get_value()),
                   updated=bool(upd))
        dz = o_dir(x, s, lda, g, zeta, S, Y, SS, L, D)
        cur['dz_raw'] = np.copy(dz)
        return dz

    def w_search(x0_, s0_, lda0_, dz, a_s, a_l):
        cur['dz'] = np.copy(dz)
        cur['nu'] = float(p.nu_dev.get_value())
        cur['alpha_smax'], cur['alpha_lmax'] = float(a_s), float(a_l)
        x, s, lda = o_search(x0_, s0_, lda0_, dz, a_s, a_l)
        cur['x_new'], cur['s_new'], cur['lda_new'] = np.copy(x), np.copy(s), np.copy(lda)
        cur['signal'] = int(p.signal)
        steps.append(dict(cur))
        return x, s, lda

    p.lbfgs_dir, p.lbfgs_update, p.search = w_dir, w_upd, w_search
    x, s, lda, fval, kkt = p.solve()
    res = dict(x=np.copy(x), s=np.copy(s), lda=np.copy(lda), fval=float(fval),
               kkt=[np.atleast_1d(np.asarray(k, dtype=np.float64)) for k in kkt], signal=int(p.signal), nsteps=len(steps))
    return p, steps, res


def pack_lbfgs(prob, x0, steps, r, kw):
    out = {'x0': np.asarray(x0, dtype=np.float64), 'dims': np.array([prob.nvar, prob.neq, prob.nineq])}
    for k, v in kw.items():
        out['kw_' + k] = np.float64(v)
    out['sol0_x'], out['sol0_s'], out['sol0_lda'], out['sol0_fval'] = r['x'], r['s'], r['lda'], np.float64(r['fval'])
    for j in range(4):
        out['sol0_kkt%d' % (j + 1)] = r['kkt'][j]
    out['sol0_signal'], out['sol0_nsteps'] = np.int64(r['signal']), np.int64(r['nsteps'])
    out['nsteps'] = np.int64(len(steps))
    if steps:
        for key in ('x', 's', 'lda', 'g', 'dz_raw', 'dz', 'x_new', 's_new', 'lda_new'):
            out['st_' + key] = np.stack([st[key] for st in steps])
        for key in ('mu', 'mu_host', 'nu_before', 'nu', 'alpha_smax', 'alpha_lmax', 'zeta'):
            out['st_' + key] = np.array([st[key] for st in steps], dtype=np.float64)
        for key in ('m', 'signal', 'updated'):
            out['st_' + key] = np.array([st[key] for st in steps], dtype=np.int64)
    return out


def pack(prob, x0, steps, results, kw, keep_H=False):
    out = {'x0': np.asarray(x0, dtype=np.float64), 'dims': np.array([prob.nvar, prob.neq, prob.nineq])}
    for k, v in kw.items():
        out['kw_' + k] = np.float64(v)
    for i, r in enumerate(results):
        pre = 'sol%d_' % i
        out[pre + 'x'], out[pre + 's'], out[pre + 'lda'] = r['x'], r['s'], r['lda']
        out[pre + 'fval'] = np.float64(r['fval'])
        for j in range(4):
            out[pre + 'kkt%d' % (j + 1)] = r['kkt'][j]
        out[pre + 'signal'] = np.int64(r['signal'])
        out[pre + 'nsteps'] = np.int64(r['nsteps'])
    n = len(steps)
    out['nsteps'] = np.int64(n)
    if n:
        for key in ('x', 's', 'lda', 'g', 'dz_raw', 'dz', 'x_new', 's_new', 'lda_new'):
            out['st_' + key] = np.stack([st[key] for st in steps])
        for key in ('mu', 'mu_host', 'nu_before', 'nu', 'delta_before', 'delta', 'alpha_smax', 'alpha_lmax', 'rcond'):
            out['st_' + key] = np.array([st[key] for st in steps], dtype=np.float64)
        for key in ('n_eig', 'n_phi_search', 'signal'):
            out['st_' + key] = np.array([st[key] for st in steps], dtype=np.int64)
        out['st_nneg_first'] = np.array([st['nneg'][0] for st in steps], dtype=np.int64)
        out['st_nneg_last'] = np.array([st['nneg'][-1] for st in steps], dtype=np.int64)
        if keep_H:
            out['st_Hfull'] = np.stack([st['Hfull'] for st in steps])
            out['st_Hreg'] = np.stack([st['Hreg'] for st in steps])
    return out


def main():
    summary = []
    kw = dict(Ftol=1.0E-8)   # unit_tests.py:50 / pyipm.py:1894
    only = set(sys.argv[1:])
    if 'nlp_rankdef' in only or not only:
        # rank-deficient equality Jacobian: 4 Newton steps of the reference (it does not converge on this problem;
        # the fixture pins the rcond <= eps branch of reghess, not a solution)
        prob = problems.make_rankdef_nlp()
        kw3 = dict(kw, niter=1, miter=4)
        p, steps, results = run_reference(prob, prob.x0, **kw3)
        np.savez_compressed(os.path.join(HERE, 'ref_nlp_rankdef.npz'), **pack(prob, prob.x0, steps, results, kw3, True))
        summary.append(('nlp_rankdef', results[0]['signal'], results[0]['nsteps'], float('nan')))
        if only == {'nlp_rankdef'}:
            print(summary)
            return
    if 'lbfgs' in only or not only:
        # L-BFGS mode: the ten examples with lbfgs=4 (unit_tests.py:49) + three synthetic problems
        cases = [('example%d' % k, problems.example_problem(k)[0], problems.example_x0(k), problems.example_problem(k)[1])
                 for k in range(1, 11)]
        for name, prob in (('qp_small', problems.make_qp(D=24, M=6, nbox=8, seed=11)),
                           ('nlp_small', problems.make_nlp(D=20, M=4, N=16, seed=13)),
                           ('nlp_eqonly', problems.make_nlp(D=24, M=6, N=0, seed=15))):
            if prob.nineq == 0:
                prob.Gt = prob.Vt = prob.r = None
            cases.append((name, prob, prob.x0, None))
        for name, prob, x0, gts in cases:
            p, steps, res = run_reference_lbfgs(prob, x0, lbfgs=4, **kw)
            err = min(np.linalg.norm(res['x'] - g) for g in gts) if gts else float('nan')
            np.savez_compressed(os.path.join(HERE, 'ref_lbfgs_%s.npz' % name), **pack_lbfgs(prob, x0, steps, res, kw))
            summary.append(('lbfgs_' + name, res['signal'], res['nsteps'], err))
        if only == {'lbfgs'}:
            for row in summary:
                print('%-18s signal=%2d steps=%3d  |x-x_gt|=%.3e' % row)
            return
    for k in range(1, 11):
        prob, gts = problems.example_problem(k)
        x0 = problems.example_x0(k)
        n_solves = 2 if k == 7 else 1   # second solve() pins quirk xi (mu_dev not reset, pyipm.py:1603)
        p, steps, results = run_reference(prob, x0, n_solves=n_solves, **kw)
        err = min(np.linalg.norm(results[0]['x'] - g) for g in gts)
        np.savez_compressed(os.path.join(HERE, 'ref_example%d.npz' % k), **pack(prob, x0, steps, results, kw, True))
        summary.append(('example%d' % k, results[0]['signal'], results[0]['nsteps'], err))

    synth = [
        ('qp_small', problems.make_qp(D=24, M=6, nbox=8, seed=11), {}),
        ('qp_mid', problems.make_qp(D=96, M=24, nbox=48, seed=12), {}),
        ('nlp_small', problems.make_nlp(D=20, M=4, N=16, seed=13), {}),
        ('nlp_mid', problems.make_nlp(D=64, M=8, N=64, seed=14), {}),
        ('nlp_eqonly', problems.make_nlp(D=24, M=6, N=0, seed=15), {}),
    ]
    for name, prob, extra in synth:
        if prob.nineq == 0:
            prob.Gt = prob.Vt = prob.r = None
        kw2 = dict(kw)
        kw2.update(extra)
        p, steps, results = run_reference(prob, prob.x0, **kw2)
        np.savez_compressed(os.path.join(HERE, 'ref_%s.npz' % name),
                            **pack(prob, prob.x0, steps, results, kw2, keep_H=(prob.nvar <= 24)))
        summary.append((name, results[0]['signal'], results[0]['nsteps'], float('nan')))

    for row in summary:
        print('%-12s signal=%2d steps=%3d  |x-x_gt|=%.3e' % row)


if __name__ == '__main__':
    main()
