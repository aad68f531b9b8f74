"""Shared helpers for the test-suite (fixtures loading, problem registry)."""
import os

import numpy as np

from pyipm_b200 import problems

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')

SYNTH = {
    'qp_small': lambda: problems.make_qp(D=24, M=6, nbox=8, seed=11),
    'qp_mid': lambda: problems.make_qp(D=96, M=24, nbox=48, seed=12),
    'nlp_small': lambda: problems.make_nlp(D=20, M=4, N=16, seed=13),
    'nlp_mid': lambda: problems.make_nlp(D=64, M=8, N=64, seed=14),
    'nlp_eqonly': lambda: _eqonly(problems.make_nlp(D=24, M=6, N=0, seed=15)),
}
EXAMPLES = ['example%d' % k for k in range(1, 11)]
ALL_GOLDEN = EXAMPLES + sorted(SYNTH)


def _eqonly(p):
    p.Gt = p.Vt = p.r = None
    return p


def get_problem(name):
    """-> (problem, x0, ground_truths or None)"""
    if name.startswith('example'):
        k = int(name[len('example'):])
        p, gts = problems.example_problem(k)
        return p, problems.example_x0(k), gts
    p = SYNTH[name]()
    return p, p.x0, None


def load_golden(name):
    return np.load(os.path.join(GOLDEN, 'ref_%s.npz' % name))
