"""bench.py contract on a machine without a GPU: the reference arm runs the CPU oracle on a bounded sample and prints ONE
JSON line with the keys the driver reads; the B200 arm refuses to run without a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    # --ref-test-size: the same code path on a toy instance (the real arm runs ONE full config-3 step: minutes of CPU)
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '10', '--warmup', '3',
                          '--ref-test-size', '128', '--ref-skip-sample'], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip().startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'newton_steps_per_sec' and d['unit'] == 'steps/s'
    assert d['higher_is_better'] is True and d['vs_baseline'] is None and d['dtype'] == 'f64' and d['data'] == 'synthetic'
    assert d['value'] > 0 and d['steps'] == 1 and d['warmup'] == 0 and d['steps_requested'] == 10      # one real step, said so
    assert 'workload' in d['config'] and 'config3' in d['config']['workload']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and 'sample' in cb
    assert d['e2e'] == {'value': d['value'], 'unit': 'steps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}


def test_b200_arm_refuses_to_run_without_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '3'],
                         capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode != 0
    assert 'no CUDA device' in (out.stderr + out.stdout)
