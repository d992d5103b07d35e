"""bench.py's reference arm runs on the CPU: check the JSON contract of its one stdout line."""
import json
import os
import subprocess
import sys

import pytest

from oracle import reflib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.ref
def test_reference_arm_prints_one_contract_line():
    if not reflib.available():
        pytest.skip('compiled reference not available')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--workload', 'c1',
                        '--steps', '2', '--warmup', '1'], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'points/s' and d['higher_is_better'] is True
    for key in ('metric', 'value', 'n_gpus', 'steps', 'warmup', 'ms_per_step', 'scaling', 'vs_baseline', 'dtype',
                'data', 'config', 'cpu_baseline', 'e2e', 'gpu_launches'):
        assert key in d, key
    assert d['value'] > 0 and d['cpu_baseline']['kind'] in ('reference', 'port') and d['cpu_baseline']['cores'] >= 1
    assert d['e2e']['h2d_bytes_per_step'] == 0 and d['e2e']['d2h_bytes_per_step'] == 0
    assert 'workload' in d['config'] and d['vs_baseline'] is None


def test_gpu_arm_refuses_to_run_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--steps', '1', '--warmup', '0'],
                       capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode != 0 and 'no CPU fallback' in (r.stderr + r.stdout)
