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


def test_clock_probe_repeat_count_is_rank_independent():
    """The probe count depends only on the all-reduced time and the step count (a lambda-sharded step
    holds a collective: ranks must agree on how often they run it)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(os.path.dirname(__file__), '..', 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.clock_probe_repeats(5000.0, 20) == 0
    for total, steps in ((12.0, 20), (2.6, 20), (575.0, 5), (150.0, 5), (0.0, 3)):
        n = bench.clock_probe_repeats(total, steps)
        assert n >= 20 and n % 20 == 0 and n <= 20000
    assert bench.clock_probe_repeats(12.0, 20) == bench.clock_probe_repeats(12.0, 20)


def test_both_arms_describe_the_same_config():
    """`config` is a pure function of the workload: the reference arm, which times a bounded sample of a
    column stack, must print exactly the `config` our arm prints (the driver compares them)."""
    import importlib.util
    spec = importlib.util.spec_from_file_location('bench_mod', os.path.join(ROOT, 'bench.py'))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ours, d0 = bench.build_workload('c3', 8, rank=0, world=1)
    sample, d1 = bench.build_workload('c3', 4, for_gpu=False, desc_columns=8)
    assert bench.workload_config('c3', d0, ours, 8) == bench.workload_config('c3', d1, sample, 8)
    a, da = bench.build_workload('c2', 4096)
    assert bench.workload_config('c2', da, a, 4096)['Ncolumns'] == 1
