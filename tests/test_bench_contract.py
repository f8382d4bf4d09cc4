"""CPU test of the measurement contract: the reference arm of bench.py (the oracle port timed on the host cores - the only
other place that may execute oracle/) prints one JSON line with the keys the driver reads, and the B200 arm refuses to run
without a GPU instead of falling back to anything."""
import json
import os
import subprocess
import sys

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args):
    return subprocess.run([sys.executable, os.path.join(REPO, 'bench.py'), *args], capture_output=True, text=True, timeout=600,
                          cwd=REPO)


def test_reference_arm_prints_one_contract_line():
    r = run_bench('--impl', 'reference', '--steps', '1', '--warmup', '1', '--cpu-seconds', '1')
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    j = json.loads(lines[0])
    assert j['impl'] == 'reference' and j['metric'] == '2-line pixels/s' and j['unit'] == 'pixels/s'
    assert j['higher_is_better'] is True and j['n_gpus'] == 1 and j['steps'] == 1 and j['warmup'] == 1
    assert j['value'] > 0 and j['ms_per_step'] > 0
    assert 'workload' in j['config'] and '256x256' in j['config']['workload']
    cb = j['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == j['value'] and cb['sample']
    e = j['e2e']
    assert e['value'] == j['value'] and e['h2d_bytes_per_step'] == 0 and e['d2h_bytes_per_step'] == 0


def test_b200_arm_needs_a_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    r = run_bench('--steps', '1', '--warmup', '1', '--no-cpu-baseline')
    assert r.returncode != 0                       # no CPU fallback: the run fails loudly
    assert not [l for l in r.stdout.splitlines() if l.startswith('{')]
