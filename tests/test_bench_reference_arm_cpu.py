"""CPU: `bench.py --impl reference` (the driver's baseline arm: the reference's CPU implementation of the path, here the
oracle port) prints ONE JSON line with the contract's keys and needs no GPU."""
import json
import os
import subprocess
import sys

from conftest import ROOT


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '0'],
                         capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.strip().splitlines() if l.startswith('{')]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'samples/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and d['steps'] == 1 and d['n_gpus'] == 1
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['value'] == d['value']
    assert d['e2e'] == dict(value=d['value'], unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config'] and d['gpu_launches'] == 0
