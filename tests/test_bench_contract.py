"""bench.py contract checks that need no GPU: the reference arm (`--impl reference`, the CPU restatement timed on the
host cores) prints exactly one JSON line with the keys the driver parses, and the product arm refuses to run without
a CUDA device (no CPU fallback)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py')] + args, capture_output=True, text=True,
                          timeout=900, env=e, cwd=ROOT)


def test_reference_arm_prints_one_json_line():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0'])
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['unit'] == 'samples/s' and d['higher_is_better'] is True
    assert d['value'] > 0 and abs(d['ms_per_step'] * d['value'] - 1000.0) < 1e-6 * 1000
    assert d['cpu_baseline']['kind'] == 'port' and d['cpu_baseline']['cores'] >= 1 and d['cpu_baseline']['sample']
    assert d['e2e'] == dict(value=d['value'], unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config'] and d['vs_baseline'] is None and d['data'] == 'synthetic'


def test_reference_arm_other_ranks_exit_quietly():
    r = _run(['--impl', 'reference', '--steps', '1', '--warmup', '0', '--gpus', '2'], env={'RANK': '1', 'WORLD_SIZE': '2'})
    assert r.returncode == 0 and r.stdout.strip() == ''


def test_product_arm_needs_cuda():
    import torch
    if torch.cuda.is_available():
        return
    r = _run(['--steps', '1', '--warmup', '3'])
    assert r.returncode != 0 and r.stdout.strip() == ''
    assert 'CUDA' in r.stderr or 'cuda' in r.stderr
