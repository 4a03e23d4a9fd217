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
    # ms_per_step is the time of one workload step (batch_per_gpu samples) at the measured CPU throughput, the same
    # unit the product arm reports; the CPU step itself runs `sample_batch` samples
    B = d['config']['batch_per_gpu']
    assert d['value'] > 0 and abs(d['ms_per_step'] * d['value'] - 1000.0 * B) < 1e-6 * 1000 * B
    assert d['impl_config']['sample_batch'] == d['cpu_baseline']['batch'] == 1
    assert set(d['config']) == {'workload', 'batch_per_gpu', 'global_batch', 'gflop_per_sample'}
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


def test_both_arms_describe_the_same_config():
    """The product arm's `config` and the reference arm's are built by ONE function from the same arguments (the driver
    compares them): check it for every selectable workload without running anything."""
    sys.path.insert(0, ROOT)
    import importlib
    import types
    saved = (os.dup(1),)
    try:
        bench = importlib.import_module('bench')
    finally:
        os.dup2(saved[0], 1)       # bench.py redirects fd 1 to stderr at import; give pytest its stdout back
        os.close(saved[0])
    for wl, bins, contract, gb, world in (('supervised', 5, 'B', 0, 1), ('ddd17', 5, 'B', 0, 1), ('dsec', 10, 'B', 64, 8),
                                          ('dsec', 5, 'A', 0, 1)):
        args = types.SimpleNamespace(workload=wl, batch=8, global_batch=gb, windows=20, bins=bins)
        w, scaling = bench.resolve_workload(args, world)
        cfg = bench.bench_config(w, contract, world)
        assert cfg == bench.bench_config(w, contract, world) and 'T=20' in cfg['workload']
        assert scaling == ('strong' if gb else 'weak') and cfg['global_batch'] == (gb or 8 * world)
    assert abs(bench.flops_per_sample(20, 5, 440, 640, 11, 'B') / 1e9 - 3292.3) < 0.1
    assert abs(bench.flops_per_sample(20, 5, 440, 640, 11, 'A') / 1e9 - 4098.1) < 0.1
    assert abs(bench.flops_per_sample(20, 5, 200, 346, 6, 'B') / 1e9 - 823.0) < 0.1
    assert abs(bench.flops_per_sample(20, 10, 440, 640, 11, 'B') / 1e9 - 3337.4) < 0.1
