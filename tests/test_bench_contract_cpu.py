"""bench.py contract pieces that need no GPU: the reference arm (`--impl reference`) prints ONE JSON line with the
keys the driver reads; non-zero ranks of a torchrun launch print nothing."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra):
    env = dict(os.environ, ARDAE_BENCH_CPU_ROWS='4', **env_extra)
    return subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--gpus', '2',
                           '--steps', '1', '--warmup', '1'], capture_output=True, text=True, env=env, timeout=600)


def test_reference_arm_prints_one_json_line():
    r = _run({'RANK': '0', 'OMP_NUM_THREADS': '1'})
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d['impl'] == 'reference' and d['metric'] == 'train_samples_per_sec' and d['unit'] == 'samples/s'
    assert d['higher_is_better'] is True and d['n_gpus'] == 2 and d['value'] > 0
    cb = d['cpu_baseline']
    assert cb['kind'] in ('reference', 'port') and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == dict(value=d['value'], unit='samples/s', h2d_bytes_per_step=0, d2h_bytes_per_step=0)
    assert 'workload' in d['config']


def test_reference_arm_other_ranks_are_silent():
    r = _run({'RANK': '1'})
    assert r.returncode == 0 and r.stdout.strip() == ''
