"""CPU suite, part 6: the bench.py output contract on the arm that needs no GPU
(``--impl reference``: the CPU oracle through the same driver, tiny mesh so it takes a second):
exactly ONE JSON line on stdout with the keys the driver reads."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--mesh', '6', '--steps', '2', '--warmup', '1'], capture_output=True,
                         text=True, timeout=300, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [ln for ln in out.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ('metric', 'value', 'unit', 'n_gpus', 'steps', 'warmup', 'ms_per_step',
                'higher_is_better', 'scaling', 'vs_baseline', 'dtype', 'data', 'config', 'impl',
                'cpu_baseline', 'e2e'):
        assert key in d, key
    assert d['impl'] == 'reference' and d['metric'] == 'dre_backward_steps_per_s'
    assert d['unit'] == 'steps/s' and d['higher_is_better'] is True and d['dtype'] == 'f64'
    assert d['steps'] == 2 and d['warmup'] == 1 and d['value'] > 0
    assert d['vs_baseline'] is None and d['data'] == 'synthetic'
    assert 'workload' in d['config'] and 'model' not in d['config']
    cb = d['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == d['value'] and cb['sample']
    assert d['e2e'] == dict(value=d['value'], unit='steps/s', h2d_bytes_per_step=0,
                            d2h_bytes_per_step=0)


def test_reference_arm_other_ranks_are_silent():
    env = dict(os.environ, RANK='1', WORLD_SIZE='2')
    out = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference',
                          '--mesh', '6', '--steps', '1', '--warmup', '0', '--gpus', '2'],
                         capture_output=True, text=True, timeout=120, cwd=ROOT, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ''
