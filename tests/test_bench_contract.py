"""bench.py's reference arm (the oracle port timed on host cores) must print ONE JSON line with the keys the driver
reads, on CPU, without touching CUDA; under torchrun only rank 0 prints."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(env_extra=None):
    env = dict(os.environ)
    env.pop('RANK', None)
    env.update(env_extra or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'bench.py'), '--impl', 'reference', '--steps', '1',
                        '--warmup', '1'], capture_output=True, text=True, timeout=600, env=env, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    return [l for l in r.stdout.splitlines() if l.startswith('{')]


def test_reference_arm_line():
    lines = _run()
    assert len(lines) == 1
    line = json.loads(lines[0])
    assert line['impl'] == 'reference' and line['metric'] == 'ddim50_cifar10_images_per_s'
    assert line['unit'] == 'images/s' and line['higher_is_better'] is True and line['n_gpus'] == 1
    assert line['value'] > 0 and line['ms_per_step'] > 0
    assert line['config']['workload'].startswith('DDIM-50 CIFAR-10 32x32 UNet')
    cb = line['cpu_baseline']
    assert cb['kind'] == 'port' and cb['cores'] >= 1 and cb['value'] == line['value'] and cb['sample']
    assert line['e2e'] == {'value': line['value'], 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
    assert line['gpu_launches'] == 0


def test_reference_arm_other_ranks_are_silent():
    assert _run({'RANK': '1', 'WORLD_SIZE': '2'}) == []
