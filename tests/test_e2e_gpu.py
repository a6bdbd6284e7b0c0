"""End-to-end parity on a B200 (tests/e2e_cases.py): UNet forward eps rel-L2 <= 1e-2, DDIM-50 PSNR >= 40 dB against
the fp32 oracle, graph replay vs eager loop, the smoke() entry point."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.parametrize('case', ['unet_forward', 'ddim50', 'ddpm_noise', 'cfg', 'families_golden', 'adm256', 'pesser256',
                                  'train_step', 'train_step_b128', 'train_step_pesser', 'train_step_adm', 'train_multi_step', 'ode_sampling',
                                  'ddim_inversion', 'engine_hygiene', 'fp32_mode'])
def test_e2e_case(case):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'e2e_cases.py'), case], capture_output=True,
                       text=True, timeout=900)
    assert r.returncode == 0, f'{case} failed:\n{r.stdout[-3000:]}\n{r.stderr[-2000:]}'


@pytest.mark.gpu
def test_smoke_entry():
    r = subprocess.run([sys.executable, '-c', 'import __graft_entry__ as g; g.smoke()'], cwd=ROOT,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]


@pytest.mark.gpu
def test_reference_scripts_replay_under_torchrun():
    """scripts/sample_uncond.py:115-195 and scripts/train_ddpm.py:100-216 call sequences (tests/script_replay.py) through the
    omegaconf / accelerate stand-ins, one process per GPU under torchrun: 2 ranks (NCCL gather, DDP gradient all-reduce)
    when the box has >= 2 GPUs, else 1 rank."""
    import socket
    import torch
    n = min(2, torch.cuda.device_count())
    assert n >= 1
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        port = s.getsockname()[1]
    r = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}',
                        '--master-addr', '127.0.0.1', '--master-port', str(port),
                        os.path.join(ROOT, 'tests', 'script_replay.py')], capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, f'script replay failed ({n} ranks):\n{r.stdout[-3000:]}\n{r.stderr[-3000:]}'
    assert '"ok": true' in r.stdout
