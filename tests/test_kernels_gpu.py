"""Per-kernel parity on a B200: every C-ABI entry point against a PyTorch fp32 restatement (tests/kernel_cases.py).
Each case runs in its own process so that a faulting kernel cannot poison the CUDA context of the others."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CASES = ['misc', 'groupnorm', 'conv_basic', 'conv_epilogue', 'conv_n256', 'conv_small_hw', 'conv_1x1',
         'conv_shortcut', 'conv_stride2', 'conv_lastconv', 'conv_up2', 'conv_tproj', 'conv_gnfuse', 'conv_gnfuse_out', 'first_conv_tc', 'attention', 'attention_bwd', 'attn_block', 'sampler',
         'sampler_cfg', 'ddim_inversion', 'precise',
         'sampler_large', 'gemm',
         'wgrad', 'groupnorm_bwd', 'backward_misc', 'optimizer', 'pack_weights', 'ode_samplers']


@pytest.mark.gpu
@pytest.mark.parametrize('case', CASES)
def test_kernel_case(case):
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'tests', 'kernel_cases.py'), case], capture_output=True,
                       text=True, timeout=600)
    assert r.returncode == 0, f'{case} failed:\n{r.stdout[-3000:]}\n{r.stderr[-2000:]}'
