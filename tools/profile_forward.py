"""Profiling driver: a few eager UNet forwards (CIFAR-10 config) so that ncu can list every kernel launch.
Usage: python tools/profile_forward.py [B] [n_forwards]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402
import models  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(2022)
m = models.UNet().cuda().eval()
x = torch.randn(B, 3, 32, 32, device='cuda')
t = torch.full((1,), 500, device='cuda').expand(B)
with torch.no_grad():
    for i in range(N):
        n0 = K.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = m(x, t)
        e1.record()
        torch.cuda.synchronize()
        print(f'forward {i}: {e0.elapsed_time(e1):.3f} ms, {K.launch_count() - n0} launches', flush=True)
print('out absmax', out.abs().max().item())
