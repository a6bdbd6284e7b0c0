"""Profiling driver: a few training steps of the CFG (AdaGN) UNet at batch 128 so that ncu can list / capture the
backward kernels.  Usage: python tools/profile_train.py [B] [steps]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
sys.path.insert(0, ROOT)
import b200diff as K  # noqa: E402
import diffusions  # noqa: E402
import models  # noqa: E402
from b200diff.optim import FusedAdam  # noqa: E402
from b200diff.train import TrainStep  # noqa: E402
from tools.bench_train import CFGC  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
N = int(sys.argv[2]) if len(sys.argv) > 2 else 3
torch.manual_seed(2022)
model = models.UNetCategorialAdaGN(**CFGC).cuda().train()
diffuser = diffusions.DDPM(total_steps=1000, beta_schedule='cosine', device='cuda')
step = TrainStep(model, diffuser, FusedAdam(model.parameters(), lr=2e-4), ema=models.EMA(model.parameters()),
                 clip_grad_norm=1.0)
x0 = (torch.randn(B, 3, 32, 32) * 0.5).clamp(-1, 1).cuda()
y = (torch.arange(B) % 10).cuda()
for i in range(N):
    n0 = K.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    loss = step(x0, y=y)
    e1.record()
    torch.cuda.synchronize()
    print(f'step {i}: {e0.elapsed_time(e1):.3f} ms, {K.launch_count() - n0} launches, loss {loss.item():.4f}', flush=True)
