"""cProfile of the host side of the training step (where do the ~28 ms of enqueue time go?)."""
import cProfile
import os
import pstats
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import diffusions  # noqa: E402
import models  # noqa: E402
from b200diff.optim import FusedAdam  # noqa: E402
from b200diff.train import TrainStep  # noqa: E402
from tools.bench_train import CFGC  # noqa: E402

torch.manual_seed(2022)
model = models.UNetCategorialAdaGN(**CFGC).cuda().train()
diffuser = diffusions.DDPM(total_steps=1000, beta_schedule='cosine', device='cuda')
step = TrainStep(model, diffuser, FusedAdam(model.parameters(), lr=2e-4), ema=models.EMA(model.parameters()), clip_grad_norm=1.0)
x0 = (torch.randn(128, 3, 32, 32) * 0.5).clamp(-1, 1).cuda()
y = (torch.arange(128) % 10).cuda()
for _ in range(3):
    step(x0, y=y)
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for _ in range(5):
    step(x0, y=y)
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats('tottime').print_stats(28)
