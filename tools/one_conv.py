"""One conv shape, a few launches (ncu target): python tools/one_conv.py Cin Cout H [k] [B]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402

Cin, Cout, H = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
k = int(sys.argv[4]) if len(sys.argv) > 4 else 3
B = int(sys.argv[5]) if len(sys.argv) > 5 else 256
a0 = torch.randn(B, H, H, Cin, device='cuda').to(torch.bfloat16)
w = K.pack_weight(torch.randn(Cout, Cin, k, k, device='cuda') / math.sqrt(k * k * Cin))
out = torch.empty(B, H, H, Cout, device='cuda')
res = torch.randn(B, H, H, Cout, device='cuda')
st = K.new_stats(B, Cout, 'cuda')
bias = torch.randn(Cout, device='cuda')
taps = K.taps_3x3_s1() if k == 3 else K.taps_1x1()
for _ in range(6):
    K.conv2d(a0, w, Cout, B, H, H, taps, a0_geom=(Cin, H, H, 1), bias=bias, residual=res, res_ld=Cout, out=out, stats=st)
torch.cuda.synchronize()
# reference check against cuDNN on the bf16-rounded operands
import torch.nn.functional as F
wr = w.float().view(Cout, k, k, Cin).permute(0, 3, 1, 2)
ref = F.conv2d(a0.float().permute(0, 3, 1, 2), wr, bias, padding=k // 2).permute(0, 2, 3, 1) + res
print('max abs err', (out - ref).abs().max().item(), 'ref absmax', ref.abs().max().item())
