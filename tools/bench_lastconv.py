"""GPU-time of the pixel-major last conv (conv_gemm_pixm_kernel) at the families' sizes: python tools/bench_lastconv.py
(B200_PIXM_VTAP=0 selects the plain 9-tap tiles).  CUDA graph of 10 launches over rotating inputs, CUDA events."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402

DEV = 'cuda'


def run(B, Cin, Cout, H):
    copies = max(2, int(300e6 // (B * H * H * Cin * 2)) + 1)
    xs = [torch.randn(B, H, H, Cin, device=DEV).to(torch.bfloat16) for _ in range(copies)]
    w = torch.randn(Cout, Cin, 3, 3, device=DEV) / math.sqrt(9 * Cin)
    wp = K.pack_weight(w)
    b = torch.randn(Cout, device=DEV)
    out = torch.empty(B, Cout, H, H, device=DEV)

    def call(x):
        K.conv2d(x, wp, Cout, B, H, H, K.taps_3x3_s1(), a0_geom=(Cin, H, H, 1), bias=b, out=out, out_mode=K.OUT_F32_NCHW)
    call(xs[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for x in xs:
            call(x)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * copies) * 1e3
    print(f'B={B:3d} {Cin}->{Cout} @{H}x{H}: {us:8.1f} us  (input {B * H * H * Cin * 2 / 1e6:.0f} MB: '
          f'{B * H * H * Cin * 2 / us / 1e6:.2f} TB/s of unique input)', flush=True)


if __name__ == '__main__':
    print('env', {k: v for k, v in os.environ.items() if k.startswith('B200_')}, flush=True)
    run(256, 128, 3, 32)      # CIFAR-10 UNet
    run(32, 128, 3, 256)      # pesser CelebA-HQ 256
    run(16, 256, 6, 256)      # ADM ImageNet-256
