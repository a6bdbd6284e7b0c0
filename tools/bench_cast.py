"""GPU-time micro-benchmark of b200_cast_bf16 (fp32 NHWC -> bf16 NHWC / 4 parity planes) at the CIFAR-10 forward's sizes:
python tools/bench_cast.py   (B200_CAST_C8=0 selects the generic kernel).  Rotating buffers (> 400 MB between reuses),
one CUDA graph per case, CUDA events around 5 replays."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402

DEV = 'cuda'


def run(B, H, C, parity):
    n = B * H * H * C
    copies = max(2, int(400e6 // (4 * n)) + 1)
    xs = [torch.randn(B, H, H, C, device=DEV) for _ in range(copies)]
    shape = (B, 4, H // 2, H // 2, C) if parity else (B, H, H, C)
    outs = [torch.empty(shape, device=DEV, dtype=torch.bfloat16) for _ in range(copies)]
    K.cast_bf16(xs[0], outs[0], B, H, H, C, parity_split=parity)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for x, o in zip(xs, outs):
            K.cast_bf16(x, o, B, H, H, C, parity_split=parity)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * copies) * 1e3
    print(f'B={B:3d} {H:3d}x{H:<3d} C={C:4d} parity={int(parity)}: {us:7.1f} us = {6.0 * n / us / 1e6:5.2f} TB/s '
          f'({6.0 * n / 1e6:.0f} MB)', flush=True)


if __name__ == '__main__':
    print('env', {k: v for k, v in os.environ.items() if k.startswith('B200_')}, flush=True)
    for (B, H, C) in ((256, 32, 128), (256, 16, 256), (256, 8, 256), (16, 256, 128), (32, 128, 256)):
        run(B, H, C, True)
        run(B, H, C, False)
