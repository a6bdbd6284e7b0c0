"""GPU-time micro-benchmark of b200_conv2d_wgrad (two-phase split-K) at the CFG training step's layer shapes, batch 128:
CUDA-graph replays, CUDA events.  python tools/bench_wgrad.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402

DEV = 'cuda'
B = int(os.environ.get('B', 128))
ONLY = os.environ.get('ONLY')


def run(Cin, Cout, H, k=3, reps=10):
    name = f'{k}x{k} {Cin}->{Cout} @{H}'
    if ONLY and ONLY not in name:
        return
    W = H
    x = torch.randn(B, H, W, Cin, device=DEV).to(torch.bfloat16)
    dy = torch.randn(B, H, W, Cout, device=DEV).to(torch.bfloat16)
    dw = torch.zeros(Cout, Cin, k, k, device=DEV)
    scratch = torch.empty(64 << 20, dtype=torch.float32, device=DEV)
    taps = (K.taps_3x3_s1() if k == 3 else K.taps_1x1())[0]

    def call():
        K.conv2d_wgrad(dy, Cout, x, (Cin, H, W, 1), B, H, W, Cout, Cin, taps, dw, scratch=scratch)
    call()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            call()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * reps) * 1e3
    fl = 2.0 * B * H * W * Cout * Cin * k * k
    print(f'wgrad {name:24s} {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    print('env', {k: v for k, v in os.environ.items() if k.startswith('B200_')}, flush=True)
    run(128, 128, 32)
    run(256, 128, 32)
    run(384, 128, 32)
    run(256, 256, 16)
    run(512, 256, 16)
    run(256, 256, 8)
    run(512, 256, 8)
    run(256, 256, 4)
    run(512, 256, 16, k=1)
