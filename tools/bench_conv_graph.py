"""GPU-time micro-benchmark of b200_conv2d_fwd on small layers: 20 launches captured in a CUDA graph (no host launch
cost in the timing), CUDA events around graph replays.  Usage: B=256 python tools/bench_conv_graph.py"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402

DEV = 'cuda'
B = int(os.environ.get('B', 256))
REPLAYS = int(os.environ.get('REPLAYS', 5))
ONLY = os.environ.get('ONLY')


def run(name, Cin, Cout, H, k=3, residual=False, out_mode=K.OUT_F32_NHWC, reps=20):
    if ONLY and ONLY not in name:
        return
    W = H
    a0 = torch.randn(B, H, W, Cin, device=DEV).to(torch.bfloat16)
    w = K.pack_weight(torch.randn(Cout, Cin, k, k, device=DEV) / math.sqrt(k * k * Cin))
    taps = K.taps_3x3_s1() if k == 3 else K.taps_1x1()
    if out_mode == K.OUT_F32_NHWC:
        out = torch.empty(B, H, W, Cout, device=DEV)
    elif out_mode == K.OUT_BF16_NHWC:
        out = torch.empty(B, H, W, Cout, device=DEV, dtype=torch.bfloat16)
    else:
        out = torch.empty(B, Cout, H, W, device=DEV, dtype=torch.bfloat16)
    res = torch.randn(B, H, W, Cout, device=DEV) if residual else None
    st = K.new_stats(B, Cout, DEV) if out_mode == K.OUT_F32_NHWC else None
    bias = torch.randn(Cout, device=DEV)

    def call():
        K.conv2d(a0, w, Cout, B, H, W, taps, a0_geom=(Cin, H, W, 1), bias=bias, residual=res, res_ld=Cout, out=out,
                 out_mode=out_mode, stats=st)
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
    for _ in range(REPLAYS):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (REPLAYS * reps) * 1e3
    macs = 1.0 * B * H * W * Cout * k * k * Cin
    print(f'{name:44s} {us:8.1f} us  {2 * macs / us / 1e6:7.1f} TFLOP/s', flush=True)


if __name__ == '__main__':
    print('env', {k: v for k, v in os.environ.items() if k.startswith('B200_')}, flush=True)
    run('3x3 256->256 @4 +res', 256, 256, 4, residual=True)
    run('3x3 512->256 @4', 512, 256, 4)
    run('3x3 256->256 @8 +res', 256, 256, 8, residual=True)
    run('3x3 512->256 @8', 512, 256, 8)
    run('3x3 256->256 @16 +res', 256, 256, 16, residual=True)
    run('3x3 128->128 @32 +res', 128, 128, 32, residual=True)
    run('1x1 256->512 @16 bf16 (qk)', 256, 512, 16, k=1, out_mode=K.OUT_BF16_NHWC)
    run('1x1 256->256 @16 bf16 NCHW (v^T)', 256, 256, 16, k=1, out_mode=K.OUT_BF16_NCHW)
    run('1x1 256->256 @16 +res (proj)', 256, 256, 16, k=1, residual=True)
