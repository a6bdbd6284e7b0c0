"""GPU-time micro-benchmark of b200_groupnorm_bwd / b200_groupnorm_apply (CUDA-graph replays): python tools/bench_gn_bwd.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402

DEV = 'cuda'
B = int(os.environ.get('B', 128))


def timeit(fn, reps=10):
    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (5 * reps) * 1e3


def run(C, H, bf16_out=True, drop=0.1):
    W = H
    x = torch.randn(B, H, W, C, device=DEV)
    st = K.stats_from_float(torch.stack([x.sum(dim=(1, 2)), (x * x).sum(dim=(1, 2))], dim=-1).contiguous())
    g = torch.randn(B, H, W, C, device=DEV).to(torch.bfloat16)
    gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    sums = torch.empty(B, 8, C, device=DEV)
    dg, db = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dxb = torch.empty(B, H, W, C, device=DEV, dtype=torch.bfloat16)
    dx = torch.empty(B, H, W, C, device=DEV)
    out = torch.empty(B, H, W, C, device=DEV, dtype=torch.bfloat16)
    n = B * H * W * C

    def bwd():
        if bf16_out:
            K.groupnorm_bwd(g, x, C, st, None, 0, None, B, H * W, W, 32, gamma, beta, 1e-5, sums, drop_p=drop, drop_seed=1,
                            dx_bf16=dxb, dgamma=dg, dbeta=db)
        else:
            K.groupnorm_bwd(g, x, C, st, None, 0, None, B, H * W, W, 32, gamma, beta, 1e-5, sums, dx0=dx, dgamma=dg, dbeta=db)

    def fwd():
        K.groupnorm_apply(x, C, st, None, 0, None, B, H * W, W, 32, gamma, beta, 1e-5, out, drop_p=drop, drop_seed=1)
    us_b, us_f = timeit(bwd), timeit(fwd)
    by_b = n * (12 + (2 if bf16_out else 4))
    print(f'C={C:4d} @{H:2d}x{W:2d} bf16_out={int(bf16_out)} drop={drop}: bwd {us_b:7.1f} us = {by_b / us_b / 1e6:6.2f} TB/s | '
          f'fwd {us_f:6.1f} us = {n * 6 / us_f / 1e6:5.2f} TB/s', flush=True)


if __name__ == '__main__':
    print('env', {k: v for k, v in os.environ.items() if k.startswith('B200_')}, flush=True)
    run(128, 32)
    run(128, 32, drop=0.0)
    run(256, 16)
    run(256, 16, bf16_out=False, drop=0.0)
    run(256, 8)
    run(256, 4)
    run(512, 16, bf16_out=False, drop=0.0)
