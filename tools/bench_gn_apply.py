"""GPU-time micro-benchmark of b200_groupnorm_apply_fwd at the sizes of the model families (inputs > L2 unless noted):
python tools/bench_gn_apply.py.  Each case: rotating over enough distinct buffers that every launch streams from HBM."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402

DEV = 'cuda'


def run(B, H, C, bf16_in=False, scale_shift=False, copies=None):
    W = H
    n = B * H * W * C
    in_bytes = n * (2 if bf16_in else 4)
    copies = copies or max(2, int(400e6 // in_bytes) + 1)   # > 400 MB of distinct input between reuses
    xs = [torch.randn(B, H, W, C, device=DEV) for _ in range(copies)]
    st = K.stats_from_float(torch.stack([xs[0].sum(dim=(1, 2)), (xs[0] * xs[0]).sum(dim=(1, 2))], dim=-1).contiguous())
    if bf16_in:
        xs = [x.to(torch.bfloat16) for x in xs]
    outs = [torch.empty(B, H, W, C, device=DEV, dtype=torch.bfloat16) for _ in range(copies)]
    gamma, beta = torch.ones(C, device=DEV), torch.zeros(C, device=DEV)
    ss = torch.randn(B, 2 * C, device=DEV) * 0.1 if scale_shift else None

    def call(i):
        K.groupnorm_apply(xs[i], C, st, None, 0, None, B, H * W, W, 32, gamma, beta, 1e-5, outs[i],
                          scale=None if ss is None else ss, shift=None if ss is None else ss[:, C:], ss_ld=2 * C)
    call(0)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(copies):
            call(i)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * copies) * 1e3
    by = in_bytes + 2 * n
    print(f'B={B:3d} {H:3d}x{W:3d} C={C:4d} in={"bf16" if bf16_in else "fp32"} ss={int(scale_shift)}: {us:8.1f} us = '
          f'{by / us / 1e6:5.2f} TB/s  ({by / 1e6:.0f} MB)', flush=True)


if __name__ == '__main__':
    print('env', {k: v for k, v in os.environ.items() if k.startswith('B200_')}, flush=True)
    for bf in (False, True):
        run(16, 256, 256, bf)
        run(16, 128, 256, bf)
        run(16, 64, 512, bf, scale_shift=True)
        run(16, 32, 512, bf)
        run(256, 32, 128, bf)
        run(256, 16, 256, bf)
        run(256, 8, 256, bf)
