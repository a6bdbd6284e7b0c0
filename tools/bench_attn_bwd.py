#!/usr/bin/env python
"""Times b200_attention_fwd_lse + b200_attention_bwd at the CFG UNet's shapes (B=128, 4 heads x 64):
python tools/bench_attn_bwd.py [B] [T] [heads] [iters].  Prints one JSON line (us per launch, TFLOP/s, GB/s)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import torch  # noqa: E402
import b200diff as K  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    heads = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    iters = int(sys.argv[4]) if len(sys.argv) > 4 else 20
    d, dev, bf = 64, 'cuda', torch.bfloat16
    C = heads * d
    nbuf = 4            # rotate over several operand sets so that the inputs do not stay in L2
    sets = []
    for i in range(nbuf):
        g = torch.Generator(device=dev).manual_seed(i)
        qk = torch.randn(B, T, 2 * C, device=dev, generator=g).to(bf)
        vt = torch.randn(B, C, T, device=dev, generator=g).to(bf)
        do = torch.randn(B, T, C, device=dev, generator=g).to(bf)
        o = torch.empty(B, T, C, device=dev, dtype=bf)
        lse = torch.empty(B, heads, T, device=dev, dtype=torch.float32)
        dqkv = torch.empty(B, T, 3 * C, device=dev, dtype=bf)
        K.attention(qk, 2 * C, 0, C, vt, o, C, B, T, heads, d, d ** -0.5, lse=lse)
        sets.append((qk, vt, o, do, lse, dqkv))

    def run(i):
        qk, vt, o, do, lse, dqkv = sets[i % nbuf]
        K.attention_bwd(qk, vt, o, do, lse, dqkv[:, :, :2 * C], dqkv[:, :, 2 * C:], B, T, heads, d, d ** -0.5)

    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(iters):
        run(i)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / iters
    flops = 10.0 * B * heads * T * T * d
    nbytes = B * T * C * 2 * (2 + 1 + 1 + 1 + 3)        # q|k, v, o, dO read; dq|dk|dv written
    print(json.dumps(dict(B=B, T=T, heads=heads, us=us, tflops=flops / us * 1e-6, gbs=nbytes / us * 1e-3)))


if __name__ == '__main__':
    main()
