#!/usr/bin/env python
"""Forward-pass timing of the larger model families on one B200 (orientation + per-kernel breakdown):
    python tools/bench_family.py adm256 [B]     ADM ImageNet-256 class-cond UNet   (2239.67 GF/img/fwd, B=16)
    python tools/bench_family.py pesser256 [B]  pesser CelebA-HQ 256 UNet          (497.03 GF/img/fwd, B=32)
    python tools/bench_family.py cfg [B]        UNetCategorialAdaGN CIFAR-10       (14.396 GF/img/fwd, B=128)
Prints one JSON line: ms per forward (CUDA events, eager launches and CUDA-graph replay), TFLOP/s (algorithmic FLOPs
of the reference network), fraction of the measured bf16 peak, per-kernel totals."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import b200diff as K  # noqa: E402
import models  # noqa: E402
from tests.e2e_cases import ADM256, PESSER256  # noqa: E402

CFGC = dict(in_channels=3, out_channels=3, dim=128, dim_mults=[1, 2, 2, 2], use_attn=[False, True, True, False],
            num_res_blocks=2, num_classes=10, attn_head_dims=64, resblock_updown=True, dropout=0.1)


def main():
    which = sys.argv[1]
    dev = 'cuda'
    torch.manual_seed(2022)
    if which == 'adm256':
        from models.adm.unet import UNetModel
        from oracle.adm_ref import randomize_zero_params
        B, gf, res = 16, 2239.67, 256
        m = UNetModel(**ADM256)
        m.load_state_dict(randomize_zero_params(m.state_dict()))
        kw = dict(y=(torch.arange(B) % 1000).to(dev))
    elif which == 'pesser256':
        from models.pesser.model import Model
        B, gf, res = 32, 497.03, 256
        m = Model(**PESSER256)
        kw = {}
    else:
        B, gf, res = 128, 14.396, 32
        m = models.UNetCategorialAdaGN(**CFGC)
        kw = dict(y=(torch.arange(B) % 10).to(dev))
    if len(sys.argv) > 2:
        B = int(sys.argv[2])
        if 'y' in kw:
            kw['y'] = kw['y'][:1].expand(B).contiguous() if B > kw['y'].shape[0] else kw['y'][:B].contiguous()
    m = m.to(dev).eval()
    x = torch.randn(B, 3, res, res, device=dev)
    t = torch.full((1,), 500, device=dev, dtype=torch.long).expand(B)
    out = torch.empty(B, m.out_channels, res, res, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    with torch.no_grad():
        for _ in range(2):
            m(x, t, out=out, **kw)
        torch.cuda.synchronize()
        mem = torch.cuda.max_memory_allocated() / 2**30
        n0 = K.direct_launch_count()
        e0, e1 = ev(), ev()
        e0.record()
        for _ in range(3):
            m(x, t, out=out, **kw)
        e1.record()
        torch.cuda.synchronize()
        ms_eager = e0.elapsed_time(e1) / 3
        launches = (K.direct_launch_count() - n0) // 3
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m(x, t, out=out, **kw)
        g.replay()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        ms_graph = e0.elapsed_time(e1) / 5
        with K.Profiler() as prof:
            for _ in range(2):
                m(x, t, out=out, **kw)
        kern = prof.summary()
    peak = 1387.7
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peak = json.load(open(p)).get('bf16_tflops_sustained', peak)
    tf = gf * 1e9 * B / (ms_graph * 1e-3) / 1e12
    print(json.dumps({
        'model': which, 'batch': B, 'ms_per_forward_eager': ms_eager, 'ms_per_forward_graph': ms_graph,
        'kernels_per_forward': launches, 'gflop_per_image': gf, 'tflops': tf, 'frac_of_bf16_sustained_peak': tf / peak,
        'images_fwd_per_s': B / (ms_graph * 1e-3), 'peak_mem_gib': mem,
        'kernels': {k: {'n': v['n'] // 2, 'ms': v['ms'] / 2,
                        **({'tflops': v['flops'] / (v['ms'] * 1e-3) / 1e12} if v['flops'] else {}),
                        **({'gbs': v['bytes'] / (v['ms'] * 1e-3) / 1e9} if v['bytes'] else {})}
                    for k, v in sorted(kern.items(), key=lambda kv: -kv[1]['ms'])}}), flush=True)


if __name__ == '__main__':
    main()
