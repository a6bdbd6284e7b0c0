"""GPU-time micro-benchmark of K4 (b200_sampler_step) against the HBM roofline: python tools/bench_sampler.py [--json F].
SURVEY §8(d): at CIFAR-10 size the update moves 9-13 MB and is launch-latency bound, so GB/s is also reported at
synthetic sizes whose tensors exceed the 126 MB L2.  Every case rotates over enough distinct buffer sets that each launch
streams from HBM, is captured into one CUDA graph and timed with CUDA events around 5 replays.
Algorithmic bytes per element = 4 B x (model_out + x_t [+ noise] [+ uncond model_out] [+ variance channels] + outputs)."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402
import diffusions  # noqa: E402

DEV = 'cuda'


def _peak():
    try:
        return float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs']), 'MEASURED_PEAKS.json'
    except Exception:  # noqa: BLE001
        return 6454.3, 'fallback (round-1 measured copy bandwidth)'


def run(name, B, C, H, *, kind='ddim', var_type='fixed_large', cfg=False, extra_outputs=(), t=500, t_prev=480):
    HW = H * H
    n = B * C * HW
    learned = var_type == 'learned_range'
    Cm = 2 * C if learned else C
    if kind == 'ddim':
        d = diffusions.DDIM(total_steps=1000, respace_type='uniform', respace_steps=50, eta=0.0, device=DEV)
    else:
        d = diffusions.DDPM(total_steps=1000, var_type=var_type, respace_type='uniform', respace_steps=50, device=DEV)
    row = d._coef_row(t, t_prev)
    streams = 2 + 1 + int(cfg) + int(learned) + 1 + len(extra_outputs)  # mo, xt, noise, [mo_u], [logvar], sample, ...
    by = 4.0 * n * streams
    copies = max(2, int(600e6 // by) + 1)
    sets = []
    for i in range(copies):
        g = torch.Generator(device=DEV).manual_seed(i)
        s = dict(mo=torch.randn(B, Cm, H, H, device=DEV, generator=g), xt=torch.randn(B, C, H, H, device=DEV, generator=g),
                 nz=torch.randn(B, C, H, H, device=DEV, generator=g), out=torch.empty(B, C, H, H, device=DEV))
        if cfg:
            s['mu'] = torch.randn(B, Cm, H, H, device=DEV, generator=g)
        for o in extra_outputs:
            s[o] = torch.empty(B, C, H, H, device=DEV)
        sets.append(s)

    def call(s):
        K.sampler_step(s['mo'], s['xt'], row, objective='pred_eps', clip=True, learned_range=learned, noise=s['nz'],
                       model_out_uncond=s.get('mu'), guidance_scale=3.0 if cfg else 1.0, sample=s['out'],
                       **{o: s[o] for o in extra_outputs})
    call(sets[0])
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for s in sets:
            call(s)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / (5 * copies) * 1e3
    gbs = by / us / 1e3
    peak, _ = _peak()
    print(f'{name:34s} B={B:4d} C={C} {H:3d}x{H:<3d} {streams} streams {by / 1e6:7.1f} MB: {us:8.1f} us = {gbs:7.0f} GB/s '
          f'({gbs / peak * 100:5.1f} % of {peak:.0f})', flush=True)
    return dict(case=name, B=B, C=C, H=H, streams=streams, algorithmic_bytes=by, us=us, gbs=gbs, frac=gbs / peak)


if __name__ == '__main__':
    if '--one' in sys.argv:   # single case for an ncu capture (101 MB tensors: > L2, quick to replay)
        run('DDIM eta=0, 101 MB tensors', 128, 3, 256)
        sys.exit(0)
    print('env', {k: v for k, v in os.environ.items() if k.startswith('B200_')}, flush=True)
    res = [
        run('DDIM eta=0, CIFAR-10 B=256', 256, 3, 32),
        run('DDIM eta=0, ADM-256 B=16', 16, 3, 256),
        run('DDIM eta=0, 403 MB tensors', 512, 3, 256),
        run('DDPM fixed_large, 403 MB tensors', 512, 3, 256, kind='ddpm'),
        run('DDPM learned_range, 403 MB', 512, 3, 256, kind='ddpm', var_type='learned_range'),
        run('DDIM CFG s=3, 403 MB tensors', 512, 3, 256, cfg=True),
        run('DDIM + pred_x0 out, 403 MB', 512, 3, 256, extra_outputs=('pred_x0',)),
        run('DDIM eta=0, 101 MB tensors', 128, 3, 256),
    ]
    peak, src = _peak()
    out = dict(kernel='sampler_step_vec4_kernel', bound='hbm', peak=peak, peak_source=src, unit='GB/s', cases=res)
    if len(sys.argv) > 2 and sys.argv[1] == '--json':
        json.dump(out, open(sys.argv[2], 'w'), indent=1)
