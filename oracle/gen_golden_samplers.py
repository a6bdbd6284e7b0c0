"""Pins the Euler / Heun restatements (oracle/diffusion_ref.py: EulerRef, HeunRef) against the live reference
(diffusions/euler.py, diffusions/heun.py) and freezes fixtures in tests/golden/ode_samplers.pt.
Build-container only (imports /root/reference):    python oracle/gen_golden_samplers.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_golden import GOLD, UNET_CFGS, import_reference  # noqa: E402


def main():
    from oracle import diffusion_ref as R
    from oracle.unet_ref import UNetRef
    _, ref_diff, RefUNet, _ = import_reference()
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(321)
    xt = torch.randn(2, 3, 4, 4, generator=g)
    mo = torch.randn(2, 3, 4, 4, generator=g) * 1.5
    d1 = torch.randn(2, 3, 4, 4, generator=g)
    steps = []
    kw0 = dict(total_steps=1000, respace_type='uniform', respace_steps=20)
    for objective in ('pred_eps', 'pred_x0', 'pred_v'):
        for clip in (True, False):
            for beta in ('linear', 'cosine'):
                rd = ref_diff.euler.EulerSampler(objective=objective, clip_denoised=clip, beta_schedule=beta, **kw0)
                od = R.EulerRef(objective=objective, clip_denoised=clip, beta_schedule=beta, **kw0)
                rh = ref_diff.heun.HeunSampler(objective=objective, clip_denoised=clip, beta_schedule=beta, **kw0)
                for (t, tp) in ((950, 900), (500, 450), (50, 0), (0, -1)):
                    r = rd.denoise(mo.clone(), xt, t, tp)
                    o = od.denoise(mo.clone(), xt, t, tp)
                    assert torch.equal(r['sample'], o['sample']) and torch.equal(r['pred_x0'], o['pred_x0']), (objective, t)
                    rec = dict(objective=objective, clip=clip, beta=beta, t=t, t_prev=tp,
                               euler={k: r[k].clone() for k in ('sample', 'pred_x0')})
                    if tp >= 0:
                        rh.denoise_1st_order(mo.clone(), xt, t, tp)
                        rh._1st_order_derivative = d1.clone()     # injected so the fixture is self-contained
                        rh._1st_order_xt = xt.clone()
                        h2 = rh.denoise_2nd_order(mo.clone(), xt * 0.9, t, tp)
                        rec['heun2'] = {k: h2[k].clone() for k in ('sample', 'pred_x0')}
                    steps.append(rec)
    # short sampling runs on the tiny UNet
    torch.manual_seed(2022)
    ref = RefUNet(**UNET_CFGS['tiny']).eval()
    orc = UNetRef(ref.state_dict(), dim=32, n_heads=1)
    x0 = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(9))
    runs = {}
    for tag, RefCls, OrcCls in (('euler10', ref_diff.euler.EulerSampler, R.EulerRef),
                                ('heun10', ref_diff.heun.HeunSampler, R.HeunRef)):
        want = RefCls(respace_type='uniform', respace_steps=10).sample(ref, x0, tqdm_kwargs=dict(disable=True))
        got = OrcCls(respace_type='uniform', respace_steps=10).sample(orc, x0)
        err = (got - want).abs().max().item()
        assert err <= 1e-4, (tag, err)
        print(f'sampling {tag}: oracle vs reference max abs err {err:.2e}')
        runs[tag] = want.clone()
    torch.save(dict(xt=xt, mo=mo, d1=d1, kw0=kw0, steps=steps, x0=x0, runs=runs), os.path.join(GOLD, 'ode_samplers.pt'))
    print(f'{len(steps)} single-step cases bit-exact; written {os.path.join(GOLD, "ode_samplers.pt")}')


if __name__ == '__main__':
    main()
