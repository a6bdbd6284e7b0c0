"""Pins the DDIM-inversion restatement (oracle/diffusion_ref.py: DDIMRef.denoise_inversion, sample_inversion_loop,
sample_inversion_loop_cfg) against the live reference (diffusions/ddim.py:88-132, 202-242) and freezes fixtures in
tests/golden/ddim_inversion.pt.  Build-container only (imports /root/reference):  python oracle/gen_golden_inversion.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.gen_golden import ADAGN_CFGS, GOLD, UNET_CFGS, import_reference  # noqa: E402


def main():
    from oracle import diffusion_ref as R
    from oracle.unet_ref import UNetRef
    _, ref_diff, RefUNet, RefAdaGN = import_reference()
    torch.set_grad_enabled(False)
    g = torch.Generator().manual_seed(777)
    xt = torch.randn(2, 3, 4, 4, generator=g)
    mo = torch.randn(2, 3, 4, 4, generator=g) * 1.5
    mo_u = torch.randn(2, 3, 4, 4, generator=g) * 1.5
    steps = []
    for objective in ('pred_eps', 'pred_x0', 'pred_v'):
        for clip in (True, False):
            for beta in ('linear', 'cosine'):
                kw = dict(total_steps=1000, beta_schedule=beta, objective=objective, clip_denoised=clip,
                          respace_type='uniform', respace_steps=50)
                rd = ref_diff.ddim.DDIM(**kw)
                od = R.DDIMRef(**kw)
                # (t, t_next): first step, mid-schedule, last respaced pair, and a target past the schedule end (ac_next = 0)
                for (t, tn) in ((0, 20), (480, 500), (960, 980), (980, 1000)):
                    r = rd.denoise_inversion(mo.clone(), xt, t, tn)
                    o = od.denoise_inversion(mo.clone(), xt, t, tn)
                    for k in ('sample', 'pred_x0', 'pred_eps'):
                        assert torch.equal(r[k], o[k]), (objective, clip, beta, t, tn, k)
                    steps.append(dict(objective=objective, clip=clip, beta=beta, t=t, t_next=tn,
                                      out={k: r[k].clone() for k in ('sample', 'pred_x0', 'pred_eps')}))
    # eta != 0 must raise in both
    for cls, kw in ((ref_diff.ddim.DDIM, {}), (R.DDIMRef, {})):
        try:
            cls(eta=0.5).denoise_inversion(mo, xt, 0, 20)
            raise AssertionError('eta != 0 accepted')
        except ValueError:
            pass

    # ---- short inversion runs on the tiny UNet: x0 -> x_T (10 respaced steps), then DDIM back ----
    torch.manual_seed(2022)
    ref = RefUNet(**UNET_CFGS['tiny']).eval()
    orc = UNetRef(ref.state_dict(), dim=32, n_heads=1)
    x0 = torch.randn(2, 3, 32, 32, generator=torch.Generator().manual_seed(11)).clamp(-1, 1)
    kw = dict(total_steps=1000, respace_type='uniform', respace_steps=10)
    want = ref_diff.ddim.DDIM(**kw).sample_inversion(ref, x0, tqdm_kwargs=dict(disable=True))
    got = R.DDIMRef(**kw).sample_inversion(orc, x0)
    err = (got - want).abs().max().item()
    assert err <= 1e-4, err
    print(f'inversion run (tiny UNet, 10 steps): oracle vs reference max abs err {err:.2e}')
    runs = {'uncond10': dict(x0=x0, latent=want.clone(), kw=kw)}

    # ---- inversion under classifier-free guidance on the tiny AdaGN UNet ----
    torch.manual_seed(2022)
    refc = RefAdaGN(**ADAGN_CFGS['tiny_adagn']).eval()
    orcc = UNetRef(refc.state_dict(), dim=64, adagn=True, attn_head_dims=64, num_res_blocks=2)
    y = torch.tensor([3, 7])
    kwc = dict(total_steps=1000, beta_schedule='cosine', respace_type='uniform', respace_steps=10)
    rc = ref_diff.ddim.DDIMCFG(guidance_scale=2.0, **kwc)
    wantc = rc.sample_inversion(refc, x0, uncond_conditioning=None, tqdm_kwargs=dict(disable=True), model_kwargs=dict(y=y))
    oc = R.DDIMRef(**kwc)
    gotc = None
    for out in oc.sample_inversion_loop_cfg(orcc, x0, 2.0, dict(y=y), dict(y=None)):
        gotc = out['sample']
    errc = (gotc - wantc).abs().max().item()
    assert errc <= 1e-4, errc
    print(f'CFG inversion run (tiny AdaGN UNet, s=2, 10 steps): oracle vs reference max abs err {errc:.2e}')
    runs['cfg10'] = dict(x0=x0, y=y, latent=wantc.clone(), kw=kwc, guidance_scale=2.0)

    torch.save(dict(xt=xt, mo=mo, mo_u=mo_u, steps=steps, runs=runs), os.path.join(GOLD, 'ddim_inversion.pt'))
    print(f'{len(steps)} single-step inversion cases bit-exact; written {os.path.join(GOLD, "ddim_inversion.pt")}')


if __name__ == '__main__':
    main()
