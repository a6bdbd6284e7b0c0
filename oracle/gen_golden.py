"""Pins the oracle against the reference itself and freezes golden vectors under tests/golden/.

Run in the build container only (it imports the unmodified reference from /root/reference, which does not
travel to the GPU box):    python oracle/gen_golden.py

For every case it (1) runs the reference, (2) runs the oracle restatement (oracle/*.py) on the same inputs and
asserts agreement (bit-exact for schedules and the sampler arithmetic, <= 1e-5 for network forwards, where op
fusion order may differ by an ulp), and (3) stores inputs + reference outputs as small .pt fixtures.
Weights are never stored: both sides build them from `torch.manual_seed(seed)` + the default initialisers, which
the product modules reproduce exactly (same registration order as the reference).
"""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, 'tests', 'golden')
REF = '/root/reference'


def import_reference():
    """Imports the reference's `models` / `diffusions` packages (needs a stub omegaconf) and returns them."""
    stub = types.ModuleType('omegaconf')
    stub.OmegaConf = type('OmegaConf', (), {})
    stub.DictConfig = dict
    sys.modules.setdefault('omegaconf', stub)
    sys.path.insert(0, REF)
    import diffusions as ref_diffusions
    import models as ref_models
    from models.unet import UNet
    from models.unet_categorial_adagn import UNetCategorialAdaGN
    return ref_models, ref_diffusions, UNet, UNetCategorialAdaGN


UNET_CFGS = {
    'tiny': dict(in_channels=3, out_channels=3, dim=32, dim_mults=[1, 2, 2, 2], use_attn=[False, True, False, False],
                 num_res_blocks=2, n_heads=1, dropout=0.1),
    'mnist': dict(in_channels=1, out_channels=1, dim=64, dim_mults=[1, 2, 2, 2], use_attn=[False, True, False, False],
                  num_res_blocks=2, n_heads=1, dropout=0.1),
    'cifar10': dict(in_channels=3, out_channels=3, dim=128, dim_mults=[1, 2, 2, 2],
                    use_attn=[False, True, False, False], num_res_blocks=2, n_heads=1, dropout=0.1),
}
ADAGN_CFGS = {
    'tiny_adagn': dict(in_channels=3, out_channels=3, dim=64, dim_mults=[1, 2, 2, 2], use_attn=[False, True, True, False],
                       num_res_blocks=2, num_classes=10, attn_head_dims=64, resblock_updown=True, dropout=0.1),
    'cfg_cifar10': dict(in_channels=3, out_channels=3, dim=128, dim_mults=[1, 2, 2, 2],
                        use_attn=[False, True, True, False], num_res_blocks=2, num_classes=10, attn_head_dims=64,
                        resblock_updown=True, dropout=0.1),
}


def main():
    sys.path.insert(0, ROOT)
    from oracle import diffusion_ref as R
    from oracle.unet_ref import UNetRef
    _, ref_diff, RefUNet, RefAdaGN = import_reference()
    os.makedirs(GOLD, exist_ok=True)
    torch.set_grad_enabled(False)

    # ------------------------------------------------------------------ schedules
    sched = {'betas_dtype': {}, 'alphas_cumprod': {}, 'respaced': {}}
    for kind in ('linear', 'quad', 'const', 'cosine'):
        for T in (1000, 200):
            rb = ref_diff.schedule.get_beta_schedule(T, kind)
            ob = R.beta_schedule(T, kind)
            assert rb.dtype == ob.dtype and torch.equal(rb, ob), (kind, T)
            d = ref_diff.ddpm.DDPM(total_steps=T, beta_schedule=kind)
            o = R.DDPMRef(total_steps=T, beta_schedule=kind)
            assert torch.equal(d.alphas_cumprod, o.alphas_cumprod)
            sched['betas_dtype'][(kind, T)] = str(rb.dtype)
            sched['alphas_cumprod'][(kind, T)] = d.alphas_cumprod.clone()
    for kind in ('uniform', 'uniform-leading', 'uniform-linspace', 'uniform-trailing', 'quad', 'none'):
        for (T, S) in ((1000, 50), (1000, 10), (1000, 250), (1000, 300), (200, 200), (200, 7)):
            rs = ref_diff.schedule.get_respaced_seq(T, kind, S)
            assert torch.equal(rs, R.respaced_seq(T, kind, S)), (kind, T, S)
            sched['respaced'][(kind, T, S)] = rs.clone()
    torch.save(sched, os.path.join(GOLD, 'schedules.pt'))
    print('schedules: oracle == reference (bit-exact); saved')

    # ------------------------------------------------------------------ sampler arithmetic
    g = torch.Generator().manual_seed(123)
    xt = torch.randn(2, 3, 4, 4, generator=g)
    noise = torch.randn(2, 3, 4, 4, generator=g)
    mo3 = torch.randn(2, 3, 4, 4, generator=g) * 1.5
    mo6 = torch.randn(2, 6, 4, 4, generator=g)
    steps = []
    kw0 = dict(total_steps=1000, respace_type='uniform', respace_steps=50)
    for kind in ('ddpm', 'ddim'):
        for var_type in (('fixed_small', 'fixed_large', 'learned_range') if kind == 'ddpm' else ('fixed_large',)):
            for eta in ((0.0, 0.5, 1.0) if kind == 'ddim' else (0.0,)):
                for objective in ('pred_eps', 'pred_x0', 'pred_v'):
                    for clip in (True, False):
                        for beta in ('linear', 'cosine'):
                            if kind == 'ddpm':
                                rd = ref_diff.ddpm.DDPM(var_type=var_type, objective=objective, clip_denoised=clip,
                                                        beta_schedule=beta, **kw0)
                                od = R.DDPMRef(var_type=var_type, objective=objective, clip_denoised=clip,
                                               beta_schedule=beta, **kw0)
                            else:
                                rd = ref_diff.ddim.DDIM(eta=eta, objective=objective, clip_denoised=clip,
                                                        beta_schedule=beta, **kw0)
                                od = R.DDIMRef(eta=eta, objective=objective, clip_denoised=clip, beta_schedule=beta,
                                               **kw0)
                            mo = mo6 if var_type == 'learned_range' else mo3
                            for (t, tp) in ((980, 960), (500, 480), (20, 0), (0, -1)):
                                orig = torch.randn_like
                                torch.randn_like = lambda x, _n=noise: _n.clone()   # inject the reference's noise draw
                                try:
                                    r = rd.denoise(mo.clone(), xt, t, tp)
                                finally:
                                    torch.randn_like = orig
                                o = od.denoise(mo.clone(), xt, t, tp, reverse_eps=noise)
                                for key in ('sample', 'mean', 'var', 'pred_x0', 'pred_eps'):
                                    assert torch.equal(r[key], o[key]), (kind, var_type, eta, objective, clip, t, key)
                                steps.append(dict(kind=kind, var_type=var_type, eta=eta, objective=objective, clip=clip,
                                                  beta=beta, t=t, t_prev=tp,
                                                  out={k: r[k].clone() for k in ('sample', 'mean', 'var', 'pred_x0',
                                                                                'pred_eps')}))
    # diffuse / get_v
    rd = ref_diff.ddpm.DDPM(total_steps=1000)
    tt = torch.tensor([0, 999])
    torch.save(dict(xt=xt, noise=noise, mo3=mo3, mo6=mo6, steps=steps, kw0=kw0,
                    diffuse=dict(t=tt, out=rd.diffuse(xt, tt, noise), v=rd.get_v(xt, noise, tt))),
               os.path.join(GOLD, 'sampler_steps.pt'))
    od = R.DDPMRef(total_steps=1000)
    assert torch.equal(rd.diffuse(xt, tt, noise), od.diffuse(xt, tt, noise))
    assert torch.equal(rd.get_v(xt, noise, tt), od.get_v(xt, noise, tt))
    print(f'sampler steps: {len(steps)} cases, oracle == reference (bit-exact); saved')

    # ------------------------------------------------------------------ UNet forwards
    fw = {}
    for name, cfg in UNET_CFGS.items():
        torch.manual_seed(2022)
        ref = RefUNet(**cfg).eval()
        B = 2
        g = torch.Generator().manual_seed(5)
        x = torch.randn(B, cfg['in_channels'], 32, 32, generator=g)
        t = torch.tensor([37, 911])
        want = ref(x, t)
        orc = UNetRef(ref.state_dict(), dim=cfg['dim'], n_heads=cfg['n_heads'])
        got = orc(x, t)
        err = (got - want).abs().max().item()
        assert err <= 1e-5, (name, err)
        fw[name] = dict(cfg=cfg, seed=2022, x=x, t=t, out=want.clone(),
                        param_sum=float(sum(p.double().sum() for p in ref.parameters())))
        print(f'unet {name}: oracle vs reference max abs err {err:.2e}')
    for name, cfg in ADAGN_CFGS.items():
        torch.manual_seed(2022)
        ref = RefAdaGN(**cfg).eval()
        B = 2
        g = torch.Generator().manual_seed(6)
        x = torch.randn(B, 3, 32, 32, generator=g)
        t = torch.tensor([37, 911])
        y = torch.tensor([3, 7])
        orc = UNetRef(ref.state_dict(), dim=cfg['dim'], adagn=True, attn_head_dims=cfg['attn_head_dims'],
                      num_res_blocks=cfg['num_res_blocks'])
        outs = {}
        for tag, yy in (('cond', y), ('uncond', None)):
            want = ref(x, t, yy)
            err = (orc(x, t, yy) - want).abs().max().item()
            assert err <= 1e-5, (name, tag, err)
            outs[tag] = want.clone()
            print(f'adagn unet {name} [{tag}]: oracle vs reference max abs err {err:.2e}')
        fw[name] = dict(cfg=cfg, seed=2022, x=x, t=t, y=y, out=outs,
                        param_sum=float(sum(p.double().sum() for p in ref.parameters())))
    torch.save(fw, os.path.join(GOLD, 'unet_forward.pt'))

    # ------------------------------------------------------------------ short sampling runs (tiny UNet)
    torch.manual_seed(2022)
    ref = RefUNet(**UNET_CFGS['tiny']).eval()
    orc = UNetRef(ref.state_dict(), dim=32, n_heads=1)
    g = torch.Generator().manual_seed(9)
    x0 = torch.randn(2, 3, 32, 32, generator=g)
    noises = [torch.randn(2, 3, 32, 32, generator=g) for _ in range(10)]
    runs = {}
    for tag, mk_ref, mk_orc in (
        ('ddim10_eta0', lambda: ref_diff.ddim.DDIM(respace_type='uniform', respace_steps=10),
         lambda: R.DDIMRef(respace_type='uniform', respace_steps=10)),
        ('ddim10_eta1', lambda: ref_diff.ddim.DDIM(respace_type='uniform', respace_steps=10, eta=1.0),
         lambda: R.DDIMRef(respace_type='uniform', respace_steps=10, eta=1.0)),
        ('ddpm10_fixed_small', lambda: ref_diff.ddpm.DDPM(respace_type='uniform', respace_steps=10, var_type='fixed_small'),
         lambda: R.DDPMRef(respace_type='uniform', respace_steps=10, var_type='fixed_small')),
    ):
        it = iter(noises)
        orig = torch.randn_like
        torch.randn_like = lambda x, _it=it: next(_it).clone()
        try:
            want = mk_ref().sample(ref, x0, tqdm_kwargs=dict(disable=True))
        finally:
            torch.randn_like = orig
        got = mk_orc().sample(orc, x0, noises=noises)
        err = (got - want).abs().max().item()
        assert err <= 1e-4, (tag, err)
        runs[tag] = want.clone()
        print(f'sampling {tag}: oracle vs reference max abs err {err:.2e}')
    # CFG (DDIM-10, s=3) on the tiny AdaGN UNet
    torch.manual_seed(2022)
    refc = RefAdaGN(**ADAGN_CFGS['tiny_adagn']).eval()
    orcc = UNetRef(refc.state_dict(), dim=64, adagn=True, attn_head_dims=64, num_res_blocks=2)
    y = torch.tensor([1, 8])
    it = iter(noises)
    orig = torch.randn_like
    torch.randn_like = lambda x, _it=it: next(_it).clone()
    try:
        want = ref_diff.ddim.DDIMCFG(guidance_scale=3.0, respace_type='uniform', respace_steps=10).sample(
            refc, x0, tqdm_kwargs=dict(disable=True), model_kwargs=dict(y=y))
    finally:
        torch.randn_like = orig
    od = R.DDIMRef(respace_type='uniform', respace_steps=10)
    got = None
    for out in od.sample_loop_cfg(orcc, x0, 3.0, dict(y=y), dict(y=None), noises=noises):
        got = out['sample']
    err = (got - want).abs().max().item()
    assert err <= 1e-4, ('cfg', err)
    print(f'sampling ddim10 cfg s=3: oracle vs reference max abs err {err:.2e}')
    runs['ddim10_cfg3'] = want.clone()
    torch.save(dict(x0=x0, noises=noises, y=y, runs=runs), os.path.join(GOLD, 'sampling_runs.pt'))
    print('golden fixtures written to', GOLD)


if __name__ == '__main__':
    main()
