"""The oracle (oracle/*.py) against the golden vectors frozen from the reference (oracle/gen_golden.py) and the
known answers recorded in SURVEY.md section 8c.  CPU only."""
import os

import pytest
import torch

from oracle import diffusion_ref as R
from oracle.unet_ref import UNetRef


@pytest.fixture(scope='module')
def gold(golden_dir):
    return {n: torch.load(os.path.join(golden_dir, n + '.pt'), weights_only=False)
            for n in ('schedules', 'sampler_steps', 'unet_forward', 'sampling_runs')}


def test_schedules_bit_exact(gold):
    s = gold['schedules']
    for (kind, T), ac in s['alphas_cumprod'].items():
        o = R.DDPMRef(total_steps=T, beta_schedule=kind)
        assert str(R.beta_schedule(T, kind).dtype) == s['betas_dtype'][(kind, T)]
        assert torch.equal(o.alphas_cumprod, ac), (kind, T)
    for (kind, T, S), seq in s['respaced'].items():
        assert torch.equal(R.respaced_seq(T, kind, S), seq), (kind, T, S)


def test_schedule_known_answers():
    """SURVEY.md section 8c constants (probed from the reference)."""
    b = R.beta_schedule(1000, 'linear')
    assert b[0].item() == 1e-4 and b[1].item() == 0.00011991991991991993 and b[999].item() == 0.02
    ac = R.DDPMRef(total_steps=1000).alphas_cumprod
    assert ac[0].item() == 0.9998999834060669
    assert ac[499].item() == 0.07858724147081375
    assert ac[999].item() == 4.0358296246267855e-05
    c = R.beta_schedule(1000, 'cosine')
    assert c.dtype == torch.float32 and c[0].item() == 4.128422369831242e-05
    assert abs(c[999].item() - 0.999) < 1e-7
    assert R.DDPMRef(total_steps=1000, beta_schedule='cosine').alphas_cumprod[499].item() == 0.4938434660434723
    assert R.DDPMRef(total_steps=1000, beta_schedule='quad').alphas_cumprod[999].item() == 0.0007334124529734254
    assert R.DDPMRef(total_steps=1000, beta_schedule='const').alphas_cumprod[0].item() == pytest.approx(0.98, abs=1e-7)
    assert R.respaced_seq(1000, 'uniform', 50).tolist() == list(range(0, 1000, 20))
    ls = R.respaced_seq(1000, 'uniform-linspace', 50).tolist()
    assert ls[:4] == [0, 20, 40, 61] and ls[-3:] == [958, 978, 999]
    assert R.respaced_seq(1000, 'uniform-trailing', 50).tolist() == list(range(19, 1000, 20))
    q = R.respaced_seq(1000, 'quad', 50).tolist()
    assert q[:4] == [0, 0, 1, 2] and q[-3:] == [736, 767, 800]
    assert R.respaced_seq(1000, 'uniform', 10).tolist() == list(range(0, 1000, 100))
    assert R.respaced_seq(1000, 'quad', 10).tolist()[:4] == [0, 9, 39, 88]
    assert len(R.respaced_seq(1000, 'uniform', 250)) == 250
    assert len(R.respaced_seq(1000, 'uniform', 300)) == 334      # S does not divide T
    with pytest.raises(ValueError):
        R.beta_schedule(10, 'nope')
    with pytest.raises(ValueError):
        R.respaced_seq(10, 'nope', 2)


def test_ddim_ddpm_known_answers():
    """DDIM/DDPM single-step constants of SURVEY.md section 8c: x_t = 0.5, eps = [0.25, -3.0]."""
    xt = torch.full((1, 2, 1, 1), 0.5)
    eps = torch.tensor([0.25, -3.0]).view(1, 2, 1, 1)
    d = R.DDIMRef(total_steps=1000, respace_type='uniform', respace_steps=50)
    o = d.denoise(eps.clone(), xt, 980, 960, reverse_eps=torch.zeros_like(xt))
    assert o['pred_x0'].flatten().tolist() == [1.0, 1.0]
    assert o['pred_eps'].flatten()[0].item() == pytest.approx(0.4923309087753296, abs=1e-7)
    assert o['sample'].flatten()[0].item() == pytest.approx(0.5016588568687439, abs=1e-7)
    assert float(o['var']) == 0.0
    o = d.denoise(eps.clone(), xt, 0, -1, reverse_eps=torch.zeros_like(xt))
    assert o['sample'].flatten().tolist() == pytest.approx([0.49752476811408997, 0.530027449131012], abs=1e-7)
    for vt, var in (('fixed_large', 0.3246144652366638), ('fixed_small', 0.324605256319046)):
        p = R.DDPMRef(total_steps=1000, respace_type='uniform', respace_steps=50, var_type=vt)
        o = p.denoise(eps.clone(), xt, 980, 960, reverse_eps=torch.zeros_like(xt))
        assert o['mean'].flatten()[0].item() == pytest.approx(0.4139326810836792, abs=1e-7)
        assert float(o['var']) == pytest.approx(var, abs=1e-7)


def test_sampler_steps_bit_exact(gold):
    g = gold['sampler_steps']
    xt, noise = g['xt'], g['noise']
    for c in g['steps']:
        kw = dict(objective=c['objective'], clip_denoised=c['clip'], beta_schedule=c['beta'], **g['kw0'])
        d = R.DDPMRef(var_type=c['var_type'], **kw) if c['kind'] == 'ddpm' else R.DDIMRef(eta=c['eta'], **kw)
        mo = g['mo6'] if c['var_type'] == 'learned_range' else g['mo3']
        o = d.denoise(mo.clone(), xt, c['t'], c['t_prev'], reverse_eps=noise)
        for k, v in c['out'].items():
            assert torch.equal(o[k], v), (c['kind'], c['var_type'], c['eta'], c['objective'], c['clip'], c['t'], k)
    d = R.DDPMRef(total_steps=1000)
    assert torch.equal(d.diffuse(xt, g['diffuse']['t'], noise), g['diffuse']['out'])
    assert torch.equal(d.get_v(xt, noise, g['diffuse']['t']), g['diffuse']['v'])


def _product_state_dict(name, cfg, seed):
    import models
    torch.manual_seed(seed)
    cls = models.UNetCategorialAdaGN if 'num_classes' in cfg else models.UNet
    return cls(**cfg).state_dict()


@pytest.mark.parametrize('name', ['tiny', 'mnist', 'cifar10', 'tiny_adagn', 'cfg_cifar10'])
def test_unet_forward_matches_reference(gold, name):
    """Oracle forward over the PRODUCT module's seeded state_dict == the reference's output with the reference's
    seeded init: pins both the oracle arithmetic and the product's parameter registration order / initialisers."""
    c = gold['unet_forward'][name]
    sd = _product_state_dict(name, c['cfg'], c['seed'])
    assert float(sum(v.double().sum() for v in sd.values())) == pytest.approx(c['param_sum'], rel=1e-12)
    with torch.no_grad():
        if 'num_classes' in c['cfg']:
            orc = UNetRef(sd, dim=c['cfg']['dim'], adagn=True, attn_head_dims=c['cfg']['attn_head_dims'],
                          num_res_blocks=c['cfg']['num_res_blocks'])
            assert torch.allclose(orc(c['x'], c['t'], c['y']), c['out']['cond'], rtol=0, atol=1e-5)
            assert torch.allclose(orc(c['x'], c['t'], None), c['out']['uncond'], rtol=0, atol=1e-5)
        else:
            orc = UNetRef(sd, dim=c['cfg']['dim'], n_heads=c['cfg']['n_heads'])
            assert torch.allclose(orc(c['x'], c['t']), c['out'], rtol=0, atol=1e-5)


def test_sampling_runs_match_reference(gold):
    g = gold['sampling_runs']
    cfg = gold['unet_forward']['tiny']['cfg']
    orc = UNetRef(_product_state_dict('tiny', cfg, 2022), dim=32, n_heads=1)
    mk = {
        'ddim10_eta0': lambda: R.DDIMRef(respace_type='uniform', respace_steps=10),
        'ddim10_eta1': lambda: R.DDIMRef(respace_type='uniform', respace_steps=10, eta=1.0),
        'ddpm10_fixed_small': lambda: R.DDPMRef(respace_type='uniform', respace_steps=10, var_type='fixed_small'),
    }
    with torch.no_grad():
        for tag, f in mk.items():
            got = f().sample(orc, g['x0'], noises=g['noises'])
            assert torch.allclose(got, g['runs'][tag], rtol=0, atol=1e-4), tag
        ccfg = gold['unet_forward']['tiny_adagn']['cfg']
        orcc = UNetRef(_product_state_dict('tiny_adagn', ccfg, 2022), dim=64, adagn=True, attn_head_dims=64,
                       num_res_blocks=2)
        out = None
        for out in R.DDIMRef(respace_type='uniform', respace_steps=10).sample_loop_cfg(
                orcc, g['x0'], 3.0, dict(y=g['y']), dict(y=None), noises=g['noises']):
            pass
        assert torch.allclose(out['sample'], g['runs']['ddim10_cfg3'], rtol=0, atol=1e-4)


# ------------------------------------------------------------------------------------------------------
# ADM / pesser families (oracle/adm_ref.py, fixtures from oracle/gen_golden_families.py)
# ------------------------------------------------------------------------------------------------------
@pytest.fixture(scope='module')
def fam(golden_dir):
    return torch.load(os.path.join(golden_dir, 'family_forward.pt'), weights_only=False)


def _family_product(c):
    from models.adm.unet import UNetModel
    from models.adm.unet_combined import UNetCombined
    from models.pesser.model import Model
    from oracle.adm_ref import randomize_zero_params
    torch.manual_seed(c['seed'])
    if c['family'] == 'pesser':
        return Model(**c['cfg'])
    m = (UNetCombined if c['family'] == 'adm_combined' else UNetModel)(**c['cfg'])
    m.load_state_dict(randomize_zero_params(m.state_dict()))
    return m


@pytest.mark.parametrize('name', ['adm_tiny_cond', 'adm_tiny_plain', 'adm_combined_cond', 'adm_combined_uncond',
                                  'pesser_tiny'])
def test_family_forward_matches_reference(fam, name):
    """The PRODUCT module has the reference's state_dict keys, shapes, registration order and seeded initial values,
    and the oracle evaluated on that state_dict reproduces the reference's output."""
    from oracle.adm_ref import FamilyRef
    c = fam[name]
    m = _family_product(c)
    sd = m.state_dict()
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == c['keys']
    if 'param_sum' in c:
        assert float(sum(p.double().sum() for p in m.parameters())) == pytest.approx(c['param_sum'], rel=1e-12)
    if c['family'] == 'adm_combined':
        sub = 'unet_uncond.' if c['y'] is None else 'unet_cond.'
        sd = {k[len(sub):]: v for k, v in sd.items() if k.startswith(sub)}
        orc = FamilyRef('adm', sd, dict(c['cfg'], num_classes=None if c['y'] is None else c['cfg']['num_classes']))
    else:
        orc = FamilyRef(c['family'], sd, c['cfg'])
    assert torch.allclose(orc(c['x'], c['t'], c['y']), c['out'], rtol=0, atol=1e-5)


def test_adm_conditioning_contract():
    """models/adm/unet.py:662-664: y must be given iff the model is class-conditional (AssertionError)."""
    from models.adm.unet import UNetModel
    m = UNetModel(image_size=32, in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
                  attention_resolutions=[2], channel_mult=[1, 2], num_classes=10, num_head_channels=64)
    x, t = torch.zeros(1, 3, 32, 32), torch.zeros(1, dtype=torch.long)
    with pytest.raises(AssertionError):
        m(x, t)
    m2 = UNetModel(image_size=32, in_channels=3, model_channels=64, out_channels=3, num_res_blocks=1,
                   attention_resolutions=[2], channel_mult=[1, 2], num_head_channels=64)
    with pytest.raises(AssertionError):
        m2(x, t, torch.zeros(1, dtype=torch.long))
    with pytest.raises(RuntimeError):      # CPU tensors: no fallback
        m2.eval()(x, t)


# ------------------------------------------------------------------------------------------------------
# Euler / Heun samplers (fixtures from oracle/gen_golden_samplers.py)
# ------------------------------------------------------------------------------------------------------
def test_ode_samplers_match_reference(golden_dir, gold):
    g = torch.load(os.path.join(golden_dir, 'ode_samplers.pt'), weights_only=False)
    for c in g['steps']:
        d = R.EulerRef(objective=c['objective'], clip_denoised=c['clip'], beta_schedule=c['beta'], **g['kw0'])
        o = d.denoise(g['mo'].clone(), g['xt'], c['t'], c['t_prev'])
        assert torch.equal(o['sample'], c['euler']['sample']) and torch.equal(o['pred_x0'], c['euler']['pred_x0'])
    cfg = gold['unet_forward']['tiny']['cfg']
    orc = UNetRef(_product_state_dict('tiny', cfg, 2022), dim=32, n_heads=1)
    with torch.no_grad():
        for tag, cls in (('euler10', R.EulerRef), ('heun10', R.HeunRef)):
            got = cls(respace_type='uniform', respace_steps=10).sample(orc, g['x0'])
            assert torch.allclose(got, g['runs'][tag], rtol=0, atol=1e-4), tag


def test_ddim_inversion_matches_reference(golden_dir, gold):
    """DDIMRef.denoise_inversion / sample_inversion(_cfg) against the fixtures frozen from the reference's
    diffusions/ddim.py:88-132, 202-242 by oracle/gen_golden_inversion.py: 48 single steps bit-exact (3 objectives x clip x
    2 beta schedules x 4 (t, t_next) pairs incl. a target past the end of the schedule), two 10-step inversion runs."""
    g = torch.load(os.path.join(golden_dir, 'ddim_inversion.pt'), weights_only=False)
    xt, mo = g['xt'], g['mo']
    assert len(g['steps']) == 48
    for c in g['steps']:
        d = R.DDIMRef(total_steps=1000, beta_schedule=c['beta'], objective=c['objective'], clip_denoised=c['clip'],
                      respace_type='uniform', respace_steps=50)
        o = d.denoise_inversion(mo.clone(), xt, c['t'], c['t_next'])
        for k, v in c['out'].items():
            assert torch.equal(o[k], v), (c['objective'], c['clip'], c['beta'], c['t'], c['t_next'], k)
    with pytest.raises(ValueError):
        R.DDIMRef(eta=0.5).denoise_inversion(mo, xt, 0, 20)
    with torch.no_grad():
        r = g['runs']['uncond10']
        cfg = gold['unet_forward']['tiny']['cfg']
        orc = UNetRef(_product_state_dict('tiny', cfg, 2022), dim=32, n_heads=1)
        got = R.DDIMRef(**r['kw']).sample_inversion(orc, r['x0'])
        assert torch.allclose(got, r['latent'], rtol=0, atol=1e-4)
        rc = g['runs']['cfg10']
        ccfg = gold['unet_forward']['tiny_adagn']['cfg']
        orcc = UNetRef(_product_state_dict('tiny_adagn', ccfg, 2022), dim=64, adagn=True, attn_head_dims=64,
                       num_res_blocks=2)
        out = None
        for out in R.DDIMRef(**rc['kw']).sample_inversion_loop_cfg(orcc, rc['x0'], rc['guidance_scale'], dict(y=rc['y']),
                                                                  dict(y=None)):
            pass
        assert torch.allclose(out['sample'], rc['latent'], rtol=0, atol=1e-4)
