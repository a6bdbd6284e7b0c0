"""Host-side logic of the product path, on CPU: schedules (bit-exact), per-step scalar coefficients, tap tables and
weight packing (through a CPU emulation of the conv kernel's addressing), state_dict compatibility, errors."""
import os

import pytest
import torch
import torch.nn.functional as F

import b200diff as K
import diffusions
import models
from diffusions.schedule import get_beta_schedule, get_respaced_seq
from oracle import diffusion_ref as R
from tests.emulate import conv_emulate


@pytest.fixture(scope='module')
def gold(golden_dir):
    return {n: torch.load(os.path.join(golden_dir, n + '.pt'), weights_only=False)
            for n in ('schedules', 'unet_forward')}


def test_product_schedules_bit_exact(gold):
    s = gold['schedules']
    for (kind, T), ac in s['alphas_cumprod'].items():
        assert str(get_beta_schedule(T, kind).dtype) == s['betas_dtype'][(kind, T)]
        d = diffusions.DDPM(total_steps=T, beta_schedule=kind)
        assert torch.equal(d.alphas_cumprod, ac), (kind, T)
        assert d.alphas_cumprod.dtype == torch.float32
    for (kind, T, S), seq in s['respaced'].items():
        assert torch.equal(get_respaced_seq(T, kind, S), seq), (kind, T, S)
        assert get_respaced_seq(T, kind, S).dtype == torch.int64
    assert torch.equal(get_respaced_seq(1000, None, 5), torch.arange(1000))


def test_error_behaviour_matches_reference():
    with pytest.raises(ValueError, match='Invalid objective'):
        diffusions.DDPM(objective='pred_nothing')
    with pytest.raises(ValueError, match='Invalid var_type'):
        diffusions.DDPM(var_type='huge')
    with pytest.raises(ValueError, match='not supported'):
        get_beta_schedule(10, 'sigmoid')
    with pytest.raises(ValueError, match='not supported'):
        get_respaced_seq(10, 'random', 2)
    with pytest.raises(AssertionError):
        diffusions.DDPM(total_steps=10, betas=torch.zeros(11))
    d = diffusions.DDIMCFG(guidance_scale=3.0, respace_type='uniform', respace_steps=10)
    with pytest.raises(ValueError, match='Condition argument'):
        next(d.sample_loop(None, torch.zeros(1, 3, 4, 4), model_kwargs=dict(z=1)))
    assert d.guidance_scale == 3.0 and d.cond_kwarg == 'y' and d.eta == 0.
    dd = diffusions.DDIM(var_type='learned_range', eta=0.3)     # var_type passes through **kwargs like the reference
    assert dd.var_type == 'learned_range' and dd.eta == 0.3
    dd.set_respaced_seq('uniform', 20)
    assert len(dd.respaced_seq) == 20
    with dd_hack(d):
        assert d.objective == 'pred_eps'


class dd_hack:
    def __init__(self, d):
        self.cm = d.hack_objective('pred_eps')

    def __enter__(self):
        return self.cm.__enter__()

    def __exit__(self, *a):
        return self.cm.__exit__(*a)


@pytest.mark.parametrize('kind', ['ddpm', 'ddim'])
def test_step_coefficients_match_oracle_scalars(kind):
    """The host-evaluated coefficient rows reproduce the reference's zero-dim tensor arithmetic exactly: applying
    them in the kernel's op order on CPU gives the oracle's outputs bit for bit."""
    g = torch.Generator().manual_seed(0)
    xt, mo, noise = (torch.randn(2, 3, 4, 4, generator=g) for _ in range(3))
    for var_type in (('fixed_small', 'fixed_large') if kind == 'ddpm' else ('fixed_large',)):
        for eta in ((0.0, 0.7) if kind == 'ddim' else (0.0,)):
            kw = dict(total_steps=1000, respace_type='uniform', respace_steps=50)
            if kind == 'ddpm':
                ours, ref = diffusions.DDPM(var_type=var_type, **kw), R.DDPMRef(var_type=var_type, **kw)
            else:
                ours, ref = diffusions.DDIM(eta=eta, **kw), R.DDIMRef(eta=eta, **kw)
            for (t, tp) in ((980, 960), (20, 0), (0, -1)):
                c = ours._predict_coefs(t) + ours._step_coefs(t, tp)
                c = [torch.as_tensor(v, dtype=torch.float32) for v in c]
                x0 = (c[0] * xt - c[1] * mo).clamp(-1, 1)
                eps = (c[0] * xt - x0) / c[1]
                mean = (c[4] * x0 + c[5] * xt) + c[6] * eps
                sample = mean if t == 0 else mean + torch.sqrt(c[7]) * noise
                o = ref.denoise(mo.clone(), xt, t, tp, reverse_eps=noise)
                assert torch.equal(x0, o['pred_x0']) and torch.equal(eps, o['pred_eps'])
                assert torch.equal(mean, o['mean']) and torch.equal(sample, o['sample'])
                assert float(c[7]) == float(o['var'])


def test_tap_tables_and_weight_packing():
    torch.manual_seed(0)
    B, C, Co, H = 2, 8, 5, 8
    x, w = torch.randn(B, C, H, H), torch.randn(Co, C, 3, 3)
    x1, w1 = torch.randn(B, 6, H, H), torch.randn(Co, 6, 1, 1)
    nhwc = lambda t: t.permute(0, 2, 3, 1).contiguous()  # noqa: E731
    f32pack = lambda *a: K.pack_weight(*a).float()  # noqa: E731  (bf16-rounded weights)
    wq = K.pack_weight(w).float().reshape(Co, 3, 3, C).permute(0, 3, 1, 2)
    w1q = K.pack_weight(w1).float().reshape(Co, 1, 1, 6).permute(0, 3, 1, 2)
    tol = dict(rtol=1e-4, atol=1e-4)
    o = conv_emulate(nhwc(x)[:, None], f32pack(w), Co, B, H, H, K.taps_3x3_s1())
    assert torch.allclose(o.permute(0, 3, 1, 2), F.conv2d(x, wq, padding=1), **tol)
    o = conv_emulate(nhwc(x)[:, None], f32pack(w, w1), Co, B, H, H, K.taps_3x3_s1(), a1=nhwc(x1)[:, None])
    assert torch.allclose(o.permute(0, 3, 1, 2), F.conv2d(x, wq, padding=1) + F.conv2d(x1, w1q), **tol)
    planes = torch.stack([nhwc(x)[:, a::2, b::2] for a in range(2) for b in range(2)], 1)
    for pad_lo, pad in ((1, (1, 1, 1, 1)), (0, (0, 1, 0, 1))):
        o = conv_emulate(planes, f32pack(w), Co, B, H // 2, H // 2, K.taps_3x3_s2(pad_lo))
        assert torch.allclose(o.permute(0, 3, 1, 2), F.conv2d(F.pad(x, pad), wq, stride=2), **tol)
    # nearest-2x + 3x3 == four 2x2 phase convolutions with summed taps (exact in fp32 before the bf16 rounding)
    wp = K.pack_weight_up2(w).float()
    o = conv_emulate(nhwc(x)[:, None], wp, Co, B, H, H, K.taps_up2_3x3(), w_rows_per_phase=Co)
    ref = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, padding=1)
    assert (o.permute(0, 3, 1, 2) - ref).abs().max() <= 0.02 * ref.abs().max()     # bf16 weight rounding only


def test_state_dict_compat_and_ema():
    torch.manual_seed(2022)
    m = models.UNet()
    sd = m.state_dict()
    assert len(sd) == 328 and sum(v.numel() for v in sd.values()) == 35746307
    assert sd['down_blocks.1.0.shortcut.weight'].shape == (256, 128, 1, 1)
    assert sd['up_blocks.0.3.1.weight'].shape == (256, 256, 3, 3)         # Upsample = Sequential(.., conv)
    assert sd['down_blocks.0.2.weight'].shape == (128, 128, 3, 3)          # Downsample = bare strided conv
    assert 'time_embed.1.weight' in sd and 'last_conv.2.bias' in sd
    m2 = models.UNet()
    m2.load_state_dict(sd)
    c = models.UNetCategorialAdaGN(num_classes=10)
    assert len(c.state_dict()) == 427 and sum(p.numel() for p in c.parameters()) == 44178947
    # EMA known answers of the reference (models/ema.py:82-117): decay 0.9, weights 0 -> 1 -> 2 => 0.1 then 0.29
    p = torch.nn.Parameter(torch.zeros(3))
    ema = models.EMA([p], decay=0.9, gradual=False)
    p.data.fill_(1.0)
    ema.update([p])
    assert torch.allclose(ema.shadow[0], torch.full((3,), 0.1))
    p.data.fill_(2.0)
    ema.update([p])
    assert torch.allclose(ema.shadow[0], torch.full((3,), 0.29))
    ema.apply_shadow([p])
    assert torch.allclose(p.data, torch.full((3,), 0.29))
    ema.restore([p])
    assert torch.allclose(p.data, torch.full((3,), 2.0))


class _PlanEngine:
    """Records the op sequence a UNet forward hands to the engine (no kernels): checks the consumer look-ahead."""
    attn_block = True
    pingpong = False

    def __init__(self):
        self.calls = []

    def begin_forward(self):
        pass

    def check_input(self, X, T, in_channels):
        return X

    def embed(self, *a, **k):
        return None, 0

    def _act(self, C, H, skipc=None):
        from models.engine import Act
        return Act(None, 2, H, H, C)

    def first_conv(self, tag, conv, X, next_gn=None):
        self.calls.append(('first', tag, None, next_gn, {}))
        return self._act(conv.out_channels, X.shape[2])

    def resblock(self, tag, blk, x, skip, *a, next_gn=None):
        self.calls.append(('res', tag, skip, next_gn, dict(x=x)))
        conv1 = blk.blk1[2]
        return self._act(conv1.out_channels, x.H)

    resblock_adagn = resblock

    def attention(self, tag, blk, x, next_gn=None):
        self.calls.append(('attn', tag, None, next_gn, dict(x=x)))
        return self._act(x.C, x.H)

    def downsample_conv(self, tag, blk, x, **k):
        self.calls.append(('down', tag, None, None, dict(x=x)))
        return self._act(x.C, x.H // 2)

    def upsample_conv(self, tag, conv, x, bf16_out=False):
        self.calls.append(('up', tag, None, None, dict(x=x, bf16_out=bf16_out)))
        return self._act(conv.out_channels, x.H * 2)

    def head(self, tag, h, norm, conv, out):
        self.calls.append(('head', tag, None, None, dict(x=h, norm=norm)))
        return 'out'


@pytest.mark.parametrize('family', ['unet', 'adagn'])
def test_forward_plan_lookahead(family):
    """models/unet.py / unet_categorial_adagn.py flatten the reference's forward (unet.py:127-150) into an op list in which
    every block knows its consumer.  The flags derived from that look-ahead decide what the kernels do NOT write:
      * next_gn[3] (`dead`): the producer's fp32 output is never written -- legal only when the consumer is the output head
        or a decoder block that concatenates a skip connection, and the producer's output is no skip connection itself;
      * next_gn[2] (skip channels of the consumer's concat) must equal the channels of the skip the consumer then pops;
      * bf16_out of an up-sampling conv: legal only when the next op is such a concatenating block."""
    import torch.nn as nn
    if family == 'unet':
        m = models.UNet(dim=32, dim_mults=[1, 2, 2], use_attn=[False, True, False], num_res_blocks=2, n_heads=1)
    else:
        m = models.UNetCategorialAdaGN(dim=32, dim_mults=[1, 2, 2], use_attn=[False, True, False], num_res_blocks=2,
                                       num_classes=10, attn_head_dims=32)
    eng = _PlanEngine()
    m.__dict__['_engine'] = eng
    X = torch.zeros(2, 3, 16, 16)
    T = torch.zeros(2, dtype=torch.long)
    out = m._forward_impl(X, T) if family == 'unet' else m._forward_impl(X, T, None)
    assert out == 'out'
    calls = eng.calls
    assert calls[0][0] == 'first' and calls[-1][0] == 'head'
    n_dead = n_cat = n_bf16 = 0
    for i, (kind, tag, skip, next_gn, kw) in enumerate(calls[:-1]):
        nkind, ntag, nskip, _, nkw = calls[i + 1]
        if kind == 'up':
            if kw['bf16_out']:
                n_bf16 += 1
                assert nkind == 'res' and nskip is not None, f'{tag}: bf16 output but the consumer {ntag} does not concatenate'
        if next_gn is None:
            continue
        norm, silu = next_gn[0], next_gn[1]
        skip_c = next_gn[2] if len(next_gn) > 2 else 0
        dead = bool(next_gn[3]) if len(next_gn) > 3 else False
        assert isinstance(norm, nn.GroupNorm)
        if skip_c:
            n_cat += 1
            assert nkind == 'res' and nskip is not None and nskip.C == skip_c, f'{tag}: concat look-ahead does not match {ntag}'
        if dead:
            n_dead += 1
            assert tag.startswith(('up_blocks', 'bottleneck_block')), f'{tag}: an encoder output (skip connection) marked dead'
            assert nkind == 'head' or (nkind == 'res' and nskip is not None), f'{tag}: dead output but {ntag} reads it as fp32'
        if nkind == 'head':
            assert norm is nkw['norm'] and silu
    assert n_dead >= 3 and n_cat >= 3 and (n_bf16 >= 1 or family == 'adagn')
