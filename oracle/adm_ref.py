"""ORACLE (test infrastructure, not a product path): plain-PyTorch fp32 restatement of the ADM / guided-diffusion
UNet and of the pesser (DDPM CelebA-HQ) UNet, functional over reference-format state_dicts.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Follows  models/adm/unet.py:244-275 (ResBlock._forward), :318-324 (AttentionBlock._forward), :347-412 (QKVAttention
Legacy / QKVAttention), :653-682 (UNetModel.forward), models/adm/nn.py:103-121 (timestep_embedding);
         models/pesser/model.py:6-24 (embedding), :114-134 (ResnetBlock), :161-187 (AttnBlock), :286-327 (forward).
Pinned against the live reference by oracle/gen_golden_families.py; fixtures in tests/golden/family_forward.pt.
"""
import math

import torch
import torch.nn.functional as F


def _gn(sd, key, x, eps=1e-5):
    return F.group_norm(x.float(), 32, sd[key + '.weight'], sd[key + '.bias'], eps)


def _conv(sd, key, x, **kw):
    return F.conv2d(x, sd[key + '.weight'], sd[key + '.bias'], **kw)


def _lin(sd, key, x):
    return F.linear(x, sd[key + '.weight'], sd[key + '.bias'])


# ------------------------------------------------------------------------------------------------------
# ADM
# ------------------------------------------------------------------------------------------------------
def adm_timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def adm_plan(*, model_channels, num_res_blocks, attention_resolutions, channel_mult=(1, 2, 4, 8), num_heads=1,
             num_head_channels=-1, num_heads_upsample=-1, resblock_updown=False, conv_resample=True, **_):
    """[(sequence prefix, [(kind, heads or updown flag), ...]), ...] for input / middle / output blocks."""
    if num_heads_upsample == -1:
        num_heads_upsample = num_heads

    def heads(ch, n):
        return n if num_head_channels == -1 else ch // num_head_channels

    inp, outp = [], []
    ch, ds = int(channel_mult[0] * model_channels), 1
    for level, mult in enumerate(channel_mult):
        for _ in range(num_res_blocks):
            ch = int(mult * model_channels)
            layers = [('res', None)]
            if ds in attention_resolutions:
                layers.append(('attn', heads(ch, num_heads)))
            inp.append(layers)
        if level != len(channel_mult) - 1:
            inp.append([('res', 'down')] if resblock_updown else [('downsample', conv_resample)])
            ds *= 2
    mid = [('res', None), ('attn', heads(ch, num_heads)), ('res', None)]
    for level, mult in list(enumerate(channel_mult))[::-1]:
        for i in range(num_res_blocks + 1):
            ch = int(model_channels * mult)
            layers = [('res', None)]
            if ds in attention_resolutions:
                layers.append(('attn', heads(ch, num_heads_upsample)))
            if level and i == num_res_blocks:
                layers.append(('res', 'up') if resblock_updown else ('upsample', conv_resample))
                ds //= 2
            outp.append(layers)
    return inp, mid, outp


def adm_resblock(sd, p, x, emb, updown, scale_shift):
    h = F.silu(_gn(sd, p + '.in_layers.0', x))
    if updown == 'up':
        h = F.interpolate(h, scale_factor=2, mode='nearest')
        x = F.interpolate(x, scale_factor=2, mode='nearest')
    elif updown == 'down':
        h = F.avg_pool2d(h, 2, 2)
        x = F.avg_pool2d(x, 2, 2)
    h = _conv(sd, p + '.in_layers.2', h, padding=1)
    e = _lin(sd, p + '.emb_layers.1', F.silu(emb))[:, :, None, None]
    if scale_shift:
        scale, shift = torch.chunk(e, 2, dim=1)
        h = _gn(sd, p + '.out_layers.0', h) * (1 + scale) + shift
    else:
        h = _gn(sd, p + '.out_layers.0', h + e)
    h = _conv(sd, p + '.out_layers.3', F.silu(h), padding=1)
    if (p + '.skip_connection.weight') in sd:
        w = sd[p + '.skip_connection.weight']
        x = _conv(sd, p + '.skip_connection', x, padding=w.shape[-1] // 2)
    return x + h


def adm_attention(sd, p, x, n_heads, new_order):
    B, C, H, W = x.shape
    xf = x.reshape(B, C, -1)
    qkv = F.conv1d(_gn(sd, p + '.norm', xf), sd[p + '.qkv.weight'], sd[p + '.qkv.bias'])
    T = xf.shape[-1]
    ch = C // n_heads
    s = 1 / math.sqrt(math.sqrt(ch))
    if new_order:
        q, k, v = qkv.chunk(3, dim=1)
        q, k, v = (z.reshape(B * n_heads, ch, T) for z in (q, k, v))
    else:
        q, k, v = qkv.reshape(B * n_heads, ch * 3, T).split(ch, dim=1)
    w = torch.einsum('bct,bcs->bts', q * s, k * s)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum('bts,bcs->bct', w, v).reshape(B, -1, T)
    h = F.conv1d(a, sd[p + '.proj_out.weight'], sd[p + '.proj_out.bias'])
    return (xf + h).reshape(B, C, H, W)


def adm_forward(sd, x, t, y=None, *, cfg):
    inp, mid, outp = adm_plan(**cfg)
    scale_shift = cfg.get('use_scale_shift_norm', False)
    new_order = cfg.get('use_new_attention_order', False)
    emb = _lin(sd, 'time_embed.2', F.silu(_lin(sd, 'time_embed.0', adm_timestep_embedding(t, cfg['model_channels']))))
    if cfg.get('num_classes') is not None:
        assert y is not None
        emb = emb + sd['label_emb.weight'][y]
    else:
        assert y is None

    def run(prefix, layers, h):
        for j, (kind, arg) in enumerate(layers):
            p = f'{prefix}.{j}'
            if kind == 'res':
                h = adm_resblock(sd, p, h, emb, arg, scale_shift)
            elif kind == 'attn':
                h = adm_attention(sd, p, h, arg, new_order)
            elif kind == 'downsample':
                h = _conv(sd, p + '.op', h, stride=2, padding=1) if arg else F.avg_pool2d(h, 2, 2)
            else:
                h = F.interpolate(h, scale_factor=2, mode='nearest')
                if arg:
                    h = _conv(sd, p + '.conv', h, padding=1)
        return h

    h = _conv(sd, 'input_blocks.0.0', x, padding=1)
    hs = [h]
    for i, layers in enumerate(inp, start=1):
        h = run(f'input_blocks.{i}', layers, h)
        hs.append(h)
    h = run('middle_block', mid, h)
    for i, layers in enumerate(outp):
        h = run(f'output_blocks.{i}', layers, torch.cat([h, hs.pop()], dim=1))
    return _conv(sd, 'out.2', F.silu(_gn(sd, 'out.0', h)), padding=1)


# ------------------------------------------------------------------------------------------------------
# pesser
# ------------------------------------------------------------------------------------------------------
def pesser_timestep_embedding(t, dim):
    half = dim // 2
    step = math.log(10000) / (half - 1)
    f = torch.exp(torch.arange(half, dtype=torch.float32) * -step).to(t.device)
    ang = t.float()[:, None] * f[None, :]
    return torch.cat([torch.sin(ang), torch.cos(ang)], dim=1)


def _swish(x):
    return x * torch.sigmoid(x)


def pesser_resblock(sd, p, x, temb):
    h = _conv(sd, p + '.conv1', _swish(_gn(sd, p + '.norm1', x, 1e-6)), padding=1)
    h = h + _lin(sd, p + '.temb_proj', _swish(temb))[:, :, None, None]
    h = _conv(sd, p + '.conv2', _swish(_gn(sd, p + '.norm2', h, 1e-6)), padding=1)
    if (p + '.nin_shortcut.weight') in sd:
        x = _conv(sd, p + '.nin_shortcut', x)
    elif (p + '.conv_shortcut.weight') in sd:
        x = _conv(sd, p + '.conv_shortcut', x, padding=1)
    return x + h


def pesser_attn(sd, p, x):
    B, C, H, W = x.shape
    n = _gn(sd, p + '.norm', x, 1e-6)
    q = _conv(sd, p + '.q', n).reshape(B, C, H * W).permute(0, 2, 1)
    k = _conv(sd, p + '.k', n).reshape(B, C, H * W)
    v = _conv(sd, p + '.v', n).reshape(B, C, H * W)
    w = torch.softmax(torch.bmm(q, k) * (int(C) ** (-0.5)), dim=2)
    h = torch.bmm(v, w.permute(0, 2, 1)).reshape(B, C, H, W)
    return x + _conv(sd, p + '.proj_out', h)


def pesser_forward(sd, x, t, *, cfg):
    ch, n_res, L = cfg['ch'], cfg['num_res_blocks'], len(cfg['ch_mult'])
    with_conv = cfg.get('resamp_with_conv', True)
    temb = _lin(sd, 'temb.dense.1', _swish(_lin(sd, 'temb.dense.0', pesser_timestep_embedding(t, ch))))
    hs = [_conv(sd, 'conv_in', x, padding=1)]
    for i in range(L):
        for j in range(n_res):
            h = pesser_resblock(sd, f'down.{i}.block.{j}', hs[-1], temb)
            if f'down.{i}.attn.{j}.norm.weight' in sd:
                h = pesser_attn(sd, f'down.{i}.attn.{j}', h)
            hs.append(h)
        if i != L - 1:
            if with_conv:
                hs.append(_conv(sd, f'down.{i}.downsample.conv', F.pad(hs[-1], (0, 1, 0, 1)), stride=2))
            else:
                hs.append(F.avg_pool2d(hs[-1], 2, 2))
    h = pesser_resblock(sd, 'mid.block_1', hs[-1], temb)
    h = pesser_attn(sd, 'mid.attn_1', h)
    h = pesser_resblock(sd, 'mid.block_2', h, temb)
    for i in reversed(range(L)):
        for j in range(n_res + 1):
            h = pesser_resblock(sd, f'up.{i}.block.{j}', torch.cat([h, hs.pop()], dim=1), temb)
            if f'up.{i}.attn.{j}.norm.weight' in sd:
                h = pesser_attn(sd, f'up.{i}.attn.{j}', h)
        if i != 0:
            h = F.interpolate(h, scale_factor=2.0, mode='nearest')
            if with_conv:
                h = _conv(sd, f'up.{i}.upsample.conv', h, padding=1)
    return _conv(sd, 'conv_out', _swish(_gn(sd, 'norm_out', h, 1e-6)), padding=1)


class FamilyRef(torch.nn.Module):
    """Callable wrapper: FamilyRef('adm' | 'pesser', state_dict, cfg)(x, t[, y])."""

    def __init__(self, family, state_dict, cfg):
        super().__init__()
        self.family, self.cfg = family, dict(cfg)
        self.sd = {k: v.detach().clone() for k, v in state_dict.items()}

    def to(self, device):  # noqa: D102
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        return self

    @torch.no_grad()
    def forward(self, x, t, y=None):
        if self.family == 'adm':
            return adm_forward(self.sd, x, t, y, cfg=self.cfg)
        return pesser_forward(self.sd, x, t, cfg=self.cfg)


def randomize_zero_params(state_dict, seed=2022, std=0.02):
    """ADM zero-initialises every ResBlock's second conv, attention proj_out and the final conv: a random-init
    network would output exactly 0.  For parity/benchmarks every all-zero tensor is re-drawn N(0, std)."""
    g = torch.Generator().manual_seed(seed)
    out = {}
    for k, v in state_dict.items():
        if v.dtype.is_floating_point and v.numel() > 0 and not bool(v.any()):
            out[k] = (torch.randn(v.shape, generator=g) * std).to(v.device, v.dtype)
        else:
            out[k] = v
    return out
