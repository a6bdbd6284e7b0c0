"""ORACLE (test infrastructure, not a product path): plain-PyTorch fp32 restatement of the reference UNets.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
Functional style over a reference-format state_dict (same key names as the reference checkpoints), following
  models/modules.py:40-57    sinusoidal embedding            models/modules.py:77-102   self-attention block
  models/unet.py:10-43       ResBlock                        models/unet.py:121-152     UNet.forward
  models/modules.py:105-123  AdaGN                           models/unet_categorial_adagn.py:12-62,165-208
Runs on whatever device the tensors live on (CPU for the reported cpu_baseline; CUDA eager for fast parity).

Pinned against the reference imported live (oracle/gen_golden.py) and the frozen fixtures in tests/golden/.
`round_hook` lets the tests emulate where the CUDA path rounds to bf16 (design studies; not used for parity).
"""
import math

import torch
import torch.nn.functional as F


def sinusoidal(t, dim):
    half = dim // 2
    step = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, device=t.device) * -step)
    ang = t[:, None] * freqs[None, :]
    return torch.cat((ang.sin(), ang.cos()), dim=-1)


def _gn(sd, key, x, eps=1e-5):
    return F.group_norm(x, 32, sd[key + '.weight'], sd[key + '.bias'], eps)


def _conv(sd, key, x, **kw):
    return F.conv2d(x, sd[key + '.weight'], sd[key + '.bias'], **kw)


def time_mlp(sd, t, dim):
    e = sinusoidal(t, dim)
    e = F.linear(e, sd['time_embed.1.weight'], sd['time_embed.1.bias'])
    return F.linear(F.silu(e), sd['time_embed.3.weight'], sd['time_embed.3.bias'])


def attention_block(sd, p, x, n_heads):
    B, C, H, W = x.shape
    n = _gn(sd, p + '.norm', x)
    q = _conv(sd, p + '.q', n).view(B * n_heads, -1, H * W)
    k = _conv(sd, p + '.k', n).view(B * n_heads, -1, H * W)
    v = _conv(sd, p + '.v', n).view(B * n_heads, -1, H * W)
    att = torch.bmm((q * (C // n_heads) ** -0.5).transpose(1, 2), k).softmax(dim=-1)
    o = torch.bmm(v, att.transpose(1, 2)).view(B, -1, H, W)
    return _conv(sd, p + '.proj', o) + x


def _drop(h, drop, p):
    """Training-mode nn.Dropout with an injected keep mask: drop = {block prefix: (mask NCHW, probability)}."""
    if drop is None or p not in drop:
        return h
    mask, prob = drop[p]
    return h * mask / (1.0 - prob)


def resblock(sd, p, x, emb, drop=None):
    """models/unet.py:30-43 (dropout is the identity in eval mode; in training mode the mask is injected)."""
    h = _conv(sd, p + '.blk1.2', F.silu(_gn(sd, p + '.blk1.0', x)), padding=1)
    h = h + F.linear(F.silu(emb), sd[p + '.proj.1.weight'], sd[p + '.proj.1.bias'])[:, :, None, None]
    h = _conv(sd, p + '.blk2.3', _drop(F.silu(_gn(sd, p + '.blk2.0', h)), drop, p), padding=1)
    sc = _conv(sd, p + '.shortcut', x) if (p + '.shortcut.weight') in sd else x
    return h + sc


def resblock_adagn(sd, p, x, emb, updown=None, drop=None):
    """models/unet_categorial_adagn.py:44-62."""
    h = F.silu(_gn(sd, p + '.blk1.0', x))
    if updown == 'up':
        h = F.interpolate(h, scale_factor=2, mode='nearest')
        x = F.interpolate(x, scale_factor=2, mode='nearest')
    elif updown == 'down':
        h = F.avg_pool2d(h, 2, 2)
        x = F.avg_pool2d(x, 2, 2)
    h = _conv(sd, p + '.blk1.2', h, padding=1)
    ss = F.linear(F.silu(emb), sd[p + '.adagn.proj.1.weight'], sd[p + '.adagn.proj.1.bias'])
    ys, yb = torch.chunk(ss, 2, dim=-1)
    h = _gn(sd, p + '.adagn.gn', h) * (1 + ys[:, :, None, None]) + yb[:, :, None, None]
    h = _conv(sd, p + '.blk2.2', _drop(F.silu(h), drop, p), padding=1)
    sc = _conv(sd, p + '.shortcut', x) if (p + '.shortcut.weight') in sd else x
    return h + sc


def _kind(sd, p):
    if (p + '.blk1.0.weight') in sd:
        return 'res'
    if (p + '.norm.weight') in sd:
        return 'attn'
    if (p + '.1.weight') in sd:
        return 'up'       # nn.Sequential(Upsample, Conv2d)
    if (p + '.weight') in sd:
        return 'down'     # bare strided Conv2d
    return None


def unet_forward(sd, x, t, *, dim, n_heads=1, y=None, adagn=False, attn_head_dims=64, num_res_blocks=2,
                 trace=None, drop=None):
    """Forward of models.unet.UNet (adagn=False) or UNetCategorialAdaGN (adagn=True) from a state_dict.

    The block structure is recovered from the key names; `num_res_blocks` is only needed to tell the
    up/down ResBlocks of the AdaGN UNet (last block of a stage) from the plain ones."""
    emb = time_mlp(sd, t, dim)
    if adagn and y is not None and 'class_embed.weight' in sd:
        emb = emb + sd['class_embed.weight'][y]

    def heads(p):
        C = sd[p + '.norm.weight'].shape[0]
        if not adagn:
            return n_heads
        return C // attn_head_dims

    def rb(p, h, updown=None):
        return resblock_adagn(sd, p, h, emb, updown, drop) if adagn else resblock(sd, p, h, emb, drop)

    h = _conv(sd, 'first_conv', x, padding=1)
    skips = [h]
    s = 0
    while _kind(sd, f'down_blocks.{s}.0'):
        j, n_res = 0, 0
        while True:
            p = f'down_blocks.{s}.{j}'
            kind = _kind(sd, p)
            if kind is None:
                break
            if kind == 'res':
                n_res += 1
                h = rb(p, h, 'down' if (adagn and n_res > num_res_blocks) else None)
                skips.append(h)
            elif kind == 'attn':
                h = attention_block(sd, p, h, heads(p))
                skips[-1] = h
            else:
                h = _conv(sd, p, h, stride=2, padding=1)
                skips.append(h)
            if trace is not None:
                trace.append((p, h))
            j += 1
        s += 1
    h = rb('bottleneck_block.0', h)
    h = attention_block(sd, 'bottleneck_block.1', h, 1)
    h = rb('bottleneck_block.2', h)
    if trace is not None:
        trace.append(('bottleneck_block.2', h))
    s = 0
    while _kind(sd, f'up_blocks.{s}.0'):
        j, n_res = 0, 0
        while True:
            p = f'up_blocks.{s}.{j}'
            kind = _kind(sd, p)
            if kind is None:
                break
            if kind == 'res':
                n_res += 1
                if adagn and n_res > num_res_blocks + 1:
                    h = rb(p, h, 'up')
                else:
                    h = rb(p, torch.cat((h, skips.pop()), dim=1))
            elif kind == 'attn':
                h = attention_block(sd, p, h, heads(p))
            else:
                h = _conv(sd, p + '.1', F.interpolate(h, scale_factor=2, mode='nearest'), padding=1)
            if trace is not None:
                trace.append((p, h))
            j += 1
        s += 1
    return _conv(sd, 'last_conv.2', F.silu(_gn(sd, 'last_conv.0', h)), padding=1)


class UNetRef(torch.nn.Module):
    """Callable wrapper so the oracle can be driven like the reference: model(x, t, **kw)."""

    def __init__(self, state_dict, **cfg):
        super().__init__()
        self.sd = {k: v.detach().clone() for k, v in state_dict.items()}
        self.cfg = cfg

    def to(self, device):  # noqa: D102
        self.sd = {k: v.to(device) for k, v in self.sd.items()}
        return self

    @torch.no_grad()
    def forward(self, x, t, y=None):
        return unet_forward(self.sd, x, t, y=y, **self.cfg)
