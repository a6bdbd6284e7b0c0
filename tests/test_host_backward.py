"""Host-side logic of the backward pass on CPU: data-gradient weight layouts and tap tables (through the CPU emulation
of the conv kernel's addressing, tests/emulate.py) against autograd, and the row mapping of ADM's fused qkv parameter."""
import pytest
import torch
import torch.nn as nn
import torch.nn.functional as F

import b200diff as K
from tests.emulate import conv_emulate


def _dgrad_ref(conv, x, dy):
    x = x.clone().requires_grad_(True)
    if conv.padding == (0, 0) and conv.stride == (2, 2):          # pesser: pad (0,1,0,1) then stride 2
        y = conv(F.pad(x, (0, 1, 0, 1)))
    else:
        y = conv(x)
    y.backward(dy)
    return x.grad


def test_dgrad_3x3_is_the_forward_kernel_on_flipped_weights():
    """models/backward.py:_w_dgrad / b200_pack_weights mode 1: dX = conv(dY, W flipped spatially, channels swapped)."""
    torch.manual_seed(0)
    conv = nn.Conv2d(6, 5, 3, padding=1)
    x, dy = torch.randn(2, 6, 8, 8), torch.randn(2, 5, 8, 8)
    want = _dgrad_ref(conv, x, dy)
    wd = K.pack_weight(conv.weight.detach().flip(2, 3).transpose(0, 1)).float()        # [Cin, 9*Cout]
    # the same layout written index by index as the pack kernel does (mode 1): dst[ci][(8 - tap)*Co + co]
    Co, Ci = 5, 6
    w = conv.weight.detach()
    manual = torch.zeros(Ci, 9 * Co)
    for co in range(Co):
        for ci in range(Ci):
            for tap in range(9):
                manual[ci, (8 - tap) * Co + co] = w[co, ci, tap // 3, tap % 3]
    assert torch.equal(manual.to(torch.bfloat16).float(), wd)
    got = conv_emulate(dy.permute(0, 2, 3, 1)[:, None].contiguous(), manual, Ci, 2, 8, 8, K.taps_3x3_s1())
    assert torch.allclose(got.permute(0, 3, 1, 2), want, atol=1e-5)


@pytest.mark.parametrize('pad_lo', [1, 0])
def test_stride2_dgrad_plan(pad_lo):
    """models/backward.py:_s2_dgrad_plan: the adjoint of the 3x3 stride-2 conv (symmetric pad 1, or pesser's (0,1,0,1)
    padding) as four 2x2-tap phase convolutions over dY."""
    from models.backward import _s2_dgrad_plan
    torch.manual_seed(1)
    conv = nn.Conv2d(4, 3, 3, stride=2, padding=1 if pad_lo == 1 else 0)
    x, dy = torch.randn(2, 4, 8, 8), torch.randn(2, 3, 4, 4)
    want = _dgrad_ref(conv, x, dy)
    taps, wd = _s2_dgrad_plan(conv, pad_lo)
    assert len(taps) == 4 and all(len(t) == 4 for t in taps) and tuple(wd.shape) == (4 * 4, 4 * 3)
    # reference in bf16-rounded weights
    convr = nn.Conv2d(4, 3, 3, stride=2, padding=1 if pad_lo == 1 else 0)
    convr.weight.data = conv.weight.detach().to(torch.bfloat16).float()
    convr.bias.data = conv.bias.detach()
    want = _dgrad_ref(convr, x, dy)
    got = conv_emulate(dy.permute(0, 2, 3, 1)[:, None].contiguous(), wd.float(), 4, 2, 4, 4, taps, w_rows_per_phase=4)
    assert torch.allclose(got.permute(0, 3, 1, 2), want, atol=1e-5)


@pytest.mark.parametrize('new_order', [False, True])
def test_adm_qkv_rows_and_gradient_targets(new_order):
    """models/adm/unet.py:AttentionBlock.packed_weights and models/backward.py:_attn_targets: the head-interleaved rows
    of the fused qkv Conv1d (reference models/adm/unet.py:366 legacy / :398 new order) map onto [q heads | k heads | v]."""
    from models.adm.unet import AttentionBlock
    torch.manual_seed(2)
    C, H = 128, 2
    d = C // H
    blk = AttentionBlock(C, num_heads=H, use_new_attention_order=new_order)
    x = torch.randn(3, C, 10)
    qkv = F.conv1d(x, blk.qkv.weight, blk.qkv.bias)
    if new_order:
        q, k, v = qkv.chunk(3, dim=1)
    else:
        q, k, v = qkv.reshape(3 * H, 3 * d, 10).split(d, dim=1)
        q, k, v = (z.reshape(3, C, 10) for z in (q, k, v))
    wqk, bqk, wv, bv, _, _ = blk.packed_weights()
    got_qk = torch.einsum('oc,bct->bot', wqk.float(), x) + bqk[None, :, None]
    got_v = torch.einsum('oc,bct->bot', wv.float(), x) + bv[None, :, None]
    assert torch.allclose(got_qk[:, :C], q, atol=2e-2) and torch.allclose(got_qk[:, C:], k, atol=2e-2)
    assert torch.allclose(got_v, v, atol=2e-2)

    class _G:                                  # gradient views like models.backward._Grads
        def __init__(self):
            self.t = {id(p): torch.zeros_like(p) for p in blk.parameters()}

        def __call__(self, p):
            return self.t[id(p)]
    from models.backward import _attn_targets
    g = _G()
    targets, proj = _attn_targets(None, dict(tag='a', mods=blk), g, C)
    assert proj is blk.proj_out
    marks = {'q': 1.0, 'k': 2.0, 'v': 3.0}
    for name, val in marks.items():
        rows = 0
        for (r0, n, wg, bg) in targets[name]:
            wg += val + r0 / 1000.0            # writes through the views
            bg += val
            rows += n
        assert rows == C
    wfull = g(blk.qkv.weight).view(3 * C, C)
    for j, name in enumerate(('q', 'k', 'v')):
        for h in range(H):
            r = (j * H + h) * d if new_order else (h * 3 + j) * d
            assert torch.allclose(wfull[r:r + d], torch.full((d, C), marks[name] + (h * d if not new_order else 0) / 1000.0))
    assert float(g(blk.qkv.bias).sum()) == pytest.approx((1 + 2 + 3) * C)
