"""Per-kernel parity cases for libb200diff (run on a B200).

Each case calls ONE C-ABI entry point through the ctypes binding and compares with a plain PyTorch fp32
restatement of the same op on the same (bf16-rounded where the kernel rounds) inputs.  Tolerances are written
next to each case.  Run as `python tests/kernel_cases.py <case> [...]` (one process per case keeps a faulting
kernel from poisoning the CUDA context of the others); `tests/test_kernels_gpu.py` drives it under pytest.
"""
import json
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))

import b200diff as K  # noqa: E402

DEV = 'cuda'


def _report(name, got, ref, rtol, atol):
    got = got.float()
    ref = ref.float()
    diff = (got - ref).abs()
    denom = ref.norm().item() + 1e-30
    rel_l2 = (got - ref).norm().item() / denom
    bad = diff > (atol + rtol * ref.abs())
    nbad = int(bad.sum().item())
    info = {
        'case': name, 'shape': list(got.shape), 'max_abs': diff.max().item(), 'rel_l2': rel_l2,
        'ref_absmax': ref.abs().max().item(), 'mismatch': nbad, 'numel': got.numel(),
        'finite': bool(torch.isfinite(got).all().item()),
    }
    if nbad:
        idx = torch.nonzero(bad)
        info['first_bad'] = idx[0].tolist()
        info['last_bad'] = idx[-1].tolist()
        flat = diff.flatten().argmax().item()
        info['argmax'] = list(torch.unravel_index(torch.tensor(flat), got.shape))
        info['argmax'] = [int(v) for v in info['argmax']]
        info['got_at_max'] = got.flatten()[flat].item()
        info['ref_at_max'] = ref.flatten()[flat].item()
        # which slices along each dim are affected
        for d in range(got.dim()):
            other = [i for i in range(got.dim()) if i != d]
            per = bad.sum(dim=other)
            info[f'bad_along_dim{d}'] = [int(i) for i in torch.nonzero(per).flatten()[:16].tolist()]
    info['ok'] = (nbad == 0) and info['finite']
    print(json.dumps(info))
    return info['ok']


def _bf16r(x):
    return x.to(torch.bfloat16).float()


def _nhwc_bf16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)


def _gen(*shape, seed=0, scale=1.0):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).to(DEV)


# ----------------------------------------------------------------------------------------------------
# conv cases
# ----------------------------------------------------------------------------------------------------
def _conv_case(name, B, Cin, Cout, H, W, ksize=3, stride=1, bias=True, rowadd=False, residual=False,
               out_mode=K.OUT_F32_NHWC, sc_cin=0, pad_lo=1):
    x = _bf16r(_gen(B, Cin, H, W, seed=1))
    w = _bf16r(_gen(Cout, Cin, ksize, ksize, seed=2, scale=1.0 / math.sqrt(Cin * ksize * ksize)))
    b = _gen(Cout, seed=3) if bias else None
    Ho, Wo = H // stride, W // stride
    ra = _gen(B, Cout + 64, seed=4) if rowadd else None  # wider row to exercise the leading dimension
    res_nchw = _gen(B, Cout, Ho, Wo, seed=5) if residual else None
    x1 = w1 = None
    if sc_cin:
        x1 = _bf16r(_gen(B, sc_cin, Ho, Wo, seed=6))
        w1 = _bf16r(_gen(Cout, sc_cin, 1, 1, seed=7, scale=1.0 / math.sqrt(sc_cin)))
    # ---- reference ----
    if stride == 1:
        ref = F.conv2d(x, w, None, stride=1, padding=ksize // 2)
    else:
        xp = F.pad(x, (1, 1, 1, 1)) if pad_lo == 1 else F.pad(x, (0, 1, 0, 1))
        ref = F.conv2d(xp, w, None, stride=2, padding=0)
    if sc_cin:
        ref = ref + F.conv2d(x1, w1)
    if bias:
        ref = ref + b[None, :, None, None]
    if rowadd:
        ref = ref + ra[:, :Cout, None, None]
    if residual:
        ref = ref + res_nchw
    # ---- kernel ----
    if stride == 1:
        a0 = _nhwc_bf16(x)
        geom = (Cin, H, W, 1)
        taps = K.taps_3x3_s1() if ksize == 3 else K.taps_1x1()
    else:
        a0 = torch.empty(B, 4, H // 2, W // 2, Cin, device=DEV, dtype=torch.bfloat16)
        K.cast_bf16(x.permute(0, 2, 3, 1).contiguous(), a0, B, H, W, Cin, parity_split=True)
        geom = (Cin, H // 2, W // 2, 4)
        taps = K.taps_3x3_s2(pad_lo)
    wp = K.pack_weight(w, w1)
    if out_mode in (K.OUT_F32_NHWC, K.OUT_BF16_NHWC):
        out = torch.full((B, Ho, Wo, Cout), float('nan'), device=DEV,
                         dtype=torch.float32 if out_mode == K.OUT_F32_NHWC else torch.bfloat16)
    else:
        out = torch.full((B, Cout, Ho, Wo), float('nan'), device=DEV,
                         dtype=torch.float32 if out_mode == K.OUT_F32_NCHW else torch.bfloat16)
    res = res_nchw.permute(0, 2, 3, 1).contiguous() if residual else None
    stats = K.new_stats(B, Cout, DEV) if (out_mode == K.OUT_F32_NHWC and Cout > 32) else None
    K.conv2d(a0, wp, Cout, B, Ho, Wo, taps, a0_geom=geom,
             a1=_nhwc_bf16(x1) if sc_cin else None, a1_geom=(sc_cin, Ho, Wo, 1) if sc_cin else None,
             bias=b, rowadd=ra, rowadd_ld=(Cout + 64) if rowadd else 0, residual=res, res_ld=Cout, out=out,
             out_mode=out_mode, stats=stats)
    torch.cuda.synchronize()
    got = out.permute(0, 3, 1, 2) if out_mode in (K.OUT_F32_NHWC, K.OUT_BF16_NHWC) else out
    if stats is not None:
        # fused GroupNorm statistics: per-(image, channel) sum and sum of squares (int64 fixed point; per-thread fp32
        # partial sums: 1e-4 relative)
        ref_st = torch.stack([ref.sum(dim=(2, 3)), (ref * ref).sum(dim=(2, 3))], dim=-1)
        if not _report(name + ' [fused GN stats]', K.stats_to_float(stats).float(), ref_st, rtol=2e-4, atol=2e-2):
            return False
        # integer atomics: a second launch must reproduce the statistics bit for bit
        stats2 = K.new_stats(B, Cout, DEV)
        K.conv2d(a0, wp, Cout, B, Ho, Wo, taps, a0_geom=geom,
                 a1=_nhwc_bf16(x1) if sc_cin else None, a1_geom=(sc_cin, Ho, Wo, 1) if sc_cin else None,
                 bias=b, rowadd=ra, rowadd_ld=(Cout + 64) if rowadd else 0, residual=res, res_ld=Cout, out=out,
                 out_mode=out_mode, stats=stats2)
        torch.cuda.synchronize()
        if not torch.equal(stats, stats2):
            print(json.dumps(dict(case=name + ' [fused GN stats reproducible]', ok=False)))
            return False
    # fp32 accumulation order differs from cuDNN/cuBLAS: 1e-4 relative on O(1) values; bf16 outputs: 1 ulp = 2^-8
    if out_mode in (K.OUT_BF16_NHWC, K.OUT_BF16_NCHW):
        return _report(name, got, ref, rtol=1e-2, atol=1e-2)
    return _report(name, got, ref, rtol=2e-4, atol=2e-4)


def case_conv_basic():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return _conv_case('conv3x3 128->128 @32 plain', 2, 128, 128, 32, 32, bias=False)


def case_conv_epilogue():
    torch.backends.cudnn.allow_tf32 = False
    return _conv_case('conv3x3 128->128 @32 bias+temb+res', 4, 128, 128, 32, 32, bias=True, rowadd=True,
                      residual=True)


def case_conv_n256():
    torch.backends.cudnn.allow_tf32 = False
    ok = _conv_case('conv3x3 256->256 @16', 3, 256, 256, 16, 16, bias=True, residual=True)
    ok &= _conv_case('conv3x3 128->256 @16', 2, 128, 256, 16, 16, bias=True, rowadd=True)
    return ok


def case_conv_small_hw():
    torch.backends.cudnn.allow_tf32 = False
    ok = _conv_case('conv3x3 256->256 @8 B=3 (2 images per tile, ragged)', 3, 256, 256, 8, 8, bias=True)
    ok &= _conv_case('conv3x3 512->256 @4 B=9 (8 images per tile, ragged)', 9, 512, 256, 4, 4, bias=True,
                     residual=True)
    ok &= _conv_case('conv3x3 64->64 @32 (MNIST width)', 2, 64, 64, 32, 32, bias=True)
    return ok


def case_conv_1x1():
    torch.backends.cudnn.allow_tf32 = False
    ok = _conv_case('conv1x1 256->512 @16 bf16 NHWC (q,k)', 2, 256, 512, 16, 16, ksize=1, bias=True,
                    out_mode=K.OUT_BF16_NHWC)
    ok &= _conv_case('conv1x1 256->256 @16 bf16 NCHW (v^T)', 2, 256, 256, 16, 16, ksize=1, bias=True,
                     out_mode=K.OUT_BF16_NCHW)
    ok &= _conv_case('conv1x1 256->256 @16 +res (attn proj)', 2, 256, 256, 16, 16, ksize=1, bias=True, residual=True)
    return ok


def case_conv_shortcut():
    torch.backends.cudnn.allow_tf32 = False
    return _conv_case('conv3x3 256->256 @16 + fused 1x1 shortcut 384->256', 2, 256, 256, 16, 16, bias=True,
                      sc_cin=384)


def case_conv_stride2():
    torch.backends.cudnn.allow_tf32 = False
    ok = _conv_case('conv3x3 s2 p1 128->128 32->16', 2, 128, 128, 32, 32, stride=2, bias=True)
    ok &= _conv_case('conv3x3 s2 p1 256->256 8->4', 9, 256, 256, 8, 8, stride=2, bias=True)
    ok &= _conv_case('conv3x3 s2 pad(0,1,0,1) 128->128 32->16 (pesser)', 2, 128, 128, 32, 32, stride=2, bias=True,
                     pad_lo=0)
    return ok


def case_conv_lastconv():
    torch.backends.cudnn.allow_tf32 = False
    ok = _conv_case('conv3x3 128->3 @32 f32 NCHW (last conv)', 3, 128, 3, 32, 32, bias=True,
                    out_mode=K.OUT_F32_NCHW)
    ok &= _conv_case('conv3x3 64->1 @32 f32 NCHW (MNIST last conv)', 2, 64, 1, 32, 32, bias=True,
                     out_mode=K.OUT_F32_NCHW)
    # vertical-tap tiles: 32x4 on a wide image (two tiles per row, top / bottom / left / right zero fill), 16x8, 64x32
    # non-square; 8x8 falls back to the plain tiles (two images per tile)
    ok &= _conv_case('conv3x3 256->6 @64 f32 NCHW (ADM-style last conv, 32x4 tiles)', 2, 256, 6, 64, 64, bias=True,
                     out_mode=K.OUT_F32_NCHW)
    ok &= _conv_case('conv3x3 128->3 @16 f32 NCHW (16x8 tiles)', 3, 128, 3, 16, 16, bias=True, out_mode=K.OUT_F32_NCHW)
    ok &= _conv_case('conv3x3 128->3 @32x64 f32 NCHW (non-square)', 2, 128, 3, 32, 64, bias=True,
                     out_mode=K.OUT_F32_NCHW)
    ok &= _conv_case('conv3x3 128->3 @8 f32 NCHW (plain tiles)', 4, 128, 3, 8, 8, bias=True, out_mode=K.OUT_F32_NCHW)
    return ok


def case_conv_up2():
    """nearest-2x + conv3x3 (models/modules.py:60-67) as the 4-phase 2x2 decomposition."""
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    # (16, 256, 256, 16): enough tiles for the CTA-pair kernel; H = 8 / 16: row-linear lean epilogue; H = 4: generic epilogue
    for (B, C, Co, H) in [(2, 256, 256, 16), (3, 256, 256, 4), (16, 256, 256, 16), (5, 256, 256, 8), (4, 128, 128, 8)]:
        x = _bf16r(_gen(B, C, H, H, seed=1))
        w = _gen(Co, C, 3, 3, seed=2, scale=1.0 / math.sqrt(C * 9))
        b = _gen(Co, seed=3)
        wp = K.pack_weight_up2(w)  # [4*Co][4*C] bf16 (sums taken in fp32, rounded once)
        # reference with the SAME effective (summed then rounded) weights: rebuild phase kernels in fp32
        ref = torch.empty(B, Co, 2 * H, 2 * H, device=DEV)
        wpf = wp.float().reshape(4, Co, 4, C)
        xp = F.pad(x, (1, 1, 1, 1))
        for a in range(2):
            for bb in range(2):
                ph = a * 2 + bb
                k2 = wpf[ph].reshape(Co, 2, 2, C).permute(0, 3, 1, 2).contiguous()  # [Co, C, 2, 2]
                # tap (i, j) reads low-res offset (i - 1 + a, j - 1 + bb): crop the padded input accordingly
                sub = xp[:, :, a:a + H + 1, bb:bb + H + 1]
                ref[:, :, a::2, bb::2] = F.conv2d(sub, k2) + b[None, :, None, None]
        # and the reference semantics proper: nearest-2x then 3x3 with the unsummed bf16-rounded weights is only
        # approximately equal (different rounding); check it loosely as well
        ref_sem = F.conv2d(F.interpolate(x, scale_factor=2, mode='nearest'), w, b, padding=1)
        out = torch.full((B, 2 * H, 2 * H, Co), float('nan'), device=DEV)
        st = K.new_stats(B, Co, DEV)
        K.conv2d(_nhwc_bf16(x), wp, Co, B, H, H, K.taps_up2_3x3(), a0_geom=(C, H, H, 1), bias=b, out=out,
                 out_mode=K.OUT_F32_NHWC, w_rows_per_phase=Co, stats=st)
        torch.cuda.synchronize()
        got = out.permute(0, 3, 1, 2)
        want_st = torch.stack([out.sum(dim=(1, 2)), (out * out).sum(dim=(1, 2))], dim=-1)
        ok &= _report(f'up2+conv3x3 {C}->{Co} @{H}->{2 * H} B={B} statistics', K.stats_to_float(st), want_st, 1e-4, 1e-2)
        ok &= _report(f'up2+conv3x3 {C}->{Co} @{H}->{2 * H} vs phase-kernel ref', got, ref, 2e-4, 2e-4)
        ok &= _report(f'up2+conv3x3 {C}->{Co} @{H}->{2 * H} vs nearest+conv fp32-weight ref', got, ref_sem, 2e-2, 2e-2)
    return ok


def case_conv_tproj():
    """Per-ResBlock time-embedding projections as one 1x1 'conv' over a [B,1,1,E] bf16 operand."""
    torch.backends.cuda.matmul.allow_tf32 = False
    B, E, N = 37, 512, 1280
    x = _bf16r(_gen(B, E, seed=1))
    w = _bf16r(_gen(N, E, seed=2, scale=1.0 / math.sqrt(E)))
    b = _gen(N, seed=3)
    ref = x @ w.t() + b
    out = torch.full((B, N), float('nan'), device=DEV)
    K.conv2d(x.to(torch.bfloat16).contiguous(), w.to(torch.bfloat16).contiguous(), N, B, 1, 1, K.taps_1x1(),
             a0_geom=(E, 1, 1, 1), bias=b, out=out, out_mode=K.OUT_F32_NHWC)
    torch.cuda.synchronize()
    return _report('time-proj linear 512->1280 B=37', out, ref, 2e-4, 2e-4)


# ----------------------------------------------------------------------------------------------------
# GroupNorm
# ----------------------------------------------------------------------------------------------------
def _gn_case(name, B, C0, C1, H, W, silu=True, adagn=False, resample=0, raw=False, eps=1e-5):
    C = C0 + C1
    x0 = _gen(B, H, W, C0, seed=1) * 2.0 + 0.5
    x1 = (_gen(B, H, W, C1, seed=2) * 0.5 - 1.0) if C1 else None
    gamma = _gen(C, seed=3) * 0.2 + 1.0
    beta = _gen(C, seed=4) * 0.2
    ys = _gen(B, 2 * C + 8, seed=5) * 0.3 if adagn else None
    xcat = torch.cat([x0, x1], dim=-1) if C1 else x0
    xn = xcat.permute(0, 3, 1, 2)
    ref = F.group_norm(xn, 32, gamma, beta, eps)
    if adagn:
        ref = ref * (1 + ys[:, :C, None, None]) + ys[:, C:2 * C, None, None]
    if silu:
        ref = F.silu(ref)
    if resample == 1:
        ref = F.avg_pool2d(ref, 2, 2)
    elif resample == 2:
        ref = F.interpolate(ref, scale_factor=2, mode='nearest')
    Ho, Wo = ref.shape[2], ref.shape[3]
    out = torch.full((B, Ho, Wo, C), float('nan'), device=DEV, dtype=torch.bfloat16)
    rawo = torch.full((B, H, W, C), float('nan'), device=DEV, dtype=torch.bfloat16) if raw else None
    K.groupnorm_silu(x0, C0, x1, C1, B, H * W, W, 32, gamma, beta, eps, out,
                     scale=ys if adagn else None, shift=ys[:, C:] if adagn else None, ss_ld=(2 * C + 8) if adagn else 0,
                     silu=silu, resample=resample, raw_out=rawo)
    torch.cuda.synchronize()
    # output is rounded to bf16 once: half-ulp = 2^-9 relative
    ok = _report(name, out.permute(0, 3, 1, 2), ref, rtol=5e-3, atol=5e-3)
    if raw:
        ok &= _report(name + ' [raw copy]', rawo, xcat.to(torch.bfloat16), 0, 0)
    # streaming variant fed with producer-side statistics
    if C0 % 4 == 0 and C1 % 4 == 0 and not (raw and resample == 1):
        def st(x):
            return K.stats_from_float(torch.stack([x.sum(dim=(1, 2)), (x * x).sum(dim=(1, 2))], dim=-1).contiguous())
        out2 = torch.full((B, Ho, Wo, C), float('nan'), device=DEV, dtype=torch.bfloat16)
        raw2 = torch.full((B, H, W, C), float('nan'), device=DEV, dtype=torch.bfloat16) if raw else None
        K.groupnorm_apply(x0, C0, st(x0), x1, C1, st(x1) if C1 else None, B, H * W, W, 32, gamma, beta, eps, out2,
                          scale=ys if adagn else None, shift=ys[:, C:] if adagn else None,
                          ss_ld=(2 * C + 8) if adagn else 0, silu=silu, resample=resample, raw_out=raw2)
        torch.cuda.synchronize()
        ok &= _report(name + ' [streaming, producer stats]', out2.permute(0, 3, 1, 2), ref, rtol=5e-3, atol=5e-3)
        if raw:
            ok &= _report(name + ' [streaming raw copy]', raw2, xcat.to(torch.bfloat16), 0, 0)
        if C1 and raw and resample == 0 and C0 % 8 == 0 and C1 % 8 == 0:
            # mixed sources: the first one stored as bf16 by its producer (an up-sampling conv whose only consumer is this
            # GroupNorm; statistics from its fp32 values), the second an fp32 skip connection
            xb = x0.to(torch.bfloat16)
            cpg = C // 32
            xg = xcat.permute(0, 3, 1, 2).reshape(B, 32, cpg * H * W)
            mean = xg.mean(dim=2)[:, :, None, None, None]
            var = xg.var(dim=2, unbiased=False)[:, :, None, None, None]
            xmix = torch.cat([xb.float(), x1], dim=-1)
            xr = xmix.permute(0, 3, 1, 2).reshape(B, 32, cpg, H, W)
            refm = ((xr - mean) * torch.rsqrt(var + eps)).reshape(B, C, H, W) * gamma[None, :, None, None] + \
                beta[None, :, None, None]
            if adagn:
                refm = refm * (1 + ys[:, :C, None, None]) + ys[:, C:2 * C, None, None]
            if silu:
                refm = F.silu(refm)
            out4 = torch.full((B, H, W, C), float('nan'), device=DEV, dtype=torch.bfloat16)
            raw4 = torch.full((B, H, W, C), float('nan'), device=DEV, dtype=torch.bfloat16)
            K.groupnorm_apply(xb, C0, st(x0), x1, C1, st(x1), B, H * W, W, 32, gamma, beta, eps, out4,
                              scale=ys if adagn else None, shift=ys[:, C:] if adagn else None,
                              ss_ld=(2 * C + 8) if adagn else 0, silu=silu, raw_out=raw4)
            torch.cuda.synchronize()
            ok &= _report(name + ' [streaming, bf16 first source + fp32 second]', out4.permute(0, 3, 1, 2), refm, rtol=5e-3, atol=5e-3)
            ok &= _report(name + ' [streaming, mixed sources: raw copy]', raw4, xmix.to(torch.bfloat16), 0, 0)
        if C1 == 0 and not raw:
            # bf16-stored input whose statistics were taken from the fp32 values by its producer
            xb = x0.to(torch.bfloat16)
            cpg = C // 32
            xg = x0.permute(0, 3, 1, 2).reshape(B, 32, cpg * H * W)
            mean = xg.mean(dim=2)[:, :, None, None, None]
            var = xg.var(dim=2, unbiased=False)[:, :, None, None, None]
            xr = xb.float().permute(0, 3, 1, 2).reshape(B, 32, cpg, H, W)
            refb = ((xr - mean) * torch.rsqrt(var + eps)).reshape(B, C, H, W) * gamma[None, :, None, None] + \
                beta[None, :, None, None]
            if adagn:
                refb = refb * (1 + ys[:, :C, None, None]) + ys[:, C:2 * C, None, None]
            if silu:
                refb = F.silu(refb)
            if resample == 1:
                refb = F.avg_pool2d(refb, 2, 2)
            elif resample == 2:
                refb = F.interpolate(refb, scale_factor=2, mode='nearest')
            out3 = torch.full((B, Ho, Wo, C), float('nan'), device=DEV, dtype=torch.bfloat16)
            K.groupnorm_apply(xb, C0, st(x0), None, 0, None, B, H * W, W, 32, gamma, beta, eps, out3,
                              scale=ys if adagn else None, shift=ys[:, C:] if adagn else None,
                              ss_ld=(2 * C + 8) if adagn else 0, silu=silu, resample=resample)
            torch.cuda.synchronize()
            ok &= _report(name + ' [streaming, bf16 input]', out3.permute(0, 3, 1, 2), refb, rtol=5e-3, atol=5e-3)
    return ok


def case_groupnorm():
    ok = _gn_case('GN+SiLU C128 @32', 3, 128, 0, 32, 32)
    ok &= _gn_case('GN+SiLU C256 @16', 3, 256, 0, 16, 16)
    ok &= _gn_case('GN+SiLU cat(256,128)=384 @16 (group straddles sources) + raw', 2, 256, 128, 16, 16, raw=True)
    ok &= _gn_case('GN+SiLU cat(256,256)=512 @4 + raw', 5, 256, 256, 4, 4, raw=True)
    ok &= _gn_case('GN+SiLU cat(256,256)=512 @16 + raw (16 ch/group: per-thread coefficients)', 3, 256, 256, 16, 16, raw=True)
    ok &= _gn_case('GN+SiLU cat(256,128)=384 @32 + raw', 2, 256, 128, 32, 32, raw=True)
    ok &= _gn_case('GN only C256 @16 (attention norm)', 2, 256, 0, 16, 16, silu=False)
    ok &= _gn_case('GN+SiLU C64 @32 (2 ch/group, float2 path)', 2, 64, 0, 32, 32)
    ok &= _gn_case('GN+SiLU cat(128,64)=192 @32 (6 ch/group)', 2, 128, 64, 32, 32, raw=True)
    ok &= _gn_case('AdaGN+SiLU C256 @8', 3, 256, 0, 8, 8, adagn=True)
    ok &= _gn_case('GN+SiLU+avgpool C128 @32->16', 2, 128, 0, 32, 32, resample=1)
    ok &= _gn_case('GN+SiLU+nearest2x C256 @8->16', 2, 256, 0, 8, 8, resample=2)
    ok &= _gn_case('GN+SiLU eps=1e-6 C128 @16 (pesser)', 2, 128, 0, 16, 16, eps=1e-6)
    return ok


# ----------------------------------------------------------------------------------------------------
# misc kernels
# ----------------------------------------------------------------------------------------------------
def case_misc():
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    for (B, Cin, Cout, H) in [(3, 3, 128, 32), (2, 1, 64, 32)]:
        x = _gen(B, Cin, H, H, seed=1)
        w = _gen(Cout, Cin, 3, 3, seed=2, scale=0.3)
        b = _gen(Cout, seed=3)
        ref = F.conv2d(x, w, b, padding=1)
        out = torch.full((B, H, H, Cout), float('nan'), device=DEV)
        st = K.new_stats(B, Cout, DEV)
        K.conv3x3_first(x, w, b, out, st)
        torch.cuda.synchronize()
        ok &= _report(f'first conv {Cin}->{Cout} @{H}', out.permute(0, 3, 1, 2), ref, 1e-5, 1e-5)
        ok &= _report(f'first conv {Cin}->{Cout} @{H} [fused GN stats]', K.stats_to_float(st).float(),
                      torch.stack([ref.sum(dim=(2, 3)), (ref * ref).sum(dim=(2, 3))], dim=-1), 2e-4, 2e-2)
    x = _gen(2, 8, 8, 128, seed=4)
    o = torch.empty(2, 8, 8, 128, device=DEV, dtype=torch.bfloat16)
    K.cast_bf16(x, o, 2, 8, 8, 128)
    ok &= _report('cast bf16', o, x.to(torch.bfloat16), 0, 0)
    o = torch.empty(2, 4, 4, 4, 128, device=DEV, dtype=torch.bfloat16)
    K.cast_bf16(x, o, 2, 8, 8, 128, parity_split=True)
    refp = torch.stack([x[:, a::2, b::2] for a in range(2) for b in range(2)], dim=1).to(torch.bfloat16)
    ok &= _report('cast bf16 parity planes', o, refp, 0, 0)
    # fast 8-channel kernel (power-of-two H, W, C/8; non-square, several grid-stride iterations) and the generic fallback
    for (Bc, Hc, Wc, Cc) in ((5, 16, 32, 256), (300, 32, 32, 128), (2, 6, 10, 12)):
        xc = _gen(Bc, Hc, Wc, Cc, seed=5)
        oc = torch.empty(Bc, Hc, Wc, Cc, device=DEV, dtype=torch.bfloat16)
        K.cast_bf16(xc, oc, Bc, Hc, Wc, Cc)
        ok &= _report(f'cast bf16 {Bc}x{Hc}x{Wc}x{Cc}', oc, xc.to(torch.bfloat16), 0, 0)
        oc = torch.empty(Bc, 4, Hc // 2, Wc // 2, Cc, device=DEV, dtype=torch.bfloat16)
        K.cast_bf16(xc, oc, Bc, Hc, Wc, Cc, parity_split=True)
        refp = torch.stack([xc[:, a::2, b::2] for a in range(2) for b in range(2)], dim=1).to(torch.bfloat16)
        ok &= _report(f'cast bf16 parity planes {Bc}x{Hc}x{Wc}x{Cc}', oc, refp, 0, 0)
    o = torch.empty(2, 4, 4, 128, device=DEV)
    K.avgpool2_f32(x, o, 2, 8, 8, 128)
    ok &= _report('avgpool2 f32', o.permute(0, 3, 1, 2), F.avg_pool2d(x.permute(0, 3, 1, 2), 2, 2), 1e-6, 1e-6)
    o = torch.empty(2, 16, 16, 128, device=DEV)
    K.upsample2_f32(x, o, 2, 8, 8, 128)
    ok &= _report('upsample2 f32', o.permute(0, 3, 1, 2),
                  F.interpolate(x.permute(0, 3, 1, 2), scale_factor=2, mode='nearest'), 0, 0)
    # time embedding MLP
    for (dim, cos_first) in [(128, False), (64, False), (256, True)]:
        E, Bn = dim * 4, 5
        t = torch.tensor([0, 1, 500, 980, 999], device=DEV)
        half = dim // 2
        if cos_first:
            freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32) / half).to(DEV)
        else:
            freqs = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1))).to(DEV)
        w1, b1 = _gen(E, dim, seed=5, scale=0.1), _gen(E, seed=6, scale=0.1)
        w2, b2 = _gen(E, E, seed=7, scale=0.05), _gen(E, seed=8, scale=0.1)
        ce = _gen(10, E, seed=9)
        y = torch.tensor([3, 0, 9, 1, 1], device=DEV)
        ang = t[:, None].float() * freqs[None]
        pe = torch.cat([ang.cos(), ang.sin()], -1) if cos_first else torch.cat([ang.sin(), ang.cos()], -1)
        ref = F.silu(pe @ w1.t() + b1) @ w2.t() + b2 + ce[y]
        out = torch.empty(Bn, E, device=DEV)
        outs = torch.empty(Bn, E, device=DEV, dtype=torch.bfloat16)
        K.time_embed(t, freqs, dim, E, cos_first, w1, b1, w2, b2, out, y=y, class_embed=ce, out_silu_bf16=outs)
        torch.cuda.synchronize()
        ok &= _report(f'time_embed dim={dim} cos_first={cos_first}', out, ref, 1e-4, 1e-4)
        ok &= _report(f'time_embed silu bf16 dim={dim}', outs, F.silu(ref), 1e-2, 1e-2)
    # negative labels = unconditional rows (batched classifier-free guidance, models/runner.py)
    dim, E, Bn = 128, 512, 6
    w1, b1, w2, b2 = _gen(E, dim, seed=31, scale=0.05), _gen(E, seed=32), _gen(E, E, seed=33, scale=0.03), _gen(E, seed=34)
    ce = _gen(10, E, seed=35)
    tt = torch.tensor([3, 500, 999, 3, 500, 999], device=DEV)
    yy = torch.tensor([1, 7, 9, -1, -1, -1], device=DEV)
    half = dim // 2
    fr = torch.exp(torch.arange(half, device=DEV) * -(math.log(10000) / (half - 1))).float().contiguous()
    o_mixed = torch.empty(Bn, E, device=DEV)
    K.time_embed(tt, fr, dim, E, False, w1, b1, w2, b2, o_mixed, y=yy, class_embed=ce)
    o_none = torch.empty(Bn, E, device=DEV)
    K.time_embed(tt, fr, dim, E, False, w1, b1, w2, b2, o_none)
    o_cond = torch.empty(Bn, E, device=DEV)
    K.time_embed(tt, fr, dim, E, False, w1, b1, w2, b2, o_cond, y=yy.clamp_min(0), class_embed=ce)
    torch.cuda.synchronize()
    neg_ok = bool(torch.equal(o_mixed[3:], o_none[3:]) and torch.equal(o_mixed[:3], o_cond[:3]))
    print(json.dumps({'case': 'time_embed: label -1 rows == unconditional rows, labelled rows unchanged (bitwise)', 'ok': neg_ok}))
    ok &= neg_ok
    # diffuse
    x0, eps = _gen(4, 3, 8, 8, seed=1), _gen(4, 3, 8, 8, seed=2)
    ac = torch.cumprod(1 - torch.linspace(1e-4, 0.02, 1000, dtype=torch.float64), 0).float().to(DEV)
    t = torch.tensor([0, 10, 500, 999], device=DEV)
    ref = (ac[t] ** 0.5)[:, None, None, None] * x0 + ((1. - ac[t]) ** 0.5)[:, None, None, None] * eps
    out = torch.empty_like(x0)
    K.diffuse(x0, eps, t, ac, out)
    ok &= _report('diffuse', out, ref, 1e-6, 1e-6)
    # out-of-range timesteps (the reference raises an IndexError): never read out of bounds, the sample becomes NaN
    t_bad = torch.tensor([0, 1000, -1, 999], device=DEV)
    out2 = torch.zeros_like(x0)
    K.diffuse(x0, eps, t_bad, ac, out2)
    torch.cuda.synchronize()
    bad_ok = bool(torch.isnan(out2[1]).all() and torch.isnan(out2[2]).all() and torch.equal(out2[0], out[0])
                  and torch.isfinite(out2[3]).all())
    print(json.dumps({'case': 'diffuse out-of-range t -> NaN sample, no OOB read', 'ok': bad_ok}))
    ok &= bad_ok
    # wrappers reject tensors the kernels would misread
    for bad in (lambda: K.diffuse(x0.double(), eps, t, ac, out), lambda: K.diffuse(x0, eps.permute(0, 1, 3, 2), t, ac, out),
                lambda: K.sampler_step(x0.half(), x0, ac[:12].contiguous(), sample=out),
                lambda: K.sampler_step(x0, x0, ac[:12].contiguous(), noise=eps[:, :, ::2], sample=out)):
        try:
            bad()
            ok = False
            print(json.dumps({'case': 'wrapper accepted a mistyped / non-contiguous tensor', 'ok': False}))
        except RuntimeError:
            pass
    return ok


# ----------------------------------------------------------------------------------------------------
# attention
# ----------------------------------------------------------------------------------------------------
def _attn_case(name, B, T, heads, d):
    torch.backends.cuda.matmul.allow_tf32 = False
    C = heads * d
    q = _bf16r(_gen(B, T, C, seed=1))
    k = _bf16r(_gen(B, T, C, seed=2))
    v = _bf16r(_gen(B, T, C, seed=3))
    scale = d ** -0.5
    qh = q.view(B, T, heads, d).transpose(1, 2)
    kh = k.view(B, T, heads, d).transpose(1, 2)
    vh = v.view(B, T, heads, d).transpose(1, 2)
    att = torch.softmax((qh @ kh.transpose(-1, -2)) * scale, dim=-1)
    ref = (att @ vh).transpose(1, 2).reshape(B, T, C)
    qk = torch.cat([q, k], dim=-1).to(torch.bfloat16).contiguous()        # [B, T, 2C]
    vt = v.transpose(1, 2).contiguous().to(torch.bfloat16)                # [B, C, T]
    out = torch.full((B, T, C), float('nan'), device=DEV, dtype=torch.bfloat16)
    K.attention(qk, 2 * C, 0, C, vt, out, C, B, T, heads, d, scale)
    torch.cuda.synchronize()
    # P and the output are rounded to bf16: 2^-8 relative each
    return _report(name, out, ref, rtol=2e-2, atol=1e-2)


def case_attention():
    ok = _attn_case('attention T=256 h=1 d=256 (CIFAR UNet)', 3, 256, 1, 256)
    ok &= _attn_case('attention T=16 h=1 d=256 (bottleneck)', 5, 16, 1, 256)
    ok &= _attn_case('attention T=256 h=4 d=64 (CFG UNet)', 2, 256, 4, 64)
    ok &= _attn_case('attention T=64 h=4 d=64 (CFG UNet 8x8)', 3, 64, 4, 64)
    ok &= _attn_case('attention T=256 h=2 d=128', 2, 256, 2, 128)
    ok &= _attn_case('attention T=1024 h=8 d=64 (ADM 32x32, KV loop)', 2, 1024, 8, 64)
    ok &= _attn_case('attention T=576 h=2 d=64 (ragged last KV tile)', 2, 576, 2, 64)
    ok &= _attn_case('attention T=256 h=1 d=512 (pesser 16x16, wide head)', 3, 256, 1, 512)
    ok &= _attn_case('attention T=64 h=1 d=512 (pesser 8x8)', 2, 64, 1, 512)
    ok &= _attn_case('attention T=144 h=2 d=320 (wide, ragged)', 2, 144, 2, 320)
    return ok


def _attn_bwd_case(name, B, T, heads):
    """b200_attention_fwd_lse + b200_attention_bwd against fp32 autograd of softmax(q k^T * scale) v on the same
    bf16-rounded q, k, v and output gradient."""
    torch.backends.cuda.matmul.allow_tf32 = False
    d = 64
    C = heads * d
    q = _bf16r(_gen(B, T, C, seed=11)).requires_grad_(True)
    k = _bf16r(_gen(B, T, C, seed=12)).requires_grad_(True)
    v = _bf16r(_gen(B, T, C, seed=13)).requires_grad_(True)
    g = _bf16r(_gen(B, T, C, seed=14))
    scale = d ** -0.5
    qh = q.view(B, T, heads, d).transpose(1, 2)
    kh = k.view(B, T, heads, d).transpose(1, 2)
    vh = v.view(B, T, heads, d).transpose(1, 2)
    s = (qh @ kh.transpose(-1, -2)) * scale
    att = torch.softmax(s, dim=-1)
    ref = (att @ vh).transpose(1, 2).reshape(B, T, C)
    ref.backward(g)
    lse_ref = torch.logsumexp(s, dim=-1) * 1.4426950408889634            # [B, heads, T], base 2
    bf = torch.bfloat16
    qk = torch.cat([q, k], dim=-1).detach().to(bf).contiguous()
    vt = v.detach().transpose(1, 2).contiguous().to(bf)
    o = torch.full((B, T, C), float('nan'), device=DEV, dtype=bf)
    lse = torch.full((B, heads, T), float('nan'), device=DEV, dtype=torch.float32)
    K.attention(qk, 2 * C, 0, C, vt, o, C, B, T, heads, d, scale, lse=lse)
    dqk = torch.full((B, T, 2 * C), float('nan'), device=DEV, dtype=bf)
    dv = torch.full((B, T, C), float('nan'), device=DEV, dtype=bf)
    K.attention_bwd(qk, vt, o, g.to(bf).contiguous(), lse, dqk, dv, B, T, heads, d, scale)
    torch.cuda.synchronize()
    ok = _report(name + ' lse', lse, lse_ref, rtol=1e-4, atol=1e-3)
    ok &= _report(name + ' out', o, ref.detach(), rtol=2e-2, atol=1e-2)
    # P, dS and the results are rounded to bf16 (2^-8 relative each); gate on the tensor-level error as well
    for nm, got, want in (('dq', dqk[:, :, :C], q.grad), ('dk', dqk[:, :, C:], k.grad), ('dv', dv, v.grad)):
        rel = ((got.float() - want).norm() / want.norm()).item()
        print(f'    {name} {nm}: rel-L2 {rel:.3e}')
        ok &= rel < 1e-2
        ok &= _report(f'{name} {nm}', got, want, rtol=3e-2, atol=3e-2 * want.abs().max().item())
    return ok


def case_attention_bwd():
    ok = _attn_bwd_case('attention_bwd T=256 h=4 (CFG UNet 16x16)', 3, 256, 4)
    ok &= _attn_bwd_case('attention_bwd T=64 h=4 (8x8)', 2, 64, 4)
    ok &= _attn_bwd_case('attention_bwd T=16 h=4 (4x4 bottleneck)', 5, 16, 4)
    ok &= _attn_bwd_case('attention_bwd T=144 h=2 (ragged second tile)', 2, 144, 2)
    ok &= _attn_bwd_case('attention_bwd T=256 h=1', 130, 256, 1)
    return ok


def _attn_block_case(name, B, debug):
    """b200_attn_block_fwd (T=256, C=256, one head) against the fp32 op sequence of models/modules.py:89-102; with
    `debug` every on-chip intermediate (xn, q, k, v^T, P, o) is dumped and checked too, which localises a failure."""
    torch.backends.cuda.matmul.allow_tf32 = False
    T = C = 256
    x = _gen(B, T, C, seed=1) * (_gen(1, 1, C, seed=2).abs() + 0.5) + _gen(1, 1, C, seed=3)
    gamma = _gen(C, seed=4) * 0.2 + 1.0
    beta = _gen(C, seed=5) * 0.2
    w = _gen(4 * C, C, seed=6) * (C ** -0.5)
    w[:2 * C] *= 1.5                      # sharper softmax than the default initialisation gives
    bias = _gen(4 * C, seed=7) * 0.3
    eps, scale = 1e-5, C ** -0.5
    wb = w.to(torch.bfloat16).contiguous()
    wf = wb.float()
    st = K.stats_from_float(torch.stack([x.sum(dim=1), (x * x).sum(dim=1)], dim=-1).contiguous())
    # fp32 restatement, rounding to bf16 exactly where the kernel does
    xn = _bf16r(F.group_norm(x.transpose(1, 2), 32, gamma, beta, eps).transpose(1, 2))       # [B, T, C]
    q = _bf16r(xn @ wf[:C].t() + bias[:C])
    k = _bf16r(xn @ wf[C:2 * C].t())      # the kernel drops the k bias: it shifts every score of a row equally
    v = _bf16r(xn @ wf[2 * C:3 * C].t() + bias[2 * C:3 * C])
    s = (q @ k.transpose(1, 2)) * scale
    pu = _bf16r(torch.exp(s - s.max(dim=-1, keepdim=True).values))                            # unnormalised, as stored
    o = _bf16r((pu @ v) / torch.exp(s - s.max(dim=-1, keepdim=True).values).sum(dim=-1, keepdim=True))
    ref = x + o @ wf[3 * C:].t() + bias[3 * C:]
    out = torch.full((B, T, C), float('nan'), device=DEV)
    ost = K.new_stats(B, C, DEV)
    dbg = [torch.full((B, T, C), float('nan'), device=DEV, dtype=torch.bfloat16) for _ in range(6)] if debug else None
    K.attn_block(x.contiguous(), st, gamma, beta, eps, wb, bias, out, ost, B, T, C, 1, 32, scale, dbg=dbg)
    torch.cuda.synchronize()
    ok = True
    if debug:
        for nm, got, want, tol in (('xn', dbg[0], xn, 1e-2), ('q', dbg[1], q, 2e-2), ('k', dbg[2], k, 2e-2),
                                   ('v^T', dbg[3].transpose(1, 2), v, 2e-2), ('P', dbg[4], pu, 3e-2), ('o', dbg[5], o, 3e-2)):
            ok &= _report(f'{name} [{nm}]', got, want, rtol=tol, atol=tol)
    ok &= _report(name, out, ref, rtol=2e-2, atol=2e-2)
    got_st = K.stats_to_float(ost)
    want_st = torch.stack([out.sum(dim=1), (out * out).sum(dim=1)], dim=-1)
    ok &= _report(name + ' [statistics of the result]', got_st, want_st, rtol=1e-4, atol=1e-2)
    return ok


def case_attn_block():
    ok = _attn_block_case('attn_block B=3 (debug dumps)', 3, True)
    ok &= _attn_block_case('attn_block B=2', 2, False)
    ok &= _attn_block_case('attn_block B=160 (persistent: 3 images per cluster, ragged)', 160, False)
    return ok


# ----------------------------------------------------------------------------------------------------
# backward-pass GEMMs: batched GEMM in all major combinations, conv weight gradient
# ----------------------------------------------------------------------------------------------------
def _gemm_case(name, B, H, M, N, Kd, a_mn, b_mn, out_dtype=torch.float32, split_k=1):
    """D[b,h] = A[b,h] (M x K) @ B[b,h]^T (N x K), operands stored with the heads side by side in the columns."""
    torch.backends.cuda.matmul.allow_tf32 = False
    A = _bf16r(_gen(B, H, M, Kd, seed=1))
    Bm = _bf16r(_gen(B, H, N, Kd, seed=2))
    ref = torch.einsum('bhmk,bhnk->bhmn', A, Bm)
    # storage: K-major operand [B][rows = M][H*K]; MN-major operand [B][rows = K][H*M]
    if a_mn:
        a_t = A.permute(0, 3, 1, 2).reshape(B, Kd, H * M).to(torch.bfloat16).contiguous()
        a = (a_t, Kd, H * M, dict(col_head=M, mn_major=True))
    else:
        a_t = A.permute(0, 2, 1, 3).reshape(B, M, H * Kd).to(torch.bfloat16).contiguous()
        a = (a_t, M, H * Kd, dict(col_head=Kd))
    if b_mn:
        b_t = Bm.permute(0, 3, 1, 2).reshape(B, Kd, H * N).to(torch.bfloat16).contiguous()
        b = (b_t, Kd, H * N, dict(col_head=N, mn_major=True))
    else:
        b_t = Bm.permute(0, 2, 1, 3).reshape(B, N, H * Kd).to(torch.bfloat16).contiguous()
        b = (b_t, N, H * Kd, dict(col_head=Kd))
    out = torch.zeros(B, H, M, N, device=DEV, dtype=out_dtype) if split_k > 1 else \
        torch.full((B, H, M, N), float('nan'), device=DEV, dtype=out_dtype)
    K.gemm_batched(a, b, out, M, N, Kd, batch=B, heads=H, out_ld=N, out_batch_stride=H * M * N, out_head_stride=M * N,
                   split_k=split_k)
    torch.cuda.synchronize()
    tol = dict(rtol=1e-2, atol=2e-2 * Kd ** 0.5) if out_dtype == torch.bfloat16 else dict(rtol=1e-4, atol=1e-3)
    return _report(name, out, ref, **tol)


def case_gemm():
    ok = True
    for a_mn in (False, True):
        for b_mn in (False, True):
            tag = f"A {'MN' if a_mn else 'K'}-major, B {'MN' if b_mn else 'K'}-major"
            ok &= _gemm_case(f'gemm 256x256x256 h=2 ({tag})', 3, 2, 256, 256, 256, a_mn, b_mn)
            ok &= _gemm_case(f'gemm 128x64x320 h=4 ({tag})', 2, 4, 128, 64, 320, a_mn, b_mn)
            ok &= _gemm_case(f'gemm ragged 200x192x64 h=1 ({tag})', 2, 1, 200, 192, 64, a_mn, b_mn)
            ok &= _gemm_case(f'gemm tiny 16x64x128 h=2 bf16 out ({tag})', 2, 2, 16, 64, 128, a_mn, b_mn,
                             out_dtype=torch.bfloat16)
    ok &= _gemm_case('gemm split-K 128x256x4096', 1, 1, 128, 256, 4096, True, True, split_k=8)
    ok &= _gemm_case('gemm ragged K=16 h=1', 4, 1, 16, 64, 16, False, True)
    return ok


def _wgrad_case(name, B, H, W, Cin, Cout, kind, scratch=False):
    """dW of a conv executed by b200_conv2d_fwd vs autograd of F.conv2d on the bf16-rounded operands."""
    torch.backends.cudnn.allow_tf32 = False
    x = _bf16r(_gen(B, Cin, H, W, seed=1))
    stride, pad, k = (2, 1, 3) if kind == 's2' else (1, 1, 3) if kind == '3x3' else (1, 0, 1)
    Ho, Wo = H // stride, W // stride
    dy = _bf16r(_gen(B, Cout, Ho, Wo, seed=2) * 0.1)
    w = torch.zeros(Cout, Cin, k, k, device=DEV, requires_grad=True)
    F.conv2d(x, w, None, stride=stride, padding=pad).backward(dy)
    ref = w.grad
    dyb = dy.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    xb = x.permute(0, 2, 3, 1).contiguous()
    dw = torch.zeros(Cout, Cin, k, k, device=DEV)
    ws = torch.full((16 << 20,), float('nan'), device=DEV) if scratch else None     # two-phase split reduction
    if kind == 's2':
        planes = torch.empty(B, 4, H // 2, W // 2, Cin, device=DEV, dtype=torch.bfloat16)
        K.cast_bf16(xb, planes, B, H, W, Cin, parity_split=True)
        K.conv2d_wgrad(dyb, Cout, planes, (Cin, H // 2, W // 2, 4), B, Ho, Wo, Cout, Cin, K.taps_3x3_s2(1)[0], dw, scratch=ws)
    else:
        taps = K.taps_3x3_s1()[0] if kind == '3x3' else K.taps_1x1()[0]
        K.conv2d_wgrad(dyb, Cout, xb.to(torch.bfloat16), (Cin, H, W, 1), B, Ho, Wo, Cout, Cin, taps, dw, scratch=ws)
    torch.cuda.synchronize()
    return _report(name, dw, ref, rtol=1e-3, atol=1e-3 * float(ref.abs().max()))


def case_wgrad():
    ok = _wgrad_case('wgrad 3x3 128->128 @32x32 B=4', 4, 32, 32, 128, 128, '3x3')
    ok &= _wgrad_case('wgrad 3x3 384->256 @16x16 B=3', 3, 16, 16, 384, 256, '3x3')
    ok &= _wgrad_case('wgrad 3x3 256->256 @4x4 B=6 (stacked images, ragged batch)', 6, 4, 4, 256, 256, '3x3')
    ok &= _wgrad_case('wgrad 3x3 64->64 @8x8 B=5', 5, 8, 8, 64, 64, '3x3')
    ok &= _wgrad_case('wgrad 1x1 512->256 @8x8 B=4', 4, 8, 8, 512, 256, '1x1')
    ok &= _wgrad_case('wgrad 3x3 stride 2 128->128 @32x32 B=2', 2, 32, 32, 128, 128, 's2')
    ok &= _wgrad_case('wgrad two-phase 3x3 256->256 @16x16 B=8', 8, 16, 16, 256, 256, '3x3', scratch=True)
    ok &= _wgrad_case('wgrad two-phase 3x3 384->128 @32x32 B=3', 3, 32, 32, 384, 128, '3x3', scratch=True)
    ok &= _wgrad_case('wgrad two-phase 1x1 512->256 @8x8 B=4', 4, 8, 8, 512, 256, '1x1', scratch=True)
    ok &= _wgrad_case('wgrad two-phase stride 2 128->128 @32x32 B=2', 2, 32, 32, 128, 128, 's2', scratch=True)
    ok &= _wgrad_case('wgrad two-phase 3x3 64->64 @8x8 B=5', 5, 8, 8, 64, 64, '3x3', scratch=True)
    return ok


# ----------------------------------------------------------------------------------------------------
# GroupNorm(+AdaGN)(+SiLU)(+dropout)(+resample) backward vs autograd of the fp32 restatement
# ----------------------------------------------------------------------------------------------------
def _gn_bwd_case(name, B, H, W, C0, C1, *, silu=True, adagn=False, resample=0, drop_p=0.0, bf16_out=False, addend=False,
                 eps=1e-5):
    C = C0 + C1
    x0 = _gen(B, H, W, C0, seed=1) * 1.5 + 0.3
    x1 = (_gen(B, H, W, C1, seed=2) * 0.7 - 0.2) if C1 else None
    gamma = (_gen(C, seed=3) * 0.2 + 1.0)
    beta = _gen(C, seed=4) * 0.1
    scale = _gen(B, C, seed=5) * 0.3 if adagn else None
    shift = _gen(B, C, seed=6) * 0.3 if adagn else None
    Ho, Wo = (H // 2, W // 2) if resample == 1 else (2 * H, 2 * W) if resample == 2 else (H, W)
    g = _bf16r(_gen(B, Ho, Wo, C, seed=7))
    ad = _gen(B, H, W, C, seed=8) if addend else None
    seed = 1234567
    mask = None
    if drop_p > 0:
        mask = K.dropout_mask(torch.empty(B, H, W, C, device=DEV), drop_p, seed)
    # ---- fp32 autograd restatement (NHWC tensors, channels last dim) ----
    xs = [x0.clone().requires_grad_(True)] + ([x1.clone().requires_grad_(True)] if C1 else [])
    gm, bt = gamma.clone().requires_grad_(True), beta.clone().requires_grad_(True)
    sc = scale.clone().requires_grad_(True) if adagn else None
    sh = shift.clone().requires_grad_(True) if adagn else None
    xc = torch.cat(xs, dim=-1).permute(0, 3, 1, 2)
    y = F.group_norm(xc, 32, gm, bt, eps)
    if adagn:
        y = y * (1 + sc[:, :, None, None]) + sh[:, :, None, None]
    if silu:
        y = F.silu(y)
    if mask is not None:
        y = y * mask.permute(0, 3, 1, 2) / (1 - drop_p)
    if resample == 1:
        y = F.avg_pool2d(y, 2, 2)
    elif resample == 2:
        y = F.interpolate(y, scale_factor=2, mode='nearest')
    y.backward(g.permute(0, 3, 1, 2))
    # ---- kernels: statistics as the conv epilogue would deliver them ----
    def stats(x):
        return K.stats_from_float(torch.stack([x.sum(dim=(1, 2)), (x * x).sum(dim=(1, 2))], dim=-1).contiguous())
    st0 = stats(x0)
    st1 = stats(x1) if C1 else None
    sums = torch.empty(B, 8, C, device=DEV)
    dgamma, dbeta = torch.zeros(C, device=DEV), torch.zeros(C, device=DEV)
    dss = torch.full((B, 2 * C), float('nan'), device=DEV) if adagn else None
    kw = dict(scale=scale, shift=shift, ss_ld=C if adagn else 0, silu=silu, resample=resample, drop_p=drop_p,
              drop_seed=seed, dgamma=dgamma, dbeta=dbeta, dscale=dss, dshift=dss[:, C:] if adagn else None,
              dss_ld=2 * C)
    gb = g.to(torch.bfloat16).contiguous()
    ok = True
    if bf16_out:
        dxb = torch.full((B, H, W, C), float('nan'), device=DEV, dtype=torch.bfloat16)
        rowsum = torch.zeros(B, C, device=DEV)
        K.groupnorm_bwd(gb, x0, C0, st0, x1, C1, st1, B, H * W, W, 32, gamma, beta, eps, sums, dx_bf16=dxb,
                        dx_rowsum=rowsum, **kw)
        torch.cuda.synchronize()
        ref = torch.cat([t.grad for t in xs], dim=-1)
        ok &= _report(name + ' dx(bf16)', dxb, ref, rtol=2e-2, atol=2e-2 * float(ref.abs().max()))
        ok &= _report(name + ' dx rowsum', rowsum, ref.sum(dim=(1, 2)), rtol=1e-2, atol=2e-2 * float(ref.abs().max()) * (H * W) ** 0.5)
    else:
        dx0 = torch.full((B, H, W, C0), 0.5, device=DEV)       # accumulate on top of an existing gradient
        dx1 = torch.full((B, H, W, C1), float('nan'), device=DEV) if C1 else None
        K.groupnorm_bwd(gb, x0, C0, st0, x1, C1, st1, B, H * W, W, 32, gamma, beta, eps, sums, dx0=dx0, dx0_acc=True,
                        dx1=dx1, dx1_acc=False, addend=ad, **kw)
        torch.cuda.synchronize()
        r0 = xs[0].grad + 0.5 + (ad[..., :C0] if addend else 0)
        ok &= _report(name + ' dx0 (accumulated)', dx0, r0, rtol=1e-3, atol=2e-3 * float(r0.abs().max()))
        if C1:
            r1 = xs[1].grad + (ad[..., C0:] if addend else 0)
            ok &= _report(name + ' dx1', dx1, r1, rtol=1e-3, atol=2e-3 * float(r1.abs().max()))
    ok &= _report(name + ' dgamma', dgamma, gm.grad, rtol=1e-3, atol=2e-3 * float(gm.grad.abs().max()))
    ok &= _report(name + ' dbeta', dbeta, bt.grad, rtol=1e-3, atol=2e-3 * float(bt.grad.abs().max()))
    if adagn:
        ok &= _report(name + ' dscale', dss[:, :C], sc.grad, rtol=1e-3, atol=2e-3 * float(sc.grad.abs().max()))
        ok &= _report(name + ' dshift', dss[:, C:], sh.grad, rtol=1e-3, atol=2e-3 * float(sh.grad.abs().max()))
    return ok


def case_groupnorm_bwd():
    ok = _gn_bwd_case('gn_bwd C128 32x32', 3, 32, 32, 128, 0)
    ok &= _gn_bwd_case('gn_bwd cat 256+128 16x16 addend', 2, 16, 16, 256, 128, addend=True)
    ok &= _gn_bwd_case('gn_bwd C256 8x8 no-silu (attention norm)', 4, 8, 8, 256, 0, silu=False, addend=True)
    ok &= _gn_bwd_case('gn_bwd adagn C256 16x16 bf16 out + rowsum', 3, 16, 16, 256, 0, adagn=True, bf16_out=True)
    ok &= _gn_bwd_case('gn_bwd dropout 0.1 C128 16x16 bf16 out', 2, 16, 16, 128, 0, drop_p=0.1, bf16_out=True)
    ok &= _gn_bwd_case('gn_bwd avgpool C128 16x16', 2, 16, 16, 128, 0, resample=1)
    ok &= _gn_bwd_case('gn_bwd nearest-up C256 8x8', 2, 8, 8, 256, 0, resample=2)
    ok &= _gn_bwd_case('gn_bwd C64 4x4 eps 1e-6', 5, 4, 4, 64, 0, eps=1e-6)
    ok &= _gn_bwd_case('gn_bwd cat 1024+512 8x8', 1, 8, 8, 1024, 512)
    # one-launch slab form (HW <= 256): concatenated sources with dropout + addend, AdaGN with dropout, a 32x32 pair that
    # keeps the four-kernel form covered with the same options
    ok &= _gn_bwd_case('gn_bwd cat 256+256 4x4 dropout addend', 6, 4, 4, 256, 256, drop_p=0.1, addend=True)
    ok &= _gn_bwd_case('gn_bwd adagn C512 8x8 dropout', 3, 8, 8, 512, 0, adagn=True, drop_p=0.2)
    ok &= _gn_bwd_case('gn_bwd adagn cat 128+128 16x16 bf16 out + rowsum', 2, 16, 16, 128, 128, adagn=True, bf16_out=True)
    ok &= _gn_bwd_case('gn_bwd adagn cat 128+128 32x32 dropout addend', 2, 32, 32, 128, 128, adagn=True, drop_p=0.1, addend=True)
    # 32x32 / 64x64 images (four-kernel form; the slab form only with B200_GNB_SLAB_HW=1024): 12-channel groups straddling
    # the boundary of the two sources, bf16 output + row sums, dropout, AdaGN
    ok &= _gn_bwd_case('gn_bwd cat 256+128 32x32 bf16 out + rowsum', 3, 32, 32, 256, 128, bf16_out=True)
    ok &= _gn_bwd_case('gn_bwd cat 256+128 32x32 dropout addend', 2, 32, 32, 256, 128, drop_p=0.1, addend=True)
    ok &= _gn_bwd_case('gn_bwd C128 64x64 dropout (four-kernel form)', 1, 64, 64, 128, 0, drop_p=0.1)
    ok &= _gn_bwd_case('gn_bwd adagn cat 128+128 64x64 addend (four-kernel form)', 1, 64, 64, 128, 128, adagn=True, addend=True)
    return ok


def case_backward_misc():
    ok = True
    # cast + column sums
    x = _gen(1000, 256, seed=1)
    out = torch.empty(1000, 256, device=DEV, dtype=torch.bfloat16)
    cs = torch.ones(256, device=DEV)
    K.cast_bf16_colsum(x, out, cs, 1000, 256)
    ok &= _report('cast_bf16_colsum cast', out, x.to(torch.bfloat16), rtol=0, atol=0)
    ok &= _report('cast_bf16_colsum sums (accumulate)', cs, x.sum(0) + 1, rtol=1e-4, atol=1e-3)
    # NCHW -> padded NHWC
    xi = _gen(3, 3, 64, seed=2)
    po = torch.full((3, 64, 64), float('nan'), device=DEV, dtype=torch.bfloat16)
    cs3 = torch.zeros(3, device=DEV)
    K.nchw_to_nhwc_pad_bf16(xi, po, cs3, 3, 3, 64, 64)
    ref = torch.zeros(3, 64, 64, device=DEV)
    ref[..., :3] = xi.permute(0, 2, 1)
    ok &= _report('nchw_to_nhwc_pad', po, ref.to(torch.bfloat16), rtol=0, atol=0)
    ok &= _report('nchw_to_nhwc_pad colsum', cs3, xi.sum(dim=(0, 2)), rtol=1e-4, atol=1e-3)
    # bf16 column sums of a window
    xb = _gen(777, 512, seed=3).to(torch.bfloat16)
    cw = torch.zeros(256, device=DEV)
    K.colsum_bf16(xb, cw, 777, 512, 128, 256)
    ok &= _report('colsum_bf16 window', cw, xb[:, 128:384].float().sum(0), rtol=1e-4, atol=1e-2)
    # resampling adjoints
    xr = _gen(2, 8, 8, 64, seed=4)
    o1 = torch.ones(2, 4, 4, 64, device=DEV)
    K.resample_f32(xr, o1, 2, 8, 8, 64, 1, scale=4.0, accumulate=True)
    ref1 = F.avg_pool2d(xr.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1) * 4 + 1
    ok &= _report('resample avg (sum-pool, accumulate)', o1, ref1, rtol=1e-5, atol=1e-5)
    o2 = torch.empty(2, 16, 16, 64, device=DEV)
    K.resample_f32(xr, o2, 2, 8, 8, 64, 2, scale=0.25)
    ref2 = F.interpolate(xr.permute(0, 3, 1, 2), scale_factor=2, mode='nearest').permute(0, 2, 3, 1) * 0.25
    ok &= _report('resample nearest x0.25', o2, ref2, rtol=1e-6, atol=1e-6)
    o3 = torch.empty(2, 16, 16, 64, device=DEV, dtype=torch.bfloat16)
    K.upsample2_bf16(xr, o3, 2, 8, 8, 64)
    ok &= _report('upsample2_bf16', o3, (ref2 * 4).to(torch.bfloat16), rtol=0, atol=0)
    # softmax rows fwd / bwd
    S = _gen(6, 100, 100, seed=5) * 3
    P = torch.empty(6, 100, 100, device=DEV, dtype=torch.bfloat16)
    K.softmax_rows(S, P, 600, 100, 0.125)
    Pr = torch.softmax(S * 0.125, dim=-1)
    ok &= _report('softmax_rows', P, Pr, rtol=1e-2, atol=1e-4)
    dP = _gen(6, 100, 100, seed=6)
    dS = torch.empty_like(P)
    K.softmax_bwd_rows(P, dP, dS, 600, 100, 0.125)
    Pf = P.float()
    ref = 0.125 * Pf * (dP - (dP * Pf).sum(-1, keepdim=True))
    ok &= _report('softmax_bwd_rows', dS, ref, rtol=1e-2, atol=1e-4)
    # MSE loss + gradient
    a, b = _gen(4, 3, 32, 32, seed=7), _gen(4, 3, 32, 32, seed=8)
    loss = torch.empty(1, device=DEV)
    K.mse_loss(a, b, loss)
    ok &= _report('mse_loss', loss, F.mse_loss(a, b).view(1), rtol=1e-5, atol=1e-6)
    da = torch.empty_like(a)
    gs = torch.tensor([0.5], device=DEV)
    K.mse_loss_grad(a, b, gs, da)
    ok &= _report('mse_loss_grad', da, 0.5 * 2 * (a - b) / a.numel(), rtol=1e-5, atol=1e-9)
    # output stage: uint8 HWC with torchvision.save_image's quantisation
    xs = _gen(5, 3, 32, 32, seed=9) * 0.8
    u8 = K.to_uint8_hwc(xs)
    ref8 = ((xs.clamp(-1, 1) + 1) / 2).mul(255).add_(0.5).clamp_(0, 255).permute(0, 2, 3, 1).to(torch.uint8)
    same = bool(torch.equal(u8, ref8))
    print(json.dumps({'case': 'to_uint8_hwc (bit-exact vs torchvision quantisation)', 'ok': same}), flush=True)
    ok &= same
    # dropout keep rate
    m = K.dropout_mask(torch.empty(1 << 20, device=DEV), 0.1, 42)
    rate = 1 - m.mean().item()
    good = abs(rate - 0.1) < 2e-3
    print(json.dumps({'case': 'dropout keep rate', 'drop_rate': rate, 'ok': good}), flush=True)
    return ok and good


# ----------------------------------------------------------------------------------------------------
# fused clip + Adam(W) + EMA vs clip_grad_norm_ + torch.optim.Adam(W) + the reference EMA update
# ----------------------------------------------------------------------------------------------------
def case_optimizer():
    sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
    from b200diff.optim import FusedAdam
    from models.ema import EMA
    ok = True
    shapes = [(256, 128, 3, 3), (128,), (70001,), (512, 512), (3, 128, 3, 3)]
    for adamw, wd, capt in ((False, 0.0, False), (False, 0.01, False), (True, 0.05, False), (False, 0.0, True),
                            (True, 0.05, True)):
        ps = [torch.nn.Parameter(_gen(*s, seed=i) * 0.1) for i, s in enumerate(shapes)]
        qs = [torch.nn.Parameter(p.detach().clone()) for p in ps]
        ours = FusedAdam(ps, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd, adamw=adamw, capturable=capt)
        ref = (torch.optim.AdamW if adamw else torch.optim.Adam)(qs, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=wd)
        ema_o, ema_r = EMA(ps, decay=0.9999, gradual=True), EMA(qs, decay=0.9999, gradual=True)
        for step in range(4):
            for i, (p, q) in enumerate(zip(ps, qs)):
                g = _gen(*p.shape, seed=100 * step + i) * (3.0 if step == 1 else 0.01)
                p.grad, q.grad = g.clone(), g.clone()
            norm_ref = torch.nn.utils.clip_grad_norm_(qs, max_norm=1.0)
            ref.step()
            ema_r.update(qs)
            v0 = ps[0]._version
            ours.step(clip_grad_norm=1.0, ema=ema_o)
            ok &= ps[0]._version > v0          # raw-pointer updates must still bump the tensors' version counters
            ok &= _report(f'fused adam{"w" if adamw else ""} wd={wd} capturable={capt} step {step} grad norm', ours.grad_norm,
                          norm_ref.view(1), rtol=1e-5, atol=1e-6)
        for i, (p, q) in enumerate(zip(ps, qs)):
            ok &= _report(f'fused adam{"w" if adamw else ""} wd={wd} param {i}', p.detach(), q.detach(), rtol=1e-5, atol=1e-7)
            ok &= _report(f'fused adam{"w" if adamw else ""} wd={wd} ema {i}', ema_o.shadow[i], ema_r.shadow[i], rtol=1e-5,
                          atol=1e-7)
        sd = ours.state_dict()
        good = set(sd['state'][0].keys()) == {'step', 'exp_avg', 'exp_avg_sq'} and ema_o.num_updates == 4
        print(json.dumps({'case': 'fused adam state_dict layout', 'ok': good}), flush=True)
        ok &= good
    return ok


# ----------------------------------------------------------------------------------------------------
# Euler / Heun ODE steps vs the reference's outputs frozen in tests/golden/ode_samplers.pt
# ----------------------------------------------------------------------------------------------------
def case_ode_samplers():
    import diffusions
    g = torch.load(os.path.join(ROOT, 'tests', 'golden', 'ode_samplers.pt'), weights_only=False)
    xt, mo, d1 = g['xt'].to(DEV), g['mo'].to(DEV), g['d1'].to(DEV)
    ok = True
    worst = 0.0
    for c in g['steps']:
        kw = dict(objective=c['objective'], clip_denoised=c['clip'], beta_schedule=c['beta'], device=DEV, **g['kw0'])
        e = diffusions.EulerSampler(**kw)
        o = e.denoise(mo.clone(), xt, c['t'], c['t_prev'])
        for k in ('sample', 'pred_x0'):
            worst = max(worst, (o[k].cpu() - c['euler'][k]).abs().max().item() / max(1.0, c['euler'][k].abs().max().item()))
        if 'heun2' in c:
            h = diffusions.HeunSampler(**kw)
            h.denoise_1st_order(mo.clone(), xt, c['t'], c['t_prev'])
            h._1st_order_derivative, h._1st_order_xt = d1.clone(), xt.clone()
            o2 = h.denoise_2nd_order(mo.clone(), xt * 0.9, c['t'], c['t_prev'])
            for k in ('sample', 'pred_x0'):
                worst = max(worst, (o2[k].cpu() - c['heun2'][k]).abs().max().item() /
                            max(1.0, c['heun2'][k].abs().max().item()))
    good = worst <= 2e-6
    print(json.dumps({'case': f'euler/heun single steps ({len(g["steps"])} cases) vs reference golden',
                      'max_rel_err': worst, 'gate': 2e-6, 'ok': good}), flush=True)
    return ok and good


def case_precise():
    """FP32-mode support kernels (csrc/precise.cu) against PyTorch: split cast layouts, GroupNorm -> split operand, row
    softmax -> split probabilities, pack mode 3, and a conv / GEMM over split operands against fp32 (TF32 off) references
    at the 1e-5 level."""
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    ok = True

    def split_ref(x):
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        return hi, lo

    # ---- split cast: channel layout with groups, both patterns, SiLU, column window
    rows, C, d = 96, 128, 32
    x = _gen(rows, 2 * C, seed=1) * 3
    for pattern in (0, 1):
        out = torch.zeros(rows, 3 * C, device=DEV, dtype=torch.bfloat16)
        K.split_cast(x, out, rows, C, in_ld=2 * C, in_col0=C, group=d, pattern=pattern)
        hi, lo = split_ref(x[:, C:])
        parts = (hi, lo, hi) if pattern == 0 else (hi, hi, lo)
        ref = torch.stack([p_.view(rows, C // d, d) for p_ in parts], dim=2).reshape(rows, 3 * C)
        ok &= _report(f'split_cast groups pattern {pattern}', out, ref, 0, 0)
    out = torch.zeros(rows, 3 * C, device=DEV, dtype=torch.bfloat16)
    K.split_cast(x[:, :C].contiguous(), out, rows, C, silu=True)
    hi, lo = split_ref(F.silu(x[:, :C]))
    rec = out[:, :C].float() + out[:, C:2 * C].float()
    ok &= _report('split_cast SiLU: hi + lo reconstructs silu(x)', rec, F.silu(x[:, :C]), 1e-5, 1e-6)
    ok &= _report('split_cast third block == hi', out[:, 2 * C:], out[:, :C], 0, 0)
    # plane layout
    B, T = 3, 16
    v = _gen(B * T, C, seed=2)
    outp = torch.zeros(B, 3, T, C, device=DEV, dtype=torch.bfloat16)
    K.split_cast(v, outp, B * T, C, pattern=1, planes_rows=T)
    hi, lo = split_ref(v.view(B, T, C))
    ok &= _report('split_cast planes', outp, torch.stack([hi, hi, lo], dim=1), 0, 0)
    # parity planes
    B, H, W, Cc = 2, 8, 8, 64
    xi = _gen(B, H, W, Cc, seed=3)
    outq = torch.zeros(B, 4, H // 2, W // 2, 3 * Cc, device=DEV, dtype=torch.bfloat16)
    K.split_cast(xi, outq, B * H * W, Cc, parity_hw=(H, W))
    hi, lo = split_ref(xi)
    full = torch.cat([hi, lo, hi], dim=-1)
    refq = torch.stack([full[:, a::2, b::2] for a in range(2) for b in range(2)], dim=1)
    ok &= _report('split_cast parity planes', outq, refq, 0, 0)

    # ---- GroupNorm -> split operand (two sources, AdaGN, resample) vs torch fp32
    for (name, B, C0, C1, H, adagn, silu, resample, raw) in (
            ('C128 @16', 2, 128, 0, 16, False, True, 0, False), ('cat(256,128) @8 + raw', 2, 256, 128, 8, False, True, 0, True),
            ('AdaGN C256 @8 avgpool', 2, 256, 0, 8, True, True, 1, False), ('AdaGN C128 @4 nearest2x', 3, 128, 0, 4, True, True, 2, False),
            ('GN only C256 @4', 2, 256, 0, 4, False, False, 0, False)):
        Cn = C0 + C1
        x0 = _gen(B, H, H, C0, seed=4) * 2 + 0.5
        x1 = _gen(B, H, H, C1, seed=5) if C1 else None
        xcat = torch.cat([x0, x1], dim=-1) if C1 else x0
        gamma, beta = _gen(Cn, seed=6) + 1, _gen(Cn, seed=7)
        ys = _gen(B, 2 * Cn, seed=8) * 0.3
        ref = F.group_norm(xcat.permute(0, 3, 1, 2), 32, gamma, beta, 1e-5)
        if adagn:
            ref = ref * (1 + ys[:, :Cn, None, None]) + ys[:, Cn:, None, None]
        if silu:
            ref = F.silu(ref)
        if resample == 1:
            ref = F.avg_pool2d(ref, 2, 2)
        elif resample == 2:
            ref = F.interpolate(ref, scale_factor=2, mode='nearest')
        Ho = H // 2 if resample == 1 else H * 2 if resample == 2 else H

        def st(t_):
            return K.stats_from_float(torch.stack([t_.double().sum(dim=(1, 2)), (t_.double() ** 2).sum(dim=(1, 2))], dim=-1))
        o3 = torch.zeros(B, Ho, Ho, 3 * Cn, device=DEV, dtype=torch.bfloat16)
        r3 = torch.zeros(B, H, H, 3 * Cn, device=DEV, dtype=torch.bfloat16) if raw else None
        K.groupnorm_apply_split(x0, C0, st(x0), x1, C1, st(x1) if C1 else None, B, H * H, H, 32, gamma, beta, 1e-5, o3,
                                scale=ys if adagn else None, shift=ys[:, Cn:] if adagn else None, ss_ld=2 * Cn if adagn else 0,
                                silu=silu, resample=resample, raw_out=r3)
        rec = (o3[..., :Cn].float() + o3[..., Cn:2 * Cn].float()).permute(0, 3, 1, 2)
        ok &= _report(f'groupnorm_apply_split {name}: hi + lo vs torch', rec, ref, 2e-5, 2e-5)
        ok &= _report(f'groupnorm_apply_split {name}: third block == hi', o3[..., 2 * Cn:], o3[..., :Cn], 0, 0)
        if raw:
            rr = r3[..., :Cn].float() + r3[..., Cn:2 * Cn].float()
            ok &= _report(f'groupnorm_apply_split {name}: raw copy', rr, xcat, 1e-5, 1e-6)

    # ---- softmax rows -> split
    S = _gen(64, 48, seed=9) * 4
    P = torch.zeros(64, 3 * 48, device=DEV, dtype=torch.bfloat16)
    K.softmax_rows_split(S, P, 64, 48, 0.37)
    ok &= _report('softmax_rows_split: hi + lo', P[:, :48].float() + P[:, 48:96].float(), torch.softmax(S * 0.37, dim=-1), 1e-5, 1e-7)
    ok &= _report('softmax_rows_split: third block == hi', P[:, 96:], P[:, :48], 0, 0)

    # ---- pack mode 3 + conv over split operands vs fp32 conv: 1e-5 level
    B, Cin, Cout, H = 4, 128, 256, 16
    x = _gen(B, Cin, H, H, seed=10)
    w = _gen(Cout, Cin, 3, 3, seed=11, scale=0.05)
    bias = _gen(Cout, seed=12)
    ref = F.conv2d(x.double(), w.double(), bias.double(), padding=1).float()
    wp = torch.zeros(Cout, 27 * Cin, device=DEV, dtype=torch.bfloat16)
    tab = torch.frombuffer(bytearray(K.pack_entry_bytes(w, wp, Cout, Cin, 9, 3, ld=27 * Cin)), dtype=torch.uint8).to(DEV)
    K.pack_weights(tab, 1)
    hi, lo = split_ref(w.permute(0, 2, 3, 1).reshape(Cout, 9, Cin))
    ok &= _report('pack mode 3', wp, torch.cat([hi, hi, lo], dim=-1).reshape(Cout, 27 * Cin), 0, 0)
    a3 = torch.zeros(B, H, H, 3 * Cin, device=DEV, dtype=torch.bfloat16)
    K.split_cast(x.permute(0, 2, 3, 1).contiguous(), a3, B * H * H, Cin)
    out = torch.zeros(B, H, H, Cout, device=DEV)
    K.conv2d(a3, wp, Cout, B, H, H, K.taps_3x3_s1(), a0_geom=(3 * Cin, H, H, 1), bias=bias, out=out)
    torch.cuda.synchronize()
    rel = ((out.permute(0, 3, 1, 2) - ref).norm() / ref.norm()).item()
    good = rel <= 2e-5
    print(json.dumps({'case': 'conv3x3 128->256 @16 over split operands vs fp64 conv', 'rel_l2': rel, 'gate': 2e-5, 'ok': good}))
    ok &= good
    # same layer with plain bf16 operands, for scale
    wb = K.pack_weight(w)
    outb = torch.zeros(B, H, H, Cout, device=DEV)
    K.conv2d(_nhwc_bf16(x), wb, Cout, B, H, H, K.taps_3x3_s1(), a0_geom=(Cin, H, H, 1), bias=bias, out=outb)
    torch.cuda.synchronize()
    print(json.dumps({'case': 'context: the same conv with single bf16 operands', 'rel_l2': ((outb.permute(0, 3, 1, 2) - ref).norm() / ref.norm()).item(), 'ok': True}))
    return ok


def case_ddim_inversion():
    """DDIM.denoise_inversion / DDIMCFG's guided inversion step (K4 with the inversion coefficient row) against the 48
    single-step fixtures frozen from the reference (diffusions/ddim.py:88-110; oracle/gen_golden_inversion.py) and,
    for the guided step, against the oracle's op sequence (ddim.py:202-232)."""
    sys.path.insert(0, ROOT)
    from oracle import diffusion_ref as R
    import diffusions
    g = torch.load(os.path.join(ROOT, 'tests', 'golden', 'ddim_inversion.pt'), weights_only=False)
    xt, mo, mo_u = g['xt'].to(DEV), g['mo'].to(DEV), g['mo_u'].to(DEV)
    worst = worst_cfg = 0.0
    for c in g['steps']:
        kw = dict(total_steps=1000, beta_schedule=c['beta'], objective=c['objective'], clip_denoised=c['clip'],
                  respace_type='uniform', respace_steps=50)
        o = diffusions.DDIM(device=DEV, **kw).denoise_inversion(mo.clone(), xt, c['t'], c['t_next'])
        for k, v in c['out'].items():
            worst = max(worst, (o[k].cpu() - v).abs().max().item() / max(1.0, v.abs().max().item()))
        # guided step: ours fuses predict(cond), predict(uncond), mix, predict(mix), step; the oracle walks them one by one
        oc = diffusions.DDIMCFG(guidance_scale=2.5, device=DEV, **kw)
        got = oc._inversion_impl(mo.clone(), xt, c['t'], c['t_next'], mo_u.clone(), 2.5)
        ref = R.DDIMRef(**kw)
        eps_c = ref.predict(g['mo'], g['xt'], c['t'])['pred_eps']
        eps_u = ref.predict(g['mo_u'], g['xt'], c['t'])['pred_eps']
        ref.objective = 'pred_eps'
        want = ref.denoise_inversion((1 - 2.5) * eps_u + 2.5 * eps_c, g['xt'], c['t'], c['t_next'])
        for k in ('sample', 'pred_x0', 'pred_eps'):
            worst_cfg = max(worst_cfg, (got[k].cpu() - want[k]).abs().max().item() / max(1.0, want[k].abs().max().item()))
    raised = False
    try:
        diffusions.DDIM(eta=0.5, device=DEV).denoise_inversion(mo, xt, 0, 20)
    except ValueError:
        raised = True
    good = worst <= 2e-6 and worst_cfg <= 1e-5 and raised
    print(json.dumps({'case': f'DDIM inversion single steps ({len(g["steps"])} cases) vs reference golden',
                      'max_rel_err': worst, 'gate': 2e-6, 'cfg_max_rel_err': worst_cfg, 'cfg_gate': 1e-5,
                      'eta_nonzero_raises': raised, 'ok': good}), flush=True)
    return good


# ----------------------------------------------------------------------------------------------------
# sampler step vs the eager op sequence of the reference (restated in oracle/diffusion_ref.py)
# ----------------------------------------------------------------------------------------------------
def case_sampler():
    sys.path.insert(0, ROOT)
    from oracle import diffusion_ref as R
    import diffusions
    ok = True
    B, C = 4, 3
    # H=8: HW % 4 == 0 -> the 16-byte streaming kernel; H=5: HW = 25 -> the generic one-element kernel
    for H in (8, 5):
        xt = _gen(B, C, H, H, seed=1)
        noise = _gen(B, C, H, H, seed=2)
        for kind in ('ddpm', 'ddim'):
            for var_type in (('fixed_small', 'fixed_large', 'learned_range') if kind == 'ddpm' else ('fixed_large',)):
                for eta in ((0.0, 0.5, 1.0) if kind == 'ddim' else (0.0,)):
                    for objective in ('pred_eps', 'pred_x0', 'pred_v'):
                        for clip in (True, False):
                            Cm = 2 * C if var_type == 'learned_range' else C
                            mo = _gen(B, Cm, H, H, seed=3)
                            kw = dict(total_steps=1000, objective=objective, clip_denoised=clip, respace_type='uniform',
                                      respace_steps=50)
                            if kind == 'ddpm':
                                ours = diffusions.DDPM(var_type=var_type, device=DEV, **kw)
                                ref = R.DDPMRef(var_type=var_type, **kw)
                            else:
                                ours = diffusions.DDIM(eta=eta, device=DEV, **kw)
                                ref = R.DDIMRef(eta=eta, **kw)
                            for (t, tp) in ((980, 960), (500, 480), (20, 0), (0, -1)):
                                o = ours.denoise(mo.clone(), xt, t, tp, reverse_eps=noise)
                                r = ref.denoise(mo.clone().cpu(), xt.cpu(), t, tp, reverse_eps=noise.cpu())
                                for key in ('sample', 'mean', 'pred_x0', 'pred_eps', 'var'):
                                    rv = r[key] if torch.is_tensor(r[key]) else torch.tensor(r[key])
                                    ov = o[key]
                                    # same fp32 op order, non-contracted: agreement to 1e-6 abs + 1e-6 rel
                                    # (learned_range goes through expf: 1e-5)
                                    tol = 1e-5 if var_type == 'learned_range' else 1e-6
                                    good = torch.allclose(ov.cpu().float().expand_as(rv) if ov.dim() == 0 else ov.cpu(),
                                                          rv, rtol=tol, atol=tol)
                                    if not good:
                                        _report(f'sampler {kind} {var_type} eta={eta} {objective} clip={clip} t={t} {key}',
                                                ov.cpu().expand_as(rv) if ov.dim() == 0 else ov.cpu(), rv, tol, tol)
                                        ok = False
    print(json.dumps({'case': 'sampler sweep', 'ok': ok}))
    return ok



def case_sampler_cfg():
    """Classifier-free-guidance step: per-branch predict (+clip), eps = (1-s) eps_u + s eps_c, then the step under
    objective pred_eps (diffusions/ddim.py:161-191, ddpm.py:319-351), fused in K4, against the oracle's op sequence."""
    sys.path.insert(0, ROOT)
    from oracle import diffusion_ref as R
    import diffusions
    ok, worst = True, 0.0
    B, C = 4, 3
    for H in (8, 5):
        xt, noise = _gen(B, C, H, H, seed=1), _gen(B, C, H, H, seed=2)
        for kind, var_type, eta in (('ddim', 'fixed_large', 0.0), ('ddim', 'fixed_large', 1.0),
                                    ('ddpm', 'fixed_small', 0.0), ('ddpm', 'learned_range', 0.0)):
            for objective in ('pred_eps', 'pred_x0', 'pred_v'):
                for clip in (True, False):
                    Cm = 2 * C if var_type == 'learned_range' else C
                    mo_c, mo_u = _gen(B, Cm, H, H, seed=3), _gen(B, Cm, H, H, seed=4)
                    kw = dict(total_steps=1000, objective=objective, clip_denoised=clip, respace_type='uniform',
                              respace_steps=50, beta_schedule='cosine')
                    rkw = dict(kw)
                    rkw['beta_schedule_kind'] = rkw.pop('beta_schedule')
                    if kind == 'ddpm':
                        ours = diffusions.DDPMCFG(guidance_scale=3.0, var_type=var_type, device=DEV, **kw)
                        ref = R.DDPMRef(var_type=var_type, **rkw)
                    else:
                        ours = diffusions.DDIMCFG(guidance_scale=3.0, eta=eta, device=DEV, **kw)
                        ref = R.DDIMRef(eta=eta, **rkw)
                    for (t, tp) in ((980, 960), (500, 480), (0, -1)):
                        o = ours._denoise_impl(mo_c.clone(), xt, t, tp, noise, mo_u.clone(), 3.0)
                        ref._pairs = lambda t=t, tp=tp: [(t, tp)]   # one step (t -> tp) of the oracle's CFG loop
                        r = next(ref.sample_loop_cfg(lambda img, tb, which: (mo_c if which else mo_u).cpu(), xt.cpu(),
                                                     3.0, dict(which=True), dict(which=False), noises=[noise.cpu()]))
                        for key in ('sample', 'mean', 'pred_x0', 'pred_eps'):
                            tol = 1e-5   # same fp32 op order; the guided eps reaches O(100) at small alpha-bar
                            worst = max(worst, ((o[key].cpu() - r[key]).abs() / (1.0 + r[key].abs())).max().item())
                            good = torch.allclose(o[key].cpu(), r[key], rtol=tol, atol=tol)
                            if not good:
                                _report(f'sampler cfg {kind} {var_type} eta={eta} {objective} clip={clip} H={H} t={t} {key}',
                                        o[key].cpu(), r[key], tol, tol)
                                ok = False
    print(json.dumps({'case': 'sampler cfg sweep', 'max_rel_err': worst, 'gate': 1e-5, 'ok': ok}))
    return ok


def case_sampler_large():
    """K4 at a size where the two-units-per-thread streaming kernel and the grid-stride tail are exercised
    (B=24, 3x256x256: 4.7 M elements, 1.18 M units > 148*8*256 threads), against the generic kernel run on the same
    data through unaligned views (bit-exact: same per-element arithmetic), for the plain, noisy, learned-variance and
    CFG forms."""
    import diffusions
    import b200diff as K
    ok = True
    B, C, H = 24, 3, 256
    n = B * C * H * H
    for var_type, cfg in (('fixed_large', False), ('learned_range', False), ('fixed_large', True)):
        Cm = 2 * C if var_type == 'learned_range' else C
        d = diffusions.DDPM(total_steps=1000, var_type=var_type, respace_type='uniform', respace_steps=50, device=DEV)
        row = d._coef_row(500, 480)
        g = torch.Generator(device=DEV).manual_seed(5)
        def mk(c):
            # one spare leading float so that [1:] views are 4-byte- but not 16-byte-aligned
            buf = torch.randn(B * c * H * H + 4, device=DEV, generator=g)
            al = buf[4:].view(B, c, H, H)
            un = torch.empty(B * c * H * H + 4, device=DEV)[1:B * c * H * H + 1].view(B, c, H, H)
            un.copy_(al)
            return al, un
        mo, mo_un = mk(Cm)
        xt, xt_un = mk(C)
        nz, nz_un = mk(C)
        mu, mu_un = mk(Cm) if cfg else (None, None)
        outs = {k: torch.empty(B, C, H, H, device=DEV) for k in ('sample', 'mean', 'pred_x0', 'pred_eps', 'var_out')}
        outs_un = {k: torch.empty(B, C, H, H, device=DEV) for k in outs}
        common = dict(objective='pred_eps', clip=True, learned_range=var_type == 'learned_range',
                      guidance_scale=3.0 if cfg else 1.0)
        K.sampler_step(mo, xt, row, noise=nz, model_out_uncond=mu, **common, **outs)
        K.sampler_step(mo_un, xt_un, row, noise=nz_un, model_out_uncond=mu_un, **common, **outs_un)
        torch.cuda.synchronize()
        for key in outs:
            if key == 'var_out' and var_type != 'learned_range':
                continue
            same = torch.equal(outs[key], outs_un[key])
            finite = bool(torch.isfinite(outs[key]).all())
            if not (same and finite):
                print(json.dumps({'case': f'sampler large {var_type} cfg={cfg} {key}', 'bit_equal': same, 'finite': finite}))
                ok = False
    print(json.dumps({'case': f'sampler large ({n} elements) streaming vs generic kernel', 'ok': ok}))
    return ok



def case_pack_weights():
    """b200_pack_weights (multi-tensor fp32 OIHW -> bf16 GEMM operands) against torch permutes: forward layout (mode 0),
    flipped / channel-swapped data-gradient layout (mode 1), bias sums (mode 2); ragged channel counts, taps 1 / 4 / 9 /
    25 (generic gather), row / column offsets inside a wider destination."""
    import b200diff as K
    ok = True
    entries, checks = [], []
    keep = []
    for i, (Co, Ci, kh, kw) in enumerate(((256, 256, 3, 3), (128, 384, 3, 3), (3, 128, 3, 3), (128, 3, 3, 3),
                                          (256, 512, 1, 1), (70, 45, 2, 2), (33, 31, 3, 3), (8, 40, 5, 5))):
        taps = kh * kw
        w = _gen(Co, Ci, kh, kw, seed=10 + i)
        # mode 0 into a wider matrix at a row / column offset
        row0, col0, extra = 2 * (i % 3), 8 * (i % 2), 16
        ld = col0 + taps * Ci + extra
        d0 = torch.full((row0 + Co + 1, ld), 7.0, device=DEV, dtype=torch.bfloat16)
        entries.append(K.pack_entry_bytes(w, d0, Co, Ci, taps, 0, row0=row0, col0=col0, ld=ld))
        r0 = torch.full_like(d0, 7.0)
        r0[row0:row0 + Co, col0:col0 + taps * Ci] = w.permute(0, 2, 3, 1).reshape(Co, taps * Ci).to(torch.bfloat16)
        checks.append((f'pack mode 0 {Co}x{Ci}x{kh}x{kw}', d0, r0))
        # mode 1
        d1 = torch.full((Ci, taps * Co), 7.0, device=DEV, dtype=torch.bfloat16)
        entries.append(K.pack_entry_bytes(w, d1, Co, Ci, taps, 1, ld=taps * Co))
        r1 = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, taps * Co).to(torch.bfloat16)
        checks.append((f'pack mode 1 {Co}x{Ci}x{kh}x{kw}', d1, r1))
        keep += [w, d0, d1]
    b1, b2 = _gen(300, seed=30), _gen(300, seed=31)
    db, db1 = torch.zeros(300, device=DEV), torch.zeros(300, device=DEV)
    entries.append(K.pack_entry_bytes(b1, db, 300, 1, 1, 2, src2=b2))
    entries.append(K.pack_entry_bytes(b1, db1, 300, 1, 1, 2))
    checks += [('pack mode 2 (bias sum)', db, b1 + b2), ('pack mode 2 (single bias)', db1, b1)]
    blob = torch.frombuffer(bytearray(b''.join(entries)), dtype=torch.uint8).to(DEV)
    K.pack_weights(blob, len(entries))
    torch.cuda.synchronize()
    for name, got, ref in checks:
        ok &= _report(name, got, ref, 0, 0)
    return ok



def case_conv_gnfuse():
    """b200_conv2d_gn_fwd (conv1 -> norm2 of a ResBlock fused in the conv epilogue): SiLU(GN(conv3x3(x) + bias + time-embedding row)) as bf16 NHWC, with
    and without AdaGN scale / shift, at the three resolutions where a tile holds whole images."""
    import b200diff as K
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    for (B, Cin, Cout, H, rowadd, ss, silu) in ((160, 256, 256, 16, True, False, True), (160, 256, 256, 8, True, False, True),
                                                 (320, 256, 256, 4, True, False, True), (160, 128, 256, 8, False, True, True),
                                                 (128, 256, 128, 16, False, False, False),
                                                 # other tile sizes of the rolled epilogue: 256-pixel tiles of sixteen 4x4
                                                 # images (4 chunks, two images per 32-pixel chunk), 64-pixel tiles = one 8x8 image
                                                 (1024, 256, 256, 4, True, True, True), (80, 256, 256, 8, True, False, True),
                                                 # 32x32 images: 4 tiles per image, statistics shared by 4 co-scheduled CTAs
                                                 (150, 128, 128, 32, True, False, True), (37, 384, 128, 32, True, True, True),
                                                 (3, 128, 128, 32, False, False, True)):
        x = _bf16r(_gen(B, Cin, H, H, seed=1))
        if H == 32:   # statistics that differ between the four tiles of an image
            x = _bf16r(x * torch.linspace(0.5, 2.0, H, device=DEV)[None, None, :, None])
        w = _bf16r(_gen(Cout, Cin, 3, 3, seed=2, scale=1.0 / math.sqrt(Cin * 9)))
        b = _gen(Cout, seed=3)
        ra = _gen(B, Cout + 64, seed=4) if rowadd else None
        gamma, beta = 1.0 + 0.1 * _gen(Cout, seed=5), 0.1 * _gen(Cout, seed=6)
        sst = 0.2 * _gen(B, 2 * Cout, seed=7) if ss else None
        ref = F.conv2d(x, w, b, padding=1)
        if rowadd:
            ref = ref + ra[:, :Cout, None, None]
        ref = F.group_norm(ref, 32, gamma, beta, eps=1e-5)
        if ss:
            ref = ref * (1 + sst[:, :Cout, None, None]) + sst[:, Cout:, None, None]
        if silu:
            ref = F.silu(ref)
        out = torch.full((B, H, H, Cout), float('nan'), device=DEV, dtype=torch.bfloat16)
        ws = dict(xstats=K.new_stats(B, Cout, DEV), xcount=K.new_stats(B, 1, DEV)) if K.conv2d_gn_needs_workspace(H, H) else {}
        K.conv2d_gn(_nhwc_bf16(x), K.pack_weight(w), Cout, B, H, H, K.taps_3x3_s1(), a0_geom=(Cin, H, H, 1), gamma=gamma,
                    beta=beta, groups=32, eps=1e-5, out_norm=out, bias=b, rowadd=ra, rowadd_ld=(Cout + 64) if rowadd else 0,
                    scale=sst, shift=None if sst is None else sst[:, Cout:], ss_ld=2 * Cout, silu=silu, **ws)
        torch.cuda.synchronize()
        if ws:   # the workspace ends up holding the statistics of the (unnormalised) conv output
            pre = F.conv2d(x, w, b, padding=1) + (ra[:, :Cout, None, None] if rowadd else 0.0)
            want_st = torch.stack([pre.sum(dim=(2, 3)), (pre * pre).sum(dim=(2, 3))], dim=-1)
            ok &= _report(f'conv3x3 {Cin}->{Cout} @{H} B={B} fused GN: image statistics workspace', K.stats_to_float(ws['xstats']), want_st, 1e-3, 5e-2)
        ok &= _report(f'conv3x3 {Cin}->{Cout} @{H} B={B} + fused GN (rowadd={rowadd}, adagn={ss}, silu={silu})',
                      out.permute(0, 3, 1, 2), ref, rtol=2e-2, atol=2e-2)
    return ok


def case_conv_gnfuse_out():
    """b200_conv2d_gn_fwd, block-output form (conv2 -> the next block's norm1 / the output head's norm): x = conv(a) + bias
    (+ residual | + 1x1 shortcut K-blocks) as fp32 NHWC with per-channel statistics, and SiLU?(GN(x)) as bf16 NHWC, in
    one launch; 3x3 and 1x1 taps, all four resolutions (32x32: four co-scheduled CTAs per image)."""
    import b200diff as K
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    #        B   Cin  Cout  H  taps residual shortcut(Csc) groups silu
    cfgs = ((150, 128, 128, 32, 3, True, 0, 32, True),       # E0 -> E1 of the CIFAR-10 UNet
            (37, 128, 128, 32, 3, False, 256, 32, True),     # last decoder block (1x1 shortcut over 256 channels) -> head
            (96, 256, 256, 16, 3, True, 0, 32, True),
            (160, 256, 256, 8, 3, True, 0, 32, True),
            (320, 256, 256, 4, 3, True, 0, 32, False),       # ResBlock -> attention norm (no SiLU)
            (64, 256, 256, 4, 1, True, 0, 32, True),         # attention output projection (1x1 + residual) -> next norm1
            (128, 128, 256, 8, 3, False, 128, 16, True),
            (5, 128, 128, 32, 3, False, 0, 32, True))
    for (B, Cin, Cout, H, k, use_res, Csc, groups, silu) in cfgs:
        x = _bf16r(_gen(B, Cin, H, H, seed=1))
        if H == 32:
            x = _bf16r(x * torch.linspace(0.5, 2.0, H, device=DEV)[None, None, :, None])
        w = _bf16r(_gen(Cout, Cin, k, k, seed=2, scale=1.0 / math.sqrt(Cin * k * k)))
        b = _gen(Cout, seed=3)
        gamma, beta = 1.0 + 0.1 * _gen(Cout, seed=5), 0.1 * _gen(Cout, seed=6)
        res = _gen(B, H, H, Cout, seed=8) if use_res else None                   # fp32 NHWC residual stream
        pre = F.conv2d(x, w, b, padding=k // 2)
        wp = K.pack_weight(w)
        a1 = None
        if Csc:
            xs = _bf16r(_gen(B, Csc, H, H, seed=9))
            wsc = _bf16r(_gen(Cout, Csc, 1, 1, seed=10, scale=1.0 / math.sqrt(Csc)))
            pre = pre + F.conv2d(xs, wsc)
            wp = torch.cat([wp, K.pack_weight(wsc)], dim=1).contiguous()
            a1 = _nhwc_bf16(xs)
        if use_res:
            pre = pre + res.permute(0, 3, 1, 2)
        ref = F.group_norm(pre, groups, gamma, beta, eps=1e-5)
        if silu:
            ref = F.silu(ref)
        out = torch.full((B, H, H, Cout), float('nan'), device=DEV, dtype=torch.float32)
        outn = torch.full((B, H, H, Cout), float('nan'), device=DEV, dtype=torch.bfloat16)
        stats = K.new_stats(B, Cout, DEV)
        multi = K.conv2d_gn_needs_workspace(H, H)
        ws = dict(xstats=stats, xcount=K.new_stats(B, 1, DEV)) if multi else {}
        taps = K.taps_3x3_s1() if k == 3 else K.taps_1x1()
        K.conv2d_gn(_nhwc_bf16(x), wp, Cout, B, H, H, taps, a0_geom=(Cin, H, H, 1), gamma=gamma, beta=beta, groups=groups,
                    eps=1e-5, out_norm=outn, bias=b, silu=silu, out=out, stats=None if multi else stats, residual=res,
                    res_ld=Cout if use_res else 0, a1=a1, a1_geom=(Csc, H, H, 1) if Csc else None, **ws)
        torch.cuda.synchronize()
        name = f'conv{k}x{k} {Cin}->{Cout} @{H} B={B} (res={use_res}, shortcut={Csc}, groups={groups}, silu={silu})'
        ok &= _report(name + ' raw fp32 output', out.permute(0, 3, 1, 2), pre, rtol=1e-3, atol=1e-3)
        want_st = torch.stack([pre.sum(dim=(2, 3)), (pre * pre).sum(dim=(2, 3))], dim=-1)
        ok &= _report(name + ' output statistics', K.stats_to_float(stats), want_st, 1e-3, 5e-2)
        ok &= _report(name + ' fused GN of the output', outn.permute(0, 3, 1, 2), ref, rtol=2e-2, atol=2e-2)

    # consumer concatenates a skip connection: the conv writes its part of the normalised + raw concat operands
    # (channel stride Cout + Cs, groups of (Cout + Cs) / 32 channels), the streaming GroupNorm kernel in window mode
    # (x0 = None) adds the skip's channels; together they must equal GroupNorm(cat(x, skip))
    for (B, Cin, Cout, Cs, H) in ((64, 128, 128, 128, 32), (96, 256, 256, 256, 8), (128, 256, 256, 256, 4), (33, 256, 256, 256, 16)):
        x = _bf16r(_gen(B, Cin, H, H, seed=21))
        w = _bf16r(_gen(Cout, Cin, 3, 3, seed=22, scale=1.0 / math.sqrt(Cin * 9)))
        b = _gen(Cout, seed=23)
        xs = _bf16r(_gen(B, 384, H, H, seed=24))
        wsc = _bf16r(_gen(Cout, 384, 1, 1, seed=25, scale=1.0 / math.sqrt(384)))
        skip = _gen(B, H, H, Cs, seed=26) * 1.5 + 0.25                             # fp32 NHWC skip connection
        Ct = Cout + Cs
        gamma, beta = 1.0 + 0.1 * _gen(Ct, seed=27), 0.1 * _gen(Ct, seed=28)
        pre = F.conv2d(x, w, b, padding=1) + F.conv2d(xs, wsc)
        cat = torch.cat([pre, skip.permute(0, 3, 1, 2)], dim=1)
        ref = F.silu(F.group_norm(cat, 32, gamma, beta, eps=1e-5))
        wp = torch.cat([K.pack_weight(w), K.pack_weight(wsc)], dim=1).contiguous()
        out = torch.full((B, H, H, Cout), float('nan'), device=DEV, dtype=torch.float32)
        outn = torch.full((B, H, H, Ct), float('nan'), device=DEV, dtype=torch.bfloat16)
        outr = torch.full((B, H, H, Ct), float('nan'), device=DEV, dtype=torch.bfloat16)
        stats = K.new_stats(B, Cout, DEV)
        multi = K.conv2d_gn_needs_workspace(H, H)
        ws = dict(xstats=stats, xcount=K.new_stats(B, 1, DEV)) if multi else {}
        cpg = Ct // 32
        K.conv2d_gn(_nhwc_bf16(x), wp, Cout, B, H, H, K.taps_3x3_s1(), a0_geom=(Cin, H, H, 1), gamma=gamma, beta=beta,
                    groups=Cout // cpg, eps=1e-5, out_norm=outn, out_norm_ld=Ct, out_raw=outr, bias=b, silu=True, out=out,
                    stats=None if multi else stats, a1=_nhwc_bf16(xs), a1_geom=(384, H, H, 1), **ws)
        sst = K.stats_from_float(torch.stack([skip.sum(dim=(1, 2)), (skip * skip).sum(dim=(1, 2))], dim=-1).contiguous())
        K.groupnorm_apply(None, Cout, None, skip, Cs, sst, B, H * H, H, 32, gamma, beta, 1e-5, outn, silu=True, raw_out=outr)
        torch.cuda.synchronize()
        name = f'conv3x3 {Cin}->{Cout} @{H} B={B} -> GN(cat(x, skip {Cs}))'
        ok &= _report(name + ' raw fp32 output', out.permute(0, 3, 1, 2), pre, rtol=1e-3, atol=1e-3)
        ok &= _report(name + ' normalised concat operand', outn.permute(0, 3, 1, 2), ref, rtol=2e-2, atol=2e-2)
        ok &= _report(name + ' raw concat operand', outr.permute(0, 3, 1, 2), cat, rtol=1e-2, atol=1e-2)
        # the same launch without the fp32 output (nobody reads it when the consumer is a concatenating block):
        # the bf16 operands must be bitwise the same as above
        outn2 = torch.full((B, H, H, Ct), float('nan'), device=DEV, dtype=torch.bfloat16)
        outr2 = torch.full((B, H, H, Ct), float('nan'), device=DEV, dtype=torch.bfloat16)
        ws2 = dict(xstats=K.new_stats(B, Cout, DEV), xcount=K.new_stats(B, 1, DEV)) if multi else {}
        K.conv2d_gn(_nhwc_bf16(x), wp, Cout, B, H, H, K.taps_3x3_s1(), a0_geom=(Cin, H, H, 1), gamma=gamma, beta=beta,
                    groups=Cout // cpg, eps=1e-5, out_norm=outn2, out_norm_ld=Ct, out_raw=outr2, bias=b, silu=True,
                    a1=_nhwc_bf16(xs), a1_geom=(384, H, H, 1), **ws2)
        torch.cuda.synchronize()
        same = (torch.equal(outn2[..., :Cout].view(torch.int16), outn[..., :Cout].view(torch.int16))
                and torch.equal(outr2[..., :Cout].view(torch.int16), outr[..., :Cout].view(torch.int16)))
        print(json.dumps({'check': name + ' without the fp32 output: operands bitwise equal', 'ok': bool(same)}))
        ok &= bool(same)
    return ok



def case_first_conv_tc():
    """First convolution on the tensor cores: b200_first_split (fp32 NCHW image -> 64-channel bf16 [hi | lo | hi | 0..]
    pixels) + pack mode 4 weights [w_hi | w_hi | w_lo | 0..] + the implicit-GEMM conv = fp32-grade products (gate 1e-4
    relative + 3e-4 absolute at |out| <= 17, measured 1.1e-4: a bf16-rounded image gives 1e-2), plain and with the first ResBlock's GroupNorm fused."""
    import b200diff as K
    torch.backends.cudnn.allow_tf32 = False
    ok = True
    for (B, Cin, Cout, H, W) in ((37, 3, 128, 32, 32), (64, 3, 256, 16, 16), (5, 1, 128, 32, 32), (3, 4, 128, 64, 64)):
        x = _gen(B, Cin, H, W, seed=1) * 1.7
        w = _gen(Cout, Cin, 3, 3, seed=2, scale=0.3)
        b = _gen(Cout, seed=3)
        ref = F.conv2d(x, w, b, padding=1)
        a = torch.full((B, H, W, 64), float('nan'), device=DEV, dtype=torch.bfloat16)
        K.first_split(x, a)
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        want_a = torch.zeros((B, 64, H, W), device=DEV, dtype=torch.bfloat16)
        want_a[:, :Cin], want_a[:, Cin:2 * Cin], want_a[:, 2 * Cin:3 * Cin] = hi, lo, hi
        ok &= _report(f'first_split {Cin} ch @{H}x{W}', a.permute(0, 3, 1, 2), want_a, 0, 0)
        wp = torch.zeros((Cout, 9 * 64), device=DEV, dtype=torch.bfloat16)
        tab = torch.frombuffer(bytearray(K.pack_entry_bytes(w, wp, Cout, Cin, 9, 4, ld=9 * 64, tap_ld=64)), dtype=torch.uint8).to(DEV)
        K.pack_weights(tab, 1)
        whi = w.to(torch.bfloat16)
        wlo = (w - whi.float()).to(torch.bfloat16)
        want_w = torch.zeros((Cout, 9, 64), device=DEV, dtype=torch.bfloat16)
        wt = lambda t: t.reshape(Cout, Cin, 9).permute(0, 2, 1)
        want_w[:, :, :Cin], want_w[:, :, Cin:2 * Cin], want_w[:, :, 2 * Cin:3 * Cin] = wt(whi), wt(whi), wt(wlo)
        ok &= _report(f'pack mode 4 {Cin}->{Cout}', wp.view(Cout, 9, 64), want_w, 0, 0)
        out = torch.full((B, H, W, Cout), float('nan'), device=DEV)
        st = K.new_stats(B, Cout, DEV)
        K.conv2d(a, wp, Cout, B, H, W, K.taps_3x3_s1(), a0_geom=(64, H, W, 1), bias=b, out=out, out_mode=K.OUT_F32_NHWC, stats=st)
        torch.cuda.synchronize()
        ok &= _report(f'first conv (tensor cores) {Cin}->{Cout} @{H}x{W}', out.permute(0, 3, 1, 2), ref, 1e-4, 3e-4)
        ok &= _report(f'first conv (tensor cores) {Cin}->{Cout} @{H}x{W} [statistics]', K.stats_to_float(st).float(),
                      torch.stack([ref.sum(dim=(2, 3)), (ref * ref).sum(dim=(2, 3))], dim=-1), 2e-4, 2e-2)
        if K.conv2d_gn_ok(B, H, W, Cout, 32):
            gamma, beta = 1.0 + 0.1 * _gen(Cout, seed=5), 0.1 * _gen(Cout, seed=6)
            out2 = torch.full((B, H, W, Cout), float('nan'), device=DEV)
            outn = torch.full((B, H, W, Cout), float('nan'), device=DEV, dtype=torch.bfloat16)
            st2 = K.new_stats(B, Cout, DEV)
            multi = K.conv2d_gn_needs_workspace(H, W)
            ws = dict(xstats=st2, xcount=K.new_stats(B, 1, DEV)) if multi else {}
            K.conv2d_gn(a, wp, Cout, B, H, W, K.taps_3x3_s1(), a0_geom=(64, H, W, 1), gamma=gamma, beta=beta, groups=32,
                        eps=1e-5, out_norm=outn, bias=b, silu=True, out=out2, stats=None if multi else st2, **ws)
            torch.cuda.synchronize()
            ok &= _report(f'first conv (tensor cores) {Cin}->{Cout} @{H}x{W} + fused GN: fp32 output', out2.permute(0, 3, 1, 2), ref, 1e-4, 3e-4)
            ok &= _report(f'first conv (tensor cores) {Cin}->{Cout} @{H}x{W} + fused GN: normalised operand', outn.permute(0, 3, 1, 2),
                          F.silu(F.group_norm(ref, 32, gamma, beta, eps=1e-5)), 1e-2, 1e-2)
    return ok


CASES = {n[5:]: f for n, f in list(globals().items()) if n.startswith('case_')}

if __name__ == '__main__':
    names = sys.argv[1:] or list(CASES)
    all_ok = True
    for n in names:
        try:
            okc = CASES[n]()
        except Exception as e:  # noqa: BLE001
            print(json.dumps({'case': n, 'ok': False, 'exception': repr(e)}))
            okc = False
        print(f'=== {n}: {"PASS" if okc else "FAIL"}', flush=True)
        all_ok &= bool(okc)
    sys.exit(0 if all_ok else 1)
