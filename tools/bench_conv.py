"""Micro-benchmark of b200_conv2d_fwd on the CIFAR-10 UNet layer shapes (B=256): CUDA-event timing, TFLOP/s."""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
import b200diff as K  # noqa: E402

DEV = 'cuda'
B = int(os.environ.get('B', 256))


def run(name, Cin, Cout, H, mode='3x3', residual=False, rowadd=False, stats=True, sc=0, out_mode=K.OUT_F32_NHWC,
        iters=20):
    W = H
    if mode == 'up2':
        a0 = torch.randn(B, H, W, Cin, device=DEV).to(torch.bfloat16)
        w = K.pack_weight_up2(torch.randn(Cout, Cin, 3, 3, device=DEV) / math.sqrt(9 * Cin))
        taps, Ho, Wo, geom, rpp = K.taps_up2_3x3(), H, W, (Cin, H, W, 1), Cout
        oh = 2 * H
        macs = 4.0 * B * H * W * Cout * 4 * Cin
    elif mode == 's2':
        a0 = torch.randn(B, 4, H // 2, W // 2, Cin, device=DEV).to(torch.bfloat16)
        w = K.pack_weight(torch.randn(Cout, Cin, 3, 3, device=DEV) / math.sqrt(9 * Cin))
        taps, Ho, Wo, geom, rpp = K.taps_3x3_s2(1), H // 2, W // 2, (Cin, H // 2, W // 2, 4), None
        oh = H // 2
        macs = 1.0 * B * Ho * Wo * Cout * 9 * Cin
    else:
        k = 3 if mode == '3x3' else 1
        a0 = torch.randn(B, H, W, Cin, device=DEV).to(torch.bfloat16)
        wsc = torch.randn(Cout, sc, 1, 1, device=DEV) if sc else None
        w = K.pack_weight(torch.randn(Cout, Cin, k, k, device=DEV) / math.sqrt(k * k * Cin), wsc)
        taps, Ho, Wo, geom, rpp = (K.taps_3x3_s1() if k == 3 else K.taps_1x1()), H, W, (Cin, H, W, 1), None
        oh = H
        macs = 1.0 * B * H * W * Cout * (k * k * Cin + sc)
    a1 = torch.randn(B, H, W, sc, device=DEV).to(torch.bfloat16) if sc else None
    if out_mode == K.OUT_F32_NHWC:
        out = torch.empty(B, oh, oh, Cout, device=DEV)
    elif out_mode == K.OUT_BF16_NHWC:
        out = torch.empty(B, oh, oh, Cout, device=DEV, dtype=torch.bfloat16)
    else:
        out = torch.empty(B, Cout, oh, oh, device=DEV, dtype=torch.bfloat16)
    res = torch.randn(B, oh, oh, Cout, device=DEV) if residual else None
    ra = torch.randn(B, Cout, device=DEV) if rowadd else None
    st = K.new_stats(B, Cout, DEV) if (stats and out_mode == K.OUT_F32_NHWC) else None
    bias = torch.randn(Cout, device=DEV)

    def call():
        K.conv2d(a0, w, Cout, B, Ho, Wo, taps, a0_geom=geom, a1=a1, a1_geom=(sc, H, W, 1) if sc else None, bias=bias,
                 rowadd=ra, rowadd_ld=Cout if rowadd else 0, residual=res, res_ld=Cout, out=out, out_mode=out_mode,
                 w_rows_per_phase=rpp, stats=st)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    print(f'{name:52s} {us:8.1f} us  {2 * macs / us / 1e6:7.1f} TFLOP/s (executed MACs)', flush=True)


if __name__ == '__main__':
    run('3x3 128->128 @32 plain', 128, 128, 32)
    run('3x3 128->128 @32 plain no-stats', 128, 128, 32, stats=False)
    run('3x3 128->128 @32 +temb', 128, 128, 32, rowadd=True)
    run('3x3 128->128 @32 +residual', 128, 128, 32, residual=True)
    run('3x3 256->128 @32', 256, 128, 32, rowadd=True)
    run('3x3 128->128 @32 + 1x1 shortcut 256', 128, 128, 32, sc=256)
    run('3x3 256->256 @16 +temb', 256, 256, 16, rowadd=True)
    run('3x3 256->256 @16 +residual', 256, 256, 16, residual=True)
    run('3x3 512->256 @16', 512, 256, 16, rowadd=True)
    run('3x3 256->256 @8', 256, 256, 8, residual=True)
    run('3x3 512->256 @8', 512, 256, 8, rowadd=True)
    run('3x3 256->256 @4', 256, 256, 4, residual=True)
    run('3x3 512->256 @4', 512, 256, 4, rowadd=True)
    run('1x1 256->512 @16 bf16 (qk)', 256, 512, 16, mode='1x1', out_mode=K.OUT_BF16_NHWC)
    run('1x1 256->256 @16 bf16 NCHW (v^T)', 256, 256, 16, mode='1x1', out_mode=K.OUT_BF16_NCHW)
    run('1x1 256->256 @16 +residual (proj)', 256, 256, 16, mode='1x1', residual=True)
    run('up2 256->256 16->32', 256, 256, 16, mode='up2')
    run('up2 256->256 16->32 no-stats', 256, 256, 16, mode='up2', stats=False)
    run('up2 256->256 8->16', 256, 256, 8, mode='up2')
    run('s2 128->128 32->16', 128, 128, 32, mode='s2')


def run_custom(name, Cin, Cout, H, taps, out_hw, up, w_rows, rpp, out_dtype=torch.float32, iters=20, pad_k=0):
    a0 = torch.randn(B, H, H, Cin, device=DEV).to(torch.bfloat16)
    Kdim = len(taps[0]) * Cin
    w = (torch.randn(w_rows, Kdim + pad_k, device=DEV) / math.sqrt(Kdim)).to(torch.bfloat16)
    out = torch.empty(B, out_hw, out_hw, Cout, device=DEV, dtype=out_dtype)
    bias = torch.randn(Cout, device=DEV)
    d = K.ConvDesc()
    d.a0 = a0.data_ptr(); d.a0_C, d.a0_H, d.a0_W, d.a0_planes = Cin, H, H, 1
    d.w = w.data_ptr(); d.w_rows, d.w_K, d.w_rows_per_phase = w_rows, Kdim, rpp
    d.B, d.Ho, d.Wo = B, H, H
    d.phases, d.N, d.ntaps0 = len(taps), Cout, len(taps[0])
    K._fill_taps(d, taps, (0, 0, 0))
    d.bias = bias.data_ptr()
    d.out = out.data_ptr(); d.out_mode = K.OUT_F32_NHWC if out_dtype == torch.float32 else K.OUT_BF16_NHWC
    d.out_ld = Cout; d.out_H = d.out_W = out_hw; d.osy = d.osx = up
    import ctypes

    def call():
        K._check(K.lib().b200_conv2d_fwd(ctypes.byref(d), K._stream()), 'conv')
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        call()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / iters * 1e3
    macs = len(taps) * B * H * H * Cout * Kdim
    print(f'{name:52s} {us:8.1f} us  {2 * macs / us / 1e6:7.1f} TFLOP/s', flush=True)


if __name__ == '__main__' and os.environ.get('UP2DIAG'):
    up = K.taps_up2_3x3()
    run_custom('up2 4 phases f32 (as is)', 256, 256, 16, up, 32, 2, 1024, 256)
    run_custom('up2 4 phases bf16 out', 256, 256, 16, up, 32, 2, 1024, 256, out_dtype=torch.bfloat16)
    run_custom('2x2-tap conv, 1 phase, out 16x16 stride 1', 256, 256, 16, up[:1], 16, 1, 256, 256)
    run_custom('2x2-tap conv, 1 phase, out 32x32 stride 2', 256, 256, 16, up[:1], 32, 2, 256, 256)
    run_custom('4 phases all writing phase-0 taps (same taps)', 256, 256, 16, [up[0]] * 4, 32, 2, 1024, 256)
    run_custom('3x3 conv K=2304 for reference', 256, 256, 16, K.taps_3x3_s1(), 16, 1, 256, 256)
    run_custom('1x1 conv K=256', 256, 256, 16, K.taps_1x1(), 16, 1, 256, 256)
    run_custom('1x1 conv K=256 bf16 out', 256, 256, 16, K.taps_1x1(), 16, 1, 256, 256, out_dtype=torch.bfloat16)
