"""pesser/pytorch_diffusion UNet (the DDPM authors' CelebA-HQ / LSUN checkpoints) on the B200 kernels: drop-in for
the reference's models/pesser/model.py:190-327 (same keyword-only constructor, state_dict keys, `model(x, t)`).

Differences from models/unet.py that the engine reproduces: GroupNorm eps 1e-6; the stride-2
downsample pads (0,1,0,1) instead of symmetrically (tap table `taps_3x3_s2(pad_lo=0)`); attention scales the
logits after the q.k product by C^-1/2 (identical inside the fused softmax); the attention flag is per
resolution, with one AttnBlock per ResnetBlock; single-head attention over all C (up to 512) channels.
"""
import torch
import torch.nn as nn

import b200diff as K
from models.engine import Act
from models.modules import SinusoidalPosEmb, _KernelOnly
from models.unet import _EngineModel


def Normalize(in_channels: int) -> nn.GroupNorm:
    return nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(_KernelOnly):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)


class Downsample(_KernelOnly):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=2, padding=0)


class ResnetBlock(_KernelOnly):
    def __init__(self, *, in_channels, out_channels=None, conv_shortcut=False, dropout, temb_channels=512):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels, self.use_conv_shortcut = in_channels, out_channels, conv_shortcut
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.temb_proj = nn.Linear(temb_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            if conv_shortcut:
                self.conv_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
            else:
                self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def shortcut_conv(self):
        if self.in_channels == self.out_channels:
            return None
        return self.conv_shortcut if self.use_conv_shortcut else self.nin_shortcut


class AttnBlock(_KernelOnly):
    def __init__(self, in_channels):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1, stride=1, padding=0)

    def packed_weights(self):
        wqk = torch.cat([K.pack_weight(self.q.weight), K.pack_weight(self.k.weight)], dim=0).contiguous()
        bqk = torch.cat([self.q.bias.detach(), self.k.bias.detach()]).float().contiguous()
        return (wqk, bqk, K.pack_weight(self.v.weight), self.v.bias.detach().float().contiguous(),
                K.pack_weight(self.proj_out.weight), self.proj_out.bias.detach().float().contiguous())


class Model(_EngineModel):
    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution):
        super().__init__()
        self.ch, self.temb_ch = ch, ch * 4
        self.num_resolutions, self.num_res_blocks = len(ch_mult), num_res_blocks
        self.resolution, self.in_channels, self.out_channels = resolution, in_channels, out_ch

        self.temb = nn.Module()
        self.temb.dense = nn.ModuleList([nn.Linear(ch, self.temb_ch), nn.Linear(self.temb_ch, self.temb_ch)])
        self.conv_in = nn.Conv2d(in_channels, ch, kernel_size=3, stride=1, padding=1)

        def res(cin, cout):
            return ResnetBlock(in_channels=cin, out_channels=cout, temb_channels=self.temb_ch, dropout=dropout)

        curr_res = resolution
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            level = nn.Module()
            blocks, attns = nn.ModuleList(), nn.ModuleList()
            block_in, block_out = ch * in_ch_mult[i_level], ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                blocks.append(res(block_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attns.append(AttnBlock(block_in))
            level.block, level.attn = blocks, attns
            if i_level != self.num_resolutions - 1:
                level.downsample = Downsample(block_in, resamp_with_conv)
                curr_res //= 2
            self.down.append(level)

        self.mid = nn.Module()
        self.mid.block_1 = res(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = res(block_in, block_in)

        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            level = nn.Module()
            blocks, attns = nn.ModuleList(), nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            skip_in = ch * ch_mult[i_level]
            for i_block in range(num_res_blocks + 1):
                if i_block == num_res_blocks:
                    skip_in = ch * in_ch_mult[i_level]
                blocks.append(res(block_in + skip_in, block_out))
                block_in = block_out
                if curr_res in attn_resolutions:
                    attns.append(AttnBlock(block_in))
            level.block, level.attn = blocks, attns
            if i_level != 0:
                level.upsample = Upsample(block_in, resamp_with_conv)
                curr_res *= 2
            self.up.insert(0, level)   # registered in ascending-resolution order, like the reference

        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)
        self.__dict__['_tfreq'] = SinusoidalPosEmb(ch)
        self._init_engine()

    def _res_blocks(self):
        blocks = []
        for i, lvl in enumerate(self.down):
            blocks += [(f'down.{i}.block.{j}', b) for j, b in enumerate(lvl.block)]
        blocks += [('mid.block_1', self.mid.block_1), ('mid.block_2', self.mid.block_2)]
        for i, lvl in enumerate(self.up):
            blocks += [(f'up.{i}.block.{j}', b) for j, b in enumerate(lvl.block)]
        return blocks

    def _forward_impl(self, x, t, out=None):
        """x: [B, C, R, R] fp32, t: [B] int64 -> [B, out_ch, R, R] fp32 (reference :286-327)."""
        assert x.shape[2] == x.shape[3] == self.resolution
        eng = self.engine
        eng.begin_forward()
        x = eng.check_input(x, t, self.in_channels)
        B, _, H, W = x.shape

        res_blocks = self._res_blocks()
        offsets, off = {}, 0
        for name, blk in res_blocks:
            offsets[name] = off
            off += blk.temb_proj.out_features
        # dense0 -> swish -> dense1, then swish again inside every block before its projection (:120)
        tproj, tld = eng.embed(t, None, B, self.__dict__['_tfreq'], self.temb.dense[0], self.temb.dense[1], None,
                               [blk.temb_proj for _, blk in res_blocks])

        def run_res(name, blk, h, skip=None):
            return eng.resblock_core(name, h, skip, norm1=blk.norm1, conv1=blk.conv1, norm2=blk.norm2, conv2=blk.conv2,
                                     shortcut=blk.shortcut_conv(), emb=tproj, emb_off=offsets[name], emb_ld=tld,
                                     scale_shift=False, dropout=blk.dropout, emb_linear=blk.temb_proj)

        def run_attn(name, blk, h):
            return eng.attention_core(name, h, blk.norm, eng.packed(('attn', name), blk.packed_weights), 1,
                                      float(int(blk.in_channels) ** (-0.5)), mods=(blk.q, blk.k, blk.v, blk.proj_out))

        hs = [eng.first_conv('conv_in', self.conv_in, x)]
        for i, lvl in enumerate(self.down):
            for j in range(self.num_res_blocks):
                h = run_res(f'down.{i}.block.{j}', lvl.block[j], hs[-1])
                if len(lvl.attn) > 0:
                    h = run_attn(f'down.{i}.attn.{j}', lvl.attn[j], h)
                hs.append(h)
            if i != self.num_resolutions - 1:
                ds = lvl.downsample
                hs.append(eng.downsample_conv(f'down.{i}.downsample.conv', ds.conv, hs[-1], pad_lo=0)
                          if ds.with_conv else eng.resample_plain(f'down.{i}.downsample', hs[-1], 1))

        eng.pingpong = True      # from here on no output is a skip connection: block outputs alternate between two buffers
        h = run_res('mid.block_1', self.mid.block_1, hs[-1])
        h = run_attn('mid.attn_1', self.mid.attn_1, h)
        h = run_res('mid.block_2', self.mid.block_2, h)

        for i in reversed(range(self.num_resolutions)):
            lvl = self.up[i]
            for j in range(self.num_res_blocks + 1):
                h = run_res(f'up.{i}.block.{j}', lvl.block[j], h, hs.pop())
                if len(lvl.attn) > 0:
                    h = run_attn(f'up.{i}.attn.{j}', lvl.attn[j], h)
            if i != 0:
                us = lvl.upsample
                # the next block concatenates a skip connection and therefore has a projection shortcut: only its GroupNorm
                # (+ the shortcut's bf16 operand) reads the up-sampled tensor, which may then be stored as bf16
                nxt = self.up[i - 1].block[0]
                h = eng.upsample_conv(f'up.{i}.upsample.conv', us.conv, h, bf16_out=nxt.shortcut_conv() is not None) \
                    if us.with_conv else eng.resample_plain(f'up.{i}.upsample', h, 2)

        return eng.head('norm_out', h, self.norm_out, self.conv_out, out)
