"""ADM / guided-diffusion UNet on the B200 kernels: drop-in for the reference's models/adm/unet.py:415-682
(same constructor arguments, state_dict keys and registration order, `model(x, timesteps, y)` call).

Like the other families the modules below are parameter containers; models/engine.py executes the forward as
libb200diff kernel launches:
  * ResBlock (reference :162-275): GN+SiLU (K3, with the 2x resample of up/down blocks fused) -> conv3x3 (K1)
    -> GN * (1 + scale) + shift + SiLU (K3, scale/shift rows from one batched projection GEMM) -> conv3x3 with the
    residual / 1x1 skip connection fused (K1);
  * AttentionBlock (:278-324) with QKVAttentionLegacy / QKVAttention (:347-412): the Conv1d qkv weight is split
    once, at pack time, into a [q heads | k heads] matrix and a v matrix (the two head/qkv interleavings differ only
    by a row permutation); the d^-1/4 scaling of q and k becomes the d^-1/2 factor of the fused softmax (K2).
Dropout is the identity in eval mode; `use_checkpoint` only changes autograd behaviour and is accepted and ignored.
"""
import math

import torch
import torch.nn as nn

import b200diff as K
from models.engine import Act
from models.modules import _KernelOnly
from models.unet import _EngineModel
from .nn import TimestepFrequencies, conv_nd, linear, normalization, zero_module


class TimestepBlock(_KernelOnly):
    """Marker base: a block that consumes the timestep embedding (reference :71-80)."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, *args, **kwargs):
        return _KernelOnly.forward(self, *args, **kwargs)


class Upsample(_KernelOnly):
    """nearest 2x (+ 3x3 conv under key `conv`), reference :99-127."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=1)


class Downsample(_KernelOnly):
    """3x3 stride-2 conv under key `op`, or a parameter-free 2x2 average pool, reference :130-159."""

    def __init__(self, channels, use_conv, dims=2, out_channels=None):
        super().__init__()
        self.channels, self.out_channels, self.use_conv, self.dims = channels, out_channels or channels, use_conv, dims
        if use_conv:
            self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=2, padding=1)
        else:
            assert self.channels == self.out_channels
            self.op = nn.AvgPool2d(kernel_size=2, stride=2)


class ResBlock(TimestepBlock):
    def __init__(self, channels, emb_channels, dropout, out_channels=None, use_conv=False,
                 use_scale_shift_norm=False, dims=2, use_checkpoint=False, up=False, down=False):
        super().__init__()
        self.channels, self.emb_channels, self.dropout = channels, emb_channels, dropout
        self.out_channels = out_channels or channels
        self.use_conv, self.use_checkpoint, self.use_scale_shift_norm = use_conv, use_checkpoint, use_scale_shift_norm
        self.in_layers = nn.Sequential(
            normalization(channels), nn.SiLU(), conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.updown = up or down
        self.resample = 2 if up else 1 if down else 0       # engine code: nearest 2x / 2x2 average pool / none
        if up:
            self.h_upd, self.x_upd = Upsample(channels, False, dims), Upsample(channels, False, dims)
        elif down:
            self.h_upd, self.x_upd = Downsample(channels, False, dims), Downsample(channels, False, dims)
        else:
            self.h_upd = self.x_upd = nn.Identity()
        self.emb_layers = nn.Sequential(
            nn.SiLU(), linear(emb_channels, 2 * self.out_channels if use_scale_shift_norm else self.out_channels))
        self.out_layers = nn.Sequential(
            normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)))
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        elif use_conv:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 3, padding=1)
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)


class AttentionBlock(_KernelOnly):
    def __init__(self, channels, num_heads=1, num_head_channels=-1, use_checkpoint=False,
                 use_new_attention_order=False):
        super().__init__()
        self.channels = channels
        if num_head_channels == -1:
            self.num_heads = num_heads
        else:
            assert channels % num_head_channels == 0, \
                f'q,k,v channels {channels} is not divisible by num_head_channels {num_head_channels}'
            self.num_heads = channels // num_head_channels
        self.use_checkpoint = use_checkpoint
        self.use_new_attention_order = use_new_attention_order
        self.norm = normalization(channels)
        self.qkv = conv_nd(1, channels, channels * 3, 1)
        self.proj_out = zero_module(conv_nd(1, channels, channels, 1))

    def packed_weights(self):
        """Conv1d qkv [3C, C, 1] -> (Wqk [2C, C] = [q of every head | k of every head], bqk, Wv [C, C], bv, Wp, bp).
        Legacy order (reference :366): output channel = (head*3 + {q,k,v})*d + i; new order (:398): ({q,k,v}*H + head)*d + i."""
        C, H = self.channels, self.num_heads
        d = C // H
        w = self.qkv.weight.detach().reshape(3 * C, C)
        b = self.qkv.bias.detach()
        if self.use_new_attention_order:
            w3, b3 = w.reshape(3, C, C), b.reshape(3, C)
        else:
            w3, b3 = w.reshape(H, 3, d, C).permute(1, 0, 2, 3).reshape(3, C, C), \
                b.reshape(H, 3, d).permute(1, 0, 2).reshape(3, C)
        wqk = torch.cat([w3[0], w3[1]], dim=0).to(torch.bfloat16).contiguous()
        bqk = torch.cat([b3[0], b3[1]]).float().contiguous()
        return (wqk, bqk, w3[2].to(torch.bfloat16).contiguous(), b3[2].float().contiguous(),
                self.proj_out.weight.detach().reshape(C, C).to(torch.bfloat16).contiguous(),
                self.proj_out.bias.detach().float().contiguous())


class UNetModel(_EngineModel):
    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False):
        super().__init__()
        if dims != 2:
            raise ValueError('the B200 kernels implement the 2-D UNet only (dims=2)')
        if use_fp16:
            raise ValueError('use_fp16 is not supported: the kernels use bf16 operands with fp32 accumulation')
        if num_heads_upsample == -1:
            num_heads_upsample = num_heads
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions, self.dropout, self.channel_mult = attention_resolutions, dropout, channel_mult
        self.conv_resample, self.num_classes, self.use_checkpoint = conv_resample, num_classes, use_checkpoint
        self.dtype = torch.float32
        self.num_heads, self.num_head_channels, self.num_heads_upsample = num_heads, num_head_channels, num_heads_upsample
        self.use_scale_shift_norm = use_scale_shift_norm

        emb_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, emb_dim), nn.SiLU(), linear(emb_dim, emb_dim))
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_classes, emb_dim)

        def res(cin, cout, **kw):
            return ResBlock(cin, emb_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint,
                            use_scale_shift_norm=use_scale_shift_norm, **kw)

        def attn(c, heads):
            return AttentionBlock(c, use_checkpoint=use_checkpoint, num_heads=heads,
                                  num_head_channels=num_head_channels,
                                  use_new_attention_order=use_new_attention_order)

        ch = input_ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, ch, 3, padding=1))])
        self._feature_size = ch
        skip_chans = [ch]
        ds = 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [res(ch, int(mult * model_channels))]
                ch = int(mult * model_channels)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch
                skip_chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(
                    res(ch, ch, down=True) if resblock_updown else Downsample(ch, conv_resample, dims=dims,
                                                                              out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2
                self._feature_size += ch

        self.middle_block = TimestepEmbedSequential(res(ch, ch), attn(ch, num_heads), res(ch, ch))
        self._feature_size += ch

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + skip_chans.pop(), int(model_channels * mult))]
                ch = int(model_channels * mult)
                if ds in attention_resolutions:
                    layers.append(attn(ch, num_heads_upsample))
                if level and i == num_res_blocks:
                    layers.append(res(ch, ch, up=True) if resblock_updown
                                  else Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))
                self._feature_size += ch

        self.out = nn.Sequential(
            normalization(ch), nn.SiLU(), zero_module(conv_nd(dims, input_ch, out_channels, 3, padding=1)))
        # not a registered submodule: carries no parameters, only the frequency table of timestep_embedding
        self.__dict__['_tfreq'] = TimestepFrequencies(model_channels)
        self._init_engine()

    def convert_to_fp16(self):
        raise RuntimeError('convert_to_fp16: not supported (bf16 operands / fp32 accumulation are built in)')

    def convert_to_fp32(self):
        return None

    # ------------------------------------------------------------------------------------------
    def _sequences(self):
        seqs = [(f'input_blocks.{i}', s) for i, s in enumerate(self.input_blocks)]
        seqs.append(('middle_block', self.middle_block))
        seqs += [(f'output_blocks.{i}', s) for i, s in enumerate(self.output_blocks)]
        return seqs

    def _res_blocks(self):
        return [(f'{n}.{j}', b) for n, s in self._sequences() for j, b in enumerate(s) if isinstance(b, ResBlock)]

    def _run_layer(self, name, layer, h, skip, emb, emb_ld, offsets):
        eng = self.engine
        if isinstance(layer, ResBlock):
            sc = layer.skip_connection if isinstance(layer.skip_connection, nn.Conv2d) else None
            return eng.resblock_core(name, h, skip, norm1=layer.in_layers[0], conv1=layer.in_layers[2],
                                     norm2=layer.out_layers[0], conv2=layer.out_layers[3], shortcut=sc, emb=emb,
                                     emb_off=offsets[name], emb_ld=emb_ld, scale_shift=layer.use_scale_shift_norm,
                                     resample=layer.resample, dropout=layer.out_layers[2],
                                     emb_linear=layer.emb_layers[1])
        assert skip is None
        if isinstance(layer, AttentionBlock):
            d = layer.channels // layer.num_heads
            return eng.attention_core(name, h, layer.norm, eng.packed(('attn', name), layer.packed_weights),
                                      layer.num_heads, 1.0 / math.sqrt(d), mods=layer)
        if isinstance(layer, Downsample):
            return eng.downsample_conv(name + '.op', layer.op, h) if layer.use_conv else eng.resample_plain(name, h, 1)
        if isinstance(layer, Upsample):
            return eng.upsample_conv(name + '.conv', layer.conv, h) if layer.use_conv else eng.resample_plain(name, h, 2)
        raise RuntimeError(f'{name}: unexpected layer {type(layer).__name__}')

    def _forward_impl(self, x, timesteps, y=None, out=None):
        """x: [N, C, H, W] fp32, timesteps: [N] int64, y: [N] int64 iff class-conditional (reference :653-682)."""
        assert (y is not None) == (self.num_classes is not None), \
            'must specify y if and only if the model is class-conditional'
        eng = self.engine
        eng.begin_forward()
        x = eng.check_input(x, timesteps, self.in_channels)
        B, _, H, W = x.shape
        if y is not None:
            assert y.shape == (B,)

        res_blocks = self._res_blocks()
        offsets, off = {}, 0
        for name, blk in res_blocks:
            offsets[name] = off
            off += blk.emb_layers[1].out_features
        emb, emb_ld = eng.embed(timesteps, y, B, self.__dict__['_tfreq'], self.time_embed[0], self.time_embed[2],
                                self.label_emb if self.num_classes is not None else None,
                                [blk.emb_layers[1] for _, blk in res_blocks])

        h = eng.first_conv('input_blocks.0', self.input_blocks[0][0], x)
        hs = [h]
        for i, seq in list(enumerate(self.input_blocks))[1:]:
            for j, layer in enumerate(seq):
                h = self._run_layer(f'input_blocks.{i}.{j}', layer, h, None, emb, emb_ld, offsets)
            hs.append(h)
        eng.pingpong = True      # from here on no output is a skip connection: block outputs alternate between two buffers
        for j, layer in enumerate(self.middle_block):
            h = self._run_layer(f'middle_block.{j}', layer, h, None, emb, emb_ld, offsets)
        for i, seq in enumerate(self.output_blocks):
            skip = hs.pop()
            for j, layer in enumerate(seq):
                h = self._run_layer(f'output_blocks.{i}.{j}', layer, h, skip if j == 0 else None, emb, emb_ld, offsets)

        return eng.head('out', h, self.out[0], self.out[2], out)
