"""DDPM-style UNet on the B200 kernels (drop-in for the reference's models/unet.py: same constructor
arguments, parameter names, registration order and `model(X, T)` call, reference models/unet.py:46-152).

The module owns fp32 parameters exactly like the reference (so `state_dict()` / `load_state_dict()` /
`models.EMA` / torch optimizers interoperate); `forward` issues libb200diff kernels through models/engine.py:
GroupNorm+SiLU (K3) -> tcgen05 implicit-GEMM convs with fused bias / time-embedding / shortcut (K1) ->
fused attention (K2).  Dropout is the identity in eval mode; training mode is rejected until the backward
kernels land (there is no autograd fallback).
"""
import weakref
from typing import List

import torch
import torch.nn as nn
from torch import Tensor

import b200diff as K
from models.engine import Act, Engine
from models.modules import Downsample, SelfAttentionBlock, SinusoidalPosEmb, Upsample, _KernelOnly
from models.runner import SamplingRunner


class ResBlock(_KernelOnly):
    """GN-SiLU-conv3x3 (+ time embedding) -> GN-SiLU-dropout-conv3x3, plus identity / 1x1 shortcut."""

    def __init__(self, in_channels: int, out_channels: int, embed_dim: int, dropout: float = 0.1):
        super().__init__()
        self.blk1 = nn.Sequential(
            nn.GroupNorm(32, in_channels), nn.SiLU(), nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1))
        self.proj = nn.Sequential(nn.SiLU(), nn.Linear(embed_dim, out_channels))
        self.blk2 = nn.Sequential(
            nn.GroupNorm(32, out_channels), nn.SiLU(), nn.Dropout(dropout),
            nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1))
        self.shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()


class _EngineModel(nn.Module):
    """Shared plumbing of the engine-backed UNets."""
    use_cuda_graph = True

    def _init_engine(self):
        self.__dict__['_engine'] = Engine(self)     # not a submodule / not in state_dict
        self.__dict__['_runners'] = weakref.WeakKeyDictionary()   # diffuser -> SamplingRunner (dies with the diffuser)

    @property
    def engine(self) -> Engine:
        return self.__dict__['_engine']

    def make_sampling_runner(self, diffuser):
        """Per-diffuser CUDA-graph sampler used by DDPM/DDIM(.CFG).sample()."""
        runners = self.__dict__['_runners']
        r = runners.get(diffuser)
        if r is None:
            r = runners[diffuser] = SamplingRunner(self, diffuser)
        return r

    def set_precision(self, precision: str):
        """'bf16' (default: bf16 tensor-core operands, fp32 accumulate; eps rel-L2 <= 1e-2 vs the fp32 reference) or
        'fp32' (bf16 hi/lo split operands, 3 MMAs per product: eps rel-L2 <= 1e-4; inference only).  Returns self."""
        self.engine.set_precision(precision)
        return self

    def forward(self, *args, **kwargs):
        """Inference: launches the forward kernels.  Training (module in train mode, grad enabled): the same forward
        with dropout and a tape, wrapped in an autograd node whose backward runs the hand-written adjoint kernels
        (models/backward.py) -- autograd itself never sees the network's interior."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            if self.engine.split:
                raise RuntimeError("precision='fp32' is an inference mode: call under torch.no_grad(), or "
                                   "set_precision('bf16') for the training step")
            from models.backward import UNetFunction
            return UNetFunction.apply(self, args, kwargs, *list(self.parameters()))
        return self._forward_impl(*args, **kwargs)


class UNet(_EngineModel):
    def __init__(
            self,
            in_channels: int = 3,
            out_channels: int = 3,
            dim: int = 128,
            dim_mults: List[int] = (1, 2, 2, 2),
            use_attn: List[int] = (False, True, False, False),
            num_res_blocks: int = 2,
            n_heads: int = 1,
            dropout: float = 0.1,
    ):
        super().__init__()
        n_stages = len(dim_mults)
        widths = [dim * m for m in dim_mults]
        embed_dim = dim * 4
        self.time_embed = nn.Sequential(
            SinusoidalPosEmb(dim), nn.Linear(dim, embed_dim), nn.SiLU(), nn.Linear(embed_dim, embed_dim))
        self.first_conv = nn.Conv2d(in_channels, dim, 3, stride=1, padding=1)

        def res(cin, cout):
            return ResBlock(cin, cout, embed_dim=embed_dim, dropout=dropout)

        # encoder: per stage [res (attn)] * num_res_blocks (+ strided conv); every output is a skip connection
        skip_widths = [dim]
        cur = dim
        self.down_blocks = nn.ModuleList()
        for i, width in enumerate(widths):
            stage = nn.ModuleList()
            for _ in range(num_res_blocks):
                stage.append(res(cur, width))
                if use_attn[i]:
                    stage.append(SelfAttentionBlock(width, n_heads=n_heads))
                skip_widths.append(width)
                cur = width
            if i < n_stages - 1:
                stage.append(Downsample(width, width))
                skip_widths.append(width)
            self.down_blocks.append(stage)

        self.bottleneck_block = nn.ModuleList([res(cur, cur), SelfAttentionBlock(cur), res(cur, cur)])

        # decoder: per stage [res(cat skip) (attn)] * (num_res_blocks + 1) (+ nearest-2x conv)
        self.up_blocks = nn.ModuleList()
        for i in reversed(range(n_stages)):
            width = widths[i]
            stage = nn.ModuleList()
            for _ in range(num_res_blocks + 1):
                stage.append(res(skip_widths.pop() + cur, width))
                if use_attn[i]:
                    stage.append(SelfAttentionBlock(width, n_heads=n_heads))
                cur = width
            if i > 0:
                stage.append(Upsample(width, width))
            self.up_blocks.append(stage)

        self.last_conv = nn.Sequential(
            nn.GroupNorm(32, cur), nn.SiLU(), nn.Conv2d(cur, out_channels, 3, stride=1, padding=1))
        self.in_channels, self.out_channels = in_channels, out_channels
        self._init_engine()

    def _res_blocks(self):
        blocks = [(f'down_blocks.{i}.{j}', b) for i, st in enumerate(self.down_blocks) for j, b in enumerate(st)]
        blocks += [(f'bottleneck_block.{j}', b) for j, b in enumerate(self.bottleneck_block)]
        blocks += [(f'up_blocks.{i}.{j}', b) for i, st in enumerate(self.up_blocks) for j, b in enumerate(st)]
        return [(n, b) for n, b in blocks if isinstance(b, ResBlock)]

    def _forward_impl(self, X: Tensor, T: Tensor, out: Tensor = None):
        """X: [B, C, H, W] fp32 (NCHW), T: [B] int64 -> [B, C_out, H, W] fp32.  (reference unet.py:121-152)"""
        eng = self.engine
        eng.begin_forward()
        X = eng.check_input(X, T, self.in_channels)
        B, _, H, W = X.shape

        res_blocks = self._res_blocks()
        offsets, off = {}, 0
        for name, blk in res_blocks:
            offsets[name] = off
            off += blk.proj[1].out_features
        tproj, tld = eng.embed(T, None, B, self.time_embed[0], self.time_embed[1], self.time_embed[3], None,
                               [blk.proj[1] for _, blk in res_blocks])

        skips = []

        # the op sequence (reference unet.py:127-150), flattened so that each op knows its consumer: a block whose output
        # is the ONLY GroupNorm input of the next op (a ResBlock without a concatenated skip and with an identity
        # shortcut, an attention block, the output head) lets its last conv apply that GroupNorm (Engine.fuse_gn1)
        ops = []
        for i, stage in enumerate(self.down_blocks):
            for j, blk in enumerate(stage):
                kind = 'res' if isinstance(blk, ResBlock) else 'attn' if isinstance(blk, SelfAttentionBlock) else 'down'
                ops.append((kind, f'down_blocks.{i}.{j}', blk, 'enc'))
        for j, blk in enumerate(self.bottleneck_block):
            ops.append(('res' if isinstance(blk, ResBlock) else 'attn', f'bottleneck_block.{j}', blk, 'mid'))
        for i, stage in enumerate(self.up_blocks):
            for j, blk in enumerate(stage):
                kind = 'res' if isinstance(blk, ResBlock) else 'attn' if isinstance(blk, SelfAttentionBlock) else 'up'
                ops.append((kind, f'up_blocks.{i}.{j}', blk, 'dec'))
        ops.append(('head', 'last_conv', None, 'dec'))

        def consumer_gn(k, x):
            kind, _, blk, part = ops[k + 1]
            dead = k >= 0 and ops[k][3] != 'enc'      # the producer's fp32 output is no skip connection: `head` / a concatenating
            if kind == 'head':             # block read only the producer-applied bf16 forms, so it need not be written
                return self.last_conv[0], True, 0, dead
            if kind == 'attn':      # the one-launch attention block normalises its input itself
                fused = eng.attn_block and K.attn_block_ok(x.H * x.W, blk.q.out_channels, blk.n_heads, blk.norm.num_groups)
                return None if fused else (blk.norm, False)
            if kind == 'res' and part != 'dec' and not isinstance(blk.shortcut, nn.Conv2d):
                return blk.blk1[0], True
            if (kind == 'res' and part == 'dec' and skips and isinstance(blk.shortcut, nn.Conv2d)
                    and blk.shortcut.kernel_size[0] == 1):
                return blk.blk1[0], True, skips[-1].C, dead      # consumer normalises cat(x, skip): x's part by its producer
            return None

        # the first convolution's consumer is ops[0]; on the tensor-core path it applies that block's norm1 too
        h = eng.first_conv('first_conv', self.first_conv, X, next_gn=consumer_gn(-1, None) if ops[0][0] == 'res' else None)
        skips.append(h)

        for k, (kind, name, blk, part) in enumerate(ops[:-1]):
            if part == 'mid' and not eng.pingpong:
                eng.pingpong = True      # from here on no output is a skip connection: block outputs alternate between two buffers
            if kind == 'res':
                skip = skips.pop() if part == 'dec' else None      # before the look-ahead: the consumer takes skips[-1]
                h = eng.resblock(name, blk, h, skip, tproj, offsets[name], tld, next_gn=consumer_gn(k, h))
                if part == 'enc':
                    skips.append(h)
            elif kind == 'attn':
                h = eng.attention(name, blk, h, next_gn=consumer_gn(k, h))
                if part == 'enc':
                    skips[-1] = h
            elif kind == 'down':
                h = eng.downsample_conv(name, blk, h)
                skips.append(h)
            else:
                # the next op is a decoder ResBlock that concatenates a skip connection (GroupNorm + 1x1 shortcut read it,
                # nobody adds it as an fp32 residual) -> the up-conv output may stay bf16
                nk, _, nblk, _ = ops[k + 1]
                cat_next = (nk == 'res' and bool(skips) and isinstance(nblk.shortcut, nn.Conv2d)
                            and nblk.shortcut.kernel_size[0] == 1 and getattr(nblk, 'updown_kind', None) is None)
                h = eng.upsample_conv(name, blk[1], h, bf16_out=cat_next)

        return eng.head('last_conv', h, self.last_conv[0], self.last_conv[2], out)
