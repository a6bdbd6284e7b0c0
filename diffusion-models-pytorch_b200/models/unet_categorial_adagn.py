"""Class-conditional UNet with AdaGN and BigGAN-style up/down residual blocks on the B200 kernels (drop-in for the
reference's models/unet_categorial_adagn.py:12-208: same constructor arguments, parameter names and order, and
the `model(X, T, y=None)` call; `y=None` skips the class embedding, which is what classifier-free guidance uses
for its unconditional branch).

Per ResBlock: GN+SiLU (+2x avg-pool / nearest-2x fused before the single bf16 rounding) -> conv3x3 ->
AdaGN (scale/shift folded into the GroupNorm kernel's per-channel coefficients) + SiLU -> conv3x3 + shortcut.
All scale/shift projections Linear(SiLU(emb)) of a forward run as ONE tensor-core GEMM.
"""
from typing import List

import torch
import torch.nn as nn
from torch import Tensor

import b200diff as K
from models.engine import Act
from models.modules import AdaGN, Downsample, SelfAttentionBlock, SinusoidalPosEmb, Upsample, _KernelOnly
from models.unet import _EngineModel


class ResBlock(_KernelOnly):
    def __init__(self, in_channels: int, out_channels: int, embed_dim: int, dropout: float = 0.1,
                 up: bool = False, down: bool = False):
        super().__init__()
        assert not (up and down), 'up and down cannot both be True'
        self.updown_kind = 'up' if up else 'down' if down else None
        self.blk1 = nn.Sequential(
            nn.GroupNorm(32, in_channels), nn.SiLU(), nn.Conv2d(in_channels, out_channels, 3, stride=1, padding=1))
        self.adagn = AdaGN(32, out_channels, embed_dim)
        self.blk2 = nn.Sequential(
            nn.SiLU(), nn.Dropout(dropout), nn.Conv2d(out_channels, out_channels, 3, stride=1, padding=1))
        self.shortcut = nn.Conv2d(in_channels, out_channels, 1) if in_channels != out_channels else nn.Identity()


class ResBlockUpsample(ResBlock):
    def __init__(self, in_channels: int, out_channels: int, embed_dim: int, dropout: float = 0.1):
        super().__init__(in_channels, out_channels, embed_dim, dropout, up=True)


class ResBlockDownsample(ResBlock):
    def __init__(self, in_channels: int, out_channels: int, embed_dim: int, dropout: float = 0.1):
        super().__init__(in_channels, out_channels, embed_dim, dropout, down=True)


class UNetCategorialAdaGN(_EngineModel):
    # one network serves the conditional and the unconditional branch of classifier-free guidance (y=None drops the class
    # embedding, reference :172-174), so the sampling runner may evaluate both as one 2B batch with labels [y ; -1]
    cfg_batch_ok = True

    """UNet conditioned on categorial labels with AdaGN."""

    def __init__(
            self,
            in_channels: int = 3,
            out_channels: int = 3,
            dim: int = 128,
            dim_mults: List[int] = (1, 2, 2, 2),
            use_attn: List[int] = (False, True, True, False),
            num_res_blocks: int = 2,
            num_classes: int = None,
            attn_head_dims: int = 64,
            resblock_updown: bool = True,
            dropout: float = 0.1,
    ):
        super().__init__()
        n_stages = len(dim_mults)
        widths = [dim * m for m in dim_mults]
        embed_dim = dim * 4
        self.time_embed = nn.Sequential(
            SinusoidalPosEmb(dim), nn.Linear(dim, embed_dim), nn.SiLU(), nn.Linear(embed_dim, embed_dim))
        self.class_embed = nn.Embedding(num_classes, embed_dim) if num_classes is not None else None
        self.first_conv = nn.Conv2d(in_channels, dim, 3, stride=1, padding=1)

        def attn(width):
            assert width % attn_head_dims == 0
            return SelfAttentionBlock(width, n_heads=width // attn_head_dims)

        skip_widths = [dim]
        cur = dim
        self.down_blocks = nn.ModuleList()
        for i, width in enumerate(widths):
            stage = nn.ModuleList()
            for _ in range(num_res_blocks):
                stage.append(ResBlock(cur, width, embed_dim=embed_dim, dropout=dropout))
                if use_attn[i]:
                    stage.append(attn(width))
                skip_widths.append(width)
                cur = width
            if i < n_stages - 1:
                stage.append(ResBlockDownsample(width, width, embed_dim=embed_dim, dropout=dropout)
                             if resblock_updown else Downsample(width, width))
                skip_widths.append(width)
            self.down_blocks.append(stage)

        self.bottleneck_block = nn.ModuleList([
            ResBlock(cur, cur, embed_dim=embed_dim, dropout=dropout),
            SelfAttentionBlock(cur),
            ResBlock(cur, cur, embed_dim=embed_dim, dropout=dropout),
        ])

        self.up_blocks = nn.ModuleList()
        for i in reversed(range(n_stages)):
            width = widths[i]
            stage = nn.ModuleList()
            for _ in range(num_res_blocks + 1):
                stage.append(ResBlock(skip_widths.pop() + cur, width, embed_dim=embed_dim, dropout=dropout))
                if use_attn[i]:
                    stage.append(attn(width))
                cur = width
            if i > 0:
                stage.append(ResBlockUpsample(width, width, embed_dim=embed_dim, dropout=dropout)
                             if resblock_updown else Upsample(width, width))
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

    def _forward_impl(self, X: Tensor, T: Tensor, y: Tensor = None, out: Tensor = None):
        """X: [B, C, H, W] fp32, T: [B] int64, y: [B] int64 labels or None -> [B, C_out, H, W] fp32."""
        eng = self.engine
        eng.begin_forward()
        X = eng.check_input(X, T, self.in_channels)
        B, _, H, W = X.shape

        res_blocks = self._res_blocks()
        offsets, off = {}, 0
        for name, blk in res_blocks:
            offsets[name] = off
            off += blk.adagn.proj[1].out_features
        ss, ss_ld = eng.embed(T, y, B, self.time_embed[0], self.time_embed[1], self.time_embed[3], self.class_embed,
                              [blk.adagn.proj[1] for _, blk in res_blocks])

        skips = []

        # op sequence with consumer look-ahead (as in models/unet.py): a block whose output is the only GroupNorm input of
        # the next op lets its last conv apply that GroupNorm (Engine.fuse_gn1)
        ops = []
        for i, stage in enumerate(self.down_blocks):
            for j, blk in enumerate(stage):
                kind = 'res' if isinstance(blk, ResBlock) else 'attn' if isinstance(blk, SelfAttentionBlock) else 'down'
                ops.append((kind, f'down_blocks.{i}.{j}', blk, 'enc'))
        for j, blk in enumerate(self.bottleneck_block):
            ops.append(('res' if isinstance(blk, ResBlock) else 'attn', f'bottleneck_block.{j}', blk, 'mid'))
        for i, stage in enumerate(self.up_blocks):
            for j, blk in enumerate(stage):
                kind = ('resup' if isinstance(blk, ResBlockUpsample) else 'res' if isinstance(blk, ResBlock)
                        else 'attn' if isinstance(blk, SelfAttentionBlock) else 'up')
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
            if (kind == 'res' and part != 'dec' and not isinstance(blk.shortcut, nn.Conv2d)
                    and getattr(blk, 'updown_kind', None) is None):
                return blk.blk1[0], True
            if (kind == 'res' and part == 'dec' and skips and isinstance(blk.shortcut, nn.Conv2d)
                    and blk.shortcut.kernel_size[0] == 1 and getattr(blk, 'updown_kind', None) is None):
                return blk.blk1[0], True, skips[-1].C, dead      # consumer normalises cat(x, skip): x's part by its producer
            return None

        # the first convolution's consumer is ops[0]; on the tensor-core path it applies that block's norm1 too
        h = eng.first_conv('first_conv', self.first_conv, X, next_gn=consumer_gn(-1, None) if ops[0][0] == 'res' else None)
        skips.append(h)

        for k, (kind, name, blk, part) in enumerate(ops[:-1]):
            if part == 'mid' and not eng.pingpong:
                eng.pingpong = True      # from here on no output is a skip connection: block outputs alternate between two buffers
            if kind == 'res' or kind == 'resup':
                skip = skips.pop() if (part == 'dec' and kind == 'res') else None
                h = eng.resblock_adagn(name, blk, h, skip, ss, offsets[name], ss_ld, next_gn=consumer_gn(k, h))
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
