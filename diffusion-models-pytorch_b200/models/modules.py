"""Building blocks of the UNets (same names and state_dict keys as the reference's models/modules.py).

These classes are *parameter containers*: they own the fp32 nn.Parameters under the reference's key names and
registration order (so checkpoints, EMA and optimizers interoperate), but they do not compute anything with
torch ops.  The arithmetic of a forward pass is issued by models/engine.py as libb200diff kernels; calling a
block on its own raises, because there is no PyTorch fallback path.
"""
import math

import torch
import torch.nn as nn


class _KernelOnly(nn.Module):
    def forward(self, *args, **kwargs):
        raise RuntimeError(f'{type(self).__name__} is executed by the b200diff engine as part of the UNet forward; '
                           'it has no standalone PyTorch path')


class SinusoidalPosEmb(_KernelOnly):
    """t -> [sin(t f_i), cos(t f_i)], f_i = exp(-ln(1e4) i / (dim/2 - 1))  (reference modules.py:40-57)."""

    def __init__(self, dim: int):
        super().__init__()
        self.dim = dim

    def frequencies(self, device) -> torch.Tensor:
        half_dim = self.dim // 2
        step = math.log(10000) / (half_dim - 1)
        return torch.exp(torch.arange(half_dim) * -step).to(device)


def Upsample(in_channels: int, out_channels: int, use_conv: bool = True):
    """nearest-2x (+ 3x3 conv): keys `<prefix>.1.weight/bias` like the reference (modules.py:60-67)."""
    if use_conv:
        return nn.Sequential(
            nn.Upsample(scale_factor=2, mode='nearest'),
            nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1),
        )
    return nn.Upsample(scale_factor=2, mode='nearest')


def Downsample(in_channels: int, out_channels: int, use_conv: bool = True):
    """3x3 stride-2 conv: keys `<prefix>.weight/bias` like the reference (modules.py:70-74)."""
    if use_conv:
        return nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=2, padding=1)
    return nn.AvgPool2d(kernel_size=2, stride=2)


class SelfAttentionBlock(_KernelOnly):
    """GroupNorm -> q,k,v 1x1 -> softmax(q^T k / sqrt(d)) v -> 1x1 proj -> + x  (reference modules.py:77-102)."""

    def __init__(self, dim: int, n_heads: int = 1, groups: int = 32):
        super().__init__()
        assert dim % n_heads == 0
        self.n_heads = n_heads
        self.norm = nn.GroupNorm(groups, dim)
        self.q = nn.Conv2d(dim, dim, kernel_size=1)
        self.k = nn.Conv2d(dim, dim, kernel_size=1)
        self.v = nn.Conv2d(dim, dim, kernel_size=1)
        self.proj = nn.Conv2d(dim, dim, kernel_size=1)
        self.scale = (dim // n_heads) ** -0.5


class AdaGN(_KernelOnly):
    """GroupNorm(x) * (1 + ys) + yb with [ys, yb] = Linear(SiLU(embed))  (reference modules.py:105-123)."""

    def __init__(self, num_groups: int, num_channels: int, embed_dim: int):
        super().__init__()
        self.gn = nn.GroupNorm(num_groups, num_channels)
        self.proj = nn.Sequential(
            nn.SiLU(),
            nn.Linear(embed_dim, num_channels * 2),
        )
