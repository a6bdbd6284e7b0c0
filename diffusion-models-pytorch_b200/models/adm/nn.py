"""Small helpers of the ADM (openai/guided-diffusion) UNet family (reference models/adm/nn.py).

Only what the hot path needs: layer constructors with the reference's names, zero initialisation and the
frequency table of `timestep_embedding` (the sinusoid itself is evaluated by b200_time_embed).
"""
import math

import torch
import torch.nn as nn


class GroupNorm32(nn.GroupNorm):
    """Parameter container; the reference's float32 upcast (nn.py:17-19) is what the kernels do anyway:
    GroupNorm statistics and the affine are always evaluated in fp32."""


def normalization(channels: int) -> nn.GroupNorm:
    return GroupNorm32(32, channels)


def conv_nd(dims: int, *args, **kwargs) -> nn.Module:
    if dims == 1:
        return nn.Conv1d(*args, **kwargs)
    if dims == 2:
        return nn.Conv2d(*args, **kwargs)
    raise ValueError(f'unsupported dimensions: {dims} (the B200 kernels implement the 2-D UNet)')


def linear(*args, **kwargs) -> nn.Linear:
    return nn.Linear(*args, **kwargs)


def zero_module(module: nn.Module) -> nn.Module:
    with torch.no_grad():
        for p in module.parameters():
            p.zero_()
    return module


class TimestepFrequencies(nn.Module):
    """Frequencies of `timestep_embedding` (nn.py:103-121): f_i = exp(-ln(max_period) i / half), embedding laid
    out [cos | sin] -- unlike models/modules.py's SinusoidalPosEmb ([sin | cos], divisor half - 1)."""
    cos_first = True

    def __init__(self, dim: int, max_period: int = 10000):
        super().__init__()
        if dim % 2:
            raise ValueError('odd embedding widths are not supported by b200_time_embed')
        self.dim, self.max_period = dim, max_period

    def frequencies(self, device) -> torch.Tensor:
        half = self.dim // 2
        return torch.exp(-math.log(self.max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(device)
