"""Beta schedules and respaced timestep sequences (drop-in for the reference's diffusions/schedule.py:5-73).

Init-time host logic, required to be bit-exact with the reference: float64 linspace for linear/quad/const,
Python-float cosine values materialised as float32, integer index arithmetic for the respacings.
"""
import math

import torch

_BETA_KINDS = ('linear', 'quad', 'const', 'cosine')
_RESPACE_KINDS = ('uniform', 'uniform-leading', 'uniform-linspace', 'uniform-trailing', 'quad', 'none', None)


def _cosine_alpha_bar(u: float) -> float:
    return math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2


def get_beta_schedule(total_steps: int = 1000, beta_schedule: str = 'linear', beta_start: float = 0.0001,
                      beta_end: float = 0.02):
    """Returns a 1-D tensor of `total_steps` betas (float64, except 'cosine' which is float32 like the reference)."""
    if beta_schedule not in _BETA_KINDS:
        raise ValueError(f'Beta schedule {beta_schedule} is not supported.')
    if beta_schedule == 'cosine':
        ratios = (_cosine_alpha_bar((i + 1) / total_steps) / _cosine_alpha_bar(i / total_steps)
                  for i in range(total_steps))
        return torch.tensor([min(1 - r, 0.999) for r in ratios])
    if beta_schedule == 'const':
        return torch.full((total_steps,), fill_value=beta_end, dtype=torch.float64)
    power = 2 if beta_schedule == 'quad' else 1
    grid = torch.linspace(beta_start ** (1 / power), beta_end ** (1 / power), total_steps, dtype=torch.float64)
    return grid ** 2 if power == 2 else grid


def get_respaced_seq(total_steps: int = 1000, respace_type: str = 'uniform', respace_steps: int = 100):
    """Returns the int64 timesteps kept by the respaced sampler (ascending)."""
    if respace_type not in _RESPACE_KINDS:
        raise ValueError(f'Respace type {respace_type} is not supported.')
    if respace_type is None or respace_type == 'none':
        return torch.arange(0, total_steps).long()
    if respace_type == 'uniform-linspace':
        return torch.linspace(0, total_steps - 1, respace_steps).long()
    if respace_type == 'quad':
        return torch.floor(torch.linspace(0, math.sqrt(total_steps * 0.8), respace_steps) ** 2).long()
    stride = total_steps // respace_steps
    if respace_type == 'uniform-trailing':
        return torch.arange(total_steps - 1, -1, -stride).long().flip(dims=[0])
    return torch.arange(0, total_steps, stride).long()
