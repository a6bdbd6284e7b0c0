"""Exponential moving average of model parameters (API of the reference's models/ema.py:7-79).

Host-side utility around the training step: a list-of-tensors shadow in `model.parameters()` order, updated
with multi-tensor (foreach) ops (b200diff.optim.FusedAdam(W) folds the same update into its kernel).
"""
from typing import Iterable

import torch
import torch.nn as nn


class EMA:
    def __init__(self, parameters: Iterable[nn.Parameter], decay: float = 0.9999, gradual: bool = True):
        self.decay = decay
        self.gradual = gradual
        self.num_updates = 0
        self.shadow = [p.detach().clone() for p in parameters]
        self.backup = []

    def get_decay(self):
        if not self.gradual:
            return self.decay
        return min(self.decay, (1 + self.num_updates) / (10 + self.num_updates))

    @torch.no_grad()
    def update(self, parameters: Iterable[nn.Parameter]):
        self.num_updates += 1
        decay = self.get_decay()
        params = list(parameters)
        live = [(s, p) for s, p in zip(self.shadow, params) if p.requires_grad]
        frozen = [(s, p) for s, p in zip(self.shadow, params) if not p.requires_grad]
        if live:
            shadows, ps = [s for s, _ in live], [p.detach() for _, p in live]
            delta = torch._foreach_sub(shadows, ps)
            torch._foreach_mul_(delta, 1. - decay)
            torch._foreach_sub_(shadows, delta)
        for s, p in frozen:
            s.copy_(p)

    @torch.no_grad()
    def apply_shadow(self, parameters: Iterable[nn.Parameter]):
        """models/ema.py:40-45 of the reference.  The copy goes through `p.copy_` (not `p.data.copy_`) so that the
        parameter's version counter moves: models/engine.py re-packs its bf16 operands and drops captured sampling
        graphs when (data_ptr, _version) of a parameter changes."""
        assert len(self.backup) == 0, 'backup is not empty'
        for s, p in zip(self.shadow, parameters):
            self.backup.append(p.detach().cpu().clone())
            p.copy_(s)

    @torch.no_grad()
    def restore(self, parameters: Iterable[nn.Parameter]):
        assert len(self.backup) > 0, 'backup is empty'
        for b, p in zip(self.backup, parameters):
            p.copy_(b.to(p.device))
        self.backup = []

    def state_dict(self):
        return dict(decay=self.decay, shadow=self.shadow, num_updates=self.num_updates)

    def load_state_dict(self, state_dict):
        self.decay = state_dict['decay']
        self.shadow = state_dict['shadow']
        self.num_updates = state_dict['num_updates']

    def to(self, device):
        self.shadow = [s.to(device) for s in self.shadow]
