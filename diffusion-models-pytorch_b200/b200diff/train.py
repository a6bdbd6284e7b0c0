"""One training step of the noise-prediction objective on the kernels -- the body of `run_step` in the reference's
scripts/train_ddpm.py:171-192 / train_ddpm_cfg.py:172-196:

    optimizer.zero_grad(); t ~ U{0..T-1}; loss = diffuser.loss_func(model, x0, t); loss.backward();
    [all-reduce(mean) of the gradients over the data-parallel ranks]; clip_grad_norm_(1.0); optimizer.step(); ema.update()

with the last three fused into b200_optimizer_step, and no host synchronisation anywhere (the loss is returned as a
device scalar; call `.item()` on it only when it is logged).
"""
from typing import Dict, Optional

import torch

from .dist import allreduce_grads_
from .optim import FusedAdam


class TrainStep:
    def __init__(self, model, diffuser, optimizer: FusedAdam, ema=None, clip_grad_norm: Optional[float] = 1.0,
                 p_uncond: float = 0.0):
        self.model, self.diffuser, self.optimizer, self.ema = model, diffuser, optimizer, ema
        self.clip_grad_norm = clip_grad_norm
        self.p_uncond = p_uncond      # train_ddpm_cfg.py:183-186: the label is dropped with this probability

    def __call__(self, x0: torch.Tensor, t: torch.Tensor = None, y: torch.Tensor = None, eps: torch.Tensor = None,
                 micro_batch: int = None) -> torch.Tensor:
        B = x0.shape[0]
        micro_batch = B if micro_batch is None else micro_batch
        self.optimizer.zero_grad(set_to_none=True)
        total = None
        for i in range(0, B, micro_batch):
            xs = x0[i:i + micro_batch].float()
            ts = t[i:i + micro_batch] if t is not None else \
                torch.randint(self.diffuser.total_steps, (xs.shape[0],), device=xs.device).long()
            kw: Dict = {}
            if y is not None:
                # classifier-free guidance training: the whole micro-batch is unconditional with probability p_uncond
                drop = self.p_uncond > 0 and float(torch.rand(())) < self.p_uncond
                kw = dict(y=None if drop else y[i:i + micro_batch])
            loss = self.diffuser.loss_func(self.model, x0=xs, t=ts, eps=None if eps is None else eps[i:i + micro_batch],
                                           model_kwargs=kw)
            scale = xs.shape[0] / B
            (loss * scale).backward()
            total = loss.detach() * scale if total is None else total + loss.detach() * scale
        allreduce_grads_(self.model)
        self.optimizer.step(clip_grad_norm=self.clip_grad_norm, ema=self.ema)
        return total
