"""Heun (2nd-order) sampler on the fused b200_ode_step kernel: drop-in for the reference's diffusions/heun.py:9-131
(two UNet forwards per step except the last; `denoise_1st_order` / `denoise_2nd_order` / `sample_loop`)."""
from typing import Dict

import torch
import tqdm
from torch import Tensor, nn as nn

from diffusions.euler import EulerSampler


class HeunSampler(EulerSampler):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._1st_order_derivative = None
        self._1st_order_xt = None

    def denoise(self, model_output, xt, t, t_prev):
        raise NotImplementedError('HeunSampler steps through denoise_1st_order / denoise_2nd_order')

    def denoise_1st_order(self, model_output: Tensor, xt: Tensor, t: int, t_prev: int):
        sample, pred_x0, deriv = self._ode(model_output, xt, t, t, t_prev, want_deriv=True)
        self._1st_order_derivative, self._1st_order_xt = deriv, xt.contiguous()
        return {'sample': sample, 'pred_x0': pred_x0}

    def denoise_2nd_order(self, model_output: Tensor, xt_prev: Tensor, t: int, t_prev: int):
        sample, pred_x0, _ = self._ode(model_output, xt_prev, t_prev, t, t_prev, d1=self._1st_order_derivative,
                                       x1=self._1st_order_xt)
        self._1st_order_derivative = self._1st_order_xt = None
        return {'sample': sample, 'pred_x0': pred_x0}

    def sample_loop(self, model: nn.Module, init_noise: Tensor, tqdm_kwargs: Dict = None, model_kwargs: Dict = None):
        tqdm_kwargs = dict() if tqdm_kwargs is None else tqdm_kwargs
        model_kwargs = dict() if model_kwargs is None else model_kwargs
        img = init_noise
        pairs = self._step_pairs()
        pbar = tqdm.tqdm(total=len(pairs), **tqdm_kwargs)
        for t, t_prev in pairs:
            t_batch = torch.full((1,), t, device=self.device, dtype=torch.long).expand(img.shape[0])
            out = self.denoise_1st_order(model(img, t_batch, **model_kwargs), img, t, t_prev)
            img = out['sample']
            if t_prev >= 0:
                tp_batch = torch.full((1,), t_prev, device=self.device, dtype=torch.long).expand(img.shape[0])
                out = self.denoise_2nd_order(model(img, tp_batch, **model_kwargs), img, t, t_prev)
                img = out['sample']
            pbar.update(1)
            yield out
        pbar.close()
