"""Euler sampler (EDM-style, sigma space) on the fused b200_ode_step kernel: drop-in for the reference's
diffusions/euler.py:7-66 (same constructor, `sigmas` attribute, `denoise` dict with `sample` / `pred_x0`; `sample` /
`sample_loop` inherited from DDPM, one UNet forward per step)."""
import torch
from torch import Tensor

import b200diff as K
from diffusions.ddpm import DDPM


class EulerSampler(DDPM):
    def __init__(self, total_steps: int = 1000, beta_schedule: str = 'linear', beta_start: float = 0.0001,
                 beta_end: float = 0.02, betas: Tensor = None, objective: str = 'pred_eps', clip_denoised: bool = True,
                 respace_type: str = None, respace_steps: int = 100, respaced_seq: Tensor = None,
                 device: torch.device = 'cpu', **kwargs):
        super().__init__(total_steps=total_steps, beta_schedule=beta_schedule, beta_start=beta_start, beta_end=beta_end,
                         betas=betas, objective=objective, clip_denoised=clip_denoised, respace_type=respace_type,
                         respace_steps=respace_steps, respaced_seq=respaced_seq, device=device, **kwargs)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod).sqrt()
        self._sigmas_host = ((1 - self._ac_host) / self._ac_host).sqrt()

    # the ODE samplers have their own update rule: no CUDA-graph runner for the DDPM posterior step
    def sample(self, model, init_noise, tqdm_kwargs=None, model_kwargs=None):
        sample = None
        for out in self.sample_loop(model, init_noise, tqdm_kwargs, model_kwargs):
            sample = out['sample']
        return sample

    def _sigma_pair(self, t: int, t_prev: int):
        return float(self._sigmas_host[t]), (float(self._sigmas_host[t_prev]) if t_prev >= 0 else 0.0)

    def _ode(self, model_output, x, t_eval, t, t_prev, d1=None, x1=None, want_deriv=False):
        x = x.contiguous()
        model_output = model_output.contiguous()
        sample, pred_x0 = torch.empty_like(x), torch.empty_like(x)
        deriv = torch.empty_like(x) if want_deriv else None
        st, sp = self._sigma_pair(t, t_prev)
        K.ode_step(model_output, x, self._predict_coefs(t_eval), st, sp, objective=self.objective, clip=self.clip_denoised,
                   d1=d1, x1=x1, sample=sample, pred_x0=pred_x0, deriv=deriv)
        return sample, pred_x0, deriv

    def denoise(self, model_output: Tensor, xt: Tensor, t: int, t_prev: int):
        """x_t -> x_{t_prev}: one Euler step of dx/dsigma = (x_bar - x0) / sigma."""
        sample, pred_x0, _ = self._ode(model_output, xt, t, t, t_prev)
        return {'sample': sample, 'pred_x0': pred_x0}
