"""DDIM sampler on the B200 kernels (drop-in for the reference's diffusions/ddim.py).

`DDIM` (reference ddim.py:12-132) and `DDIMCFG` (:135-250) keep their constructor, attributes and methods; the
per-step update x_{t-1} = sqrt(ac_prev) x0 + sqrt(1 - ac_prev - var) eps + sqrt(var) z runs in the fused
`b200_sampler_step` kernel with host-evaluated scalar coefficients (see diffusions/ddpm.py).
"""
from typing import Any, Dict

import torch
import torch.nn as nn
import tqdm
from torch import Tensor

import b200diff as K
from diffusions.ddpm import DDPM, _CFGMixin, _graph_runner


class DDIM(DDPM):
    def __init__(
            self,
            total_steps: int = 1000,
            beta_schedule: str = 'linear',
            beta_start: float = 0.0001,
            beta_end: float = 0.02,
            betas: Tensor = None,
            objective: str = 'pred_eps',

            clip_denoised: bool = True,
            respace_type: str = None,
            respace_steps: int = 100,
            respaced_seq: Tensor = None,
            eta: float = 0.,

            device: torch.device = 'cpu',
            **kwargs,
    ):
        super().__init__(
            total_steps=total_steps, beta_schedule=beta_schedule, beta_start=beta_start, beta_end=beta_end,
            betas=betas, objective=objective, clip_denoised=clip_denoised, respace_type=respace_type,
            respace_steps=respace_steps, respaced_seq=respaced_seq, device=device, **kwargs,
        )
        self.eta = eta

    def _uses_learned_var(self) -> bool:
        return False   # DDIM ignores the learned-variance channels (reference ddim.py:57-86)

    def _step_coefs(self, t: int, t_prev: int):
        """[x0_coef, xt_coef, eps_coef, var, -, -] of the DDIM update (ddim.py:65-73), reference op order."""
        ac_t, ac_prev = self._ac_pair(t, t_prev)
        var = ((self.eta ** 2) * (1. - ac_prev) / (1. - ac_t) * (1. - ac_t / ac_prev))
        zero = torch.zeros(())
        return [torch.sqrt(ac_prev), zero, torch.sqrt(1. - ac_prev - var), var, zero, zero]

    def denoise(self, model_output: Tensor, xt: Tensor, t: int, t_prev: int, reverse_eps: Tensor = None):
        """Sample from p_theta(x{t-1} | xt) (ddim.py:57-86)."""
        return self._denoise_impl(model_output, xt, t, t_prev, reverse_eps)

    # ---- DDIM inversion (ddim.py:88-132): x_{t+1} = sqrt(ac_next) x0 + sqrt(1 - ac_next) eps ----
    def _inversion_row(self, t: int, t_next: int) -> Tensor:
        key = ('inv', t, t_next)
        row = self._coef_rows.get(key)
        if row is None:
            ac_next = self._ac_host[t_next] if t_next < self.total_steps else torch.tensor(0.0)
            zero = torch.zeros(())
            vals = self._predict_coefs(t) + [torch.sqrt(ac_next), zero, torch.sqrt(1. - ac_next), zero, zero, zero,
                                             zero]
            host = torch.stack([torch.as_tensor(v, dtype=torch.float32).reshape(()) for v in vals] +
                               [torch.zeros(())] * (K.SC_COUNT - len(vals)))
            row = host.to(self.device)
            self._coef_rows[key] = row
        return row

    def _inversion_impl(self, model_output, xt, t, t_next, model_output_uncond=None, guidance_scale=1.0):
        if self.eta != 0.:
            raise ValueError(f'DDIM inversion is only valid when eta=0, get {self.eta}')
        xt = xt.contiguous()
        sample, pred_x0, pred_eps = torch.empty_like(xt), torch.empty_like(xt), torch.empty_like(xt)
        K.sampler_step(model_output.contiguous(), xt, self._inversion_row(t, t_next), objective=self.objective,
                       clip=self.clip_denoised,
                       model_out_uncond=None if model_output_uncond is None else model_output_uncond.contiguous(),
                       guidance_scale=guidance_scale, sample=sample, pred_x0=pred_x0, pred_eps=pred_eps)
        return {'sample': sample, 'pred_x0': pred_x0, 'pred_eps': pred_eps}

    def denoise_inversion(self, model_output: Tensor, xt: Tensor, t: int, t_next: int):
        """Sample x{t+1} from xt, only valid for DDIM (eta=0)."""
        return self._inversion_impl(model_output, xt, t, t_next)

    def sample_inversion_loop(self, model: nn.Module, img: Tensor, tqdm_kwargs: Dict = None,
                              model_kwargs: Dict = None):
        tqdm_kwargs = dict() if tqdm_kwargs is None else tqdm_kwargs
        model_kwargs = dict() if model_kwargs is None else model_kwargs
        sample_seq = self.respaced_seq[:-1].tolist()
        sample_seq_next = self.respaced_seq[1:].tolist()
        pbar = tqdm.tqdm(total=len(sample_seq), **tqdm_kwargs)
        for t, t_next in zip(sample_seq, sample_seq_next):
            t_batch = torch.full((1, ), t, device=img.device, dtype=torch.long).expand(img.shape[0])
            model_output = model(img, t_batch, **model_kwargs)
            out = self.denoise_inversion(model_output, img, t, t_next)
            img = out['sample']
            pbar.update(1)
            yield out
        pbar.close()

    def sample_inversion(self, model: nn.Module, img: Tensor, tqdm_kwargs: Dict = None, model_kwargs: Dict = None):
        sample = None
        for out in self.sample_inversion_loop(model, img, tqdm_kwargs, model_kwargs):
            sample = out['sample']
        return sample


class DDIMCFG(_CFGMixin, DDIM):
    def __init__(self, guidance_scale: float = 1., cond_kwarg: str = 'y', *args, **kwargs):
        """DDIM with classifier-free guidance (reference ddim.py:135-159); guidance_scale s: 0 = unconditional,
        1 = conditional, > 1 = guided."""
        DDIM.__init__(self, *args, **kwargs)
        self.guidance_scale = guidance_scale
        self.cond_kwarg = cond_kwarg

    def sample_loop(self, model: nn.Module, init_noise: Tensor, uncond_conditioning: Any = None,
                    tqdm_kwargs: Dict = None, model_kwargs: Dict = None):
        yield from self._cfg_loop(model, init_noise, uncond_conditioning, tqdm_kwargs, model_kwargs)

    def sample(self, model: nn.Module, init_noise: Tensor, uncond_conditioning: Any = None,
               tqdm_kwargs: Dict = None, model_kwargs: Dict = None):
        runner = _graph_runner(self, model)
        if runner is not None:
            return runner.run(init_noise, tqdm_kwargs, model_kwargs, self.guidance_scale, uncond_conditioning)
        sample = None
        for out in self.sample_loop(model, init_noise, uncond_conditioning, tqdm_kwargs, model_kwargs):
            sample = out['sample']
        return sample

    def sample_inversion_loop(self, model: nn.Module, img: Tensor, uncond_conditioning: Any = None,
                              tqdm_kwargs: Dict = None, model_kwargs: Dict = None):
        tqdm_kwargs = dict() if tqdm_kwargs is None else tqdm_kwargs
        if self.cond_kwarg not in model_kwargs.keys():
            raise ValueError(f'Condition argument `{self.cond_kwarg}` not found in model_kwargs.')
        uncond_model_kwargs = model_kwargs.copy()
        uncond_model_kwargs[self.cond_kwarg] = uncond_conditioning
        sample_seq = self.respaced_seq[:-1].tolist()
        sample_seq_next = self.respaced_seq[1:].tolist()
        pbar = tqdm.tqdm(total=len(sample_seq), **tqdm_kwargs)
        for t, t_next in zip(sample_seq, sample_seq_next):
            t_batch = torch.full((1, ), t, device=img.device, dtype=torch.long).expand(img.shape[0])
            cond = model(img, t_batch, **model_kwargs)
            uncond = model(img, t_batch, **uncond_model_kwargs)
            out = self._inversion_impl(cond, img, t, t_next, uncond, self.guidance_scale)
            img = out['sample']
            pbar.update(1)
            yield out
        pbar.close()

    def sample_inversion(self, model: nn.Module, img: Tensor, clip_denoised: bool = None, eta: float = None,
                         guidance_scale: float = None, uncond_conditioning: Any = None,
                         tqdm_kwargs: Dict = None, model_kwargs: Dict = None):
        sample = None
        for out in self.sample_inversion_loop(model, img, uncond_conditioning, tqdm_kwargs, model_kwargs):
            sample = out['sample']
        return sample
