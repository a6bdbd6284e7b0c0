"""DDPM forward/reverse process on the B200 kernels (drop-in for the reference's diffusions/ddpm.py).

Same constructor, attributes and methods as the reference classes `DDPM` (diffusions/ddpm.py:13-290) and
`DDPMCFG` (:293-368); the arithmetic is different plumbing:

* every per-step scalar (the ~35 zero-dim tensor ops of ddpm.py:102-110,218-246) is evaluated once on the host,
  in fp32 and in the reference's operation order, into a row of a device-resident coefficient table;
* predict -> clip -> posterior mean -> variance -> noise injection (and the classifier-free-guidance mix)
  is ONE fused elementwise kernel, `b200_sampler_step`;
* `sample()` replays a CUDA graph of {UNet forward(s), noise draw, sampler step} per timestep when the model
  supports it; `sample_loop()` stays a generator yielding the reference's dict per step.

There is no CPU path: tensors must live on a CUDA device (RuntimeError otherwise).
"""
from contextlib import contextmanager
from typing import Any, Dict

import torch
import torch.nn as nn
import tqdm
from torch import Tensor

import b200diff as K
from diffusions.schedule import get_beta_schedule, get_respaced_seq


class DDPM:
    def __init__(
            self,
            total_steps: int = 1000,
            beta_schedule: str = 'linear',
            beta_start: float = 0.0001,
            beta_end: float = 0.02,
            betas: Tensor = None,
            objective: str = 'pred_eps',

            var_type: str = 'fixed_large',
            clip_denoised: bool = True,
            respace_type: str = None,
            respace_steps: int = 100,
            respaced_seq: Tensor = None,

            device: torch.device = 'cpu',
    ):
        if objective not in ['pred_eps', 'pred_x0', 'pred_v']:
            raise ValueError(f'Invalid objective: {objective}')
        if var_type not in ['fixed_small', 'fixed_large', 'learned_range']:
            raise ValueError(f'Invalid var_type: {var_type}')

        self.total_steps = total_steps
        self.objective = objective
        self.var_type = var_type
        self.clip_denoised = clip_denoised
        self.device = device

        if betas is None:
            betas = get_beta_schedule(total_steps=total_steps, beta_schedule=beta_schedule,
                                      beta_start=beta_start, beta_end=beta_end)
        assert isinstance(betas, Tensor)
        assert betas.shape == (total_steps, )
        cumprod = torch.cumprod(1. - betas, dim=0)
        self._ac_host = cumprod.to('cpu', torch.float)          # host copy: scalar coefficient arithmetic
        self.alphas_cumprod = cumprod.to(device, torch.float)   # public attribute, as in the reference

        if respaced_seq is None:
            respaced_seq = get_respaced_seq(total_steps=total_steps, respace_type=respace_type,
                                            respace_steps=respace_steps)
        assert isinstance(respaced_seq, Tensor)
        assert respaced_seq.ndim == 1
        self.respaced_seq = respaced_seq.to(device)
        self._coef_rows: Dict[Any, Tensor] = {}

    def set_respaced_seq(self, respace_type: str = 'uniform', respace_steps: int = 100):
        self.respaced_seq = get_respaced_seq(total_steps=self.total_steps, respace_type=respace_type,
                                             respace_steps=respace_steps).to(self.device)

    # ------------------------------------------------------------------------------------------
    # host-side scalar coefficients (reference op order, fp32 zero-dim CPU tensors)
    # ------------------------------------------------------------------------------------------
    def _ac_pair(self, t: int, t_prev: int):
        ac_t = self._ac_host[t]
        ac_prev = self._ac_host[t_prev] if t_prev >= 0 else torch.tensor(1.0)
        return ac_t, ac_prev

    def _predict_coefs(self, t: int):
        ac_t = self._ac_host[t]
        return [(1. / ac_t) ** 0.5, (1. / ac_t - 1.) ** 0.5, ac_t ** 0.5, (1. - ac_t) ** 0.5]

    def _step_coefs(self, t: int, t_prev: int):
        """[x0_coef, xt_coef, eps_coef, var, min_logvar, max_logvar] of the DDPM posterior (ddpm.py:221-246)."""
        ac_t, ac_prev = self._ac_pair(t, t_prev)
        alphas_t = ac_t / ac_prev
        betas_t = 1. - alphas_t
        coef1 = (ac_prev ** 0.5) * betas_t / (1. - ac_t)
        coef2 = (alphas_t ** 0.5) * (1. - ac_prev) / (1. - ac_t)
        small = betas_t * (1. - ac_prev) / (1. - ac_t)
        zero = torch.zeros(())
        if t == 0:
            var = zero
        elif self.var_type == 'fixed_small':
            var = small
        elif self.var_type == 'fixed_large':
            var = betas_t
        else:
            var = zero  # per-pixel, computed in the kernel
        min_logvar = torch.log(torch.clamp_min(small, 1e-20))
        max_logvar = torch.log(betas_t) if float(betas_t) > 0 else zero
        return [coef1, coef2, zero, var, min_logvar, max_logvar]

    def _coef_row(self, t: int, t_prev: int) -> Tensor:
        """Device row of K.SC_COUNT floats for step (t -> t_prev); cached."""
        key = (type(self).__name__, t, t_prev, self.var_type, getattr(self, 'eta', None))
        row = self._coef_rows.get(key)
        if row is None:
            step = self._step_coefs(t, t_prev)
            # the noise term sqrt(var) * z is skipped (and z not read) at t == 0 like the reference (ddpm.py:252,
            # ddim.py:77) and whenever the fixed variance is exactly 0 (DDIM eta = 0: mean + 0 * z == mean), which
            # saves the 4 B/element noise stream; the learned-range variance is per pixel and always reads it
            learned = self.var_type == 'learned_range' and self._uses_learned_var()
            add_noise = t != 0 and (learned or float(step[3]) != 0.0)
            vals = self._predict_coefs(t) + step + [torch.tensor(1.0 if add_noise else 0.0)]
            host = torch.stack([torch.as_tensor(v, dtype=torch.float32).reshape(()) for v in vals] +
                               [torch.zeros(())] * (K.SC_COUNT - len(vals)))
            row = host.to(self.device)
            self._coef_rows[key] = row
        return row

    def _coef_table(self, pairs) -> Tensor:
        return torch.stack([self._coef_row(t, tp) for t, tp in pairs]).contiguous()

    # ------------------------------------------------------------------------------------------
    # reference helpers kept for API compatibility (ddpm.py:102-120)
    # ------------------------------------------------------------------------------------------
    def pred_x0_from_eps(self, xt: Tensor, t: int, eps: Tensor):
        return self._predict_kernel(eps, xt, t, 'pred_eps', clip=False)['pred_x0']

    def pred_eps_from_x0(self, xt: Tensor, t: int, x0: Tensor):
        return self._predict_kernel(x0, xt, t, 'pred_x0', clip=False)['pred_eps']

    def pred_x0_from_v(self, xt: Tensor, t: int, v: Tensor):
        return self._predict_kernel(v, xt, t, 'pred_v', clip=False)['pred_x0']

    def pred_eps_from_v(self, xt: Tensor, t: int, v: Tensor):
        return self._predict_kernel(v, xt, t, 'pred_v', clip=False)['pred_eps']

    def _predict_kernel(self, model_output: Tensor, xt: Tensor, t: int, objective: str, clip: bool):
        xt = xt.contiguous()
        model_output = model_output.contiguous()
        pred_x0, pred_eps = torch.empty_like(xt), torch.empty_like(xt)
        K.sampler_step(model_output, xt, self._coef_row(t, -1), objective=objective, clip=clip,
                       pred_x0=pred_x0, pred_eps=pred_eps)
        return {'pred_x0': pred_x0, 'pred_eps': pred_eps}

    # ------------------------------------------------------------------------------------------
    # training-side entry points (ddpm.py:122-172)
    # ------------------------------------------------------------------------------------------
    def loss_func(self, model: nn.Module, x0: Tensor, t: Tensor, eps: Tensor = None, model_kwargs: Dict = None):
        model_kwargs = dict() if model_kwargs is None else model_kwargs
        eps = torch.randn_like(x0) if eps is None else eps
        xt = self.diffuse(x0, t, eps)
        if self.objective == 'pred_eps':
            target = eps
        elif self.objective == 'pred_x0':
            target = x0
        elif self.objective == 'pred_v':
            target = self.get_v(x0, eps, t)
        else:
            raise ValueError(f'Objective {self.objective} is not supported.')
        pred = model(xt, t, **model_kwargs)
        if not pred.is_cuda:
            raise RuntimeError('loss_func needs a b200diff model on a CUDA device: the training step has no '
                               'PyTorch fallback')
        from models.backward import mse_loss
        return mse_loss(pred, target)

    def get_v(self, x0: Tensor, eps: Tensor, t: Tensor):
        # v = sqrt(ac) eps - sqrt(1-ac) x0 == diffuse(eps, t, -x0)
        return self.diffuse(eps, t, -x0)

    def diffuse(self, x0: Tensor, t: Tensor, eps: Tensor = None):
        """Sample from q(xt | x0); t is a [B] tensor (per-sample timesteps) or an int."""
        eps = torch.randn_like(x0) if eps is None else eps
        if not torch.is_tensor(t):
            if not 0 <= int(t) < self.total_steps:     # the reference's alphas_cumprod[t] raises the same way
                raise IndexError(f'timestep {int(t)} is out of range for {self.total_steps} diffusion steps')
            t = torch.full((x0.shape[0],), int(t), device=x0.device, dtype=torch.long)
        t = t.to(device=x0.device, dtype=torch.long).contiguous()
        x0 = x0.float().contiguous()
        out = torch.empty_like(x0)
        # device-side timesteps are not synchronised for a range check: the kernel never reads alphas_cumprod out of
        # bounds and turns an out-of-range sample into NaN
        K.diffuse(x0, eps.to(device=x0.device, dtype=torch.float32).contiguous(), t, self.alphas_cumprod, out)
        return out

    # ------------------------------------------------------------------------------------------
    # reverse process
    # ------------------------------------------------------------------------------------------
    def predict(self, model_output: Tensor, xt: Tensor, t: int):
        """x0 / eps prediction from the network output (ddpm.py:174-203)."""
        C = xt.shape[1]
        out = self._predict_kernel(model_output, xt, t, self.objective, self.clip_denoised)
        out['learned_var'] = model_output[:, C:] if model_output.shape[1] > C else None
        return out

    def _denoise_impl(self, model_output, xt, t, t_prev, reverse_eps=None, model_output_uncond=None,
                      guidance_scale=1.0, objective=None):
        xt = xt.contiguous()
        model_output = model_output.contiguous()
        learned = self.var_type == 'learned_range' and self._uses_learned_var()
        if reverse_eps is None:
            reverse_eps = torch.randn_like(xt)   # drawn every step, like the reference (ddpm.py:251, ddim.py:76)
        elif reverse_eps.dtype != xt.dtype or not reverse_eps.is_contiguous() or reverse_eps.shape != xt.shape:
            reverse_eps = reverse_eps.to(xt.dtype).expand_as(xt).contiguous()   # the kernel reads raw fp32 pointers
        sample, mean = torch.empty_like(xt), torch.empty_like(xt)
        pred_x0, pred_eps = torch.empty_like(xt), torch.empty_like(xt)
        row = self._coef_row(t, t_prev)
        var = None
        if learned and t != 0:
            var = torch.empty_like(xt)
        K.sampler_step(model_output, xt, row, objective=objective or self.objective, clip=self.clip_denoised,
                       learned_range=learned, noise=reverse_eps,
                       model_out_uncond=None if model_output_uncond is None else model_output_uncond.contiguous(),
                       guidance_scale=guidance_scale, sample=sample, mean=mean, pred_x0=pred_x0, pred_eps=pred_eps,
                       var_out=var)
        if var is None:
            var = row[K.SC_VAR]
        return {'sample': sample, 'mean': mean, 'var': var, 'pred_x0': pred_x0, 'pred_eps': pred_eps,
                'reverse_eps': reverse_eps}

    def _uses_learned_var(self) -> bool:
        return True

    def denoise(self, model_output: Tensor, xt: Tensor, t: int, t_prev: int, reverse_eps: Tensor = None):
        """Sample from p_theta(x{t-1} | xt) (ddpm.py:205-261).  `reverse_eps` optionally injects the noise."""
        return self._denoise_impl(model_output, xt, t, t_prev, reverse_eps)

    def _step_pairs(self):
        sample_seq = self.respaced_seq.tolist()
        sample_seq_prev = [-1] + self.respaced_seq[:-1].tolist()
        return list(zip(reversed(sample_seq), reversed(sample_seq_prev)))

    def sample_loop(self, model: nn.Module, init_noise: Tensor, tqdm_kwargs: Dict = None, model_kwargs: Dict = None):
        tqdm_kwargs = dict() if tqdm_kwargs is None else tqdm_kwargs
        model_kwargs = dict() if model_kwargs is None else model_kwargs
        img = init_noise
        pairs = self._step_pairs()
        pbar = tqdm.tqdm(total=len(pairs), **tqdm_kwargs)
        for t, t_prev in pairs:
            # stride-0 expand: same values as torch.full, and tells the UNet that t is uniform over the batch
            t_batch = torch.full((1, ), t, device=self.device, dtype=torch.long).expand(img.shape[0])
            model_output = model(img, t_batch, **model_kwargs)
            out = self.denoise(model_output, img, t, t_prev)
            img = out['sample']
            pbar.update(1)
            yield out
        pbar.close()

    def sample(self, model: nn.Module, init_noise: Tensor, tqdm_kwargs: Dict = None, model_kwargs: Dict = None):
        runner = _graph_runner(self, model)
        if runner is not None:
            return runner.run(init_noise, tqdm_kwargs, model_kwargs, None, None)
        sample = None
        for out in self.sample_loop(model, init_noise, tqdm_kwargs, model_kwargs):
            sample = out['sample']
        return sample


_STEP_METHODS = ('denoise', 'predict', '_denoise_impl', '_predict_kernel', '_step_coefs', '_predict_coefs', '_coef_row',
                 '_coef_table', '_step_pairs', '_uses_learned_var', 'sample_loop', '_cfg_loop')


def _graph_runner(diffuser, model):
    """CUDA-graph replay of the per-timestep work, when the model is a b200diff engine model AND the diffuser's
    per-step arithmetic is the stock one.  The runner hard-wires `b200_sampler_step` with `_coef_table`, so a subclass
    that overrides any of the step methods (guidance samplers, DDPM-IP-style variants, user subclasses) must not be
    routed through it: `sample()` then walks `sample_loop()` like the reference (ddpm.py:283-290), override included."""
    make = getattr(model, 'make_sampling_runner', None)
    if make is None or not getattr(model, 'use_cuda_graph', True):
        return None
    from diffusions.ddim import DDIM, DDIMCFG
    cls = type(diffuser)
    if cls not in (DDPM, DDPMCFG, DDIM, DDIMCFG):
        stock = next(b for b in cls.__mro__ if b in (DDIMCFG, DDPMCFG, DDIM, DDPM))
        if any(getattr(cls, m, None) is not getattr(stock, m, None) for m in _STEP_METHODS):
            return None
    return make(diffuser)


class _CFGMixin:
    """Classifier-free guidance loop shared by DDPMCFG / DDIMCFG (ddpm.py:319-351, ddim.py:161-191)."""
    guidance_scale: float
    cond_kwarg: str

    def _cfg_loop(self, model, init_noise, uncond_conditioning, tqdm_kwargs, model_kwargs):
        tqdm_kwargs = dict() if tqdm_kwargs is None else tqdm_kwargs
        if self.cond_kwarg not in model_kwargs.keys():
            raise ValueError(f'Condition argument `{self.cond_kwarg}` not found in model_kwargs.')
        uncond_model_kwargs = model_kwargs.copy()
        uncond_model_kwargs[self.cond_kwarg] = uncond_conditioning

        img = init_noise
        pairs = self._step_pairs()
        pbar = tqdm.tqdm(total=len(pairs), **tqdm_kwargs)
        for t, t_prev in pairs:
            t_batch = torch.full((1, ), t, device=self.device).expand(img.shape[0])
            model_output_cond = model(img, t_batch, **model_kwargs)
            model_output_uncond = model(img, t_batch, **uncond_model_kwargs)
            # per-branch predict+clip, the (1-s)/s mix and the final predict+clip+step are fused in one kernel
            out = self._denoise_impl(model_output_cond, img, t, t_prev, None, model_output_uncond,
                                     self.guidance_scale)
            img = out['sample']
            pbar.update(1)
            yield out
        pbar.close()

    @contextmanager
    def hack_objective(self, objective: str):
        """Hack objective temporarily."""
        tmp = self.objective
        self.objective = objective
        yield
        self.objective = tmp


class DDPMCFG(_CFGMixin, DDPM):
    def __init__(self, guidance_scale: float = 1., cond_kwarg: str = 'y', *args, **kwargs):
        """DDPM with classifier-free guidance; `guidance_scale` follows the classifier-guidance convention
        (0 = unconditional, 1 = conditional, >1 = guided), as in the reference (ddpm.py:294-317)."""
        DDPM.__init__(self, *args, **kwargs)
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
