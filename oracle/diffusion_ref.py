"""ORACLE (test infrastructure, not a product path): CPU restatement of the reference's diffusion arithmetic.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
It restates, in plain PyTorch fp32 on the CPU, what the reference computes in
  diffusions/schedule.py:5-73      (beta schedules, respaced timestep sequences)
  diffusions/ddpm.py:72-93         (alphas_cumprod in float64 -> float32)
  diffusions/ddpm.py:102-120       (x0 / eps / v conversions)
  diffusions/ddpm.py:152-172       (diffuse)
  diffusions/ddpm.py:174-261       (predict + DDPM posterior step, three variance types)
  diffusions/ddim.py:57-86         (DDIM step)
  diffusions/ddim.py:88-132        (DDIM inversion step and loop), :202-242 (inversion under classifier-free guidance)
  diffusions/ddpm.py:263-351, diffusions/ddim.py:161-200  (sampling loops incl. classifier-free guidance)
with the random draws made injectable so that two implementations can consume identical noise.

Pinning: the reference ships no golden vectors for this path ("parity unpinned" by the
reference's own tests).  This restatement is pinned instead against the reference itself, imported live from
/root/reference by oracle/gen_golden.py, which also froze tests/golden/*.pt; tests/test_oracle.py re-checks
the restatement against those fixtures and against the known answers listed in SURVEY.md section 8c.
"""
import math

import torch


# ------------------------------------------------------------------------------------------------
# schedules  (diffusions/schedule.py:5-73)
# ------------------------------------------------------------------------------------------------
def beta_schedule(total_steps=1000, kind='linear', beta_start=1e-4, beta_end=0.02):
    if kind == 'linear':
        return torch.linspace(beta_start, beta_end, total_steps, dtype=torch.float64)
    if kind == 'quad':
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, total_steps, dtype=torch.float64) ** 2
    if kind == 'const':
        return torch.full((total_steps,), beta_end, dtype=torch.float64)
    if kind == 'cosine':
        f = lambda u: math.cos((u + 0.008) / 1.008 * math.pi / 2) ** 2  # noqa: E731
        vals = [min(1 - f((i + 1) / total_steps) / f(i / total_steps), 0.999) for i in range(total_steps)]
        return torch.tensor(vals)  # float32, like the reference
    raise ValueError(f'Beta schedule {kind} is not supported.')


def respaced_seq(total_steps=1000, kind='uniform', steps=100):
    if kind in ('uniform', 'uniform-leading'):
        return torch.arange(0, total_steps, total_steps // steps).long()
    if kind == 'uniform-linspace':
        return torch.linspace(0, total_steps - 1, steps).long()
    if kind == 'uniform-trailing':
        return torch.arange(total_steps - 1, -1, -(total_steps // steps)).long().flip(dims=[0])
    if kind == 'quad':
        return torch.floor(torch.linspace(0, math.sqrt(total_steps * 0.8), steps) ** 2).long()
    if kind is None or kind == 'none':
        return torch.arange(0, total_steps).long()
    raise ValueError(f'Respace type {kind} is not supported.')


# ------------------------------------------------------------------------------------------------
# DDPM / DDIM
# ------------------------------------------------------------------------------------------------
class DDPMRef:
    """Restates diffusions/ddpm.py:13-290 on CPU tensors."""

    def __init__(self, total_steps=1000, beta_schedule_kind='linear', beta_start=1e-4, beta_end=0.02, betas=None,
                 objective='pred_eps', var_type='fixed_large', clip_denoised=True, respace_type=None,
                 respace_steps=100, respaced=None, beta_schedule=None):
        if beta_schedule is not None:
            beta_schedule_kind = beta_schedule
        if objective not in ('pred_eps', 'pred_x0', 'pred_v'):
            raise ValueError(f'Invalid objective: {objective}')
        if var_type not in ('fixed_small', 'fixed_large', 'learned_range'):
            raise ValueError(f'Invalid var_type: {var_type}')
        self.total_steps, self.objective, self.var_type, self.clip_denoised = total_steps, objective, var_type, clip_denoised
        if betas is None:
            betas = globals()['beta_schedule'](total_steps, beta_schedule_kind, beta_start, beta_end)
        self.alphas_cumprod = torch.cumprod(1. - betas, dim=0).to(torch.float)
        self.respaced_seq = respaced if respaced is not None else respaced_seq(total_steps, respace_type, respace_steps)

    # ---- conversions (ddpm.py:102-120) ----
    def _x0_from_eps(self, xt, t, eps):
        ac = self.alphas_cumprod[t]
        return (1. / ac) ** 0.5 * xt - (1. / ac - 1.) ** 0.5 * eps

    def _eps_from_x0(self, xt, t, x0):
        ac = self.alphas_cumprod[t]
        return ((1. / ac) ** 0.5 * xt - x0) / (1. / ac - 1.) ** 0.5

    def _x0_from_v(self, xt, t, v):
        ac = self.alphas_cumprod[t]
        return ac ** 0.5 * xt - (1. - ac) ** 0.5 * v

    @staticmethod
    def _bcast(coef, like):
        while coef.ndim < like.ndim:
            coef = coef.unsqueeze(-1)
        return coef

    def get_v(self, x0, eps, t):
        ac = self.alphas_cumprod[t]
        return self._bcast(ac ** 0.5, x0) * eps - self._bcast((1. - ac) ** 0.5, x0) * x0

    def diffuse(self, x0, t, eps):
        ac = self.alphas_cumprod[t]
        return self._bcast(ac ** 0.5, x0) * x0 + self._bcast((1. - ac) ** 0.5, x0) * eps

    def loss(self, model, x0, t, eps, model_kwargs=None):
        """ddpm.py:122-138 with the noise passed in."""
        kw = model_kwargs or {}
        out = model(self.diffuse(x0, t, eps), t, **kw)
        target = {'pred_eps': eps, 'pred_x0': x0, 'pred_v': None}[self.objective]
        if target is None:
            target = self.get_v(x0, eps, t)
        return torch.nn.functional.mse_loss(out, target)

    # ---- predict (ddpm.py:174-203) ----
    def predict(self, model_output, xt, t):
        learned = None
        C = xt.shape[1]
        if model_output.shape[1] > C:
            model_output, learned = model_output[:, :C], model_output[:, C:]
        if self.objective == 'pred_eps':
            x0 = self._x0_from_eps(xt, t, model_output)
        elif self.objective == 'pred_x0':
            x0 = model_output.clone()
        else:
            x0 = self._x0_from_v(xt, t, model_output)
        if self.clip_denoised:
            x0 = x0.clamp(-1., 1.)
        return {'pred_x0': x0, 'pred_eps': self._eps_from_x0(xt, t, x0), 'learned_var': learned}

    def _ac_pair(self, t, t_prev):
        return self.alphas_cumprod[t], (self.alphas_cumprod[t_prev] if t_prev >= 0 else torch.tensor(1.0))

    # ---- posterior step (ddpm.py:205-261) ----
    def denoise(self, model_output, xt, t, t_prev, reverse_eps=None):
        pr = self.predict(model_output, xt, t)
        x0, eps, learned = pr['pred_x0'], pr['pred_eps'], pr['learned_var']
        ac_t, ac_p = self._ac_pair(t, t_prev)
        alpha_t = ac_t / ac_p
        beta_t = 1. - alpha_t
        mean = (ac_p ** 0.5) * beta_t / (1. - ac_t) * x0 + (alpha_t ** 0.5) * (1. - ac_p) / (1. - ac_t) * xt
        if t == 0:
            var = torch.zeros_like(beta_t)
        elif self.var_type == 'fixed_small':
            var = beta_t * (1. - ac_p) / (1. - ac_t)
        elif self.var_type == 'fixed_large':
            var = beta_t
        else:
            lo = torch.log(torch.clamp_min(beta_t * (1. - ac_p) / (1. - ac_t), 1e-20))
            hi = torch.log(beta_t)
            frac = (learned + 1) / 2
            var = torch.exp(frac * hi + (1 - frac) * lo)
        if reverse_eps is None:
            reverse_eps = torch.randn_like(xt)
        sample = mean if t == 0 else mean + torch.sqrt(var) * reverse_eps
        return {'sample': sample, 'mean': mean, 'var': var, 'pred_x0': x0, 'pred_eps': eps, 'reverse_eps': reverse_eps}

    # ---- loops (ddpm.py:263-290) ----
    def _pairs(self):
        seq = self.respaced_seq.tolist()
        prev = [-1] + seq[:-1]
        return list(zip(reversed(seq), reversed(prev)))

    def sample_loop(self, model, init_noise, noises=None, model_kwargs=None):
        kw = model_kwargs or {}
        img = init_noise
        for i, (t, tp) in enumerate(self._pairs()):
            tb = torch.full((img.shape[0],), t, dtype=torch.long, device=img.device)
            out = self.denoise(model(img, tb, **kw), img, t, tp, None if noises is None else noises[i])
            img = out['sample']
            yield out

    def sample(self, model, init_noise, noises=None, model_kwargs=None):
        out = None
        for out in self.sample_loop(model, init_noise, noises, model_kwargs):
            pass
        return out['sample']

    # ---- classifier-free guidance (ddpm.py:319-351 / ddim.py:161-191) ----
    def sample_loop_cfg(self, model, init_noise, guidance_scale, cond_kwargs, uncond_kwargs, noises=None):
        img = init_noise
        C = init_noise.shape[1]
        for i, (t, tp) in enumerate(self._pairs()):
            tb = torch.full((img.shape[0],), t, dtype=torch.long, device=img.device)
            out_c = model(img, tb, **cond_kwargs)
            eps_c = self.predict(out_c, img, t)['pred_eps']
            eps_u = self.predict(model(img, tb, **uncond_kwargs), img, t)['pred_eps']
            mix = (1 - guidance_scale) * eps_u + guidance_scale * eps_c
            if self.var_type == 'learned_range' and type(self) is DDPMRef:
                mix = torch.cat([mix, out_c[:, C:]], dim=1)
            keep, self.objective = self.objective, 'pred_eps'
            try:
                out = self.denoise(mix, img, t, tp, None if noises is None else noises[i])
            finally:
                self.objective = keep
            img = out['sample']
            yield out


class DDIMRef(DDPMRef):
    """Restates diffusions/ddim.py:12-86."""

    def __init__(self, eta=0., **kw):
        super().__init__(**kw)
        self.eta = eta

    def denoise(self, model_output, xt, t, t_prev, reverse_eps=None):
        pr = self.predict(model_output, xt, t)
        x0, eps = pr['pred_x0'], pr['pred_eps']
        ac_t, ac_p = self._ac_pair(t, t_prev)
        var = (self.eta ** 2) * (1. - ac_p) / (1. - ac_t) * (1. - ac_t / ac_p)
        mean = torch.sqrt(ac_p) * x0 + torch.sqrt(1. - ac_p - var) * eps
        if reverse_eps is None:
            reverse_eps = torch.randn_like(xt)
        sample = mean if t == 0 else mean + torch.sqrt(var) * reverse_eps
        return {'sample': sample, 'mean': mean, 'var': var, 'pred_x0': x0, 'pred_eps': eps, 'reverse_eps': reverse_eps}

    # ---- DDIM inversion (ddim.py:88-132): the deterministic update run towards higher noise levels ----
    def denoise_inversion(self, model_output, xt, t, t_next):
        """x_t -> x_{t_next}: sqrt(ac_next) x0 + sqrt(1 - ac_next) eps; ac_next = 0 past the last step (ddim.py:99)."""
        if self.eta != 0.:
            raise ValueError(f'DDIM inversion is only valid when eta=0, get {self.eta}')
        pr = self.predict(model_output, xt, t)
        x0, eps = pr['pred_x0'], pr['pred_eps']
        ac_next = self.alphas_cumprod[t_next] if t_next < self.total_steps else torch.tensor(0.0)
        sample = torch.sqrt(ac_next) * x0 + torch.sqrt(1. - ac_next) * eps
        return {'sample': sample, 'pred_x0': x0, 'pred_eps': eps}

    def _inversion_pairs(self):
        seq = self.respaced_seq.tolist()
        return list(zip(seq[:-1], seq[1:]))      # ddim.py:113-114: the last respaced step is only ever a target

    def sample_inversion_loop(self, model, img, model_kwargs=None):
        kw = model_kwargs or {}
        for (t, tn) in self._inversion_pairs():
            tb = torch.full((img.shape[0],), t, dtype=torch.long, device=img.device)
            out = self.denoise_inversion(model(img, tb, **kw), img, t, tn)
            img = out['sample']
            yield out

    def sample_inversion(self, model, img, model_kwargs=None):
        out = None
        for out in self.sample_inversion_loop(model, img, model_kwargs):
            pass
        return out['sample']

    def sample_inversion_loop_cfg(self, model, img, guidance_scale, cond_kwargs, uncond_kwargs):
        """ddim.py:202-232: per-branch predict (with clip), the (1 - s) / s mix, then the inversion step on the mix."""
        for (t, tn) in self._inversion_pairs():
            tb = torch.full((img.shape[0],), t, dtype=torch.long, device=img.device)
            eps_c = self.predict(model(img, tb, **cond_kwargs), img, t)['pred_eps']
            eps_u = self.predict(model(img, tb, **uncond_kwargs), img, t)['pred_eps']
            mix = (1 - guidance_scale) * eps_u + guidance_scale * eps_c
            keep, self.objective = self.objective, 'pred_eps'
            try:
                out = self.denoise_inversion(mix, img, t, tn)
            finally:
                self.objective = keep
            img = out['sample']
            yield out


# ------------------------------------------------------------------------------------------------------
# Euler / Heun samplers (diffusions/euler.py:46-66, diffusions/heun.py:48-131)
# ------------------------------------------------------------------------------------------------------
class EulerRef(DDPMRef):
    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.sigmas = ((1 - self.alphas_cumprod) / self.alphas_cumprod).sqrt()

    def _sig(self, t, t_prev):
        return self.sigmas[t], (self.sigmas[t_prev] if t_prev >= 0 else torch.tensor(0.0))

    def denoise(self, model_output, xt, t, t_prev, reverse_eps=None):
        st, sp = self._sig(t, t_prev)
        x0 = self.predict(model_output, xt, t)['pred_x0']
        bar_xt = (1 + st ** 2).sqrt() * xt
        derivative = (bar_xt - x0) / st
        bar_sample = bar_xt + derivative * (sp - st)
        return {'sample': bar_sample / (1 + sp ** 2).sqrt(), 'pred_x0': x0, 'derivative': derivative}


class HeunRef(EulerRef):
    def sample_loop(self, model, init_noise, noises=None, model_kwargs=None):
        kw = model_kwargs or {}
        img = init_noise
        for (t, tp) in self._pairs():
            tb = torch.full((img.shape[0],), t, dtype=torch.long, device=img.device)
            first = EulerRef.denoise(self, model(img, tb, **kw), img, t, tp)
            out, x_first, img = first, img, first['sample']
            if tp >= 0:
                st, sp = self._sig(t, tp)
                tpb = torch.full((img.shape[0],), tp, dtype=torch.long, device=img.device)
                x0 = self.predict(model(img, tpb, **kw), img, tp)['pred_x0']
                bar_prev = (1 + sp ** 2).sqrt() * img
                derivative = ((bar_prev - x0) / sp + first['derivative']) / 2
                bar_xt = (1 + st ** 2).sqrt() * x_first
                out = {'sample': (bar_xt + derivative * (sp - st)) / (1 + sp ** 2).sqrt(), 'pred_x0': x0}
                img = out['sample']
            yield out
