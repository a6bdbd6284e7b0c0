"""Fused Adam / AdamW (+ gradient-norm clip + EMA) on the b200_optimizer_step kernels.

Drop-in for the `torch.optim.Adam` / `AdamW` the reference's training scripts build from `conf.train.optim`
(scripts/train_ddpm.py:128-131): same constructor arguments, `param_groups`, `state_dict()` layout ('step', 'exp_avg',
'exp_avg_sq' per parameter) and update rule.  `step()` alone replaces `optimizer.step()`; passing `clip_grad_norm=` and
`ema=` additionally fuses `accelerator.clip_grad_norm_` and `ema.update(model.parameters())` (train_ddpm.py:186-188)
into the same two kernels, with no host synchronisation (the clip coefficient is computed on the device).

`capturable=True` keeps the step count, bias corrections, learning rate and (gradual) EMA decay in device memory, so
that a whole training step can be captured in a CUDA graph and replayed (b200diff.train.TrainStep(use_cuda_graph=True)).
"""
import ctypes
import struct

import torch

from . import OptimDesc, _check, _stream, lib

_CHUNK = 65536


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, adamw=False, capturable=False):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError('FusedAdam: invalid hyper-parameter')
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, adamw=adamw))
        self.capturable = capturable
        self._tables = {}
        self._norm = None
        self._dev_state = None       # b200_optim_dev_state (capturable mode)
        self._dev_lr = None
        self.grad_norm = None
        if capturable and len(self.param_groups) != 1:
            raise ValueError('FusedAdam(capturable=True) supports a single param group')

    def _table(self, gi, params, ema_shadow):
        """Device chunk table of one param group; rebuilt only when a tensor moved (new .grad buffers, new shadow)."""
        sig = tuple((p.data_ptr(), p.grad.data_ptr(), self.state[p]['exp_avg'].data_ptr(),
                     self.state[p]['exp_avg_sq'].data_ptr()) for p in params) + \
            (tuple(s.data_ptr() for s in ema_shadow) if ema_shadow is not None else ())
        cached = self._tables.get(gi)
        if cached is not None and cached[0] == sig:
            return cached[1], cached[2]
        if torch.cuda.is_current_stream_capturing():
            raise RuntimeError('FusedAdam: parameter / gradient storage changed during CUDA-graph capture '
                               '(run warm-up steps with the same gradient buffers first)')
        rows = []
        for i, p in enumerate(params):
            st = self.state[p]
            e = ema_shadow[i] if ema_shadow is not None else None
            n = p.numel()
            for o in range(0, n, _CHUNK):
                rows.append(struct.pack('<5Q2i', p.data_ptr() + 4 * o, p.grad.data_ptr() + 4 * o,
                                        st['exp_avg'].data_ptr() + 4 * o, st['exp_avg_sq'].data_ptr() + 4 * o,
                                        0 if e is None else e.data_ptr() + 4 * o, min(_CHUNK, n - o), 0))
        blob = torch.frombuffer(bytearray(b''.join(rows)), dtype=torch.uint8).to(params[0].device)
        self._tables[gi] = (sig, blob, len(rows))
        return blob, len(rows)

    def _device_state(self, device, step0, lr, ema, ema_updates0):
        """(Re)creates the device-resident scalar state from the host-side PRE-step counts; lr is refreshed on every call."""
        if self._dev_state is None or self._dev_state.device != device:
            host = struct.pack('<2i5fi', int(step0), int(ema_updates0), float(lr), 0.0, 0.0, 0.0,
                               0.0 if ema is None else float(ema.decay), 0 if ema is None else int(bool(ema.gradual)))
            self._dev_state = torch.frombuffer(bytearray(host), dtype=torch.uint8).to(device)
            self._dev_lr = self._dev_state[8:12].view(torch.float32)
            self._lr_host = float(lr)
        elif float(lr) != self._lr_host and not torch.cuda.is_current_stream_capturing():
            self._dev_lr.fill_(float(lr))
            self._lr_host = float(lr)
        return self._dev_state

    def load_state_dict(self, state_dict):
        """torch.optim.Optimizer.load_state_dict + drop everything derived from the old state tensors: the device chunk
        tables (they hold raw exp_avg / exp_avg_sq pointers) and the device-resident step / EMA counters (capturable
        mode re-creates them from the loaded host-side step counts at the next step)."""
        super().load_state_dict(state_dict)
        self._tables.clear()
        self._dev_state = None
        self._dev_lr = None
        for st in self.state.values():      # torch moves 'step' to the parameter's device; this class keeps it on the host
            if torch.is_tensor(st.get('step')) and st['step'].is_cuda:
                st['step'] = st['step'].cpu()

    def reset_device_state(self):
        """Call after `ema.load_state_dict(...)` in capturable mode: the EMA update count lives in device memory."""
        self._tables.clear()
        self._dev_state = None
        self._dev_lr = None

    def set_lr(self, lr: float):
        """Learning-rate schedulers: updates param_groups and (capturable mode) the device copy read by graph replays."""
        for g in self.param_groups:
            g['lr'] = lr
        if self._dev_state is not None:
            self._dev_lr.fill_(float(lr))
            self._lr_host = float(lr)

    def mirror_replayed_step(self, ema=None):
        """Host-side bookkeeping after a CUDA-graph replay of `step` (the device advanced its own counters): step counts
        in `state`, EMA update count, and the parameters' version counters (so packed-weight caches notice)."""
        params = [p for g in self.param_groups for p in g['params'] if p.grad is not None]
        for p in params:
            self.state[p]['step'] += 1
        if ema is not None:
            ema.num_updates += 1
        torch.autograd.graph.increment_version(params)

    @torch.no_grad()
    def step(self, closure=None, *, clip_grad_norm=None, ema=None):
        """Returns the closure's loss (torch convention).  `self.grad_norm` (a device scalar) holds the global gradient
        norm of the last step when `clip_grad_norm` was given."""
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        ema_decay = -1.0
        ema_updates0 = 0
        if ema is not None:
            ema_updates0 = ema.num_updates
            ema.num_updates += 1
            ema_decay = float(ema.get_decay())
        self.grad_norm = None
        shadow_off = 0
        n_groups = len(self.param_groups)
        for gi, group in enumerate(self.param_groups):
            all_params = group['params']
            shadow = None
            if ema is not None:
                shadow_all = ema.shadow[shadow_off:shadow_off + len(all_params)]
                shadow_off += len(all_params)
            idx = [i for i, p in enumerate(all_params) if p.grad is not None]
            if ema is not None:
                # models/ema.py:31-38: parameters that are frozen (or received no gradient) are copied, not averaged
                for i, p in enumerate(all_params):
                    if p.grad is None:
                        shadow_all[i].copy_(p)
            if not idx:
                continue
            params = [all_params[i] for i in idx]
            if ema is not None:
                shadow = [shadow_all[i] for i in idx]
            step_before = 0
            for p in params:
                if not p.is_cuda or p.dtype != torch.float32 or not p.is_contiguous() or not p.grad.is_contiguous():
                    raise RuntimeError('FusedAdam: parameters and gradients must be contiguous float32 CUDA tensors')
                st = self.state[p]
                if not st:
                    st['step'] = torch.zeros((), dtype=torch.float32)
                    st['exp_avg'] = torch.zeros_like(p)
                    st['exp_avg_sq'] = torch.zeros_like(p)
                step_before = int(st['step'])
                st['step'] += 1
            table, n_chunks = self._table(gi, params, shadow)
            if self._norm is None or self._norm.device != params[0].device:
                self._norm = torch.zeros(1, dtype=torch.float32, device=params[0].device)
            d = OptimDesc()
            d.chunks, d.n_chunks = table.data_ptr(), n_chunks
            d.lr, (d.beta1, d.beta2), d.eps = float(group['lr']), group['betas'], float(group['eps'])
            d.weight_decay, d.adamw = float(group['weight_decay']), int(bool(group.get('adamw', False)))
            d.step = step_before + 1
            d.max_grad_norm = float(clip_grad_norm) if clip_grad_norm is not None else 0.0
            d.want_norm = 0
            d.gnorm_sq = self._norm.data_ptr()
            d.ema_decay = ema_decay
            if self.capturable:
                dev = self._device_state(params[0].device, step_before, group['lr'], ema, ema_updates0)
                d.dev_state = dev.data_ptr()
                d.ema_decay = 0.0 if ema is not None else -1.0      # only the flag; the value is derived on the device
            if clip_grad_norm is not None and n_groups > 1:
                raise RuntimeError('FusedAdam: fused clipping supports a single param group (global norm)')
            _check(lib().b200_optimizer_step(ctypes.byref(d), _stream()), 'optimizer_step')
            # the kernels wrote the parameters through raw pointers: bump their version counters so that consumers
            # keyed on tensor versions (the engine's packed bf16 weights) see the update
            torch.autograd.graph.increment_version(params)
            if clip_grad_norm is not None:
                self.grad_norm = self._norm.sqrt()
        return loss


class FusedAdamW(FusedAdam):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, capturable=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, adamw=True, capturable=capturable)
