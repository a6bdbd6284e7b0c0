"""One training step of the noise-prediction objective on the kernels -- the body of `run_step` in the reference's
scripts/train_ddpm.py:171-192 / train_ddpm_cfg.py:172-196:

    optimizer.zero_grad(); t ~ U{0..T-1}; loss = diffuser.loss_func(model, x0, t); loss.backward();
    [all-reduce(mean) of the gradients over the data-parallel ranks, in buckets overlapped with the backward]; clip_grad_norm_(1.0); optimizer.step(); ema.update()

with the last three fused into b200_optimizer_step, and no host synchronisation anywhere (the loss is returned as a
device scalar; call `.item()` on it only when it is logged).

`use_cuda_graph=True` (FusedAdam(capturable=True); with world_size > 1 the gradient all-reduce is captured too): after two eager warm-up steps the whole step -- timestep and
noise draws, diffuse, UNet forward with fresh dropout masks, MSE, every backward kernel, clip + Adam + EMA, and the bf16
re-pack of the weights -- is captured once per (batch shape, conditional / unconditional) and replayed: the ~830
kernel launches of a step then cost one graph launch on the host instead of ~25 ms of Python.
"""
from typing import Dict, Optional

import torch
import torch.distributed as dist

import b200diff as K

from .dist import allreduce_grads_
from .optim import FusedAdam


class TrainStep:
    def __init__(self, model, diffuser, optimizer: FusedAdam, ema=None, clip_grad_norm: Optional[float] = 1.0,
                 p_uncond: float = 0.0, use_cuda_graph: bool = False, bucket_bytes: int = 32 << 20):
        self.model, self.diffuser, self.optimizer, self.ema = model, diffuser, optimizer, ema
        self.clip_grad_norm = clip_grad_norm
        self.p_uncond = p_uncond      # train_ddpm_cfg.py:183-186: the label is dropped with this probability
        self.use_cuda_graph = use_cuda_graph
        self.bucket_bytes = bucket_bytes     # size of the gradient buckets all-reduced while the backward still runs
        if use_cuda_graph and not getattr(optimizer, 'capturable', False):
            raise ValueError('TrainStep(use_cuda_graph=True) needs FusedAdam(capturable=True)')
        self._graphs: Dict = {}
        self._warm: Dict = {}

    # ------------------------------------------------------------------------------------------
    def _body(self, x0, t, y, eps, micro_batch):
        B = x0.shape[0]
        micro_batch = B if micro_batch is None else micro_batch
        self.optimizer.zero_grad(set_to_none=True)
        total = None
        # data-parallel: average the gradients INSIDE the backward, in buckets of finished weight gradients that overlap
        # the remaining backward kernels (models/backward.py: _OverlappedAllReduce); with several micro-batches per step
        # the single flat all-reduce after the last backward is used instead (B200_DDP_OVERLAP=0 forces that path)
        eng = getattr(getattr(self.model, 'module', self.model), 'engine', None)
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        overlap = (eng is not None and world > 1 and micro_batch >= B
                   and __import__('os').environ.get('B200_DDP_OVERLAP', '1') != '0')
        if eng is not None:
            eng.ddp_overlap_bytes = self.bucket_bytes if overlap else 0
            eng.grads_reduced = False
        for i in range(0, B, micro_batch):
            xs = x0[i:i + micro_batch].float()
            ts = t[i:i + micro_batch] if t is not None else \
                torch.randint(self.diffuser.total_steps, (xs.shape[0],), device=xs.device).long()
            kw: Dict = {}
            if y is not None:
                kw = dict(y=y[i:i + micro_batch])
            loss = self.diffuser.loss_func(self.model, x0=xs, t=ts, eps=None if eps is None else eps[i:i + micro_batch],
                                           model_kwargs=kw)
            scale = xs.shape[0] / B
            (loss * scale).backward()
            total = loss.detach() * scale if total is None else total + loss.detach() * scale
        if eng is not None:
            eng.ddp_overlap_bytes = 0
        if not (eng is not None and eng.grads_reduced):
            allreduce_grads_(self.model)
        self.optimizer.step(clip_grad_norm=self.clip_grad_norm, ema=self.ema)
        return total

    def __call__(self, x0: torch.Tensor, t: torch.Tensor = None, y: torch.Tensor = None, eps: torch.Tensor = None,
                 micro_batch: int = None) -> torch.Tensor:
        if y is not None and self.p_uncond > 0 and float(torch.rand(())) < self.p_uncond:
            y = None       # classifier-free guidance training: the whole batch is unconditional with probability p_uncond
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        # data-parallel runs capture the NCCL all-reduce of the flat gradient buffer inside the same graph
        # (B200_TRAIN_GRAPH_DDP=0 falls back to eager launches when world > 1)
        ddp_ok = world == 1 or __import__('os').environ.get('B200_TRAIN_GRAPH_DDP', '1') != '0'
        graphable = (self.use_cuda_graph and ddp_ok and t is None and eps is None
                     and (micro_batch is None or micro_batch >= x0.shape[0]))
        if not graphable:
            return self._body(x0, t, y, eps, micro_batch)
        key = (tuple(x0.shape), x0.dtype, None if y is None else tuple(y.shape))
        g = self._graphs.get(key)
        if g is None:
            if self._warm.get(key, 0) < 2:     # eager warm-up: fills the arena, weight / optimizer tables, .grad storage
                self._warm[key] = self._warm.get(key, 0) + 1
                return self._body(x0, None, y, None, None)
            g = self._capture(x0, y)
            self._graphs[key] = g
        g['x0'].copy_(x0)
        if y is not None:
            g['y'].copy_(y)
        lr = self.optimizer.param_groups[0]['lr']
        if lr != self.optimizer._lr_host:
            self.optimizer.set_lr(lr)
        g['graph'].replay()
        K.GRAPH_LAUNCHES += g['kernels']
        self.optimizer.mirror_replayed_step(self.ema)
        return g['loss'].clone()

    def close(self):
        """Drops the captured graphs.  Call it (then `torch.cuda.synchronize()`) before `dist.destroy_process_group()`:
        graphs that captured NCCL kernels must not outlive the communicator."""
        self._graphs.clear()
        self._warm.clear()

    def warmup(self, x0, y=None):
        """Runs (and, in CUDA-graph mode, captures) every variant of the step once: conditional and -- when labels may be
        dropped -- unconditional.  These are real optimizer steps."""
        p, self.p_uncond = self.p_uncond, 0.0
        try:
            variants = [y] + ([None] if (y is not None and p > 0) else [])
            for yy in variants:
                for _ in range(3):
                    self(x0, y=yy)
        finally:
            self.p_uncond = p

    def _capture(self, x0, y):
        st = {'x0': x0.clone(), 'y': None if y is None else y.clone()}
        # the captured forward must contain the bf16 re-pack of the weights (every replay follows an optimizer step):
        # make sure the engine sees "parameters changed" while capturing
        torch.autograd.graph.increment_version([p for g_ in self.optimizer.param_groups for p in g_['params']])
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        n0 = K.direct_launch_count()
        # the host-side counters advance once here (capture executes the Python body once); the captured kernels are
        # not executed during capture, so undo that bookkeeping afterwards
        world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
        # thread-local capture mode: the NCCL watchdog thread of torch.distributed keeps querying events while we capture
        with torch.cuda.graph(graph, capture_error_mode='thread_local' if world > 1 else 'global'):
            st['loss'] = self._body(st['x0'], None, st['y'], None, None)
        params = [p for g_ in self.optimizer.param_groups for p in g_['params'] if p.grad is not None]
        for p in params:
            self.optimizer.state[p]['step'] -= 1
        if self.ema is not None:
            self.ema.num_updates -= 1
        st['graph'] = graph
        st['kernels'] = K.direct_launch_count() - n0      # libb200diff kernels replayed per step
        return st
