"""Minimal `accelerate` stand-in on top of torch.distributed: the Accelerator surface the reference's scripts touch
(scripts/train_ddpm.py:36-192, scripts/sample_uncond.py:118-200): one process per GPU launched by torchrun
(RANK / LOCAL_RANK / WORLD_SIZE), NCCL on GPUs, gloo on CPU."""
import contextlib
import os

import torch
import torch.distributed as dist

from . import utils  # noqa: F401
from .utils import DistributedDataParallelKwargs, DistributedType, set_seed  # noqa: F401

__all__ = ['Accelerator', 'DistributedDataParallelKwargs']


class Accelerator:
    def __init__(self, kwargs_handlers=None, mixed_precision='no', gradient_accumulation_steps=1, **_):
        self.mixed_precision = mixed_precision or 'no'
        self.ddp_kwargs = {}
        for h in kwargs_handlers or []:
            if isinstance(h, DistributedDataParallelKwargs):
                self.ddp_kwargs = h.to_kwargs()
        world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local_process_index = int(os.environ.get('LOCAL_RANK', '0'))
        if torch.cuda.is_available():
            torch.cuda.set_device(self.local_process_index)
            self.device = torch.device('cuda', self.local_process_index)
        else:
            self.device = torch.device('cpu')
        if world > 1 and not dist.is_initialized():
            dist.init_process_group('nccl' if self.device.type == 'cuda' else 'gloo')
        self.num_processes = dist.get_world_size() if dist.is_initialized() else 1
        self.process_index = dist.get_rank() if dist.is_initialized() else 0
        self.distributed_type = DistributedType.MULTI_GPU if self.num_processes > 1 else DistributedType.NO

    # ---- topology ----
    @property
    def is_main_process(self):
        return self.process_index == 0

    @property
    def is_local_main_process(self):
        return self.local_process_index == 0

    def wait_for_everyone(self):
        if self.num_processes > 1:
            dist.barrier()

    def on_main_process(self, fn):
        def wrapper(*a, **k):
            if self.is_main_process:
                return fn(*a, **k)
        return wrapper

    def print(self, *a, **k):
        if self.is_main_process:
            print(*a, **k)

    # ---- objects ----
    def prepare(self, *objs):
        out = []
        for o in objs:
            if isinstance(o, torch.nn.Module):
                o = o.to(self.device)
                if self.num_processes > 1:
                    ids = [self.local_process_index] if self.device.type == 'cuda' else None
                    o = torch.nn.parallel.DistributedDataParallel(o, device_ids=ids, **self.ddp_kwargs)
            out.append(o)       # optimizers / dataloaders pass through (every rank draws its own batches)
        return out[0] if len(out) == 1 else tuple(out)

    def unwrap_model(self, model):
        return model.module if isinstance(model, torch.nn.parallel.DistributedDataParallel) else model

    def no_sync(self, model):
        return model.no_sync() if hasattr(model, 'no_sync') else contextlib.nullcontext()

    @contextlib.contextmanager
    def accumulate(self, *models):
        yield

    # ---- step ----
    def backward(self, loss, **kwargs):
        loss.backward(**kwargs)

    def clip_grad_norm_(self, parameters, max_norm, norm_type=2):
        return torch.nn.utils.clip_grad_norm_(parameters, max_norm, norm_type=norm_type)

    # ---- collectives ----
    def gather(self, tensor):
        if self.num_processes == 1:
            return tensor
        parts = [torch.empty_like(tensor) for _ in range(self.num_processes)]
        dist.all_gather(parts, tensor.contiguous())
        return torch.cat(parts, dim=0)

    def gather_for_metrics(self, tensor):
        return self.gather(tensor)

    def reduce(self, tensor, reduction='sum'):
        if self.num_processes == 1:
            return tensor
        t = tensor.clone()
        dist.all_reduce(t)
        return t / self.num_processes if reduction == 'mean' else t

    # ---- io ----
    def save(self, obj, path):
        if self.is_main_process:
            torch.save(obj, path)

    def end_training(self):
        if dist.is_initialized():
            dist.destroy_process_group()
