"""Data-parallel plumbing of the hot path (one process per GPU, torch.distributed; NCCL on GPUs, gloo in CPU tests).

Sampling shards by batch with NO traffic inside the loop -- every image's trajectory is independent, exactly like
the reference (scripts/sample_uncond.py:182-190): per-rank batch `bspp`, folds of `bspp * world` images, a
rank-specific seed, and one terminal all_gather per fold.  The training step is data-parallel: per-rank
micro-batches and one mean all-reduce of the gradients per step (scripts/train_ddpm.py:180-186).
"""
import math
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_plan(n_samples: int, batch_size: int, world: int):
    """(per-rank batch, fold sizes) as in scripts/sample_uncond.py:182-183 + utils/misc.py:74-77 (amortize)."""
    bspp = min(batch_size, math.ceil(n_samples / world))
    per_fold = bspp * world
    k, r = divmod(n_samples, per_fold)
    return bspp, k * [per_fold] + ([r] if r else [])


def rank_seed(seed: int, rank: int) -> int:
    """accelerate.utils.set_seed(seed, device_specific=True) semantics: seed + process index."""
    return seed + rank


def gather_samples(samples: torch.Tensor, keep: int = None) -> torch.Tensor:
    """Terminal all_gather of the per-rank samples (concatenated in rank order), truncated to `keep` images."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        out = samples
    else:
        parts = [torch.empty_like(samples) for _ in range(dist.get_world_size())]
        dist.all_gather(parts, samples.contiguous())
        out = torch.cat(parts, dim=0)
    return out if keep is None else out[:keep]


def allreduce_mean_(tensors: Sequence[torch.Tensor], bucket_bytes: int = 64 << 20) -> None:
    """In-place mean all-reduce of a list of (gradient) tensors, flattened into buckets so that NVLink/NVSwitch
    sees few large messages (177 MB of fp32 gradients for the CFG UNet = 3 buckets)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    world = dist.get_world_size()
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([t.reshape(-1) for t in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for t in bucket:
            n = t.numel()
            t.copy_(flat[off:off + n].view_as(t))
            off += n
        bucket, size = [], 0

    for t in tensors:
        bucket.append(t)
        size += t.numel() * t.element_size()
        if size >= bucket_bytes:
            flush()
    flush()


def allreduce_grads_(model: torch.nn.Module) -> None:
    """Mean all-reduce of a b200diff model's gradients after `loss.backward()`.  The backward pass returns every
    parameter gradient as a view into one flat fp32 buffer (models/backward.py), so a single in-place NCCL all-reduce
    over NVLink covers the whole model (143 MB for the CIFAR-10 UNet) with no flatten / unflatten copies; if the .grad
    tensors no longer alias that buffer (e.g. accumulated over micro-batches into older storage) the bucketed path runs."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    inner = getattr(model, 'module', model)
    flat = getattr(getattr(inner, 'engine', None), 'flat_grad', None)
    grads = [p.grad for p in inner.parameters() if p.grad is not None]
    if flat is not None and grads:
        lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
        if all(lo <= g.data_ptr() < hi for g in grads) and sum(g.numel() for g in grads) == flat.numel():
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat.div_(dist.get_world_size())
            return
    allreduce_mean_(grads)
