"""Replays the call sequences of the reference's entry scripts against the product packages on real GPUs, one process per
GPU under torchrun, through the `omegaconf` / `accelerate` stand-ins (diffusion-models-pytorch_b200/shims) -- i.e. what a
user of the reference gets when `PYTHONPATH` puts this package first.  Nothing of the reference is imported or vendored
(`/root/reference` does not exist on the GPU box); the statements below restate, in order and with the reference's own
names, the lines cited next to them:

  sampling : scripts/sample_uncond.py:115-195   (Accelerator, set_seed(device_specific), diffuser / model from the yaml
             through instantiate_from_config, load_state_dict, accelerator.prepare, per-rank folds, diffuser.sample,
             clamp, accelerator.gather()[:bs], image_norm_to_float)
  training : scripts/train_ddpm.py:100-192,194-216  (EMA, optimizer from conf.train.optim, accelerator.prepare -> DDP when
             world > 1, run_step with micro-batches / no_sync / accelerator.backward / clip_grad_norm_ / optimizer.step /
             ema.update, then the sample-with-EMA flow: ema.apply_shadow -> diffuser.sample -> gather -> ema.restore)

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tests/script_replay.py

Checks (rank 0 prints one JSON line, exit code 0 / 1): gathered sample count and finiteness, every rank's shard equals a
direct single-process `sample()` with that rank's seed (bitwise: the forward is deterministic), losses finite, parameters
identical on all ranks after the data-parallel steps, the EMA flow restores the live weights bit for bit."""
import importlib
import json
import math
import os
import sys
from contextlib import nullcontext

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, 'diffusion-models-pytorch_b200')
for p_ in (os.path.join(PKG, 'shims'), PKG, ROOT):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

import torch  # noqa: E402

import accelerate  # noqa: E402  (shim)
from omegaconf import DictConfig, OmegaConf  # noqa: E402  (shim)

import diffusions  # noqa: E402
from models import EMA  # noqa: E402

YAML = """
seed: 2022
data:
  img_channels: 1
  params:
    img_size: 32
model:
  target: models.unet.UNet
  params:
    in_channels: 1
    out_channels: 1
    dim: 64
    dim_mults: [1, 2, 2, 2]
    use_attn: [false, true, false, false]
    num_res_blocks: 2
    n_heads: 1
    dropout: 0.1
diffusion:
  target: diffusions.ddpm.DDPM
  params:
    total_steps: 1000
    beta_schedule: linear
    beta_start: 0.0001
    beta_end: 0.02
    objective: pred_eps
    var_type: fixed_large
train:
  n_steps: 3
  batch_size: 16
  micro_batch: 4
  clip_grad_norm: 1.0
  ema_decay: 0.9999
  ema_gradual: true
  n_samples: 6
  optim:
    target: torch.optim.Adam
    params:
      lr: 0.0002
"""


def instantiate_from_config(conf, **extra_params):          # utils/misc.py:71-78
    if isinstance(conf, DictConfig):
        conf = OmegaConf.to_container(conf)
    module, cls = conf['target'].rsplit('.', 1)
    cls = getattr(importlib.import_module(module, package=None), cls)
    params = conf.get('params', dict())
    params.update(extra_params)
    return cls(**params)


def amortize(n_samples: int, batch_size: int):               # utils/misc.py:41-44
    k = n_samples // batch_size
    r = n_samples % batch_size
    return k * [batch_size] if r == 0 else k * [batch_size] + [r]


def image_norm_to_float(image):                              # utils/misc.py image_norm_to_float
    return (image + 1) / 2


def main():
    import tempfile
    with tempfile.NamedTemporaryFile('w', suffix='.yaml', delete=False) as f:
        f.write(YAML)
    conf = OmegaConf.load(f.name)                                             # sample_uncond.py:115
    conf = OmegaConf.merge(conf, OmegaConf.from_dotlist(['train.n_samples=6']))    # :116
    os.unlink(f.name)
    seed, n_samples, batch_size, respace_steps = conf.seed, 11, 4, 5          # --seed --n_samples --batch_size --respace_steps

    # ------------------------------------------------------------------ scripts/sample_uncond.py:118-195
    accelerator = accelerate.Accelerator()                                    # :118
    device = accelerator.device
    accelerator.wait_for_everyone()
    accelerate.utils.set_seed(seed, device_specific=True)                     # :131
    world, rank = accelerator.num_processes, accelerator.process_index
    params = dict(total_steps=conf.diffusion.params.total_steps, beta_schedule=conf.diffusion.params.beta_schedule,
                  beta_start=conf.diffusion.params.beta_start, beta_end=conf.diffusion.params.beta_end,
                  objective=conf.diffusion.params.objective, respace_type='uniform', respace_steps=respace_steps,
                  device=device)                                              # :140-149
    diffuser = diffusions.ddim.DDIM(eta=0.0, **params)                        # :153-154 (--sampler ddim)
    model = instantiate_from_config(conf.model)                               # :163
    with torch.random.fork_rng(devices=[device]):      # stands in for load_weights(args.weights): must not touch the
        torch.manual_seed(2022)                        # rank-specific RNG streams set up above
        weights = {k: v.clone() for k, v in instantiate_from_config(conf.model).state_dict().items()}
    model.load_state_dict(weights)                                            # :166-167
    model = accelerator.prepare(model)                                        # :173
    model.eval()
    accelerator.wait_for_everyone()

    result = {'world': world}
    ok = True
    with torch.no_grad():                                                     # sample(): :178-195
        img_shape = (conf.data.img_channels, conf.data.params.img_size, conf.data.params.img_size)
        bspp = min(batch_size, math.ceil(n_samples / accelerator.num_processes))
        folds = amortize(n_samples, bspp * accelerator.num_processes)
        kept, mine = [], []
        for i, bs in enumerate(folds):
            init_noise = torch.randn((bspp, *img_shape), device=device)
            samples = diffuser.sample(
                model=accelerator.unwrap_model(model), init_noise=init_noise,
                tqdm_kwargs=dict(desc=f'Fold {i}/{len(folds)}', disable=True),
            ).clamp(-1, 1)
            mine.append((init_noise, samples))
            samples = accelerator.gather(samples)[:bs]
            if accelerator.is_main_process:
                for x in samples:
                    x = image_norm_to_float(x).cpu()
                    kept.append(x)
        if accelerator.is_main_process:
            ok &= len(kept) == n_samples and all(bool(torch.isfinite(x).all()) and float(x.min()) >= 0 and float(x.max()) <= 1
                                                 for x in kept)
            result['sampling'] = {'folds': folds, 'images': len(kept)}
        # this rank's shard, recomputed directly with the rank's seed: bitwise equal (deterministic forward)
        accelerate.utils.set_seed(seed, device_specific=True)
        same = True
        for (noise, smp) in mine:
            n2 = torch.randn((bspp, *img_shape), device=device)
            s2 = diffuser.sample(model=accelerator.unwrap_model(model), init_noise=n2, tqdm_kwargs=dict(disable=True)).clamp(-1, 1)
            same &= bool(torch.equal(noise, n2)) and bool(torch.equal(smp, s2))
        flags = accelerator.gather(torch.tensor([int(same)], device=device))
        ok &= bool(flags.all())
        result['shard_reproducible_on_every_rank'] = bool(flags.all())
    accelerator.wait_for_everyone()
    del model

    # ------------------------------------------------------------------ scripts/train_ddpm.py:100-192
    accelerator = accelerate.Accelerator(kwargs_handlers=[accelerate.DistributedDataParallelKwargs(find_unused_parameters=True)])
    torch.manual_seed(conf.seed)       # identical initial weights on every rank (the reference relies on DDP's broadcast)
    batch_size_per_process = conf.train.batch_size // accelerator.num_processes          # :101
    micro_batch = conf.train.micro_batch or batch_size_per_process                       # :102
    diffuser = instantiate_from_config(conf.diffusion, device=device)                    # :115
    model = instantiate_from_config(conf.model)                                          # :118
    ema = EMA(model.parameters(), decay=conf.train.ema_decay, gradual=conf.train.ema_gradual)   # :119
    optimizer = instantiate_from_config(conf.train.optim, params=model.parameters())     # :120
    step = 0
    model, optimizer, train_loader = accelerator.prepare(model, optimizer, [None])       # :166
    ema.to(device)                                                                       # :167
    accelerator.wait_for_everyone()
    accelerate.utils.set_seed(conf.seed, device_specific=True)

    def run_step(_batch):                                                                # :171-192
        optimizer.zero_grad()
        _batch = _batch[0] if isinstance(_batch, (tuple, list)) else _batch
        batch_size = _batch.shape[0]
        losses = []
        for i in range(0, batch_size, micro_batch):
            X = _batch[i:i + micro_batch].float()
            t = torch.randint(conf.diffusion.params.total_steps, (X.shape[0], ), device=device).long()
            loss_scale = X.shape[0] / batch_size
            no_sync = (i + micro_batch) < batch_size
            cm = accelerator.no_sync(model) if no_sync else nullcontext()
            with cm:
                loss = diffuser.loss_func(model, x0=X, t=t)
                accelerator.backward(loss * loss_scale)
            losses.append(loss.item())
        accelerator.clip_grad_norm_(model.parameters(), max_norm=conf.train.clip_grad_norm)
        optimizer.step()
        ema.update(model.parameters())
        return dict(loss=sum(losses) / len(losses), lr=optimizer.param_groups[0]['lr'])

    losses = []
    while step < conf.train.n_steps:                                                     # :222-243
        batch = (torch.rand(batch_size_per_process, *img_shape, device=device) * 2 - 1).clamp(-1, 1)
        model.train()
        losses.append(run_step(batch)['loss'])
        accelerator.wait_for_everyone()
        model.eval()
        step += 1
    fin = all(math.isfinite(v) for v in losses)
    # data-parallel invariant: identical parameters on every rank after the all-reduced steps
    csum = torch.stack([p.detach().double().sum() for p in accelerator.unwrap_model(model).parameters()]).sum().reshape(1)
    allc = accelerator.gather(csum)
    in_sync = bool((allc - allc[0]).abs().max() <= 1e-9 * allc[0].abs().clamp_min(1.0))
    ok &= fin and in_sync
    result['training'] = {'losses': losses, 'params_in_sync': in_sync, 'micro_batches_per_step': batch_size_per_process // micro_batch}

    with torch.no_grad():                                                                # sample(): :194-216
        unwrapped_model = accelerator.unwrap_model(model)
        before = [p.detach().clone() for p in model.parameters()]
        out_live = unwrapped_model(batch[:2], torch.tensor([10, 500], device=device)).clone()
        ema.apply_shadow(model.parameters())
        all_samples = []
        mb = min(micro_batch, math.ceil(conf.train.n_samples / accelerator.num_processes))
        folds = amortize(conf.train.n_samples, mb * accelerator.num_processes)
        sampler = diffusions.ddim.DDIM(eta=0.0, **params)
        for i, bs in enumerate(folds):
            init_noise = torch.randn((mb, *img_shape), device=device)
            samples = sampler.sample(model=unwrapped_model, init_noise=init_noise, tqdm_kwargs=dict(disable=True)).clamp(-1, 1)
            samples = accelerator.gather(samples)[:bs]
            all_samples.append(samples)
        all_samples = torch.cat(all_samples, dim=0)
        ema.restore(model.parameters())
        restored = all(bool(torch.equal(a, b)) for a, b in zip(before, model.parameters()))
        out_back = unwrapped_model(batch[:2], torch.tensor([10, 500], device=device))
        ema_ok = (all_samples.shape[0] == conf.train.n_samples and bool(torch.isfinite(all_samples).all()) and restored
                  and bool(torch.equal(out_live, out_back)))
    ok &= ema_ok
    result['ema_sampling'] = {'images': int(all_samples.shape[0]), 'weights_restored_bitwise': restored,
                              'forward_after_restore_bitwise': bool(torch.equal(out_live, out_back))}
    flags = accelerator.gather(torch.tensor([int(ok)], device=device))
    ok = bool(flags.all())
    result['ok'] = ok
    accelerator.wait_for_everyone()
    if accelerator.is_main_process:
        print(json.dumps(result), flush=True)
    accelerator.end_training()
    return 0 if ok else 1


if __name__ == '__main__':
    sys.exit(main())
