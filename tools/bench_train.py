#!/usr/bin/env python
"""Training-step timing on one B200 (or N under torchrun): python tools/bench_train.py [cfg|cifar] [B] [steps]
One step = zero_grad -> loss_func (diffuse + UNet forward with dropout + MSE) -> backward (hand-written adjoints) ->
[gradient all-reduce] -> fused clip + Adam + EMA.  Prints one JSON line: ms/step, images/s, TFLOP/s counting the
step as 3x the forward's algorithmic FLOPs, per-kernel totals of one profiled step."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, 'diffusion-models-pytorch_b200'))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import b200diff as K  # noqa: E402
import diffusions  # noqa: E402
import models  # noqa: E402
from b200diff.optim import FusedAdam  # noqa: E402
from b200diff.train import TrainStep  # noqa: E402

CFGC = dict(in_channels=3, out_channels=3, dim=128, dim_mults=[1, 2, 2, 2], use_attn=[False, True, True, False],
            num_res_blocks=2, num_classes=10, attn_head_dims=64, resblock_updown=True, dropout=0.1)
CIFAR = dict(in_channels=3, out_channels=3, dim=128, dim_mults=[1, 2, 2, 2], use_attn=[False, True, False, False],
             num_res_blocks=2, n_heads=1, dropout=0.1)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else 'cfg'
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    steps = int(sys.argv[3]) if len(sys.argv) > 3 else 10
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    torch.manual_seed(2022)
    if which == 'cfg':
        model, gf = models.UNetCategorialAdaGN(**CFGC).to(dev).train(), 14.396
        diffuser = diffusions.DDPM(total_steps=1000, beta_schedule='cosine', device=dev)
        y = (torch.arange(B) % 10).to(dev)
    else:
        model, gf = models.UNet(**CIFAR).to(dev).train(), 12.444
        diffuser = diffusions.DDPM(total_steps=1000, device=dev)
        y = None
    ema = models.EMA(model.parameters(), decay=0.9999)
    opt = FusedAdam(model.parameters(), lr=2e-4, capturable=True)
    use_graph = os.environ.get('B200_TRAIN_GRAPH', '1') != '0'
    step = TrainStep(model, diffuser, opt, ema=ema, clip_grad_norm=1.0, p_uncond=0.2 if y is not None else 0.0,
                     use_cuda_graph=use_graph)
    g = torch.Generator(device='cpu').manual_seed(2022 + rank)
    x0 = (torch.randn(B, 3, 32, 32, generator=g) * 0.5).clamp(-1, 1).to(dev)
    losses = []
    step.warmup(x0, y)        # eager steps, then the CUDA-graph captures (conditional and unconditional)
    losses.append(step(x0, y=y))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n0 = K.launch_count()
    import time
    e0.record()
    h0 = time.perf_counter()
    for _ in range(steps):
        losses.append(step(x0, y=y))
    host_ms = (time.perf_counter() - h0) * 1e3 / steps      # time the host needs to ENQUEUE one step
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    launches = (K.launch_count() - n0) // steps
    step.use_cuda_graph = False       # per-kernel CUDA-event profile of one eagerly launched step
    with K.Profiler() as prof:
        step(x0, y=y)
    kern = prof.summary()
    peak = 1387.7
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        peak = json.load(open(p)).get('bf16_tflops_sustained', peak)
    tf = 3 * gf * 1e9 * B / (ms * 1e-3) / 1e12
    if rank == 0:
        print(json.dumps({
            'model': which, 'batch_per_gpu': B, 'n_gpus': world, 'ms_per_step': ms, 'host_enqueue_ms_per_step': host_ms,
            'images_per_s': world * B / (ms * 1e-3), 'tflops_per_gpu_3x_fwd': tf, 'frac_of_bf16_sustained_peak': tf / peak,
            'kernels_per_step': launches, 'loss_first': float(losses[0]), 'loss_last': float(losses[-1]),
            'peak_mem_gib': torch.cuda.max_memory_allocated() / 2 ** 30,
            'kernels': {k: {'n': v['n'], 'ms': v['ms'],
                            **({'tflops': v['flops'] / (v['ms'] * 1e-3) / 1e12} if v['flops'] else {}),
                            **({'gbs': v['bytes'] / (v['ms'] * 1e-3) / 1e9} if v['bytes'] else {})}
                        for k, v in sorted(kern.items(), key=lambda kv: -kv[1]['ms'])}}), flush=True)
    if world > 1:
        # graphs that captured NCCL kernels must be released before the communicator goes away; a process that still
        # hangs in teardown is ended hard (the measurement is already printed)
        step.close()
        torch.cuda.synchronize()
        dist.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
