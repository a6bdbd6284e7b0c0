#!/usr/bin/env python
"""bench.py -- DDIM-50 CIFAR-10 sampling throughput (BASELINE.json metric) on N B200s of one node.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --steps K --warmup W    (the reference's CPU path = the oracle port)

One "step" = one complete DDIM-50 sampling run (50 UNet forwards + 50 fused sampler updates) of a batch of 256
CIFAR-10-shaped images per GPU through the reference-facing API `DDIM.sample(model, init_noise)`.
  value : whole-job images/s with the initial noise already resident in HBM (CUDA events, max over ranks)
  e2e   : same call with HOST buffers: pinned-host noise -> device, sample, result -> pinned host, per step
Weights: reference default initialisers under seed 2022 (no checkpoints offline); inputs: seeded Gaussian noise.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, 'diffusion-models-pytorch_b200')
for _p in (PKG, ROOT):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

CIFAR = dict(in_channels=3, out_channels=3, dim=128, dim_mults=[1, 2, 2, 2], use_attn=[False, True, False, False],
             num_res_blocks=2, n_heads=1, dropout=0.1)
GFLOP_PER_IMAGE_FWD = 12.444          # BASELINE.md section 2 (2*MAC of every conv/linear/bmm of the reference UNet)
METRIC = 'ddim50_cifar10_images_per_s'


def _peaks():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return dict(bf16_burst=p['bf16_tflops'], bf16_sustained=p.get('bf16_tflops_sustained', p['bf16_tflops']),
                    hbm=p['hbm_gbs'], source='measured (MEASURED_PEAKS.json)')
    return dict(bf16_burst=1590.0, bf16_sustained=1400.0, hbm=6650.0, source='fallback (B200_PROFILING.md)')


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons while the timed region runs."""
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.FIELDS}',
                                          '--format=csv,noheader,nounits', '-lms', '200'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(',')])

    def __exit__(self, *exc):
        if self.proc is not None:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(names, r[3:7]):
                if v.lower().startswith('active'):
                    reasons.add(name)
        return dict(sm_mhz=statistics.median(sm) if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


# --------------------------------------------------------------------------------------------------------------
# CPU baseline = the oracle port of the reference's CPU path (oracle/*.py), a bounded sample of the same workload
# --------------------------------------------------------------------------------------------------------------
def cpu_ddim_rate(batch, substeps, warmup, repeats, sample_steps=50):
    """images/s of DDIM-`sample_steps` on the host cores, extrapolated from `substeps` timed sampler steps
    (every step is identical work: one UNet forward + the sampler arithmetic)."""
    import models
    from oracle import diffusion_ref as R
    from oracle.unet_ref import UNetRef
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    torch.manual_seed(2022)
    sd = models.UNet(**CIFAR).state_dict()
    net = UNetRef(sd, dim=CIFAR['dim'], n_heads=1)
    d = R.DDIMRef(total_steps=1000, respace_type='uniform', respace_steps=sample_steps)
    x = torch.randn(batch, 3, 32, 32, generator=torch.Generator().manual_seed(2022))
    pairs = d._pairs()
    times = []
    with torch.no_grad():
        for rep in range(warmup + repeats):
            img = x
            t0 = time.perf_counter()
            for (t, tp) in pairs[:substeps]:
                tb = torch.full((batch,), t, dtype=torch.long)
                img = d.denoise(net(img, tb), img, t, tp, torch.zeros_like(img))['sample']
            dt = time.perf_counter() - t0
            if rep >= warmup:
                times.append(dt)
    per_step = statistics.mean(times) / substeps
    return batch / (per_step * sample_steps), cores, times


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    batch, substeps = 64, 2     # ~2-4 s of host work per bench step: the default K=5 / W=3 run takes well under a minute
    t0 = time.perf_counter()
    rate, cores, times = cpu_ddim_rate(batch, substeps, args.warmup, args.steps)
    ms_per_step = statistics.mean(times) * 1e3
    sample = (f'oracle port (oracle/unet_ref.py + diffusion_ref.py, PyTorch fp32 CPU), batch {batch}, {substeps} of 50 '
              f'DDIM steps per bench step, extrapolated linearly to 50 (identical work per step)')
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': 'images/s', 'n_gpus': args.gpus,
        'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms_per_step, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        # the same workload as our arm; one bench step = a bounded CPU sample of it (see cpu_baseline.sample)
        'config': {'workload': 'DDIM-50 CIFAR-10 32x32 UNet (configs/ddpm_cifar10.yaml), batch 256/GPU, eta=0, uniform '
                               'respacing, random-init weights (seed 2022)',
                   'batch_per_gpu': 256, 'sampler_steps': 50,
                   'sample': f'batch {batch}, {substeps} of 50 DDIM steps per bench step, host cores only'},
        'cpu_baseline': {'value': rate, 'unit': 'images/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': rate, 'unit': 'images/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.perf_counter() - t0,
    }
    print(json.dumps(line), flush=True)



# --------------------------------------------------------------------------------------------------------------
# Checks and baselines around the timed region (never inside it)
# --------------------------------------------------------------------------------------------------------------
def parity_check(model, diffuser, noise_dev, B, S):
    """Parity of the EXACT benchmarked call (batch 256, graph-replayed DDIM-50) against the fp32 oracle on the same GPU
    (PyTorch eager, TF32 off): eps rel-L2 at one timestep and PSNR of the final samples.  Outside the timed region."""
    import math
    import diffusions  # noqa: F401
    from oracle import diffusion_ref as R
    from oracle.unet_ref import UNetRef
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = UNetRef(model.state_dict(), dim=CIFAR['dim'], n_heads=1).to(noise_dev.device)
        orc = R.DDIMRef(total_steps=1000, respace_type='uniform', respace_steps=S)
        orc.alphas_cumprod = orc.alphas_cumprod.to(noise_dev.device)
        with torch.no_grad():
            t = torch.full((1,), 500, device=noise_dev.device, dtype=torch.long).expand(B)
            e_our, e_ref = model(noise_dev, t), ref(noise_dev, t.contiguous())
            rel = ((e_our - e_ref).norm() / e_ref.norm()).item()
            got = diffuser.sample(model, noise_dev, tqdm_kwargs=dict(disable=True)).clamp(-1, 1)
            got2 = diffuser.sample(model, noise_dev, tqdm_kwargs=dict(disable=True)).clamp(-1, 1)
            want = orc.sample(ref, noise_dev, noises=[torch.zeros_like(noise_dev)] * S).clamp(-1, 1)
        mse = ((got - want) ** 2).mean().item()
        per_img = ((got - want) ** 2).flatten(1).mean(dim=1)
        return {'checked': f'the benchmarked call: batch {B}, DDIM-{S} through DDIM.sample (CUDA-graph replay) vs oracle fp32 '
                           f'eager on the same GPU, TF32 off, same initial noise',
                'eps_rel_l2': rel, 'eps_gate': 1e-2,
                'ddim_psnr_db': 10 * math.log10(4.0 / max(mse, 1e-20)), 'psnr_gate_db': 40.0,
                'worst_image_psnr_db': 10 * math.log10(4.0 / max(per_img.max().item(), 1e-20)),
                'bitwise_reproducible': bool(torch.equal(got, got2)),
                'ok': bool(rel <= 1e-2 and mse <= 4.0 / 10 ** 4.0 and torch.equal(got, got2))}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = tf


def gpu_reference(model, noise_dev, B, S, substeps=3):
    """The kernel-for-kernel bar of BASELINE.md section 3: the reference's own op sequence (oracle/unet_ref.py = the same
    F.conv2d / group_norm / bmm calls, PyTorch eager, cuDNN / cuBLAS) on THIS GPU, same batch: fp32, TF32, and
    autocast(bf16) + channels_last.  `substeps` timed DDIM steps after one warm-up step, extrapolated to S (identical
    work per step).  Reported beside our number, not the target."""
    from oracle import diffusion_ref as R
    from oracle.unet_ref import UNetRef
    dev = noise_dev.device
    out = {'kind': 'port (oracle/unet_ref.py + diffusion_ref.py: the reference\'s ATen op sequence in PyTorch eager on this GPU)',
           'sample': f'batch {B}, {substeps} of {S} DDIM steps timed after 1 warm-up step, extrapolated linearly'}
    tf = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
    d = R.DDIMRef(total_steps=1000, respace_type='uniform', respace_steps=S)
    d.alphas_cumprod = d.alphas_cumprod.to(dev)
    pairs = d._pairs()
    sd = model.state_dict()
    try:
        torch.backends.cudnn.benchmark = True
        for tag, tf32, autocast in (('fp32', False, False), ('tf32', True, False), ('bf16_autocast_channels_last', True, True)):
            torch.backends.cudnn.allow_tf32 = tf32
            torch.backends.cuda.matmul.allow_tf32 = tf32
            sdx = {k: (v.contiguous(memory_format=torch.channels_last) if (autocast and v.dim() == 4) else v) for k, v in sd.items()}
            net = UNetRef(sdx, dim=CIFAR['dim'], n_heads=1).to(dev)
            x = noise_dev.contiguous(memory_format=torch.channels_last) if autocast else noise_dev

            def steps(n, img):
                for (t, tp) in pairs[:n]:
                    tb = torch.full((B,), t, dtype=torch.long, device=dev)
                    if autocast:
                        with torch.autocast('cuda', dtype=torch.bfloat16):
                            eps = net(img, tb).float()
                    else:
                        eps = net(img, tb)
                    img = d.denoise(eps, img, t, tp, torch.zeros_like(img))['sample']
                return img
            with torch.no_grad():
                steps(1, x)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                steps(substeps, x)
                e1.record()
                torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / substeps
            out[tag] = {'images_per_s': B / (ms * 1e-3 * S), 'ms_per_ddim_step': ms}
            del net
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = tf
    return out


# --------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import b200diff as K
    import diffusions
    import models

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the product path has no CPU fallback '
                         '(use --impl reference for the CPU baseline)')
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    stdout_fd = None
    if world > 1:
        # NCCL prints its version banner (and NCCL_DEBUG output) on stdout: route fd 1 to stderr while the job runs and
        # restore it for the single JSON line
        sys.stdout.flush()
        stdout_fd = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group('nccl', device_id=dev)
    B, S = args.batch, args.sample_steps

    torch.manual_seed(2022)
    model = models.UNet(**CIFAR).to(dev).eval()
    diffuser = diffusions.DDIM(total_steps=1000, beta_schedule='linear', respace_type='uniform', respace_steps=S,
                               eta=0.0, device=dev)
    gen = torch.Generator(device='cpu').manual_seed(2022 + rank)
    noise_host = torch.randn(B, 3, 32, 32, generator=gen).pin_memory()
    out_host = torch.empty(B, 3, 32, 32).pin_memory()
    noise_dev = noise_host.to(dev)
    gathered = [torch.empty_like(noise_dev) for _ in range(world)] if world > 1 else None
    quiet = dict(disable=True)

    def step_resident():
        s = diffuser.sample(model, noise_dev, tqdm_kwargs=quiet).clamp_(-1, 1)
        if world > 1:   # the reference's terminal gather (scripts/sample_uncond.py:190)
            dist.all_gather(gathered, s)
        return s

    def step_e2e():
        x = noise_host.to(dev, non_blocking=True)
        s = diffuser.sample(model, x, tqdm_kwargs=quiet).clamp_(-1, 1)
        if world > 1:
            dist.all_gather(gathered, s)
        out_host.copy_(s, non_blocking=True)
        torch.cuda.current_stream().synchronize()   # the caller owns the images on the host after each step

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n0 = K.launch_count()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item(), K.launch_count() - n0

    with torch.no_grad():
        for _ in range(max(args.warmup, 3)):
            step_resident()
        with ClockSampler(local) as clocks:
            ms_total, launches = timed(step_resident, args.steps)
        for _ in range(2):
            step_e2e()
        ms_e2e, _ = timed(step_e2e, args.steps)

        # ---- roofline of the dominant kernel: per-launch CUDA events over eager forwards, right after the
        # long timed region (GPU in its sustained, power-capped state) ----
        t_mid = torch.full((1,), 500, device=dev, dtype=torch.long).expand(B)
        model(noise_dev, t_mid)
        with K.Profiler() as prof:
            for _ in range(5):
                model(noise_dev, t_mid)
        kern = prof.summary()

    peaks = _peaks()
    total_ms = sum(v['ms'] for v in kern.values())
    conv = kern.get('conv_gemm', dict(n=1, ms=1.0, flops=0.0))
    conv_tflops = conv['flops'] / (conv['ms'] * 1e-3) / 1e12
    gn = kern.get('groupnorm_apply')
    traffic = gn_traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'conv_gemm_traffic.json')   # written by tools/capture_traffic.sh (ncu)
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        traffic = tj.get('dram_bytes_per_launch')
        gn_traffic = tj.get('kernels', {}).get('groupnorm_apply', {}).get('dram_bytes_per_launch')
    images = world * B * args.steps
    value = images / (ms_total * 1e-3)
    e2e_value = images / (ms_e2e * 1e-3)
    line = {
        'metric': METRIC, 'value': value, 'unit': 'images/s', 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': ms_total / args.steps, 'higher_is_better': True,
        'scaling': 'weak', 'vs_baseline': None, 'dtype': 'bf16', 'data': 'synthetic',
        'config': {
            'workload': f'DDIM-{S} CIFAR-10 32x32 UNet (configs/ddpm_cifar10.yaml), batch {B}/GPU, eta=0, uniform '
                        f'respacing, random-init weights (seed 2022)',
            'batch_per_gpu': B, 'sampler_steps': S, 'sharding': f'batch x{world}, no traffic inside the loop, '
                                                               f'terminal all_gather',
            'l2': 'each step streams ~4 GB of activations per forward x 50 forwards (>> 126 MB L2); no explicit flush',
            'accumulate': 'fp32 (TMEM), bf16 operands, fp32 residual stream / GroupNorm statistics / sampler',
        },
        'unet_fwd_tflops_per_gpu': GFLOP_PER_IMAGE_FWD * 1e9 * B * S * args.steps / (ms_total * 1e-3) / 1e12,
        'unet_fwd_frac_of_bf16_peak': GFLOP_PER_IMAGE_FWD * 1e9 * B * S * args.steps / (ms_total * 1e-3) / 1e12 /
        peaks['bf16_sustained'],
        'e2e': {'value': e2e_value, 'unit': 'images/s', 'h2d_bytes_per_step': noise_host.numel() * 4,
                'd2h_bytes_per_step': out_host.numel() * 4},
        'gpu_launches': launches,
        'roofline': {
            'kernel': 'conv_gemm_kernel (tcgen05 implicit-GEMM conv)', 'bound': 'tensor',
            'achieved': conv_tflops, 'peak': peaks['bf16_sustained'], 'unit': 'TFLOP/s',
            'frac': conv_tflops / peaks['bf16_sustained'], 'traffic': traffic,
            'peak_source': peaks['source'] + ', sustained bf16 figure (kernel timed inside a long step)',
            'launches_timed': conv['n'], 'avg_launch_us': conv['ms'] * 1e3 / max(conv['n'], 1),
            'share_of_forward': conv['ms'] / total_ms,
            'algorithmic_flops_per_forward': conv['flops'] / 5,
        },
        'kernels': {k: {'n_per_forward': v['n'] // 5, 'ms_per_forward': v['ms'] / 5,
                        **({'tflops': v['flops'] / (v['ms'] * 1e-3) / 1e12} if v['flops'] else {}),
                        **({'gbs': v['bytes'] / (v['ms'] * 1e-3) / 1e9} if v['bytes'] else {})}
                    for k, v in sorted(kern.items(), key=lambda kv: -kv[1]['ms'])},
        'clocks': clocks.summary(),
    }
    if gn:
        gbs = gn['bytes'] / (gn['ms'] * 1e-3) / 1e9
        line['roofline_groupnorm'] = {'kernel': 'groupnorm_apply_kernel', 'bound': 'hbm', 'achieved': gbs,
                                      'peak': peaks['hbm'], 'unit': 'GB/s', 'frac': gbs / peaks['hbm'],
                                      'traffic': gn_traffic,
                                      'algorithmic_bytes_per_launch': gn['bytes'] / max(gn['n'], 1)}
    if rank == 0 and not args.no_parity:
        try:
            line['parity'] = parity_check(model, diffuser, noise_dev, B, S)
        except Exception as e:  # noqa: BLE001
            line['parity'] = {'error': repr(e)}
    if world == 1 and not args.no_extras:
        try:
            line['gpu_reference'] = gpu_reference(model, noise_dev, B, S)
            line['gpu_reference']['ours_over_bf16_autocast'] = value / line['gpu_reference']['bf16_autocast_channels_last']['images_per_s']
        except Exception as e:  # noqa: BLE001
            line['gpu_reference'] = {'error': repr(e)}
    if not args.no_extras:
        # free the headline model's arena before the (larger) secondary workloads
        del model, diffuser
        torch.cuda.empty_cache()
        try:
            # world > 1: only the data-parallel training step (collective: every rank takes part)
            line['extras'] = _extras(dev, peaks, world, rank)
            if 'sampler_update' in line['extras']:
                line['roofline_sampler'] = line['extras'].pop('sampler_update')
        except Exception as e:  # noqa: BLE001  (secondary numbers must never cost the headline line)
            line['extras'] = {'error': repr(e)}
    if stdout_fd is not None:
        torch.cuda.synchronize()
        dist.destroy_process_group()
        sys.stdout.flush()
        os.dup2(stdout_fd, 1)
        os.close(stdout_fd)
    if rank == 0:
        # CPU baseline on rank 0: a bounded sample (batch 64, 1 warm-up + 3 timed pairs of DDIM steps: 10-20 s of host work)
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, _ = cpu_ddim_rate(batch=64, substeps=2, warmup=1, repeats=3)
            line['cpu_baseline'] = {
                'value': rate, 'unit': 'images/s', 'cores': cores, 'kind': 'port',
                'sample': 'oracle port (PyTorch fp32 CPU), batch 64, 1 warm-up + 3 timed repeats of 2 DDIM steps (UNet '
                          'forward + sampler arithmetic), extrapolated linearly to 50 steps'}
        print(json.dumps(line), flush=True)


def _graph_time(fn, iters, warm=1):
    """ms per call of `fn` (device time, CUDA events), after `warm` untimed calls."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


ADM256 = dict(image_size=256, in_channels=3, model_channels=256, out_channels=6, num_res_blocks=2,
              attention_resolutions=[32, 16, 8], dropout=0.0, channel_mult=[1, 1, 2, 2, 4, 4], num_classes=1000,
              num_heads=4, num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True)
CFGC = dict(in_channels=3, out_channels=3, dim=128, dim_mults=[1, 2, 2, 2], use_attn=[False, True, True, False],
            num_res_blocks=2, num_classes=10, attn_head_dims=64, resblock_updown=True, dropout=0.1)
MNIST = dict(in_channels=1, out_channels=1, dim=64, dim_mults=[1, 2, 2, 2], use_attn=[False, True, False, False],
             num_res_blocks=2, n_heads=1, dropout=0.1)
PESSER = dict(in_channels=3, out_ch=3, ch=128, ch_mult=[1, 1, 2, 2, 4, 4], num_res_blocks=2, attn_resolutions=[16],
              dropout=0.0, resamp_with_conv=True, resolution=256)


def _extra_train(dev, peaks, world, rank):
    """One classifier-free-guidance training step (configs[2]: UNetCategorialAdaGN, batch 128 PER GPU, dropout 0.1):
    loss_func -> hand-written backward -> [NCCL all-reduce of the flat gradient buffer over the data-parallel ranks,
    captured inside the step's CUDA graph] -> fused clip + Adam + EMA.  Device time, max over ranks."""
    import torch.distributed as dist
    import b200diff as K
    import diffusions
    import models
    from b200diff.optim import FusedAdam
    from b200diff.train import TrainStep
    torch.manual_seed(2022)
    model = models.UNetCategorialAdaGN(**CFGC).to(dev).train()
    diffuser = diffusions.DDPM(total_steps=1000, beta_schedule='cosine', device=dev)
    step = TrainStep(model, diffuser, FusedAdam(model.parameters(), lr=2e-4, capturable=True),
                     ema=models.EMA(model.parameters()), clip_grad_norm=1.0, p_uncond=0.2, use_cuda_graph=True)
    B = 128
    x0 = (torch.randn(B, 3, 32, 32, generator=torch.Generator().manual_seed(2022 + rank)) * 0.5).clamp(-1, 1).to(dev)
    y = (torch.arange(B, device=dev) + rank) % 10
    first = step(x0, y=y)
    step.warmup(x0, y)        # eager steps, then the CUDA-graph captures (conditional and unconditional)
    # p_uncond draws come from the host RNG: seed it identically on every rank so that all ranks replay the same graph
    torch.manual_seed(1234)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    n0 = K.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    iters = 10
    e0.record()
    for _ in range(iters):
        loss = step(x0, y=y)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms_t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
    if world > 1:
        dist.all_reduce(ms_t, op=dist.ReduceOp.MAX)
    ms = ms_t.item()
    tf = 3 * 14.396e9 * B / (ms * 1e-3) / 1e12
    res = {'workload': 'UNetCategorialAdaGN (configs/ddpm_cfg_cifar10.yaml) noise-prediction training step, batch 128 per '
                       f'GPU, dropout 0.1, p_uncond 0.2, clip 1.0 + Adam + EMA fused, whole step (incl. the gradient '
                       f'all-reduce when data-parallel) replayed as one CUDA graph, {world} GPU(s)',
           'n_gpus': world, 'ms_per_step': ms, 'images_per_s': world * B / (ms * 1e-3), 'tflops_3x_fwd_per_gpu': tf,
           'frac_of_bf16_sustained_peak': tf / peaks['bf16_sustained'],
           'kernels_per_step': (K.launch_count() - n0) // iters,
           'allreduce_bytes_per_step': 4 * sum(p.numel() for p in model.parameters()) if world > 1 else 0,
           'loss_first': float(first), 'loss_last': float(loss)}
    step.close()
    torch.cuda.synchronize()
    del step, model
    torch.cuda.empty_cache()
    return res


def _extras(dev, peaks, world=1, rank=0):
    """Secondary workloads of BASELINE.json on the same GPU(s), same timing rules (CUDA events, warm-up, inputs resident).
    world > 1: the data-parallel training step only.  world == 1: ADM-256 forward (configs[4]; north_star >= 60 % of bf16
    peak), the training step, the K4 roofline, and sampling throughput of configs[0] (DDPM-1000 MNIST, batch 16),
    configs[2] (DDIMCFG-50, s = 3, batch 128), configs[3] (pesser 256x256 DDIM-100, batch 32) and configs[4]
    (ADM-256 DDIM-250, batch 16; 25 of 250 steps timed) through the public `sample()` API."""
    import b200diff as K
    import diffusions
    import models
    from models.adm.unet import UNetModel
    out = {}
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    quiet = dict(disable=True)
    if world > 1:
        out['cfg_train_step'] = _extra_train(dev, peaks, world, rank)
        return out
    # ---- (1) ADM-256 forward + DDIM-250 sampling ----
    torch.manual_seed(2022)
    m = UNetModel(**ADM256)
    gen = torch.Generator().manual_seed(2022)
    with torch.no_grad():      # ADM zero-initialises conv2 / proj_out / out: re-draw them N(0, 0.02)
        for p_ in m.parameters():
            if not bool(p_.any()):
                p_.copy_(torch.randn(p_.shape, generator=gen) * 0.02)
    m = m.to(dev).eval()
    B = 16
    x = torch.randn(B, 3, 256, 256, device=dev)
    t = torch.full((1,), 500, device=dev, dtype=torch.long).expand(B)
    y = (torch.arange(B, device=dev) % 1000)
    o = torch.empty(B, 6, 256, 256, device=dev)
    with torch.no_grad():
        for _ in range(3):
            m(x, t, y, out=o)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            m(x, t, y, out=o)
        ms = _graph_time(g.replay, 5)
    tf = 2239.67e9 * B / (ms * 1e-3) / 1e12
    out['adm256_forward'] = {'workload': 'ADM ImageNet-256 class-cond UNet forward (guided-diffusion 256x256_diffusion.yaml), '
                                         'batch 16, random weights', 'ms_per_forward': ms, 'tflops': tf,
                             'frac_of_bf16_sustained_peak': tf / peaks['bf16_sustained'],
                             'gflop_per_image': 2239.67, 'images_fwd_per_s': B / (ms * 1e-3),
                             'arena_gb': m.engine.arena_bytes() / 2 ** 30}
    del g
    try:
        steps_timed, S_adm = 25, 250
        d_adm = diffusions.DDIM(total_steps=1000, beta_schedule='linear', var_type='learned_range', respace_type='uniform',
                                respace_steps=steps_timed, device=dev)
        with torch.no_grad():
            ms_s = _graph_time(lambda: d_adm.sample(m, x, tqdm_kwargs=quiet, model_kwargs=dict(y=y)), 1)
        out['adm256_ddim250_sampling'] = {
            'workload': 'configs[4]: ADM ImageNet-256 class-conditional UNet, DDIM-250, batch 16 through DDIM.sample (CUDA-graph '
                        f'replay per step); {steps_timed} of {S_adm} steps timed (identical work per step), extrapolated',
            'ms_per_ddim_step': ms_s / steps_timed, 'images_per_s': B / (ms_s / steps_timed * S_adm * 1e-3)}
    except Exception as e:  # noqa: BLE001
        out['adm256_ddim250_sampling'] = {'error': repr(e)}
    del m, x, o
    torch.cuda.empty_cache()
    # ---- (2) CFG training step ----
    out['cfg_train_step'] = _extra_train(dev, peaks, 1, 0)
    # ---- (3) sampling throughput of the other BASELINE configs through the public API ----
    try:
        torch.manual_seed(2022)
        mc = models.UNetCategorialAdaGN(**CFGC).to(dev).eval()
        dc = diffusions.DDIMCFG(guidance_scale=3.0, total_steps=1000, beta_schedule='cosine', respace_type='uniform',
                                respace_steps=50, device=dev)
        Bc = 128
        xc = torch.randn(Bc, 3, 32, 32, device=dev)
        yc = torch.arange(Bc, device=dev) % 10
        with torch.no_grad():
            ms_c = _graph_time(lambda: dc.sample(mc, xc, tqdm_kwargs=quiet, model_kwargs=dict(y=yc)), 3)
        out['cfg_ddim50_sampling'] = {
            'workload': 'configs[2]: UNetCategorialAdaGN (ddpm_cfg_cifar10.yaml), DDIMCFG-50, guidance scale 3, batch 128, '
                        'two forwards per step (cond + uncond) + fused guidance mix / sampler step, full 50-step runs',
            'ms_per_run': ms_c, 'images_per_s': Bc / (ms_c * 1e-3),
            'unet_fwd_tflops': 2 * 50 * 14.396e9 * Bc / (ms_c * 1e-3) / 1e12,
            'frac_of_bf16_sustained_peak': 2 * 50 * 14.396e9 * Bc / (ms_c * 1e-3) / 1e12 / peaks['bf16_sustained']}
        del mc, dc
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out['cfg_ddim50_sampling'] = {'error': repr(e)}
    try:
        torch.manual_seed(2022)
        mm = models.UNet(**MNIST).to(dev).eval()
        dm = diffusions.DDPM(total_steps=1000, var_type='fixed_large', device=dev)
        xm = torch.randn(16, 1, 32, 32, device=dev)
        with torch.no_grad():
            ms_m = _graph_time(lambda: dm.sample(mm, xm, tqdm_kwargs=quiet), 1)
        out['mnist_ddpm1000_sampling'] = {
            'workload': 'configs[0]: DDPM MNIST UNet (ddpm_mnist.yaml, 32x32 padded input), 1000 steps, batch 16 (launch-bound: '
                        '~115 kernels per step replayed as one CUDA graph)', 'ms_per_run': ms_m,
            'images_per_s': 16 / (ms_m * 1e-3), 'us_per_step': ms_m}
        del mm, dm
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out['mnist_ddpm1000_sampling'] = {'error': repr(e)}
    try:
        from models.pesser.model import Model as Pesser
        torch.manual_seed(2022)
        mp = Pesser(**PESSER).to(dev).eval()
        dp = diffusions.DDIM(total_steps=1000, respace_type='uniform', respace_steps=100, device=dev)
        xp = torch.randn(32, 3, 256, 256, device=dev)
        with torch.no_grad():
            ms_p = _graph_time(lambda: dp.sample(mp, xp, tqdm_kwargs=quiet), 1)
        out['pesser256_ddim100_sampling'] = {
            'workload': 'configs[3]: CelebA-HQ 256x256 pesser UNet (ddpm_celebahq.yaml), DDIM-100, batch 32, one full 100-step run',
            'ms_per_run': ms_p, 'images_per_s': 32 / (ms_p * 1e-3), 'ms_per_ddim_step': ms_p / 100}
        del mp, dp, xp
        torch.cuda.empty_cache()
    except Exception as e:  # noqa: BLE001
        out['pesser256_ddim100_sampling'] = {'error': repr(e)}
    try:
        # ---- (4) K4 sampler update against the HBM roofline (CIFAR-size tensors are launch-latency bound,
        # so the GB/s figure is taken on tensors whose working set exceeds the 126 MB L2; three buffer sets rotate).
        # DDIM eta = 0: the noise term is sqrt(0) * z, so the kernel reads model output and x_t and writes x_{t-1}:
        # 12 B / element; the eta = 1 line adds the 4 B / element noise stream.
        Bk, Hk = 128, 256     # the shape of the ncu capture in profiles/sampler_traffic.json
        sets = [[torch.randn(Bk, 3, Hk, Hk, device=dev) for _ in range(3)] + [torch.empty(Bk, 3, Hk, Hk, device=dev)]
                for _ in range(3)]
        res = {}
        for tag, eta, streams in (('eta0', 0.0, 3), ('eta1', 1.0, 4)):
            dd = diffusions.DDIM(total_steps=1000, respace_type='uniform', respace_steps=50, eta=eta, device=dev)
            row = dd._coef_row(500, 480)
            call = lambda s_: K.sampler_step(s_[0], s_[1], row, objective='pred_eps', clip=True, noise=s_[2], sample=s_[3])  # noqa: E731
            for s_ in sets:
                call(s_)
            torch.cuda.synchronize()
            e0, e1 = ev(), ev()
            e0.record()
            for _ in range(10):
                for s_ in sets:
                    call(s_)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 30 * 1e3
            by = 4.0 * Bk * 3 * Hk * Hk * streams
            res[tag] = (us, by)
        k4_traffic = None
        tp = os.path.join(ROOT, 'profiles', 'sampler_traffic.json')    # ncu dram__bytes_read/write.sum of this launch shape
        if os.path.exists(tp):
            with open(tp) as f:
                k4_traffic = json.load(f).get('dram_bytes_per_launch')
        us, by = res['eta0']
        out['sampler_update'] = {'kernel': 'sampler_step_vec4_kernel (K4: predict + clip + DDIM/DDPM step, fused)',
                                 'bound': 'hbm', 'achieved': by / us / 1e3, 'peak': peaks['hbm'], 'unit': 'GB/s',
                                 'frac': by / us / 1e3 / peaks['hbm'], 'traffic': k4_traffic,
                                 'traffic_note': 'ncu capture of the eta = 1 (4-stream) launch of round 1',
                                 'algorithmic_bytes_per_launch': by, 'avg_launch_us': us,
                                 'eta1_4_streams': {'achieved': res['eta1'][1] / res['eta1'][0] / 1e3, 'avg_launch_us': res['eta1'][0],
                                                    'frac': res['eta1'][1] / res['eta1'][0] / 1e3 / peaks['hbm']},
                                 'workload': f'DDIM eta=0 step on [{Bk},3,{Hk},{Hk}] fp32 tensors (101 MB each, 3 streams = 12 B/element: '
                                             f'model output + x_t read, x_(t-1) written; the zero-variance noise is not read), '
                                             f'30 launches over 3 buffer sets (> L2 between reuses)'}
    except Exception as e:  # noqa: BLE001  (must not cost the numbers above)
        out['sampler_update_error'] = repr(e)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--batch', type=int, default=256)
    ap.add_argument('--sample-steps', type=int, default=50)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extras', action='store_true', help='skip the secondary workloads (ADM-256, training step, configs A/C/D/E, GPU reference)')
    ap.add_argument('--no-parity', action='store_true', help='skip the parity check of the benchmarked call against the oracle')
    args = ap.parse_args()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
