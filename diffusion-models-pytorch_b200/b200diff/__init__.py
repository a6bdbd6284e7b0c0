"""ctypes binding of libb200diff.so (C ABI in include/b200diff.h).

PyTorch is plumbing here: it owns device memory and the stream; every arithmetic op on the diffusion hot
path is a kernel of the shared library.  There is deliberately no fallback: if the library is missing or a
kernel fails, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import POINTER, Structure, c_char_p, c_double, c_float, c_int, c_int8, c_longlong, c_void_p

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get('B200DIFF_LIB') or os.path.join(_HERE, 'libb200diff.so')   # override: A/B experiments

OUT_F32_NHWC, OUT_BF16_NHWC, OUT_F32_NCHW, OUT_BF16_NCHW = 0, 1, 2, 3
OBJ = {'pred_eps': 0, 'pred_x0': 1, 'pred_v': 2}
(SC_SQRT_RECIP_AC, SC_SQRT_RECIPM1_AC, SC_SQRT_AC, SC_SQRT_1M_AC, SC_X0_COEF, SC_XT_COEF, SC_EPS_COEF, SC_VAR,
 SC_MIN_LOGVAR, SC_MAX_LOGVAR, SC_ADD_NOISE) = range(11)
SC_COUNT = 12
STAT_Q1, STAT_Q2 = float(2 ** 30), float(2 ** 22)   # fixed-point scales of the [B, C, 2] int64 GroupNorm statistics


def new_stats(B, C, device):
    """Zeroed [B, C, 2] int64 accumulator for a producer's fused GroupNorm statistics (b200_conv_desc.stats)."""
    return torch.zeros((B, C, 2), dtype=torch.int64, device=device)


def stats_to_float(stats):
    """[B, C, 2] int64 fixed point -> float64 (sum, sum of squares); for tests and diagnostics."""
    out = stats.to(torch.float64)
    out[..., 0] /= STAT_Q1
    out[..., 1] /= STAT_Q2
    return out


def stats_from_float(st):
    """float (sum, sum of squares) [B, C, 2] -> the int64 fixed-point layout the kernels read (tests)."""
    out = torch.empty(st.shape, dtype=torch.int64, device=st.device)
    out[..., 0] = torch.round(st[..., 0].to(torch.float64) * STAT_Q1).to(torch.int64)
    out[..., 1] = torch.round(st[..., 1].to(torch.float64) * STAT_Q2).to(torch.int64)
    return out


class ConvDesc(Structure):
    _fields_ = [
        ('a0', c_void_p), ('a0_C', c_int), ('a0_H', c_int), ('a0_W', c_int), ('a0_planes', c_int),
        ('a1', c_void_p), ('a1_C', c_int), ('a1_H', c_int), ('a1_W', c_int), ('a1_planes', c_int),
        ('w', c_void_p), ('w_rows', c_int), ('w_K', c_int), ('w_rows_per_phase', c_int),
        ('B', c_int), ('Ho', c_int), ('Wo', c_int),
        ('phases', c_int), ('N', c_int), ('ntaps0', c_int),
        ('taps0', c_int8 * 144), ('tap1', c_int8 * 4),
        ('bias', c_void_p), ('rowadd', c_void_p), ('rowadd_ld', c_int),
        ('residual', c_void_p), ('res_ld', c_int),
        ('stats', c_void_p),
        ('out', c_void_p), ('out_mode', c_int), ('out_ld', c_int), ('out_H', c_int), ('out_W', c_int),
        ('osy', c_int), ('osx', c_int),
    ]


class SamplerDesc(Structure):
    _fields_ = [
        ('model_out', c_void_p), ('model_out_uncond', c_void_p), ('xt', c_void_p), ('noise', c_void_p),
        ('coef', c_void_p),
        ('B', c_int), ('C', c_int), ('Cm', c_int), ('HW', c_int),
        ('objective', c_int), ('clip', c_int), ('learned_range', c_int),
        ('guidance_scale', c_double),
        ('sample', c_void_p), ('mean', c_void_p), ('pred_x0', c_void_p), ('pred_eps', c_void_p),
        ('var_out', c_void_p),
    ]


class GnFuseDesc(Structure):      # mirrors b200_gn_fuse_desc (fused conv + next GroupNorm)
    _fields_ = [
        ('gamma', c_void_p), ('beta', c_void_p), ('scale', c_void_p), ('shift', c_void_p), ('out_norm', c_void_p),
        ('out_norm_ld', c_int), ('out_raw_bf16', c_void_p), ('ss_ld', c_int), ('groups', c_int), ('apply_silu', c_int),
        ('eps', ctypes.c_float),
        ('xstats', c_void_p), ('xcount', c_void_p),
    ]


class AttnBlockDesc(ctypes.Structure):
    """Mirror of b200_attn_block_desc (include/b200diff.h)."""
    _fields_ = [('x', c_void_p), ('x_stats', c_void_p), ('gamma', c_void_p), ('beta', c_void_p), ('w', c_void_p),
                ('bias', c_void_p), ('out', c_void_p), ('out_stats', c_void_p), ('dbg', c_void_p * 6),
                ('B', c_int), ('T', c_int), ('C', c_int), ('heads', c_int), ('groups', c_int),
                ('eps', c_float), ('scale', c_float), ('pad_', c_int)]


class SplitDesc(Structure):       # mirrors b200_split_desc (FP32 mode: fp32 -> bf16 hi/lo split operands)
    _fields_ = [('in_', c_void_p), ('rows', c_longlong), ('in_ld', c_int), ('in_col0', c_int), ('C', c_int), ('group', c_int),
                ('pattern', c_int), ('act', c_int), ('out', c_void_p), ('out_ld', c_int), ('out_col0', c_int),
                ('planes_rows', c_int), ('parity_H', c_int), ('parity_W', c_int)]


class GemmOperand(Structure):
    _fields_ = [('ptr', c_void_p), ('rows', c_int), ('ld', c_int), ('batch_stride', c_longlong),
                ('col_base', c_int), ('col_head', c_int), ('mn_major', c_int), ('per_head_batch', c_int)]


class GemmDesc(Structure):
    _fields_ = [('a', GemmOperand), ('b', GemmOperand), ('M', c_int), ('N', c_int), ('K', c_int),
                ('batch', c_int), ('heads', c_int), ('out', c_void_p), ('out_bf16', c_int), ('out_ld', c_int),
                ('out_batch_stride', c_longlong), ('out_head_stride', c_longlong), ('accumulate', c_int),
                ('split_k', c_int), ('alpha', c_float)]


class WgradDesc(Structure):
    _fields_ = [('dy', c_void_p), ('dy_C', c_int), ('dy_c0', c_int),
                ('x', c_void_p), ('x_C', c_int), ('x_H', c_int), ('x_W', c_int), ('x_planes', c_int), ('x_c0', c_int),
                ('B', c_int), ('Ho', c_int), ('Wo', c_int), ('Cout', c_int), ('Cin', c_int), ('ntaps', c_int),
                ('taps', c_int8 * 36), ('dw', c_void_p), ('dw_co_stride', c_longlong), ('dw_ci_stride', c_longlong),
                ('dw_tap_stride', c_longlong), ('scratch', c_void_p), ('scratch_bytes', c_longlong)]


class GnBwdDesc(Structure):
    _fields_ = [('g', c_void_p),
                ('x0', c_void_p), ('C0', c_int), ('stats0', c_void_p),
                ('x1', c_void_p), ('C1', c_int), ('stats1', c_void_p),
                ('B', c_int), ('HW', c_int), ('W', c_int), ('groups', c_int),
                ('gamma', c_void_p), ('beta', c_void_p), ('eps', c_float),
                ('scale', c_void_p), ('shift', c_void_p), ('ss_ld', c_int),
                ('apply_silu', c_int), ('resample', c_int),
                ('drop_p', c_float), ('drop_seed', ctypes.c_ulonglong), ('drop_seed_dev', c_void_p),
                ('sums', c_void_p),
                ('dx0', c_void_p), ('dx0_accumulate', c_int),
                ('dx1', c_void_p), ('dx1_accumulate', c_int),
                ('addend', c_void_p), ('dx_bf16', c_void_p), ('dx_rowsum', c_void_p), ('dx_rowsum_ld', c_int),
                ('dx_colsum', c_void_p), ('dgamma', c_void_p), ('dbeta', c_void_p),
                ('dscale', c_void_p), ('dshift', c_void_p), ('dss_ld', c_int)]


class OdeDesc(Structure):
    _fields_ = [('model_out', c_void_p), ('x', c_void_p), ('d1', c_void_p), ('x1', c_void_p),
                ('B', c_int), ('C', c_int), ('Cm', c_int), ('HW', c_int),
                ('objective', c_int), ('clip', c_int), ('second_order', c_int),
                ('sqrt_recip_ac', c_float), ('sqrt_recipm1_ac', c_float), ('sqrt_ac', c_float), ('sqrt_1m_ac', c_float),
                ('sigma_t', c_float), ('sigma_prev', c_float),
                ('sample', c_void_p), ('pred_x0', c_void_p), ('deriv', c_void_p)]


class OptimDesc(Structure):
    _fields_ = [('chunks', c_void_p), ('n_chunks', c_int), ('lr', c_float), ('beta1', c_float), ('beta2', c_float),
                ('eps', c_float), ('weight_decay', c_float), ('adamw', c_int), ('step', c_int),
                ('max_grad_norm', c_float), ('want_norm', c_int), ('gnorm_sq', c_void_p), ('ema_decay', c_float),
                ('dev_state', c_void_p)]


_lib = None


def lib():
    """Loads the shared library (once).  Raises if it has not been built: there is no CPU/PyTorch fallback."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f'{LIB_PATH} not found: build it with `python -c "import __graft_entry__ as g; g.build()"` '
            f'(or `make -C diffusion-models-pytorch_b200/csrc`). There is no fallback path.')
    L = ctypes.CDLL(LIB_PATH)
    L.b200_version.restype = c_int
    L.b200_last_error.restype = c_char_p
    L.b200_launch_count.restype = c_longlong
    L.b200_conv2d_fwd.argtypes = [POINTER(ConvDesc), c_void_p]
    L.b200_conv2d_gn_fwd.argtypes = [POINTER(ConvDesc), POINTER(GnFuseDesc), c_void_p]
    L.b200_conv2d_gn_fwd.restype = c_int
    L.b200_attn_block_fwd.argtypes = [POINTER(AttnBlockDesc), c_void_p]
    L.b200_attn_block_fwd.restype = c_int
    L.b200_conv3x3_first.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_void_p]
    L.b200_first_split.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    L.b200_first_split.restype = c_int
    L.b200_groupnorm_apply_fwd.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int,
                                           c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int,
                                           c_int, c_int, c_void_p, c_void_p, c_void_p]
    L.b200_groupnorm_silu_fwd.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                          c_void_p, c_float, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p,
                                          c_void_p, c_void_p]
    L.b200_cast_bf16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_void_p]
    L.b200_avgpool2_f32.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    L.b200_upsample2_f32.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    L.b200_attention_fwd.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                     c_int, c_float, c_void_p]
    L.b200_attention_fwd_lse.argtypes = [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_int, c_int, c_int, c_int,
                                         c_int, c_float, c_void_p, c_void_p]
    L.b200_attention_bwd.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                     c_int, c_int, c_int, c_int, c_float, c_void_p]
    L.b200_attention_fwd_lse.restype = c_int
    L.b200_attention_bwd.restype = c_int
    L.b200_time_embed.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                  c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]
    L.b200_sampler_step.argtypes = [POINTER(SamplerDesc), c_void_p]
    L.b200_diffuse.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    ull = ctypes.c_ulonglong
    L.b200_groupnorm_apply_train_fwd.argtypes = [c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int,
                                                 c_int, c_int, c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p,
                                                 c_int, c_int, c_int, c_float, ull, c_void_p, c_void_p, c_void_p, c_void_p]
    L.b200_dropout_mask.argtypes = [c_void_p, c_longlong, c_float, ull, c_void_p]
    L.b200_groupnorm_bwd.argtypes = [POINTER(GnBwdDesc), c_void_p]
    L.b200_cast_bf16_colsum.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_void_p]
    L.b200_nchw_to_nhwc_pad_bf16.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    L.b200_colsum_bf16.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_int, c_int, c_void_p]
    L.b200_resample_f32.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p]
    L.b200_upsample2_bf16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]
    L.b200_softmax_rows.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p]
    L.b200_softmax_bwd_rows.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p]
    L.b200_time_embed_bwd.argtypes = [c_void_p, c_int, c_void_p, c_int, c_int, c_int] + [c_void_p] * 14
    L.b200_optimizer_step.argtypes = [POINTER(OptimDesc), c_void_p]
    L.b200_ode_step.argtypes = [POINTER(OdeDesc), c_void_p]
    L.b200_pack_weights.argtypes = [c_void_p, c_int, c_void_p]
    L.b200_to_uint8_hwc.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]
    L.b200_mse_loss.argtypes = [c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]
    L.b200_mse_loss_grad.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_longlong, c_void_p]
    for name in BACKWARD_SYMBOLS:
        getattr(L, name).restype = c_int
    L.b200_split_cast.argtypes = [POINTER(SplitDesc), c_void_p]
    L.b200_groupnorm_apply_split_fwd.argtypes = [c_void_p, c_int, c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                                 c_int, c_void_p, c_void_p, c_float, c_void_p, c_void_p, c_int, c_int,
                                                 c_int, c_void_p, c_void_p, c_void_p]
    L.b200_softmax_rows_split.argtypes = [c_void_p, c_void_p, c_longlong, c_int, c_float, c_void_p]
    for name in PRECISE_SYMBOLS:
        getattr(L, name).restype = c_int
    L.b200_gemm_batched.argtypes = [POINTER(GemmDesc), c_void_p]
    L.b200_conv2d_wgrad.argtypes = [POINTER(WgradDesc), c_void_p]
    for name in ('b200_gemm_batched', 'b200_conv2d_wgrad', 'b200_conv2d_fwd', 'b200_conv3x3_first', 'b200_groupnorm_silu_fwd', 'b200_groupnorm_apply_fwd',
                 'b200_cast_bf16',
                 'b200_avgpool2_f32', 'b200_upsample2_f32', 'b200_attention_fwd', 'b200_time_embed',
                 'b200_sampler_step', 'b200_diffuse'):
        getattr(L, name).restype = c_int
    _lib = L
    return L


BACKWARD_SYMBOLS = (
    'b200_groupnorm_apply_train_fwd', 'b200_dropout_mask', 'b200_groupnorm_bwd', 'b200_cast_bf16_colsum',
    'b200_nchw_to_nhwc_pad_bf16', 'b200_colsum_bf16', 'b200_resample_f32', 'b200_upsample2_bf16', 'b200_softmax_rows',
    'b200_softmax_bwd_rows', 'b200_mse_loss', 'b200_mse_loss_grad', 'b200_time_embed_bwd',
    'b200_optimizer_step', 'b200_ode_step', 'b200_to_uint8_hwc', 'b200_pack_weights',
)

PRECISE_SYMBOLS = ('b200_split_cast', 'b200_groupnorm_apply_split_fwd', 'b200_softmax_rows_split')

EXPORTED_SYMBOLS = BACKWARD_SYMBOLS + PRECISE_SYMBOLS + (
    'b200_version', 'b200_last_error', 'b200_launch_count', 'b200_conv2d_fwd', 'b200_conv2d_gn_fwd', 'b200_attn_block_fwd', 'b200_conv3x3_first',
    'b200_first_split', 'b200_groupnorm_silu_fwd', 'b200_groupnorm_apply_fwd', 'b200_cast_bf16', 'b200_avgpool2_f32', 'b200_upsample2_f32', 'b200_attention_fwd',
    'b200_time_embed', 'b200_sampler_step', 'b200_diffuse', 'b200_gemm_batched', 'b200_conv2d_wgrad',
    'b200_attention_fwd_lse', 'b200_attention_bwd',
)


GRAPH_LAUNCHES = 0   # kernels executed through CUDA-graph replays (captured once, replayed per timestep)


def launch_count() -> int:
    """Kernels of this library launched so far: direct launches plus launches replayed from captured graphs."""
    return int(lib().b200_launch_count()) + GRAPH_LAUNCHES


def direct_launch_count() -> int:
    return int(lib().b200_launch_count())


class Profiler:
    """Per-launch CUDA-event timing of the library's kernels on the current stream (bench.py's roofline leg).
    Usage: `with K.Profiler() as prof: model(x, t)`; `prof.summary()` -> {kind: dict(n, ms, flops, bytes)}."""

    def __init__(self):
        self.records = []

    def __enter__(self):
        global _PROFILER
        _PROFILER = self
        return self

    def __exit__(self, *exc):
        global _PROFILER
        _PROFILER = None

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for kind, flops, nbytes, e0, e1 in self.records:
            r = out.setdefault(kind, dict(n=0, ms=0.0, flops=0.0, bytes=0.0))
            r['n'] += 1
            r['ms'] += e0.elapsed_time(e1)
            r['flops'] += flops
            r['bytes'] += nbytes
        return out


_PROFILER = None


def _launch(kind, call, flops=0.0, nbytes=0.0):
    if _PROFILER is None:
        call()
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    call()
    e1.record()
    _PROFILER.records.append((kind, flops, nbytes, e0, e1))


def _check(rc: int, what: str):
    if rc != 0:
        msg = lib().b200_last_error().decode('utf-8', 'replace')
        raise RuntimeError(f'libb200diff {what} failed (code {rc}): {msg}')


_raw_stream = getattr(torch._C, '_cuda_getCurrentRawStream', None)


def _stream() -> int:
    """Raw cudaStream_t of torch's current stream (the fast private accessor when available: this is called once per
    kernel launch, ~800 times per training step)."""
    if _raw_stream is not None:
        return _raw_stream(torch.cuda.current_device())
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return None if t is None else t.data_ptr()


def _need_stats(*ts):
    for t in ts:
        if t is not None and (t.dtype != torch.int64 or not t.is_contiguous()):
            raise RuntimeError('GroupNorm statistics buffers are contiguous int64 [B, C, 2] tensors (b200diff.new_stats); '
                               f'got {t.dtype}')


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError('b200diff kernels run on CUDA tensors only (no CPU fallback); got a '
                               f'{t.device} tensor')


# --------------------------------------------------------------------------------------------------
# tap tables
# --------------------------------------------------------------------------------------------------
_T3 = [[(s - 1, r - 1, 0) for r in range(3) for s in range(3)]]
_T1 = [[(0, 0, 0)]]


def taps_3x3_s1():
    """3x3 stride 1 pad 1: tap k = r*3+s reads (ho + r - 1, wo + s - 1)."""
    return _T3


def taps_1x1():
    return _T1


def taps_3x3_s2(pad_lo: int):
    """3x3 stride 2 over 4 parity planes.  pad_lo = 1: symmetric pad 1 (models/modules.py:72);
    pad_lo = 0: pad (0,1,0,1) then stride 2 (models/pesser/model.py:66-70).
    Input row 2*ho + r - pad_lo lives in plane parity (r - pad_lo) & 1 at index ho + floor((r - pad_lo) / 2)."""
    taps = []
    for r in range(3):
        for s in range(3):
            ry, rx = r - pad_lo, s - pad_lo
            taps.append((rx // 2, ry // 2, ((ry & 1) << 1) | (rx & 1)))
    return [taps]


def taps_up2_3x3():
    """nearest-2x followed by 3x3 pad 1 == four 2x2-tap convolutions on the low-resolution grid.
    Phase (a, b) = output pixel parity; tap (i, j) in {0,1}^2 reads low-res offset (dh, dw) with
    dh = i - 1 + a, dw = j - 1 + b.  The matching weights are sums of the 3x3 taps (see pack_weight_up2)."""
    out = []
    for a in range(2):
        for b in range(2):
            out.append([(j - 1 + b, i - 1 + a, 0) for i in range(2) for j in range(2)])
    return out


_TAPS_CACHE = {}


def _fill_taps(desc, taps0, tap1):
    key = (tuple(tuple(t) for t in taps0), tuple(tap1))
    hit = _TAPS_CACHE.get(key)
    if hit is None:
        flat = [0] * 144
        for ph, taps in enumerate(taps0):
            for k, (dw, dh, pl) in enumerate(taps):
                base = (ph * 9 + k) * 4
                flat[base], flat[base + 1], flat[base + 2] = dw, dh, pl
        hit = _TAPS_CACHE[key] = ((c_int8 * 144)(*flat), (c_int8 * 4)(tap1[0], tap1[1], tap1[2], 0))
    desc.taps0, desc.tap1 = hit


# --------------------------------------------------------------------------------------------------
# weight packing (host side, once per weight update): OIHW fp32 -> [Cout][K] bf16, K = tap-major, channel-minor
# --------------------------------------------------------------------------------------------------
def pack_weight(w: torch.Tensor, shortcut_w: torch.Tensor = None) -> torch.Tensor:
    """[Cout, Cin, kh, kw] -> bf16 [Cout, kh*kw*Cin (+ Cin_sc)] with k = (r*kw + s)*Cin + c."""
    co, ci, kh, kw = w.shape
    m = w.detach().permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
    if shortcut_w is not None:
        m = torch.cat([m, shortcut_w.detach().reshape(co, -1)], dim=1)
    return m.to(torch.bfloat16).contiguous()


def pack_weight_up2(w: torch.Tensor, split: bool = False) -> torch.Tensor:
    """3x3 weights -> the four 2x2 phase kernels of 'nearest-2x then 3x3' stacked as [4*Cout][4*Cin] bf16.
    Row taps: phase a=0 -> {W[0], W[1]+W[2]}, a=1 -> {W[0]+W[1], W[2]}; same for columns (summed in fp32).
    split=True (FP32 mode): [4*Cout][4*3*Cin], per tap [w_hi | w_hi | w_lo] of the fp32 phase weights."""
    co, ci, _, _ = w.shape
    w = w.detach().float()
    rows = {0: [w[:, :, 0:1], w[:, :, 1:2] + w[:, :, 2:3]], 1: [w[:, :, 0:1] + w[:, :, 1:2], w[:, :, 2:3]]}
    mats = []
    for a in range(2):
        for b in range(2):
            taps = []
            for i in range(2):
                wr = rows[a][i]  # [co, ci, 1, 3]
                cols = {0: [wr[..., 0], wr[..., 1] + wr[..., 2]], 1: [wr[..., 0] + wr[..., 1], wr[..., 2]]}
                for j in range(2):
                    t32 = cols[b][j].reshape(co, ci)
                    if split:
                        hi = t32.to(torch.bfloat16)
                        lo = (t32 - hi.float()).to(torch.bfloat16)
                        taps += [hi, hi, lo]
                    else:
                        taps.append(t32)
            mats.append(torch.cat(taps, dim=1))
    return torch.cat(mats, dim=0).to(torch.bfloat16).contiguous()


# --------------------------------------------------------------------------------------------------
# kernel wrappers
# --------------------------------------------------------------------------------------------------
def conv2d(a0, w_packed, N, B, Ho, Wo, taps0, *, a0_geom, a1=None, a1_geom=None, tap1=(0, 0, 0), bias=None,
           rowadd=None, rowadd_ld=0, residual=None, res_ld=0, out=None, out_mode=OUT_F32_NHWC, out_ld=None,
           out_H=None, out_W=None, w_rows_per_phase=None, stats=None, alg_macs=None):
    """a0_geom = (C, H, W, planes) of the bf16 source tensor [B][planes][H][W][C]."""
    _need_cuda(a0, w_packed, out)
    _need_stats(stats)
    d = ConvDesc()
    d.a0 = a0.data_ptr()
    d.a0_C, d.a0_H, d.a0_W, d.a0_planes = a0_geom
    if a1 is not None:
        d.a1 = a1.data_ptr()
        d.a1_C, d.a1_H, d.a1_W, d.a1_planes = a1_geom
    d.w = w_packed.data_ptr()
    d.w_rows, d.w_K = w_packed.shape
    phases = len(taps0)
    d.w_rows_per_phase = w_rows_per_phase if w_rows_per_phase is not None else N
    d.B, d.Ho, d.Wo = B, Ho, Wo
    d.phases, d.N, d.ntaps0 = phases, N, len(taps0[0])
    _fill_taps(d, taps0, tap1)
    d.bias = _ptr(bias)
    d.rowadd, d.rowadd_ld = _ptr(rowadd), rowadd_ld
    d.residual, d.res_ld = _ptr(residual), res_ld
    d.stats = _ptr(stats)
    d.out, d.out_mode = out.data_ptr(), out_mode
    d.out_ld = out_ld if out_ld is not None else N
    up = 2 if phases == 4 else 1
    d.out_H = out_H if out_H is not None else Ho * up
    d.out_W = out_W if out_W is not None else Wo * up
    d.osy = d.osx = up
    # algorithmic work = MACs of the reference convolution (alg_macs) or, by default, the MACs executed
    macs = alg_macs if alg_macs is not None else float(phases) * B * Ho * Wo * N * d.w_K
    _launch('conv_gemm', lambda: _check(lib().b200_conv2d_fwd(ctypes.byref(d), _stream()), 'conv2d_fwd'),
            flops=2.0 * macs)
    return out


def conv2d_gn_ok(B, Ho, Wo, N, groups) -> bool:
    """Static eligibility of a layer for the fused conv + next-GroupNorm entry (mirrors the checks of
    b200_conv2d_gn_fwd, so the engine decides per layer up front instead of catching a rejected launch)."""
    hw = Ho * Wo
    if N % 128 != 0 or groups < 1 or N % groups != 0 or (_GN_FUSE_HW is not None and hw not in _GN_FUSE_HW):
        return False
    cpg = N // groups
    if cpg > 32 or cpg & (cpg - 1):
        return False
    if hw in (512, 1024):        # multi-tile variant: 2 / 4 co-scheduled CTAs share an image; 128 output channels, full-width 256-pixel tiles
        return _GN_CLUSTER and N == 128 and Wo <= 256 and 256 % Wo == 0 and Ho % (256 // Wo) == 0
    if hw not in (16, 64, 256):
        return False
    return hw >= 64 or B % (64 // hw) == 0      # the smallest tile is 64 pixels = 4 whole 4x4 images


_GN_FUSE_HW = ({int(v) for v in os.environ['B200_FUSE_GN2_HW'].split(',') if v} if 'B200_FUSE_GN2_HW' in os.environ
               else None)   # experiment knob: image sizes (pixels) allowed to take the fused entry, e.g. "16,256,1024"
_GN_CLUSTER = os.environ.get('B200_FUSE_GN2_CLUSTER', '1') != '0'   # =0: 32x32 layers keep conv + GroupNorm launches (A/B)


def conv2d_gn_needs_workspace(Ho, Wo) -> bool:
    """Images of 512 / 1024 pixels span several tiles: the fused entry then needs zeroed `xstats` / `xcount` buffers."""
    return Ho * Wo in (512, 1024)


def conv2d_gn(a0, w_packed, N, B, Ho, Wo, taps0, *, a0_geom, gamma, beta, groups, eps, out_norm, bias=None, rowadd=None,
              rowadd_ld=0, scale=None, shift=None, ss_ld=0, silu=True, xstats=None, xcount=None, out=None, stats=None,
              residual=None, res_ld=0, a1=None, a1_geom=None, tap1=(0, 0, 0), out_norm_ld=0, out_raw=None, alg_macs=None):
    """b200_conv2d_gn_fwd: out_norm = SiLU(GN(conv(a0) + bias + rowadd)) as the bf16 NHWC operand of the next convolution.
    Eligibility (`conv2d_gn_ok`): a tile must hold whole images (Ho*Wo in {16, 64, 256}; 512 / 1024 with workspaces),
    N % 128 == 0, power-of-two channels per group <= 32.
    out=None: nothing else is written (conv1 -> norm2; or a block output whose fp32 form nobody reads: a1 = fused 1x1
    shortcut, out_raw = its bf16 copy at the consumer's concat stride).  out = fp32 NHWC tensor: block-output form -- x = conv + bias
    (+ residual) goes to `out` with its statistics in `stats` (multi-tile images: in `xstats`), GN(x) to out_norm;
    a1 = second source (the fused 1x1 shortcut's K-blocks)."""
    _need_cuda(a0, w_packed, out_norm)
    if out is not None:
        _need_cuda(out)
        if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() != B * Ho * Wo * N:
            raise RuntimeError('conv2d_gn: out must be a contiguous float32 [B, Ho, Wo, N] tensor')
        if residual is not None and (residual.dtype != torch.float32 or not residual.is_cuda):
            raise RuntimeError('conv2d_gn: residual must be a float32 CUDA tensor')
        _need_stats(stats)
    elif residual is not None or stats is not None:
        raise RuntimeError('conv2d_gn: residual / stats need the fp32 output `out`')
    d = ConvDesc()
    d.a0 = a0.data_ptr()
    d.a0_C, d.a0_H, d.a0_W, d.a0_planes = a0_geom
    d.w = w_packed.data_ptr()
    d.w_rows, d.w_K = w_packed.shape
    d.w_rows_per_phase = N
    d.B, d.Ho, d.Wo = B, Ho, Wo
    d.phases, d.N, d.ntaps0 = 1, N, len(taps0[0])
    _fill_taps(d, taps0, tap1)
    if a1 is not None:
        _need_cuda(a1)
        d.a1 = a1.data_ptr()
        d.a1_C, d.a1_H, d.a1_W, d.a1_planes = a1_geom
    d.bias = _ptr(bias)
    d.rowadd, d.rowadd_ld = _ptr(rowadd), rowadd_ld
    if out is None:
        d.out, d.out_mode, d.out_ld = None, OUT_BF16_NHWC, N
    else:
        d.out, d.out_mode, d.out_ld = out.data_ptr(), OUT_F32_NHWC, N
        d.residual, d.res_ld = _ptr(residual), res_ld
        d.stats = _ptr(stats)
    d.out_H, d.out_W, d.osy, d.osx = Ho, Wo, 1, 1
    g = GnFuseDesc()
    g.gamma, g.beta, g.scale, g.shift = _ptr(gamma), _ptr(beta), _ptr(scale), _ptr(shift)
    g.out_norm, g.ss_ld, g.groups, g.apply_silu, g.eps = out_norm.data_ptr(), ss_ld, groups, int(silu), float(eps)
    g.out_norm_ld = out_norm_ld
    if out_raw is not None:
        _need_cuda(out_raw)
        if out_raw.dtype != torch.bfloat16 or (out is None and (rowadd is not None or scale is not None)):
            raise RuntimeError('conv2d_gn: out_raw must be bf16 and, without `out`, takes no embedding row / scale / shift')
    g.out_raw_bf16 = _ptr(out_raw)
    _need_stats(xstats, xcount)
    if conv2d_gn_needs_workspace(Ho, Wo) and (xstats is None or xcount is None or xstats.numel() < B * N * 2
                                              or xcount.numel() < B):
        raise RuntimeError('conv2d_gn: images of 512 / 1024 pixels need zeroed int64 workspaces xstats [B, N, 2] and xcount [B]')
    g.xstats, g.xcount = _ptr(xstats), _ptr(xcount)
    _launch('conv_gemm', lambda: _check(lib().b200_conv2d_gn_fwd(ctypes.byref(d), ctypes.byref(g), _stream()),
                                        'conv2d_gn_fwd'),
            flops=2.0 * (alg_macs if alg_macs is not None else float(B) * Ho * Wo * N * d.w_K))
    return out_norm


def conv3x3_first(x, w, bias, out, stats=None):
    _need_cuda(x, w, out)
    _need_stats(stats)
    B, Cin, H, W = x.shape
    Cout = w.shape[0]
    _launch('conv3x3_first',
            lambda: _check(lib().b200_conv3x3_first(x.data_ptr(), w.data_ptr(), _ptr(bias), out.data_ptr(),
                                                    _ptr(stats), B, Cin, H, W, Cout, _stream()), 'conv3x3_first'),
            flops=2.0 * B * H * W * Cout * Cin * 9, nbytes=4.0 * B * H * W * (Cin + Cout))
    return out


def first_split(x, out):
    """b200_first_split: fp32 NCHW image -> bf16 NHWC [B, H, W, 64] pixels [hi | lo | hi | 0 ...], the tensor-core operand
    of the first convolution (weights: pack mode 4, tap_ld = 64)."""
    _need_cuda(x, out)
    B, Cin, H, W = x.shape
    if x.dtype != torch.float32 or not x.is_contiguous():
        raise RuntimeError('first_split: x must be a contiguous float32 [B, Cin, H, W] tensor')
    if out.dtype != torch.bfloat16 or not out.is_contiguous() or out.numel() != B * H * W * 64:
        raise RuntimeError('first_split: out must be a contiguous bfloat16 [B, H, W, 64] tensor')
    _launch('first_split', lambda: _check(lib().b200_first_split(x.data_ptr(), out.data_ptr(), B, Cin, H, W, _stream()),
                                          'first_split'), nbytes=4.0 * x.numel() + 2.0 * out.numel())
    return out


def groupnorm_silu(x0, C0, x1, C1, B, HW, W, groups, gamma, beta, eps, out, *, scale=None, shift=None, ss_ld=0,
                   silu=True, resample=0, raw_out=None):
    _need_cuda(x0, out)
    _launch('groupnorm_slab',
            lambda: _check(lib().b200_groupnorm_silu_fwd(x0.data_ptr(), C0, _ptr(x1), C1, B, HW, W, groups, _ptr(gamma),
                                                         _ptr(beta), float(eps), _ptr(scale), _ptr(shift), ss_ld,
                                                         int(silu), resample, out.data_ptr(), _ptr(raw_out),
                                                         _stream()), 'groupnorm_silu_fwd'),
            nbytes=_gn_bytes(B, HW, C0 + (C1 if x1 is not None else 0), resample, raw_out is not None))
    return out


def _gn_bytes(B, HW, C, resample, raw):
    """Algorithmic HBM bytes of GroupNorm+SiLU: fp32 read once, bf16 written once (+ bf16 raw copy)."""
    out_elems = B * HW * C * ({0: 1.0, 1: 0.25, 2: 4.0}[resample])
    return 4.0 * B * HW * C + 2.0 * out_elems + (2.0 * B * HW * C if raw else 0.0)


def groupnorm_apply(x0, C0, stats0, x1, C1, stats1, B, HW, W, groups, gamma, beta, eps, out, *, scale=None,
                    shift=None, ss_ld=0, silu=True, resample=0, raw_out=None, drop_p=0.0, drop_seed=0,
                    drop_seed_dev=None):
    """Streaming GroupNorm(+SiLU)(+dropout) for inputs whose [B][C][2] statistics came from the producing kernel.
    x0=None with C0 > 0: window mode -- channels [0, C0) of `out` / `raw_out` were already written by the first source's
    producer (conv2d_gn block-output form); only the second source's channels [C0, C0 + C1) are normalised here."""
    if x0 is None:
        if x1 is None or C0 <= 0:
            raise RuntimeError('groupnorm_apply: window mode (x0=None) needs C0 > 0 and a second source')
        _need_cuda(x1, stats1, out)
        _need_stats(stats1)
        _launch('groupnorm_apply',
                lambda: _check(lib().b200_groupnorm_apply_train_fwd(
                    None, 0, C0, None, x1.data_ptr(), C1, stats1.data_ptr(), B, HW, W, groups, _ptr(gamma), _ptr(beta),
                    float(eps), _ptr(scale), _ptr(shift), ss_ld, int(silu), resample, 0.0, 0, None, out.data_ptr(),
                    _ptr(raw_out), _stream()), 'groupnorm_apply_fwd'),
                nbytes=_gn_bytes(B, HW, C1, resample, raw_out is not None))
        return out
    _need_cuda(x0, stats0, out)
    _need_stats(stats0, stats1)
    _launch('groupnorm_apply',
            lambda: _check(lib().b200_groupnorm_apply_train_fwd(
                x0.data_ptr(), int(x0.dtype == torch.bfloat16), C0, stats0.data_ptr(), _ptr(x1), C1, _ptr(stats1), B,
                HW, W, groups, _ptr(gamma), _ptr(beta), float(eps), _ptr(scale), _ptr(shift), ss_ld, int(silu),
                resample, float(drop_p), int(drop_seed), _ptr(drop_seed_dev), out.data_ptr(), _ptr(raw_out), _stream()),
                'groupnorm_apply_fwd'),
            nbytes=_gn_bytes(B, HW, C0 + (C1 if x1 is not None else 0), resample, raw_out is not None) -
            (2.0 * B * HW * C0 if x0.dtype == torch.bfloat16 else 0.0))
    return out


# --------------------------------------------------------------------------------------------------
# FP32 mode ("bf16x3", csrc/precise.cu): split operands
# --------------------------------------------------------------------------------------------------
SPLIT_ACT, SPLIT_WEIGHT = 0, 1    # [hi | lo | hi] (activation side) / [hi | hi | lo] (weight side)


def split_cast(x, out, rows, C, *, in_ld=None, in_col0=0, group=None, pattern=SPLIT_ACT, silu=False, out_ld=None,
               out_col0=0, planes_rows=0, parity_hw=None):
    """fp32 [rows][in_ld] window -> bf16 split operand (b200_split_cast)."""
    _need_cuda(x, out)
    _need_f32(x=x)
    if out.dtype != torch.bfloat16 or not out.is_contiguous():
        raise RuntimeError('split_cast: out must be a contiguous bfloat16 tensor')
    d = SplitDesc()
    d.in_, d.rows, d.in_ld, d.in_col0, d.C = x.data_ptr(), rows, (in_ld if in_ld is not None else C), in_col0, C
    d.group, d.pattern, d.act = (group if group is not None else C), pattern, int(silu)
    d.out, d.out_ld, d.out_col0 = out.data_ptr(), (out_ld if out_ld is not None else 3 * C), out_col0
    d.planes_rows = planes_rows
    d.parity_H, d.parity_W = parity_hw if parity_hw is not None else (0, 0)
    _launch('split_cast', lambda: _check(lib().b200_split_cast(ctypes.byref(d), _stream()), 'split_cast'),
            nbytes=10.0 * rows * C)
    return out


def groupnorm_apply_split(x0, C0, stats0, x1, C1, stats1, B, HW, W, groups, gamma, beta, eps, out, *, scale=None,
                          shift=None, ss_ld=0, silu=True, resample=0, raw_out=None):
    """FP32-mode GroupNorm(+AdaGN)(+SiLU)(+resample): out = split operand [B][HW_out][3C] (b200_groupnorm_apply_split_fwd)."""
    _need_cuda(x0, stats0, out)
    _need_stats(stats0, stats1)
    _need_f32(x0=x0, x1=x1, gamma=gamma, beta=beta)
    C = C0 + (C1 if x1 is not None else 0)
    _launch('groupnorm_apply', lambda: _check(lib().b200_groupnorm_apply_split_fwd(
        x0.data_ptr(), C0, stats0.data_ptr(), _ptr(x1), C1, _ptr(stats1), B, HW, W, groups, _ptr(gamma), _ptr(beta),
        float(eps), _ptr(scale), _ptr(shift), ss_ld, int(silu), resample, out.data_ptr(), _ptr(raw_out), _stream()),
        'groupnorm_apply_split_fwd'),
        nbytes=4.0 * B * HW * C + 6.0 * B * HW * C * ({0: 1.0, 1: 0.25, 2: 4.0}[resample]) +
        (6.0 * B * HW * C if raw_out is not None else 0.0))
    return out


def softmax_rows_split(S, P, rows, T, scale):
    _need_cuda(S, P)
    _need_f32(S=S)
    _check(lib().b200_softmax_rows_split(S.data_ptr(), P.data_ptr(), rows, T, float(scale), _stream()), 'softmax_rows_split')


def cast_bf16(x, out, B, H, W, C, parity_split=False):
    _need_cuda(x, out)
    _launch('cast_bf16', lambda: _check(lib().b200_cast_bf16(x.data_ptr(), out.data_ptr(), B, H, W, C,
                                                             int(parity_split), _stream()), 'cast_bf16'),
            nbytes=6.0 * B * H * W * C)
    return out


def avgpool2_f32(x, out, B, H, W, C):
    _check(lib().b200_avgpool2_f32(x.data_ptr(), out.data_ptr(), B, H, W, C, _stream()), 'avgpool2_f32')
    return out


def upsample2_f32(x, out, B, H, W, C):
    _check(lib().b200_upsample2_f32(x.data_ptr(), out.data_ptr(), B, H, W, C, _stream()), 'upsample2_f32')
    return out


def attention(qk, ld_qk, q_off, k_off, vt, out, ld_out, B, T, heads, d, scale, lse=None):
    """lse: optional fp32 [B, heads, T] output (log2-sum-exp of the scaled score rows) for `attention_bwd`."""
    _need_cuda(qk, vt, out)
    if lse is not None:
        _need_cuda(lse)
        if lse.dtype != torch.float32 or not lse.is_contiguous() or lse.numel() != B * heads * T:
            raise RuntimeError('attention: lse must be a contiguous float32 tensor of B*heads*T elements')
    _launch('attention',
            lambda: _check(lib().b200_attention_fwd_lse(qk.data_ptr(), ld_qk, q_off, k_off, vt.data_ptr(), out.data_ptr(),
                                                        ld_out, B, T, heads, d, float(scale), _ptr(lse), _stream()),
                           'attention_fwd'),
            flops=4.0 * B * heads * T * T * d)
    return out


def attention_bwd_ok(T, d) -> bool:
    """Static eligibility for the one-launch attention adjoint (mirrors b200_attention_bwd's checks)."""
    return d == 64 and 8 <= T <= 256 and T % 8 == 0


def attention_bwd(qk, vt, o, d_o, lse, dqk, dv, B, T, heads, d, scale):
    """[dQ | dK] -> the first 2C columns of dqk (bf16 [B, T, ld]), dV -> dv (bf16 [B, T, ld']) from q|k [B, T, 2C],
    v^T [B, C, T], o, dO [B, T, C].  dqk / dv may be column windows of one [B, T, 3C] tensor."""
    _need_cuda(qk, vt, o, d_o, lse, dqk, dv)
    C = heads * d
    for name, t, n in (('qk', qk, 2 * B * T * C), ('vt', vt, B * T * C), ('o', o, B * T * C), ('d_o', d_o, B * T * C)):
        if t.dtype != torch.bfloat16 or not t.is_contiguous() or t.numel() != n:
            raise RuntimeError(f'attention_bwd: {name} must be a contiguous bfloat16 tensor of {n} elements')
    for name, t, n in (('dqk', dqk, 2 * C), ('dv', dv, C)):
        if (t.dtype != torch.bfloat16 or t.dim() != 3 or tuple(t.shape) != (B, T, n) or t.stride(2) != 1
                or t.stride(0) != T * t.stride(1)):
            raise RuntimeError(f'attention_bwd: {name} must be a bfloat16 [B, T, {n}] tensor or column window')
    if lse.dtype != torch.float32 or not lse.is_contiguous() or lse.numel() != B * heads * T:
        raise RuntimeError('attention_bwd: lse must be a contiguous float32 tensor of B*heads*T elements')
    _launch('attention_bwd',
            lambda: _check(lib().b200_attention_bwd(qk.data_ptr(), vt.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr(),
                                                    dqk.data_ptr(), dqk.stride(1), dv.data_ptr(), dv.stride(1), B, T, heads,
                                                    d, float(scale), _stream()),
                           'attention_bwd'),
            flops=10.0 * B * heads * T * T * d)
    return dqk, dv


def attn_block_ok(T, C, heads, groups) -> bool:
    """Static eligibility for the one-launch attention block (mirrors b200_attn_block_fwd's checks)."""
    return T == 256 and C == 256 and heads == 1 and groups == 32


def attn_block(x, x_stats, gamma, beta, eps, w, bias, out, out_stats, B, T, C, heads, groups, scale, dbg=None):
    """b200_attn_block_fwd: out = x + proj(softmax(q k^T scale) v) with q, k, v = 1x1 convs of GroupNorm(x), one launch.
    x / out fp32 [B, T, C]; w bf16 [4C, C] = [Wq; Wk; Wv; Wproj]; bias fp32 [4C].  `dbg`: up to six bf16 [B, 256, 256]
    tensors (or None) receiving xn, q, k, v^T, P, o (tests)."""
    _need_cuda(x, w, out)
    _need_stats(x_stats)
    _need_stats(out_stats)
    for t_, dt in ((x, torch.float32), (out, torch.float32), (w, torch.bfloat16), (bias, torch.float32),
                   (gamma, torch.float32), (beta, torch.float32)):
        if t_.dtype != dt or not t_.is_contiguous():
            raise RuntimeError(f'attn_block: expected contiguous {dt} tensors, got {t_.dtype} (contiguous={t_.is_contiguous()})')
    if tuple(w.shape) != (4 * C, C) or bias.numel() != 4 * C or x.numel() != B * T * C or out.numel() != B * T * C:
        raise RuntimeError('attn_block: shape mismatch')
    d = AttnBlockDesc()
    d.x, d.x_stats, d.gamma, d.beta = x.data_ptr(), x_stats.data_ptr(), gamma.data_ptr(), beta.data_ptr()
    d.w, d.bias, d.out, d.out_stats = w.data_ptr(), bias.data_ptr(), out.data_ptr(), _ptr(out_stats)
    for i in range(6):
        d.dbg[i] = _ptr(dbg[i]) if dbg is not None and i < len(dbg) else None
    d.B, d.T, d.C, d.heads, d.groups = B, T, C, heads, groups
    d.eps, d.scale = float(eps), float(scale)
    _launch('attn_block', lambda: _check(lib().b200_attn_block_fwd(ctypes.byref(d), _stream()), 'attn_block_fwd'),
            flops=2.0 * B * T * C * (4 * C + 2 * T), nbytes=8.0 * B * T * C)
    return out


def time_embed(t, freqs, dim, E, cos_first, w1, b1, w2, b2, out, *, y=None, class_embed=None, out_silu_bf16=None):
    _need_cuda(t, out)
    _launch('time_embed',
            lambda: _check(lib().b200_time_embed(t.data_ptr(), t.shape[0], freqs.data_ptr(), dim, E, int(cos_first),
                                                 w1.data_ptr(), b1.data_ptr(), w2.data_ptr(), b2.data_ptr(), _ptr(y),
                                                 _ptr(class_embed), out.data_ptr(), _ptr(out_silu_bf16), _stream()),
                           'time_embed'),
            flops=2.0 * t.shape[0] * (dim * E + E * E))
    return out


def sampler_step(model_out, xt, coef_row, *, objective='pred_eps', clip=True, learned_range=False, noise=None,
                 model_out_uncond=None, guidance_scale=1.0, sample=None, mean=None, pred_x0=None, pred_eps=None,
                 var_out=None):
    _need_cuda(model_out, xt, coef_row)
    _need_f32(model_out=model_out, xt=xt, coef_row=coef_row, noise=noise, model_out_uncond=model_out_uncond,
              sample=sample, mean=mean, pred_x0=pred_x0, pred_eps=pred_eps, var_out=var_out)
    for name, v in (('noise', noise), ('sample', sample), ('mean', mean), ('pred_x0', pred_x0), ('pred_eps', pred_eps)):
        if v is not None and v.shape != xt.shape:
            raise RuntimeError(f'sampler_step: {name} has shape {tuple(v.shape)}, expected {tuple(xt.shape)}')
    if model_out_uncond is not None and model_out_uncond.shape != model_out.shape:
        raise RuntimeError('sampler_step: model_out_uncond must have the shape of model_out')
    if model_out.shape[0] != xt.shape[0] or model_out.shape[2:] != xt.shape[2:] or model_out.shape[1] < xt.shape[1]:
        raise RuntimeError(f'sampler_step: model output {tuple(model_out.shape)} does not match x_t {tuple(xt.shape)}')
    d = SamplerDesc()
    d.model_out, d.model_out_uncond = model_out.data_ptr(), _ptr(model_out_uncond)
    d.xt, d.noise, d.coef = xt.data_ptr(), _ptr(noise), coef_row.data_ptr()
    B, C = xt.shape[0], xt.shape[1]
    d.B, d.C, d.Cm, d.HW = B, C, model_out.shape[1], xt[0, 0].numel()
    d.objective, d.clip, d.learned_range = OBJ[objective], int(clip), int(learned_range)
    d.guidance_scale = float(guidance_scale)
    d.sample, d.mean, d.pred_x0, d.pred_eps, d.var_out = _ptr(sample), _ptr(mean), _ptr(pred_x0), _ptr(pred_eps), \
        _ptr(var_out)
    n_streams = 2 + sum(v is not None for v in (noise, model_out_uncond, sample, mean, pred_x0, pred_eps, var_out))
    _launch('sampler_step', lambda: _check(lib().b200_sampler_step(ctypes.byref(d), _stream()), 'sampler_step'),
            nbytes=4.0 * xt.numel() * n_streams)


def _need_f32(**ts):
    """The kernels read raw pointers: a tensor of another dtype / layout / device would be silently misread."""
    dev = None
    for name, t in ts.items():
        if t is None:
            continue
        if t.dtype != torch.float32 or not t.is_contiguous():
            raise RuntimeError(f'{name} must be a contiguous float32 tensor, got dtype {t.dtype}, '
                               f'contiguous={t.is_contiguous()}')
        if dev is not None and t.device != dev:
            raise RuntimeError(f'{name} is on {t.device}, expected {dev}: all operands of a kernel live on one device')
        dev = t.device


def diffuse(x0, eps, t, alphas_cumprod, out):
    _need_cuda(x0, eps, t, alphas_cumprod, out)
    _need_f32(x0=x0, eps=eps, alphas_cumprod=alphas_cumprod, out=out)
    if t.dtype != torch.int64 or not t.is_contiguous() or t.numel() != x0.shape[0] or t.device != x0.device:
        raise RuntimeError('diffuse: t must be a contiguous int64 tensor [B] on the device of x0')
    if eps.shape != x0.shape or out.shape != x0.shape:
        raise RuntimeError('diffuse: x0, eps and out must have the same shape')
    _check(lib().b200_diffuse(x0.data_ptr(), eps.data_ptr(), t.data_ptr(), alphas_cumprod.data_ptr(), out.data_ptr(),
                              x0.shape[0], x0[0].numel(), alphas_cumprod.numel(), _stream()), 'diffuse')
    return out


# --------------------------------------------------------------------------------------------------
# backward-pass kernels
# --------------------------------------------------------------------------------------------------
def _operand(t, rows, ld, *, col_base=0, col_head=0, mn_major=False, per_head_batch=False, batch_stride=0):
    o = GemmOperand()
    o.ptr, o.rows, o.ld, o.batch_stride = t.data_ptr(), rows, ld, batch_stride
    o.col_base, o.col_head, o.mn_major, o.per_head_batch = col_base, col_head, int(mn_major), int(per_head_batch)
    return o


def gemm_batched(a, b, out, M, N, Kdim, *, batch=1, heads=1, out_ld=None, out_batch_stride=0, out_head_stride=0,
                 accumulate=False, split_k=1, alpha=1.0):
    """a, b: (tensor, rows, ld, kwargs of _operand).  out: fp32 or bf16 tensor."""
    d = GemmDesc()
    d.a = _operand(a[0], a[1], a[2], **(a[3] if len(a) > 3 else {}))
    d.b = _operand(b[0], b[1], b[2], **(b[3] if len(b) > 3 else {}))
    _need_cuda(a[0], b[0], out)
    d.M, d.N, d.K, d.batch, d.heads = M, N, Kdim, batch, heads
    d.out, d.out_bf16 = out.data_ptr(), int(out.dtype == torch.bfloat16)
    d.out_ld = out_ld if out_ld is not None else N
    d.out_batch_stride, d.out_head_stride = out_batch_stride, out_head_stride
    d.accumulate, d.split_k, d.alpha = int(accumulate), split_k, float(alpha)
    _launch('gemm_batched', lambda: _check(lib().b200_gemm_batched(ctypes.byref(d), _stream()), 'gemm_batched'),
            flops=2.0 * batch * heads * M * N * Kdim)
    return out


def conv2d_wgrad(dy, dy_C, x, x_geom, B, Ho, Wo, Cout, Cin, taps, dw, *, x_c0=0, dy_c0=0, co_stride=None,
                 ci_stride=None, tap_stride=1, scratch=None):
    """Accumulates the weight gradient into `dw` (fp32; default strides = an OIHW tensor [Cout][Cin][ntaps]).
    x_geom = (C, H, W, planes) of the bf16 tensor the forward conv read, taps = its tap table (one phase)."""
    _need_cuda(dy, x, dw)
    d = WgradDesc()
    d.dy, d.dy_C, d.dy_c0 = dy.data_ptr(), dy_C, dy_c0
    d.x = x.data_ptr()
    d.x_C, d.x_H, d.x_W, d.x_planes = x_geom
    d.x_c0 = x_c0
    d.B, d.Ho, d.Wo, d.Cout, d.Cin, d.ntaps = B, Ho, Wo, Cout, Cin, len(taps)
    flat = [0] * 36
    for k, (dwx, dhy, pl) in enumerate(taps):
        flat[4 * k], flat[4 * k + 1], flat[4 * k + 2] = dwx, dhy, pl
    d.taps = (c_int8 * 36)(*flat)
    d.dw = dw.data_ptr()
    nt = len(taps)
    d.dw_co_stride = co_stride if co_stride is not None else Cin * nt
    d.dw_ci_stride = ci_stride if ci_stride is not None else nt
    d.dw_tap_stride = tap_stride
    if scratch is not None:
        d.scratch, d.scratch_bytes = scratch.data_ptr(), scratch.numel() * scratch.element_size()
    _launch('conv_wgrad', lambda: _check(lib().b200_conv2d_wgrad(ctypes.byref(d), _stream()), 'conv2d_wgrad'),
            flops=2.0 * B * Ho * Wo * Cout * Cin * nt)
    return dw


def groupnorm_bwd(g, x0, C0, stats0, x1, C1, stats1, B, HW, W, groups, gamma, beta, eps, sums, *, scale=None,
                  shift=None, ss_ld=0, silu=True, resample=0, drop_p=0.0, drop_seed=0, drop_seed_dev=None, dx0=None,
                  dx0_acc=False,
                  dx1=None, dx1_acc=False, addend=None, dx_bf16=None, dx_rowsum=None, dx_rowsum_ld=0, dx_colsum=None,
                  dgamma=None, dbeta=None, dscale=None, dshift=None, dss_ld=0):
    _need_cuda(g, x0, stats0, sums)
    _need_stats(stats0, stats1)
    d = GnBwdDesc()
    d.g, d.x0, d.C0, d.stats0 = g.data_ptr(), x0.data_ptr(), C0, stats0.data_ptr()
    d.x1, d.C1, d.stats1 = _ptr(x1), (C1 if x1 is not None else 0), _ptr(stats1)
    d.B, d.HW, d.W, d.groups = B, HW, W, groups
    d.gamma, d.beta, d.eps = _ptr(gamma), _ptr(beta), float(eps)
    d.scale, d.shift, d.ss_ld = _ptr(scale), _ptr(shift), ss_ld
    d.apply_silu, d.resample, d.drop_p, d.drop_seed = int(silu), resample, float(drop_p), int(drop_seed)
    d.drop_seed_dev = _ptr(drop_seed_dev)
    d.sums = sums.data_ptr()
    d.dx0, d.dx0_accumulate, d.dx1, d.dx1_accumulate = _ptr(dx0), int(dx0_acc), _ptr(dx1), int(dx1_acc)
    d.addend, d.dx_bf16, d.dx_rowsum = _ptr(addend), _ptr(dx_bf16), _ptr(dx_rowsum)
    d.dx_rowsum_ld, d.dx_colsum = dx_rowsum_ld, _ptr(dx_colsum)
    d.dgamma, d.dbeta, d.dscale, d.dshift, d.dss_ld = _ptr(dgamma), _ptr(dbeta), _ptr(dscale), _ptr(dshift), dss_ld
    C = C0 + (C1 if x1 is not None else 0)
    _launch('groupnorm_bwd', lambda: _check(lib().b200_groupnorm_bwd(ctypes.byref(d), _stream()), 'groupnorm_bwd'),
            nbytes=float(B) * HW * C * (12.0 + (2.0 if dx_bf16 is not None else 4.0)))


def dropout_mask(out, p, seed):
    _check(lib().b200_dropout_mask(out.data_ptr(), out.numel(), float(p), int(seed), _stream()), 'dropout_mask')
    return out


def cast_bf16_colsum(x, out, colsum, rows, C):
    _launch('grad_cast', lambda: _check(lib().b200_cast_bf16_colsum(x.data_ptr(), out.data_ptr(), _ptr(colsum), rows, C,
                                                                    _stream()), 'cast_bf16_colsum'),
            nbytes=6.0 * rows * C)
    return out


def nchw_to_nhwc_pad_bf16(x, out, colsum, B, C, HW, Cpad):
    _check(lib().b200_nchw_to_nhwc_pad_bf16(x.data_ptr(), out.data_ptr(), _ptr(colsum), B, C, HW, Cpad, _stream()),
           'nchw_to_nhwc_pad_bf16')
    return out


def colsum_bf16(x, colsum, rows, ld, c0, C):
    _check(lib().b200_colsum_bf16(x.data_ptr(), colsum.data_ptr(), rows, ld, c0, C, _stream()), 'colsum_bf16')


def resample_f32(x, out, B, H, W, C, mode, scale=1.0, accumulate=False):
    _check(lib().b200_resample_f32(x.data_ptr(), out.data_ptr(), B, H, W, C, mode, float(scale), int(accumulate),
                                   _stream()), 'resample_f32')
    return out


def upsample2_bf16(x, out, B, H, W, C):
    _check(lib().b200_upsample2_bf16(x.data_ptr(), out.data_ptr(), B, H, W, C, _stream()), 'upsample2_bf16')
    return out


def softmax_rows(S, P, rows, T, scale):
    _check(lib().b200_softmax_rows(S.data_ptr(), P.data_ptr(), rows, T, float(scale), _stream()), 'softmax_rows')


def softmax_bwd_rows(P, dP, dS, rows, T, scale):
    _check(lib().b200_softmax_bwd_rows(P.data_ptr(), dP.data_ptr(), dS.data_ptr(), rows, T, float(scale), _stream()),
           'softmax_bwd_rows')


def mse_loss(a, b, loss):
    _check(lib().b200_mse_loss(a.data_ptr(), b.data_ptr(), loss.data_ptr(), a.numel(), _stream()), 'mse_loss')
    return loss


def mse_loss_grad(a, b, grad_scale, da):
    _check(lib().b200_mse_loss_grad(a.data_ptr(), b.data_ptr(), _ptr(grad_scale), da.data_ptr(), a.numel(), _stream()),
           'mse_loss_grad')
    return da


def time_embed_bwd(t, freqs, dim, E, cos_first, w1, b1, w2, emb, d_semb, y, pe, hid, demb, dpre, db1, db2, dclass):
    _check(lib().b200_time_embed_bwd(t.data_ptr(), t.shape[0], freqs.data_ptr(), dim, E, int(cos_first), w1.data_ptr(),
                                     b1.data_ptr(), w2.data_ptr(), emb.data_ptr(), d_semb.data_ptr(), _ptr(y),
                                     pe.data_ptr(), hid.data_ptr(), demb.data_ptr(), dpre.data_ptr(), db1.data_ptr(),
                                     db2.data_ptr(), _ptr(dclass), _stream()), 'time_embed_bwd')


def ode_step(model_out, x, coefs, sigma_t, sigma_prev, *, objective='pred_eps', clip=True, d1=None, x1=None, sample=None,
             pred_x0=None, deriv=None):
    """coefs = (sqrt(1/ac), sqrt(1/ac - 1), sqrt(ac), sqrt(1 - ac)) of the evaluation timestep, as Python floats."""
    _need_cuda(model_out, x)
    d = OdeDesc()
    d.model_out, d.x, d.d1, d.x1 = model_out.data_ptr(), x.data_ptr(), _ptr(d1), _ptr(x1)
    d.B, d.C, d.Cm, d.HW = x.shape[0], x.shape[1], model_out.shape[1], x[0, 0].numel()
    d.objective, d.clip, d.second_order = OBJ[objective], int(clip), int(d1 is not None)
    d.sqrt_recip_ac, d.sqrt_recipm1_ac, d.sqrt_ac, d.sqrt_1m_ac = [float(c) for c in coefs]
    d.sigma_t, d.sigma_prev = float(sigma_t), float(sigma_prev)
    d.sample, d.pred_x0, d.deriv = _ptr(sample), _ptr(pred_x0), _ptr(deriv)
    _launch('ode_step', lambda: _check(lib().b200_ode_step(ctypes.byref(d), _stream()), 'ode_step'),
            nbytes=4.0 * x.numel() * (2 + sum(v is not None for v in (d1, x1, sample, pred_x0, deriv))))


def to_uint8_hwc(x, out=None):
    """fp32 NCHW samples in [-1, 1] -> uint8 NHWC pixels with torchvision.save_image's rounding (the output stage)."""
    _need_cuda(x)
    B, C, H, W = x.shape
    if out is None:
        out = torch.empty((B, H, W, C), dtype=torch.uint8, device=x.device)
    _check(lib().b200_to_uint8_hwc(x.contiguous().data_ptr(), out.data_ptr(), B, C, H * W, _stream()), 'to_uint8_hwc')
    return out


def pack_entry_bytes(src, dst, Co, Ci, taps, mode, row0=0, col0=0, ld=0, src2=None, tap_ld=0):
    """One b200_pack_entry as bytes (3 pointers, 8 ints).  mode 3 = FP32-mode split operand [hi | hi | lo] per tap;
    mode 4 = the same with a tap stride of tap_ld columns (first convolution on 64-channel split pixels)."""
    import struct
    if mode >= 3 and taps > 9:
        raise RuntimeError('pack modes 3 / 4 (split operands) support at most 9 taps')
    if mode == 4 and tap_ld < 3 * Ci:
        raise RuntimeError('pack mode 4 needs tap_ld >= 3 * Ci')
    return struct.pack('<3Q8i', src.data_ptr(), 0 if src2 is None else src2.data_ptr(), dst.data_ptr(), Co, Ci, taps, mode,
                       row0, col0, ld, tap_ld)


def pack_weights(table_dev, n_entries):
    _check(lib().b200_pack_weights(table_dev.data_ptr(), n_entries, _stream()), 'pack_weights')
