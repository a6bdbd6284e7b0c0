// Fused optimizer tail of the training step (scripts/train_ddpm.py:186-188 +
// models/ema.py:44-52 of the reference): global gradient-norm clip + Adam / AdamW + EMA of the weights as two
// multi-tensor kernels over a device-resident chunk table, instead of clip_grad_norm_ (2 passes), Adam.step
// (foreach, ~6 passes) and the per-tensor EMA loop.  Nothing synchronises with the host: the clip coefficient is
// computed on the device from the squared norm accumulated by the first kernel.
#include "common.cuh"
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

__global__ void __launch_bounds__(256) optim_sumsq_kernel(const b200_optim_chunk* __restrict__ chunks,
                                                          float* __restrict__ gnorm_sq) {
  const b200_optim_chunk c = chunks[blockIdx.x];
  float s = 0.f;
  for (int i = threadIdx.x; i < c.n; i += 256) {
    const float g = c.g[i];
    s = fmaf(g, g, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += ws[w];
    atomicAdd(gnorm_sq, t);
  }
}

struct OptimK {
  float lr, beta1, beta2, eps, wd, bc1, bc2_sqrt, max_norm, ema_decay;
  int adamw, use_ema;
  const b200_optim_dev_state* dev;   // when set, lr / bias corrections / EMA decay come from device memory
};

// One thread: advances the device-resident step counters and derives the per-step scalars, so that a captured CUDA
// graph of the training step needs no host-computed constants (bias corrections, gradual EMA decay, lr).
__global__ void optim_prepare_kernel(b200_optim_dev_state* st, float beta1, float beta2, int use_ema) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  st->step += 1;
  st->bc1 = (float)(1.0 - pow((double)beta1, (double)st->step));
  st->bc2_sqrt = (float)sqrt(1.0 - pow((double)beta2, (double)st->step));
  if (use_ema) {
    st->ema_updates += 1;
    const float n = (float)st->ema_updates;
    st->ema_decay = st->ema_gradual ? fminf(st->ema_decay_max, (1.f + n) / (10.f + n)) : st->ema_decay_max;
  }
}

__global__ void __launch_bounds__(256) optim_adam_kernel(const b200_optim_chunk* __restrict__ chunks,
                                                         const float* __restrict__ gnorm_sq, const OptimK k) {
  const b200_optim_chunk c = chunks[blockIdx.x];
  float lr = k.lr, bc1 = k.bc1, bc2_sqrt = k.bc2_sqrt, ema_decay = k.ema_decay;
  if (k.dev) { lr = k.dev->lr; bc1 = k.dev->bc1; bc2_sqrt = k.dev->bc2_sqrt; ema_decay = k.dev->ema_decay; }
  float coef = 1.f;
  if (k.max_norm > 0.f) {   // torch.nn.utils.clip_grad_norm_: coef = max_norm / (norm + 1e-6), clamped to 1
    coef = fminf(1.f, k.max_norm / (sqrtf(gnorm_sq[0]) + 1e-6f));
  }
  const float step_size = lr / bc1;
  for (int i = threadIdx.x; i < c.n; i += 256) {
    float p = c.p[i];
    float g = c.g[i] * coef;
    if (k.wd != 0.f) {
      if (k.adamw) p *= 1.f - lr * k.wd;
      else g = fmaf(k.wd, p, g);
    }
    const float m = fmaf(k.beta1, c.m[i], (1.f - k.beta1) * g);
    const float v = fmaf(k.beta2, c.v[i], (1.f - k.beta2) * g * g);
    c.m[i] = m;
    c.v[i] = v;
    p -= step_size * m / (sqrtf(v) / bc2_sqrt + k.eps);
    c.p[i] = p;
    if (k.use_ema && c.ema) {
      const float e = c.ema[i];
      c.ema[i] = e - (1.f - ema_decay) * (e - p);
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_optimizer_step(const b200_optim_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d && d->chunks && d->n_chunks >= 1 && d->gnorm_sq, "optimizer_step: null chunk table / workspace");
  B200_REQUIRE((d->step >= 1 || d->dev_state) && d->beta1 >= 0.f && d->beta1 < 1.f && d->beta2 >= 0.f && d->beta2 < 1.f,
               "optimizer_step: bad step / betas");
  const b200_optim_chunk* chunks = reinterpret_cast<const b200_optim_chunk*>(d->chunks);
  B200_CHECK(cudaMemsetAsync(d->gnorm_sq, 0, sizeof(float), stream));
  if (d->max_grad_norm > 0.f || d->want_norm) {
    optim_sumsq_kernel<<<d->n_chunks, 256, 0, stream>>>(chunks, d->gnorm_sq);
    ++g_launch_count;
    B200_CHECK(cudaGetLastError());
  }
  if (d->dev_state) {
    optim_prepare_kernel<<<1, 32, 0, stream>>>(reinterpret_cast<b200_optim_dev_state*>(d->dev_state), d->beta1, d->beta2,
                                               d->ema_decay >= 0.f ? 1 : 0);
    ++g_launch_count;
    B200_CHECK(cudaGetLastError());
  }
  OptimK k;
  k.dev = reinterpret_cast<const b200_optim_dev_state*>(d->dev_state);
  k.lr = d->lr; k.beta1 = d->beta1; k.beta2 = d->beta2; k.eps = d->eps; k.wd = d->weight_decay;
  const int step = d->step >= 1 ? d->step : 1;
  k.bc1 = (float)(1.0 - pow((double)d->beta1, (double)step));
  k.bc2_sqrt = (float)sqrt(1.0 - pow((double)d->beta2, (double)step));
  k.max_norm = d->max_grad_norm; k.adamw = d->adamw;
  k.use_ema = d->ema_decay >= 0.f ? 1 : 0; k.ema_decay = d->ema_decay;
  optim_adam_kernel<<<d->n_chunks, 256, 0, stream>>>(chunks, d->gnorm_sq, k);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "optim_adam_kernel launch");
}

// ================================================================================================
// Multi-tensor weight pack: after an optimizer step (or load_state_dict) every fp32 OIHW parameter is re-laid-out to
// the bf16 GEMM operand(s) the kernels consume, for ALL layers in one launch over a device-resident table:
//   mode 0 (forward operand)      dst[row0 + co][col0 + tap*Ci + ci]          = src[co][ci][tap]
//   mode 1 (data-gradient operand) dst[row0 + ci][col0 + (taps-1-tap)*Co + co] = src[co][ci][tap]   (flipped taps)
//   mode 2 (fp32 bias sum)         dstf[i] = src[i] + src2[i]                  (conv bias + fused-shortcut bias)
//   mode 3 (FP32-mode operand)     dst[row0 + co][col0 + tap*3*Ci + {0, Ci, 2Ci} + ci] = {hi, hi, lo}(src[co][ci][tap])
//   mode 4 (first-conv operand)    as mode 3 with a tap stride of tap_ld columns (b200_first_split's 64-channel pixels;
//                                  the columns in between are never written: the caller zero-fills dst once)
// Replaces ~1000 tiny torch permute / cast / cat kernels per training step.
// ================================================================================================
namespace b200 {

// Tiled form (taps <= 9): a CTA walks 32 (co) x 32 (ci) x taps tiles of its entry through shared memory.  Loads are
// whole 32*taps-float row segments of the OIHW source (coalesced), stores are 32 consecutive bf16 of a packed row; the
// thread mapping is (32 x 8), so no per-element division is needed.  The first version gathered single floats at a
// stride of taps (mode 0) or Ci*taps (mode 1) elements with two integer divisions each: 448 us per training step for
// 354 MB of traffic.
constexpr int PW_T = 32;
constexpr int PW_MAX_TAPS = 9;
constexpr int PW_LD = PW_T * PW_MAX_TAPS + 1;   // odd row stride: conflict-free column reads in mode 1

__global__ void __launch_bounds__(256) pack_weights_kernel(const b200_pack_entry* __restrict__ table, int ctas_per_entry) {
  __shared__ float tile[PW_T * PW_LD];
  const b200_pack_entry e = table[blockIdx.x / ctas_per_entry];
  const int part = blockIdx.x % ctas_per_entry;
  const long long total = (long long)e.Co * e.Ci * e.taps;
  if (e.mode == 2) {
    const long long per = (total + ctas_per_entry - 1) / ctas_per_entry;
    const long long j0 = part * per, j1 = min(total, j0 + per);
    float* dst = reinterpret_cast<float*>(e.dst);
    for (long long j = j0 + threadIdx.x; j < j1; j += 256) dst[j] = e.src[j] + (e.src2 ? e.src2[j] : 0.f);
    return;
  }
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(e.dst);
  if (e.taps > PW_MAX_TAPS && e.mode >= 3) return;   // rejected on the host (b200_pack_weights cannot see device tables)
  if (e.taps > PW_MAX_TAPS) {   // generic gather (no layer of the supported families takes it)
    const long long per = (total + ctas_per_entry - 1) / ctas_per_entry;
    const long long j0 = part * per, j1 = min(total, j0 + per);
    const int tc = e.taps * (e.mode == 0 ? e.Ci : e.Co);    // packed row length of this entry
    for (long long j = j0 + threadIdx.x; j < j1; j += 256) {
      const int r = (int)(j / tc);
      const int rem = (int)(j - (long long)r * tc);
      int co, ci, tap;
      if (e.mode == 0) {
        co = r; tap = rem / e.Ci; ci = rem - tap * e.Ci;
      } else {
        ci = r;
        const int tp = rem / e.Co;
        co = rem - tp * e.Co;
        tap = e.taps - 1 - tp;
      }
      const float v = __ldg(e.src + ((long long)co * e.Ci + ci) * e.taps + tap);
      dst[(long long)(e.row0 + r) * e.ld + e.col0 + rem] = __float2bfloat16_rn(v);
    }
    return;
  }
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int taps = e.taps;
  const int tiles_ci = (e.Ci + PW_T - 1) / PW_T, tiles_co = (e.Co + PW_T - 1) / PW_T;
  for (int t = part; t < tiles_co * tiles_ci; t += ctas_per_entry) {
    const int co0 = (t / tiles_ci) * PW_T, ci0 = (t % tiles_ci) * PW_T;
    const int nco = min(PW_T, e.Co - co0), nci = min(PW_T, e.Ci - ci0);
    const int seg = nci * taps;   // contiguous floats of one co row inside this tile
    for (int r = ty; r < nco; r += 8) {
      const float* srow = e.src + ((long long)(co0 + r) * e.Ci + ci0) * taps;
      for (int c = tx; c < seg; c += 32) tile[r * PW_LD + c] = __ldg(srow + c);
    }
    __syncthreads();
    if (e.mode == 0) {
      // dst[row0 + co][col0 + tap*Ci + ci]
      for (int r = ty; r < nco; r += 8) {
        __nv_bfloat16* drow = dst + (long long)(e.row0 + co0 + r) * e.ld + e.col0 + ci0;
        if (tx < nci)
          for (int tap = 0; tap < taps; ++tap)
            drow[(long long)tap * e.Ci + tx] = __float2bfloat16_rn(tile[r * PW_LD + tx * taps + tap]);
      }
    } else if (e.mode >= 3) {
      // FP32 mode (precise.cu): weight side of the 3-term split product, per tap [w_hi | w_hi | w_lo]:
      // dst[row0 + co][col0 + tap*3*Ci + {0, Ci, 2*Ci} + ci]
      for (int r = ty; r < nco; r += 8) {
        __nv_bfloat16* drow = dst + (long long)(e.row0 + co0 + r) * e.ld + e.col0 + ci0;
        if (tx < nci)
          for (int tap = 0; tap < taps; ++tap) {
            const float w = tile[r * PW_LD + tx * taps + tap];
            const __nv_bfloat16 hi = __float2bfloat16_rn(w);
            const __nv_bfloat16 lo = __float2bfloat16_rn(w - __bfloat162float(hi));
            __nv_bfloat16* d3 = drow + (long long)tap * (e.mode == 4 ? e.tap_ld : 3 * e.Ci) + tx;
            d3[0] = hi; d3[e.Ci] = hi; d3[2 * e.Ci] = lo;
          }
      }
    } else {
      // dst[row0 + ci][col0 + (taps-1-tap)*Co + co]
      for (int q = ty; q < nci; q += 8) {
        __nv_bfloat16* drow = dst + (long long)(e.row0 + ci0 + q) * e.ld + e.col0 + co0;
        if (tx < nco)
          for (int tp = 0; tp < taps; ++tp)
            drow[(long long)tp * e.Co + tx] = __float2bfloat16_rn(tile[tx * PW_LD + q * taps + (taps - 1 - tp)]);
      }
    }
    __syncthreads();
  }
}

}  // namespace b200

extern "C" int b200_pack_weights(const void* table_dev, int n_entries, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(table_dev && n_entries >= 1, "pack_weights: empty table");
  const int cpe = 8;
  pack_weights_kernel<<<n_entries * cpe, 256, 0, stream>>>(reinterpret_cast<const b200_pack_entry*>(table_dev), cpe);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "pack_weights_kernel launch");
}
