// FP32 mode ("bf16x3"): support kernels for running the UNet forward at fp32-level accuracy on the bf16 tensor cores.
//
// The reference runs this path in fp32 (models/unet.py:121-152, fp32 bmm / softmax in models/modules.py:92-97);
// BASELINE.json's north_star asks for a mode whose per-step eps prediction agrees with it to 1e-4 relative L2.
// tcgen05 has no fp32 operand type, so every operand x is split into two bf16 numbers
//     x = hi + lo + r,   hi = bf16(x),  lo = bf16(x - hi),  |r| <= 2^-18 |x|
// and a product a * w is evaluated as the three tensor-core terms a_hi w_hi + a_lo w_hi + a_hi w_lo (the dropped
// a_lo w_lo term is <= 2^-18 relative), accumulated in fp32 in TMEM.  The three terms are laid out along the GEMM K
// dimension, so the UNCHANGED implicit-GEMM kernels (conv_gemm*.cu, gemm_tc.cu) execute them as a convolution with 3x
// the input channels:
//     "activation" side (pattern 0):  [ hi | lo | hi ]  per group of C channels
//     "weight" side     (pattern 1):  [ hi | hi | lo ]
// This file holds the producers of those layouts: a generic fp32 -> split cast (+ optional exact SiLU, parity-plane and
// plane layouts), GroupNorm(+AdaGN)(+SiLU)(+2x resample) with fp64 statistics and an exact SiLU writing the split
// operand, and a row softmax writing split probabilities.  The weight-side split lives in b200_pack_weights (mode 3).
// Throughput is secondary here (3x the MMAs by construction); accuracy is the product: no approximate intrinsics.
#include "common.cuh"
#include "../../include/b200diff.h"
#include <math.h>

namespace b200 {
extern long long g_launch_count;

__device__ __forceinline__ void split_bf16(float v, __nv_bfloat16& hi, __nv_bfloat16& lo) {
  hi = __float2bfloat16_rn(v);
  lo = __float2bfloat16_rn(v - __bfloat162float(hi));
}

__device__ __forceinline__ float silu_exact(float x) { return x / (1.0f + expf(-x)); }

// writes the three terms of one channel quad; `plane_stride` = distance (elements) between the three copies
__device__ __forceinline__ void store_split4(__nv_bfloat16* dst, long long plane_stride, int pattern, const float (&v)[4]) {
  __nv_bfloat16 hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) split_bf16(v[i], hi[i], lo[i]);
  uint2 uh, ul;
  uh.x = (uint32_t)__bfloat16_as_ushort(hi[0]) | ((uint32_t)__bfloat16_as_ushort(hi[1]) << 16);
  uh.y = (uint32_t)__bfloat16_as_ushort(hi[2]) | ((uint32_t)__bfloat16_as_ushort(hi[3]) << 16);
  ul.x = (uint32_t)__bfloat16_as_ushort(lo[0]) | ((uint32_t)__bfloat16_as_ushort(lo[1]) << 16);
  ul.y = (uint32_t)__bfloat16_as_ushort(lo[2]) | ((uint32_t)__bfloat16_as_ushort(lo[3]) << 16);
  *reinterpret_cast<uint2*>(dst) = uh;
  *reinterpret_cast<uint2*>(dst + plane_stride) = pattern == 0 ? ul : uh;
  *reinterpret_cast<uint2*>(dst + 2 * plane_stride) = pattern == 0 ? uh : ul;
}

// ------------------------------------------------------------------------------------------------------------------
// generic split cast
// ------------------------------------------------------------------------------------------------------------------
struct SplitK {
  const float* in; long long rows; int in_ld, in_col0, C, group, pattern, act;
  __nv_bfloat16* out; int out_ld, out_col0;
  int planes_rows;          // > 0: plane layout out[b][3][planes_rows][C], b = row / planes_rows
  int par_H, par_W;         // > 0: rows are NHWC pixels of [B][par_H][par_W]; output rows follow the 4 parity planes
};

__global__ void __launch_bounds__(256) split_cast_kernel(const SplitK k) {
  const int quads = k.C >> 2;
  const long long units = k.rows * quads;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < units; u += (long long)gridDim.x * blockDim.x) {
    const long long r = u / quads;
    const int c = (int)(u - r * quads) << 2;
    const float4 x = *reinterpret_cast<const float4*>(k.in + r * k.in_ld + k.in_col0 + c);
    float v[4] = {x.x, x.y, x.z, x.w};
    if (k.act == 1) {
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = silu_exact(v[i]);
    }
    if (k.planes_rows > 0) {
      const long long b = r / k.planes_rows, t = r - b * k.planes_rows;
      store_split4(k.out + ((b * 3) * k.planes_rows + t) * k.C + c, (long long)k.planes_rows * k.C, k.pattern, v);
      continue;
    }
    long long ro = r;
    if (k.par_H > 0) {
      const int hw = k.par_H * k.par_W;
      const long long b = r / hw;
      const int pix = (int)(r - b * hw);
      const int h = pix / k.par_W, w = pix - h * k.par_W;
      ro = ((b * 4 + ((h & 1) << 1) + (w & 1)) * (k.par_H >> 1) + (h >> 1)) * (k.par_W >> 1) + (w >> 1);
    }
    const int g = c / k.group, j = c - g * k.group;
    store_split4(k.out + ro * k.out_ld + k.out_col0 + g * 3 * k.group + j, k.group, k.pattern, v);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// GroupNorm (+ AdaGN scale/shift) (+ SiLU) (+ 2x2 average pool / nearest 2x) -> split operand [B][HWo][3C] (pattern 0),
// optional raw copy of the un-normalised concatenation as a split operand (the 1x1 / 3x3 shortcut conv's input).
// Statistics: the producers' int64 fixed-point sums (common.cuh) evaluated in fp64 (no E[x^2] - mean^2 cancellation
// at fp32 level).  Same semantics as b200_groupnorm_apply_fwd (groupnorm.cu) otherwise.
// ------------------------------------------------------------------------------------------------------------------
struct GnSplitK {
  const float* x0; int C0; const long long* st0;
  const float* x1; int C1; const long long* st1;
  int HW, W, cpg;
  const float* gamma; const float* beta; float eps;
  const float* scale; const float* shift; int ss_ld;
  int apply_silu, resample;
  __nv_bfloat16* out; __nv_bfloat16* raw;
};

__global__ void __launch_bounds__(256) gn_apply_split_kernel(const GnSplitK p) {
  extern __shared__ float gsm[];
  const int C = p.C0 + p.C1;
  float* cA = gsm; float* cB = gsm + C;
  const int n = blockIdx.y;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / p.cpg) * p.cpg;
    long long s = 0, q = 0;
    for (int i = 0; i < p.cpg; ++i) {
      const int cc = g0 + i;
      const long long* st = (cc < p.C0) ? p.st0 + ((size_t)n * p.C0 + cc) * 2 : p.st1 + ((size_t)n * p.C1 + (cc - p.C0)) * 2;
      s += st[0]; q += st[1];
    }
    const double cnt = (double)p.HW * p.cpg;
    const double mean = (double)s / (double)kStatQ1 / cnt;
    double var = (double)q / (double)kStatQ2 / cnt - mean * mean;
    if (var < 0.0) var = 0.0;
    const double rstd = 1.0 / sqrt(var + (double)p.eps);
    double ga = p.gamma ? (double)p.gamma[c] : 1.0;
    double be = p.beta ? (double)p.beta[c] : 0.0;
    if (p.scale) {
      const double sc = 1.0 + (double)p.scale[(size_t)n * p.ss_ld + c];
      ga *= sc;
      be = be * sc + (double)p.shift[(size_t)n * p.ss_ld + c];
    }
    cA[c] = (float)(rstd * ga);
    cB[c] = (float)(be - mean * rstd * ga);
  }
  __syncthreads();
  const int quads = C >> 2;
  const int Wo = p.resample == 1 ? p.W >> 1 : p.resample == 2 ? p.W << 1 : p.W;
  const int HWo = p.resample == 1 ? p.HW >> 2 : p.resample == 2 ? p.HW << 2 : p.HW;
  auto load4 = [&](int pix, int c, float (&v)[4]) {
    const float* src = (c < p.C0) ? p.x0 + ((size_t)n * p.HW + pix) * p.C0 + c : p.x1 + ((size_t)n * p.HW + pix) * p.C1 + (c - p.C0);
    const float4 x = *reinterpret_cast<const float4*>(src);
    v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
  };
  auto act4 = [&](int c, float (&v)[4]) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float y = fmaf(v[i], cA[c + i], cB[c + i]);
      v[i] = p.apply_silu ? silu_exact(y) : y;
    }
  };
  // work items = (output pixel, channel quad); for nearest-2x the item is an INPUT pixel that writes 4 outputs
  const int items_px = p.resample == 2 ? p.HW : HWo;
  const long long items = (long long)items_px * quads;
  for (long long u = (long long)blockIdx.x * blockDim.x + threadIdx.x; u < items; u += (long long)gridDim.x * blockDim.x) {
    const int pix = (int)(u / quads);
    const int c = (int)(u - (long long)pix * quads) << 2;
    float v[4];
    if (p.resample == 0) {
      load4(pix, c, v);
      if (p.raw) store_split4(p.raw + ((size_t)n * p.HW + pix) * 3 * C + c, C, 0, v);
      act4(c, v);
      store_split4(p.out + ((size_t)n * HWo + pix) * 3 * C + c, C, 0, v);
    } else if (p.resample == 1) {
      const int oy = pix / Wo, ox = pix - oy * Wo;
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          load4((2 * oy + dy) * p.W + 2 * ox + dx, c, v);
          act4(c, v);
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] += v[i];
        }
#pragma unroll
      for (int i = 0; i < 4; ++i) acc[i] *= 0.25f;
      store_split4(p.out + ((size_t)n * HWo + pix) * 3 * C + c, C, 0, acc);
    } else {
      const int iy = pix / p.W, ix = pix - iy * p.W;
      load4(pix, c, v);
      act4(c, v);
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx)
          store_split4(p.out + ((size_t)n * HWo + (size_t)(2 * iy + dy) * Wo + 2 * ix + dx) * 3 * C + c, C, 0, v);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// row softmax of fp32 scores -> split probabilities [rows][3T] = [p_hi | p_lo | p_hi] (the A operand of P V)
// (models/modules.py:95: softmax over keys of q k^T * d^-1/2; exact expf, one warp per row)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) softmax_rows_split_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P,
                                                                 long long rows, int T, float scale) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* s = S + r * T;
  float m = -INFINITY;
  for (int j = lane; j < T; j += 32) m = fmaxf(m, s[j] * scale);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) sum += expf(s[j] * scale - m);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __nv_bfloat16* prow = P + r * 3 * T;
  for (int j = lane; j < T; j += 32) {
    const float pv = expf(s[j] * scale - m) / sum;
    __nv_bfloat16 hi, lo;
    split_bf16(pv, hi, lo);
    prow[j] = hi; prow[T + j] = lo; prow[2 * T + j] = hi;
  }
}

static inline int ew_grid2(long long units, int block) {
  long long g = (units + block - 1) / block;
  const long long cap = 148 * 8;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_split_cast(const b200_split_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d && d->in && d->out && d->rows >= 1, "split_cast: null pointer / empty input");
  B200_REQUIRE(d->C >= 4 && d->C % 4 == 0 && d->in_ld % 4 == 0 && d->in_col0 % 4 == 0 && ((uintptr_t)d->in & 15) == 0,
               "split_cast: C=%d, in_ld=%d, in_col0=%d must be multiples of 4 (16-byte loads)", d->C, d->in_ld, d->in_col0);
  B200_REQUIRE(d->group >= 4 && d->group % 4 == 0 && d->C % d->group == 0, "split_cast: group=%d must divide C=%d", d->group, d->C);
  B200_REQUIRE(d->pattern == 0 || d->pattern == 1, "split_cast: pattern must be 0 ([hi|lo|hi]) or 1 ([hi|hi|lo])");
  B200_REQUIRE(((uintptr_t)d->out & 7) == 0, "split_cast: out alignment");
  SplitK k;
  k.in = d->in; k.rows = d->rows; k.in_ld = d->in_ld; k.in_col0 = d->in_col0; k.C = d->C; k.group = d->group;
  k.pattern = d->pattern; k.act = d->act;
  k.out = reinterpret_cast<__nv_bfloat16*>(d->out); k.out_ld = d->out_ld; k.out_col0 = d->out_col0;
  k.planes_rows = d->planes_rows; k.par_H = d->parity_H; k.par_W = d->parity_W;
  if (d->planes_rows > 0) {
    B200_REQUIRE(d->rows % d->planes_rows == 0 && d->parity_H == 0, "split_cast: plane layout needs rows %% planes_rows == 0");
  } else {
    B200_REQUIRE(d->out_ld >= 3 * d->C + d->out_col0 && d->out_ld % 4 == 0 && d->out_col0 % 4 == 0,
                 "split_cast: out_ld=%d too small for 3*C=%d columns", d->out_ld, 3 * d->C);
    if (d->parity_H > 0)
      B200_REQUIRE(d->parity_H % 2 == 0 && d->parity_W % 2 == 0 && d->rows % ((long long)d->parity_H * d->parity_W) == 0,
                   "split_cast: parity split needs even H, W and whole images");
  }
  split_cast_kernel<<<ew_grid2(d->rows * (d->C / 4), 256), 256, 0, stream>>>(k);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "split_cast_kernel launch");
}

extern "C" int b200_groupnorm_apply_split_fwd(const float* x0, int C0, const long long* stats0, const float* x1, int C1,
                                              const long long* stats1, int B, int HW, int W, int groups,
                                              const float* gamma, const float* beta, float eps, const float* scale,
                                              const float* shift, int ss_ld, int apply_silu, int resample, void* out_split,
                                              void* raw_out_split, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x0 && stats0 && out_split, "groupnorm_apply_split: null x0/stats0/out");
  if (x1 == nullptr) C1 = 0;
  const int C = C0 + C1;
  B200_REQUIRE(x1 == nullptr || stats1 != nullptr, "groupnorm_apply_split: second source needs its statistics");
  B200_REQUIRE(groups >= 1 && C % groups == 0, "groupnorm_apply_split: C=%d is not a multiple of groups=%d", C, groups);
  B200_REQUIRE(C0 % 4 == 0 && C1 % 4 == 0 && C <= 4096, "groupnorm_apply_split: channel counts must be multiples of 4 (<= 4096)");
  B200_REQUIRE(resample >= 0 && resample <= 2 && W >= 1 && HW % W == 0, "groupnorm_apply_split: bad resample / W");
  B200_REQUIRE(resample != 1 || (W % 2 == 0 && (HW / W) % 2 == 0), "groupnorm_apply_split: average pooling needs even H, W");
  B200_REQUIRE(!(raw_out_split && resample), "groupnorm_apply_split: the raw copy is not available with resampling");
  B200_REQUIRE((scale == nullptr) == (shift == nullptr), "groupnorm_apply_split: scale and shift come together");
  GnSplitK p;
  p.x0 = x0; p.C0 = C0; p.st0 = stats0; p.x1 = x1; p.C1 = C1; p.st1 = stats1;
  p.HW = HW; p.W = W; p.cpg = C / groups;
  p.gamma = gamma; p.beta = beta; p.eps = eps; p.scale = scale; p.shift = shift; p.ss_ld = ss_ld;
  p.apply_silu = apply_silu; p.resample = resample;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_split); p.raw = reinterpret_cast<__nv_bfloat16*>(raw_out_split);
  const long long items = (long long)(resample == 1 ? HW / 4 : HW) * (C / 4);
  int gx = (int)((items + 256 * 4 - 1) / (256 * 4));
  if (gx < 1) gx = 1;
  if (gx > 64) gx = 64;
  gn_apply_split_kernel<<<dim3(gx, B), 256, 2 * C * sizeof(float), stream>>>(p);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "gn_apply_split_kernel launch");
}

extern "C" int b200_softmax_rows_split(const float* S, void* P_split, long long rows, int T, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(S && P_split && rows >= 1 && T >= 1, "softmax_rows_split: bad arguments");
  softmax_rows_split_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(S, reinterpret_cast<__nv_bfloat16*>(P_split), rows, T, scale);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "softmax_rows_split_kernel launch");
}
