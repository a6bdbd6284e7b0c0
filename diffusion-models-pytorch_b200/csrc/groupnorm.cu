// K3: GroupNorm (+AdaGN scale/shift) (+SiLU) (+2x avg-pool / nearest-2x) over fp32 NHWC input(s), bf16 NHWC out.
// Replaces nn.GroupNorm + nn.SiLU + torch.cat of the reference (models/unet.py:14-15,23-24,116-117,145;
// models/modules.py:82,105-123; models/unet_categorial_adagn.py:52-57).
//
// One CTA owns (image n, a chunk of whole groups): the [HW x CC] fp32 slab is read from HBM exactly once into
// shared memory, statistics are an exact two-pass (mean, then centred variance) over the slab, and the
// normalised, activated values are rounded to bf16 exactly once on the way out.  HBM traffic = 4 B read +
// 2 B written per element (+2 B when the raw bf16 copy for the 1x1 shortcut conv is requested).
#include "common.cuh"
#include <stdlib.h>
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

struct GnParams {
  const float* x0; int C0;
  const float* x1; int C1;
  int HW, W, groups, cpg, gpc;  // gpc = groups per CTA chunk
  int chunks;
  const float* gamma; const float* beta; float eps;
  const float* scale; const float* shift; int ss_ld;
  int apply_silu, resample;
  __nv_bfloat16* out; __nv_bfloat16* raw;
};

template <int V> struct VecT;
template <> struct VecT<4> { using type = float4; };
template <> struct VecT<2> { using type = float2; };

template <int V>
__device__ __forceinline__ void load_vec(const float* p, float (&v)[V]) {
  if constexpr (V == 4) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(p));
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    const float2 t = __ldg(reinterpret_cast<const float2*>(p));
    v[0] = t.x; v[1] = t.y;
  }
}
template <int V>
__device__ __forceinline__ void lds_vec(const float* p, float (&v)[V]) {
  if constexpr (V == 4) {
    const float4 t = *reinterpret_cast<const float4*>(p);
    v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
  } else {
    const float2 t = *reinterpret_cast<const float2*>(p);
    v[0] = t.x; v[1] = t.y;
  }
}
template <int V>
__device__ __forceinline__ void sts_vec(float* p, const float (&v)[V]) {
  if constexpr (V == 4) *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
  else *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}
template <int V>
__device__ __forceinline__ void store_bf16_vec(__nv_bfloat16* p, const float (&v)[V]) {
  if constexpr (V == 4) {
    uint2 u;
    u.x = pack_bf16x2(v[0], v[1]);
    u.y = pack_bf16x2(v[2], v[3]);
    *reinterpret_cast<uint2*>(p) = u;
  } else {
    *reinterpret_cast<uint32_t*>(p) = pack_bf16x2(v[0], v[1]);
  }
}

template <int V>
__global__ void __launch_bounds__(256) groupnorm_kernel(const GnParams p) {
  extern __shared__ float gsm[];
  const int n = blockIdx.x / p.chunks;
  const int chunk = blockIdx.x - n * p.chunks;
  const int g0 = chunk * p.gpc;
  const int ng = min(p.gpc, p.groups - g0);
  const int c0 = g0 * p.cpg;
  const int CC = ng * p.cpg;          // channels handled by this CTA
  const int pitch = p.gpc * p.cpg + 4;  // slab row pitch in floats (16 B pad -> conflict-free float4 columns)
  const int nv = CC / V;
  const int C = p.C0 + p.C1;
  float* slab = gsm;
  float* coefA = slab + (size_t)p.HW * pitch;  // [CC] multiplicative coefficient
  float* coefB = coefA + p.gpc * p.cpg;        // [CC] additive coefficient
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5;

  // ---- pass 0: HBM -> smem (single read of the input), optional raw bf16 copy ----
  const int total = p.HW * nv;
  for (int idx = tid; idx < total; idx += 256) {
    const int px = idx / nv;
    const int j = idx - px * nv;
    const int c = c0 + j * V;
    const float* src = (c < p.C0) ? p.x0 + ((size_t)n * p.HW + px) * p.C0 + c
                                  : p.x1 + ((size_t)n * p.HW + px) * p.C1 + (c - p.C0);
    float v[V];
    load_vec<V>(src, v);
    sts_vec<V>(slab + px * pitch + j * V, v);
    if (p.raw) store_bf16_vec<V>(p.raw + ((size_t)n * p.HW + px) * C + c, v);
  }
  __syncthreads();

  // ---- pass 1+2: per-group mean and centred variance from smem; one warp per group ----
  const int vpg = p.cpg / V;
  const float inv_cnt = 1.0f / (float)(p.HW * p.cpg);
  for (int g = warp; g < ng; g += 8) {
    const float* base = slab + g * p.cpg;
    float s = 0.f;
    for (int px = lane; px < p.HW; px += 32)
      for (int i = 0; i < vpg; ++i) {
        float t[V];
        lds_vec<V>(base + px * pitch + i * V, t);
#pragma unroll
        for (int e = 0; e < V; ++e) s += t[e];
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s * inv_cnt;
    float q = 0.f;
    for (int px = lane; px < p.HW; px += 32)
      for (int i = 0; i < vpg; ++i) {
        float t[V];
        lds_vec<V>(base + px * pitch + i * V, t);
#pragma unroll
        for (int e = 0; e < V; ++e) {
          const float d = t[e] - mean;
          q += d * d;
        }
      }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float rstd = rsqrtf(q * inv_cnt + p.eps);
    // fold mean/rstd, affine and the AdaGN scale/shift into y = x * A + B
    for (int cc = lane; cc < p.cpg; cc += 32) {
      const int c = c0 + g * p.cpg + cc;
      float ga = p.gamma ? __ldg(p.gamma + c) : 1.f;
      float be = p.beta ? __ldg(p.beta + c) : 0.f;
      if (p.scale) {
        const float sc = 1.f + __ldg(p.scale + (size_t)n * p.ss_ld + c);
        ga *= sc;
        be = be * sc + __ldg(p.shift + (size_t)n * p.ss_ld + c);
      }
      coefA[g * p.cpg + cc] = rstd * ga;
      coefB[g * p.cpg + cc] = be - mean * rstd * ga;
    }
  }
  __syncthreads();

  // ---- pass 3: normalise (+SiLU) (+resample) and write bf16 ----
  if (p.resample == 0) {
    for (int idx = tid; idx < total; idx += 256) {
      const int px = idx / nv;
      const int j = idx - px * nv;
      float v[V], a[V], b[V];
      lds_vec<V>(slab + px * pitch + j * V, v);
      lds_vec<V>(coefA + j * V, a);
      lds_vec<V>(coefB + j * V, b);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float y = v[i] * a[i] + b[i];
        v[i] = p.apply_silu ? silu_f(y) : y;
      }
      store_bf16_vec<V>(p.out + ((size_t)n * p.HW + px) * C + c0 + j * V, v);
    }
  } else if (p.resample == 1) {  // 2x2 average pool of the activated values
    const int H = p.HW / p.W, Wo = p.W / 2, Ho = H / 2;
    const int total_o = Ho * Wo * nv;
    for (int idx = tid; idx < total_o; idx += 256) {
      const int po = idx / nv;
      const int j = idx - po * nv;
      const int oy = po / Wo, ox = po - oy * Wo;
      float v[V], a[V], b[V];
      lds_vec<V>(coefA + j * V, a);
      lds_vec<V>(coefB + j * V, b);
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = 0.f;
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int px = (2 * oy + dy) * p.W + 2 * ox + dx;
          float t[V];
          lds_vec<V>(slab + px * pitch + j * V, t);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            const float y = t[i] * a[i] + b[i];
            v[i] += p.apply_silu ? silu_f(y) : y;
          }
        }
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] *= 0.25f;
      store_bf16_vec<V>(p.out + ((size_t)n * (Ho * Wo) + po) * C + c0 + j * V, v);
    }
  } else {  // nearest 2x
    const int Wo = p.W * 2;
    for (int idx = tid; idx < total; idx += 256) {
      const int px = idx / nv;
      const int j = idx - px * nv;
      const int iy = px / p.W, ix = px - iy * p.W;
      float v[V], a[V], b[V];
      lds_vec<V>(slab + px * pitch + j * V, v);
      lds_vec<V>(coefA + j * V, a);
      lds_vec<V>(coefB + j * V, b);
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const float y = v[i] * a[i] + b[i];
        v[i] = p.apply_silu ? silu_f(y) : y;
      }
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const size_t po = (size_t)(2 * iy + dy) * Wo + 2 * ix + dx;
          store_bf16_vec<V>(p.out + ((size_t)n * (4 * p.HW) + po) * C + c0 + j * V, v);
        }
    }
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_groupnorm_silu_fwd(const float* x0, int C0, const float* x1, int C1, int B, int HW, int W,
                                       int groups, const float* gamma, const float* beta, float eps,
                                       const float* scale, const float* shift, int ss_ld, int apply_silu,
                                       int resample, void* out_bf16, void* raw_out_bf16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x0 && out_bf16, "groupnorm: null x0/out");
  if (!x1) C1 = 0;
  const int C = C0 + C1;
  B200_REQUIRE(groups > 0 && C % groups == 0, "groupnorm: C=%d not divisible by groups=%d", C, groups);
  const int cpg = C / groups;
  B200_REQUIRE(cpg % 2 == 0 && C0 % 2 == 0, "groupnorm: channels per group (%d) and C0 (%d) must be even", cpg, C0);
  B200_REQUIRE(resample >= 0 && resample <= 2, "groupnorm: bad resample mode");
  B200_REQUIRE(W > 0 && HW % W == 0, "groupnorm: HW=%d not a multiple of W=%d", HW, W);
  if (resample == 1) B200_REQUIRE(W % 2 == 0 && (HW / W) % 2 == 0, "groupnorm: avg-pool needs even H, W");
  B200_REQUIRE((scale == nullptr) == (shift == nullptr), "groupnorm: scale and shift must be given together");
  const int V = (cpg % 4 == 0 && C0 % 4 == 0) ? 4 : 2;
  // groups per CTA: largest power of two (<= groups) whose slab stays <= 64 KB; fall back to 1 group up to 200 KB
  int gpc = 1;
  for (int g = groups; g >= 1; g >>= 1) {
    if ((size_t)HW * (g * cpg + 4) * 4 <= 64 * 1024) { gpc = g; break; }
  }
  const size_t smem = ((size_t)HW * (gpc * cpg + 4) + 2 * (size_t)gpc * cpg) * 4;
  B200_REQUIRE(smem <= 200 * 1024, "groupnorm: slab of %zu bytes (HW=%d, %d ch/group) exceeds the single-CTA path",
               smem, HW, cpg);
  GnParams p;
  p.x0 = x0; p.C0 = C0; p.x1 = x1; p.C1 = C1;
  p.HW = HW; p.W = W; p.groups = groups; p.cpg = cpg; p.gpc = gpc;
  p.chunks = (groups + gpc - 1) / gpc;
  p.gamma = gamma; p.beta = beta; p.eps = eps;
  p.scale = scale; p.shift = shift; p.ss_ld = ss_ld;
  p.apply_silu = apply_silu; p.resample = resample;
  p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.raw = reinterpret_cast<__nv_bfloat16*>(raw_out_bf16);
  const int grid = B * p.chunks;
  if (V == 4) {
    static bool attr4 = false;
    if (!attr4) {
      B200_CHECK(cudaFuncSetAttribute(groupnorm_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr4 = true;
    }
    groupnorm_kernel<4><<<grid, 256, smem, stream>>>(p);
  } else {
    static bool attr2 = false;
    if (!attr2) {
      B200_CHECK(cudaFuncSetAttribute(groupnorm_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      attr2 = true;
    }
    groupnorm_kernel<2><<<grid, 256, smem, stream>>>(p);
  }
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "groupnorm_kernel launch");
}

// ================================================================================================
// K3, streaming variant: the per-(image, channel) sum / sum-of-squares of the input were already accumulated by
// the kernel that produced it (conv epilogue, conv_gemm.cu), so GroupNorm+SiLU is ONE coalesced pass:
// read fp32 once, write bf16 once.  A CTA owns (image n, a contiguous range of pixels, all channels); its
// prologue turns the channel statistics into per-channel y = x*A + B coefficients in shared memory.
// ================================================================================================
namespace b200 {

struct GnApplyParams {
  const float* x0; int C0; const long long* st0;   // statistics: [B][C][2] int64 fixed point (common.cuh: stat_load)
  const float* x1; int C1; const long long* st1;
  int HW, W, groups, cpg, pix_per_cta;
  const float* gamma; const float* beta; float eps;
  const float* scale; const float* shift; int ss_ld;
  int apply_silu, resample;
  int cols8;    // 8-channel columns: C0, C1 multiples of 8 and C / 8 <= 256; 2 = coefficients computed per thread
  int reverse;  // walk images / pixel ranges from the end: the producer's most recent writes are still in L2
  int in_bf16;  // source 0 is bf16 (a conv output consumed only by this GroupNorm), single source only
  int c_begin;  // first channel this launch handles: C0 when the first source's part of `out` / `raw` was already written by
                // its producer (fused conv epilogue, b200_conv2d_gn_fwd block-output form; x0 == NULL), else 0
  __nv_bfloat16* out; __nv_bfloat16* raw;
  float drop_p, drop_scale;       // training-mode dropout after the activation (resample == 0 only)
  uint32_t drop_thresh; unsigned long long drop_seed; const unsigned long long* drop_seed_dev;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 ldg4_bf16(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  float4 r;
  r.x = __uint_as_float(u.x << 16); r.y = __uint_as_float(u.x & 0xffff0000u);
  r.z = __uint_as_float(u.y << 16); r.w = __uint_as_float(u.y & 0xffff0000u);
  return r;
}
// SiLU through one MUFU op: x * sigmoid(x) = x * (0.5 * tanh(x / 2) + 0.5); tanh.approx error (2^-11) is below the
// half-ulp of the bf16 the value is rounded to.
__device__ __forceinline__ float silu_tanh(float x) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * x));
  return x * fmaf(0.5f, th, 0.5f);
}

// Fast path of the streaming kernel: a thread owns one 8-channel column (coefficients in registers) and walks down the
// CTA's pixels with kU independent 16-byte (bf16 source) / 2 x 16-byte (fp32 source) loads in flight; every store is
// a full 16-byte bf16x8.  (With 4-channel columns a bf16 source moved only 8 bytes per load and ran latency-bound.)
// kOwnCoef: the thread derives the y = x*a + b coefficients of ITS 8 channels straight from the producer's statistics
// (its GroupNorm groups lie inside / are made of whole 8-channel columns), after its first data loads were issued:
// no shared-memory prologue, no __syncthreads, the statistics' latency hides behind the data loads.
template <bool kInBf16, int kU, bool kOwnCoef>
__device__ __forceinline__ void gn_apply_cols8(const GnApplyParams& p, const float* coefA, const float* coefB, int n,
                                               int px0, int px1, unsigned long long drop_seed) {
  const int C = p.C0 + p.C1;
  const int nv8 = (C - p.c_begin) >> 3;
  const int pstep = 256 / nv8;
  const int tid = threadIdx.x;
  if (tid >= pstep * nv8) return;
  const int j = tid % nv8, prow = tid / nv8;
  const int c = p.c_begin + (j << 3);
  const bool from0 = c < p.C0;
  const int sld = from0 ? p.C0 : p.C1;
  const float* src = from0 ? p.x0 + (size_t)n * p.HW * p.C0 + c : p.x1 + (size_t)n * p.HW * p.C1 + (c - p.C0);
  const __nv_bfloat16* srcb = reinterpret_cast<const __nv_bfloat16*>(p.x0) + (size_t)n * p.HW * p.C0 + c;
  __nv_bfloat16* dst = p.out + (size_t)n * p.HW * C + c;
  __nv_bfloat16* rdst = p.raw ? p.raw + (size_t)n * p.HW * C + c : nullptr;
  uint4 raw0[kU], raw1[kInBf16 ? 1 : kU];   // loads stay packed until they are consumed (register budget: 4 CTAs / SM)
  // mixed sources (p.in_bf16 == 2: first source bf16 -- e.g. an up-sampling conv's output kept in bf16 --, second source an
  // fp32 skip connection): the dtype is a per-thread property of the column's source
  const bool b16 = kInBf16 || (p.in_bf16 == 2 && from0);
  auto issue = [&](int px) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = px + u * pstep;
      if (q < px1) {
        if (kInBf16 || b16) {
          raw0[u] = __ldg(reinterpret_cast<const uint4*>(srcb + (size_t)q * sld));
        } else {
          raw0[u] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)q * sld));
          raw1[kInBf16 ? 0 : u] = __ldg(reinterpret_cast<const uint4*>(src + (size_t)q * sld + 4));
        }
      }
    }
  };
  int px = px0 + prow;
  if (px >= px1) return;
  issue(px);
  float a[8], b[8];
  if (kOwnCoef) {
    const float inv_cnt = 1.0f / (float)(p.HW * p.cpg);
    float mean[8], rstd[8];
    if (p.cpg <= 8) {
      // statistics of the own 8 channels: 16 consecutive floats of [n][c][2]
      const long long* st = from0 ? p.st0 + ((size_t)n * p.C0 + c) * 2 : p.st1 + ((size_t)n * p.C1 + (c - p.C0)) * 2;
      longlong2 sq[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) sq[i] = __ldg(reinterpret_cast<const longlong2*>(st) + i);
      // group sums are exact in the fixed-point domain; one int64 -> float conversion pair per GROUP (a 64-bit
      // conversion is a multi-instruction sequence), carried to the group's other channels
      float gm = 0.f, gr = 0.f;
#pragma unroll
      for (int i0 = 0; i0 < 8; i0 += 1) {
        if ((i0 & (p.cpg - 1)) == 0) {             // first channel of its group (cpg is 1, 2, 4 or 8 here)
          long long sm = 0, qm = 0;
#pragma unroll
          for (int i = 0; i < 8; ++i)
            if (i >= i0 && i < i0 + p.cpg) { sm += sq[i].x; qm += sq[i].y; }
          gm = __ll2float_rn(sm) * (1.0f / kStatQ1) * inv_cnt;
          gr = rsqrtf(fmaxf(__ll2float_rn(qm) * (1.0f / kStatQ2) * inv_cnt - gm * gm, 0.f) + p.eps);
        }
        mean[i0] = gm;
        rstd[i0] = gr;
      }
    } else {
      // groups of 16, 24, ... channels: whole 8-channel columns of one source
      const int g0 = (c / p.cpg) * p.cpg;
      const long long* st = (g0 < p.C0) ? p.st0 + ((size_t)n * p.C0 + g0) * 2 : p.st1 + ((size_t)n * p.C1 + (g0 - p.C0)) * 2;
      const float2 gs = stat_load_group(st, p.cpg);
      const float m = gs.x * inv_cnt;
      const float r = rsqrtf(fmaxf(gs.y * inv_cnt - m * m, 0.f) + p.eps);
#pragma unroll
      for (int i = 0; i < 8; ++i) { mean[i] = m; rstd[i] = r; }
    }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float4 ga = p.gamma ? ldg4(p.gamma + c + 4 * h) : make_float4(1.f, 1.f, 1.f, 1.f);
      const float4 be = p.beta ? ldg4(p.beta + c + 4 * h) : make_float4(0.f, 0.f, 0.f, 0.f);
      float gv[4] = {ga.x, ga.y, ga.z, ga.w}, bv[4] = {be.x, be.y, be.z, be.w};
      if (p.scale) {
        const float4 sc = ldg4(p.scale + (size_t)n * p.ss_ld + c + 4 * h), sh = ldg4(p.shift + (size_t)n * p.ss_ld + c + 4 * h);
        const float scv[4] = {sc.x, sc.y, sc.z, sc.w}, shv[4] = {sh.x, sh.y, sh.z, sh.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) { gv[i] *= 1.f + scv[i]; bv[i] = bv[i] * (1.f + scv[i]) + shv[i]; }
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        a[4 * h + i] = rstd[4 * h + i] * gv[i];
        b[4 * h + i] = bv[i] - mean[4 * h + i] * rstd[4 * h + i] * gv[i];
      }
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = coefA[c + i]; b[i] = coefB[c + i]; }
  }
  for (;;) {
#pragma unroll
    for (int u = 0; u < kU; ++u) {
      const int q = px + u * pstep;
      if (q >= px1) break;
      float v[8];
      if (kInBf16 || b16) {
        const uint32_t ww[4] = {raw0[u].x, raw0[u].y, raw0[u].z, raw0[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          v[2 * i] = __uint_as_float(ww[i] << 16);
          v[2 * i + 1] = __uint_as_float(ww[i] & 0xffff0000u);
        }
      } else {
        v[0] = __uint_as_float(raw0[u].x); v[1] = __uint_as_float(raw0[u].y);
        v[2] = __uint_as_float(raw0[u].z); v[3] = __uint_as_float(raw0[u].w);
        v[4] = __uint_as_float(raw1[kInBf16 ? 0 : u].x); v[5] = __uint_as_float(raw1[kInBf16 ? 0 : u].y);
        v[6] = __uint_as_float(raw1[kInBf16 ? 0 : u].z); v[7] = __uint_as_float(raw1[kInBf16 ? 0 : u].w);
      }
      float y[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        y[i] = fmaf(v[i], a[i], b[i]);
        if (p.apply_silu) y[i] = silu_tanh(y[i]);
      }
      if (p.drop_thresh) {
        const unsigned long long e = ((unsigned long long)n * p.HW + q) * C + c;
        const uint32_t keep = dropout_keep4(drop_seed, e >> 2, p.drop_thresh) |
                              (dropout_keep4(drop_seed, (e >> 2) + 1, p.drop_thresh) << 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = ((keep >> i) & 1u) ? y[i] * p.drop_scale : 0.f;
      }
      uint4 uo;
      uo.x = pack_bf16x2(y[0], y[1]); uo.y = pack_bf16x2(y[2], y[3]);
      uo.z = pack_bf16x2(y[4], y[5]); uo.w = pack_bf16x2(y[6], y[7]);
      *reinterpret_cast<uint4*>(dst + (size_t)q * C) = uo;
      if (rdst) {
        uint4 ur;
        ur.x = pack_bf16x2(v[0], v[1]); ur.y = pack_bf16x2(v[2], v[3]);
        ur.z = pack_bf16x2(v[4], v[5]); ur.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(rdst + (size_t)q * C) = ur;
      }
    }
    px += kU * pstep;
    if (px >= px1) break;
    issue(px);
  }
}

__global__ void __launch_bounds__(256, 4) groupnorm_apply_kernel(const GnApplyParams p) {
  extern __shared__ float gsm[];
  // dropout seed = per-block offset (by value) + per-forward base read from device memory (CUDA-graph replayable)
  const unsigned long long drop_seed = p.drop_seed + ((p.drop_thresh && p.drop_seed_dev) ? *p.drop_seed_dev : 0ull);
  const int C = p.C0 + p.C1;
  float* coefA = gsm;            // [C]
  float* coefB = gsm + C;        // [C]
  float* chS = gsm + 2 * C;      // [C] channel sums
  float* chQ = gsm + 3 * C;      // [C] channel sums of squares
  const int n = p.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y;
  const int bx = p.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x;
  const int tid = threadIdx.x;
  griddep_sync();
  if (p.resample == 0 && p.cols8 == 2) {   // per-thread coefficients: no shared-memory prologue
    const int px0 = bx * p.pix_per_cta;
    const int px1 = min(p.HW, px0 + p.pix_per_cta);
    if (p.in_bf16 == 1) gn_apply_cols8<true, 4, true>(p, nullptr, nullptr, n, px0, px1, drop_seed);
    else gn_apply_cols8<false, 2, true>(p, nullptr, nullptr, n, px0, px1, drop_seed);
    return;
  }
  for (int c = tid; c < C; c += 256) {
    const long long* st = (c < p.C0) ? p.st0 + ((size_t)n * p.C0 + c) * 2 : p.st1 + ((size_t)n * p.C1 + (c - p.C0)) * 2;
    const float2 sv = stat_load(st);
    chS[c] = sv.x;
    chQ[c] = sv.y;
  }
  __syncthreads();
  const float inv_cnt = 1.0f / (float)(p.HW * p.cpg);
  for (int c = tid; c < C; c += 256) {
    const int g0 = (c / p.cpg) * p.cpg;
    float s = 0.f, q = 0.f;
    for (int i = 0; i < p.cpg; ++i) { s += chS[g0 + i]; q += chQ[g0 + i]; }
    const float mean = s * inv_cnt;
    const float var = fmaxf(q * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    float ga = p.gamma ? __ldg(p.gamma + c) : 1.f;
    float be = p.beta ? __ldg(p.beta + c) : 0.f;
    if (p.scale) {
      const float sc = 1.f + __ldg(p.scale + (size_t)n * p.ss_ld + c);
      ga *= sc;
      be = be * sc + __ldg(p.shift + (size_t)n * p.ss_ld + c);
    }
    coefA[c] = rstd * ga;
    coefB[c] = be - mean * rstd * ga;
  }
  __syncthreads();

  const int nv = C >> 2;
  if (p.resample != 1) {
    const int px0 = bx * p.pix_per_cta;
    const int px1 = min(p.HW, px0 + p.pix_per_cta);
    if (p.resample == 0 && p.cols8) {
      if (p.in_bf16 == 1) gn_apply_cols8<true, 4, false>(p, coefA, coefB, n, px0, px1, drop_seed);
      else gn_apply_cols8<false, 2, false>(p, coefA, coefB, n, px0, px1, drop_seed);
      return;
    }
    if (p.resample == 0 && (256 % nv) == 0) {
      // fast path: a thread owns one 4-channel column (coefficients in registers) and walks down the pixels
      const int j = tid % nv, prow = tid / nv, pstep = 256 / nv;
      const int c = j << 2;
      const float4 a = *reinterpret_cast<const float4*>(coefA + c);
      const float4 b = *reinterpret_cast<const float4*>(coefB + c);
      const bool from0 = c < p.C0;
      const float* src = from0 ? p.x0 + (size_t)n * p.HW * p.C0 + c : p.x1 + (size_t)n * p.HW * p.C1 + (c - p.C0);
      const int sld = from0 ? p.C0 : p.C1;
      __nv_bfloat16* dst = p.out + (size_t)n * p.HW * C + c;
      __nv_bfloat16* rdst = p.raw ? p.raw + (size_t)n * p.HW * C + c : nullptr;
      const __nv_bfloat16* srcb = reinterpret_cast<const __nv_bfloat16*>(p.x0) + (size_t)n * p.HW * p.C0 + c;
      for (int px = px0 + prow; px < px1; px += 4 * pstep) {
        float4 vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u)
          if (px + u * pstep < px1)
            vv[u] = p.in_bf16 ? ldg4_bf16(srcb + (size_t)(px + u * pstep) * sld) : ldg4(src + (size_t)(px + u * pstep) * sld);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int q = px + u * pstep;
          if (q >= px1) break;
          const float4 v = vv[u];
          float y0 = fmaf(v.x, a.x, b.x), y1 = fmaf(v.y, a.y, b.y), y2 = fmaf(v.z, a.z, b.z), y3 = fmaf(v.w, a.w, b.w);
          if (p.apply_silu) { y0 = silu_tanh(y0); y1 = silu_tanh(y1); y2 = silu_tanh(y2); y3 = silu_tanh(y3); }
          if (p.drop_thresh) {
            const unsigned long long e = ((unsigned long long)n * p.HW + q) * C + c;
            const uint32_t keep = dropout_keep4(drop_seed, e >> 2, p.drop_thresh);
            y0 = (keep & 1u) ? y0 * p.drop_scale : 0.f;
            y1 = (keep & 2u) ? y1 * p.drop_scale : 0.f;
            y2 = (keep & 4u) ? y2 * p.drop_scale : 0.f;
            y3 = (keep & 8u) ? y3 * p.drop_scale : 0.f;
          }
          uint2 uo;
          uo.x = pack_bf16x2(y0, y1);
          uo.y = pack_bf16x2(y2, y3);
          *reinterpret_cast<uint2*>(dst + (size_t)q * C) = uo;
          if (rdst) {
            uint2 ur;
            ur.x = pack_bf16x2(v.x, v.y);
            ur.y = pack_bf16x2(v.z, v.w);
            *reinterpret_cast<uint2*>(rdst + (size_t)q * C) = ur;
          }
        }
      }
      return;
    }
    const int total = (px1 - px0) * nv;
    // 4 independent 16-byte loads in flight per thread before any dependent math
    for (int idx0 = tid; idx0 < total; idx0 += 1024) {
      float4 vv[4];
      int pxs[4], cs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int idx = idx0 + u * 256;
        const int pl = idx / nv;
        const int j = idx - pl * nv;
        pxs[u] = px0 + pl;
        cs[u] = j << 2;
        if (idx < total) {
          if (p.in_bf16)
            vv[u] = ldg4_bf16(reinterpret_cast<const __nv_bfloat16*>(p.x0) + ((size_t)n * p.HW + pxs[u]) * p.C0 + cs[u]);
          else
            vv[u] = (cs[u] < p.C0) ? ldg4(p.x0 + ((size_t)n * p.HW + pxs[u]) * p.C0 + cs[u])
                                   : ldg4(p.x1 + ((size_t)n * p.HW + pxs[u]) * p.C1 + (cs[u] - p.C0));
        }
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (idx0 + u * 256 >= total) break;
        const float4 v = vv[u];
        const int px = pxs[u], c = cs[u];
        const float4 a = *reinterpret_cast<const float4*>(coefA + c);
        const float4 b = *reinterpret_cast<const float4*>(coefB + c);
        float y[4] = {v.x * a.x + b.x, v.y * a.y + b.y, v.z * a.z + b.z, v.w * a.w + b.w};
        if (p.apply_silu) {
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = silu_tanh(y[i]);
        }
        if (p.drop_thresh) {
          const unsigned long long e = ((unsigned long long)n * p.HW + px) * C + c;
          const uint32_t keep = dropout_keep4(drop_seed, e >> 2, p.drop_thresh);
#pragma unroll
          for (int i = 0; i < 4; ++i) y[i] = ((keep >> i) & 1u) ? y[i] * p.drop_scale : 0.f;
        }
        uint2 uo;
        uo.x = pack_bf16x2(y[0], y[1]);
        uo.y = pack_bf16x2(y[2], y[3]);
        if (p.resample == 0) {
          *reinterpret_cast<uint2*>(p.out + ((size_t)n * p.HW + px) * C + c) = uo;
        } else {  // nearest 2x
          const int iy = px / p.W, ix = px - iy * p.W;
          const int Wo = p.W * 2;
#pragma unroll
          for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx) {
              const size_t po = (size_t)(2 * iy + dy) * Wo + 2 * ix + dx;
              *reinterpret_cast<uint2*>(p.out + ((size_t)n * (4 * p.HW) + po) * C + c) = uo;
            }
        }
        if (p.raw) {
          uint2 ur;
          ur.x = pack_bf16x2(v.x, v.y);
          ur.y = pack_bf16x2(v.z, v.w);
          *reinterpret_cast<uint2*>(p.raw + ((size_t)n * p.HW + px) * C + c) = ur;
        }
      }
    }
  } else {  // 2x2 average pool of the activated values; the CTA's pixel range is over OUTPUT pixels
    const int Wo = p.W >> 1, HWo = p.HW >> 2;
    const int po0 = bx * p.pix_per_cta;
    const int po1 = min(HWo, po0 + p.pix_per_cta);
    const int total = (po1 - po0) * nv;
    for (int idx = tid; idx < total; idx += 256) {
      const int pl = idx / nv;
      const int j = idx - pl * nv;
      const int po = po0 + pl;
      const int c = j << 2;
      const int oy = po / Wo, ox = po - oy * Wo;
      const float4 a = *reinterpret_cast<const float4*>(coefA + c);
      const float4 b = *reinterpret_cast<const float4*>(coefB + c);
      float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int dy = 0; dy < 2; ++dy)
#pragma unroll
        for (int dx = 0; dx < 2; ++dx) {
          const int px = (2 * oy + dy) * p.W + 2 * ox + dx;
          float4 v;
          if (p.in_bf16) v = ldg4_bf16(reinterpret_cast<const __nv_bfloat16*>(p.x0) + ((size_t)n * p.HW + px) * p.C0 + c);
          else v = (c < p.C0) ? ldg4(p.x0 + ((size_t)n * p.HW + px) * p.C0 + c)
                              : ldg4(p.x1 + ((size_t)n * p.HW + px) * p.C1 + (c - p.C0));
          float y[4] = {v.x * a.x + b.x, v.y * a.y + b.y, v.z * a.z + b.z, v.w * a.w + b.w};
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[i] += p.apply_silu ? silu_tanh(y[i]) : y[i];
        }
      uint2 u;
      u.x = pack_bf16x2(0.25f * acc[0], 0.25f * acc[1]);
      u.y = pack_bf16x2(0.25f * acc[2], 0.25f * acc[3]);
      *reinterpret_cast<uint2*>(p.out + ((size_t)n * HWo + po) * C + c) = u;
    }
  }
}

}  // namespace b200

extern "C" int b200_groupnorm_apply_train_fwd(const void* x0_, int x0_is_bf16, int C0, const long long* stats0,
                                              const float* x1, int C1, const long long* stats1, int B, int HW, int W,
                                              int groups, const float* gamma, const float* beta, float eps,
                                              const float* scale, const float* shift, int ss_ld, int apply_silu,
                                              int resample, float drop_p, unsigned long long drop_seed,
                                              const unsigned long long* drop_seed_dev, void* out_bf16,
                                              void* raw_out_bf16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  const float* x0 = reinterpret_cast<const float*>(x0_);
  // x0 == NULL with C0 > 0: window mode -- channels [0, C0) of out / raw_out were written by the first source's producer
  // (b200_conv2d_gn_fwd, block-output form); this launch normalises the second source into channels [C0, C0 + C1)
  const bool window = x0 == nullptr && C0 > 0 && x1 != nullptr;
  B200_REQUIRE((window || (x0 && stats0)) && out_bf16, "groupnorm_apply: null x0/stats0/out");
  if (!x1) C1 = 0;

  B200_REQUIRE(x1 == nullptr || stats1 != nullptr, "groupnorm_apply: second source needs its statistics");
  const int C = C0 + C1;
  B200_REQUIRE(groups > 0 && C % groups == 0, "groupnorm_apply: C=%d not divisible by groups=%d", C, groups);
  B200_REQUIRE(C0 % 4 == 0 && C1 % 4 == 0, "groupnorm_apply: channel counts (%d, %d) must be multiples of 4", C0, C1);
  B200_REQUIRE(resample >= 0 && resample <= 2, "groupnorm_apply: bad resample mode");
  B200_REQUIRE(W > 0 && HW % W == 0, "groupnorm_apply: HW=%d not a multiple of W=%d", HW, W);
  if (resample == 1) B200_REQUIRE(W % 2 == 0 && (HW / W) % 2 == 0 && raw_out_bf16 == nullptr, "groupnorm_apply: avg-pool needs even H, W and no raw copy");
  B200_REQUIRE((scale == nullptr) == (shift == nullptr), "groupnorm_apply: scale and shift must be given together");
  const size_t smem = (size_t)4 * C * 4;
  B200_REQUIRE(smem <= 48 * 1024, "groupnorm_apply: C=%d too large", C);
  GnApplyParams p;
  p.x0 = x0; p.C0 = C0; p.st0 = stats0; p.x1 = x1; p.C1 = C1; p.st1 = stats1;
  p.HW = HW; p.W = W; p.groups = groups; p.cpg = C / groups;
  p.gamma = gamma; p.beta = beta; p.eps = eps; p.scale = scale; p.shift = shift; p.ss_ld = ss_ld;
  p.apply_silu = apply_silu; p.resample = resample;
  p.in_bf16 = x0_is_bf16 ? ((x1 != nullptr || raw_out_bf16 != nullptr) ? 2 : 1) : 0;   // 2: per-thread dtype (8-channel columns only)
  static const char* env_c8 = getenv("B200_GN_COLS8");
  p.cols8 = (C0 % 8 == 0 && C1 % 8 == 0 && C / 8 <= 256 && !(env_c8 && atoi(env_c8) == 0)) ? 1 : 0;
  {
    const int cpg = C / groups;
    const bool own = (cpg == 1 || cpg == 2 || cpg == 4 || cpg == 8) ||
                     (cpg % 8 == 0 && (C1 == 0 || C0 % cpg == 0));   // a group never straddles the two sources
    if (p.cols8 && own && !(env_c8 && atoi(env_c8) == 1)) p.cols8 = 2;
  }
  B200_REQUIRE(p.in_bf16 != 2 || (p.cols8 && resample == 0 && !window),
               "groupnorm_apply: a bf16 first source with a second source / raw copy needs 8-channel columns and no resampling");
  p.c_begin = window ? C0 : 0;
  if (window)
    B200_REQUIRE(p.cols8 == 2 && resample == 0 && drop_p == 0.f && !x0_is_bf16 && C0 % 8 == 0 && C1 % 8 == 0,
                 "groupnorm_apply: window mode needs whole-group 8-channel columns, no resampling / dropout");
  static const char* env_rev = getenv("B200_L2_REVERSE");
  p.reverse = (env_rev && atoi(env_rev) == 0) ? 0 : 1;
  B200_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "groupnorm_apply: dropout probability %f out of [0,1)", (double)drop_p);
  B200_REQUIRE(drop_p == 0.f || resample == 0, "groupnorm_apply: dropout is not combined with resampling");
  p.drop_p = drop_p; p.drop_scale = 1.0f / (1.0f - drop_p); p.drop_seed = drop_seed; p.drop_seed_dev = drop_seed_dev;
  p.drop_thresh = dropout_threshold(drop_p);
  p.out = reinterpret_cast<__nv_bfloat16*>(out_bf16);
  p.raw = reinterpret_cast<__nv_bfloat16*>(raw_out_bf16);
  const int work_pix = resample == 1 ? HW / 4 : HW;
  int ppc = (resample == 1 ? 8192 : x0_is_bf16 ? 65536 : 32768) / (window ? C1 : C);  // ~128 KB of input per CTA
  if (ppc < 1) ppc = 1;
  if (ppc > work_pix) ppc = work_pix;
  // small tensors: shrink the per-CTA range (down to ~16 KB of input) until the grid covers the SMs once
  // (4 resident CTAs per SM), otherwise a few long CTAs leave most of the memory system idle
  static const char* env_fill = getenv("B200_GN_FILL");
  const long want = env_fill ? atol(env_fill) : 1;   // measured (ncu launch list): shrinking the ranges made the 16x16 / 8x8 layers slower, so off by default
  const int ppc_min = (resample == 1 ? 1024 : x0_is_bf16 ? 8192 : 4096) / C > 8 ? (resample == 1 ? 1024 : x0_is_bf16 ? 8192 : 4096) / C : 8;
  while (ppc > ppc_min && (long)B * ((work_pix + ppc - 1) / ppc) < want) ppc = (ppc + 1) / 2;
  p.pix_per_cta = ppc;
  dim3 grid((work_pix + ppc - 1) / ppc, B);
  B200_CHECK(launch_pdl(groupnorm_apply_kernel, grid, dim3(256), smem, stream, p));
  ++g_launch_count;
  return 0;
}

extern "C" int b200_groupnorm_apply_fwd(const void* x0_, int x0_is_bf16, int C0, const long long* stats0, const float* x1,
                                        int C1, const long long* stats1, int B, int HW, int W, int groups, const float* gamma,
                                        const float* beta, float eps, const float* scale, const float* shift,
                                        int ss_ld, int apply_silu, int resample, void* out_bf16, void* raw_out_bf16,
                                        void* stream_) {
  return b200_groupnorm_apply_train_fwd(x0_, x0_is_bf16, C0, stats0, x1, C1, stats1, B, HW, W, groups, gamma, beta, eps,
                                        scale, shift, ss_ld, apply_silu, resample, 0.f, 0ull, nullptr, out_bf16, raw_out_bf16,
                                        stream_);
}
