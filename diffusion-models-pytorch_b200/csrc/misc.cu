// Small memory-bound kernels around the tensor-core path: the Cin<=4 first convolution, fp32->bf16 casts
// (with the parity split that turns the stride-2 conv into unit-stride TMA boxes), 2x resampling of the fp32
// residual stream, and the time-embedding MLP.
#include "common.cuh"
#include <stdlib.h>
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

// ------------------------------------------------------------------------------------------------
// First convolution (models/unet.py:72,123): x NCHW fp32 [B][Cin][H][W] -> NHWC fp32 [B][H][W][Cout].
// One CTA = (R output rows, image, group of 128 output channels).  Each lane owns 4 output channels and keeps
// their 9*Cin*4 weights in registers; the (R+2) x (W+2) x Cin zero-padded input patch sits in shared memory and is
// read as warp-wide broadcasts; a warp emits one pixel's 128 channels per iteration (coalesced 512-byte store).
// The per-(image, channel) sum / sum of squares for the GroupNorm that follows are accumulated on the fly.
// ------------------------------------------------------------------------------------------------
template <int CIN>
__global__ void __launch_bounds__(256) conv3x3_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            long long* __restrict__ stats, int B, int H, int W, int Cout,
                                                            int R) {
  constexpr int K = 9 * CIN;
  extern __shared__ float fsm[];
  unsigned long long* ssm = reinterpret_cast<unsigned long long*>(fsm);   // [2*128] fixed-point statistics (stat_add)
  float* patch = fsm + 512;       // [CIN][R+2][W+2], zero padded
  const int n = blockIdx.y;
  const int h0 = blockIdx.x * R;
  const int PW = W + 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int co = blockIdx.z * 128 + 4 * lane;  // first of this lane's 4 output channels
  const bool co_ok = co < Cout;
  griddep_sync();
  for (int i = threadIdx.x; i < 256; i += blockDim.x) ssm[i] = 0ull;
  for (int i = threadIdx.x; i < CIN * (R + 2) * PW; i += blockDim.x) {
    const int ci = i / ((R + 2) * PW);
    const int rem = i - ci * (R + 2) * PW;
    const int py = rem / PW, pxx = rem - py * PW;
    const int yy = h0 + py - 1, xx = pxx - 1;
    patch[i] = (yy >= 0 && yy < H && xx >= 0 && xx < W) ? __ldg(x + (((size_t)n * CIN + ci) * H + yy) * W + xx) : 0.f;
  }
  // weights of channels co..co+3: rows of K contiguous floats in the OIHW tensor (k = (ci*3 + r)*3 + s)
  float wr[4][K];
  float bv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int k = 0; k < K; ++k) wr[e][k] = co_ok ? __ldg(w + (size_t)(co + e) * K + k) : 0.f;
    bv[e] = (co_ok && bias) ? __ldg(bias + co + e) : 0.f;
  }
  __syncthreads();
  const int rows = min(R, H - h0);
  float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
  for (int lp = warp; lp < rows * W; lp += nwarps) {
    const int ly = lp / W, lx = lp - ly * W;
    float acc[4] = {bv[0], bv[1], bv[2], bv[3]};
#pragma unroll
    for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const float v = patch[(ci * (R + 2) + ly + r) * PW + lx + s];
          const int k = (ci * 3 + r) * 3 + s;
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[e] += v * wr[e][k];
        }
    if (co_ok) {
      const size_t pix = ((size_t)n * H + h0 + ly) * W + lx;
      *reinterpret_cast<float4*>(out + pix * Cout + co) = make_float4(acc[0], acc[1], acc[2], acc[3]);
#pragma unroll
      for (int e = 0; e < 4; ++e) { s1[e] += acc[e]; s2[e] += acc[e] * acc[e]; }
    }
  }
  if (stats) {
    if (co_ok) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        atomicAdd(ssm + 2 * (4 * lane + e), stat_fix1(s1[e]));
        atomicAdd(ssm + 2 * (4 * lane + e) + 1, stat_fix2(s2[e]));
      }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 256; i += blockDim.x) {
      const int ch = blockIdx.z * 128 + (i >> 1);
      if (ch < Cout) atomicAdd(reinterpret_cast<unsigned long long*>(stats) + ((size_t)n * Cout + ch) * 2 + (i & 1), ssm[i]);
    }
  }
}

// Persistent variant for W % 4 == 0 (every model family): a CTA keeps its lanes' weights in registers across
// (image, row block) units; a lane computes 4 consecutive pixels x 4 channels, so each (ci, r) patch row is read as
// one 16-byte + one 8-byte warp-broadcast shared-memory load feeding 48 FMAs (the 1-pixel form issued one 4-byte
// load per 4 FMAs and ran at 22 % of the FP32 rate).
// kF2: packed FP32 FMAs (FFMA2, sm_100: two fp32 FMAs per instruction).  The accumulators and the weights of a lane's
// channel pairs (0, 1) / (2, 3) are natural register pairs; the pixel operand has to carry the same value in both halves,
// so the patch is stored DUPLICATED in shared memory ([x, x] per pixel) and a patch row is three 16-byte broadcast
// loads feeding 24 FFMA2 (= the 48 FMAs of the scalar form, same rounding: bitwise identical results).
__device__ __forceinline__ void ffma2_acc(float2& d, const float2 a, const float2 b) {
  unsigned long long dd = *reinterpret_cast<unsigned long long*>(&d);
  asm("fma.rn.f32x2 %0, %1, %2, %0;"
      : "+l"(dd)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)));
  d = *reinterpret_cast<float2*>(&dd);
}

template <int CIN, bool kF2>
__global__ void __launch_bounds__(256, 1) conv3x3_first_px4_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ out,
                                                                   long long* __restrict__ stats, int B, int H, int W, int Cout,
                                                                   int R, int units_per_image, int total_units) {
  constexpr int K = 9 * CIN;
  extern __shared__ float fsm[];
  unsigned long long* ssm = reinterpret_cast<unsigned long long*>(fsm);   // [2*128] fixed-point statistics of the unit
  float* patch = fsm + 512;       // [CIN][R+2][PWp], zero padded, rows 16-byte aligned
  const int PWp = (W + 2 + 3) & ~3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int co = blockIdx.z * 128 + 4 * lane;
  const bool co_ok = co < Cout;
  float wr[4][K];
  float bv[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
#pragma unroll
    for (int k = 0; k < K; ++k) wr[e][k] = co_ok ? __ldg(w + (size_t)(co + e) * K + k) : 0.f;
    bv[e] = (co_ok && bias) ? __ldg(bias + co + e) : 0.f;
  }
  griddep_sync();
  const int gpr = W >> 2;   // 4-pixel groups per row
  for (int unit = blockIdx.x; unit < total_units; unit += gridDim.x) {
    const int n = unit / units_per_image;
    const int h0 = (unit - n * units_per_image) * R;
    __syncthreads();   // the previous unit's patch and statistics are consumed
    for (int i = threadIdx.x; i < 256; i += blockDim.x) ssm[i] = 0ull;
    // patch fill in batches of 8 independent global loads per thread (one load per loop iteration left every thread
    // waiting for 4-5 serial L2 / DRAM round trips per unit: ncu's top stall site of this kernel, 14 % of the samples)
    const int ptotal = CIN * (R + 2) * PWp;
    for (int i0 = threadIdx.x; i0 < ptotal; i0 += 8 * blockDim.x) {
      float pvv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        pvv[u] = 0.f;
        if (i < ptotal) {
          const int ci = i / ((R + 2) * PWp);
          const int rem = i - ci * (R + 2) * PWp;
          const int py = rem / PWp, pxx = rem - py * PWp;
          const int yy = h0 + py - 1, xx = pxx - 1;
          if (yy >= 0 && yy < H && xx >= 0 && xx < W) pvv[u] = __ldg(x + (((size_t)n * CIN + ci) * H + yy) * W + xx);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int i = i0 + u * blockDim.x;
        if (i < ptotal) {
          if (kF2) reinterpret_cast<float2*>(patch)[i] = make_float2(pvv[u], pvv[u]);
          else patch[i] = pvv[u];
        }
      }
    }
    __syncthreads();
    const int rows = min(R, H - h0);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    for (int g = warp; g < rows * gpr; g += nwarps) {
      const int ly = g / gpr, lx = (g - ly * gpr) << 2;
      float acc[4][4];
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[q][e] = bv[e];
      if (kF2) {
        float2 a2[4][2];
#pragma unroll
        for (int q = 0; q < 4; ++q) { a2[q][0] = make_float2(bv[0], bv[1]); a2[q][1] = make_float2(bv[2], bv[3]); }
#pragma unroll
        for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
          for (int r = 0; r < 3; ++r) {
            const float4* pr = reinterpret_cast<const float4*>(reinterpret_cast<const float2*>(patch) +
                                                               (ci * (R + 2) + ly + r) * PWp + lx);
            const float4 u0 = pr[0], u1 = pr[1], u2 = pr[2];      // six pixels, each as [x, x]
            const float2 pv2[6] = {make_float2(u0.x, u0.y), make_float2(u0.z, u0.w), make_float2(u1.x, u1.y),
                                   make_float2(u1.z, u1.w), make_float2(u2.x, u2.y), make_float2(u2.z, u2.w)};
#pragma unroll
            for (int s = 0; s < 3; ++s) {
              const int k = (ci * 3 + r) * 3 + s;
              const float2 w01 = make_float2(wr[0][k], wr[1][k]), w23 = make_float2(wr[2][k], wr[3][k]);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                ffma2_acc(a2[q][0], pv2[q + s], w01);
                ffma2_acc(a2[q][1], pv2[q + s], w23);
              }
            }
          }
#pragma unroll
        for (int q = 0; q < 4; ++q) { acc[q][0] = a2[q][0].x; acc[q][1] = a2[q][0].y; acc[q][2] = a2[q][1].x; acc[q][3] = a2[q][1].y; }
      } else
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci)
#pragma unroll
        for (int r = 0; r < 3; ++r) {
          const float* pr = patch + (ci * (R + 2) + ly + r) * PWp + lx;
          const float4 v0 = *reinterpret_cast<const float4*>(pr);
          const float2 v1 = *reinterpret_cast<const float2*>(pr + 4);
          const float pv[6] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y};
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int k = (ci * 3 + r) * 3 + s;
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
              for (int e = 0; e < 4; ++e) acc[q][e] = fmaf(pv[q + s], wr[e][k], acc[q][e]);
          }
        }
      if (co_ok) {
        const size_t pix = ((size_t)n * H + h0 + ly) * W + lx;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          *reinterpret_cast<float4*>(out + (pix + q) * Cout + co) = make_float4(acc[q][0], acc[q][1], acc[q][2], acc[q][3]);
#pragma unroll
          for (int e = 0; e < 4; ++e) { s1[e] += acc[q][e]; s2[e] = fmaf(acc[q][e], acc[q][e], s2[e]); }
        }
      }
    }
    if (stats) {
      if (co_ok) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          atomicAdd(ssm + 2 * (4 * lane + e), stat_fix1(s1[e]));
          atomicAdd(ssm + 2 * (4 * lane + e) + 1, stat_fix2(s2[e]));
        }
      }
      __syncthreads();
      for (int i = threadIdx.x; i < 256; i += blockDim.x) {
        const int ch = blockIdx.z * 128 + (i >> 1);
        if (ch < Cout) atomicAdd(reinterpret_cast<unsigned long long*>(stats) + ((size_t)n * Cout + ch) * 2 + (i & 1), ssm[i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 NHWC -> bf16 NHWC (optionally into 4 parity planes)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        int B, int H, int W, int C, int parity, int reverse) {
  const int cv = C >> 2;
  const size_t total = (size_t)B * H * W * cv;
  griddep_sync();
  // 4 independent 16-byte loads in flight per thread (one load per thread left most of the HBM latency exposed)
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < total; i0 += 4 * stride) {
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t ii = i0 + u * stride;
      // reverse: start from the end of the tensor, where the producer's most recent writes are still in L2
      if (ii < total) v[u] = __ldg(reinterpret_cast<const float4*>(x) + (reverse ? total - 1 - ii : ii));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const size_t ii = i0 + u * stride;
      if (ii >= total) break;
      const size_t i = reverse ? total - 1 - ii : ii;
      uint2 w2;
      w2.x = pack_bf16x2(v[u].x, v[u].y);
      w2.y = pack_bf16x2(v[u].z, v[u].w);
      size_t o = i;
      if (parity) {
        const int c4 = (int)(i % cv);
        const size_t pix = i / cv;
        const int wq = (int)(pix % W);
        const int h = (int)((pix / W) % H);
        const int n = (int)(pix / ((size_t)W * H));
        const int pl = ((h & 1) << 1) | (wq & 1);
        o = ((((size_t)n * 4 + pl) * (H >> 1) + (h >> 1)) * (W >> 1) + (wq >> 1)) * cv + c4;
      }
      reinterpret_cast<uint2*>(out)[o] = w2;
    }
  }
}

// Fast form for power-of-two H, W and C/8 (every UNet family here): a thread owns 8-channel units (two 16-byte loads ->
// one 16-byte store, 4 units = 128 B of loads in flight), all index arithmetic is 32-bit shifts and masks (the generic
// kernel above spends four 64-bit divisions per 16 bytes in the parity mode and ran at 3.0 TB/s), one-wave grid.
__global__ void __launch_bounds__(256, 4) cast_bf16_c8_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                              uint32_t total, int hs, int ws, int cs, int parity,
                                                              int reverse) {
  const float4* __restrict__ src = reinterpret_cast<const float4*>(x);
  uint4* __restrict__ dst = reinterpret_cast<uint4*>(out);
  const uint32_t stride = gridDim.x * 256u;
  const uint32_t cmask = (1u << cs) - 1u, wmask = (1u << ws) - 1u, hmask = (1u << hs) - 1u;
  griddep_sync();
  for (uint32_t u0 = blockIdx.x * 256u + threadIdx.x; u0 < total; u0 += 4u * stride) {
    float4 a[4], b[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t uu = u0 + (uint32_t)k * stride;
      if (uu < total) {
        // reverse: start from the end of the tensor, where the producer's most recent writes are still in L2
        const uint32_t i = reverse ? total - 1u - uu : uu;
        a[k] = __ldg(src + 2 * (size_t)i);
        b[k] = __ldg(src + 2 * (size_t)i + 1);
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t uu = u0 + (uint32_t)k * stride;
      if (uu >= total) break;
      const uint32_t i = reverse ? total - 1u - uu : uu;
      uint4 w4;
      w4.x = pack_bf16x2(a[k].x, a[k].y);
      w4.y = pack_bf16x2(a[k].z, a[k].w);
      w4.z = pack_bf16x2(b[k].x, b[k].y);
      w4.w = pack_bf16x2(b[k].z, b[k].w);
      uint32_t o = i;
      if (parity) {
        const uint32_t cc = i & cmask, pix = i >> cs;
        const uint32_t wq = pix & wmask, h = (pix >> ws) & hmask, n = pix >> (ws + hs);
        const uint32_t pl = ((h & 1u) << 1) | (wq & 1u);
        o = ((((((n << 2) | pl) << (hs - 1)) | (h >> 1)) << (ws - 1)) | (wq >> 1)) << cs | cc;
      }
      dst[o] = w4;
    }
  }
}

__global__ void __launch_bounds__(256) avgpool2_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int B,
                                                           int H, int W, int C) {
  const int cv = C >> 2, Ho = H >> 1, Wo = W >> 1;
  const size_t total = (size_t)B * Ho * Wo * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cv);
    const size_t po = i / cv;
    const int ox = (int)(po % Wo);
    const int oy = (int)((po / Wo) % Ho);
    const int n = (int)(po / ((size_t)Wo * Ho));
    const float4* src = reinterpret_cast<const float4*>(x) + (((size_t)n * H + 2 * oy) * W + 2 * ox) * cv + c4;
    const float4 a = __ldg(src), b = __ldg(src + cv), c = __ldg(src + (size_t)W * cv), d = __ldg(src + (size_t)W * cv + cv);
    float4 r;
    r.x = 0.25f * (a.x + b.x + c.x + d.x);
    r.y = 0.25f * (a.y + b.y + c.y + d.y);
    r.z = 0.25f * (a.z + b.z + c.z + d.z);
    r.w = 0.25f * (a.w + b.w + c.w + d.w);
    reinterpret_cast<float4*>(out)[i] = r;
  }
}

__global__ void __launch_bounds__(256) upsample2_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int B,
                                                            int H, int W, int C) {
  const int cv = C >> 2, Ho = H * 2, Wo = W * 2;
  const size_t total = (size_t)B * Ho * Wo * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cv);
    const size_t po = i / cv;
    const int ox = (int)(po % Wo);
    const int oy = (int)((po / Wo) % Ho);
    const int n = (int)(po / ((size_t)Wo * Ho));
    reinterpret_cast<float4*>(out)[i] =
        __ldg(reinterpret_cast<const float4*>(x) + (((size_t)n * H + (oy >> 1)) * W + (ox >> 1)) * cv + c4);
  }
}

// ------------------------------------------------------------------------------------------------
// Time embedding (models/modules.py:40-57 + models/unet.py:64-69): one CTA per row of t.
//   pe = [sin(t f), cos(t f)] (or [cos, sin] for ADM), h = SiLU(W1 pe + b1), e = W2 h + b2 (+ class row)
// Outputs e as fp32 [rows][E] and SiLU(e) as bf16 [rows][E] (the operand of the per-block projections).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) time_embed_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                                         int dim, int E, int cos_first, const float* __restrict__ w1,
                                                         const float* __restrict__ b1, const float* __restrict__ w2,
                                                         const float* __restrict__ b2, const int64_t* __restrict__ y,
                                                         const float* __restrict__ class_embed, float* __restrict__ out,
                                                         __nv_bfloat16* __restrict__ out_silu_bf16) {
  // grid = (splits, rows): every CTA recomputes the cheap hidden layer and produces E/splits outputs of layer 2
  extern __shared__ float tsm[];  // pe[dim] + h[E]
  float* pe = tsm;
  float* hid = tsm + dim;
  const int b = blockIdx.y;
  const int half = dim >> 1;
  const float tv = (float)t[b];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = tv * freqs[i];
    const float s = sinf(a), c = cosf(a);
    pe[i] = cos_first ? c : s;
    pe[half + i] = cos_first ? s : c;
  }
  __syncthreads();
  // layer 1: thread-per-output, each thread streams its own weight row with float4 loads (dim % 4 == 0)
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float4* wr = reinterpret_cast<const float4*>(w1 + (size_t)e * dim);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k = 0; k < dim / 4; ++k) {
      const float4 wv = __ldg(wr + k);
      a0 += wv.x * pe[4 * k]; a1 += wv.y * pe[4 * k + 1]; a2 += wv.z * pe[4 * k + 2]; a3 += wv.w * pe[4 * k + 3];
    }
    const float v = (a0 + a1) + (a2 + a3) + b1[e];
    hid[e] = v / (1.0f + expf(-v));
  }
  __syncthreads();
  // layer 2: 8 lanes per output, coalesced float4 reads of the weight row, shuffle reduction
  const int per_cta = (E + gridDim.x - 1) / gridDim.x;
  const int e_begin = blockIdx.x * per_cta;
  const int e_end = min(E, e_begin + per_cta);
  const int sub = threadIdx.x & 7, grp = threadIdx.x >> 3;  // 32 groups of 8 lanes
  for (int e0 = e_begin; e0 < e_end; e0 += 32) {
    const int e = e0 + grp;
    float acc = 0.f;
    if (e < e_end) {
      const float4* wr = reinterpret_cast<const float4*>(w2 + (size_t)e * E);
      for (int k = sub; k < E / 4; k += 8) {
        const float4 wv = __ldg(wr + k);
        acc += wv.x * hid[4 * k] + wv.y * hid[4 * k + 1] + wv.z * hid[4 * k + 2] + wv.w * hid[4 * k + 3];
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 4);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    if (sub == 0 && e < e_end) {
      float v = acc + b2[e];
      if (y && class_embed && y[b] >= 0) v += class_embed[(size_t)y[b] * E + e];   // y < 0: unconditional row (batched CFG)
      out[(size_t)b * E + e] = v;
      if (out_silu_bf16) out_silu_bf16[(size_t)b * E + e] = __float2bfloat16_rn(v / (1.0f + expf(-v)));
    }
  }
}

// q(x_t | x_0), per-sample t (diffusions/ddpm.py:152-172)
__global__ void __launch_bounds__(256) diffuse_kernel(const float* __restrict__ x0, const float* __restrict__ eps,
                                                      const int64_t* __restrict__ t, const float* __restrict__ ac,
                                                      float* __restrict__ xt, int B, int CHW, int total_steps) {
  const int b = blockIdx.y;
  // the reference indexes alphas_cumprod[t] and raises on an out-of-range timestep (ddpm.py:166); a kernel cannot
  // raise, so it never reads out of bounds and poisons that sample with NaN (the loss then shows it at once)
  const int64_t tb = t[b];
  const bool in_range = tb >= 0 && tb < (int64_t)total_steps;
  const float a = in_range ? ac[tb] : __int_as_float(0x7fc00000);
  const float sa = sqrtf(a), sb = sqrtf(1.0f - a);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < CHW; i += gridDim.x * blockDim.x) {
    const size_t o = (size_t)b * CHW + i;
    xt[o] = sa * x0[o] + sb * eps[o];
  }
}

static inline int ew_grid(size_t total, int block) {
  size_t g = (total + block - 1) / block;
  const size_t cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace b200

using namespace b200;

template <int CIN>
static int launch_first(const float* x, const float* w, const float* bias, float* out, long long* stats, int B, int H,
                        int W, int Cout, cudaStream_t stream) {
  static const char* env_v = getenv("B200_FIRST_PX4");
  if (W % 4 == 0 && !(env_v && atoi(env_v) == 0)) {
    int R = 8;
    if (R > H) R = H;
    const int PWp = (W + 2 + 3) & ~3;
    static const char* env_f2 = getenv("B200_FIRST_F2");       // =0: scalar FMAs (A/B)
    const bool f2 = !(env_f2 && atoi(env_f2) == 0);
    const size_t smem = (512 + (size_t)CIN * (R + 2) * PWp * (f2 ? 2 : 1)) * 4;
    B200_REQUIRE(smem <= 100 * 1024, "conv3x3_first: input patch does not fit in shared memory (W=%d)", W);
    static bool attr4 = false;
    static int sms = 0;
    if (!attr4) {
      int dev = 0;
      B200_CHECK(cudaGetDevice(&dev));
      B200_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
      B200_CHECK(cudaFuncSetAttribute(conv3x3_first_px4_kernel<CIN, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      B200_CHECK(cudaFuncSetAttribute(conv3x3_first_px4_kernel<CIN, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
      attr4 = true;
    }
    const int upi = (H + R - 1) / R;
    const int total = upi * B;
    dim3 grid(total < sms ? total : sms, 1, (Cout + 127) / 128);
    if (f2)
      B200_CHECK(launch_pdl(conv3x3_first_px4_kernel<CIN, true>, grid, dim3(256), smem, stream, x, w, bias, out, stats, B, H, W,
                            Cout, R, upi, total));
    else
      B200_CHECK(launch_pdl(conv3x3_first_px4_kernel<CIN, false>, grid, dim3(256), smem, stream, x, w, bias, out, stats, B, H, W,
                            Cout, R, upi, total));
    ++g_launch_count;
    return 0;
  }
  // rows per CTA: enough pixels per warp (R * W / 8) to amortise the per-lane weight fetch (108 floats for Cin = 3)
  static const char* env_r = getenv("B200_FIRST_ROWS");
  int R = env_r ? atoi(env_r) : 32;
  while (R > 1 && (size_t)CIN * (R + 2) * (W + 2) * 4 > 96 * 1024) R >>= 1;
  if (R > H) R = H;
  const size_t smem = (512 + (size_t)CIN * (R + 2) * (W + 2)) * 4;
  B200_REQUIRE(smem <= 100 * 1024, "conv3x3_first: input patch does not fit in shared memory (W=%d)", W);
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(conv3x3_first_kernel<CIN>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr = true;
  }
  dim3 grid((H + R - 1) / R, B, (Cout + 127) / 128);
  B200_CHECK(launch_pdl(conv3x3_first_kernel<CIN>, grid, dim3(256), smem, stream, x, w, bias, out, stats, B, H, W, Cout, R));
  ++g_launch_count;
  return 0;
}

extern "C" int b200_conv3x3_first(const float* x, const float* w, const float* bias, float* out, long long* stats, int B,
                                  int Cin, int H, int W, int Cout, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && w && out, "conv3x3_first: null pointer");
  B200_REQUIRE(Cin >= 1 && Cin <= 4, "conv3x3_first: Cin=%d must be in [1,4]", Cin);
  B200_REQUIRE(Cout % 4 == 0, "conv3x3_first: Cout=%d must be a multiple of 4", Cout);
  switch (Cin) {
    case 1: return launch_first<1>(x, w, bias, out, stats, B, H, W, Cout, stream);
    case 2: return launch_first<2>(x, w, bias, out, stats, B, H, W, Cout, stream);
    case 3: return launch_first<3>(x, w, bias, out, stats, B, H, W, Cout, stream);
    default: return launch_first<4>(x, w, bias, out, stats, B, H, W, Cout, stream);
  }
}

// Input image -> the tensor-core form of the first convolution (models/unet.py:72,123 on K1 instead of FP32 FMAs):
// NCHW fp32 [B][Cin][H][W] -> bf16 NHWC [B][H][W][64], channels [hi(Cin) | lo(Cin) | hi(Cin) | 0 ...] with
// hi = bf16(x), lo = bf16(x - hi).  Against weights packed [w_hi | w_hi | w_lo | 0 ...] per tap (pack mode 4) the
// 64-channel contraction evaluates x*w as x_hi*w_hi + x_lo*w_hi + x_hi*w_lo: fp32-grade products on the bf16 tensor
// pipe (the image is the one operand of the network that is NOT rounded to bf16 by the reference-faithful path).
// Eight threads per pixel, one 16-byte store each: a warp writes 512 contiguous bytes.
namespace b200 {
__global__ void __launch_bounds__(256) first_split_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                          long long npix, int HW, int Cin) {
  const long long gid = (long long)blockIdx.x * 256 + threadIdx.x;
  const long long pix = gid >> 3;
  if (pix >= npix) return;
  const int part = (int)(gid & 7);
  uint4 u = make_uint4(0u, 0u, 0u, 0u);
  if (part * 8 < 3 * Cin) {
    const long long n = pix / HW;
    const long long base = n * (long long)Cin * HW + (pix - n * HW);
    __nv_bfloat16 v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = part * 8 + j;
      float r = 0.f;
      if (k < 3 * Cin) {
        const int kind = k / Cin, c = k - kind * Cin;
        const float f = __ldg(x + base + (long long)c * HW);
        const __nv_bfloat16 hi = __float2bfloat16_rn(f);
        r = kind == 1 ? f - __bfloat162float(hi) : __bfloat162float(hi);
      }
      v[j] = __float2bfloat16_rn(r);
    }
    u = *reinterpret_cast<const uint4*>(v);
  }
  *reinterpret_cast<uint4*>(out + gid * 8) = u;
}
}  // namespace b200

extern "C" int b200_first_split(const float* x, void* out, int B, int Cin, int H, int W, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out, "first_split: null pointer");
  B200_REQUIRE(Cin >= 1 && 3 * Cin <= 64, "first_split: Cin=%d must be in [1,21]", Cin);
  B200_REQUIRE(((uintptr_t)out & 15) == 0, "first_split: out must be 16-byte aligned");
  const long long npix = (long long)B * H * W;
  const long long ctas = (npix * 8 + 255) / 256;
  B200_REQUIRE(ctas <= 0x7fffffffLL, "first_split: tensor too large");
  b200::first_split_kernel<<<(unsigned)ctas, 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), npix, H * W, Cin);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "first_split_kernel launch");
}

extern "C" int b200_cast_bf16(const float* x, void* out, int B, int H, int W, int C, int parity_split, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out, "cast_bf16: null pointer");
  B200_REQUIRE(C % 4 == 0, "cast_bf16: C=%d must be a multiple of 4", C);
  if (parity_split) B200_REQUIRE(H % 2 == 0 && W % 2 == 0, "cast_bf16: parity split needs even H, W");
  const size_t total = (size_t)B * H * W * (C / 4);
  static const char* env_rev = getenv("B200_L2_REVERSE");
  const int reverse = (env_rev && atoi(env_rev) == 0) ? 0 : 1;
  auto log2_exact = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (1 << s) == v ? s : -1; };
  const int hs = log2_exact(H), ws = log2_exact(W), cs = C % 8 == 0 ? log2_exact(C / 8) : -1;
  const size_t units = (size_t)B * H * W * (C / 8);
  static const char* env_c8 = getenv("B200_CAST_C8");
  const bool fast = hs >= (parity_split ? 1 : 0) && ws >= (parity_split ? 1 : 0) && cs >= 0 && units < (1ull << 31) &&
                    (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 &&
                    !(env_c8 && atoi(env_c8) == 0);
  if (fast) {
    size_t g = (units + 1023) / 1024;      // 4 units per thread per iteration
    if (g > 148 * 4) g = 148 * 4;
    if (g == 0) g = 1;
    B200_CHECK(launch_pdl(cast_bf16_c8_kernel, dim3((unsigned)g), dim3(256), 0, stream, x,
                          reinterpret_cast<__nv_bfloat16*>(out), (uint32_t)units, hs, ws, cs, parity_split, reverse));
  } else {
    B200_CHECK(launch_pdl(cast_bf16_kernel, dim3(ew_grid(total, 256)), dim3(256), 0, stream, x,
                          reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C, parity_split, reverse));
  }
  ++g_launch_count;
  return 0;
}

extern "C" int b200_avgpool2_f32(const float* x, float* out, int B, int H, int W, int C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out && C % 4 == 0 && H % 2 == 0 && W % 2 == 0, "avgpool2: bad arguments");
  const size_t total = (size_t)B * (H / 2) * (W / 2) * (C / 4);
  avgpool2_f32_kernel<<<ew_grid(total, 256), 256, 0, stream>>>(x, out, B, H, W, C);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "avgpool2 launch");
}

extern "C" int b200_upsample2_f32(const float* x, float* out, int B, int H, int W, int C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out && C % 4 == 0, "upsample2: bad arguments");
  const size_t total = (size_t)B * (H * 2) * (W * 2) * (C / 4);
  upsample2_f32_kernel<<<ew_grid(total, 256), 256, 0, stream>>>(x, out, B, H, W, C);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "upsample2 launch");
}

extern "C" int b200_time_embed(const int64_t* t, int rows, const float* freqs, int dim, int E, int cos_first,
                               const float* w1, const float* b1, const float* w2, const float* b2, const int64_t* y,
                               const float* class_embed, float* out, void* out_silu_bf16, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(t && freqs && w1 && b1 && w2 && b2 && out, "time_embed: null pointer");
  B200_REQUIRE(dim % 8 == 0 && E % 4 == 0 && rows >= 1, "time_embed: dim must be a multiple of 8, E of 4");
  const size_t smem = (size_t)(dim + E) * 4;
  B200_REQUIRE(smem <= 48 * 1024, "time_embed: dim+E too large");
  const int splits = rows >= 64 ? 2 : 8;
  time_embed_kernel<<<dim3(splits, rows), 256, smem, stream>>>(t, freqs, dim, E, cos_first, w1, b1, w2, b2, y, class_embed, out,
                                                 reinterpret_cast<__nv_bfloat16*>(out_silu_bf16));
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "time_embed launch");
}

extern "C" int b200_diffuse(const float* x0, const float* eps, const int64_t* t, const float* alphas_cumprod, float* xt,
                            int B, int CHW, int total_steps, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x0 && eps && t && alphas_cumprod && xt, "diffuse: null pointer");
  B200_REQUIRE(total_steps >= 1, "diffuse: total_steps must be positive");
  dim3 grid((CHW + 1023) / 1024 > 64 ? 64 : (CHW + 1023) / 1024, B);
  diffuse_kernel<<<grid, 256, 0, stream>>>(x0, eps, t, alphas_cumprod, xt, B, CHW, total_steps);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "diffuse launch");
}
