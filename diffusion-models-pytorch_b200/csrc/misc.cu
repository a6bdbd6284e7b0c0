// Small memory-bound kernels around the tensor-core path: the Cin<=4 first convolution, fp32->bf16 casts
// (with the parity split that turns the stride-2 conv into unit-stride TMA boxes), 2x resampling of the fp32
// residual stream, and the time-embedding MLP.
#include "common.cuh"
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

// ------------------------------------------------------------------------------------------------
// First convolution (models/unet.py:72,123): x NCHW fp32 [B][Cin][H][W] -> NHWC fp32 [B][H][W][Cout].
// A warp produces all Cout channels of one pixel per iteration (coalesced 4*Cout-byte row store);
// the 9*Cin inputs of that pixel are warp-uniform broadcast loads, the weights sit in smem as [9*Cin][Cout].
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) conv3x3_first_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                            const float* __restrict__ bias, float* __restrict__ out,
                                                            int B, int Cin, int H, int W, int Cout) {
  extern __shared__ float wsm[];  // [9*Cin][Cout] + bias[Cout]
  const int K = 9 * Cin;
  for (int i = threadIdx.x; i < K * Cout; i += blockDim.x) {
    const int k = i / Cout, co = i - k * Cout;  // k = (ci*3 + r)*3 + s  (OIHW inner order)
    wsm[i] = w[(size_t)co * K + k];
  }
  float* bsm = wsm + K * Cout;
  for (int i = threadIdx.x; i < Cout; i += blockDim.x) bsm[i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_total = gridDim.x * (blockDim.x >> 5);
  const int gwarp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int npix = B * H * W;
  const int quads = Cout >> 2;
  for (int pix = gwarp; pix < npix; pix += warps_total) {
    const int n = pix / (H * W);
    const int hw = pix - n * H * W;
    const int h = hw / W, wq = hw - h * W;
    float in[36];
#pragma unroll
    for (int ci = 0; ci < 4; ++ci)
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int s = 0; s < 3; ++s) {
          const int yy = h + r - 1, xx = wq + s - 1;
          float v = 0.f;
          if (ci < Cin && yy >= 0 && yy < H && xx >= 0 && xx < W)
            v = __ldg(x + (((size_t)n * Cin + ci) * H + yy) * W + xx);
          in[(ci * 3 + r) * 3 + s] = v;
        }
    for (int qd = lane; qd < quads; qd += 32) {
      float4 acc = *reinterpret_cast<const float4*>(bsm + 4 * qd);
#pragma unroll
      for (int k = 0; k < 36; ++k) {
        if (k < K) {
          const float4 wv = *reinterpret_cast<const float4*>(wsm + k * Cout + 4 * qd);
          acc.x += in[k] * wv.x; acc.y += in[k] * wv.y; acc.z += in[k] * wv.z; acc.w += in[k] * wv.w;
        }
      }
      *reinterpret_cast<float4*>(out + (size_t)pix * Cout + 4 * qd) = acc;
    }
  }
}

// ------------------------------------------------------------------------------------------------
// fp32 NHWC -> bf16 NHWC (optionally into 4 parity planes)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cast_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                        int B, int H, int W, int C, int parity) {
  const int cv = C >> 2;
  const size_t total = (size_t)B * H * W * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + i);
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
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
    reinterpret_cast<uint2*>(out)[o] = u;
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
  extern __shared__ float tsm[];  // pe[dim] + h[E]
  float* pe = tsm;
  float* hid = tsm + dim;
  const int b = blockIdx.x;
  const int half = dim >> 1;
  const float tv = (float)t[b];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = tv * freqs[i];
    const float s = sinf(a), c = cosf(a);
    pe[i] = cos_first ? c : s;
    pe[half + i] = cos_first ? s : c;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int e = warp; e < E; e += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < dim; k += 32) acc += __ldg(w1 + (size_t)e * dim + k) * pe[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      const float v = acc + b1[e];
      hid[e] = v / (1.0f + expf(-v));
    }
  }
  __syncthreads();
  for (int e = warp; e < E; e += nwarps) {
    float acc = 0.f;
    for (int k = lane; k < E; k += 32) acc += __ldg(w2 + (size_t)e * E + k) * hid[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) {
      float v = acc + b2[e];
      if (y && class_embed) v += class_embed[(size_t)y[b] * E + e];
      out[(size_t)b * E + e] = v;
      if (out_silu_bf16) out_silu_bf16[(size_t)b * E + e] = __float2bfloat16_rn(v / (1.0f + expf(-v)));
    }
  }
}

// q(x_t | x_0), per-sample t (diffusions/ddpm.py:152-172)
__global__ void __launch_bounds__(256) diffuse_kernel(const float* __restrict__ x0, const float* __restrict__ eps,
                                                      const int64_t* __restrict__ t, const float* __restrict__ ac,
                                                      float* __restrict__ xt, int B, int CHW) {
  const int b = blockIdx.y;
  const float a = ac[t[b]];
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

extern "C" int b200_conv3x3_first(const float* x, const float* w, const float* bias, float* out, int B, int Cin, int H,
                                  int W, int Cout, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && w && out, "conv3x3_first: null pointer");
  B200_REQUIRE(Cin >= 1 && Cin <= 4, "conv3x3_first: Cin=%d must be in [1,4]", Cin);
  B200_REQUIRE(Cout % 4 == 0 && Cout <= 1024, "conv3x3_first: Cout=%d must be a multiple of 4", Cout);
  const size_t smem = ((size_t)9 * Cin * Cout + Cout) * 4;
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(conv3x3_first_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr = true;
  }
  B200_REQUIRE(smem <= 160 * 1024, "conv3x3_first: weights do not fit in shared memory");
  const int npix = B * H * W;
  int grid = (npix + 63) / 64;  // 8 warps per CTA, ~8 pixels per warp
  if (grid > 148 * 4) grid = 148 * 4;
  conv3x3_first_kernel<<<grid, 256, smem, stream>>>(x, w, bias, out, B, Cin, H, W, Cout);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "conv3x3_first launch");
}

extern "C" int b200_cast_bf16(const float* x, void* out, int B, int H, int W, int C, int parity_split, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out, "cast_bf16: null pointer");
  B200_REQUIRE(C % 4 == 0, "cast_bf16: C=%d must be a multiple of 4", C);
  if (parity_split) B200_REQUIRE(H % 2 == 0 && W % 2 == 0, "cast_bf16: parity split needs even H, W");
  const size_t total = (size_t)B * H * W * (C / 4);
  cast_bf16_kernel<<<ew_grid(total, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out), B, H, W, C,
                                                            parity_split);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "cast_bf16 launch");
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
  B200_REQUIRE(dim % 2 == 0 && rows >= 1, "time_embed: bad dim/rows");
  const size_t smem = (size_t)(dim + E) * 4;
  B200_REQUIRE(smem <= 48 * 1024, "time_embed: dim+E too large");
  time_embed_kernel<<<rows, 256, smem, stream>>>(t, freqs, dim, E, cos_first, w1, b1, w2, b2, y, class_embed, out,
                                                 reinterpret_cast<__nv_bfloat16*>(out_silu_bf16));
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "time_embed launch");
}

extern "C" int b200_diffuse(const float* x0, const float* eps, const int64_t* t, const float* alphas_cumprod, float* xt,
                            int B, int CHW, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x0 && eps && t && alphas_cumprod && xt, "diffuse: null pointer");
  dim3 grid((CHW + 1023) / 1024 > 64 ? 64 : (CHW + 1023) / 1024, B);
  diffuse_kernel<<<grid, 256, 0, stream>>>(x0, eps, t, alphas_cumprod, xt, B, CHW);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "diffuse launch");
}
