// Memory-bound kernels of the backward pass (training step, scripts/train_ddpm.py:171-192): the adjoint of
// GroupNorm(+AdaGN)(+SiLU)(+dropout)(+2x resample) as two streaming passes, gradient casts with fused bias-gradient
// column sums, resampling adjoints, the row softmax of the unfused attention backward and the MSE loss.
// Everything is fp32 arithmetic on NHWC tensors; tensor-core operands of the following GEMMs are emitted as bf16.
#include "common.cuh"
#include <stdlib.h>
#include <string.h>
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

__device__ __forceinline__ float4 bw_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 bw_ldg4_bf16(const __nv_bfloat16* p) {
  const uint2 u = __ldg(reinterpret_cast<const uint2*>(p));
  float4 r;
  r.x = __uint_as_float(u.x << 16); r.y = __uint_as_float(u.x & 0xffff0000u);
  r.z = __uint_as_float(u.y << 16); r.w = __uint_as_float(u.y & 0xffff0000u);
  return r;
}
__device__ __forceinline__ float sigmoid_tanh(float x) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * x));
  return fmaf(0.5f, th, 0.5f);
}
// d/dz [z * sigmoid(z)]
__device__ __forceinline__ float silu_grad(float z) {
  const float s = sigmoid_tanh(z);
  return s * fmaf(z, 1.0f - s, 1.0f);
}

// ================================================================================================
// GroupNorm backward.  Forward (groupnorm.cu): y = resample(drop(act(xh * g' + b'))), xh = (x - mean) * rstd,
// g' = gamma (1 + scale_n), b' = beta (1 + scale_n) + shift_n.  With dz = dL/dz (z = xh g' + b'):
//   S1[n,c] = sum_p dz, S2[n,c] = sum_p dz * xh                                   (pass 1, "reduce")
//   dx = rstd * (g' dz - mean_group(g' S1) - xh * mean_group(g' S2))               (pass 2, "apply")
//   dgamma_c = sum_n S2 (1 + scale), dbeta_c = sum_n S1 (1 + scale), dscale[n,c] = gamma S2 + beta S1, dshift = S1.
// Both passes stream x (fp32) and g (bf16) once: 6 B + 6 B read, 2-4 B written per element.
// ================================================================================================
struct GnBwdParams {
  const __nv_bfloat16* g;
  const float* x0; int C0; const long long* st0;   // [B][C][2] int64 fixed-point statistics (common.cuh: stat_load)
  const float* x1; int C1; const long long* st1;
  int HW, W, groups, cpg, pix_per_cta, reverse;
  const float* gamma; const float* beta; float eps;
  const float* scale; const float* shift; int ss_ld;
  int apply_silu, resample;
  uint32_t drop_thresh; float drop_scale; unsigned long long drop_seed; const unsigned long long* drop_seed_dev;
  float* sums;
  float* dx0; int acc0; float* dx1; int acc1; const float* addend;
  __nv_bfloat16* dx_bf16; float* dx_rowsum; int rowsum_ld; float* dx_colsum;
  float* dgamma; float* dbeta; float* dscale; float* dshift; int dss_ld;
};

// Workspace layout (d->sums, [B][8][C] floats): rows 0,1 = S1, S2 (channel sums of dz and dz*xh); rows 2..5 = the
// forward coefficients z = x*cA + cB, xh = x*rA + rB; rows 6,7 = k2, k3 of dx = dz*cA - k2 - xh*k3.
// The coefficient rows are computed ONCE per image by small kernels, so the two streaming passes start loading data
// immediately (no per-CTA prologue: with it the passes ran at 25 % of HBM bandwidth on the 16x16 / 32x32 slabs).
__device__ __forceinline__ float* gn_ws(const GnBwdParams& p, int n, int row) {
  return p.sums + ((size_t)n * 8 + row) * (p.C0 + p.C1);
}

__global__ void __launch_bounds__(256) gn_bwd_coef_kernel(const GnBwdParams p) {
  extern __shared__ float bsm[];
  const int C = p.C0 + p.C1;
  float* tmpS = bsm; float* tmpQ = bsm + C;
  const int n = blockIdx.x;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const long long* st = (c < p.C0) ? p.st0 + ((size_t)n * p.C0 + c) * 2 : p.st1 + ((size_t)n * p.C1 + (c - p.C0)) * 2;
    const float2 sv = stat_load(st);
    tmpS[c] = sv.x;
    tmpQ[c] = sv.y;
  }
  __syncthreads();
  const float inv_cnt = 1.0f / (float)(p.HW * p.cpg);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / p.cpg) * p.cpg;
    float s = 0.f, q = 0.f;
    for (int i = 0; i < p.cpg; ++i) { s += tmpS[g0 + i]; q += tmpQ[g0 + i]; }
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
    gn_ws(p, n, 0)[c] = 0.f;
    gn_ws(p, n, 1)[c] = 0.f;
    gn_ws(p, n, 2)[c] = rstd * ga;
    gn_ws(p, n, 3)[c] = be - mean * rstd * ga;
    gn_ws(p, n, 4)[c] = rstd;
    gn_ws(p, n, 5)[c] = -mean * rstd;
  }
}

// k2 / k3 from the channel sums, parameter gradients of image n
__global__ void __launch_bounds__(256) gn_bwd_final_kernel(const GnBwdParams p) {
  const int C = p.C0 + p.C1;
  const int n = blockIdx.x;
  const float* S1 = gn_ws(p, n, 0); const float* S2 = gn_ws(p, n, 1);
  const float* cA = gn_ws(p, n, 2); const float* rA = gn_ws(p, n, 4);
  const float inv_m = 1.0f / (float)(p.HW * p.cpg);
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    const int g0 = (c / p.cpg) * p.cpg;
    float A = 0.f, Bq = 0.f;
    for (int i = 0; i < p.cpg; ++i) {
      const float gp = cA[g0 + i] / rA[g0 + i];     // g' = gamma (1 + scale)
      A = fmaf(gp, S1[g0 + i], A);
      Bq = fmaf(gp, S2[g0 + i], Bq);
    }
    gn_ws(p, n, 6)[c] = rA[c] * A * inv_m;
    gn_ws(p, n, 7)[c] = rA[c] * Bq * inv_m;
    const float s1 = S1[c], s2 = S2[c];
    const float sc = p.scale ? 1.f + __ldg(p.scale + (size_t)n * p.ss_ld + c) : 1.f;
    if (p.dgamma) atomicAdd(p.dgamma + c, s2 * sc);
    if (p.dbeta) atomicAdd(p.dbeta + c, s1 * sc);
    if (p.dscale) {
      const float ga = p.gamma ? __ldg(p.gamma + c) : 1.f, be = p.beta ? __ldg(p.beta + c) : 0.f;
      p.dscale[(size_t)n * p.dss_ld + c] = ga * s2 + be * s1;
      p.dshift[(size_t)n * p.dss_ld + c] = s1;
    }
  }
}

// gradient w.r.t. the activated output at input pixel px (adjoint of the forward's resampling), 4 channels
__device__ __forceinline__ float4 gn_bwd_load_g(const GnBwdParams& p, int n, int px, int C, int c) {
  if (p.resample == 0) return bw_ldg4_bf16(p.g + ((size_t)n * p.HW + px) * C + c);
  const int py = px / p.W, pxx = px - py * p.W;
  if (p.resample == 1) {   // forward averaged 2x2 blocks: every input pixel receives a quarter of its block's gradient
    const int Wo = p.W >> 1;
    float4 v = bw_ldg4_bf16(p.g + ((size_t)n * (p.HW >> 2) + (size_t)(py >> 1) * Wo + (pxx >> 1)) * C + c);
    v.x *= 0.25f; v.y *= 0.25f; v.z *= 0.25f; v.w *= 0.25f;
    return v;
  }
  const int Wo = p.W * 2;   // forward replicated the pixel 2x2: sum the four gradients
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      const float4 v = bw_ldg4_bf16(p.g + ((size_t)n * (4 * (size_t)p.HW) + (size_t)(2 * py + dy) * Wo + 2 * pxx + dx) * C + c);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
  return a;
}

__device__ __forceinline__ float4 gn_bwd_load_x(const GnBwdParams& p, int n, int px, int c) {
  return (c < p.C0) ? bw_ldg4(p.x0 + ((size_t)n * p.HW + px) * p.C0 + c)
                    : bw_ldg4(p.x1 + ((size_t)n * p.HW + px) * p.C1 + (c - p.C0));
}

// dz for 4 channels of one pixel; also returns xh
__device__ __forceinline__ void gn_bwd_dz(const GnBwdParams& p, int n, int px, int C, int c, const float4 x, const float4 g,
                                          const float4 a, const float4 b, const float4 ra, const float4 rb, float (&dz)[4],
                                          float (&xh)[4]) {
  const float xv[4] = {x.x, x.y, x.z, x.w}, gv[4] = {g.x, g.y, g.z, g.w};
  const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
  const float rav[4] = {ra.x, ra.y, ra.z, ra.w}, rbv[4] = {rb.x, rb.y, rb.z, rb.w};
  uint32_t keep = 15u;
  if (p.drop_thresh) keep = dropout_keep4(p.drop_seed + (p.drop_seed_dev ? __ldg(p.drop_seed_dev) : 0ull), (((unsigned long long)n * p.HW + px) * C + c) >> 2, p.drop_thresh);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xh[i] = fmaf(xv[i], rav[i], rbv[i]);
    float d = gv[i];
    if (p.apply_silu) d *= silu_grad(fmaf(xv[i], av[i], bv[i]));
    if (p.drop_thresh) d = ((keep >> i) & 1u) ? d * p.drop_scale : 0.f;
    dz[i] = d;
  }
}


// dz / xh of 4 channels from already fetched operands and an already drawn keep mask (the hot loops below step
// pointers and the dropout counter instead of re-deriving 64-bit addresses per pixel)
__device__ __forceinline__ void gn_bwd_dz_core(const float4 x, const float4 g, const float4 a, const float4 b,
                                               const float4 ra, const float4 rb, const uint32_t keep, const bool silu,
                                               const bool drop, const float drop_scale, float (&dz)[4], float (&xh)[4]) {
  const float xv[4] = {x.x, x.y, x.z, x.w}, gv[4] = {g.x, g.y, g.z, g.w};
  const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
  const float rav[4] = {ra.x, ra.y, ra.z, ra.w}, rbv[4] = {rb.x, rb.y, rb.z, rb.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    xh[i] = fmaf(xv[i], rav[i], rbv[i]);
    float d = gv[i];
    if (silu) d *= silu_grad(fmaf(xv[i], av[i], bv[i]));
    if (drop) d = ((keep >> i) & 1u) ? d * drop_scale : 0.f;
    dz[i] = d;
  }
}

// no-resampling columns (every ResBlock GroupNorm except the up/down ones; concatenated inputs included).
// One 32-bit element offset is stepped (the launcher takes this path only for tensors below 2^32 elements); the
// base pointers stay in the constant bank and the dropout counter is offset >> 2.
template <bool kCat>
__device__ __forceinline__ void gn_bwd_reduce_simple(const GnBwdParams& p, const int n, const int c, const int prow,
                                                     const int pstep, const int px0, const int px1, const float4 a,
                                                     const float4 b, const float4 ra, const float4 rb, float (&s1)[4],
                                                     float (&s2)[4]) {
  const int C = p.C0 + (kCat ? p.C1 : 0);
  const bool from0 = !kCat || c < p.C0;
  const float* xsrc = from0 ? p.x0 : p.x1;          // concatenated inputs: the column lives in one of the two sources
  const uint32_t sld = from0 ? p.C0 : p.C1, sc = from0 ? c : c - p.C0;
  const uint32_t pix = (uint32_t)n * (uint32_t)p.HW + (uint32_t)(px0 + prow);
  uint32_t off = pix * (uint32_t)C + (uint32_t)c, xoff = kCat ? pix * sld + sc : 0u;
  const uint32_t step = (uint32_t)pstep * (uint32_t)C, xstep = kCat ? (uint32_t)pstep * sld : 0u;
  const bool drop = p.drop_thresh != 0, silu = p.apply_silu != 0;
  const unsigned long long seed = p.drop_seed + ((drop && p.drop_seed_dev) ? __ldg(p.drop_seed_dev) : 0ull);
  for (int px = px0 + prow; px < px1; px += 4 * pstep, off += 4 * step, xoff += 4 * xstep) {
    float4 xs[4];
    uint2 gs[4];       // bf16 x 4, unpacked when consumed (register budget)
#pragma unroll
    for (int u = 0; u < 4; ++u)
      if (px + u * pstep < px1) {
        xs[u] = kCat ? bw_ldg4(xsrc + (size_t)(xoff + u * xstep)) : bw_ldg4(p.x0 + (size_t)(off + u * step));
        gs[u] = __ldg(reinterpret_cast<const uint2*>(p.g + (size_t)(off + u * step)));
      }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (px + u * pstep >= px1) break;
      const uint32_t keep = drop ? dropout_keep4(seed, (unsigned long long)((off + u * step) >> 2), p.drop_thresh) : 15u;
      float dz[4], xh[4];
      const float4 g4 = make_float4(__uint_as_float(gs[u].x << 16), __uint_as_float(gs[u].x & 0xffff0000u),
                                    __uint_as_float(gs[u].y << 16), __uint_as_float(gs[u].y & 0xffff0000u));
      gn_bwd_dz_core(xs[u], g4, a, b, ra, rb, keep, silu, drop, p.drop_scale, dz, xh);
#pragma unroll
      for (int i = 0; i < 4; ++i) { s1[i] += dz[i]; s2[i] = fmaf(dz[i], xh[i], s2[i]); }
    }
  }
}

template <int kMode>   // 0 general (resampling), 1 single source, 2 concatenated sources
__global__ void __launch_bounds__(256, 4) gn_bwd_reduce_kernel(const GnBwdParams p) {
  extern __shared__ float bsm[];
  const int C = p.C0 + p.C1;
  float* accS = bsm; float* accQ = bsm + C;
  const int n = blockIdx.y, tid = threadIdx.x;
  const float* cA = gn_ws(p, n, 2); const float* cB = gn_ws(p, n, 3);
  const float* rA = gn_ws(p, n, 4); const float* rB = gn_ws(p, n, 5);
  for (int c = tid; c < C; c += 256) { accS[c] = 0.f; accQ[c] = 0.f; }
  __syncthreads();
  const int nv = C >> 2;
  const int cols = nv < 256 ? nv : 256;
  const int pstep = 256 / cols;
  const int px0 = blockIdx.x * p.pix_per_cta;
  const int px1 = min(p.HW, px0 + p.pix_per_cta);
  for (int j0 = 0; j0 < nv; j0 += cols) {
    const int j = j0 + tid % cols, prow = tid / cols;
    if (prow >= pstep || j >= nv) continue;
    const int c = j << 2;
    const float4 a = bw_ldg4(cA + c), b = bw_ldg4(cB + c);
    const float4 ra = bw_ldg4(rA + c), rb = bw_ldg4(rB + c);
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (kMode == 1) gn_bwd_reduce_simple<false>(p, n, c, prow, pstep, px0, px1, a, b, ra, rb, s1, s2);
    else if (kMode == 2) gn_bwd_reduce_simple<true>(p, n, c, prow, pstep, px0, px1, a, b, ra, rb, s1, s2);
    else
    for (int px = px0 + prow; px < px1; px += 4 * pstep) {
      float4 xs[4], gs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (px + u * pstep < px1) {
          xs[u] = gn_bwd_load_x(p, n, px + u * pstep, c);
          gs[u] = gn_bwd_load_g(p, n, px + u * pstep, C, c);
        }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (px + u * pstep >= px1) break;
        float dz[4], xh[4];
        gn_bwd_dz(p, n, px + u * pstep, C, c, xs[u], gs[u], a, b, ra, rb, dz, xh);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s1[i] += dz[i]; s2[i] = fmaf(dz[i], xh[i], s2[i]); }
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (pstep > 1) { atomicAdd(accS + c + i, s1[i]); atomicAdd(accQ + c + i, s2[i]); }
      else { accS[c + i] = s1[i]; accQ[c + i] = s2[i]; }
    }
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) {
    atomicAdd(gn_ws(p, n, 0) + c, accS[c]);
    atomicAdd(gn_ws(p, n, 1) + c, accQ[c]);
  }
}


// pass 2 for single-source, no-resampling columns: dx = dz*cA - k2 - xh*k3 (+ addend), bf16 or fp32 (optionally
// accumulating) output, per-thread row sums of dx; pointers and the dropout counter are stepped, not re-derived
constexpr int kUA = 2;   // pixels in flight per thread in the apply pass (4 spill at 64 registers)
// pass 2 for single-source, no-resampling columns: dx = dz*cA - k2 - xh*k3 (+ addend), bf16 or fp32 (optionally
// accumulating) output, per-thread row sums of dx; same 32-bit offset stepping as the reduce pass
template <bool kCat>
__device__ __forceinline__ void gn_bwd_apply_simple(const GnBwdParams& p, const int n, const int c, const int prow,
                                                    const int pstep, const int px0, const int px1, const float4 a,
                                                    const float4 b, const float4 ra, const float4 rb, const float4 q2,
                                                    const float4 q3, float (&rs)[4]) {
  const int C = p.C0 + (kCat ? p.C1 : 0);
  const bool from0 = !kCat || c < p.C0;
  const float* xsrc = from0 ? p.x0 : p.x1;
  float* dst = from0 ? p.dx0 : p.dx1;
  const uint32_t sld = from0 ? p.C0 : p.C1, sc = from0 ? c : c - p.C0;
  const uint32_t pix = (uint32_t)n * (uint32_t)p.HW + (uint32_t)(px0 + prow);
  uint32_t off = pix * (uint32_t)C + (uint32_t)c, xoff = kCat ? pix * sld + sc : 0u;
  const uint32_t step = (uint32_t)pstep * (uint32_t)C, xstep = kCat ? (uint32_t)pstep * sld : 0u;
  const bool acc = (from0 ? p.acc0 : p.acc1) != 0;
  const bool drop = p.drop_thresh != 0, silu = p.apply_silu != 0;
  const unsigned long long seed = p.drop_seed + ((drop && p.drop_seed_dev) ? __ldg(p.drop_seed_dev) : 0ull);
  const float av[4] = {a.x, a.y, a.z, a.w}, q2v[4] = {q2.x, q2.y, q2.z, q2.w}, q3v[4] = {q3.x, q3.y, q3.z, q3.w};
  for (int px = px0 + prow; px < px1; px += kUA * pstep, off += kUA * step, xoff += kUA * xstep) {
    float4 xs[kUA], ads[kUA];
    uint2 gs[kUA];
#pragma unroll
    for (int u = 0; u < kUA; ++u)
      if (px + u * pstep < px1) {
        xs[u] = kCat ? bw_ldg4(xsrc + (size_t)(xoff + u * xstep)) : bw_ldg4(p.x0 + (size_t)(off + u * step));
        gs[u] = __ldg(reinterpret_cast<const uint2*>(p.g + (size_t)(off + u * step)));
        if (p.addend) ads[u] = bw_ldg4(p.addend + (size_t)(off + u * step));
      }
#pragma unroll
    for (int u = 0; u < kUA; ++u) {
      if (px + u * pstep >= px1) break;
      const size_t o = (size_t)(off + u * step);
      const uint32_t keep = drop ? dropout_keep4(seed, (unsigned long long)((off + u * step) >> 2), p.drop_thresh) : 15u;
      float dz[4], xh[4], dx[4];
      const float4 g4 = make_float4(__uint_as_float(gs[u].x << 16), __uint_as_float(gs[u].x & 0xffff0000u),
                                    __uint_as_float(gs[u].y << 16), __uint_as_float(gs[u].y & 0xffff0000u));
      gn_bwd_dz_core(xs[u], g4, a, b, ra, rb, keep, silu, drop, p.drop_scale, dz, xh);
#pragma unroll
      for (int i = 0; i < 4; ++i) dx[i] = fmaf(dz[i], av[i], -q2v[i]) - xh[i] * q3v[i];
      if (p.addend) { dx[0] += ads[u].x; dx[1] += ads[u].y; dx[2] += ads[u].z; dx[3] += ads[u].w; }
      if (p.dx_bf16) {
        uint2 u2;
        u2.x = pack_bf16x2(dx[0], dx[1]);
        u2.y = pack_bf16x2(dx[2], dx[3]);
        *reinterpret_cast<uint2*>(p.dx_bf16 + o) = u2;
      } else if (kCat ? dst != nullptr : p.dx0 != nullptr) {
        float4* op = reinterpret_cast<float4*>(kCat ? dst + (size_t)(xoff + u * xstep) : p.dx0 + o);
        if (acc) {
          const float4 old = *op;
          dx[0] += old.x; dx[1] += old.y; dx[2] += old.z; dx[3] += old.w;
        }
        *op = make_float4(dx[0], dx[1], dx[2], dx[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) rs[i] += dx[i];
    }
  }
}

template <int kMode>
__global__ void __launch_bounds__(256, 4) gn_bwd_apply_kernel(const GnBwdParams p) {
  extern __shared__ float bsm[];
  const int C = p.C0 + p.C1;
  float* rsum = bsm;
  // second streaming pass: walk images / pixel ranges from the END -- the reduce pass just read x and g front to back, so
  // the tail of the tensors is what the 126 MB L2 still holds (p.reverse, B200_GNB_REVERSE=0: front to back)
  const int n = p.reverse ? (int)(gridDim.y - 1 - blockIdx.y) : (int)blockIdx.y, tid = threadIdx.x;
  const float* cA = gn_ws(p, n, 2); const float* cB = gn_ws(p, n, 3);
  const float* rA = gn_ws(p, n, 4); const float* rB = gn_ws(p, n, 5);
  const float* k2 = gn_ws(p, n, 6); const float* k3 = gn_ws(p, n, 7);
  if (p.dx_rowsum || p.dx_colsum) {
    for (int c = tid; c < C; c += 256) rsum[c] = 0.f;
    __syncthreads();
  }
  const int nv = C >> 2;
  const int cols = nv < 256 ? nv : 256;
  const int pstep = 256 / cols;
  const int px0 = (p.reverse ? (int)(gridDim.x - 1 - blockIdx.x) : (int)blockIdx.x) * p.pix_per_cta;
  const int px1 = min(p.HW, px0 + p.pix_per_cta);
  for (int j0 = 0; j0 < nv; j0 += cols) {
    const int j = j0 + tid % cols, prow = tid / cols;
    if (prow >= pstep || j >= nv) continue;
    const int c = j << 2;
    const float4 a = bw_ldg4(cA + c), b = bw_ldg4(cB + c);
    const float4 ra = bw_ldg4(rA + c), rb = bw_ldg4(rB + c);
    const float4 q2 = bw_ldg4(k2 + c), q3 = bw_ldg4(k3 + c);
    const float av[4] = {a.x, a.y, a.z, a.w}, q2v[4] = {q2.x, q2.y, q2.z, q2.w}, q3v[4] = {q3.x, q3.y, q3.z, q3.w};
    const bool from0 = c < p.C0;
    float* dst = from0 ? p.dx0 : p.dx1;
    const int dld = from0 ? p.C0 : p.C1, dc = from0 ? c : c - p.C0;
    const int acc = from0 ? p.acc0 : p.acc1;
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    if (kMode == 1) gn_bwd_apply_simple<false>(p, n, c, prow, pstep, px0, px1, a, b, ra, rb, q2, q3, rs);
    else if (kMode == 2) gn_bwd_apply_simple<true>(p, n, c, prow, pstep, px0, px1, a, b, ra, rb, q2, q3, rs);
    else
    for (int pxb = px0 + prow; pxb < px1; pxb += 2 * pstep) {
      float4 xs[2], gs[2], ads[2];
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (pxb + u * pstep < px1) {
          xs[u] = gn_bwd_load_x(p, n, pxb + u * pstep, c);
          gs[u] = gn_bwd_load_g(p, n, pxb + u * pstep, C, c);
          if (p.addend) ads[u] = bw_ldg4(p.addend + ((size_t)n * p.HW + pxb + u * pstep) * C + c);
        }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int px = pxb + u * pstep;
        if (px >= px1) break;
        const float4 x = xs[u], g = gs[u];
      float dz[4], xh[4], dx[4];
      gn_bwd_dz(p, n, px, C, c, x, g, a, b, ra, rb, dz, xh);
#pragma unroll
      for (int i = 0; i < 4; ++i) dx[i] = fmaf(dz[i], av[i], -q2v[i]) - xh[i] * q3v[i];
      if (p.addend) {
        const float4 ad = ads[u];
        dx[0] += ad.x; dx[1] += ad.y; dx[2] += ad.z; dx[3] += ad.w;
      }
      if (p.dx_bf16) {
        uint2 u2;
        u2.x = pack_bf16x2(dx[0], dx[1]);
        u2.y = pack_bf16x2(dx[2], dx[3]);
        *reinterpret_cast<uint2*>(p.dx_bf16 + ((size_t)n * p.HW + px) * C + c) = u2;
      } else if (dst) {
        float4* o = reinterpret_cast<float4*>(dst + ((size_t)n * p.HW + px) * dld + dc);
        if (acc) {
          const float4 old = *o;
          dx[0] += old.x; dx[1] += old.y; dx[2] += old.z; dx[3] += old.w;
        }
        *o = make_float4(dx[0], dx[1], dx[2], dx[3]);
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) rs[i] += dx[i];
      }
    }
    if (p.dx_rowsum || p.dx_colsum) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(rsum + c + i, rs[i]);
    }
  }
  if (p.dx_rowsum || p.dx_colsum) {
    __syncthreads();
    for (int c = tid; c < C; c += 256) {
      if (p.dx_rowsum) atomicAdd(p.dx_rowsum + (size_t)n * p.rowsum_ld + c, rsum[c]);
      if (p.dx_colsum) atomicAdd(p.dx_colsum + c, rsum[c]);
    }
  }
}

// ================================================================================================
// One-launch GroupNorm adjoint for small images (HW <= 256 by default, no resampling): the four-kernel form above is bound by its
// launch chain there (ncu, CFG UNet at batch 128: coef 4 us -> reduce 14-19 us -> final 5 us -> apply 19-30 us for 3-25 MB
// of traffic, 128 CTAs at 12 % occupancy on the 8x8 / 4x4 levels).  A CTA owns ALL pixels of one image for a slice of
// CS channels made of whole groups, so nothing has to be exchanged between CTAs: x (fp32) and g (bf16) are read from
// HBM ONCE into shared memory, the forward coefficients come straight from the producer's statistics, S1 / S2 are
// reduced inside the CTA in a fixed order (no atomics), k2 / k3 and the parameter gradients follow, and the second pass
// runs from shared memory.  Thread (q, prow) owns channel quad q of the slice and pixels prow, prow + R, ... in both
// passes, so the data needs no barrier of its own.  Same formulas, dropout counters and outputs as the kernels above.
// ================================================================================================
__global__ void __launch_bounds__(256, 3) gn_bwd_slab_kernel(const GnBwdParams p, const int CS) {
  extern __shared__ __align__(16) uint8_t slab_sm[];
  const int C = p.C0 + p.C1, HW = p.HW;
  const int n = blockIdx.y, c0 = blockIdx.x * CS;
  const int nq = CS >> 2;            // channel quads of the slice
  const int R = 256 / nq;            // pixel rows processed in parallel
  float* xs = reinterpret_cast<float*>(slab_sm);                              // [HW][CS]
  __nv_bfloat16* gs = reinterpret_cast<__nv_bfloat16*>(xs + (size_t)HW * CS);   // [HW][CS]
  float* cf = reinterpret_cast<float*>(gs + (size_t)HW * CS);                   // [8][CS]: cA cB rA rB k2 k3 S1 S2
  float* part = cf + 8 * CS;                                                  // [2][R][CS] per-row partial sums
  float* cA = cf, *cB = cf + CS, *rA = cf + 2 * CS, *rB = cf + 3 * CS, *k2 = cf + 4 * CS, *k3 = cf + 5 * CS;
  float* S1 = cf + 6 * CS, *S2 = cf + 7 * CS;
  const int tid = threadIdx.x;
  const int q = tid % nq, prow = tid / nq;
  const bool active = prow < R;
  const int c = c0 + (q << 2);                 // first channel of this thread's quad (concatenated index)
  const bool from0 = c < p.C0;
  const float* xsrc = from0 ? p.x0 + (size_t)n * HW * p.C0 + c : p.x1 + (size_t)n * HW * p.C1 + (c - p.C0);
  const int sld = from0 ? p.C0 : p.C1;
  const size_t goff = (size_t)n * HW * C + c;  // offset of (n, pixel 0, c) in the concatenated tensors (g, addend, dx_bf16)

  // ---- 1. x, g -> shared memory (all loads of a thread are issued back to back) ----
  if (active) {
    for (int px = prow; px < HW; px += 4 * R) {
      float4 xv[4];
      uint2 gv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (px + u * R < HW) {
          xv[u] = bw_ldg4(xsrc + (size_t)(px + u * R) * sld);
          gv[u] = __ldg(reinterpret_cast<const uint2*>(p.g + goff + (size_t)(px + u * R) * C));
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (px + u * R < HW) {
          *reinterpret_cast<float4*>(xs + (size_t)(px + u * R) * CS + (q << 2)) = xv[u];
          *reinterpret_cast<uint2*>(gs + (size_t)(px + u * R) * CS + (q << 2)) = gv[u];
        }
    }
  }
  // ---- 2. forward coefficients of the slice's channels from the producer statistics ----
  if (tid < CS) {
    const int cc = c0 + tid;
    const long long* st = (cc < p.C0) ? p.st0 + ((size_t)n * p.C0 + cc) * 2 : p.st1 + ((size_t)n * p.C1 + (cc - p.C0)) * 2;
    const float2 sv = stat_load(st);
    S1[tid] = sv.x;
    S2[tid] = sv.y;
  }
  __syncthreads();
  if (tid < CS) {
    const int cc = c0 + tid;
    const int g0 = (tid / p.cpg) * p.cpg;
    float sm = 0.f, qm = 0.f;
    for (int i = 0; i < p.cpg; ++i) { sm += S1[g0 + i]; qm += S2[g0 + i]; }
    const float inv_cnt = 1.0f / (float)(HW * p.cpg);
    const float mean = sm * inv_cnt;
    const float var = fmaxf(qm * inv_cnt - mean * mean, 0.f);
    const float rstd = rsqrtf(var + p.eps);
    float ga = p.gamma ? __ldg(p.gamma + cc) : 1.f;
    float be = p.beta ? __ldg(p.beta + cc) : 0.f;
    if (p.scale) {
      const float sc = 1.f + __ldg(p.scale + (size_t)n * p.ss_ld + cc);
      ga *= sc;
      be = be * sc + __ldg(p.shift + (size_t)n * p.ss_ld + cc);
    }
    cA[tid] = rstd * ga;
    cB[tid] = be - mean * rstd * ga;
    rA[tid] = rstd;
    rB[tid] = -mean * rstd;
  }
  __syncthreads();

  const bool drop = p.drop_thresh != 0, silu = p.apply_silu != 0;
  const unsigned long long seed = p.drop_seed + ((drop && p.drop_seed_dev) ? __ldg(p.drop_seed_dev) : 0ull);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a, ra = a, rb = a;
  if (active) {
    a = *reinterpret_cast<const float4*>(cA + (q << 2)); b = *reinterpret_cast<const float4*>(cB + (q << 2));
    ra = *reinterpret_cast<const float4*>(rA + (q << 2)); rb = *reinterpret_cast<const float4*>(rB + (q << 2));
  }
  // ---- 3. pass 1: S1 = sum dz, S2 = sum dz * xh over the image ----
  {
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (active)
      for (int px = prow; px < HW; px += R) {
        const float4 x = *reinterpret_cast<const float4*>(xs + (size_t)px * CS + (q << 2));
        const uint2 gu = *reinterpret_cast<const uint2*>(gs + (size_t)px * CS + (q << 2));
        const float4 g4 = make_float4(__uint_as_float(gu.x << 16), __uint_as_float(gu.x & 0xffff0000u),
                                      __uint_as_float(gu.y << 16), __uint_as_float(gu.y & 0xffff0000u));
        const uint32_t keep = drop ? dropout_keep4(seed, (unsigned long long)((goff + (size_t)px * C) >> 2), p.drop_thresh) : 15u;
        float dz[4], xh[4];
        gn_bwd_dz_core(x, g4, a, b, ra, rb, keep, silu, drop, p.drop_scale, dz, xh);
#pragma unroll
        for (int i = 0; i < 4; ++i) { s1[i] += dz[i]; s2[i] = fmaf(dz[i], xh[i], s2[i]); }
        // pass 2 needs dz and xh again: keep them in this thread's own slab slots (dz fp32 over x, xh bf16 over g; xh only
        // multiplies k3, a mean over the group, so its bf16 rounding is ~1e-4 of dx) instead of re-evaluating the SiLU
        // derivative and the dropout hash
        *reinterpret_cast<float4*>(xs + (size_t)px * CS + (q << 2)) = make_float4(dz[0], dz[1], dz[2], dz[3]);
        uint2 xu;
        xu.x = pack_bf16x2(xh[0], xh[1]);
        xu.y = pack_bf16x2(xh[2], xh[3]);
        *reinterpret_cast<uint2*>(gs + (size_t)px * CS + (q << 2)) = xu;
      }
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        part[(size_t)prow * CS + (q << 2) + i] = s1[i];
        part[(size_t)(R + prow) * CS + (q << 2) + i] = s2[i];
      }
    }
  }
  __syncthreads();
  if (tid < CS) {
    float u = 0.f, v = 0.f;
    for (int r = 0; r < R; ++r) { u += part[(size_t)r * CS + tid]; v += part[(size_t)(R + r) * CS + tid]; }
    S1[tid] = u;
    S2[tid] = v;
  }
  __syncthreads();
  // ---- 4. k2 / k3 of the groups, parameter gradients ----
  if (tid < CS) {
    const int cc = c0 + tid;
    const int g0 = (tid / p.cpg) * p.cpg;
    float A = 0.f, Bq = 0.f;
    for (int i = 0; i < p.cpg; ++i) {
      const float gp = cA[g0 + i] / rA[g0 + i];     // g' = gamma (1 + scale)
      A = fmaf(gp, S1[g0 + i], A);
      Bq = fmaf(gp, S2[g0 + i], Bq);
    }
    const float inv_m = 1.0f / (float)(HW * p.cpg);
    k2[tid] = rA[tid] * A * inv_m;
    k3[tid] = rA[tid] * Bq * inv_m;
    const float s1 = S1[tid], s2 = S2[tid];
    const float sc = p.scale ? 1.f + __ldg(p.scale + (size_t)n * p.ss_ld + cc) : 1.f;
    if (p.dgamma) atomicAdd(p.dgamma + cc, s2 * sc);
    if (p.dbeta) atomicAdd(p.dbeta + cc, s1 * sc);
    if (p.dscale) {
      const float ga = p.gamma ? __ldg(p.gamma + cc) : 1.f, be = p.beta ? __ldg(p.beta + cc) : 0.f;
      p.dscale[(size_t)n * p.dss_ld + cc] = ga * s2 + be * s1;
      p.dshift[(size_t)n * p.dss_ld + cc] = s1;
    }
  }
  __syncthreads();
  // ---- 5. pass 2: dx = dz*cA - k2 - xh*k3 (+ addend) ----
  float rs[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
    const float4 q2 = *reinterpret_cast<const float4*>(k2 + (q << 2)), q3 = *reinterpret_cast<const float4*>(k3 + (q << 2));
    const float av[4] = {a.x, a.y, a.z, a.w}, q2v[4] = {q2.x, q2.y, q2.z, q2.w}, q3v[4] = {q3.x, q3.y, q3.z, q3.w};
    float* dst = from0 ? p.dx0 : p.dx1;
    float* dbase = dst ? (from0 ? dst + (size_t)n * HW * p.C0 + c : dst + (size_t)n * HW * p.C1 + (c - p.C0)) : nullptr;
    const bool acc = (from0 ? p.acc0 : p.acc1) != 0;
    for (int px0 = prow; px0 < HW; px0 += 2 * R) {
      float4 ads[2], olds[2];
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int px = px0 + u * R;
        if (px < HW) {
          if (p.addend) ads[u] = bw_ldg4(p.addend + goff + (size_t)px * C);
          if (!p.dx_bf16 && dbase && acc) olds[u] = *reinterpret_cast<const float4*>(dbase + (size_t)px * sld);
        }
      }
#pragma unroll
      for (int u = 0; u < 2; ++u) {
        const int px = px0 + u * R;
        if (px >= HW) break;
        const float4 dz4 = *reinterpret_cast<const float4*>(xs + (size_t)px * CS + (q << 2));      // dz, xh of pass 1
        const uint2 xu = *reinterpret_cast<const uint2*>(gs + (size_t)px * CS + (q << 2));
        const float dz[4] = {dz4.x, dz4.y, dz4.z, dz4.w};
        const float xh[4] = {__uint_as_float(xu.x << 16), __uint_as_float(xu.x & 0xffff0000u),
                             __uint_as_float(xu.y << 16), __uint_as_float(xu.y & 0xffff0000u)};
        float dx[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) dx[i] = fmaf(dz[i], av[i], -q2v[i]) - xh[i] * q3v[i];
        if (p.addend) { dx[0] += ads[u].x; dx[1] += ads[u].y; dx[2] += ads[u].z; dx[3] += ads[u].w; }
        if (p.dx_bf16) {
          uint2 u2;
          u2.x = pack_bf16x2(dx[0], dx[1]);
          u2.y = pack_bf16x2(dx[2], dx[3]);
          *reinterpret_cast<uint2*>(p.dx_bf16 + goff + (size_t)px * C) = u2;
        } else if (dbase) {
          if (acc) { dx[0] += olds[u].x; dx[1] += olds[u].y; dx[2] += olds[u].z; dx[3] += olds[u].w; }
          *reinterpret_cast<float4*>(dbase + (size_t)px * sld) = make_float4(dx[0], dx[1], dx[2], dx[3]);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) rs[i] += dx[i];
      }
    }
  }
  if (p.dx_rowsum || p.dx_colsum) {
    if (active) {
#pragma unroll
      for (int i = 0; i < 4; ++i) part[(size_t)prow * CS + (q << 2) + i] = rs[i];
    }
    __syncthreads();
    if (tid < CS) {
      float u = 0.f;
      for (int r = 0; r < R; ++r) u += part[(size_t)r * CS + tid];
      if (p.dx_rowsum) atomicAdd(p.dx_rowsum + (size_t)n * p.rowsum_ld + c0 + tid, u);
      if (p.dx_colsum) atomicAdd(p.dx_colsum + c0 + tid, u);
    }
  }
}

// ================================================================================================
// Gradient casts with fused column sums (bias gradients)
// ================================================================================================
// fp32 [rows][C] -> bf16 [rows][C]; colsum[c] += sum over rows (optional)
__global__ void __launch_bounds__(256) cast_colsum_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                          float* __restrict__ colsum, long long rows, int C,
                                                          int rows_per_cta) {
  extern __shared__ float csm[];
  const int tid = threadIdx.x;
  const int nv = C >> 2;
  const int cols = nv < 256 ? nv : 256;
  const int pstep = 256 / cols;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  if (colsum) {
    for (int c = tid; c < C; c += 256) csm[c] = 0.f;
    __syncthreads();
  }
  for (int j0 = 0; j0 < nv; j0 += cols) {
    const int j = j0 + tid % cols, prow = tid / cols;
    if (prow >= pstep || j >= nv) continue;
    const int c = j << 2;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (long long r = r0 + prow; r < r1; r += pstep) {
      const float4 v = bw_ldg4(x + r * C + c);
      uint2 u;
      u.x = pack_bf16x2(v.x, v.y);
      u.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(out + r * C + c) = u;
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
    if (colsum) {
#pragma unroll
      for (int i = 0; i < 4; ++i) atomicAdd(csm + c + i, s[i]);
    }
  }
  if (colsum) {
    __syncthreads();
    for (int c = tid; c < C; c += 256) atomicAdd(colsum + c, csm[c]);
  }
}

// fp32 NCHW [B][C][HW] (C small) -> bf16 NHWC [B][HW][Cpad] zero padded; colsum[c] += sum (optional)
__global__ void __launch_bounds__(256) nchw_pad_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                       float* __restrict__ colsum, int B, int C, int HW, int Cpad) {
  __shared__ float ssum[8];
  if (threadIdx.x < 8) ssum[threadIdx.x] = 0.f;
  __syncthreads();
  const long long total = (long long)B * HW;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i / HW), px = (int)(i - (long long)n * HW);
    float v[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      v[c] = c < C ? __ldg(x + ((size_t)n * C + c) * HW + px) : 0.f;
      s[c] += v[c];
    }
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    uint4* o = reinterpret_cast<uint4*>(out + (size_t)i * Cpad);
    o[0] = u;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (int k = 1; k < Cpad / 8; ++k) o[k] = z;
  }
  if (colsum) {
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      float t = s[c];
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      if ((threadIdx.x & 31) == 0 && c < C) atomicAdd(ssum + c, t);
    }
    __syncthreads();
    if (threadIdx.x < C) atomicAdd(colsum + threadIdx.x, ssum[threadIdx.x]);
  }
}

// colsum[c] += sum over rows of a bf16 matrix window [rows][ld] at columns [c0, c0 + C)
__global__ void __launch_bounds__(256) colsum_bf16_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ colsum,
                                                          long long rows, int ld, int c0, int C, int rows_per_cta) {
  extern __shared__ float csm[];
  const int tid = threadIdx.x;
  for (int c = tid; c < C; c += 256) csm[c] = 0.f;
  __syncthreads();
  const int nv = C >> 2;
  const int cols = nv < 256 ? nv : 256;
  const int pstep = 256 / cols;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  for (int j0 = 0; j0 < nv; j0 += cols) {
    const int j = j0 + tid % cols, prow = tid / cols;
    if (prow >= pstep || j >= nv) continue;
    const int c = j << 2;
    float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll 4
    for (long long r = r0 + prow; r < r1; r += pstep) {
      const float4 v = bw_ldg4_bf16(x + r * ld + c0 + c);
      s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) atomicAdd(csm + c + i, s[i]);
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) atomicAdd(colsum + c, csm[c]);
}

// 8-column variant (16-byte loads, 8 independent rows in flight per thread): the 4-column form kept only 32 bytes
// per thread in flight and ran at ~1 TB/s
__global__ void __launch_bounds__(256) colsum_bf16x8_kernel(const __nv_bfloat16* __restrict__ x, float* __restrict__ colsum,
                                                            long long rows, int ld, int c0, int C, int rows_per_cta) {
  extern __shared__ float csm[];
  const int tid = threadIdx.x;
  for (int c = tid; c < C; c += 256) csm[c] = 0.f;
  __syncthreads();
  const int nv = C >> 3;
  const int cols = nv < 256 ? nv : 256;
  const int pstep = 256 / cols;
  const long long r0 = (long long)blockIdx.x * rows_per_cta;
  const long long r1 = min(rows, r0 + rows_per_cta);
  for (int j0 = 0; j0 < nv; j0 += cols) {
    const int j = j0 + tid % cols, prow = tid / cols;
    if (prow >= pstep || j >= nv) continue;
    const int c = j << 3;
    float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const __nv_bfloat16* xp = x + c0 + c;
    for (long long r = r0 + prow; r < r1; r += 8ll * pstep) {
      uint4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u)
        if (r + (long long)u * pstep < r1) v[u] = __ldg(reinterpret_cast<const uint4*>(xp + (r + (long long)u * pstep) * ld));
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        if (r + (long long)u * pstep >= r1) break;
        const uint32_t w[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          s[2 * i] += __uint_as_float(w[i] << 16);
          s[2 * i + 1] += __uint_as_float(w[i] & 0xffff0000u);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) atomicAdd(csm + c + i, s[i]);
  }
  __syncthreads();
  for (int c = tid; c < C; c += 256) atomicAdd(colsum + c, csm[c]);
}

// out (+)= scale * resample(x): mode 1 = 2x2 average (out is half size), mode 2 = nearest 2x (out is double size)
__global__ void __launch_bounds__(256) resample_f32_kernel(const float* __restrict__ x, float* __restrict__ out, int B,
                                                           int H, int W, int C, int mode, float scale, int accumulate) {
  const int cv = C >> 2;
  const int Ho = mode == 1 ? H >> 1 : H * 2, Wo = mode == 1 ? W >> 1 : W * 2;
  const size_t total = (size_t)B * Ho * Wo * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cv);
    const size_t po = i / cv;
    const int ox = (int)(po % Wo);
    const int oy = (int)((po / Wo) % Ho);
    const int n = (int)(po / ((size_t)Wo * Ho));
    float4 r;
    if (mode == 1) {
      const float4* src = reinterpret_cast<const float4*>(x) + (((size_t)n * H + 2 * oy) * W + 2 * ox) * cv + c4;
      const float4 a = __ldg(src), b = __ldg(src + cv), c = __ldg(src + (size_t)W * cv), d = __ldg(src + (size_t)W * cv + cv);
      r.x = 0.25f * (a.x + b.x + c.x + d.x); r.y = 0.25f * (a.y + b.y + c.y + d.y);
      r.z = 0.25f * (a.z + b.z + c.z + d.z); r.w = 0.25f * (a.w + b.w + c.w + d.w);
    } else {
      r = __ldg(reinterpret_cast<const float4*>(x) + (((size_t)n * H + (oy >> 1)) * W + (ox >> 1)) * cv + c4);
    }
    r.x *= scale; r.y *= scale; r.z *= scale; r.w *= scale;
    float4* o = reinterpret_cast<float4*>(out) + i;
    if (accumulate) {
      const float4 old = *o;
      r.x += old.x; r.y += old.y; r.z += old.z; r.w += old.w;
    }
    *o = r;
  }
}

// fp32 NHWC -> bf16 NHWC nearest 2x (training-mode Upsample: the 3x3 conv then runs on the materialised tensor)
__global__ void __launch_bounds__(256) upsample2_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ out,
                                                             int B, int H, int W, int C) {
  const int cv = C >> 2, Ho = H * 2, Wo = W * 2;
  const size_t total = (size_t)B * Ho * Wo * cv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c4 = (int)(i % cv);
    const size_t po = i / cv;
    const int ox = (int)(po % Wo);
    const int oy = (int)((po / Wo) % Ho);
    const int n = (int)(po / ((size_t)Wo * Ho));
    const float4 v = __ldg(reinterpret_cast<const float4*>(x) + (((size_t)n * H + (oy >> 1)) * W + (ox >> 1)) * cv + c4);
    uint2 u;
    u.x = pack_bf16x2(v.x, v.y);
    u.y = pack_bf16x2(v.z, v.w);
    reinterpret_cast<uint2*>(out)[i] = u;
  }
}

// ================================================================================================
// Row softmax of the unfused attention backward: P = softmax(scale * S) (bf16), dS = scale * P o (dP - sum_j dP_j P_j)
// One warp per row.
// ================================================================================================
__global__ void __launch_bounds__(256) softmax_rows_kernel(const float* __restrict__ S, __nv_bfloat16* __restrict__ P,
                                                           long long rows, int T, float scale) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* s = S + r * T;
  float m = -INFINITY;
  for (int j = lane; j < T; j += 32) m = fmaxf(m, s[j]);
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  float sum = 0.f;
  for (int j = lane; j < T; j += 32) sum += __expf((s[j] - m) * scale);
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  const float inv = 1.0f / sum;
  for (int j = lane; j < T; j += 32) P[r * T + j] = __float2bfloat16_rn(__expf((s[j] - m) * scale) * inv);
}

__global__ void __launch_bounds__(256) softmax_bwd_rows_kernel(const __nv_bfloat16* __restrict__ P,
                                                               const float* __restrict__ dP,
                                                               __nv_bfloat16* __restrict__ dS, long long rows, int T,
                                                               float scale) {
  const long long r = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float dot = 0.f;
  for (int j = lane; j < T; j += 32) dot = fmaf(__bfloat162float(P[r * T + j]), dP[r * T + j], dot);
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  for (int j = lane; j < T; j += 32)
    dS[r * T + j] = __float2bfloat16_rn(scale * __bfloat162float(P[r * T + j]) * (dP[r * T + j] - dot));
}

// ================================================================================================
// MSE loss (F.mse_loss(mean), diffusions/ddpm.py:136-138) and its gradient
// ================================================================================================
__global__ void __launch_bounds__(256) mse_loss_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       float* __restrict__ loss, long long n, float inv_n) {
  float s = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float d = a[i] - b[i];
    s = fmaf(d, d, s);
  }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  __shared__ float ws[8];
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += ws[w];
    atomicAdd(loss, t * inv_n);
  }
}

__global__ void __launch_bounds__(256) mse_grad_kernel(const float* __restrict__ a, const float* __restrict__ b,
                                                       const float* __restrict__ gscale, float* __restrict__ da,
                                                       long long n, float two_inv_n) {
  const float k = two_inv_n * (gscale ? gscale[0] : 1.f);
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    da[i] = k * (a[i] - b[i]);
}

// keep-mask of the counter-based dropout (for the parity tests: the oracle applies exactly this mask)
__global__ void __launch_bounds__(256) dropout_mask_kernel(float* __restrict__ out, long long n, unsigned long long seed,
                                                           uint32_t thresh) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    out[i] = dropout_keep(seed, (unsigned long long)i, thresh) ? 1.f : 0.f;
}

// ================================================================================================
// Time-embedding MLP backward, per-row part (adjoint of time_embed_kernel in misc.cu): recomputes the sinusoid and the
// hidden layer, and emits the bf16 operands of the two weight-gradient GEMMs:
//   demb = d_semb o silu'(emb);  dhid = demb W2;  dpre = dhid o silu'(hid_pre)
//   db2 += demb, db1 += dpre, dclass[y] += demb (atomics);  pe / hid / demb / dpre rows as bf16.
// ================================================================================================
__global__ void __launch_bounds__(256) time_embed_bwd_kernel(const int64_t* __restrict__ t, const float* __restrict__ freqs,
                                                             int dim, int E, int cos_first, const float* __restrict__ w1,
                                                             const float* __restrict__ b1, const float* __restrict__ w2,
                                                             const float* __restrict__ emb, const float* __restrict__ d_semb,
                                                             const int64_t* __restrict__ y, __nv_bfloat16* __restrict__ pe_o,
                                                             __nv_bfloat16* __restrict__ hid_o, __nv_bfloat16* __restrict__ demb_o,
                                                             __nv_bfloat16* __restrict__ dpre_o, float* __restrict__ db1,
                                                             float* __restrict__ db2, float* __restrict__ dclass) {
  extern __shared__ float tsm[];   // pe[dim] + hid_pre[E] + demb[E]
  float* pe = tsm;
  float* hpre = tsm + dim;
  float* demb = hpre + E;
  const int r = blockIdx.x;
  const int half = dim >> 1;
  const float tv = (float)t[r];
  for (int i = threadIdx.x; i < half; i += blockDim.x) {
    const float a = tv * freqs[i];
    const float s = sinf(a), c = cosf(a);
    pe[i] = cos_first ? c : s;
    pe[half + i] = cos_first ? s : c;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < dim; i += blockDim.x) pe_o[(size_t)r * dim + i] = __float2bfloat16_rn(pe[i]);
  for (int e = threadIdx.x; e < E; e += blockDim.x) {
    const float4* wr = reinterpret_cast<const float4*>(w1 + (size_t)e * dim);
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    for (int k = 0; k < dim / 4; ++k) {
      const float4 wv = __ldg(wr + k);
      a0 += wv.x * pe[4 * k]; a1 += wv.y * pe[4 * k + 1]; a2 += wv.z * pe[4 * k + 2]; a3 += wv.w * pe[4 * k + 3];
    }
    const float v = (a0 + a1) + (a2 + a3) + b1[e];
    hpre[e] = v;
    hid_o[(size_t)r * E + e] = __float2bfloat16_rn(v / (1.0f + expf(-v)));
    const float em = emb[(size_t)r * E + e];
    const float sg = 1.0f / (1.0f + expf(-em));
    const float de = d_semb[(size_t)r * E + e] * sg * (1.0f + em * (1.0f - sg));
    demb[e] = de;
    demb_o[(size_t)r * E + e] = __float2bfloat16_rn(de);
    atomicAdd(db2 + e, de);
    if (y && dclass) atomicAdd(dclass + (size_t)y[r] * E + e, de);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < E; k += blockDim.x) {   // dhid[k] = sum_e demb[e] w2[e][k]  (coalesced over k)
    float acc = 0.f;
    for (int e = 0; e < E; ++e) acc = fmaf(demb[e], __ldg(w2 + (size_t)e * E + k), acc);
    const float v = hpre[k];
    const float sg = 1.0f / (1.0f + expf(-v));
    const float dp = acc * sg * (1.0f + v * (1.0f - sg));
    dpre_o[(size_t)r * E + k] = __float2bfloat16_rn(dp);
    atomicAdd(db1 + k, dp);
  }
}

static inline int bw_grid(long long total, int block) {
  long long g = (total + block - 1) / block;
  const long long cap = 148 * 16;
  return (int)(g < cap ? (g ? g : 1) : cap);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_groupnorm_bwd(const b200_gn_bwd_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d && d->g && d->x0 && d->stats0 && d->sums, "groupnorm_bwd: null g/x0/stats0/sums");
  const int C1 = d->x1 ? d->C1 : 0;
  const int C = d->C0 + C1;
  B200_REQUIRE(d->x1 == nullptr || d->stats1 != nullptr, "groupnorm_bwd: second source needs its statistics");
  B200_REQUIRE(d->groups > 0 && C % d->groups == 0, "groupnorm_bwd: C=%d not divisible by groups=%d", C, d->groups);
  B200_REQUIRE(d->C0 % 4 == 0 && C1 % 4 == 0, "groupnorm_bwd: channel counts must be multiples of 4");
  B200_REQUIRE(C <= 2048, "groupnorm_bwd: C=%d too large", C);
  B200_REQUIRE(d->resample >= 0 && d->resample <= 2, "groupnorm_bwd: bad resample mode");
  B200_REQUIRE(d->W > 0 && d->HW % d->W == 0, "groupnorm_bwd: HW not a multiple of W");
  if (d->resample == 1) B200_REQUIRE(d->W % 2 == 0 && (d->HW / d->W) % 2 == 0, "groupnorm_bwd: avg-pool needs even H, W");
  B200_REQUIRE((d->scale == nullptr) == (d->shift == nullptr), "groupnorm_bwd: scale and shift go together");
  B200_REQUIRE((d->dscale == nullptr) == (d->dshift == nullptr), "groupnorm_bwd: dscale and dshift go together");
  B200_REQUIRE(d->drop_p >= 0.f && d->drop_p < 1.f && (d->drop_p == 0.f || d->resample == 0), "groupnorm_bwd: bad dropout");
  B200_REQUIRE(d->dx_bf16 == nullptr || (d->dx0 == nullptr && d->dx1 == nullptr), "groupnorm_bwd: choose bf16 or fp32 dx");
  GnBwdParams p;
  memset(&p, 0, sizeof(p));
  p.g = reinterpret_cast<const __nv_bfloat16*>(d->g);
  p.x0 = d->x0; p.C0 = d->C0; p.st0 = d->stats0; p.x1 = d->x1; p.C1 = C1; p.st1 = d->stats1;
  p.HW = d->HW; p.W = d->W; p.groups = d->groups; p.cpg = C / d->groups;
  p.gamma = d->gamma; p.beta = d->beta; p.eps = d->eps; p.scale = d->scale; p.shift = d->shift; p.ss_ld = d->ss_ld;
  p.apply_silu = d->apply_silu; p.resample = d->resample;
  p.drop_thresh = dropout_threshold(d->drop_p); p.drop_scale = 1.0f / (1.0f - d->drop_p); p.drop_seed = d->drop_seed;
  p.drop_seed_dev = d->drop_seed_dev;
  p.sums = d->sums;
  p.dx0 = d->dx0; p.acc0 = d->dx0_accumulate; p.dx1 = d->dx1; p.acc1 = d->dx1_accumulate; p.addend = d->addend;
  p.dx_bf16 = reinterpret_cast<__nv_bfloat16*>(d->dx_bf16); p.dx_rowsum = d->dx_rowsum;
  p.rowsum_ld = d->dx_rowsum_ld ? d->dx_rowsum_ld : C; p.dx_colsum = d->dx_colsum;
  p.dgamma = d->dgamma; p.dbeta = d->dbeta; p.dscale = d->dscale; p.dshift = d->dshift; p.dss_ld = d->dss_ld;
  static const char* env_ppc = getenv("B200_GNB_ELEMS");
  int ppc = (env_ppc ? atoi(env_ppc) : 16384) / C;
  if (ppc < 1) ppc = 1;
  if (ppc > d->HW) ppc = d->HW;
  p.pix_per_cta = ppc;
  static const char* env_rev = getenv("B200_GNB_REVERSE");
  p.reverse = !(env_rev && atoi(env_rev) == 0);
  dim3 grid((d->HW + ppc - 1) / ppc, d->B);
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(gn_bwd_reduce_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B200_CHECK(cudaFuncSetAttribute(gn_bwd_reduce_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B200_CHECK(cudaFuncSetAttribute(gn_bwd_reduce_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B200_CHECK(cudaFuncSetAttribute(gn_bwd_apply_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B200_CHECK(cudaFuncSetAttribute(gn_bwd_apply_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    B200_CHECK(cudaFuncSetAttribute(gn_bwd_apply_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    attr = true;
  }
  const int mode = (p.resample == 0 && (unsigned long long)d->B * d->HW * C < (1ull << 32)) ? (p.C1 == 0 ? 1 : 2) : 0;
  // small images: everything in one launch, one CTA per (image, slice of whole groups), data read once (gn_bwd_slab_kernel)
  static const char* env_slab = getenv("B200_GNB_SLAB");      // =0: the four-kernel form everywhere (A/B timing)
  // largest image (pixels) that takes it.  32x32 images fit too (a 32-channel slice = 201 KB of shared memory) and were
  // measured (B200_GNB_SLAB_HW=1024): one CTA of 8 warps per SM cannot hide the load latency -- the twelve 32x32 adjoints
  // of the CFG UNet step got 64 us SLOWER each (training step 14.69 -> 15.53 ms), so they keep the four-kernel form.
  static const char* env_slab_hw = getenv("B200_GNB_SLAB_HW");
  const int slab_hw = env_slab_hw ? atoi(env_slab_hw) : 256;
  if (mode != 0 && d->HW <= slab_hw && !(env_slab && atoi(env_slab) == 0)) {
    // slice = the fewest whole groups that give >= 32 channels (32 groups per tensor: 1, 2, 4, ... groups divide it evenly);
    // 32x32 images: x + g of a 32-channel slice are 192 KB, so the slice shrinks (whole groups, >= 16 channels) until one
    // CTA's slab fits the 227 KB of shared memory
    const size_t smem_max = 227 * 1024;
    int gsl = 1;
    while (p.cpg * gsl < 32 && gsl < d->groups && d->groups % (gsl * 2) == 0) gsl *= 2;
    auto slab_bytes = [&](int cs) {
      const int nq_ = cs / 4, R_ = 256 / (nq_ > 0 ? nq_ : 1);
      return (size_t)d->HW * cs * 6 + (size_t)(8 + 2 * R_) * cs * 4;
    };
    while (gsl > 1 && p.cpg * (gsl / 2) >= 16 && slab_bytes(p.cpg * gsl) > smem_max) gsl /= 2;
    const int CS = p.cpg * gsl;
    const int nq = CS / 4;
    const size_t smem = slab_bytes(CS);
    if (CS % 4 == 0 && nq >= 1 && nq <= 64 && C % CS == 0 && ((size_t)d->HW * CS * 6) % 16 == 0 && smem <= smem_max) {
      static bool slab_attr = false;
      if (!slab_attr) {
        B200_CHECK(cudaFuncSetAttribute(gn_bwd_slab_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max));
        slab_attr = true;
      }
      gn_bwd_slab_kernel<<<dim3(C / CS, d->B), 256, smem, stream>>>(p, CS);
      g_launch_count += 1;
      return check_cuda(cudaGetLastError(), "gn_bwd slab kernel launch");
    }
  }
  gn_bwd_coef_kernel<<<d->B, 256, (size_t)2 * C * 4, stream>>>(p);
  if (mode == 1) gn_bwd_reduce_kernel<1><<<grid, 256, (size_t)2 * C * 4, stream>>>(p);
  else if (mode == 2) gn_bwd_reduce_kernel<2><<<grid, 256, (size_t)2 * C * 4, stream>>>(p);
  else gn_bwd_reduce_kernel<0><<<grid, 256, (size_t)2 * C * 4, stream>>>(p);
  gn_bwd_final_kernel<<<d->B, 256, 0, stream>>>(p);
  B200_CHECK(cudaGetLastError());
  if (mode == 1) gn_bwd_apply_kernel<1><<<grid, 256, (size_t)C * 4, stream>>>(p);
  else if (mode == 2) gn_bwd_apply_kernel<2><<<grid, 256, (size_t)C * 4, stream>>>(p);
  else gn_bwd_apply_kernel<0><<<grid, 256, (size_t)C * 4, stream>>>(p);
  g_launch_count += 4;
  return check_cuda(cudaGetLastError(), "gn_bwd kernels launch");
}

extern "C" int b200_cast_bf16_colsum(const float* x, void* out_bf16, float* colsum, long long rows, int C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out_bf16 && rows >= 1 && C >= 4 && C % 4 == 0 && C <= 32768, "cast_bf16_colsum: bad arguments");
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(cast_colsum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024));
    attr = true;
  }
  int rpc = 32768 / C;
  if (rpc < 1) rpc = 1;
  const long long grid = (rows + rpc - 1) / rpc;
  cast_colsum_kernel<<<(unsigned)grid, 256, (size_t)C * 4, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out_bf16), colsum,
                                                                   rows, C, rpc);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "cast_colsum_kernel launch");
}

extern "C" int b200_nchw_to_nhwc_pad_bf16(const float* x, void* out_bf16, float* colsum, int B, int C, int HW, int Cpad,
                                          void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out_bf16 && C >= 1 && C <= 8 && Cpad % 8 == 0 && Cpad >= 8, "nchw_to_nhwc_pad: needs C <= 8, Cpad % 8 == 0");
  nchw_pad_kernel<<<bw_grid((long long)B * HW, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out_bf16), colsum, B,
                                                                     C, HW, Cpad);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "nchw_pad_kernel launch");
}

extern "C" int b200_colsum_bf16(const void* x, float* colsum, long long rows, int ld, int c0, int C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && colsum && rows >= 1 && C >= 4 && C % 4 == 0 && c0 % 4 == 0 && ld % 4 == 0 && C <= 8192,
               "colsum_bf16: bad arguments");
  if (C % 8 == 0 && c0 % 8 == 0 && ld % 8 == 0 && ((uintptr_t)x & 15) == 0) {
    int rpc = 32768 / C;          // 64 KB of bf16 per CTA
    if (rpc < 8) rpc = 8;
    const long long grid = (rows + rpc - 1) / rpc;
    colsum_bf16x8_kernel<<<(unsigned)grid, 256, (size_t)C * 4, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), colsum,
                                                                       rows, ld, c0, C, rpc);
    ++g_launch_count;
    return check_cuda(cudaGetLastError(), "colsum_bf16x8_kernel launch");
  }
  int rpc = 65536 / C;
  if (rpc < 1) rpc = 1;
  const long long grid = (rows + rpc - 1) / rpc;
  colsum_bf16_kernel<<<(unsigned)grid, 256, (size_t)C * 4, stream>>>(reinterpret_cast<const __nv_bfloat16*>(x), colsum, rows,
                                                                   ld, c0, C, rpc);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "colsum_bf16_kernel launch");
}

extern "C" int b200_resample_f32(const float* x, float* out, int B, int H, int W, int C, int mode, float scale,
                                 int accumulate, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out && C % 4 == 0 && (mode == 1 || mode == 2), "resample_f32: bad arguments");
  if (mode == 1) B200_REQUIRE(H % 2 == 0 && W % 2 == 0, "resample_f32: average pool needs even H, W");
  const size_t total = (size_t)B * (mode == 1 ? (H / 2) * (W / 2) : 4 * H * W) * (C / 4);
  resample_f32_kernel<<<bw_grid((long long)total, 256), 256, 0, stream>>>(x, out, B, H, W, C, mode, scale, accumulate);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "resample_f32_kernel launch");
}

extern "C" int b200_upsample2_bf16(const float* x, void* out_bf16, int B, int H, int W, int C, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out_bf16 && C % 4 == 0, "upsample2_bf16: bad arguments");
  const size_t total = (size_t)B * 4 * H * W * (C / 4);
  upsample2_bf16_kernel<<<bw_grid((long long)total, 256), 256, 0, stream>>>(x, reinterpret_cast<__nv_bfloat16*>(out_bf16), B, H,
                                                                          W, C);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "upsample2_bf16_kernel launch");
}

extern "C" int b200_softmax_rows(const float* S, void* P_bf16, long long rows, int T, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(S && P_bf16 && rows >= 1 && T >= 1, "softmax_rows: bad arguments");
  softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(S, reinterpret_cast<__nv_bfloat16*>(P_bf16), rows, T, scale);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "softmax_rows_kernel launch");
}

extern "C" int b200_softmax_bwd_rows(const void* P_bf16, const float* dP, void* dS_bf16, long long rows, int T, float scale,
                                     void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(P_bf16 && dP && dS_bf16 && rows >= 1 && T >= 1, "softmax_bwd_rows: bad arguments");
  softmax_bwd_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, stream>>>(reinterpret_cast<const __nv_bfloat16*>(P_bf16), dP,
                                                                       reinterpret_cast<__nv_bfloat16*>(dS_bf16), rows, T, scale);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "softmax_bwd_rows_kernel launch");
}

extern "C" int b200_mse_loss(const float* a, const float* b, float* loss, long long n, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(a && b && loss && n >= 1, "mse_loss: bad arguments");
  B200_CHECK(cudaMemsetAsync(loss, 0, sizeof(float), stream));
  mse_loss_kernel<<<bw_grid(n, 256), 256, 0, stream>>>(a, b, loss, n, 1.0f / (float)n);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "mse_loss_kernel launch");
}

extern "C" int b200_mse_loss_grad(const float* a, const float* b, const float* grad_scale, float* da, long long n,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(a && b && da && n >= 1, "mse_loss_grad: bad arguments");
  mse_grad_kernel<<<bw_grid(n, 256), 256, 0, stream>>>(a, b, grad_scale, da, n, 2.0f / (float)n);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "mse_grad_kernel launch");
}

extern "C" int b200_dropout_mask(float* out, long long n, float p, unsigned long long seed, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(out && n >= 1 && p >= 0.f && p < 1.f, "dropout_mask: bad arguments");
  dropout_mask_kernel<<<bw_grid(n, 256), 256, 0, stream>>>(out, n, seed, dropout_threshold(p));
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "dropout_mask_kernel launch");
}

extern "C" int b200_time_embed_bwd(const int64_t* t, int rows, const float* freqs, int dim, int E, int cos_first,
                                   const float* w1, const float* b1, const float* w2, const float* emb, const float* d_semb,
                                   const int64_t* y, void* pe_bf16, void* hid_bf16, void* demb_bf16, void* dpre_bf16,
                                   float* db1, float* db2, float* dclass, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(t && freqs && w1 && b1 && w2 && emb && d_semb && pe_bf16 && hid_bf16 && demb_bf16 && dpre_bf16 && db1 && db2,
               "time_embed_bwd: null pointer");
  B200_REQUIRE(dim % 8 == 0 && E % 4 == 0 && rows >= 1, "time_embed_bwd: dim must be a multiple of 8, E of 4");
  const size_t smem = (size_t)(dim + 2 * E) * 4;
  B200_REQUIRE(smem <= 48 * 1024, "time_embed_bwd: dim+2E too large");
  time_embed_bwd_kernel<<<rows, 256, smem, stream>>>(
      t, freqs, dim, E, cos_first, w1, b1, w2, emb, d_semb, y, reinterpret_cast<__nv_bfloat16*>(pe_bf16),
      reinterpret_cast<__nv_bfloat16*>(hid_bf16), reinterpret_cast<__nv_bfloat16*>(demb_bf16),
      reinterpret_cast<__nv_bfloat16*>(dpre_bf16), db1, db2, dclass);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "time_embed_bwd_kernel launch");
}
