// Shared pieces of the implicit-GEMM convolution kernels (conv_gemm.cu: one CTA per tile; conv_gemm2.cu: CTA pairs).
#pragma once
#include "common.cuh"
#include "../../include/b200diff.h"

namespace b200 {

struct ConvKParams {
  int B, NP;
  int bw, bh, bn, lg_bw, lg_bhw;
  int tiles_w, tiles_h;
  int lg_tiles_w, lg_tiles_hw;   // log2(tiles_w), log2(tiles_w * tiles_h) when both are powers of two, else lg_tiles_hw = -1
  int phases;
  int p_tiles, c_tiles, total_tiles;
  int N, w_rows_per_phase;
  int cpb0, nkb0, nkb1;
  int stages;
  int8_t taps0[4][9][4];
  int8_t tap1[4];
  const float* bias;
  const float* rowadd;
  int rowadd_ld;
  const float* residual;
  int res_ld;
  long long* stats;  // [B][N][2] int64 fixed point (common.cuh: stat_add) sum / sum of squares of the fp32 output, or NULL
  void* out;
  int out_mode, out_ld, out_H, out_W, osy, osx, vec8_ok;
  int group4;    // bw >= 4: 4 consecutive tile pixels are 4 output pixels osx apart in one row
  int epi_halves;  // 1 or 2 epilogue warps per TMEM lane quarter
  int k_rotate;    // rotate the K-block order per CTA (spreads weight-tile requests over L2)
  int fast_epi;  // output pixel index is linear in the tile pixel index, all tiles full, >= 16 pixels per image
  int row_epi;   // 8 consecutive tile pixels are 8 output pixels osx apart in one row (strided / phase outputs), tiles full
  int vtap;      // vertical-tap reuse (conv_gemm.cu): one haloed pixel tile + 3 weight tiles per stage
  int dbg;       // timing experiments, builds with -DB200_DEBUG only (B200_EPI_DBG: 1 = no residual loads, 2 = no output
                 // stores, 4 = epilogue does nothing, 8 / 16 = no weight / pixel TMA); always 0 in the shipping library
  // fused next-GroupNorm epilogue (b200_conv2d_gn_fwd)
  const float* gn_gamma;
  const float* gn_beta;
  const float* gn_scale;
  const float* gn_shift;
  void* gn_out;
  int gn_out_ld;   // channel stride of gn_out (>= N: the normalised copy may be a column window of a wider tensor)
  int gn_raw;      // 1: conv2-style use -- the fp32 result (+ residual) is ALSO written to `out` (+ `stats`), see below
  void* gn_rawcopy;  // optional: bf16 copy of the un-normalised result, laid out like gn_out (the operand of
                     // the consumer's fused 1x1 shortcut when the consumer concatenates a skip connection)
  int gn_ss_ld, gn_lg_cpg, gn_silu;
  int gn_late_out;   // multi-tile block-output form: write the fp32 output AFTER the statistics arrival (see conv_epilogue_gnfuse)
  int gn_cl;     // tiles (= co-scheduled CTAs) per image in the multi-tile variant of the fused epilogue, else 0
  long long* gn_xstats;            // multi-tile variant: zeroed [B][N][2] int64 statistics of the conv output
  unsigned long long* gn_xcount;   // multi-tile variant: zeroed per-image arrival counters
  float gn_eps;
};

// The B200_EPI_DBG timing knob makes kernels skip loads / stores (wrong results by design): it exists only in builds with
// -DB200_DEBUG; the shipping library compiles every use to a constant 0.
#ifdef B200_DEBUG
#define B200_DBG(p) ((p).dbg)
#else
#define B200_DBG(p) 0
#endif

constexpr int kBlockC = 128;  // output channels per tile (UMMA M)
constexpr int kBlockK = 64;
constexpr int kWBytes = kBlockC * kBlockK * 2;  // 16 KB weight tile
constexpr int kThreads = 320;  // TMA warp + MMA warp + 8 epilogue warps
constexpr int kThreadsWide = 576;  // ... + 16 epilogue warps (epi_halves == 4; fewer registers per thread)
constexpr int kMaxStages = 8;

struct __align__(8) ConvBarriers {
  uint64_t full[kMaxStages];
  uint64_t empty[kMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

struct TileCoord {
  int ph, ct, w0, h0, n0;
};

__device__ __forceinline__ TileCoord decode_tile(const ConvKParams& p, int tile) {
  TileCoord t;
  // every warp of the CTA decodes every tile: the common cases (one phase, one or two channel tiles, power-of-two tile
  // grids) avoid the five integer divisions of the general form (~25 dependent instructions each)
  int rem = tile;
  t.ph = 0;
  if (p.phases > 1) {
    const int per_phase = p.p_tiles * p.c_tiles;
    t.ph = tile / per_phase;
    rem = tile - t.ph * per_phase;
  }
  int pt = rem;
  t.ct = 0;
  if (p.c_tiles == 2) { pt = rem >> 1; t.ct = rem & 1; }
  else if (p.c_tiles > 2) { pt = rem / p.c_tiles; t.ct = rem - pt * p.c_tiles; }
  int tw, th, tn;
  if (p.lg_tiles_hw >= 0) {
    tw = pt & (p.tiles_w - 1);
    th = (pt >> p.lg_tiles_w) & (p.tiles_h - 1);
    tn = pt >> p.lg_tiles_hw;
  } else {
    tw = pt % p.tiles_w;
    th = (pt / p.tiles_w) % p.tiles_h;
    tn = pt / (p.tiles_w * p.tiles_h);
  }
  t.w0 = tw * p.bw;
  t.h0 = th * p.bh;
  t.n0 = tn * p.bn;
  return t;
}

// Image index and pixel offset (row-major within the output image) of tile pixel pp.
__device__ __forceinline__ void decode_pixel(const ConvKParams& p, const TileCoord& t, int pp, int pa, int pb, int& n,
                                             int& po) {
  n = t.n0 + (pp >> p.lg_bhw);
  const int oy = (t.h0 + ((pp >> p.lg_bw) & (p.bh - 1))) * p.osy + pa;
  const int ox = (t.w0 + (pp & (p.bw - 1))) * p.osx + pb;
  po = oy * p.out_W + ox;
}

// Epilogue of one tile for one warp: TMEM lanes = 32 consecutive output channels (this thread owns channel c), TMEM
// columns = the tile's pixels.  Shared by the 1-CTA and the CTA-pair kernels.  (Fetching the residual a chunk ahead
// was measured: no gain -- the epilogue's cost is its HBM traffic competing with the operand stream, not its latency --
// and the doubled code slowed the short-K 1x1 layers, so one chunk is processed at a time.)
__device__ __forceinline__ void conv_epilogue_tile(const ConvKParams& p, const TileCoord& t, const uint32_t taddr, const int c,
                                                   const bool c_ok, const int half, uint64_t* acc_full_bar,
                                                   const uint32_t acc_parity) {
  const int hw_out = p.out_H * p.out_W;
  const size_t img_out = (size_t)hw_out * p.out_ld;   // elements per image of an NHWC output
  const size_t img_res = (size_t)hw_out * p.res_ld;
  const int ostep = p.osx * p.out_ld, rstep = p.osx * p.res_ld;
  const float* __restrict__ residual = (B200_DBG(p) & 1) ? nullptr : p.residual;
  const int pa = t.ph >> 1, pb = t.ph & 1;
  const float bias_c = (p.bias && c_ok) ? __ldg(p.bias + c) : 0.f;
  float s1 = 0.f, s2 = 0.f, ra_c = 0.f;
  int cur_n = -1;
  // fast path: the tile's pixels are consecutive output pixels (full-width rows / whole images, stride 1)
  const size_t pix0 = ((size_t)t.n0 * p.out_H + t.h0) * p.out_W + t.w0;
  const float* __restrict__ rbase = residual ? residual + pix0 * (size_t)p.res_ld + c : nullptr;

  auto load_res = [&](const int ch, float (&r)[32]) {
          if (p.fast_epi) {
#pragma unroll
            for (int j = 0; j < 32; ++j) r[j] = c_ok ? __ldg(rbase + (ch + j) * p.res_ld) : 0.f;
          } else if (p.group4) {
#pragma unroll
            for (int g = 0; g < 8; ++g) {
              int n, po;
              decode_pixel(p, t, ch + 4 * g, pa, pb, n, po);
              const float* rp = residual + (size_t)n * img_res + c + po * p.res_ld;
              const bool ok = c_ok && n < p.B;
#pragma unroll
              for (int e = 0; e < 4; ++e) r[4 * g + e] = ok ? __ldg(rp + e * rstep) : 0.f;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              int n, po;
              decode_pixel(p, t, ch + j, pa, pb, n, po);
              r[j] = (c_ok && n < p.B) ? __ldg(residual + (size_t)n * img_res + c + po * p.res_ld) : 0.f;
            }
          }
  };

  auto process = [&](const int ch) {
        uint32_t v[32];
        float r[32];
        __syncwarp();
        tmem_ld_x32(taddr + (uint32_t)ch, v);
        if (residual) load_res(ch, r);      // overlaps the TMEM load
        tmem_ld_wait();
        float acc[32];
        if (p.fast_epi) {
          // >= 16 pixels per image: the image index is constant over each half of the chunk; two independent
          // accumulator pairs shorten the dependent FADD/FFMA chains of the statistics
#pragma unroll
          for (int hf = 0; hf < 2; ++hf) {
            const int n = t.n0 + ((ch + 16 * hf) >> p.lg_bhw);
            if (n != cur_n) {  // warp-uniform: new image -> flush statistics, fetch the time-embedding value
              if (p.stats && c_ok && cur_n >= 0) {
                stat_add(p.stats + ((size_t)cur_n * p.N + c) * 2, s1, s2);
              }
              s1 = 0.f; s2 = 0.f;
              cur_n = n;
              ra_c = (p.rowadd && c_ok) ? __ldg(p.rowadd + (size_t)n * p.rowadd_ld + c) : 0.f;
            }
            const float add_c = bias_c + ra_c;
            float t1 = 0.f, t2 = 0.f;
#pragma unroll
            for (int j = 16 * hf; j < 16 * hf + 16; j += 2) {
              float a0 = __uint_as_float(v[j]) + add_c;
              float a1 = __uint_as_float(v[j + 1]) + add_c;
              if (residual) { a0 += r[j]; a1 += r[j + 1]; }
              acc[j] = a0;
              acc[j + 1] = a1;
              s1 += a0; t1 += a1;
              s2 = fmaf(a0, a0, s2); t2 = fmaf(a1, a1, t2);
            }
            s1 += t1;
            s2 += t2;
          }
        } else {
  #pragma unroll
          for (int g = 0; g < 8; ++g) {
            // the image index can only change between groups of 4 pixels when every image has >= 4 pixels per tile
            const int n = t.n0 + ((ch + 4 * g) >> p.lg_bhw);
            if (p.lg_bhw >= 2) {
              if (n != cur_n) {  // warp-uniform: new image -> flush statistics, fetch the time-embedding value
                if (p.stats && c_ok && cur_n >= 0 && cur_n < p.B) {
                  stat_add(p.stats + ((size_t)cur_n * p.N + c) * 2, s1, s2);
                }
                s1 = 0.f; s2 = 0.f;
                cur_n = n;
                ra_c = (p.rowadd && c_ok && n < p.B) ? __ldg(p.rowadd + (size_t)n * p.rowadd_ld + c) : 0.f;
              }
              const float add_c = bias_c + ra_c;
              const bool n_ok = n < p.B;
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                float a = __uint_as_float(v[4 * g + e]) + add_c;
                if (residual) a += r[4 * g + e];
                acc[4 * g + e] = a;
                if (n_ok) { s1 += a; s2 += a * a; }
              }
            } else {
  #pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int ne = t.n0 + ((ch + 4 * g + e) >> p.lg_bhw);
                if (ne != cur_n) {
                  if (p.stats && c_ok && cur_n >= 0 && cur_n < p.B) {
                    stat_add(p.stats + ((size_t)cur_n * p.N + c) * 2, s1, s2);
                  }
                  s1 = 0.f; s2 = 0.f;
                  cur_n = ne;
                  ra_c = (p.rowadd && c_ok && ne < p.B) ? __ldg(p.rowadd + (size_t)ne * p.rowadd_ld + c) : 0.f;
                }
                float a = __uint_as_float(v[4 * g + e]) + bias_c + ra_c;
                if (residual) a += r[4 * g + e];
                acc[4 * g + e] = a;
                if (ne < p.B) { s1 += a; s2 += a * a; }
              }
            }
          }
        }
        if (B200_DBG(p) & 2) return;
        if (p.fast_epi && p.out_mode == B200_OUT_F32_NHWC) {
          float* __restrict__ ob = reinterpret_cast<float*>(p.out) + pix0 * (size_t)p.out_ld + c;
          if (c_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) ob[(ch + j) * p.out_ld] = acc[j];
          }
        } else if (p.fast_epi && p.out_mode == B200_OUT_BF16_NHWC) {
          __nv_bfloat16* __restrict__ ob = reinterpret_cast<__nv_bfloat16*>(p.out) + pix0 * (size_t)p.out_ld + c;
          if (c_ok) {
#pragma unroll
            for (int j = 0; j < 32; ++j) ob[(ch + j) * p.out_ld] = __float2bfloat16_rn(acc[j]);
          }
        } else if (p.group4 && p.out_mode <= B200_OUT_BF16_NHWC) {
#pragma unroll
          for (int g = 0; g < 8; ++g) {
            int n, po;
            decode_pixel(p, t, ch + 4 * g, pa, pb, n, po);
            if (c_ok && n < p.B) {
              if (p.out_mode == B200_OUT_F32_NHWC) {
                float* __restrict__ o = reinterpret_cast<float*>(p.out) + (size_t)n * img_out + c + po * p.out_ld;
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e * ostep] = acc[4 * g + e];
              } else {
                __nv_bfloat16* __restrict__ o =
                    reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)n * img_out + c + po * p.out_ld;
#pragma unroll
                for (int e = 0; e < 4; ++e) o[e * ostep] = __float2bfloat16_rn(acc[4 * g + e]);
              }
            }
          }
        } else if (p.out_mode <= B200_OUT_BF16_NHWC) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            int n, po;
            decode_pixel(p, t, ch + j, pa, pb, n, po);
            if (c_ok && n < p.B) {
              const size_t e = (size_t)n * img_out + c + po * p.out_ld;
              if (p.out_mode == B200_OUT_F32_NHWC) reinterpret_cast<float*>(p.out)[e] = acc[j];
              else reinterpret_cast<__nv_bfloat16*>(p.out)[e] = __float2bfloat16_rn(acc[j]);
            }
          }
        } else if (p.vec8_ok) {
          // channel-major outputs: this thread owns a row of consecutive pixels -> 8-pixel vector stores
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            int n, po;
            decode_pixel(p, t, ch + 8 * g, 0, 0, n, po);
            const size_t e = ((size_t)n * p.out_ld + c) * hw_out + po;
            if (c_ok && n < p.B) {
              if (p.out_mode == B200_OUT_F32_NCHW) {
                float4* o = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.out) + e);
                o[0] = make_float4(acc[8 * g], acc[8 * g + 1], acc[8 * g + 2], acc[8 * g + 3]);
                o[1] = make_float4(acc[8 * g + 4], acc[8 * g + 5], acc[8 * g + 6], acc[8 * g + 7]);
              } else {
                uint4 u;
                u.x = pack_bf16x2(acc[8 * g], acc[8 * g + 1]);
                u.y = pack_bf16x2(acc[8 * g + 2], acc[8 * g + 3]);
                u.z = pack_bf16x2(acc[8 * g + 4], acc[8 * g + 5]);
                u.w = pack_bf16x2(acc[8 * g + 6], acc[8 * g + 7]);
                *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.out) + e) = u;
              }
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            int n, po;
            decode_pixel(p, t, ch + j, pa, pb, n, po);
            const size_t e = ((size_t)n * p.out_ld + c) * hw_out + po;
            if (c_ok && n < p.B) {
              if (p.out_mode == B200_OUT_F32_NCHW) reinterpret_cast<float*>(p.out)[e] = acc[j];
              else reinterpret_cast<__nv_bfloat16*>(p.out)[e] = __float2bfloat16_rn(acc[j]);
            }
          }
        }
  };

  mbar_wait(acc_full_bar, acc_parity);
  tc_fence_after();
  if (B200_DBG(p) & 4) return;
#pragma unroll 1
  for (int ch = half * 32; ch < p.NP; ch += 32 * p.epi_halves) process(ch);
      if (p.stats && c_ok && cur_n >= 0 && cur_n < p.B) {
        stat_add(p.stats + ((size_t)cur_n * p.N + c) * 2, s1, s2);
      }
}

// ------------------------------------------------------------------------------------------------------------------
// Lean epilogue (kernel instantiations <kT, 1>; the default for every layer whose tile pixels are linear in the output,
// B200_EPI_LEAN=0 selects the generic epilogue for A/B timing).  Motivation: ncu on the attention blocks' 1x1 projections (4 K-blocks per tile, so the
// epilogue is the critical path) shows ~560 warp instructions per 32-pixel chunk, 217 of them 64-bit address
// arithmetic for the 32 stores and ~100 statistics FADD/FFMA that run even when no statistics are requested.  Here the
// uniform options are template parameters and every store / residual address is ONE IMAD.WIDE (32-bit byte stride x
// compile-time pixel index + 64-bit chunk base).  Covers the fast_epi tiles (pixels linear in the output) with NHWC
// fp32 / bf16 outputs; everything else takes conv_epilogue_tile.  Same arithmetic, same
// order of the per-thread statistics sums as the standard fast path.
// Measured (round 2, B200): every conv parity case and the forward / DDIM-50 / CFG / training e2e cases pass; DDIM-50
// CIFAR-10 1035 -> 1040 images/s, CFG training step 17.12 -> 16.54 ms.
// ------------------------------------------------------------------------------------------------------------------
template <bool BF16_OUT, bool HAS_RES, bool HAS_STATS, bool HAS_ROW>
__device__ __forceinline__ void conv_epilogue_lean(const ConvKParams& p, const TileCoord& t, const uint32_t taddr,
                                                   const int c, const bool c_ok, const int half,
                                                   uint64_t* acc_full_bar, const uint32_t acc_parity) {
  constexpr int kOutBytes = BF16_OUT ? 2 : 4;
  const size_t pix0 = ((size_t)t.n0 * p.out_H + t.h0) * p.out_W + t.w0;
  const int ost = p.out_ld * kOutBytes;        // byte strides between consecutive pixels
  const int rst = p.res_ld * 4;
  char* const obase = reinterpret_cast<char*>(p.out) + (pix0 * (size_t)p.out_ld + c) * kOutBytes;
  const char* const rbase = HAS_RES ? reinterpret_cast<const char*>(p.residual + pix0 * (size_t)p.res_ld + c) : nullptr;
  const float bias_c = (p.bias && c_ok) ? __ldg(p.bias + c) : 0.f;
  float s1 = 0.f, s2 = 0.f, ra_c = 0.f;
  int cur_n = -1;
  if (HAS_RES) {
    // The residual of the whole tile is requested into L2 BEFORE waiting for the accumulator, so the DRAM round trip
    // overlaps the tile's MMAs: one prefetch instruction per 32-pixel chunk (lane j asks for pixel j's 128-byte line of
    // this warp's 32 channels), no registers held.  ncu on the attention blocks' 1x1 output projection (4 K-blocks per
    // tile; profiles/r02_prof_conv.details.txt id 6) had shown the tensor pipe 7 % active, DRAM at 31 % and the residual
    // FADDs as the top stall sites: eight epilogue warps with one 32-load batch each in flight and ~1.5 us per DRAM
    // round trip cap such a layer at ~3 TB/s.  (Holding a second chunk's loads in registers instead spilled: 168
    // registers is the cap for 10 warps.)
    const int lane = threadIdx.x & 31;
    const char* pf = rbase - (long long)lane * 4;      // channel of lane 0: start of the 128-byte line
    for (int ch = half * 32; ch < p.NP; ch += 32 * p.epi_halves)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (long long)(ch + lane) * rst));
  }
  mbar_wait(acc_full_bar, acc_parity);
  tc_fence_after();
#pragma unroll 1
  for (int ch = half * 32; ch < p.NP; ch += 32 * p.epi_halves) {
    uint32_t v[32];
    float r[32];
    __syncwarp();
    tmem_ld_x32(taddr + (uint32_t)ch, v);
    if (HAS_RES) {      // overlaps the TMEM load
      const char* rp = rbase + (long long)ch * rst;
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = c_ok ? __ldg(reinterpret_cast<const float*>(rp + (long long)j * rst)) : 0.f;
    }
    tmem_ld_wait();
    float acc[32];
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      if (HAS_STATS || HAS_ROW) {
        const int n = t.n0 + ((ch + 16 * hf) >> p.lg_bhw);
        if (n != cur_n) {  // warp-uniform: new image -> flush statistics, fetch the time-embedding value
          if (HAS_STATS && c_ok && cur_n >= 0) {
            stat_add(p.stats + ((size_t)cur_n * p.N + c) * 2, s1, s2);
          }
          s1 = 0.f; s2 = 0.f;
          cur_n = n;
          if (HAS_ROW) ra_c = c_ok ? __ldg(p.rowadd + (size_t)n * p.rowadd_ld + c) : 0.f;
        }
      }
      const float add_c = bias_c + ra_c;
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int j = 16 * hf; j < 16 * hf + 16; j += 2) {
        float a0 = __uint_as_float(v[j]) + add_c;
        float a1 = __uint_as_float(v[j + 1]) + add_c;
        if (HAS_RES) { a0 += r[j]; a1 += r[j + 1]; }
        acc[j] = a0;
        acc[j + 1] = a1;
        if (HAS_STATS) {
          s1 += a0; t1 += a1;
          s2 = fmaf(a0, a0, s2); t2 = fmaf(a1, a1, t2);
        }
      }
      if (HAS_STATS) { s1 += t1; s2 += t2; }
    }
    if (c_ok) {
      char* op = obase + (long long)ch * ost;
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (BF16_OUT) *reinterpret_cast<__nv_bfloat16*>(op + (long long)j * ost) = __float2bfloat16_rn(acc[j]);
        else *reinterpret_cast<float*>(op + (long long)j * ost) = acc[j];
      }
    }
  }
  if (HAS_STATS && c_ok && cur_n >= 0 && cur_n < p.B) {
    stat_add(p.stats + ((size_t)cur_n * p.N + c) * 2, s1, s2);
  }
}

// Lean epilogue for outputs that are only ROW-wise linear in the tile pixel index: the four phase convolutions of the
// nearest-2x + 3x3 up-sampling layers (output pixel = 2 * tile pixel + phase offset) and other strided writes.  Groups of 8
// consecutive tile pixels lie in one tile row (bw >= 8), so a group costs one base-address computation and its 8 stores
// are compile-time multiples of the pixel stride.  No residual / time-embedding row (the up-sampling convs have neither).
// The generic epilogue decodes every group of 4 pixels with 64-bit arithmetic per store: the 256 -> 256 up-conv at
// 16x16 -> 32x32 (B = 256) ran 25 k cycles per tile against 8 k cycles of MMA work.
template <bool BF16_OUT, bool HAS_STATS>
__device__ __forceinline__ void conv_epilogue_lean_rows(const ConvKParams& p, const TileCoord& t, const uint32_t taddr,
                                                        const int c, const bool c_ok, const int half,
                                                        uint64_t* acc_full_bar, const uint32_t acc_parity) {
  constexpr int kOutBytes = BF16_OUT ? 2 : 4;
  const int pa = t.ph >> 1, pb = t.ph & 1;
  const long long pst = (long long)p.osx * p.out_ld * kOutBytes;     // bytes between consecutive tile pixels of a row
  char* const obase = reinterpret_cast<char*>(p.out) + (size_t)c * kOutBytes;
  const float bias_c = (p.bias && c_ok) ? __ldg(p.bias + c) : 0.f;
  float s1 = 0.f, s2 = 0.f;
  int cur_n = -1;
  mbar_wait(acc_full_bar, acc_parity);
  tc_fence_after();
#pragma unroll 1
  for (int ch = half * 32; ch < p.NP; ch += 32 * p.epi_halves) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(taddr + (uint32_t)ch, v);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int pp = ch + 8 * g;
      const int n = t.n0 + (pp >> p.lg_bhw);
      if (HAS_STATS && n != cur_n) {      // warp-uniform: new image -> flush statistics
        if (c_ok && cur_n >= 0) stat_add(p.stats + ((size_t)cur_n * p.N + c) * 2, s1, s2);
        s1 = 0.f; s2 = 0.f;
        cur_n = n;
      }
      const int oy = (t.h0 + ((pp >> p.lg_bw) & (p.bh - 1))) * p.osy + pa;
      const int ox = (t.w0 + (pp & (p.bw - 1))) * p.osx + pb;
      char* op = obase + ((size_t)(n * p.out_H + oy) * p.out_W + ox) * (size_t)(p.out_ld * kOutBytes);
      float t1 = 0.f, t2 = 0.f;
#pragma unroll
      for (int j = 0; j < 8; j += 2) {
        const float a0 = __uint_as_float(v[8 * g + j]) + bias_c, a1 = __uint_as_float(v[8 * g + j + 1]) + bias_c;
        if (c_ok) {
          if (BF16_OUT) {
            *reinterpret_cast<__nv_bfloat16*>(op + (long long)j * pst) = __float2bfloat16_rn(a0);
            *reinterpret_cast<__nv_bfloat16*>(op + (long long)(j + 1) * pst) = __float2bfloat16_rn(a1);
          } else {
            *reinterpret_cast<float*>(op + (long long)j * pst) = a0;
            *reinterpret_cast<float*>(op + (long long)(j + 1) * pst) = a1;
          }
        }
        if (HAS_STATS) {
          s1 += a0; t1 += a1;
          s2 = fmaf(a0, a0, s2); t2 = fmaf(a1, a1, t2);
        }
      }
      if (HAS_STATS) { s1 += t1; s2 += t2; }
    }
  }
  if (HAS_STATS && c_ok && cur_n >= 0 && cur_n < p.B) stat_add(p.stats + ((size_t)cur_n * p.N + c) * 2, s1, s2);
}

// Host-side eligibility of a layer for the lean epilogue (also asserted by the device dispatch below).
inline bool conv_epilogue_lean_ok(const ConvKParams& p) {
  return (p.fast_epi || p.row_epi) && p.out_mode <= B200_OUT_BF16_NHWC && B200_DBG(p) == 0;
}

__device__ __forceinline__ void conv_epilogue_lean_dispatch(const ConvKParams& p, const TileCoord& t, const uint32_t taddr,
                                                            const int c, const bool c_ok, const int half,
                                                            uint64_t* bar, const uint32_t parity) {
  const bool bf = p.out_mode == B200_OUT_BF16_NHWC, rs = p.residual != nullptr, st = p.stats != nullptr;
  const bool rw = p.rowadd != nullptr;
  if (!p.fast_epi && p.row_epi && p.out_mode <= B200_OUT_BF16_NHWC && B200_DBG(p) == 0) {
    if (bf) { if (st) conv_epilogue_lean_rows<true, true>(p, t, taddr, c, c_ok, half, bar, parity);
              else conv_epilogue_lean_rows<true, false>(p, t, taddr, c, c_ok, half, bar, parity); }
    else { if (st) conv_epilogue_lean_rows<false, true>(p, t, taddr, c, c_ok, half, bar, parity);
           else conv_epilogue_lean_rows<false, false>(p, t, taddr, c, c_ok, half, bar, parity); }
    return;
  }
  if (!(p.fast_epi && p.out_mode <= B200_OUT_BF16_NHWC && B200_DBG(p) == 0)) {
    conv_epilogue_tile(p, t, taddr, c, c_ok, half, bar, parity);
    return;
  }
#define B200_LEAN_CASE(BF, RS, ST, RW) \
  if (bf == BF && rs == RS && st == ST && rw == RW) { conv_epilogue_lean<BF, RS, ST, RW>(p, t, taddr, c, c_ok, half, bar, parity); return; }
  B200_LEAN_CASE(true, false, false, false) B200_LEAN_CASE(true, false, false, true)
  B200_LEAN_CASE(true, false, true, false)  B200_LEAN_CASE(true, false, true, true)
  B200_LEAN_CASE(true, true, false, false)  B200_LEAN_CASE(true, true, false, true)
  B200_LEAN_CASE(true, true, true, false)   B200_LEAN_CASE(true, true, true, true)
  B200_LEAN_CASE(false, false, false, false) B200_LEAN_CASE(false, false, false, true)
  B200_LEAN_CASE(false, false, true, false)  B200_LEAN_CASE(false, false, true, true)
  B200_LEAN_CASE(false, true, false, false)  B200_LEAN_CASE(false, true, false, true)
  B200_LEAN_CASE(false, true, true, false)   B200_LEAN_CASE(false, true, true, true)
#undef B200_LEAN_CASE
}

// ------------------------------------------------------------------------------------------------------------------
// Fused next-GroupNorm epilogue (kernel instantiation <kThreads, 2>, entry b200_conv2d_gn_fwd; validated on a B200 in
// round 2: tests/kernel_cases.py conv_gnfuse, DDIM-50 CIFAR-10 1035 -> 1057 images/s).  The tile is 64 / 128 / 256 pixels = whole images (2^lg_bhw pixels
// each, lg_bhw in {4, 6, 8}) x 128 channels = whole groups, so the statistics of the GroupNorm that consumes this conv's output are
// complete inside the CTA:
//   pass 1  a = acc + bias (+ time-embedding row); per-thread (= per-channel) sums over each 16-pixel half chunk
//   exchange the two warps that share a TMEM lane quarter own alternating 32-pixel chunks: per-chunk sums go through
//           shared memory (double-buffered by tile parity, one named barrier of the 8 epilogue warps per tile)
//   reduce  butterfly over the 2^lg_cpg lanes of a group -> mean / rstd -> A = rstd*gamma', B = beta' - mean*A
//   pass 2  the accumulators are read from TMEM again (they stay valid until tmem_empty is signalled by the caller):
//           y = SiLU(a*A + B) -> bf16 NHWC, the operand of the next convolution.  Nothing else is written.
// Same formulas as groupnorm_apply_kernel (biased variance clamped at 0, rsqrtf, SiLU through tanh.approx).
// ------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float epi_silu_tanh(float x) {
  float th;
  asm("tanh.approx.f32 %0, %1;" : "=f"(th) : "f"(0.5f * x));
  return x * fmaf(0.5f, th, 0.5f);   // identical to silu_tanh in groupnorm.cu
}

// Multi-tile variant (MULTI = true, kernel instantiation <kThreads, 3>): an image spans p.gn_cl tiles (32x32 images = 4
// tiles of 256 pixels).  The persistent grid is a multiple of p.gn_cl, so the p.gn_cl tiles of an image are processed in
// the same iteration by p.gn_cl CTAs, which are co-resident because the kernel is launched cooperatively.  Every epilogue
// warp adds its per-channel sums to the image's [N][2] int64 fixed-point statistics in global memory (the same
// order-independent accumulation the separate GroupNorm kernel consumes: bitwise reproducible), announces itself on a
// per-image counter and waits until all 8 * gn_cl warps of the image have arrived; the accumulators stay in TMEM
// meanwhile (the other TMEM buffer keeps the MMA warp busy with the next tile), then pass 2 normalises as usual.
// (A first version exchanged the sums between the CTAs of a 4-CTA thread-block cluster through DSMEM: correct, but only
// 33 clusters = 132 of the 148 SMs can be resident, which cost more than the GroupNorm launches saved.)
__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
// group sums read through L2 (written by other SMs during this kernel: no read-only / L1 path)
__device__ __forceinline__ float2 stat_load_group_cg(const long long* pair, int cnt) {
  long long a = 0, b = 0;
  for (int i = 0; i < cnt; ++i) {
    const longlong2 v = __ldcg(reinterpret_cast<const longlong2*>(pair) + i);
    a += v.x; b += v.y;
  }
  return make_float2(__ll2float_rn(a) * (1.0f / kStatQ1), __ll2float_rn(b) * (1.0f / kStatQ2));
}

// RAW = true (conv2 -> the NEXT block's norm1, models/unet.py:30-43 across two ResBlocks; the last block -> the output
// head's norm): the conv result x = acc + bias (+ residual) is a block output, so it is also written as fp32 NHWC to
// p.out with its per-channel statistics in p.stats (skip connections, attention blocks and the next block's residual
// read them), and pass 1 stores x back into TMEM so that pass 2 normalises the final values without a second read of
// the residual.  The normalised copy goes to p.gn_out with channel stride p.gn_out_ld.
// Loops over the 32-pixel chunks are ROLLED and a thread keeps scalar accumulators and one or two (A, B) coefficient pairs:
// the tile's chunks map to images uniformly -- the whole tile inside ONE image (MULTI: 32x32 images over 2 / 4
// co-scheduled tiles; one image per tile: 16x16), one image per 64-pixel chunk pair (8x8: chunk i of this warp + chunk i of
// its partner = image n0 + i; per-chunk sums travel through the shared-memory exchange buffer), or one image per 16-pixel
// half chunk (4x4: no partner at all).  A first, fully unrolled form ran 2445 straight-line instructions per warp and
// tile, once each: ncu attributed 23 % of the epilogue warps' samples to instruction fetch (stall_no_inst) on the 32x32
// layers, which are bound by this epilogue and not by the tensor pipe (DDIM-50 +3.7 % with the rolled form).
template <bool HAS_ROW, bool HAS_SS, bool MULTI, bool RAW>
__device__ __forceinline__ void conv_epilogue_gnfuse(const ConvKParams& p, const TileCoord& t, const uint32_t taddr,
                                                             const int c, const int half, const int cl, float* xbuf,
                                                             float* xbuf16, uint64_t* acc_full_bar, const uint32_t acc_parity) {
  const size_t pix0 = ((size_t)t.n0 * p.out_H + t.h0) * p.out_W + t.w0;
  const bool per = !MULTI && p.lg_bhw == 6;      // image n0 + i per chunk index i
  const bool p16 = !MULTI && p.lg_bhw == 4;      // 4x4 images: every 16-pixel half chunk is a whole image of this warp
  const int nch = p.NP >> 6;
  const int rst = p.res_ld * 4, wst = p.out_ld * 4;
  const char* const rbase = (RAW && p.residual) ? reinterpret_cast<const char*>(p.residual + pix0 * (size_t)p.res_ld + c) : nullptr;
  char* const wbase = (RAW && p.out) ? reinterpret_cast<char*>(reinterpret_cast<float*>(p.out) + pix0 * (size_t)p.out_ld + c) : nullptr;
  if (RAW && rbase != nullptr) {
    const int lane = threadIdx.x & 31;
    const char* pf = rbase - (long long)lane * 4;
    for (int i = 0; i < nch; ++i)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(pf + (long long)(half * 32 + 64 * i + lane) * rst));
  }
  const float bias_c = p.bias ? __ldg(p.bias + c) : 0.f;
  const bool early_out = RAW && wbase != nullptr && !(MULTI && p.gn_late_out);
  mbar_wait(acc_full_bar, acc_parity);
  tc_fence_after();
  // ---- pass 1: x = acc + bias (+ row) (+ residual); sums; RAW: x back to TMEM ----
  float S1 = 0.f, S2 = 0.f;      // tile-uniform case: this warp's share of the image's per-channel sums
#pragma unroll 1
  for (int i = 0; i < nch; ++i) {
    const int ch = half * 32 + 64 * i;
    const int n = t.n0 + (per ? i : p16 ? (ch >> 4) : 0);      // p16: image of the first half chunk, the second is n + 1
    uint32_t v[32];
    float r[32];
    __syncwarp();
    tmem_ld_x32(taddr + (uint32_t)ch, v);
    if (RAW && rbase != nullptr) {
      const char* rp = rbase + (long long)ch * rst;
#pragma unroll
      for (int j = 0; j < 32; ++j) r[j] = __ldg(reinterpret_cast<const float*>(rp + (long long)j * rst));
    }
    const float add_c = bias_c + (HAS_ROW ? __ldg(p.rowadd + (size_t)n * p.rowadd_ld + c) : 0.f);
    const float add_d = (HAS_ROW && p16) ? bias_c + __ldg(p.rowadd + (size_t)(n + 1) * p.rowadd_ld + c) : add_c;
    tmem_ld_wait();
    // (s1a, s2a): columns 0..15, (s1b, s2b): columns 16..31 of the chunk (two images in the p16 case)
    float s1a = 0.f, s1b = 0.f, s2a = 0.f, s2b = 0.f;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      float a0 = __uint_as_float(v[j]) + add_c, a1 = __uint_as_float(v[j + 16]) + add_d;
      if (RAW) {
        if (rbase != nullptr) { a0 += r[j]; a1 += r[j + 16]; }
        v[j] = __float_as_uint(a0);
        v[j + 16] = __float_as_uint(a1);
      }
      s1a += a0; s1b += a1;
      s2a = fmaf(a0, a0, s2a); s2b = fmaf(a1, a1, s2b);
    }
    if (RAW) {
      tmem_st_x32(taddr + (uint32_t)ch, v);
      if (early_out) {
        char* wp = wbase + (long long)ch * wst;
#pragma unroll
        for (int j = 0; j < 32; ++j) *reinterpret_cast<float*>(wp + (long long)j * wst) = __uint_as_float(v[j]);
      }
    }
    const float c1 = s1a + s1b, c2 = s2a + s2b;
    if (p16) {
      // no partner: both parity halves of the exchange buffer hold this tile's [half][chunk][half chunk][channel] sums
      if (RAW && p.stats != nullptr) {
        stat_add(p.stats + ((size_t)n * p.N + c) * 2, s1a, s2a);
        stat_add(p.stats + ((size_t)(n + 1) * p.N + c) * 2, s1b, s2b);
      }
      float* slot = xbuf16 + ((((half * 4 + i) * 2) * 128 + cl) << 1);
      *reinterpret_cast<float2*>(slot) = make_float2(s1a, s2a);
      *reinterpret_cast<float2*>(slot + 256) = make_float2(s1b, s2b);
    } else if (per) {
      if (RAW && p.stats != nullptr) stat_add(p.stats + ((size_t)n * p.N + c) * 2, c1, c2);
      *reinterpret_cast<float2*>(xbuf + (((half * 4 + i) * 128 + cl) << 1)) = make_float2(c1, c2);
    } else {
      S1 += c1; S2 += c2;
    }
  }
  if (RAW) tmem_st_wait();
  if (MULTI) {
    // the image's statistics live in global memory (gn_xstats doubles as the output statistics in the RAW form)
    const int n = t.n0;
    stat_add(p.gn_xstats + ((size_t)n * p.N + c) * 2, S1, S2);
    __syncwarp();
    unsigned long long* cnt = p.gn_xcount + n;
    if ((threadIdx.x & 31) == 0)
      asm volatile("red.release.gpu.global.add.u64 [%0], %1;" ::"l"(cnt), "l"(1ULL) : "memory");
    if (RAW && wbase != nullptr && !early_out) {     // fp32 output after the arrival: fills the wait for the other tiles
#pragma unroll 1
      for (int i = 0; i < nch; ++i) {
        const int ch = half * 32 + 64 * i;
        uint32_t v[32];
        __syncwarp();
        tmem_ld_x32(taddr + (uint32_t)ch, v);
        tmem_ld_wait();
        char* wp = wbase + (long long)ch * wst;
#pragma unroll
        for (int j = 0; j < 32; ++j) *reinterpret_cast<float*>(wp + (long long)j * wst) = __uint_as_float(v[j]);
      }
    }
    if ((threadIdx.x & 31) == 0) {
      const unsigned long long target = (unsigned long long)(8 * p.gn_cl);
      uint64_t t0 = 0;
      uint32_t spins = 0;
      while (ld_relaxed_u64(cnt) < target) {
        if ((++spins & 0x3ff) == 0) {
          const uint64_t now = globaltimer_ns();
          if (t0 == 0) t0 = now;
          else if (now - t0 > 4000000000ull) {
            printf("b200diff: fused GroupNorm statistics wait timeout (block %d image %d)\n", blockIdx.x, n);
            __trap();
          }
        }
      }
      // relaxed polling, then ONE acquire load of the final value (synchronises with the release-adds it observes).
      // A fence.acq_rel here also orders this warp's earlier writes, i.e. waits for the 128 fp32 output stores issued
      // just above: ncu showed 7 % of the epilogue warps' samples in that MEMBAR.
      (void)ld_acquire_u64(cnt);
    }
    __syncwarp();
    const int g0 = (c >> p.gn_lg_cpg) << p.gn_lg_cpg;
    const float2 gs = stat_load_group_cg(p.gn_xstats + ((size_t)n * p.N + g0) * 2, 1 << p.gn_lg_cpg);
    S1 = gs.x; S2 = gs.y;
  } else {
    if (!per && !p16) {
      if (RAW && p.stats != nullptr) stat_add(p.stats + ((size_t)t.n0 * p.N + c) * 2, S1, S2);
      *reinterpret_cast<float2*>(xbuf + ((half * 4 * 128 + cl) << 1)) = make_float2(S1, S2);
    }
    // the partner warp (same channels, the other pixels of the image[s]) through shared memory; p16: own values only,
    // but a warp may not overwrite them for the NEXT tile while its own pass 2 still reads -- program order gives that
    if (!p16) asm volatile("bar.sync 1, 256;" ::: "memory");
    else __syncwarp();
  }
  const float inv_cnt = 1.0f / (float)(((1 << p.lg_bhw) << p.gn_lg_cpg) * (MULTI ? p.gn_cl : 1));
  const float gamma_c = p.gn_gamma ? __ldg(p.gn_gamma + c) : 1.f, beta_c = p.gn_beta ? __ldg(p.gn_beta + c) : 0.f;
  __nv_bfloat16* const obase = reinterpret_cast<__nv_bfloat16*>(p.gn_out) + pix0 * (size_t)p.gn_out_ld + c;
  const int ost = p.gn_out_ld * 2;
  const bool rawcopy = p.gn_rawcopy != nullptr;
  __nv_bfloat16* const rcbase = rawcopy ? reinterpret_cast<__nv_bfloat16*>(p.gn_rawcopy) + pix0 * (size_t)p.gn_out_ld + c : nullptr;
  const bool silu = p.gn_silu != 0;
  float A = 0.f, B = 0.f, radd = 0.f;
  // ---- pass 2: y = (x [+ bias + row]) * A + B ----
#pragma unroll 1
  for (int i = 0; i < nch; ++i) {
    const int ch = half * 32 + 64 * i;
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(taddr + (uint32_t)ch, v);
    float A2 = 0.f, B2 = 0.f, radd2 = 0.f;      // p16: second half chunk
    if (per || p16 || i == 0) {      // coefficients of this chunk's image(s) (tile-uniform: once)
      const int n = t.n0 + (per ? i : p16 ? (ch >> 4) : 0);
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        if (hf == 1 && !p16) break;
        if (p16) {
          const float2 own = *reinterpret_cast<const float2*>(xbuf16 + ((((half * 4 + i) * 2 + hf) * 128 + cl) << 1));
          S1 = own.x; S2 = own.y;
        } else if (!MULTI) {
          const float2 own = *reinterpret_cast<const float2*>(xbuf + (((half * 4 + (per ? i : 0)) * 128 + cl) << 1));
          const float2 oth = *reinterpret_cast<const float2*>(xbuf + ((((1 - half) * 4 + (per ? i : 0)) * 128 + cl) << 1));
          S1 = own.x + oth.x; S2 = own.y + oth.y;
        }
        if (!MULTI)
          for (int m = 1; m < (1 << p.gn_lg_cpg); m <<= 1) {
            S1 += __shfl_xor_sync(0xffffffffu, S1, m);
            S2 += __shfl_xor_sync(0xffffffffu, S2, m);
          }
        const float mean = S1 * inv_cnt;
        const float var = fmaxf(S2 * inv_cnt - mean * mean, 0.f);
        const float rstd = rsqrtf(var + p.gn_eps);
        float ga = gamma_c, be = beta_c;
        if (HAS_SS) {
          const float sc = 1.f + __ldg(p.gn_scale + (size_t)(n + hf) * p.gn_ss_ld + c);
          ga *= sc;
          be = be * sc + __ldg(p.gn_shift + (size_t)(n + hf) * p.gn_ss_ld + c);
        }
        // RAW: TMEM already holds the final x; else x = acc + radd, rounded as in pass 1 (both forms give the same bits)
        const float ra_ = RAW ? 0.f : bias_c + (HAS_ROW ? __ldg(p.rowadd + (size_t)(n + hf) * p.rowadd_ld + c) : 0.f);
        if (hf == 0) { A = rstd * ga; B = be - mean * A; radd = ra_; }
        else { A2 = rstd * ga; B2 = be - mean * A2; radd2 = ra_; }
      }
    }
    if (!p16) { A2 = A; B2 = B; radd2 = radd; }
    tmem_ld_wait();
    char* op = reinterpret_cast<char*>(obase) + (long long)ch * ost;
    if (rawcopy) {
      char* rp = reinterpret_cast<char*>(rcbase) + (long long)ch * ost;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        *reinterpret_cast<__nv_bfloat16*>(rp + (long long)j * ost) =
            __float2bfloat16_rn(__uint_as_float(v[j]) + (j < 16 ? radd : radd2));
    }
    if (silu) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        *reinterpret_cast<__nv_bfloat16*>(op + (long long)j * ost) = __float2bfloat16_rn(
            epi_silu_tanh(fmaf(__uint_as_float(v[j]) + (j < 16 ? radd : radd2), j < 16 ? A : A2, j < 16 ? B : B2)));
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        *reinterpret_cast<__nv_bfloat16*>(op + (long long)j * ost) =
            __float2bfloat16_rn(fmaf(__uint_as_float(v[j]) + (j < 16 ? radd : radd2), j < 16 ? A : A2, j < 16 ? B : B2));
    }
  }
}

}  // namespace b200
