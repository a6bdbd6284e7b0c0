// Generic tcgen05 GEMM with selectable operand majorness, used by the backward pass:
//   * weight gradients of the convolutions (b200_conv2d_wgrad): dW[co][tap][ci] = sum over pixels of
//     dY[pixel][co] * X[pixel + tap][ci].  The contraction runs over PIXELS, which are the slow dimension of the
//     NHWC tensors, so both operands are fed to the tensor core "MN-major" (UMMA a_major = b_major = 1): a TMA box
//     of (64 channels, 64 pixels) lands in shared memory as [pixel][128 B] and is consumed without any transpose.
//     The tap shift and the zero padding are, as in the forward kernel, the TMA box origin and its out-of-bounds fill.
//   * the batched matrix products of the attention backward and of the embedding-projection backward
//     (b200_gemm_batched), in all four major combinations (A K-major / MN-major x B K-major / MN-major).
// Tile: 128 (M) x BN <= 256 (N) fp32 accumulators in TMEM, K-blocks of 64, TMA ring as deep as shared memory allows, split-K with fp32
// atomics.  Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2-5 epilogue (one TMEM lane quarter each).
#include "common.cuh"
#include <string.h>
#include <stdlib.h>
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;
int make_a_map(CUtensorMap* m, const void* base, int C, int H, int W, int planes, int B, int bw, int bh, int bn);

struct GemmOperandK {
  int mn_major;        // 0: K contiguous, 1: M/N contiguous
  int c_base, c_head;  // inner-dimension (column) origin: c_base + (g % heads) * c_head
  int b_mul, h_mul;    // batch coordinate = (g / heads) * b_mul + (g % heads) * h_mul
};

struct GemmKParams {
  GemmOperandK a, b;
  int conv;                          // 1: K-blocks are pixel tiles of a convolution (wgrad)
  int BN;                            // N tile
  int m_tiles, n_tiles, groups, heads, ksplit, kblocks, stages, k_rotate;
  int kpx;                           // K extent of one stage: 64, or 128 pixels in conv mode (half as many TMA operations per byte)
  int M, N;                          // valid extents (masking in the epilogue)
  // conv mode: pixel tile = bw x bh x bn = 64 pixels, g = tap
  int bw, bh, bn, tiles_w, tiles_h;
  int8_t taps[9][4];                 // {dw, dh, plane, 0} of the B operand (the activation)
  // epilogue
  void* out;
  int out_bf16, atomic, split_out;
  long long s_batch, s_head, s_m, s_n;
  float alpha;
};

constexpr int kGemmThreads = 192;
constexpr int kGemmMaxStages = 8;   // ring depth = as many (16 KB + BN x 128 B) stages as fit in shared memory

struct __align__(8) GemmBars {
  uint64_t full[kGemmMaxStages], empty[kGemmMaxStages], acc_full;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
               const __grid_constant__ GemmKParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int chunk_bytes = p.kpx * 128;                  // one 64-channel slab of kpx K-rows
  const int a_bytes = 2 * chunk_bytes;
  const int stage_bytes = a_bytes + p.BN * p.kpx * 2;
  GemmBars* bars = reinterpret_cast<GemmBars*>(smem + (size_t)p.stages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile decode: blockIdx.x = ((g * ksplit + ks) * m_tiles + mt) * n_tiles + nt
  int idx = blockIdx.x;
  const int nt = idx % p.n_tiles; idx /= p.n_tiles;
  const int mt = idx % p.m_tiles; idx /= p.m_tiles;
  const int ks = idx % p.ksplit;
  const int g = idx / p.ksplit;
  const int gb = g / p.heads, gh = g - gb * p.heads;
  const int per = (p.kblocks + p.ksplit - 1) / p.ksplit;
  const int kb0 = ks * per;
  const int kb1 = min(p.kblocks, kb0 + per);
  const int nkb = kb1 - kb0;
  uint32_t tmem_cols = 32;
  while (tmem_cols < (uint32_t)p.BN) tmem_cols <<= 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bars->tmem_base;

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const int m0 = mt * 128, n0 = nt * p.BN;
        const int ca = p.a.c_base + gh * p.a.c_head, cb = p.b.c_base + gh * p.b.c_head;
        const int ba = gb * p.a.b_mul + gh * p.a.h_mul, bb = gb * p.b.b_mul + gh * p.b.h_mul;
        int stage = 0;
        uint32_t phase = 0;
        // conv mode: the CTAs of the 9 taps of one pixel range would otherwise request the same dY / activation lines
        // from L2 at the same moment; each tap starts at a different K-block of the range (the sum is order-free)
        const int rot = (p.conv && p.k_rotate) ? (int)(((long)g * nkb) / p.groups) : 0;
        // conv mode: pixel-tile coordinates of the K-block, stepped incrementally (three integer divisions per K-block
        // sat between the slot wait and the TMA issue of this single producer thread)
        int tw = 0, th = 0, tn = 0;
        bool recompute = true;
        for (int kbi = 0; kbi < nkb; ++kbi) {
          int kb = kb0 + kbi + rot;
          if (kb >= kb1) kb -= nkb;
          if (p.conv) {
            if (recompute || kb == kb0) {
              tw = kb % p.tiles_w;
              th = (kb / p.tiles_w) % p.tiles_h;
              tn = kb / (p.tiles_w * p.tiles_h);
              recompute = false;
            } else if (++tw == p.tiles_w) {
              tw = 0;
              if (++th == p.tiles_h) { th = 0; ++tn; }
            }
          }
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          uint8_t* sA = smem + (size_t)stage * stage_bytes;
          uint8_t* sB = sA + a_bytes;
          mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
          if (p.conv) {
            // K-block = pixel tile (tw, th, tn); A = dY (unshifted), B = activation shifted by the tap g
            const int w0 = tw * p.bw, h0 = th * p.bh, i0 = tn * p.bn;
            for (int c = 0; c < 2; ++c)
              tma_load_5d(sA + c * chunk_bytes, &mapA, &bars->full[stage], ca + m0 + c * 64, w0, h0, 0, i0);
            for (int c = 0; c < p.BN / 64; ++c)
              tma_load_5d(sB + c * chunk_bytes, &mapB, &bars->full[stage], cb + n0 + c * 64, w0 + p.taps[g][0],
                          h0 + p.taps[g][1], p.taps[g][2], i0);
          } else {
            const int k0 = kb * 64;
            if (p.a.mn_major) {
              for (int c = 0; c < 2; ++c)
                tma_load_5d(sA + c * 8192, &mapA, &bars->full[stage], ca + m0 + c * 64, k0, 0, 0, ba);
            } else {
              tma_load_5d(sA, &mapA, &bars->full[stage], ca + k0, m0, 0, 0, ba);
            }
            if (p.b.mn_major) {
              for (int c = 0; c < p.BN / 64; ++c)
                tma_load_5d(sB + c * 8192, &mapB, &bars->full[stage], cb + n0 + c * 64, k0, 0, 0, bb);
            } else {
              tma_load_5d(sB, &mapB, &bars->full[stage], cb + k0, n0, 0, 0, bb);
            }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        uint32_t idesc = umma_idesc_bf16_m128((uint32_t)p.BN);
        if (p.a.mn_major) idesc |= 1u << 15;
        if (p.b.mn_major) idesc |= 1u << 16;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + (uint32_t)a_bytes;
          const uint64_t adesc = p.a.mn_major ? umma_desc_mnmajor_sw128(a_addr, (uint32_t)chunk_bytes) : umma_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = p.b.mn_major ? umma_desc_mnmajor_sw128(b_addr, (uint32_t)chunk_bytes) : umma_desc_kmajor_sw128(b_addr);
          // 16 k per instruction: K-major advances 32 B inside the swizzle row, MN-major advances 16 rows = 2048 B
          const uint64_t astep = p.a.mn_major ? 128u : 2u;
          const uint64_t bstep = p.b.mn_major ? 128u : 2u;
          const int ksteps = p.kpx >> 4;
#pragma unroll 4
          for (int k = 0; k < ksteps; ++k)
            umma_bf16(tmem_d, adesc + astep * (uint64_t)k, bdesc + bstep * (uint64_t)k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&bars->acc_full);
      }
    } else {
      // ================================ epilogue ================================
      const int q = warp & 3;
      const int m = mt * 128 + q * 32 + lane;
      const bool m_ok = m < p.M;
      mbar_wait(&bars->acc_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
      const long long base = (long long)(p.split_out ? ks : gb) * p.s_batch + (long long)gh * p.s_head + (long long)m * p.s_m;
      for (int c0 = 0; c0 < p.BN; c0 += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_x32(taddr + (uint32_t)c0, v);
        tmem_ld_wait();
        const int n0 = nt * p.BN + c0;
        if (!m_ok || n0 >= p.N) continue;
        if (p.out_bf16) {
          __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + base + (long long)n0 * p.s_n;
          if (p.s_n == 1 && n0 + 32 <= p.N && (((uintptr_t)o) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              uint4 u;
              u.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) * p.alpha, __uint_as_float(v[8 * i + 1]) * p.alpha);
              u.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * p.alpha, __uint_as_float(v[8 * i + 3]) * p.alpha);
              u.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * p.alpha, __uint_as_float(v[8 * i + 5]) * p.alpha);
              u.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * p.alpha, __uint_as_float(v[8 * i + 7]) * p.alpha);
              reinterpret_cast<uint4*>(o)[i] = u;
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) o[(long long)j * p.s_n] = __float2bfloat16_rn(__uint_as_float(v[j]) * p.alpha);
          }
        } else {
          float* o = reinterpret_cast<float*>(p.out) + base + (long long)n0 * p.s_n;
          if (p.atomic) {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) atomicAdd(o + (long long)j * p.s_n, __uint_as_float(v[j]) * p.alpha);
          } else if (p.s_n == 1 && n0 + 32 <= p.N && (((uintptr_t)o) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              reinterpret_cast<float4*>(o)[i] =
                  make_float4(__uint_as_float(v[4 * i]) * p.alpha, __uint_as_float(v[4 * i + 1]) * p.alpha,
                              __uint_as_float(v[4 * i + 2]) * p.alpha, __uint_as_float(v[4 * i + 3]) * p.alpha);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) o[(long long)j * p.s_n] = __uint_as_float(v[j]) * p.alpha;
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_d, tmem_cols);
  }
}

// ================================================================================================
// 3x3 stride-1 weight gradient with VERTICAL TAP FUSION: a CTA owns one horizontal tap offset dx and accumulates the
// three taps dy = -1, 0, +1 in three TMEM accumulators.  Per K-block (bw x bh pixels of one image) it loads the dY
// tile once and ONE activation tile with a one-row halo above and below ((bh + 2) x bw pixels, TMA zero fill = the
// conv padding); the three taps are three UMMA views of that tile whose start addresses differ by dy * bw rows of
// 128 B -- a multiple of the 1024-byte swizzle repeat for bw >= 8, so plain descriptor offsets suffice.
// 4 TMA operations then feed 3x the tensor-core work of the tap-per-CTA form, which was bound by the TMA operation rate.
// ================================================================================================
struct Wgrad3Params {
  int BN;                        // ci tile (<= 128), multiple of 64
  int m_tiles, n_tiles, ksplit, kblocks, stages;
  int bw, bh, tiles_w, tiles_h;  // K-block = bw x bh pixels of one image
  int dy_c0, x_c0, M, N;
  float* out;                    // partials [split][tap = (dy+1)*3 + (dx+1)][co][ci]
  long long s_split, s_tap;
};

__global__ void __launch_bounds__(kGemmThreads, 1)
wgrad3_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
              const __grid_constant__ Wgrad3Params p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int kp = p.bw * p.bh;                          // pixels per K-block
  const int a_slab = kp * 128;                         // dY: [kp rows][128 B] per 64-channel slab
  const int b_slab = (p.bh + 2) * p.bw * 128;          // activation with halo rows
  const int a_bytes = 2 * a_slab;
  const int stage_bytes = a_bytes + (p.BN / 64) * b_slab;
  GemmBars* bars = reinterpret_cast<GemmBars*>(smem + (size_t)p.stages * stage_bytes);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // blockIdx.x = ((dxi * ksplit + ks) * m_tiles + mt) * n_tiles + nt
  int idx = blockIdx.x;
  const int nt = idx % p.n_tiles; idx /= p.n_tiles;
  const int mt = idx % p.m_tiles; idx /= p.m_tiles;
  const int ks = idx % p.ksplit;
  const int dxi = idx / p.ksplit;                      // 0..2  ->  dx = dxi - 1
  const int per = (p.kblocks + p.ksplit - 1) / p.ksplit;
  const int kb0 = ks * per;
  const int kb1 = min(p.kblocks, kb0 + per);
  const int nkb = kb1 - kb0;
  const uint32_t tmem_cols = 512;                      // 3 accumulators of BN <= 128 columns

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA);
    tma_prefetch_desc(&mapB);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = bars->tmem_base;

  if (nkb > 0) {
    if (warp == 0) {
      if (lane == 0) {
        const int m0 = mt * 128, n0 = nt * p.BN;
        int stage = 0;
        uint32_t phase = 0;
        // pixel-tile coordinates of the K-block, stepped incrementally (no integer divisions inside the loop)
        int tw = kb0 % p.tiles_w, th = (kb0 / p.tiles_w) % p.tiles_h, tn = kb0 / (p.tiles_w * p.tiles_h);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          uint8_t* sA = smem + (size_t)stage * stage_bytes;
          uint8_t* sB = sA + a_bytes;
          mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
          const int w0 = tw * p.bw, h0 = th * p.bh;
          for (int c = 0; c < 2; ++c)
            tma_load_5d(sA + c * a_slab, &mapA, &bars->full[stage], p.dy_c0 + m0 + c * 64, w0, h0, 0, tn);
          for (int c = 0; c < p.BN / 64; ++c)
            tma_load_5d(sB + c * b_slab, &mapB, &bars->full[stage], p.x_c0 + n0 + c * 64, w0 + dxi - 1, h0 - 1, 0, tn);
          if (++tw == p.tiles_w) {
            tw = 0;
            if (++th == p.tiles_h) { th = 0; ++tn; }
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    } else if (warp == 1) {
      if (lane == 0) {
        const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)p.BN) | (1u << 15) | (1u << 16);   // both MN-major
        const uint32_t dy_bytes = (uint32_t)p.bw * 128u;       // one image row of the tile
        const int ksteps = kp >> 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint32_t b_addr = a_addr + (uint32_t)a_bytes;
          const uint64_t adesc = umma_desc_mnmajor_sw128(a_addr, (uint32_t)a_slab);
#pragma unroll
          for (int dyi = 0; dyi < 3; ++dyi) {
            const uint64_t bdesc = umma_desc_mnmajor_sw128(b_addr + (uint32_t)dyi * dy_bytes, (uint32_t)b_slab);
            const uint32_t td = tmem_d + (uint32_t)(dyi * p.BN);
            for (int k = 0; k < ksteps; ++k)   // 16 K-rows = 2048 B per step
              umma_bf16(td, adesc + 128ull * (uint64_t)k, bdesc + 128ull * (uint64_t)k, idesc, (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&bars->acc_full);
      }
    } else {
      const int q = warp & 3;
      const int m = mt * 128 + q * 32 + lane;
      const bool m_ok = m < p.M;
      mbar_wait(&bars->acc_full, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
      for (int dyi = 0; dyi < 3; ++dyi) {
        float* obase = p.out + (long long)ks * p.s_split + (long long)(dyi * 3 + dxi) * p.s_tap + (long long)m * p.N;
        for (int c0 = 0; c0 < p.BN; c0 += 32) {
          uint32_t v[32];
          __syncwarp();
          tmem_ld_x32(taddr + (uint32_t)(dyi * p.BN + c0), v);
          tmem_ld_wait();
          const int n0 = nt * p.BN + c0;
          if (!m_ok || n0 >= p.N) continue;
          float* o = obase + n0;
          if (n0 + 32 <= p.N && (((uintptr_t)o) & 15) == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              reinterpret_cast<float4*>(o)[i] = make_float4(__uint_as_float(v[4 * i]), __uint_as_float(v[4 * i + 1]),
                                                            __uint_as_float(v[4 * i + 2]), __uint_as_float(v[4 * i + 3]));
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
              if (n0 + j < p.N) o[j] = __uint_as_float(v[j]);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_d, tmem_cols);
  }
}

static int launch_gemm(const CUtensorMap& mapA, const CUtensorMap& mapB, const GemmKParams& p_in, cudaStream_t stream) {
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  GemmKParams p = p_in;
  // as many stages as fit: with 4 stages of 32 KB (BN = 128) only 128 KB were in flight and the K loop ran at the
  // TMA round-trip latency, not at the tensor-core rate
  static const char* env_st = getenv("B200_GEMM_STAGES");
  if (p.kpx == 0) p.kpx = 64;
  const int stage_bytes = 2 * p.kpx * 128 + p.BN * p.kpx * 2;
  int stages = 4;   // measured: deeper rings (up to 7 x 32 KB) are not faster -- the K loop is not latency-bound
  const int fit = (int)((227 * 1024 - 2048 - sizeof(GemmBars)) / stage_bytes);
  if (stages > kGemmMaxStages) stages = kGemmMaxStages;
  if (env_st && atoi(env_st) >= 2) stages = atoi(env_st);
  if (stages > fit) stages = fit;
  p.stages = stages;
  static const char* env_rot = getenv("B200_WGRAD_ROTATE");
  p.k_rotate = (env_rot && atoi(env_rot) == 0) ? 0 : 1;
  const size_t smem = (size_t)stages * stage_bytes + sizeof(GemmBars) + 1024;
  const long long grid = (long long)p.groups * p.ksplit * p.m_tiles * p.n_tiles;
  B200_REQUIRE(grid >= 1 && grid < (1ll << 31), "gemm: bad grid %lld", grid);
  gemm_tc_kernel<<<(unsigned)grid, kGemmThreads, smem, stream>>>(mapA, mapB, p);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "gemm_tc_kernel launch");
}

// dw[co][ci][tap] (arbitrary strides) += sum over splits of scratch[split][tap][co][ci]: the second phase of the
// weight gradient when the pixel range is split over CTAs (plain coalesced partial stores instead of atomics)
template <int NT>   // taps known at compile time (9, 4, 1): NT accumulators, NT x 4 independent loads per round
__global__ void __launch_bounds__(256) wgrad_reduce_nt_kernel(const float* __restrict__ scratch, float* __restrict__ dw,
                                                              int splits, int Cout, int Cin, long long s_co,
                                                              long long s_ci, long long s_tap) {
  // a thread owns one (co, ci) and writes its NT consecutive outputs (a tap-parallel grid was measured 1.4-2x
  // slower: nine CTAs then read-modify-write interleaved 4-byte words of the same sectors).  With one load chain
  // per tap the kernel was pure latency (72 dependent-ish loads = 12.6 us for 256x256x9 over 8 splits).
  const long long per_tap = (long long)Cout * Cin;
  const long long per_split = per_tap * NT;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_tap; i += (long long)gridDim.x * blockDim.x) {
    float acc[NT];
#pragma unroll
    for (int t = 0; t < NT; ++t) acc[t] = 0.f;
    const float* src = scratch + i;
    int sp = 0;
    for (; sp + 4 <= splits; sp += 4) {
      float v[4][NT];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < NT; ++t) v[u][t] = __ldg(src + (long long)(sp + u) * per_split + (long long)t * per_tap);
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int t = 0; t < NT; ++t) acc[t] += v[u][t];
    }
    for (; sp < splits; ++sp) {
#pragma unroll
      for (int t = 0; t < NT; ++t) acc[t] += __ldg(src + (long long)sp * per_split + (long long)t * per_tap);
    }
    const int co = (int)(i / Cin), ci = (int)(i - (long long)co * Cin);
    float* o = dw + co * s_co + ci * s_ci;
#pragma unroll
    for (int t = 0; t < NT; ++t) o[t * s_tap] += acc[t];
  }
}

__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ scratch, float* __restrict__ dw,
                                                           int splits, int ntaps, int Cout, int Cin, long long s_co,
                                                           long long s_ci, long long s_tap) {
  const long long per_tap = (long long)Cout * Cin;
  const long long per_split = per_tap * ntaps;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < per_tap; i += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(i / Cin), ci = (int)(i - (long long)co * Cin);
    float* o = dw + co * s_co + ci * s_ci;
    for (int t = 0; t < ntaps; ++t) {
      const float* src = scratch + (long long)t * per_tap + i;
      float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
      int sp = 0;
      for (; sp + 4 <= splits; sp += 4) {
        a0 += __ldg(src + (long long)sp * per_split);
        a1 += __ldg(src + (long long)(sp + 1) * per_split);
        a2 += __ldg(src + (long long)(sp + 2) * per_split);
        a3 += __ldg(src + (long long)(sp + 3) * per_split);
      }
      for (; sp < splits; ++sp) a0 += __ldg(src + (long long)sp * per_split);
      o[t * s_tap] += (a0 + a1) + (a2 + a3);
    }
  }
}

// Second phase of the split-K weight gradient, parallel over splits: a block of 256 threads = 32 consecutive (co, ci)
// entries x 8 split groups; a thread sums its share of the splits for all taps (ntaps independent accumulators),
// the 8 groups are combined in shared memory and the 32 x ntaps results are written with consecutive threads on
// consecutive (entry, tap) pairs.  (One thread per entry walking every split serially took 12 us for 10 MB and
// 45 us once the tap-fused kernel split the pixel range ~50 ways.)
constexpr int kRedE = 32, kRedG = 8;
__global__ void __launch_bounds__(256) wgrad_reduce2_kernel(const float* __restrict__ scratch, float* __restrict__ dw,
                                                            int splits, int ntaps, int Cout, int Cin, long long s_co,
                                                            long long s_ci, long long s_tap) {
  __shared__ float red[kRedG][9][kRedE + 1];
  const long long per_tap = (long long)Cout * Cin;
  const long long per_split = per_tap * ntaps;
  const int e = threadIdx.x % kRedE, g = threadIdx.x / kRedE;
  const long long i = (long long)blockIdx.x * kRedE + e;
  float acc[9];
#pragma unroll
  for (int t = 0; t < 9; ++t) acc[t] = 0.f;
  if (i < per_tap) {
    for (int sp = g; sp < splits; sp += kRedG) {
      const float* src = scratch + (long long)sp * per_split + i;
#pragma unroll
      for (int t = 0; t < 9; ++t)
        if (t < ntaps) acc[t] += __ldg(src + (long long)t * per_tap);
    }
  }
#pragma unroll
  for (int t = 0; t < 9; ++t) red[g][t][e] = acc[t];
  __syncthreads();
  for (int j = threadIdx.x; j < kRedE * ntaps; j += 256) {
    const int ee = j / ntaps, t = j - ee * ntaps;
    const long long ie = (long long)blockIdx.x * kRedE + ee;
    if (ie >= per_tap) continue;
    float v = 0.f;
#pragma unroll
    for (int gg = 0; gg < kRedG; ++gg) v += red[gg][t][ee];
    const int co = (int)(ie / Cin), ci = (int)(ie - (long long)co * Cin);
    dw[co * s_co + ci * s_ci + t * s_tap] += v;
  }
}

static int launch_wgrad_reduce(const float* scratch, float* dw, int splits, int ntaps, int Cout, int Cin, long long s_co,
                               long long s_ci, long long s_tap, cudaStream_t stream) {
  const long long per_tap = (long long)Cout * Cin;
  static const char* env_r = getenv("B200_WGRAD_REDUCE");
  // measured: the split-parallel form wins from ~12 splits on (128->128@32x32: 55 -> 51 us), loses below
  // (512->256@16x16 with 6 splits: 72 -> 80 us)
  // and only helps while one thread per entry leaves the machine empty (<= 32 K entries; 1x1 512->256: 22 -> 38 us)
  if (ntaps <= 9 && splits >= 12 && per_tap <= 32768 && !(env_r && atoi(env_r) == 1)) {
    const long long g = (per_tap + kRedE - 1) / kRedE;
    wgrad_reduce2_kernel<<<(unsigned)g, 256, 0, stream>>>(scratch, dw, splits, ntaps, Cout, Cin, s_co, s_ci, s_tap);
  } else {
    long long g = (per_tap + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    static const char* env_nt = getenv("B200_WGRAD_REDUCE_NT");
    const bool nt_ok = !(env_nt && atoi(env_nt) == 0);
    if (nt_ok && ntaps == 9)
      wgrad_reduce_nt_kernel<9><<<(unsigned)g, 256, 0, stream>>>(scratch, dw, splits, Cout, Cin, s_co, s_ci, s_tap);
    else if (nt_ok && ntaps == 4)
      wgrad_reduce_nt_kernel<4><<<(unsigned)g, 256, 0, stream>>>(scratch, dw, splits, Cout, Cin, s_co, s_ci, s_tap);
    else if (nt_ok && ntaps == 1)
      wgrad_reduce_nt_kernel<1><<<(unsigned)g, 256, 0, stream>>>(scratch, dw, splits, Cout, Cin, s_co, s_ci, s_tap);
    else
      wgrad_reduce_kernel<<<(unsigned)g, 256, 0, stream>>>(scratch, dw, splits, ntaps, Cout, Cin, s_co, s_ci, s_tap);
  }
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "wgrad_reduce launch");
}

static int g_sms = 0;
static int num_sms() {
  if (g_sms == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess)
      g_sms = 148;
  }
  return g_sms;
}

// 5-D map over a plain [batch][rows][ld] bf16 matrix: dims (ld, rows, 1, 1, batch)
static int make_matrix_map(CUtensorMap* m, const b200_gemm_operand* o, int batch, int box_rows) {
  uint64_t dims[5] = {(uint64_t)o->ld, (uint64_t)o->rows, 1, 1, (uint64_t)batch};
  const uint64_t bs = (uint64_t)(o->batch_stride ? o->batch_stride : (long long)o->rows * o->ld) * 2;
  uint64_t strides[4] = {(uint64_t)o->ld * 2, bs, bs, bs};
  uint32_t box[5] = {64, (uint32_t)box_rows, 1, 1, 1};
  return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, o->ptr, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_gemm_batched(const b200_gemm_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d && d->a.ptr && d->b.ptr && d->out, "gemm_batched: null pointer");
  B200_REQUIRE(d->M >= 1 && d->N >= 1 && d->K >= 1 && d->batch >= 1 && d->heads >= 1, "gemm_batched: bad sizes");
  B200_REQUIRE(d->a.ld % 8 == 0 && d->b.ld % 8 == 0, "gemm_batched: leading dimensions must be multiples of 8");
  B200_REQUIRE(((uintptr_t)d->a.ptr & 127) == 0 && ((uintptr_t)d->b.ptr & 127) == 0, "gemm_batched: operand alignment");
  B200_REQUIRE((d->a.batch_stride % 8) == 0 && (d->b.batch_stride % 8) == 0, "gemm_batched: batch strides must be multiples of 8");
  B200_REQUIRE(d->K % 8 == 0, "gemm_batched: K=%d must be a multiple of 8", d->K);
  // A K tail that is not a multiple of 64 relies on TMA zero fill, which only triggers at the tensor edge: the
  // contraction window must then end exactly at the operand's last row / column
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.a.mn_major = d->a.mn_major; p.a.c_base = d->a.col_base; p.a.c_head = d->a.col_head;
  p.b.mn_major = d->b.mn_major; p.b.c_base = d->b.col_base; p.b.c_head = d->b.col_head;
  p.a.b_mul = d->a.per_head_batch ? d->heads : 1; p.a.h_mul = d->a.per_head_batch ? 1 : 0;
  p.b.b_mul = d->b.per_head_batch ? d->heads : 1; p.b.h_mul = d->b.per_head_batch ? 1 : 0;
  if (d->K % 64 != 0) {
    if (d->a.mn_major) B200_REQUIRE(d->a.rows == d->K, "gemm_batched: ragged K needs A rows == K");
    else B200_REQUIRE((d->heads == 1 || d->a.col_head == 0) && d->a.col_base + d->K == d->a.ld, "gemm_batched: ragged K needs the A window at the row end");
    if (d->b.mn_major) B200_REQUIRE(d->b.rows == d->K, "gemm_batched: ragged K needs B rows == K");
    else B200_REQUIRE((d->heads == 1 || d->b.col_head == 0) && d->b.col_base + d->K == d->b.ld, "gemm_batched: ragged K needs the B window at the row end");
  }
  // an MN-major operand is loaded in 64-column chunks: columns past M/N inside the last chunk must either not exist
  // (tensor edge, zero fill) or be harmless: they only produce rows/columns the epilogue masks.
  int BN = d->b.mn_major ? ((d->N + 63) / 64) * 64 : ((d->N + 15) / 16) * 16;
  if (BN > 256) BN = 256;
  p.BN = BN;
  p.m_tiles = (d->M + 127) / 128;
  p.n_tiles = (d->N + BN - 1) / BN;
  p.groups = d->batch * d->heads;
  p.heads = d->heads;
  p.kblocks = (d->K + 63) / 64;
  p.M = d->M; p.N = d->N;
  p.out = d->out; p.out_bf16 = d->out_bf16;
  p.s_batch = d->out_batch_stride; p.s_head = d->out_head_stride; p.s_m = d->out_ld; p.s_n = 1;
  p.alpha = d->alpha == 0.f ? 1.f : d->alpha;
  int ksplit = 1;
  if (d->split_k > 1) {
    B200_REQUIRE(!d->out_bf16, "gemm_batched: split-K needs an fp32 output");
    ksplit = d->split_k < p.kblocks ? d->split_k : p.kblocks;
  }
  p.ksplit = ksplit;
  p.atomic = (d->accumulate || ksplit > 1) ? 1 : 0;
  B200_REQUIRE(!(p.atomic && d->out_bf16), "gemm_batched: accumulate needs an fp32 output");
  CUtensorMap mapA, mapB;
  int rc = make_matrix_map(&mapA, &d->a, d->batch * (d->a.per_head_batch ? d->heads : 1), d->a.mn_major ? 64 : 128);
  if (rc) return rc;
  rc = make_matrix_map(&mapB, &d->b, d->batch * (d->b.per_head_batch ? d->heads : 1), d->b.mn_major ? 64 : BN);
  if (rc) return rc;
  return launch_gemm(mapA, mapB, p, stream);
}

extern "C" int b200_conv2d_wgrad(const b200_wgrad_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d && d->dy && d->x && d->dw, "conv2d_wgrad: null pointer");
  B200_REQUIRE(d->dy_C % 64 == 0 && d->x_C % 64 == 0, "conv2d_wgrad: channel counts (%d, %d) must be multiples of 64", d->dy_C, d->x_C);
  B200_REQUIRE(d->Cout >= 1 && d->dy_c0 >= 0 && d->dy_c0 + d->Cout <= d->dy_C && d->Cin >= 1 && d->Cin <= d->x_C, "conv2d_wgrad: bad Cout/Cin");
  B200_REQUIRE(d->ntaps >= 1 && d->ntaps <= 9, "conv2d_wgrad: ntaps=%d out of range", d->ntaps);
  B200_REQUIRE(((uintptr_t)d->dy & 127) == 0 && ((uintptr_t)d->x & 127) == 0, "conv2d_wgrad: operand alignment");
  // ---- vertical-tap-fused path: plain 3x3 stride-1 pad-1 convolution, two-phase reduction available ----
  {
    static const char* env_v3 = getenv("B200_WGRAD_V3");
    bool plain3x3 = d->ntaps == 9 && d->x_planes == 1 && d->x_H == d->Ho && d->x_W == d->Wo && d->scratch != nullptr &&
                    d->Wo >= 8 && (d->Wo & (d->Wo - 1)) == 0 && (d->Ho & (d->Ho - 1)) == 0;
    for (int k = 0; k < 9 && plain3x3; ++k)
      plain3x3 = d->taps[k][0] == (k % 3) - 1 && d->taps[k][1] == (k / 3) - 1 && d->taps[k][2] == 0;
    if (plain3x3 && !(env_v3 && atoi(env_v3) == 0)) {
      Wgrad3Params q;
      memset(&q, 0, sizeof(q));
      static const char* env_kp = getenv("B200_WGRAD_V3_KP");
      const int kp_want = env_kp ? atoi(env_kp) : 128;   // measured: 128-pixel K-blocks beat 64 on every layer
      q.bw = d->Wo < kp_want ? d->Wo : kp_want;
      q.bh = kp_want / q.bw;
      if (q.bh > d->Ho) q.bh = d->Ho;
      if (q.bh < 1) q.bh = 1;
      q.tiles_w = d->Wo / q.bw; q.tiles_h = d->Ho / q.bh;
      q.kblocks = q.tiles_w * q.tiles_h * d->B;
      const int n_split = (d->Cin + 127) / 128;
      q.BN = (((d->Cin + n_split - 1) / n_split + 63) / 64) * 64;
      q.m_tiles = (d->Cout + 127) / 128;
      q.n_tiles = (d->Cin + q.BN - 1) / q.BN;
      q.dy_c0 = d->dy_c0; q.x_c0 = d->x_c0; q.M = d->Cout; q.N = d->Cin;
      const long tiles3 = 3l * q.m_tiles * q.n_tiles;
      long ks = num_sms() / tiles3;
      if (ks > q.kblocks) ks = q.kblocks;
      if (ks < 1) ks = 1;
      {
        const long per = (q.kblocks + ks - 1) / ks;
        ks = (q.kblocks + per - 1) / per;
      }
      const long long per_split3 = 9ll * d->Cout * d->Cin;
      const int kp = q.bw * q.bh;
      const int stage_bytes = 2 * kp * 128 + (q.BN / 64) * (q.bh + 2) * q.bw * 128;
      int stages = (int)((227 * 1024 - 2048 - sizeof(GemmBars)) / stage_bytes);
      if (stages > kGemmMaxStages) stages = kGemmMaxStages;
      // measured (tools/bench_wgrad.py): with few tiles the pixel range is split ~50 ways and the partial-sum traffic
      // eats the gain (128->128@32x32: 56 -> 70 us), small images have too few K-blocks; otherwise 1.1-1.8x faster
      const bool pays = (tiles3 >= 6 && d->Ho * d->Wo >= 256) || (env_v3 && atoi(env_v3) == 2);
      if (pays && kp % 16 == 0 && stages >= 2 && d->scratch_bytes >= ks * per_split3 * 4 && tiles3 <= num_sms()) {
        q.ksplit = (int)ks; q.stages = stages;
        q.out = reinterpret_cast<float*>(d->scratch);
        q.s_split = per_split3; q.s_tap = (long long)d->Cout * d->Cin;
        CUtensorMap mapA, mapB;
        int rc = make_a_map(&mapA, d->dy, d->dy_C, d->Ho, d->Wo, 1, d->B, q.bw, q.bh, 1);
        if (rc) return rc;
        rc = make_a_map(&mapB, d->x, d->x_C, d->x_H, d->x_W, 1, d->B, q.bw, q.bh + 2, 1);
        if (rc) return rc;
        static bool attr3 = false;
        if (!attr3) {
          B200_CHECK(cudaFuncSetAttribute(wgrad3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
          attr3 = true;
        }
        const size_t smem = (size_t)stages * stage_bytes + sizeof(GemmBars) + 1024;
        const long long grid = 3ll * q.ksplit * q.m_tiles * q.n_tiles;
        wgrad3_kernel<<<(unsigned)grid, kGemmThreads, smem, stream>>>(mapA, mapB, q);
        ++g_launch_count;
        B200_CHECK(cudaGetLastError());
        return launch_wgrad_reduce(reinterpret_cast<const float*>(d->scratch), d->dw, q.ksplit, 9, d->Cout, d->Cin,
                                   d->dw_co_stride, d->dw_ci_stride, d->dw_tap_stride, stream);
      }
    }
  }
  // K-block = kpx pixels per stage: 128 when the tensor has enough pixels for every CTA (one TMA operation then
  // brings 16 KB instead of 8 KB: the 64-pixel form was bound by the TMA operation rate), else 64
  static const char* env_kpx = getenv("B200_WGRAD_KPX");
  int kpx = env_kpx ? atoi(env_kpx) : 128;
  if (kpx != 128 || (long)d->B * d->Ho * d->Wo < 4096) kpx = 64;   // (N tiles wider than 128 then run on 2 stages)
  int bw = 1;
  while (bw * 2 <= kpx && d->Wo % (bw * 2) == 0) bw *= 2;
  int bh = 1;
  while (bw * bh * 2 <= kpx && d->Ho % (bh * 2) == 0) bh *= 2;
  int bn = kpx / (bw * bh);
  if (!(bn == 1 || (bw == d->Wo && bh == d->Ho)) && kpx == 128) {   // fall back to 64-pixel blocks
    kpx = 64;
    bw = 1;
    while (bw * 2 <= kpx && d->Wo % (bw * 2) == 0) bw *= 2;
    bh = 1;
    while (bw * bh * 2 <= kpx && d->Ho % (bh * 2) == 0) bh *= 2;
    bn = kpx / (bw * bh);
  }
  B200_REQUIRE(bn == 1 || (bw == d->Wo && bh == d->Ho), "conv2d_wgrad: unsupported spatial size %dx%d", d->Ho, d->Wo);
  GemmKParams p;
  memset(&p, 0, sizeof(p));
  p.kpx = kpx;
  p.conv = 1;
  p.a.mn_major = 1; p.b.mn_major = 1;
  p.a.c_base = d->dy_c0; p.b.c_base = d->x_c0;
  // N tile: Cin split evenly over ceil(Cin / 256) tiles (384 -> 2 x 192, not 256 + a half-empty 128)
  const int n_split = (d->Cin + 255) / 256;
  int BN = (((d->Cin + n_split - 1) / n_split + 63) / 64) * 64;
  if (BN > 256) BN = 256;
  p.BN = BN;
  p.m_tiles = (d->Cout + 127) / 128;
  p.n_tiles = (d->Cin + BN - 1) / BN;
  p.groups = d->ntaps;
  p.heads = d->ntaps;      // g = tap: gb = 0, gh = tap
  p.bw = bw; p.bh = bh; p.bn = bn;
  p.tiles_w = d->Wo / bw; p.tiles_h = d->Ho / bh;
  p.kblocks = p.tiles_w * p.tiles_h * ((d->B + bn - 1) / bn);
  p.M = d->Cout; p.N = d->Cin;
  memcpy(p.taps, d->taps, sizeof(p.taps));
  p.alpha = 1.f;
  // split the pixel range so that the grid fills the machine (one wave of CTAs)
  // (one CTA per SM: rounding the split count UP gave e.g. 9 taps x 17 splits = 153 CTAs on 148 SMs, i.e. a second
  // wave of 5 CTAs that doubled the kernel's duration; round DOWN so that the grid never exceeds the SM count)
  const long tiles = (long)p.groups * p.m_tiles * p.n_tiles;
  static const char* env_up = getenv("B200_WGRAD_SPLIT_UP");
  long ks = (env_up && atoi(env_up) == 1) ? (num_sms() + tiles - 1) / tiles : num_sms() / tiles;
  if (ks > p.kblocks) ks = p.kblocks;
  if (ks < 1) ks = 1;
  {   // no empty split: the kernel gives every split ceil(kblocks / ksplit) K-blocks
    const long per = (p.kblocks + ks - 1) / ks;
    ks = (p.kblocks + per - 1) / per;
  }
  const long long per_split = (long long)d->ntaps * d->Cout * d->Cin;
  const bool two_phase = d->scratch != nullptr && d->scratch_bytes >= (long long)ks * per_split * 4;
  p.ksplit = (int)ks;
  if (two_phase) {
    // phase 1: every split stores its partial [tap][co][ci] tile rows with plain vector stores (split = "batch" of
    // the generic epilogue addressing: blockIdx decode gives ks; we fold it into the output pointer per CTA below)
    p.out = d->scratch; p.out_bf16 = 0; p.atomic = 0;
    p.s_batch = per_split;                       // indexed by the split (see kernel: conv mode uses ks as "gb")
    p.s_head = (long long)d->Cout * d->Cin;      // tap
    p.s_m = d->Cin; p.s_n = 1;
    p.split_out = 1;
  } else {
    // single phase: atomics straight into the (strided) gradient tensor; also what lets repeated backward passes
    // accumulate into an existing .grad
    p.out = d->dw; p.out_bf16 = 0; p.atomic = 1;
    p.s_batch = 0; p.s_head = d->dw_tap_stride; p.s_m = d->dw_co_stride; p.s_n = d->dw_ci_stride;
  }
  CUtensorMap mapA, mapB;
  int rc = make_a_map(&mapA, d->dy, d->dy_C, d->Ho, d->Wo, 1, d->B, bw, bh, bn);
  if (rc) return rc;
  rc = make_a_map(&mapB, d->x, d->x_C, d->x_H, d->x_W, d->x_planes, d->B, bw, bh, bn);
  if (rc) return rc;
  rc = launch_gemm(mapA, mapB, p, stream);
  if (rc || !two_phase) return rc;
  return launch_wgrad_reduce(reinterpret_cast<const float*>(d->scratch), d->dw, p.ksplit, d->ntaps, d->Cout, d->Cin,
                             d->dw_co_stride, d->dw_ci_stride, d->dw_tap_stride, stream);
}
