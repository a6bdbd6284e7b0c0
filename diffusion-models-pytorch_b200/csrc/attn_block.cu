// K2b: the whole self-attention block of the CIFAR-10 UNet in ONE kernel (models/modules.py:77-102):
//     out = x + proj(softmax(q k^T d^-1/2) v),   q, k, v = 1x1 convs of GroupNorm(x)
// specialised to T = 256 tokens (16x16), C = 256 channels, one head, 32 groups.  q, k, v, the scores and the attention
// output never leave the SM: per image the kernel reads x once (fp32 residual stream) and writes the block output
// once, 0.5 MB instead of the 2.2 MB the five-launch form (GroupNorm, [q|k] conv, v conv, attention, proj conv) moves.
//
// One 2-CTA cluster (the two SMs of a TPC) per image, persistent over images.  CTA r owns tokens [128 r, 128 r + 128).
// All six GEMMs are tcgen05 cta_group::2 MMAs (M = 256 split over the two CTAs' TMEM, N = 256 with each CTA holding
// half of the B rows in its own shared memory), so every operand half is produced exactly where the MMA wants it and
// no tile is ever copied between the CTAs:
//     Q   [tok x d]   = xn  Wq^T     A = xn   (own tokens),       B = Wq (own 128 rows)      -> TMEM cols [0,256)
//     K   [tok x d]   = xn  Wk^T     A = xn,                      B = Wk (own 128 rows)      -> TMEM cols [256,512)
//     V^T [d x tok]   = Wv  xn^T     A = Wv (own 128 rows),       B = xn (own tokens)        -> TMEM cols [0,256)
//     S   [q x key]   = Q   K^T      A = Q  (own tokens),         B = K  (own keys)          -> TMEM cols [256,512)
//     O   [q x d]     = P   V        A = P  (own queries),        B = V^T (own 128 d rows)   -> TMEM cols [0,256)
//     Y^T [ch x tok]  = Wp  O^T      A = Wp (own 128 rows),       B = O  (own tokens)        -> TMEM cols [256,512)
// The 512 worker threads (16 warps) of each CTA produce the shared-memory operands in between: GroupNorm of the fp32
// tile -> xn (bf16), TMEM -> bf16 drains of Q / K / V^T (+ bias), the row softmax -> P, O / rowsum, and the final
// epilogue (Y^T + bias + residual -> fp32 NHWC, coalesced because TMEM lanes = channels, + GroupNorm statistics of the
// result for the next block).  Three 64 KB operand regions are reused by liveness: R0 = xn -> V^T, R1 = Q -> P,
// R2 = K -> O; weights stream through a 2 x 16 KB TMA ring (each CTA loads its 128 rows of every 64-wide K chunk).
// Warp 16 = TMA producer (both CTAs), warp 17 = MMA issuer (leader CTA only).  Worker -> MMA hand-offs are mbarriers
// in the LEADER's shared memory (32 warp arrivals after a proxy fence), MMA -> worker hand-offs are multicast
// tcgen05.commit arrivals on both CTAs.
#include "pair.cuh"
#include <stdlib.h>
#include <string.h>
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

constexpr int kAbT = 256, kAbC = 256;
constexpr int kAbWorkers = 512;
constexpr int kAbThreads = kAbWorkers + 64;
constexpr uint32_t kAbRegion = 65536;                    // one bf16 [128 rows][256] operand half, 4 chunks of 16 KB
constexpr uint32_t kAbWStage = 16384;                    // [128 rows][64 k] bf16
constexpr int kAbWStages = 2;
constexpr uint32_t kAbBarsOff = 3 * kAbRegion + kAbWStages * kAbWStage;
constexpr uint32_t kAbBiasOff = kAbBarsOff + 256;            // q bias [256] fp32 (the only column-indexed bias: see below)
constexpr size_t kAbSmemBytes = kAbBiasOff + 1024 + 1024;

struct AttnBlockParams {
  const float* x;
  const long long* x_stats;
  const float* gamma;
  const float* beta;
  const float* bias;          // [4][256]: q, k, v, proj
  float* out;
  long long* out_stats;       // may be null
  float eps;
  float scale_log2e;
  int B;
  __nv_bfloat16* dbg[6];      // test builds of the launch: xn, q, k, v^T, p, o as bf16 [B][256][256] (null = off)
};

struct __align__(8) AbBars {
  uint64_t w_full[kAbWStages], w_empty[kAbWStages];
  uint64_t q_done, k_done, v_done, s_done, o_done, y_done;                              // multicast commits
  uint64_t xn_ready, q_drained, k_drained, v_drained, p_ready, o_drained, y_drained;    // leader's copies, 32 warps
  uint32_t tmem_base;
};

__device__ __forceinline__ void ab_workers_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kAbWorkers) : "memory"); }

__device__ __forceinline__ float4 ab_ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// One warp's share of a TMEM [128 lanes x 256 columns] fp32 block -> bf16 K-major swizzled operand half in `region`
// (row = TMEM lane, 64-column chunk `part`): value * row_scale + col_bias[column] + row_bias.
template <bool kDbg>
__device__ __forceinline__ void ab_drain(uint32_t taddr, uint8_t* region, int r, int part, const float* col_bias,
                                         float row_bias, float row_scale, __nv_bfloat16* dbg_row) {
  uint8_t* rowp = region + (size_t)part * 16384 + (size_t)r * 128;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int c0 = part * 64 + h * 32;
    uint32_t v[32];
    tmem_ld_x32(taddr + (uint32_t)c0, v);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float f[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] = __uint_as_float(v[8 * i + j]) * row_scale + row_bias;
      if (col_bias != nullptr) {      // shared memory (warp-uniform address: broadcast)
        const float4 b0 = *reinterpret_cast<const float4*>(col_bias + c0 + 8 * i);
        const float4 b1 = *reinterpret_cast<const float4*>(col_bias + c0 + 8 * i + 4);
        f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
        f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
      }
      uint4 u;
      u.x = pack_bf16x2(f[0], f[1]);
      u.y = pack_bf16x2(f[2], f[3]);
      u.z = pack_bf16x2(f[4], f[5]);
      u.w = pack_bf16x2(f[6], f[7]);
      *reinterpret_cast<uint4*>(rowp + (((h * 4 + i) ^ (r & 7)) << 4)) = u;
      if (kDbg && dbg_row != nullptr) *reinterpret_cast<uint4*>(dbg_row + c0 + 8 * i) = u;
    }
  }
}

template <bool kDbg>
__global__ void __launch_bounds__(kAbThreads, 1)
attn_block_pair_kernel(const __grid_constant__ CUtensorMap mapW, const __grid_constant__ AttnBlockParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* R0 = smem;
  uint8_t* R1 = smem + kAbRegion;
  uint8_t* R2 = smem + 2 * kAbRegion;
  uint8_t* WS = smem + 3 * kAbRegion;
  AbBars* bars = reinterpret_cast<AbBars*>(smem + kAbBarsOff);
  float* s_bq = reinterpret_cast<float*>(smem + kAbBiasOff);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;

  if (warp == 16 && lane == 0) tma_prefetch_desc(&mapW);
  if (warp == 17 && lane == 0) {
    for (int s = 0; s < kAbWStages; ++s) {
      mbar_init(&bars->w_full[s], 1);
      mbar_init(&bars->w_empty[s], 1);
    }
    mbar_init(&bars->q_done, 1);
    mbar_init(&bars->k_done, 1);
    mbar_init(&bars->v_done, 1);
    mbar_init(&bars->s_done, 1);
    mbar_init(&bars->o_done, 1);
    mbar_init(&bars->y_done, 1);
    const uint32_t nw = 2 * (kAbWorkers / 32);
    mbar_init(&bars->xn_ready, nw);
    mbar_init(&bars->q_drained, nw);
    mbar_init(&bars->k_drained, nw);
    mbar_init(&bars->v_drained, nw);
    mbar_init(&bars->p_ready, nw);
    mbar_init(&bars->o_drained, nw);
    mbar_init(&bars->y_drained, nw);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(&bars->tmem_base, 512);
  if (threadIdx.x < kAbC) s_bq[threadIdx.x] = __ldg(p.bias + threadIdx.x);     // parameters: not written by the preceding grid
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  griddep_sync();

  if (warp == 16) {
    // ================================ weight TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int b = cluster_id; b < p.B; b += n_clusters) {
        for (int mat = 0; mat < 4; ++mat) {
          for (int kc = 0; kc < 4; ++kc) {
            mbar_wait(&bars->w_empty[stage], phase ^ 1u);
            if (rank == 0) mbar_arrive_expect_tx(&bars->w_full[stage], 2u * kAbWStage);
            tma_load_2d_2sm(WS + (size_t)stage * kAbWStage, &mapW, &bars->w_full[stage], kc * 64, mat * 256 + rank * 128);
            if (++stage == kAbWStages) { stage = 0; phase ^= 1u; }
          }
        }
      }
    }
  } else if (warp == 17) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16_m256(256u);
      const uint32_t tm0 = tmem_base, tm1 = tmem_base + 256u;
      const uint32_t r0 = smem_u32(R0), r1 = smem_u32(R1), r2 = smem_u32(R2), ws = smem_u32(WS);
      int stage = 0;
      uint32_t phase = 0;
      // one operand streams through the weight ring: w_is_a selects which side
      auto gemm_w = [&](uint32_t tmem_d, uint32_t act, bool w_is_a) {
        for (int kc = 0; kc < 4; ++kc) {
          mbar_wait(&bars->w_full[stage], phase);
          tc_fence_after();
          const uint64_t wdesc = umma_desc_kmajor_sw128(ws + (uint32_t)stage * kAbWStage);
          const uint64_t xdesc = umma_desc_kmajor_sw128(act + (uint32_t)kc * 16384u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_2sm(tmem_d, (w_is_a ? wdesc : xdesc) + (uint64_t)(2 * k), (w_is_a ? xdesc : wdesc) + (uint64_t)(2 * k),
                          idesc, (kc > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&bars->w_empty[stage]);
          if (++stage == kAbWStages) { stage = 0; phase ^= 1u; }
        }
      };
      auto gemm_ss = [&](uint32_t tmem_d, uint32_t a, uint32_t bop) {
        for (int kc = 0; kc < 4; ++kc) {
          const uint64_t adesc = umma_desc_kmajor_sw128(a + (uint32_t)kc * 16384u);
          const uint64_t bdesc = umma_desc_kmajor_sw128(bop + (uint32_t)kc * 16384u);
#pragma unroll
          for (int k = 0; k < 4; ++k)
            umma_bf16_2sm(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc, (kc > 0 || k > 0) ? 1u : 0u);
        }
      };
      int it = 0;
      for (int b = cluster_id; b < p.B; b += n_clusters, ++it) {
        const uint32_t ph = (uint32_t)it & 1u;
        mbar_wait_cluster(&bars->xn_ready, ph);
        tc_fence_after();
        gemm_w(tm0, r0, false);                       // Q = xn Wq^T   (TMEM [0,256) was drained: o_drained of it-1)
        umma_commit_2sm(&bars->q_done);
        if (it > 0) {
          mbar_wait_cluster(&bars->y_drained, ph ^ 1u);
          tc_fence_after();
        }
        gemm_w(tm1, r0, false);                       // K = xn Wk^T
        umma_commit_2sm(&bars->k_done);
        mbar_wait_cluster(&bars->q_drained, ph);
        tc_fence_after();
        gemm_w(tm0, r0, true);                        // V^T = Wv xn^T
        umma_commit_2sm(&bars->v_done);
        mbar_wait_cluster(&bars->k_drained, ph);
        tc_fence_after();
        gemm_ss(tm1, r1, r2);                         // S = Q K^T
        umma_commit_2sm(&bars->s_done);
        mbar_wait_cluster(&bars->v_drained, ph);
        mbar_wait_cluster(&bars->p_ready, ph);
        tc_fence_after();
        gemm_ss(tm0, r1, r0);                         // O = P V
        umma_commit_2sm(&bars->o_done);
        mbar_wait_cluster(&bars->o_drained, ph);
        tc_fence_after();
        gemm_w(tm1, r2, true);                        // Y^T = Wp O^T
        umma_commit_2sm(&bars->y_done);
      }
    }
  } else {
    // ================================ workers (both CTAs) ================================
    const int wq = warp & 3, part = warp >> 2;
    const int r = wq * 32 + lane;                       // TMEM lane = row of the own 128-row half
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    const uint32_t tm0 = tmem_base + lane_base, tm1 = tmem_base + 256u + lane_base;
    float* red_max = reinterpret_cast<float*>(R2);      // [4][128], alive between s_done and p_ready (K is dead)
    float* red_sum = red_max + 4 * 128;

    // GroupNorm of the own 128 tokens of image b -> xn (bf16, R0).  A warp takes 8 token rows; lane = 8-channel group
    // (= one GroupNorm group).  The fp32 loads are issued in two halves of 4 rows so that the first half can be in
    // flight across the wait for the P V MMA and the O drain (R0 must not be written before o_done).
    auto gn_issue = [&](int b, int i0, float4 (&raw)[8]) {
      const float* xb = p.x + ((size_t)b * kAbT + rank * 128 + warp * 8 + i0) * kAbC + 8 * lane;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        raw[2 * i] = ab_ldg4(xb + (size_t)i * kAbC);
        raw[2 * i + 1] = ab_ldg4(xb + (size_t)i * kAbC + 4);
      }
    };
    // (mean, rstd) of this lane's group; fetched an image ahead so that the statistics' round trip is off the critical path
    auto gn_moments = [&](int b) {
      const float2 gs = stat_load_group(p.x_stats + ((size_t)b * kAbC + 8 * lane) * 2, 8);
      const float inv_cnt = 1.0f / (float)(8 * kAbT);
      const float mean = gs.x * inv_cnt;
      return make_float2(mean, rsqrtf(fmaxf(gs.y * inv_cnt - mean * mean, 0.f) + p.eps));
    };
    auto gn_finish = [&](int b, float4 (&rawA)[8], float4 (&rawB)[8], const float2 mom) {
      const float mean = mom.x, rstd = mom.y;
      float a[8], bb[8];
      {
        const float4 g0 = ab_ldg4(p.gamma + 8 * lane), g1 = ab_ldg4(p.gamma + 8 * lane + 4);
        const float4 b0 = ab_ldg4(p.beta + 8 * lane), b1 = ab_ldg4(p.beta + 8 * lane + 4);
        const float gv[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
        const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int j = 0; j < 8; ++j) { a[j] = rstd * gv[j]; bb[j] = bv[j] - mean * rstd * gv[j]; }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = warp * 8 + i;
        const float4 v0 = i < 4 ? rawA[2 * i] : rawB[2 * (i - 4)], v1 = i < 4 ? rawA[2 * i + 1] : rawB[2 * (i - 4) + 1];
        uint4 u;
        u.x = pack_bf16x2(fmaf(v0.x, a[0], bb[0]), fmaf(v0.y, a[1], bb[1]));
        u.y = pack_bf16x2(fmaf(v0.z, a[2], bb[2]), fmaf(v0.w, a[3], bb[3]));
        u.z = pack_bf16x2(fmaf(v1.x, a[4], bb[4]), fmaf(v1.y, a[5], bb[5]));
        u.w = pack_bf16x2(fmaf(v1.z, a[6], bb[6]), fmaf(v1.w, a[7], bb[7]));
        *reinterpret_cast<uint4*>(R0 + (size_t)(lane >> 3) * 16384 + (size_t)row * 128 + (((lane & 7) ^ (row & 7)) << 4)) = u;
        if (kDbg && p.dbg[0] != nullptr)
          *reinterpret_cast<uint4*>(p.dbg[0] + ((size_t)b * kAbT + rank * 128 + row) * kAbC + 8 * lane) = u;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&bars->xn_ready);
    };
    // hand a finished shared-memory operand (and the TMEM block it was read from) to the MMA issuer: the proxy fence
    // orders this thread's shared-memory writes before the tensor core's (async-proxy) reads, the arrival on the leader's
    // barrier is observed by the MMA thread before it issues (same hand-off as a CUTLASS transform -> UMMA pipeline; a
    // cluster-scope release here costs MEMBAR.GPU + ERRBAR per arrival: 12 % of the kernel in the first version)
    auto publish = [&](uint64_t* bar) {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(bar);
    };
    // L2 prefetch of the own fp32 tile (128 tokens x 1 KB, contiguous) of a later image: 2 lines per thread
    auto prefetch_tile = [&](int b) {
      const char* base = reinterpret_cast<const char*>(p.x + ((size_t)b * kAbT + rank * 128) * kAbC);
      const int tid = threadIdx.x;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)tid * 128));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(base + (size_t)(tid + kAbWorkers) * 128));
    };

    int it = 0;
    if (cluster_id < p.B) {
      float4 rawA[8], rawB[8];
      gn_issue(cluster_id, 0, rawA);
      gn_issue(cluster_id, 4, rawB);
      if (cluster_id + n_clusters < p.B) prefetch_tile(cluster_id + n_clusters);
      gn_finish(cluster_id, rawA, rawB, gn_moments(cluster_id));
    }
    for (int b = cluster_id; b < p.B; b += n_clusters, ++it) {
      const uint32_t ph = (uint32_t)it & 1u;
      const size_t row_g = (size_t)b * kAbT + rank * 128 + r;     // this thread's row of the [B][256][256] debug dumps
      const bool has_next = b + n_clusters < p.B;
      if (b + 2 * n_clusters < p.B) prefetch_tile(b + 2 * n_clusters);
      float2 mom_next = make_float2(0.f, 0.f);
      if (has_next) mom_next = gn_moments(b + n_clusters);

      // ---- Q, K: TMEM -> bf16 operands (q + bias) ----
      mbar_wait(&bars->q_done, ph);
      tc_fence_after();
      ab_drain<kDbg>(tm0, R1, r, part, s_bq, 0.f, 1.f, kDbg && p.dbg[1] ? p.dbg[1] + row_g * kAbC : nullptr);
      publish(&bars->q_drained);
      mbar_wait(&bars->k_done, ph);
      tc_fence_after();
      // the k bias adds q_i . b_k to every score of row i: softmax is invariant to it, so it is never applied
      ab_drain<kDbg>(tm1, R2, r, part, nullptr, 0.f, 1.f, kDbg && p.dbg[2] ? p.dbg[2] + row_g * kAbC : nullptr);
      publish(&bars->k_drained);

      // ---- V^T: rows = this CTA's 128 d channels, columns = all keys; overwrites xn ----
      mbar_wait(&bars->v_done, ph);
      tc_fence_after();
      ab_drain<kDbg>(tm0, R0, r, part, nullptr, __ldg(p.bias + 512 + rank * 128 + r), 1.f,
                     kDbg && p.dbg[3] ? p.dbg[3] + row_g * kAbC : nullptr);
      publish(&bars->v_drained);

      // ---- softmax over the own 128 score rows (4 warps per row quarter share a row: 64 keys each) ----
      mbar_wait(&bars->s_done, ph);
      tc_fence_after();
      float m = -INFINITY;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_x32(tm1 + (uint32_t)(part * 64 + h * 32), v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) m = fmaxf(m, __uint_as_float(v[j]));
      }
      red_max[part * 128 + r] = m;
      ab_workers_sync();
#pragma unroll
      for (int i = 0; i < 4; ++i) m = fmaxf(m, red_max[i * 128 + r]);
      const float ms = m * p.scale_log2e;
      float sum = 0.f;
      {
        uint8_t* prow = R1 + (size_t)part * 16384 + (size_t)r * 128;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_x32(tm1 + (uint32_t)(part * 64 + h * 32), v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              e[j] = exp2f(fmaf(__uint_as_float(v[8 * i + j]), p.scale_log2e, -ms));
              sum += e[j];
            }
            uint4 u;
            u.x = pack_bf16x2(e[0], e[1]);
            u.y = pack_bf16x2(e[2], e[3]);
            u.z = pack_bf16x2(e[4], e[5]);
            u.w = pack_bf16x2(e[6], e[7]);
            *reinterpret_cast<uint4*>(prow + (((h * 4 + i) ^ (r & 7)) << 4)) = u;
            if (kDbg && p.dbg[4] != nullptr)
              *reinterpret_cast<uint4*>(p.dbg[4] + row_g * kAbC + part * 64 + h * 32 + 8 * i) = u;
          }
        }
      }
      red_sum[part * 128 + r] = sum;
      ab_workers_sync();
      sum = red_sum[r] + red_sum[128 + r] + red_sum[256 + r] + red_sum[384 + r];
      publish(&bars->p_ready);

      // ---- O / rowsum -> bf16 operand of the output projection (overwrites K and the reduction scratch) ----
      float4 rawA[8];
      if (has_next) gn_issue(b + n_clusters, 0, rawA);          // in flight across the P V MMA and the O drain
      mbar_wait(&bars->o_done, ph);
      tc_fence_after();
      ab_drain<kDbg>(tm0, R2, r, part, nullptr, 0.f, 1.0f / sum, kDbg && p.dbg[5] ? p.dbg[5] + row_g * kAbC : nullptr);
      publish(&bars->o_drained);

      // ---- next image's GroupNorm while the projection MMA runs (R0 = V^T is dead since o_done) ----
      if (has_next) {
        float4 rawB[8];
        gn_issue(b + n_clusters, 4, rawB);
        gn_finish(b + n_clusters, rawA, rawB, mom_next);
      }

      // ---- epilogue: Y^T + bias + residual -> fp32 NHWC, GroupNorm statistics of the result.  16-token chunks, the
      // residual loads of chunk i + 1 (and of chunk 0 across the wait for the MMA) in flight while chunk i is stored ----
      {
        const int c = rank * 128 + r;                               // TMEM lane = output channel
        const float bp = __ldg(p.bias + 768 + c);
        const size_t base0 = ((size_t)b * kAbT + part * 64) * kAbC + c;
        float s1 = 0.f, s2 = 0.f;
        float res[2][16];
#pragma unroll
        for (int j = 0; j < 16; ++j) res[0][j] = __ldg(p.x + base0 + (size_t)j * kAbC);
        mbar_wait(&bars->y_done, ph);
        tc_fence_after();
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          const size_t base = base0 + (size_t)(ch * 16) * kAbC;
          if (ch < 3) {
#pragma unroll
            for (int j = 0; j < 16; ++j) res[(ch + 1) & 1][j] = __ldg(p.x + base + (size_t)(16 + j) * kAbC);
          }
          uint32_t v[16];
          tmem_ld_x16(tm1 + (uint32_t)(part * 64 + ch * 16), v);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float y = __uint_as_float(v[j]) + bp + res[ch & 1][j];
            p.out[base + (size_t)j * kAbC] = y;
            s1 += y;
            s2 = fmaf(y, y, s2);
          }
        }
        if (p.out_stats != nullptr) stat_add(p.out_stats + ((size_t)b * kAbC + c) * 2, s1, s2);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&bars->y_drained);
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // the peer may still be reading this CTA's smem / arriving on its barriers
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attn_block_fwd(const b200_attn_block_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d != nullptr, "attn_block_fwd: null descriptor");
  B200_REQUIRE(d->x && d->x_stats && d->gamma && d->beta && d->w && d->bias && d->out, "attn_block_fwd: null pointer");
  B200_REQUIRE(d->T == kAbT && d->C == kAbC && d->heads == 1 && d->groups == 32,
               "attn_block_fwd: specialised to T=256, C=256, one head, 32 groups (got T=%d C=%d heads=%d groups=%d)", d->T,
               d->C, d->heads, d->groups);
  B200_REQUIRE(d->B >= 1, "attn_block_fwd: B=%d", d->B);
  B200_REQUIRE(((uintptr_t)d->x & 15) == 0 && ((uintptr_t)d->out & 15) == 0 && ((uintptr_t)d->w & 127) == 0 &&
               ((uintptr_t)d->bias & 15) == 0 && ((uintptr_t)d->gamma & 15) == 0 && ((uintptr_t)d->beta & 15) == 0 &&
               ((uintptr_t)d->x_stats & 15) == 0, "attn_block_fwd: alignment");
  B200_REQUIRE(d->x != d->out, "attn_block_fwd: in-place operation is not supported (the residual is re-read)");
  AttnBlockParams p;
  memset(&p, 0, sizeof(p));
  p.x = d->x; p.x_stats = d->x_stats; p.gamma = d->gamma; p.beta = d->beta; p.bias = d->bias;
  p.out = d->out; p.out_stats = d->out_stats; p.eps = d->eps;
  p.scale_log2e = d->scale * 1.4426950408889634f;
  p.B = d->B;
  bool dbg = false;
  for (int i = 0; i < 6; ++i) {
    p.dbg[i] = reinterpret_cast<__nv_bfloat16*>(d->dbg[i]);
    dbg |= d->dbg[i] != nullptr;
  }
  CUtensorMap mapW;
  {
    uint64_t dims[2] = {(uint64_t)kAbC, (uint64_t)4 * kAbC};
    uint64_t strides[1] = {(uint64_t)kAbC * 2};
    uint32_t box[2] = {64, 128};
    int rc = encode_tmap(&mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    B200_CHECK(cudaGetDevice(&dev));
    B200_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B200_CHECK(cudaFuncSetAttribute(attn_block_pair_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAbSmemBytes));
    B200_CHECK(cudaFuncSetAttribute(attn_block_pair_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAbSmemBytes));
  }
  const int max_clusters = sms / 2;
  const int clusters = d->B < max_clusters ? d->B : max_clusters;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(kAbThreads);
  cfg.dynamicSmemBytes = kAbSmemBytes;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (dbg) B200_CHECK(cudaLaunchKernelEx(&cfg, attn_block_pair_kernel<true>, mapW, p));
  else B200_CHECK(cudaLaunchKernelEx(&cfg, attn_block_pair_kernel<false>, mapW, p));
  ++g_launch_count;
  return 0;
}
