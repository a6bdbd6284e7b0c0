// Fused attention backward (training step, models/modules.py:92-97 under autograd) for head dim 64 and T <= 256:
//     P  = softmax(scale * Q K^T)          (recomputed from the forward's per-row log-sum-exp, never stored)
//     dV = P^T dO,   dP = dO V^T,   dS = scale * P o (dP - rowsum(dO o O)),   dQ = dS K,   dK = dS^T Q
// One CTA per (image, head).  Q, K, dO ([T][64] bf16) and V^T ([64][T]) are TMA-loaded ONCE into shared memory and serve
// every GEMM untransposed: the same tile is a K-major operand where d is contracted (S, dP) and an MN-major operand
// where tokens are contracted (dK, dV, dQ) -- only the UMMA descriptor differs.  P and dS are written once per
// 128 x 128 block as bf16 [query][key] tiles: K-major A of dQ = dS K, MN-major A of dV = P^T dO and dK = dS^T Q.
// Loop over key tiles (outer) x query tiles (inner), flash-attention-2 style; TMEM (512 columns):
//     S [0,128)  dP [128,256)  dK [256,320)  dV [320,384)  dQ(query tile 0 / 1) [384,448) / [448,512)
// It replaces 5 batched GEMMs + 2 row-softmax launches per block that moved the [B*h, T, T] score tensors through HBM
// four times (S fp32, P bf16, dP fp32, dS bf16: ~1 GB per 16x16 block of the CFG UNet at batch 128).
#include "common.cuh"
#include <string.h>
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

constexpr int kAbwWorkers = 256;
constexpr int kAbwThreads = kAbwWorkers + 32;
constexpr uint32_t kAbwTile = 16384;            // [128 rows][128 B]

struct AttnBwdParams {
  int T, heads, C, nt;
  float scale, scale_log2e;
  const __nv_bfloat16* o;        // [B][T][C]
  const __nv_bfloat16* d_o;      // [B][T][C]
  const float* lse;              // [B][heads][T]
  __nv_bfloat16* dqk;            // [B][T][ld_dqk]: dQ at columns h*64, dK at C + h*64
  __nv_bfloat16* dv;             // [B][T][ld_dv]
  int ld_dqk, ld_dv;
};

struct __align__(8) AttnBwdBars {
  uint64_t load_full, sdp_full, sdp_free, pds_ready, g2_done, dkv_drained;
  uint32_t tmem_base;
};

__device__ __forceinline__ float abw_bf16_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float abw_bf16_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float abw_ex2(float x) {      // 2^x, one MUFU (exp2f adds a denormal-range rescale: 4 more instructions)
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(kAbwThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap mapQK, const __grid_constant__ CUtensorMap mapDO,
                const __grid_constant__ CUtensorMap mapV, const __grid_constant__ AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                       // [2 tiles][128 tokens][128 B]
  uint8_t* sK = sQ + 2 * kAbwTile;
  uint8_t* sDO = sK + 2 * kAbwTile;
  uint8_t* sV = sDO + 2 * kAbwTile;         // [4 key chunks][64 d rows][128 B]
  uint8_t* sP = sV + 2 * kAbwTile;          // [2 key chunks][128 queries][128 B]
  uint8_t* sDS = sP + 2 * kAbwTile;
  AttnBwdBars* bars = reinterpret_cast<AttnBwdBars*>(sDS + 2 * kAbwTile);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.x, b = blockIdx.y;
  const int nt = p.nt;

  if (threadIdx.x == kAbwWorkers) {
    tma_prefetch_desc(&mapQK);
    tma_prefetch_desc(&mapDO);
    tma_prefetch_desc(&mapV);
    mbar_init(&bars->load_full, 1);
    mbar_init(&bars->sdp_full, 1);
    mbar_init(&bars->sdp_free, kAbwWorkers / 32);
    mbar_init(&bars->pds_ready, kAbwWorkers / 32);
    mbar_init(&bars->g2_done, 1);
    mbar_init(&bars->dkv_drained, kAbwWorkers / 32);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = bars->tmem_base;
  griddep_sync();

  if (warp == kAbwWorkers / 32) {
    // ================================ TMA loads + MMA issue (one thread) ================================
    if (lane == 0) {
      const int vchunks = (p.T + 63) >> 6;
      mbar_arrive_expect_tx(&bars->load_full, (uint32_t)(3 * nt) * kAbwTile + (uint32_t)vchunks * 8192u);
      for (int t = 0; t < nt; ++t) {
        tma_load_3d(sQ + t * kAbwTile, &mapQK, &bars->load_full, h * 64, t * 128, b);
        tma_load_3d(sK + t * kAbwTile, &mapQK, &bars->load_full, p.C + h * 64, t * 128, b);
        tma_load_3d(sDO + t * kAbwTile, &mapDO, &bars->load_full, h * 64, t * 128, b);
      }
      for (int c = 0; c < vchunks; ++c) tma_load_3d(sV + c * 8192, &mapV, &bars->load_full, c * 64, h * 64, b);
      mbar_wait(&bars->load_full, 0);
      tc_fence_after();
      const uint32_t i_s = umma_idesc_bf16_m128(128u);                             // S: both K-major
      const uint32_t i_dp = umma_idesc_bf16_m128(128u) | (1u << 16);               // dP: B = V^T MN-major
      const uint32_t i_kv = umma_idesc_bf16_m128(64u) | (1u << 15) | (1u << 16);   // dV / dK: both MN-major
      const uint32_t i_dq = umma_idesc_bf16_m128(64u) | (1u << 16);                // dQ: A = dS K-major, B = K MN-major
      const uint32_t aQ = smem_u32(sQ), aK = smem_u32(sK), aDO = smem_u32(sDO), aV = smem_u32(sV), aP = smem_u32(sP),
                     aDS = smem_u32(sDS);
      // S = Q_qt K_kt^T, dP = dO_qt V_kt^T of block (kt, qt) into TMEM columns [0, 256)
      auto issue_sdp = [&](int kt, int qt) {
        const uint64_t qd = umma_desc_kmajor_sw128(aQ + (uint32_t)qt * kAbwTile);
        const uint64_t kd = umma_desc_kmajor_sw128(aK + (uint32_t)kt * kAbwTile);
        const uint64_t dod = umma_desc_kmajor_sw128(aDO + (uint32_t)qt * kAbwTile);
        const uint64_t vd = umma_desc_mnmajor_sw128(aV + (uint32_t)kt * 2u * 8192u, 8192u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem, qd + (uint64_t)(2 * k), kd + (uint64_t)(2 * k), i_s, k > 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16(tmem + 128u, dod + (uint64_t)(2 * k), vd + (uint64_t)(128 * k), i_dp, k > 0 ? 1u : 0u);
        umma_commit(&bars->sdp_full);
      };
      issue_sdp(0, 0);
      int it = 0;
      for (int kt = 0; kt < nt; ++kt) {
        for (int qt = 0; qt < nt; ++qt, ++it) {
          // the next block's S / dP are issued as soon as the workers hold this block's in registers, so they run
          // on the tensor pipe while the workers compute P / dS (and the workers' next block overlaps this block's
          // dV / dK / dQ MMAs)
          if (it + 1 < nt * nt) {
            mbar_wait(&bars->sdp_free, (uint32_t)it & 1u);
            tc_fence_after();
            const int nqt = qt + 1 < nt ? qt + 1 : 0;
            issue_sdp(nqt == 0 ? kt + 1 : kt, nqt);
          }
          mbar_wait(&bars->pds_ready, (uint32_t)it & 1u);
          if (qt == 0 && kt > 0) mbar_wait(&bars->dkv_drained, (uint32_t)(kt - 1) & 1u);
          tc_fence_after();
          {
            const uint64_t pd = umma_desc_mnmajor_sw128(aP, kAbwTile);             // A = P^T: M = keys, K = queries
            const uint64_t dsd_mn = umma_desc_mnmajor_sw128(aDS, kAbwTile);        // A = dS^T
            const uint64_t dob = umma_desc_mnmajor_sw128(aDO + (uint32_t)qt * kAbwTile, kAbwTile);   // B: N = d, K = queries
            const uint64_t qb = umma_desc_mnmajor_sw128(aQ + (uint32_t)qt * kAbwTile, kAbwTile);
            const uint64_t kb = umma_desc_mnmajor_sw128(aK + (uint32_t)kt * kAbwTile, kAbwTile);     // B: N = d, K = keys
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem + 320u, pd + (uint64_t)(128 * k), dob + (uint64_t)(128 * k), i_kv, (qt > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k)
              umma_bf16(tmem + 256u, dsd_mn + (uint64_t)(128 * k), qb + (uint64_t)(128 * k), i_kv, (qt > 0 || k > 0) ? 1u : 0u);
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              // A = dS K-major: key chunk k / 4 (16 KB apart), 32 B per 16 keys inside it; B = K rows 16 k .. 16 k + 15
              const uint64_t dsd = umma_desc_kmajor_sw128(aDS + (uint32_t)(k >> 2) * kAbwTile) + (uint64_t)(2 * (k & 3));
              umma_bf16(tmem + 384u + (uint32_t)qt * 64u, dsd, kb + (uint64_t)(128 * k), i_dq, (kt > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&bars->g2_done);
          }
        }
      }
    }
  } else {
    // ================================ workers: softmax recompute, dS, drains ================================
    const int wq = warp & 3, half = warp >> 2;
    const int r = wq * 32 + lane;
    const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
    // per query row of both query tiles: log-sum-exp of the forward and D = sum_d dO * O
    // rows past T get lse = +inf: their P and dS come out as exact zeros without a per-element test
    float lse[2] = {INFINITY, INFINITY}, drow[2] = {0.f, 0.f};
    for (int qt = 0; qt < nt; ++qt) {
      const int q = qt * 128 + r;
      if (q < p.T) {
        lse[qt] = __ldg(p.lse + ((size_t)b * p.heads + h) * p.T + q);
        const uint4* po = reinterpret_cast<const uint4*>(p.o + ((size_t)b * p.T + q) * p.C + h * 64);
        const uint4* pd = reinterpret_cast<const uint4*>(p.d_o + ((size_t)b * p.T + q) * p.C + h * 64);
        float acc = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint4 a = __ldg(po + i), g = __ldg(pd + i);
          acc = fmaf(abw_bf16_lo(a.x), abw_bf16_lo(g.x), acc); acc = fmaf(abw_bf16_hi(a.x), abw_bf16_hi(g.x), acc);
          acc = fmaf(abw_bf16_lo(a.y), abw_bf16_lo(g.y), acc); acc = fmaf(abw_bf16_hi(a.y), abw_bf16_hi(g.y), acc);
          acc = fmaf(abw_bf16_lo(a.z), abw_bf16_lo(g.z), acc); acc = fmaf(abw_bf16_hi(a.z), abw_bf16_hi(g.z), acc);
          acc = fmaf(abw_bf16_lo(a.w), abw_bf16_lo(g.w), acc); acc = fmaf(abw_bf16_hi(a.w), abw_bf16_hi(g.w), acc);
        }
        drow[qt] = acc;
      }
    }
    int it = 0;
    for (int kt = 0; kt < nt; ++kt) {
      for (int qt = 0; qt < nt; ++qt, ++it) {
        mbar_wait(&bars->sdp_full, (uint32_t)it & 1u);
        tc_fence_after();
        // p = 2^(s * scale * log2e - lse), ds = p * (dP - D) * scale = p * (dP * scale - D * scale)
        const float l2 = lse[qt], drs = drow[qt] * p.scale, sl2 = p.scale_log2e, sc = p.scale;
        uint32_t pp[32], dd[32];            // this thread's 64 keys of P and dS, packed bf16 pairs
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t sv[32], dv[32];
          tmem_ld_x32(tmem + lane_base + (uint32_t)(half * 64 + c * 32), sv);
          tmem_ld_x32(tmem + 128u + lane_base + (uint32_t)(half * 64 + c * 32), dv);
          tmem_ld_wait();
          if (c == 1) {          // S / dP of this block are in registers: their TMEM columns may be overwritten
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->sdp_free);
          }
          const int key0 = kt * 128 + half * 64 + c * 32;
          if (key0 + 32 <= p.T) {           // warp-uniform: all 32 keys of the chunk exist (always, when T % 128 == 0)
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              const float p0 = abw_ex2(fmaf(__uint_as_float(sv[j]), sl2, -l2));
              const float p1 = abw_ex2(fmaf(__uint_as_float(sv[j + 1]), sl2, -l2));
              pp[c * 16 + (j >> 1)] = pack_bf16x2(p0, p1);
              dd[c * 16 + (j >> 1)] = pack_bf16x2(p0 * fmaf(__uint_as_float(dv[j]), sc, -drs),
                                                  p1 * fmaf(__uint_as_float(dv[j + 1]), sc, -drs));
            }
          } else if (key0 >= p.T) {         // no key of the chunk exists (T <= 64 in a 128-key tile)
#pragma unroll
            for (int j = 0; j < 16; ++j) pp[c * 16 + j] = dd[c * 16 + j] = 0u;
          } else {
#pragma unroll
            for (int j = 0; j < 32; j += 2) {
              float p0 = abw_ex2(fmaf(__uint_as_float(sv[j]), sl2, -l2));
              float p1 = abw_ex2(fmaf(__uint_as_float(sv[j + 1]), sl2, -l2));
              if (key0 + j >= p.T) p0 = 0.f;
              if (key0 + j + 1 >= p.T) p1 = 0.f;
              pp[c * 16 + (j >> 1)] = pack_bf16x2(p0, p1);
              dd[c * 16 + (j >> 1)] = pack_bf16x2(p0 * fmaf(__uint_as_float(dv[j]), sc, -drs),
                                                  p1 * fmaf(__uint_as_float(dv[j + 1]), sc, -drs));
            }
          }
        }
        if (it > 0) mbar_wait(&bars->g2_done, (uint32_t)(it - 1) & 1u);     // the previous block's MMAs have read P / dS
        {
          uint8_t* prow = sP + (size_t)half * kAbwTile + (size_t)r * 128;
          uint8_t* drw = sDS + (size_t)half * kAbwTile + (size_t)r * 128;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int sw = ((i ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(prow + sw) = make_uint4(pp[4 * i], pp[4 * i + 1], pp[4 * i + 2], pp[4 * i + 3]);
            *reinterpret_cast<uint4*>(drw + sw) = make_uint4(dd[4 * i], dd[4 * i + 1], dd[4 * i + 2], dd[4 * i + 3]);
          }
        }
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&bars->pds_ready);

        if (qt == nt - 1) {
          // dK / dV of key tile kt are complete: TMEM lanes = keys, 64 columns = d; this warp takes 32 of them
          mbar_wait(&bars->g2_done, (uint32_t)it & 1u);
          tc_fence_after();
          const int key = kt * 128 + r;
          uint32_t v[32];
          tmem_ld_x32(tmem + 256u + lane_base + (uint32_t)(half * 32), v);
          tmem_ld_wait();
          if (key < p.T) {
            uint4* o = reinterpret_cast<uint4*>(p.dqk + ((size_t)b * p.T + key) * p.ld_dqk + p.C + h * 64 + half * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o[i] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1])),
                                pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3])),
                                pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                                pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
          }
          tmem_ld_x32(tmem + 320u + lane_base + (uint32_t)(half * 32), v);
          tmem_ld_wait();
          if (key < p.T) {
            uint4* o = reinterpret_cast<uint4*>(p.dv + ((size_t)b * p.T + key) * p.ld_dv + h * 64 + half * 32);
#pragma unroll
            for (int i = 0; i < 4; ++i)
              o[i] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1])),
                                pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3])),
                                pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                                pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
          }
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&bars->dkv_drained);
        }
      }
    }
    // dQ of both query tiles (the last g2_done was awaited in the last block's drain)
    for (int qt = 0; qt < nt; ++qt) {
      const int q = qt * 128 + r;
      uint32_t v[32];
      tmem_ld_x32(tmem + 384u + (uint32_t)qt * 64u + lane_base + (uint32_t)(half * 32), v);
      tmem_ld_wait();
      if (q < p.T) {
        uint4* o = reinterpret_cast<uint4*>(p.dqk + ((size_t)b * p.T + q) * p.ld_dqk + h * 64 + half * 32);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          o[i] = make_uint4(pack_bf16x2(__uint_as_float(v[8 * i]), __uint_as_float(v[8 * i + 1])),
                            pack_bf16x2(__uint_as_float(v[8 * i + 2]), __uint_as_float(v[8 * i + 3])),
                            pack_bf16x2(__uint_as_float(v[8 * i + 4]), __uint_as_float(v[8 * i + 5])),
                            pack_bf16x2(__uint_as_float(v[8 * i + 6]), __uint_as_float(v[8 * i + 7])));
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attention_bwd(const void* qk, const void* vt, const void* o, const void* d_o, const float* lse, void* dqk,
                                  int ld_dqk, void* dv, int ld_dv, int B, int T, int heads, int d, float scale, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(qk && vt && o && d_o && lse && dqk && dv, "attention_bwd: null pointer");
  B200_REQUIRE(d == 64, "attention_bwd: head dim %d (only 64 is fused; other shapes take the batched-GEMM path)", d);
  B200_REQUIRE(T >= 8 && T <= 256 && T % 8 == 0, "attention_bwd: T=%d must be a multiple of 8 in [8, 256]", T);
  B200_REQUIRE(B >= 1 && heads >= 1, "attention_bwd: bad B / heads");
  const int C = heads * 64;
  B200_REQUIRE(((uintptr_t)qk & 127) == 0 && ((uintptr_t)vt & 127) == 0 && ((uintptr_t)d_o & 127) == 0 &&
               ((uintptr_t)o & 15) == 0 && ((uintptr_t)dqk & 15) == 0 && ((uintptr_t)dv & 15) == 0, "attention_bwd: alignment");
  AttnBwdParams p;
  p.T = T; p.heads = heads; p.C = C; p.nt = (T + 127) / 128;
  p.scale = scale; p.scale_log2e = scale * 1.4426950408889634f;
  p.o = reinterpret_cast<const __nv_bfloat16*>(o);
  p.d_o = reinterpret_cast<const __nv_bfloat16*>(d_o);
  p.lse = lse;
  p.dqk = reinterpret_cast<__nv_bfloat16*>(dqk);
  p.dv = reinterpret_cast<__nv_bfloat16*>(dv);
  B200_REQUIRE(ld_dqk >= 2 * C && ld_dqk % 8 == 0 && ld_dv >= C && ld_dv % 8 == 0,
               "attention_bwd: ld_dqk=%d / ld_dv=%d must be multiples of 8 covering 2C / C columns", ld_dqk, ld_dv);
  p.ld_dqk = ld_dqk; p.ld_dv = ld_dv;
  CUtensorMap mapQK, mapDO, mapV;
  {
    uint64_t dims[3] = {(uint64_t)2 * C, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)2 * C * 2, (uint64_t)T * 2 * C * 2};
    uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap(&mapQK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qk, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)C * 2, (uint64_t)T * C * 2};
    uint32_t box[3] = {64, 128, 1};
    int rc = encode_tmap(&mapDO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, d_o, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)T, (uint64_t)C, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)T * 2, (uint64_t)C * T * 2};
    uint32_t box[3] = {64, 64, 1};
    int rc = encode_tmap(&mapV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const size_t smem = 12 * (size_t)kAbwTile + sizeof(AttnBwdBars) + 1024;
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr = true;
  }
  B200_CHECK(launch_pdl(attn_bwd_kernel, dim3(heads, B), dim3(kAbwThreads), smem, stream, mapQK, mapDO, mapV, p));
  ++g_launch_count;
  return 0;
}
