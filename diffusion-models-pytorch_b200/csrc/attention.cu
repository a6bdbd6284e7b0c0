// K2: fused attention softmax(q k^T * scale) v for the low-resolution self-attention blocks
// (models/modules.py:92-97: q*scale, bmm, softmax, bmm; [B*h, T, T] is never materialised in HBM here).
//
// One CTA per (128-query tile, head, image), T <= 256 keys so the whole score row fits in TMEM:
//   1. TMA: Q tile [128 x d] and K [Tp x d] (Tp = T rounded up to 64; out-of-range rows are zero-filled).
//   2. tcgen05.mma: S[128 x Tp] = Q K^T   (fp32, TMEM columns [0, Tp)).
//   3. 128 threads, one score row each: tcgen05.ld S, row max, p = exp2((s - max) * scale * log2e), row sum,
//      P rounded to bf16 and written to smem in the 128B-swizzled K-major layout UMMA expects.
//      Meanwhile TMA brings V^T [d x Tp] into the smem that held K.
//   4. tcgen05.mma: O[128 x d] = P V      (TMEM columns [Tp, Tp + d)).
//   5. tcgen05.ld O, scale by 1/rowsum, store bf16 [B][T][heads*d].
#include "common.cuh"
#include <stdlib.h>
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

struct AttnParams {
  int T, Tp, d;
  int q_off, k_off;
  float scale_log2e;
  __nv_bfloat16* out;
  int ld_out;
  uint32_t tmem_cols;
  float* lse;      // optional [B][heads][T]: log2-sum-exp of the scaled score rows (for the fused backward), or NULL
};

struct __align__(8) AttnBars {
  uint64_t qk_full, v_full, s_done, o_done;
  uint32_t tmem_base;
};

// NS = threads per score row: the row's Tp columns (softmax) and d columns (output) are split over NS warps that share
// a TMEM lane quarter (warp % 4).  With one thread per row the softmax + output stage ran as ONE warp per scheduler
// (~6 k dependent instructions at IPC 0.8 per SM, about half of the CTA's lifetime); NS warps per scheduler both
// divide the work and hide each other's latency.  Row max / row sum are combined through shared memory.
template <int NS>
__global__ void __launch_bounds__(128 * NS, 1)
attention_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                 const __grid_constant__ CUtensorMap mapV, const __grid_constant__ AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int dk = p.d >> 6;    // 64-wide chunks of the head dim
  const int tk = p.Tp >> 6;   // 64-wide chunks of the keys
  const int qp_bytes = (dk > tk ? dk : tk) * 16384;  // Q tile, later P tile
  uint8_t* sQ = smem;                 // [dk][128 rows][128 B]   (aliased by P: [tk][128 rows][128 B])
  uint8_t* sK = smem + qp_bytes;      // [dk][Tp rows][128 B]    (aliased by V^T: [tk][d rows][128 B])
  AttnBars* bars = reinterpret_cast<AttnBars*>(sK + (size_t)dk * p.Tp * 128);
  float* red_max = reinterpret_cast<float*>(bars + 1);   // [NS][128]
  float* red_sum = red_max + NS * 128;                   // [NS][128]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wq = warp & 3, part = warp >> 2;             // TMEM lane quarter, column split owned by this warp
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(&bars->qk_full, 1);
    mbar_init(&bars->v_full, 1);
    mbar_init(&bars->s_done, 1);
    mbar_init(&bars->o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, p.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = bars->tmem_base;
  const uint32_t tmem_o = tmem_s + (uint32_t)p.Tp;
  griddep_sync();

  if (threadIdx.x == 0) {
    // ---- 1. loads + 2. S = Q K^T ----
    mbar_arrive_expect_tx(&bars->qk_full, (uint32_t)(dk * 16384 + dk * p.Tp * 128));
    for (int c = 0; c < dk; ++c) {
      tma_load_3d(sQ + c * 16384, &mapQ, &bars->qk_full, p.q_off + h * p.d + c * 64, q0, b);
      tma_load_3d(sK + (size_t)c * p.Tp * 128, &mapK, &bars->qk_full, p.k_off + h * p.d + c * 64, 0, b);
    }
    mbar_wait(&bars->qk_full, 0);
    tc_fence_after();
    const uint32_t idesc_s = umma_idesc_bf16_m128((uint32_t)p.Tp);
    for (int c = 0; c < dk; ++c) {
      const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sQ + c * 16384));
      const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(sK + (size_t)c * p.Tp * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_s, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
    }
    umma_commit(&bars->s_done);
  }
  __syncwarp();

  // ---- 3. softmax over the score row held by this thread ----
  mbar_wait(&bars->s_done, 0);
  tc_fence_after();
  if (threadIdx.x == 0) {
    // K is consumed: bring V^T into its place while the softmax runs
    mbar_arrive_expect_tx(&bars->v_full, (uint32_t)(tk * p.d * 128));
    for (int c = 0; c < tk; ++c)
      tma_load_3d(sK + (size_t)c * p.d * 128, &mapV, &bars->v_full, c * 64, h * p.d, b);
  }
  __syncwarp();
  const int r = wq * 32 + lane;
  const uint32_t lane_base = (uint32_t)(wq * 32) << 16;
  const int c_lo = part * (p.Tp / NS), c_hi = c_lo + p.Tp / NS;
  float m = -INFINITY;
  for (int c = c_lo; c < c_hi; c += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem_s + lane_base + (uint32_t)c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c + j < p.T) m = fmaxf(m, __uint_as_float(v[j]));
  }
  if (NS > 1) {
    red_max[part * 128 + r] = m;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NS; ++i) m = fmaxf(m, red_max[i * 128 + r]);
  }
  const float ms = m * p.scale_log2e;
  float sum = 0.f;
  for (int c = c_lo; c < c_hi; c += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem_s + lane_base + (uint32_t)c, v);
    tmem_ld_wait();
    float e[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = (c + j < p.T) ? exp2f(__uint_as_float(v[j]) * p.scale_log2e - ms) : 0.f;
      e[j] = x;
      sum += x;
    }
    // P tile kt = c/64: row r at r*128 B, 16-byte chunk index ((c%64)/8 + i) XOR (r & 7)
    uint8_t* prow = sQ + (size_t)(c >> 6) * 16384 + (size_t)r * 128;
    const int chunk0 = (c & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(e[8 * i + 0], e[8 * i + 1]);
      u.y = pack_bf16x2(e[8 * i + 2], e[8 * i + 3]);
      u.z = pack_bf16x2(e[8 * i + 4], e[8 * i + 5]);
      u.w = pack_bf16x2(e[8 * i + 6], e[8 * i + 7]);
      *reinterpret_cast<uint4*>(prow + (((chunk0 + i) ^ (r & 7)) << 4)) = u;
    }
  }
  if (NS > 1) red_sum[part * 128 + r] = sum;
  // generic-proxy writes of P -> visible to the tensor-core (async) proxy
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (NS > 1) {
    sum = 0.f;
#pragma unroll
    for (int i = 0; i < NS; ++i) sum += red_sum[i * 128 + r];
  }
  // softmax_ij = exp2(s_ij * scale * log2e - lse_i): what the fused backward (attn_bwd.cu) recomputes P from
  if (p.lse != nullptr && part == 0 && q0 + r < p.T)
    p.lse[((size_t)b * gridDim.y + h) * p.T + q0 + r] = ms + log2f(sum);

  // ---- 4. O = P V ----
  if (threadIdx.x == 0) {
    mbar_wait(&bars->v_full, 0);
    tc_fence_after();
    const uint32_t idesc_o = umma_idesc_bf16_m128((uint32_t)p.d);
    for (int c = 0; c < tk; ++c) {
      const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sQ + c * 16384));
      const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(sK + (size_t)c * p.d * 128));
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_o, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_o, (c > 0 || k > 0) ? 1u : 0u);
    }
    umma_commit(&bars->o_done);
  }
  __syncwarp();

  // ---- 5. normalise and store ----
  mbar_wait(&bars->o_done, 0);
  tc_fence_after();
  const float inv = 1.0f / sum;
  const bool row_ok = (q0 + r) < p.T;
  __nv_bfloat16* orow = p.out + ((size_t)b * p.T + q0 + r) * p.ld_out + h * p.d;
  const int d_lo = part * (p.d / NS), d_hi = d_lo + p.d / NS;
  for (int c = d_lo; c < d_hi; c += 32) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(tmem_o + lane_base + (uint32_t)c, v);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + c + 8 * i) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_s, p.tmem_cols);
  }
}

}  // namespace b200

namespace b200 {
int attention_kv64(const void* qk, int ld_qk, int q_off, int k_off, const void* vt, void* out, int ld_out, int B, int T,
                   int heads, float scale, cudaStream_t stream);
int attention_wide(const void* qk, int ld_qk, int q_off, int k_off, const void* vt, void* out, int ld_out, int B, int T,
                   int heads, int d, float scale, cudaStream_t stream);
}
using namespace b200;

extern "C" int b200_attention_fwd(const void* qk, int ld_qk, int q_off, int k_off, const void* vt, void* out, int ld_out,
                                  int B, int T, int heads, int d, float scale, void* stream_) {
  return b200_attention_fwd_lse(qk, ld_qk, q_off, k_off, vt, out, ld_out, B, T, heads, d, scale, nullptr, stream_);
}

extern "C" int b200_attention_fwd_lse(const void* qk, int ld_qk, int q_off, int k_off, const void* vt, void* out, int ld_out,
                                      int B, int T, int heads, int d, float scale, float* lse, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(lse == nullptr || (T <= 256 && (d == 64 || d == 128 || d == 256)),
               "attention_fwd_lse: the log-sum-exp output is available for T <= 256 and head dim 64 / 128 / 256");
  B200_REQUIRE(qk && vt && out, "attention_fwd: null pointer");
  B200_REQUIRE(d >= 64 && d <= 512 && d % 64 == 0, "attention_fwd: head dim %d must be a multiple of 64 in [64,512]", d);
  B200_REQUIRE(T >= 8 && T % 8 == 0, "attention_fwd: T=%d must be a positive multiple of 8", T);
  static const char* env_wide = getenv("B200_ATTN_WIDE");   // experiment: chunk-streaming kernel for single-head d = 256
  const bool force_wide = env_wide && atoi(env_wide) == 1 && heads == 1 && d == 256 && lse == nullptr;
  if (T <= 256 && (force_wide || !(d == 64 || d == 128 || d == 256))) {
    B200_REQUIRE(ld_qk % 8 == 0 && ld_out % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0, "attention_fwd: ld/offsets must be multiples of 8");
    B200_REQUIRE(((uintptr_t)qk & 127) == 0 && ((uintptr_t)vt & 127) == 0 && ((uintptr_t)out & 15) == 0, "attention_fwd: alignment");
    return attention_wide(qk, ld_qk, q_off, k_off, vt, out, ld_out, B, T, heads, d, scale, stream);
  }
  if (T > 256) {
    B200_REQUIRE(d == 64, "attention_fwd: T=%d > 256 is supported for head dim 64 only (got %d)", T, d);
    B200_REQUIRE(ld_qk % 8 == 0 && ld_out % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0, "attention_fwd: ld/offsets must be multiples of 8");
    return attention_kv64(qk, ld_qk, q_off, k_off, vt, out, ld_out, B, T, heads, scale, stream);
  }
  B200_REQUIRE(ld_qk % 8 == 0 && ld_out % 8 == 0 && q_off % 8 == 0 && k_off % 8 == 0, "attention_fwd: ld/offsets must be multiples of 8");
  B200_REQUIRE(((uintptr_t)qk & 127) == 0 && ((uintptr_t)vt & 127) == 0 && ((uintptr_t)out & 15) == 0, "attention_fwd: alignment");
  const int Tp = (T + 63) / 64 * 64;
  AttnParams p;
  p.T = T; p.Tp = Tp; p.d = d; p.q_off = q_off; p.k_off = k_off;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  p.lse = lse;
  uint32_t cols = 32;
  while (cols < (uint32_t)(Tp + d)) cols <<= 1;
  p.tmem_cols = cols;
  CUtensorMap mapQ, mapK, mapV;
  {
    uint64_t dims[3] = {(uint64_t)ld_qk, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)ld_qk * 2, (uint64_t)T * ld_qk * 2};
    uint32_t boxq[3] = {64, 128, 1};
    uint32_t boxk[3] = {64, (uint32_t)Tp, 1};
    int rc = encode_tmap(&mapQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qk, dims, strides, boxq, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&mapK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qk, dims, strides, boxk, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)T, (uint64_t)heads * d, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)T * 2, (uint64_t)heads * d * T * 2};
    uint32_t box[3] = {64, (uint32_t)d, 1};
    int rc = encode_tmap(&mapV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int dk = d / 64, tk = Tp / 64;
  // threads per score row: 4 when both the key and the head dimension give every warp >= 32 columns, else 2
  static const char* env_ns = getenv("B200_ATTN_SPLIT");   // experiment knob: 1 = one thread per row (the first version)
  int ns = (Tp >= 128 && d >= 128) ? 4 : 2;
  if (env_ns && (atoi(env_ns) == 1 || atoi(env_ns) == 2)) ns = atoi(env_ns);
  const size_t smem = (size_t)(dk > tk ? dk : tk) * 16384 + (size_t)dk * Tp * 128 + sizeof(AttnBars) +
                      (size_t)2 * ns * 128 * sizeof(float) + 1024;
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(attention_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CHECK(cudaFuncSetAttribute(attention_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CHECK(cudaFuncSetAttribute(attention_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  B200_REQUIRE(smem <= 227 * 1024, "attention_fwd: smem %zu too large", smem);
  dim3 grid((T + 127) / 128, heads, B);
  if (ns == 4) B200_CHECK(launch_pdl(attention_kernel<4>, grid, dim3(512), smem, stream, mapQ, mapK, mapV, p));
  else if (ns == 2) B200_CHECK(launch_pdl(attention_kernel<2>, grid, dim3(256), smem, stream, mapQ, mapK, mapV, p));
  else B200_CHECK(launch_pdl(attention_kernel<1>, grid, dim3(128), smem, stream, mapQ, mapK, mapV, p));
  ++g_launch_count;
  return 0;
}

// ================================================================================================
// K2, long-sequence variant (T > 256, head dim 64: ADM 32x32 attention, T = 1024): flash-style loop over key/value
// tiles of 256 keys with an online softmax.  Per tile: S = Q K_j^T (tcgen05, TMEM cols [0,256)), row max / exp2 /
// row sum in registers, P (bf16) to swizzled smem, O_j = P V_j (TMEM cols [256, 320)), and the running output
// O = O * alpha + O_j is kept in 64 fp32 registers per thread (one query row each).  K/V tiles are double-buffered:
// the TMA for tile j+1 is in flight while tile j is processed.
// ================================================================================================
namespace b200 {

struct AttnKvParams {
  int T, q_off, k_off, n_tiles;
  float scale_log2e;
  __nv_bfloat16* out;
  int ld_out;
};

struct __align__(8) AttnKvBars {
  uint64_t q_full, kv_full[2], s_done, o_done;
  uint32_t tmem_base;
};

constexpr int kKvTile = 256;

__global__ void __launch_bounds__(128, 1)
attention_kv64_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapV, const __grid_constant__ AttnKvParams p) {
  constexpr int D = 64;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  uint8_t* sQ = smem;                                // [128 rows][128 B]                    16 KB
  uint8_t* sK = sQ + 16384;                          // [2][256 rows][128 B]                 64 KB
  uint8_t* sV = sK + 2 * 32768;                      // [2][4 key chunks][64 rows][128 B]    64 KB
  uint8_t* sP = sV + 2 * 32768;                      // [4 key chunks][128 rows][128 B]      64 KB
  AttnKvBars* bars = reinterpret_cast<AttnKvBars*>(sP + 65536);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    mbar_init(&bars->q_full, 1);
    mbar_init(&bars->kv_full[0], 1);
    mbar_init(&bars->kv_full[1], 1);
    mbar_init(&bars->s_done, 1);
    mbar_init(&bars->o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_s = bars->tmem_base;
  const uint32_t tmem_o = tmem_s + 256;
  griddep_sync();

  auto load_kv = [&](int j) {   // thread 0 only
    const int buf = j & 1;
    mbar_arrive_expect_tx(&bars->kv_full[buf], 32768u + 32768u);
    tma_load_3d(sK + buf * 32768, &mapK, &bars->kv_full[buf], p.k_off + h * D, j * kKvTile, b);
    for (int c = 0; c < 4; ++c)
      tma_load_3d(sV + buf * 32768 + c * 8192, &mapV, &bars->kv_full[buf], j * kKvTile + c * 64, h * D, b);
  };

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(&bars->q_full, 16384u);
    tma_load_3d(sQ, &mapQ, &bars->q_full, p.q_off + h * D, q0, b);
    load_kv(0);
    if (p.n_tiles > 1) load_kv(1);
    mbar_wait(&bars->q_full, 0);
  }
  __syncwarp();

  const int r = warp * 32 + lane;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  const uint32_t idesc_s = umma_idesc_bf16_m128(kKvTile);
  const uint32_t idesc_o = umma_idesc_bf16_m128(D);
  float o_acc[D];
#pragma unroll
  for (int i = 0; i < D; ++i) o_acc[i] = 0.f;
  float m_run = -INFINITY, l_run = 0.f;

  for (int j = 0; j < p.n_tiles; ++j) {
    const int buf = j & 1;
    const uint32_t par = (uint32_t)j & 1u;
    if (threadIdx.x == 0) {
      mbar_wait(&bars->kv_full[buf], (uint32_t)(j >> 1) & 1u);
      tc_fence_after();
      const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sQ));
      const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(sK + buf * 32768));
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_s, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_s, k > 0 ? 1u : 0u);
      umma_commit(&bars->s_done);
    }
    __syncwarp();
    mbar_wait(&bars->s_done, par);
    tc_fence_after();

    // ---- online softmax over this tile's 256 score columns ----
    const int kbase = j * kKvTile;
    float m_tile = -INFINITY;
    for (int c = 0; c < kKvTile; c += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem_s + lane_base + (uint32_t)c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i)
        if (kbase + c + i < p.T) m_tile = fmaxf(m_tile, __uint_as_float(v[i]));
    }
    const float m_new = fmaxf(m_run, m_tile);
    const float alpha = (m_run == -INFINITY) ? 0.f : exp2f((m_run - m_new) * p.scale_log2e);
    const float ms = m_new * p.scale_log2e;
    float sum = 0.f;
    for (int c = 0; c < kKvTile; c += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem_s + lane_base + (uint32_t)c, v);
      tmem_ld_wait();
      float e[32];
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float x = (kbase + c + i < p.T) ? exp2f(__uint_as_float(v[i]) * p.scale_log2e - ms) : 0.f;
        e[i] = x;
        sum += x;
      }
      uint8_t* prow = sP + (size_t)(c >> 6) * 16384 + (size_t)r * 128;
      const int chunk0 = (c & 63) >> 3;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(e[8 * i + 0], e[8 * i + 1]);
        u.y = pack_bf16x2(e[8 * i + 2], e[8 * i + 3]);
        u.z = pack_bf16x2(e[8 * i + 4], e[8 * i + 5]);
        u.w = pack_bf16x2(e[8 * i + 6], e[8 * i + 7]);
        *reinterpret_cast<uint4*>(prow + (((chunk0 + i) ^ (r & 7)) << 4)) = u;
      }
    }
    l_run = l_run * alpha + sum;
    m_run = m_new;
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // ---- O_j = P V_j ----
    if (threadIdx.x == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sP + c * 16384));
        const uint64_t bdesc = umma_desc_kmajor_sw128(smem_u32(sV + buf * 32768 + c * 8192));
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_o, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_o, (c > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&bars->o_done);
    }
    __syncwarp();
    mbar_wait(&bars->o_done, par);
    tc_fence_after();
    if (threadIdx.x == 0 && j + 2 < p.n_tiles) load_kv(j + 2);   // buffer `buf` is free again
    __syncwarp();
#pragma unroll
    for (int c = 0; c < D; c += 32) {
      uint32_t v[32];
      tmem_ld_x32(tmem_o + lane_base + (uint32_t)c, v);
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 32; ++i) o_acc[c + i] = o_acc[c + i] * alpha + __uint_as_float(v[i]);
    }
    // all threads must have read S / O of this tile before the next tile's MMAs overwrite them
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }

  const float inv = 1.0f / l_run;
  if (q0 + r < p.T) {
    __nv_bfloat16* orow = p.out + ((size_t)b * p.T + q0 + r) * p.ld_out + h * D;
#pragma unroll
    for (int i = 0; i < D; i += 8) {
      uint4 u;
      u.x = pack_bf16x2(o_acc[i] * inv, o_acc[i + 1] * inv);
      u.y = pack_bf16x2(o_acc[i + 2] * inv, o_acc[i + 3] * inv);
      u.z = pack_bf16x2(o_acc[i + 4] * inv, o_acc[i + 5] * inv);
      u.w = pack_bf16x2(o_acc[i + 6] * inv, o_acc[i + 7] * inv);
      *reinterpret_cast<uint4*>(orow + i) = u;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_s, 512);
  }
}

int attention_kv64(const void* qk, int ld_qk, int q_off, int k_off, const void* vt, void* out, int ld_out, int B, int T,
                   int heads, float scale, cudaStream_t stream) {
  constexpr int D = 64;
  AttnKvParams p;
  p.T = T; p.q_off = q_off; p.k_off = k_off;
  p.n_tiles = (T + kKvTile - 1) / kKvTile;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  CUtensorMap mapQ, mapK, mapV;
  {
    uint64_t dims[3] = {(uint64_t)ld_qk, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)ld_qk * 2, (uint64_t)T * ld_qk * 2};
    uint32_t boxq[3] = {64, 128, 1};
    uint32_t boxk[3] = {64, (uint32_t)kKvTile, 1};
    int rc = encode_tmap(&mapQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qk, dims, strides, boxq, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&mapK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qk, dims, strides, boxk, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)T, (uint64_t)heads * D, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)T * 2, (uint64_t)heads * D * T * 2};
    uint32_t box[3] = {64, (uint32_t)D, 1};
    int rc = encode_tmap(&mapV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const size_t smem = 16384 + 65536 + 65536 + 65536 + sizeof(AttnKvBars) + 1024;
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(attention_kv64_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  dim3 grid((T + 127) / 128, heads, B);
  B200_CHECK(launch_pdl(attention_kv64_kernel, grid, dim3(128), smem, stream, mapQ, mapK, mapV, p));
  ++g_launch_count;
  return 0;
}

}  // namespace b200


// ================================================================================================
// K2, wide-head variant (T <= 256, head dim up to 512: the single-head attention of the pesser / DDPM CelebA-HQ UNet,
// C = 512 at 16x16).  Q/K do not fit in shared memory at once, so S = sum over 64-wide channel chunks of Q_c K_c^T is
// accumulated from a 2-stage TMA ring; after the softmax (P in smem) the accumulator columns are reused for
// O[128 x d] (up to 512 fp32 columns = all of TMEM), computed as two N <= 256 halves per 64-key chunk of V^T, again
// from a 2-stage ring.
// ================================================================================================
namespace b200 {

struct AttnWideParams {
  int T, Tp, d, q_off, k_off;
  float scale_log2e;
  __nv_bfloat16* out;
  int ld_out;
};

struct __align__(8) AttnWideBars {
  uint64_t full[2], empty[2], vfull[2], vempty[2], s_done, o_done;
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(128, 1)
attention_wide_kernel(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
                      const __grid_constant__ CUtensorMap mapV, const __grid_constant__ AttnWideParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  const int dk = p.d >> 6, tk = p.Tp >> 6;
  const int qk_stage = 16384 + p.Tp * 128;          // Q chunk [128][128 B] + K chunk [Tp][128 B]
  const int v_stage = p.d * 128;                    // V^T chunk [d][128 B] (64 keys)
  const int ring = 2 * (qk_stage > v_stage ? qk_stage : v_stage);
  uint8_t* sRing = smem;
  uint8_t* sP = smem + ring;                        // [tk][128 rows][128 B]
  AttnWideBars* bars = reinterpret_cast<AttnWideBars*>(sP + (size_t)tk * 16384);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = blockIdx.x * 128, h = blockIdx.y, b = blockIdx.z;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapV);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
      mbar_init(&bars->vfull[s], 1);
      mbar_init(&bars->vempty[s], 1);
    }
    mbar_init(&bars->s_done, 1);
    mbar_init(&bars->o_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_acc = bars->tmem_base;
  griddep_sync();

  auto load_qk = [&](int c) {   // thread 0 only
    uint8_t* st = sRing + (size_t)(c & 1) * qk_stage;
    mbar_arrive_expect_tx(&bars->full[c & 1], (uint32_t)qk_stage);
    tma_load_3d(st, &mapQ, &bars->full[c & 1], p.q_off + h * p.d + c * 64, q0, b);
    tma_load_3d(st + 16384, &mapK, &bars->full[c & 1], p.k_off + h * p.d + c * 64, 0, b);
  };
  auto load_v = [&](int c) {    // thread 0 only: V^T rows [h*d, h*d + d), keys [64c, 64c + 64), in boxes of <= 256 rows
    uint8_t* st = sRing + (size_t)(c & 1) * v_stage;
    mbar_arrive_expect_tx(&bars->vfull[c & 1], (uint32_t)v_stage);
    for (int r0 = 0; r0 < p.d; r0 += 64)
      tma_load_3d(st + (size_t)r0 * 128, &mapV, &bars->vfull[c & 1], c * 64, h * p.d + r0, b);
  };

  if (threadIdx.x == 0) {
    // ---- S = sum_c Q_c K_c^T ----
    load_qk(0);
    if (dk > 1) load_qk(1);
    const uint32_t idesc_s = umma_idesc_bf16_m128((uint32_t)p.Tp);
    for (int c = 0; c < dk; ++c) {
      mbar_wait(&bars->full[c & 1], (uint32_t)(c >> 1) & 1u);
      tc_fence_after();
      const uint32_t base = smem_u32(sRing + (size_t)(c & 1) * qk_stage);
      const uint64_t adesc = umma_desc_kmajor_sw128(base);
      const uint64_t bdesc = umma_desc_kmajor_sw128(base + 16384);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        umma_bf16(tmem_acc, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_s, (c > 0 || k > 0) ? 1u : 0u);
      umma_commit(&bars->empty[c & 1]);
      if (c + 2 < dk) {
        mbar_wait(&bars->empty[c & 1], (uint32_t)(c >> 1) & 1u);
        load_qk(c + 2);
      }
    }
    umma_commit(&bars->s_done);
  }
  __syncwarp();
  mbar_wait(&bars->s_done, 0);
  tc_fence_after();
  if (threadIdx.x == 0) {   // all Q/K chunks are consumed: the ring is free for V^T
    load_v(0);
    if (tk > 1) load_v(1);
  }
  __syncwarp();

  // ---- softmax over the score row held by this thread ----
  const int r = warp * 32 + lane;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
  float m = -INFINITY;
  for (int c = 0; c < p.Tp; c += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem_acc + lane_base + (uint32_t)c, v);
    tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (c + j < p.T) m = fmaxf(m, __uint_as_float(v[j]));
  }
  const float ms = m * p.scale_log2e;
  float sum = 0.f;
  for (int c = 0; c < p.Tp; c += 32) {
    uint32_t v[32];
    tmem_ld_x32(tmem_acc + lane_base + (uint32_t)c, v);
    tmem_ld_wait();
    float e[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float x = (c + j < p.T) ? exp2f(__uint_as_float(v[j]) * p.scale_log2e - ms) : 0.f;
      e[j] = x;
      sum += x;
    }
    uint8_t* prow = sP + (size_t)(c >> 6) * 16384 + (size_t)r * 128;
    const int chunk0 = (c & 63) >> 3;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      uint4 u;
      u.x = pack_bf16x2(e[8 * i + 0], e[8 * i + 1]);
      u.y = pack_bf16x2(e[8 * i + 2], e[8 * i + 3]);
      u.z = pack_bf16x2(e[8 * i + 4], e[8 * i + 5]);
      u.w = pack_bf16x2(e[8 * i + 6], e[8 * i + 7]);
      *reinterpret_cast<uint4*>(prow + (((chunk0 + i) ^ (r & 7)) << 4)) = u;
    }
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();       // every thread has read its S row: the accumulator columns may be overwritten by O
  tc_fence_after();

  // ---- O = P V, N halves of <= 256 columns ----
  if (threadIdx.x == 0) {
    for (int c = 0; c < tk; ++c) {
      mbar_wait(&bars->vfull[c & 1], (uint32_t)(c >> 1) & 1u);
      tc_fence_after();
      const uint64_t adesc = umma_desc_kmajor_sw128(smem_u32(sP + (size_t)c * 16384));
      const uint32_t vbase = smem_u32(sRing + (size_t)(c & 1) * v_stage);
      for (int n0 = 0; n0 < p.d; n0 += 256) {
        const int nn = (p.d - n0) < 256 ? (p.d - n0) : 256;
        const uint32_t idesc_o = umma_idesc_bf16_m128((uint32_t)nn);
        const uint64_t bdesc = umma_desc_kmajor_sw128(vbase + (uint32_t)n0 * 128u);
#pragma unroll
        for (int k = 0; k < 4; ++k)
          umma_bf16(tmem_acc + (uint32_t)n0, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc_o,
                    (c > 0 || k > 0) ? 1u : 0u);
      }
      umma_commit(&bars->vempty[c & 1]);
      if (c + 2 < tk) {
        mbar_wait(&bars->vempty[c & 1], (uint32_t)(c >> 1) & 1u);
        load_v(c + 2);
      }
    }
    umma_commit(&bars->o_done);
  }
  __syncwarp();
  mbar_wait(&bars->o_done, 0);
  tc_fence_after();
  const float inv = 1.0f / sum;
  const bool row_ok = (q0 + r) < p.T;
  __nv_bfloat16* orow = p.out + ((size_t)b * p.T + q0 + r) * p.ld_out + h * p.d;
  for (int c = 0; c < p.d; c += 32) {
    uint32_t v[32];
    __syncwarp();
    tmem_ld_x32(tmem_acc + lane_base + (uint32_t)c, v);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        uint4 u;
        u.x = pack_bf16x2(__uint_as_float(v[8 * i + 0]) * inv, __uint_as_float(v[8 * i + 1]) * inv);
        u.y = pack_bf16x2(__uint_as_float(v[8 * i + 2]) * inv, __uint_as_float(v[8 * i + 3]) * inv);
        u.z = pack_bf16x2(__uint_as_float(v[8 * i + 4]) * inv, __uint_as_float(v[8 * i + 5]) * inv);
        u.w = pack_bf16x2(__uint_as_float(v[8 * i + 6]) * inv, __uint_as_float(v[8 * i + 7]) * inv);
        *reinterpret_cast<uint4*>(orow + c + 8 * i) = u;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_acc, 512);
  }
}

int attention_wide(const void* qk, int ld_qk, int q_off, int k_off, const void* vt, void* out, int ld_out, int B, int T,
                   int heads, int d, float scale, cudaStream_t stream) {
  const int Tp = (T + 63) / 64 * 64;
  AttnWideParams p;
  p.T = T; p.Tp = Tp; p.d = d; p.q_off = q_off; p.k_off = k_off;
  p.scale_log2e = scale * 1.4426950408889634f;
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld_out = ld_out;
  CUtensorMap mapQ, mapK, mapV;
  {
    uint64_t dims[3] = {(uint64_t)ld_qk, (uint64_t)T, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)ld_qk * 2, (uint64_t)T * ld_qk * 2};
    uint32_t boxq[3] = {64, 128, 1};
    uint32_t boxk[3] = {64, (uint32_t)Tp, 1};
    int rc = encode_tmap(&mapQ, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qk, dims, strides, boxq, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    rc = encode_tmap(&mapK, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, qk, dims, strides, boxk, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  {
    uint64_t dims[3] = {(uint64_t)T, (uint64_t)heads * d, (uint64_t)B};
    uint64_t strides[2] = {(uint64_t)T * 2, (uint64_t)heads * d * T * 2};
    uint32_t box[3] = {64, 64, 1};
    int rc = encode_tmap(&mapV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, vt, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  const int qk_stage = 16384 + Tp * 128, v_stage = d * 128;
  const size_t smem = (size_t)2 * (qk_stage > v_stage ? qk_stage : v_stage) + (size_t)(Tp / 64) * 16384 +
                      sizeof(AttnWideBars) + 1024;
  B200_REQUIRE(smem <= 227 * 1024, "attention_fwd: smem %zu too large", smem);
  static bool attr = false;
  if (!attr) {
    B200_CHECK(cudaFuncSetAttribute(attention_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr = true;
  }
  dim3 grid((T + 127) / 128, heads, B);
  B200_CHECK(launch_pdl(attention_wide_kernel, grid, dim3(128), smem, stream, mapQ, mapK, mapV, p));
  ++g_launch_count;
  return 0;
}

}  // namespace b200
