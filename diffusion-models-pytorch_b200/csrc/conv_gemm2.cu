// K1, CTA-pair variant (tcgen05 cta_group::2) for convolutions with Cout % 256 == 0.
//
// The single-CTA kernel (conv_gemm.cu) is bound by how fast one SM can pull operand tiles out of L2 (~42 B/clk/SM):
// 48 KB per 128x256x64 MAC block.  Here two CTAs of a cluster (the two SMs of a TPC) share one 256-channel x 256-pixel
// tile: each loads ITS 128 weight rows (16 KB) and ITS half of the pixel tile (128 pixels, 16 KB), and the leader CTA
// issues one M=256 x N=256 UMMA per 16 k that reads both halves of B from both shared memories.  Every SM now ingests
// 32 KB for the same 128x256x64 MACs: 1.5x more math per operand byte.  Accumulators: TMEM lanes = the CTA's own
// 128 channels, 256 columns = all pixels of the tile, so the channel-major epilogue is shared with conv_gemm.cu.
//
// Pipeline: per CTA warp 0 = TMA producer (cta_group::2 loads complete on the LEADER's mbarrier), leader's warp 1 =
// MMA issuer, tcgen05.commit multicasts stage-free / accumulator-ready to both CTAs, both CTAs' epilogue warps arrive
// on the leader's accumulator-free barrier.
#include "conv_epilogue.cuh"
#include "pair.cuh"
#include <stdlib.h>
#include <string.h>

namespace b200 {

int make_a_map(CUtensorMap* m, const void* base, int C, int H, int W, int planes, int B, int bw, int bh, int bn);
extern long long g_launch_count;

struct PairParams {
  int half_dim;      // which box dimension the pixel tile is split over: 0 = w, 1 = h, 2 = n
  int half_ext;      // extent of one half along that dimension
  int pair_tiles;    // phases * p_tiles * (c_tiles / 2)
};

constexpr int kPairStageBytes = kWBytes + 128 * 128;   // 16 KB weights + 128 pixels x 128 B per CTA

template <int kT, int kLean = 0>
__global__ void __launch_bounds__(kT, 1)
conv_gemm_pair_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                      const __grid_constant__ CUtensorMap mapW, const __grid_constant__ ConvKParams p,
                      const __grid_constant__ PairParams pp) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);
  ConvBarriers* bars = reinterpret_cast<ConvBarriers*>(smem + (size_t)p.stages * kPairStageBytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = (int)cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int n_clusters = gridDim.x >> 1;
  const int nkb = p.nkb0 + p.nkb1;
  const int c_pairs = p.c_tiles >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapW);
    if (p.nkb1 > 0) tma_prefetch_desc(&mapA1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);       // leader: its own arrive.expect_tx; both CTAs' TMA bytes complete on it
      mbar_init(&bars->empty[s], 1);      // one multicast commit per use
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], 2 * 4 * p.epi_halves);   // epilogue warps of BOTH CTAs (leader's copy is used)
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_2sm(&bars->tmem_base, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();          // barriers of both CTAs initialised before any remote arrival / multicast commit
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  griddep_sync();

  // pair tile -> (phase, pixel tile, channel pair); this CTA's own tile coordinate has ct = 2 * cp + rank
  auto my_tile = [&](int ptile) {
    const int cp = ptile % c_pairs;
    const int rest = ptile / c_pairs;          // = ph * p_tiles + pt
    return decode_tile(p, rest * p.c_tiles + 2 * cp + rank);
  };

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int ptile = cluster_id; ptile < pp.pair_tiles; ptile += n_clusters) {
        const TileCoord t = my_tile(ptile);
        const int wrow = t.ph * p.w_rows_per_phase + t.ct * kBlockC;
        // this CTA's half of the pixel tile
        const int w0 = t.w0 + (pp.half_dim == 0 ? rank * pp.half_ext : 0);
        const int h0 = t.h0 + (pp.half_dim == 1 ? rank * pp.half_ext : 0);
        const int n0 = t.n0 + (pp.half_dim == 2 ? rank * pp.half_ext : 0);
        const int rot = p.k_rotate ? (int)(((unsigned)cluster_id * 5u + (unsigned)ptile) % (unsigned)nkb) : 0;
        for (int kbi = 0; kbi < nkb; ++kbi) {
          int kb = kbi + rot;
          if (kb >= nkb) kb -= nkb;
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          uint8_t* sW = smem + (size_t)stage * kPairStageBytes;
          uint8_t* sP = sW + kWBytes;
          if (rank == 0) mbar_arrive_expect_tx(&bars->full[stage], 2u * (uint32_t)kPairStageBytes);
          if (kb < p.nkb0) {
            const int tap = kb / p.cpb0;
            const int c0 = (kb - tap * p.cpb0) * kBlockK;
            tma_load_5d_2sm(sP, &mapA0, &bars->full[stage], c0, w0 + p.taps0[t.ph][tap][0], h0 + p.taps0[t.ph][tap][1],
                            p.taps0[t.ph][tap][2], n0);
          } else {
            const int c0 = (kb - p.nkb0) * kBlockK;
            tma_load_5d_2sm(sP, &mapA1, &bars->full[stage], c0, w0 + p.tap1[0], h0 + p.tap1[1], p.tap1[2], n0);
          }
          tma_load_2d_2sm(sW, &mapW, &bars->full[stage], kb * kBlockK, wrow);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    if (lane == 0 && rank == 0) {
      const uint32_t idesc = umma_idesc_bf16_m256(256u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int ptile = cluster_id; ptile < pp.pair_tiles; ptile += n_clusters, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * 256);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t w_addr = smem_u32(smem + (size_t)stage * kPairStageBytes);
          const uint64_t wdesc = umma_desc_kmajor_sw128(w_addr);
          const uint64_t pdesc = umma_desc_kmajor_sw128(w_addr + kWBytes);
#pragma unroll
          for (int k = 0; k < kBlockK / 16; ++k)
            umma_bf16_2sm(tmem_d, wdesc + (uint64_t)(2 * k), pdesc + (uint64_t)(2 * k), idesc, (kb > 0 || k > 0) ? 1u : 0u);
          umma_commit_2sm(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit_2sm(&bars->tmem_full[as]);
      }
    }
  } else {
    // ================================ epilogue (both CTAs) ================================
    const int q = warp & 3;
    const int half = (warp - 2) >> 2;
    int it = 0;
    for (int ptile = cluster_id; ptile < pp.pair_tiles; ptile += n_clusters, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const TileCoord t = my_tile(ptile);
      const int c = t.ct * kBlockC + q * 32 + lane;
      const bool c_ok = c < p.N;
      const uint32_t taddr = tmem_base + (uint32_t)(as * 256) + ((uint32_t)(q * 32) << 16);
      if (kLean) conv_epilogue_lean_dispatch(p, t, taddr, c, c_ok, half, &bars->tmem_full[as], aphase);
      else conv_epilogue_tile(p, t, taddr, c, c_ok, half, &bars->tmem_full[as], aphase);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_leader(&bars->tmem_empty[as]);
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

// Called by b200_conv2d_fwd with the fully populated single-CTA parameters (NP == 256, N % 256 == 0).
int conv2d_fwd_pair(const b200_conv_desc* d, const ConvKParams& p1, cudaStream_t stream) {
  ConvKParams p = p1;
  PairParams pp;
  int hbw = p.bw, hbh = p.bh, hbn = p.bn;
  if (p.bn > 1) { pp.half_dim = 2; hbn = p.bn / 2; pp.half_ext = hbn; }
  else if (p.bh > 1) { pp.half_dim = 1; hbh = p.bh / 2; pp.half_ext = hbh; }
  else { pp.half_dim = 0; hbw = p.bw / 2; pp.half_ext = hbw; }
  pp.pair_tiles = d->phases * p.p_tiles * (p.c_tiles / 2);
  int stages = (227 * 1024 - 2048) / kPairStageBytes;
  if (stages > kMaxStages) stages = kMaxStages;
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * kPairStageBytes + sizeof(ConvBarriers) + 1024;

  CUtensorMap mapA0, mapA1, mapW;
  int rc = make_a_map(&mapA0, d->a0, d->a0_C, d->a0_H, d->a0_W, d->a0_planes, d->B, hbw, hbh, hbn);
  if (rc) return rc;
  if (d->a1) {
    rc = make_a_map(&mapA1, d->a1, d->a1_C, d->a1_H, d->a1_W, d->a1_planes, d->B, hbw, hbh, hbn);
    if (rc) return rc;
  } else {
    mapA1 = mapA0;
  }
  {
    uint64_t dims[2] = {(uint64_t)d->w_K, (uint64_t)d->w_rows};
    uint64_t strides[1] = {(uint64_t)d->w_K * 2};
    uint32_t box[2] = {64, (uint32_t)kBlockC};
    rc = encode_tmap(&mapW, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }
  static int sms = 0;
  if (sms == 0) {
    int dev = 0;
    B200_CHECK(cudaGetDevice(&dev));
    B200_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    B200_CHECK(cudaFuncSetAttribute(conv_gemm_pair_kernel<kThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    B200_CHECK(cudaFuncSetAttribute(conv_gemm_pair_kernel<kThreadsWide>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int max_clusters = sms / 2;
  const int clusters = pp.pair_tiles < max_clusters ? pp.pair_tiles : max_clusters;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(2 * clusters);
  cfg.blockDim = dim3(64 + 128 * p.epi_halves);
  cfg.dynamicSmemBytes = smem_bytes;
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
  static const char* env_lean = getenv("B200_EPI_LEAN");   // =0: generic epilogue everywhere (A/B timing)
  if (p.epi_halves == 4)
    B200_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm_pair_kernel<kThreadsWide>, mapA0, mapA1, mapW, p, pp));
  else if (!(env_lean && atoi(env_lean) == 0) && conv_epilogue_lean_ok(p)) {
    static bool lean_attr = false;    // set lazily: the default path never touches the experimental instantiation
    if (!lean_attr) {
      B200_CHECK(cudaFuncSetAttribute(conv_gemm_pair_kernel<kThreads, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
      lean_attr = true;
    }
    B200_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm_pair_kernel<kThreads, 1>, mapA0, mapA1, mapW, p, pp));
  }
  else
    B200_CHECK(cudaLaunchKernelEx(&cfg, conv_gemm_pair_kernel<kThreads>, mapA0, mapA1, mapW, p, pp));
  ++g_launch_count;
  return 0;
}

}  // namespace b200
