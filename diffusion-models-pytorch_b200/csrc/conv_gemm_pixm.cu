// K1 (pixel-major variant, used for tiny Cout <= 32 such as the last 128->3 conv; conv_gemm.cu holds the main
// channel-major kernel): convolution as implicit GEMM on the 5th-gen tensor cores (tcgen05.mma, accumulators in
// TMEM), operands staged by TMA with 128-byte swizzle.  Replaces nn.Conv2d + the adds around it in
// models/unet.py:10-43, models/modules.py:60-102 of the reference (see include/b200diff.h).
//
// GEMM view: D[128 pixels x block_n channels] per tile, K-blocks of 64 bf16 = (tap, 64-channel chunk).
//   A tile  = TMA 5-D box (64 ch, bw, bh, 1 plane, bn images) of an NHWC bf16 activation, shifted by the
//             tap offset; out-of-image elements are zero-filled by TMA, which *is* the conv padding.
//   B tile  = TMA 2-D box (64, block_n) of the packed weight matrix [Cout][K].
// Warp roles (192 threads, persistent over tiles):
//   warp 0   TMA producer      (ring of `stages` smem slots, full/empty mbarriers)
//   warp 1   MMA issuer        (single thread; tcgen05.commit releases slots and publishes accumulators)
//   warps 2-5 epilogue         (tcgen05.ld -> +bias +time-embedding row +residual -> global), overlapped
//                              with the next tile's MMAs through two TMEM accumulator stages.
#include "common.cuh"
#include <string.h>
#include "../../include/b200diff.h"

namespace b200 {

struct PixmParams {
  int B, bw, bh, bn;
  int tiles_w, tiles_h;
  int m_tiles, n_tiles, total_tiles;
  int N, block_n, w_rows_per_phase;
  int cpb0, nkb0, nkb1;
  int stages;
  int vtap;   // vertical-tap reuse: one haloed pixel tile + three weight tiles per stage (see the host side)
  int8_t taps0[4][9][4];
  int8_t tap1[4];
  const float* bias;
  const float* rowadd;
  int rowadd_ld;
  const float* residual;
  int res_ld;
  void* out;
  int out_mode, out_ld, out_H, out_W, osy, osx;
};

constexpr int kPmBlockM = 128;
constexpr int kPmBlockK = 64;
constexpr int kPmABytes = kPmBlockM * kPmBlockK * 2;  // 16 KB
constexpr int kPmThreads = 192;
constexpr int kPmMaxStages = 8;

struct __align__(8) PixmBarriers {
  uint64_t full[kPmMaxStages];
  uint64_t empty[kPmMaxStages];
  uint64_t tmem_full[2];
  uint64_t tmem_empty[2];
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kPmThreads, 1)
conv_gemm_pixm_kernel(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
                 const __grid_constant__ CUtensorMap mapB, const __grid_constant__ PixmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + (((raw_addr + 1023u) & ~1023u) - raw_addr);

  const int b_bytes = p.block_n * kPmBlockK * 2;
  const int halo_bytes = (p.bh + 2) * p.bw * 128;     // vtap: bh + 2 image rows of bw pixels x 64 channels
  const int stage_bytes = p.vtap ? halo_bytes + 3 * b_bytes : kPmABytes + b_bytes;
  const int ngroups = 3 * p.cpb0;                      // vtap: (horizontal offset, 64-channel chunk) units per tile
  PixmBarriers* bars = reinterpret_cast<PixmBarriers*>(smem + (size_t)p.stages * stage_bytes);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int nkb = p.nkb0 + p.nkb1;
  const uint32_t tmem_cols = (2 * p.block_n <= 32) ? 32u : (2 * p.block_n <= 64)  ? 64u
                             : (2 * p.block_n <= 128) ? 128u : (2 * p.block_n <= 256) ? 256u : 512u;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapA0);
    tma_prefetch_desc(&mapB);
    if (p.nkb1 > 0) tma_prefetch_desc(&mapA1);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], 1);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->tmem_full[s], 1);
      mbar_init(&bars->tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(&bars->tmem_base, tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  griddep_sync();

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x) {
        const int ph = tile / (p.m_tiles * p.n_tiles);
        const int rem = tile - ph * (p.m_tiles * p.n_tiles);
        const int mt = rem / p.n_tiles;
        const int nt = rem - mt * p.n_tiles;
        const int tw = mt % p.tiles_w;
        const int th = (mt / p.tiles_w) % p.tiles_h;
        const int tn = mt / (p.tiles_w * p.tiles_h);
        const int w0 = tw * p.bw, h0 = th * p.bh, n0 = tn * p.bn;
        const int wrow = ph * p.w_rows_per_phase + nt * p.block_n;
        if (p.vtap) {
          for (int g = 0; g < ngroups; ++g) {
            const int dxi = g / p.cpb0, cb = g - dxi * p.cpb0;
            mbar_wait(&bars->empty[stage], phase ^ 1u);
            uint8_t* sA = smem + (size_t)stage * stage_bytes;
            uint8_t* sB = sA + halo_bytes;
            mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
            // rows h0-1 .. h0+bh of the column-shifted tile; out-of-image rows / columns are zero-filled (= padding)
            tma_load_5d(sA, &mapA0, &bars->full[stage], cb * kPmBlockK, w0 + dxi - 1, h0 - 1, 0, n0);
#pragma unroll
            for (int r = 0; r < 3; ++r)
              tma_load_2d(sB + r * b_bytes, &mapB, &bars->full[stage], ((r * 3 + dxi) * p.cpb0 + cb) * kPmBlockK, wrow);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          continue;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->empty[stage], phase ^ 1u);
          uint8_t* sA = smem + (size_t)stage * stage_bytes;
          uint8_t* sB = sA + kPmABytes;
          mbar_arrive_expect_tx(&bars->full[stage], (uint32_t)stage_bytes);
          if (kb < p.nkb0) {
            const int tap = kb / p.cpb0;
            const int c0 = (kb - tap * p.cpb0) * kPmBlockK;
            tma_load_5d(sA, &mapA0, &bars->full[stage], c0, w0 + p.taps0[ph][tap][0], h0 + p.taps0[ph][tap][1],
                        p.taps0[ph][tap][2], n0);
          } else {
            const int c0 = (kb - p.nkb0) * kPmBlockK;
            tma_load_5d(sA, &mapA1, &bars->full[stage], c0, w0 + p.tap1[0], h0 + p.tap1[1], p.tap1[2], n0);
          }
          tma_load_2d(sB, &mapB, &bars->full[stage], kb * kPmBlockK, wrow);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    if (lane == 0) {
      const uint32_t idesc = umma_idesc_bf16_m128((uint32_t)p.block_n);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
        const int as = it & 1;
        const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
        mbar_wait(&bars->tmem_empty[as], aphase ^ 1u);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + (uint32_t)(as * p.block_n);
        if (p.vtap) {
          // the three vertical taps are three views of the haloed tile whose first rows differ by one image row
          // (bw * 128 B, a multiple of the 1 KB swizzle repeat for bw >= 8)
          const uint32_t row_step = (uint32_t)(p.bw * 128) >> 4;
          for (int g = 0; g < ngroups; ++g) {
            mbar_wait(&bars->full[stage], phase);
            tc_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
            const uint64_t adesc0 = umma_desc_kmajor_sw128(a_addr);
            for (int r = 0; r < 3; ++r) {
              const uint64_t ad = adesc0 + (uint64_t)(r * row_step);
              const uint64_t bd = umma_desc_kmajor_sw128(a_addr + (uint32_t)(halo_bytes + r * b_bytes));
#pragma unroll
              for (int k = 0; k < kPmBlockK / 16; ++k)
                umma_bf16(tmem_d, ad + (uint64_t)(2 * k), bd + (uint64_t)(2 * k), idesc, (g > 0 || r > 0 || k > 0) ? 1u : 0u);
            }
            umma_commit(&bars->empty[stage]);
            if (++stage == p.stages) { stage = 0; phase ^= 1u; }
          }
          umma_commit(&bars->tmem_full[as]);
          continue;
        }
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&bars->full[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + (size_t)stage * stage_bytes);
          const uint64_t adesc = umma_desc_kmajor_sw128(a_addr);
          const uint64_t bdesc = umma_desc_kmajor_sw128(a_addr + kPmABytes);
#pragma unroll
          for (int k = 0; k < kPmBlockK / 16; ++k) {
            // advance 16 bf16 = 32 B inside the 128 B swizzle row: +2 in the (addr >> 4) field
            umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), idesc,
                      (kb > 0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&bars->empty[stage]);
          if (++stage == p.stages) { stage = 0; phase ^= 1u; }
        }
        umma_commit(&bars->tmem_full[as]);
      }
    }
  } else {
    // ================================ epilogue ================================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;
    const int w_l = r % p.bw;
    const int h_l = (r / p.bw) % p.bh;
    const int n_l = r / (p.bw * p.bh);
    const int hw_out = p.out_H * p.out_W;
    int it = 0;
    for (int tile = blockIdx.x; tile < p.total_tiles; tile += gridDim.x, ++it) {
      const int as = it & 1;
      const uint32_t aphase = (uint32_t)(it >> 1) & 1u;
      const int ph = tile / (p.m_tiles * p.n_tiles);
      const int rem = tile - ph * (p.m_tiles * p.n_tiles);
      const int mt = rem / p.n_tiles;
      const int nt = rem - mt * p.n_tiles;
      const int tw = mt % p.tiles_w;
      const int th = (mt / p.tiles_w) % p.tiles_h;
      const int tn = mt / (p.tiles_w * p.tiles_h);
      const int n = tn * p.bn + n_l;
      const int oy = (th * p.bh + h_l) * p.osy + (ph >> 1);
      const int ox = (tw * p.bw + w_l) * p.osx + (ph & 1);
      const bool row_ok = n < p.B;
      const size_t pix = ((size_t)n * p.out_H + oy) * p.out_W + ox;

      mbar_wait(&bars->tmem_full[as], aphase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (uint32_t)(as * p.block_n) + ((uint32_t)(q * 32) << 16);
      for (int c = 0; c < p.block_n; c += 16) {
        uint32_t v[16];
        __syncwarp();
        tmem_ld_x16(taddr + (uint32_t)c, v);
        tmem_ld_wait();
        const int col0 = nt * p.block_n + c;
        const int ncols = p.N - col0;  // valid columns in this chunk (may be <= 0 or >= 16)
        if (row_ok && ncols > 0) {
          float acc[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) acc[j] = __uint_as_float(v[j]);
          if (ncols >= 16) {
            if (p.bias) {
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j));
                acc[j] += b4.x; acc[j + 1] += b4.y; acc[j + 2] += b4.z; acc[j + 3] += b4.w;
              }
            }
            if (p.rowadd) {
              const float* ra = p.rowadd + (size_t)n * p.rowadd_ld + col0;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(ra + j));
                acc[j] += b4.x; acc[j + 1] += b4.y; acc[j + 2] += b4.z; acc[j + 3] += b4.w;
              }
            }
            if (p.residual) {
              const float* rs = p.residual + pix * (size_t)p.res_ld + col0;
#pragma unroll
              for (int j = 0; j < 16; j += 4) {
                const float4 b4 = __ldg(reinterpret_cast<const float4*>(rs + j));
                acc[j] += b4.x; acc[j + 1] += b4.y; acc[j + 2] += b4.z; acc[j + 3] += b4.w;
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              if (j < ncols) {
                if (p.bias) acc[j] += __ldg(p.bias + col0 + j);
                if (p.rowadd) acc[j] += __ldg(p.rowadd + (size_t)n * p.rowadd_ld + col0 + j);
                if (p.residual) acc[j] += __ldg(p.residual + pix * (size_t)p.res_ld + col0 + j);
              }
            }
          }
          if (p.out_mode == B200_OUT_F32_NHWC) {
            float* o = reinterpret_cast<float*>(p.out) + pix * (size_t)p.out_ld + col0;
            if (ncols >= 16) {
#pragma unroll
              for (int j = 0; j < 16; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < ncols) o[j] = acc[j];
            }
          } else if (p.out_mode == B200_OUT_BF16_NHWC) {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + pix * (size_t)p.out_ld + col0;
            if (ncols >= 16) {
              uint4 u0, u1;
              u0.x = pack_bf16x2(acc[0], acc[1]);   u0.y = pack_bf16x2(acc[2], acc[3]);
              u0.z = pack_bf16x2(acc[4], acc[5]);   u0.w = pack_bf16x2(acc[6], acc[7]);
              u1.x = pack_bf16x2(acc[8], acc[9]);   u1.y = pack_bf16x2(acc[10], acc[11]);
              u1.z = pack_bf16x2(acc[12], acc[13]); u1.w = pack_bf16x2(acc[14], acc[15]);
              *reinterpret_cast<uint4*>(o) = u0;
              *reinterpret_cast<uint4*>(o + 8) = u1;
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (j < ncols) o[j] = __float2bfloat16_rn(acc[j]);
            }
          } else if (p.out_mode == B200_OUT_F32_NCHW) {
            // channel-major per image: consecutive lanes hold consecutive pixels -> coalesced per channel
            float* o = reinterpret_cast<float*>(p.out) + ((size_t)n * p.out_ld + col0) * hw_out + oy * p.out_W + ox;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < ncols) o[(size_t)j * hw_out] = acc[j];
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) +
                               ((size_t)n * p.out_ld + col0) * hw_out + oy * p.out_W + ox;
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (j < ncols) o[(size_t)j * hw_out] = __float2bfloat16_rn(acc[j]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&bars->tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, tmem_cols);
  }
}

static int g_num_sms_pixm = 0;
extern long long g_launch_count;

int make_a_map(CUtensorMap* m, const void* base, int C, int H, int W, int planes, int B, int bw, int bh,
                      int bn) {
  uint64_t dims[5] = {(uint64_t)C, (uint64_t)W, (uint64_t)H, (uint64_t)planes, (uint64_t)B};
  uint64_t strides[4] = {(uint64_t)C * 2, (uint64_t)W * C * 2, (uint64_t)H * W * C * 2,
                         (uint64_t)planes * H * W * C * 2};
  uint32_t box[5] = {64, (uint32_t)bw, (uint32_t)bh, 1, (uint32_t)bn};
  return encode_tmap(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, base, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

}  // namespace b200


namespace b200 {
int conv2d_fwd_pixm(const b200_conv_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d != nullptr, "conv2d_fwd: null descriptor");
  B200_REQUIRE(d->a0 && d->w && d->out, "conv2d_fwd: null a0/w/out");
  B200_REQUIRE(d->a0_C > 0 && d->a0_C % 64 == 0, "conv2d_fwd: a0_C=%d must be a positive multiple of 64", d->a0_C);
  B200_REQUIRE(d->a1 == nullptr || (d->a1_C > 0 && d->a1_C % 64 == 0), "conv2d_fwd: a1_C=%d must be a multiple of 64",
               d->a1_C);
  B200_REQUIRE(d->phases == 1 || d->phases == 4, "conv2d_fwd: phases must be 1 or 4");
  B200_REQUIRE(d->ntaps0 >= 1 && d->ntaps0 <= 9, "conv2d_fwd: ntaps0=%d out of range", d->ntaps0);
  B200_REQUIRE(d->N >= 1, "conv2d_fwd: N must be positive");
  B200_REQUIRE(d->B >= 1 && d->Ho >= 1 && d->Wo >= 1, "conv2d_fwd: bad B/Ho/Wo");
  const int K = d->ntaps0 * d->a0_C + (d->a1 ? d->a1_C : 0);
  B200_REQUIRE(d->w_K == K, "conv2d_fwd: w_K=%d does not match ntaps0*a0_C+a1_C=%d", d->w_K, K);
  B200_REQUIRE(d->out_mode >= 0 && d->out_mode <= 3, "conv2d_fwd: bad out_mode");
  B200_REQUIRE(((uintptr_t)d->a0 & 127) == 0 && ((uintptr_t)d->w & 127) == 0 && ((uintptr_t)d->out & 15) == 0,
               "conv2d_fwd: a0/w must be 128-byte aligned and out 16-byte aligned");
  if (d->out_mode <= B200_OUT_BF16_NHWC)
    B200_REQUIRE(d->out_ld % 8 == 0 || d->N < 16, "conv2d_fwd: NHWC out_ld=%d must be a multiple of 8", d->out_ld);
  if (d->residual) B200_REQUIRE(d->res_ld % 4 == 0 && ((uintptr_t)d->residual & 15) == 0, "conv2d_fwd: residual alignment");
  if (d->rowadd) B200_REQUIRE(d->rowadd_ld % 4 == 0 && ((uintptr_t)d->rowadd & 15) == 0, "conv2d_fwd: rowadd alignment");
  if (d->bias) B200_REQUIRE(((uintptr_t)d->bias & 15) == 0, "conv2d_fwd: bias alignment");

  // ---- M tiling: 128 output pixels = bw x bh x bn ----
  int bw = 1;
  while (bw * 2 <= 128 && d->Wo % (bw * 2) == 0) bw *= 2;
  int bh = 1;
  while (bw * bh * 2 <= 128 && d->Ho % (bh * 2) == 0) bh *= 2;
  int bn = 128 / (bw * bh);
  B200_REQUIRE(bn == 1 || (bw == d->Wo && bh == d->Ho),
               "conv2d_fwd: unsupported spatial size %dx%d (need power-of-two factors to fill a 128-pixel tile)",
               d->Ho, d->Wo);
  // Vertical-tap reuse (B200_PIXM_VTAP=0 disables): for plain 3x3 stride-1 layers the tile becomes (<= 32) x (>= 4)
  // pixels and a pipeline stage holds ONE haloed pixel tile (bh + 2 rows) for a horizontal offset and channel chunk plus
  // the three weight tiles of its vertical taps: (bh + 2) / (3 bh) of the pixel operand bytes (0.5 at bh = 4).  The
  // one-row tiles chosen above for wide images (bw = 128, bh = 1) cannot share rows between taps.
  static const char* env_vtap = getenv("B200_PIXM_VTAP");
  bool vtap = !(env_vtap && atoi(env_vtap) == 0) && d->phases == 1 && d->ntaps0 == 9 && d->a0_planes == 1 &&
              d->a1 == nullptr && d->a0_H == d->Ho && d->a0_W == d->Wo;
  for (int k = 0; vtap && k < 9; ++k)
    vtap = d->taps0[0][k][0] == (k % 3) - 1 && d->taps0[0][k][1] == (k / 3) - 1 && d->taps0[0][k][2] == 0;
  if (vtap) {
    // experiment knob: tile width cap 8 / 16 / 32.  Measured: 16x8 and 8x16 tiles (fewer halo bytes) run exactly as
    // fast as 32x4 (50 us at 128->3@32x32, B=256), i.e. after the 2x cut the pixel operand stream is no longer the limit
    static const char* env_vbw = getenv("B200_PIXM_VBW");
    const int cap_env = env_vbw ? atoi(env_vbw) : 32;
    const int cap = (cap_env == 8 || cap_env == 16) ? cap_env : 32;
    const int vbw = bw > cap ? cap : bw, vbh = 128 / vbw;
    if (vbw >= 8 && d->Ho % vbh == 0) { bw = vbw; bh = vbh; bn = 1; }
    else vtap = false;
  }
  PixmParams p;
  memset(&p, 0, sizeof(p));
  p.B = d->B;
  p.vtap = vtap ? 1 : 0;
  p.bw = bw; p.bh = bh; p.bn = bn;
  p.tiles_w = d->Wo / bw;
  p.tiles_h = d->Ho / bh;
  const int tiles_n = (d->B + bn - 1) / bn;
  p.m_tiles = p.tiles_w * p.tiles_h * tiles_n;
  // ---- N tiling ----
  int block_n = d->N >= 256 ? 256 : ((d->N + 15) / 16) * 16;
  p.block_n = block_n;
  p.n_tiles = (d->N + block_n - 1) / block_n;
  p.total_tiles = d->phases * p.m_tiles * p.n_tiles;
  p.N = d->N;
  p.w_rows_per_phase = d->w_rows_per_phase;
  p.cpb0 = d->a0_C / 64;
  p.nkb0 = d->ntaps0 * p.cpb0;
  p.nkb1 = d->a1 ? d->a1_C / 64 : 0;
  memcpy(p.taps0, d->taps0, sizeof(p.taps0));
  memcpy(p.tap1, d->tap1, sizeof(p.tap1));
  p.bias = d->bias; p.rowadd = d->rowadd; p.rowadd_ld = d->rowadd_ld;
  p.residual = d->residual; p.res_ld = d->res_ld;
  p.out = d->out; p.out_mode = d->out_mode; p.out_ld = d->out_ld;
  p.out_H = d->out_H; p.out_W = d->out_W; p.osy = d->osy; p.osx = d->osx;

  const int stage_bytes = vtap ? (bh + 2) * bw * 128 + 3 * block_n * 128 : kPmABytes + block_n * 128;
  int stages = (227 * 1024 - 2048) / stage_bytes;
  if (stages > kPmMaxStages) stages = kPmMaxStages;
  if (stages < 2) stages = 2;
  p.stages = stages;
  const size_t smem_bytes = (size_t)stages * stage_bytes + sizeof(PixmBarriers) + 1024;

  CUtensorMap mapA0, mapA1, mapB;
  int rc = make_a_map(&mapA0, d->a0, d->a0_C, d->a0_H, d->a0_W, d->a0_planes, d->B, bw, vtap ? bh + 2 : bh, bn);
  if (rc) return rc;
  if (d->a1) {
    rc = make_a_map(&mapA1, d->a1, d->a1_C, d->a1_H, d->a1_W, d->a1_planes, d->B, bw, bh, bn);
    if (rc) return rc;
  } else {
    mapA1 = mapA0;
  }
  {
    uint64_t dims[2] = {(uint64_t)d->w_K, (uint64_t)d->w_rows};
    uint64_t strides[1] = {(uint64_t)d->w_K * 2};
    uint32_t box[2] = {64, (uint32_t)block_n};
    rc = encode_tmap(&mapB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d->w, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
  }

  if (g_num_sms_pixm == 0) {
    int dev = 0;
    B200_CHECK(cudaGetDevice(&dev));
    B200_CHECK(cudaDeviceGetAttribute(&g_num_sms_pixm, cudaDevAttrMultiProcessorCount, dev));
    B200_CHECK(cudaFuncSetAttribute(conv_gemm_pixm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
  }
  const int grid = p.total_tiles < g_num_sms_pixm ? p.total_tiles : g_num_sms_pixm;
  B200_CHECK(launch_pdl(conv_gemm_pixm_kernel, dim3(grid), dim3(kPmThreads), smem_bytes, stream, mapA0, mapA1, mapB, p));
  ++g_launch_count;
  return 0;
}
}  // namespace b200
