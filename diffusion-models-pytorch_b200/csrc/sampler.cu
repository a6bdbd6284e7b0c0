// K4: fused sampler update.  One elementwise pass replaces the ~45 ATen ops of DDPM.predict + DDPM/DDIM.denoise
// (diffusions/ddpm.py:174-261, diffusions/ddim.py:57-86) and the CFG mix (diffusions/ddim.py:177-187,
// diffusions/ddpm.py:335-347).  Arithmetic follows the reference's operation order with explicitly rounded
// (non-contracted) fp32 ops, so for fixed variances the result is bit-identical to the eager sequence given
// identical scalar coefficients.  Coefficients are read from a row in device memory (CUDA-graph replayable).
#include "common.cuh"
#include "../../include/b200diff.h"

namespace b200 {
extern long long g_launch_count;

struct SamplerK {
  const float* mo; const float* mo_u; const float* xt; const float* noise; const float* coef;
  int C, Cm, HW, objective, clip, learned_range;
  float gs_c, gs_u;  // guidance: s and (1 - s), rounded to fp32 like the reference's scalar operands
  float* sample; float* mean; float* pred_x0; float* pred_eps; float* var_out;
  size_t total;
};

__device__ __forceinline__ float predict_eps(float out, float xt, int objective, int clip, float c_rx, float c_rm1,
                                             float c_sa, float c_s1ma, float& x0) {
  if (objective == B200_OBJ_EPS) x0 = __fsub_rn(__fmul_rn(c_rx, xt), __fmul_rn(c_rm1, out));
  else if (objective == B200_OBJ_X0) x0 = out;
  else x0 = __fsub_rn(__fmul_rn(c_sa, xt), __fmul_rn(c_s1ma, out));
  if (clip) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  return __fdiv_rn(__fsub_rn(__fmul_rn(c_rx, xt), x0), c_rm1);
}

// Compile-time form of the same arithmetic for the streaming kernel (no per-element branches on uniform flags).
template <int OBJ, bool CLIP>
__device__ __forceinline__ float predict_eps_t(float out, float xt, const float c_rx, const float c_rm1, const float c_sa,
                                               const float c_s1ma, float& x0) {
  if (OBJ == B200_OBJ_EPS) x0 = __fsub_rn(__fmul_rn(c_rx, xt), __fmul_rn(c_rm1, out));
  else if (OBJ == B200_OBJ_X0) x0 = out;
  else x0 = __fsub_rn(__fmul_rn(c_sa, xt), __fmul_rn(c_s1ma, out));
  if (CLIP) x0 = fminf(fmaxf(x0, -1.0f), 1.0f);
  return __fdiv_rn(__fsub_rn(__fmul_rn(c_rx, xt), x0), c_rm1);
}

// Scalars of one step: the coefficient row is read once per thread, before the streaming loop.
struct StepCoef {
  float c_rx, c_rm1, c_sa, c_s1ma, c_x0, c_xt, c_eps, sd_fixed, min_lv, max_lv;
  bool add_noise;
};

__device__ __forceinline__ StepCoef load_step_coef(const float* __restrict__ coef) {
  StepCoef c;
  c.c_rx = coef[B200_SC_SQRT_RECIP_AC]; c.c_rm1 = coef[B200_SC_SQRT_RECIPM1_AC];
  c.c_sa = coef[B200_SC_SQRT_AC]; c.c_s1ma = coef[B200_SC_SQRT_1M_AC];
  c.c_x0 = coef[B200_SC_X0_COEF]; c.c_xt = coef[B200_SC_XT_COEF]; c.c_eps = coef[B200_SC_EPS_COEF];
  c.sd_fixed = __fsqrt_rn(coef[B200_SC_VAR]);
  c.min_lv = coef[B200_SC_MIN_LOGVAR]; c.max_lv = coef[B200_SC_MAX_LOGVAR];
  c.add_noise = coef[B200_SC_ADD_NOISE] != 0.0f;
  return c;
}

// One element of the update.  mo_u is only read when cfg, lv only when learned_range (and noise is being added).
__device__ __forceinline__ void step_element(const SamplerK& k, const StepCoef& c, bool cfg, float mo, float mo_u, float xt,
                                             float nz, float lv, float& out, float& mean, float& x0, float& eps, float& var) {
  if (cfg) {
    float tmp;
    const float eps_c = predict_eps(mo, xt, k.objective, k.clip, c.c_rx, c.c_rm1, c.c_sa, c.c_s1ma, tmp);
    const float eps_u = predict_eps(mo_u, xt, k.objective, k.clip, c.c_rx, c.c_rm1, c.c_sa, c.c_s1ma, tmp);
    const float mix = __fadd_rn(__fmul_rn(k.gs_u, eps_u), __fmul_rn(k.gs_c, eps_c));
    eps = predict_eps(mix, xt, B200_OBJ_EPS, k.clip, c.c_rx, c.c_rm1, c.c_sa, c.c_s1ma, x0);
  } else {
    eps = predict_eps(mo, xt, k.objective, k.clip, c.c_rx, c.c_rm1, c.c_sa, c.c_s1ma, x0);
  }
  mean = __fadd_rn(__fadd_rn(__fmul_rn(c.c_x0, x0), __fmul_rn(c.c_xt, xt)), __fmul_rn(c.c_eps, eps));
  out = mean;
  var = 0.0f;
  if (c.add_noise) {
    float sd = c.sd_fixed;
    if (k.learned_range) {
      const float frac = __fdiv_rn(__fadd_rn(lv, 1.0f), 2.0f);
      const float logvar = __fadd_rn(__fmul_rn(frac, c.max_lv), __fmul_rn(__fsub_rn(1.0f, frac), c.min_lv));
      var = expf(logvar);
      sd = __fsqrt_rn(var);
    }
    out = __fadd_rn(mean, __fmul_rn(sd, nz));
  }
}

// Generic form: any HW / alignment, one element per iteration.
__global__ void __launch_bounds__(256) sampler_step_kernel(const SamplerK k) {
  const StepCoef c = load_step_coef(k.coef);
  const bool cfg = k.mo_u != nullptr;
  const bool want_lv = c.add_noise && k.learned_range;
  const size_t chw = (size_t)k.C * k.HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < k.total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / chw;
    const size_t r = i - b * chw;  // c*HW + hw
    const size_t mi = b * (size_t)k.Cm * k.HW + r;
    float out, mean, x0, eps, var;
    step_element(k, c, cfg, k.mo[mi], cfg ? k.mo_u[mi] : 0.0f, k.xt[i], (c.add_noise && k.noise) ? k.noise[i] : 0.0f,
                 want_lv ? k.mo[mi + chw] : 0.0f, out, mean, x0, eps, var);
    if (want_lv && k.var_out) k.var_out[i] = var;
    if (k.sample) k.sample[i] = out;
    if (k.mean) k.mean[i] = mean;
    if (k.pred_x0) k.pred_x0[i] = x0;
    if (k.pred_eps) k.pred_eps[i] = eps;
  }
}

// Streaming form (HW % 4 == 0, 16-byte aligned tensors): a thread owns U 16-byte units per iteration, a grid stride
// apart, and issues every load of the iteration (3..5 streams x U) before the first dependent instruction, so
// >= 64 B per thread are in flight; outputs leave as 16-byte stores.  The uniform flags (CFG, objective, clip,
// learned variance) are template parameters: the first version branched on them per element and was issue-bound
// (ncu: 61 % issue slots busy at 55 % of DRAM peak, ~73 instructions per element).  Same per-element arithmetic as the
// generic kernel, hence the same bits.  The batch index (needed only when the model has 2C channels) costs one
// division per unit.  Grid = 4 resident CTAs x 148 SMs (one wave).
template <int U, bool CFG, int OBJ, bool CLIP, bool LEARNED>
__global__ void __launch_bounds__(256, 4) sampler_step_vec4_kernel(const SamplerK k) {
  const StepCoef c = load_step_coef(k.coef);
  const bool want_lv = LEARNED && c.add_noise;
  const bool want_nz = c.add_noise && k.noise != nullptr;
  const size_t chw4 = ((size_t)k.C * k.HW) >> 2, mchw4 = ((size_t)k.Cm * k.HW) >> 2;
  const size_t units = k.total >> 2;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  // The graph runner updates x in place (sample == xt): every unit is read and then written by the same thread only,
  // and all loads of an iteration are issued before its stores.  xt is therefore NOT declared __restrict__ / read-only
  // (a .nc load of memory the same kernel writes would be formally undefined); the other inputs never alias an output.
  const float4* __restrict__ mo4 = reinterpret_cast<const float4*>(k.mo);
  const float4* __restrict__ mou4 = reinterpret_cast<const float4*>(k.mo_u);
  const float4* xt4 = reinterpret_cast<const float4*>(k.xt);
  const float4* __restrict__ nz4 = reinterpret_cast<const float4*>(k.noise);
  const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
  for (size_t base = (size_t)blockIdx.x * blockDim.x + threadIdx.x; base < units; base += stride * U) {
    float4 vmo[U], vmu[U], vxt[U], vnz[U], vlv[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + (size_t)u * stride;
      vmo[u] = vmu[u] = vxt[u] = vnz[u] = vlv[u] = zero;
      if (i < units) {
        size_t mi = i;
        if (LEARNED || mchw4 != chw4) { const size_t b = i / chw4; mi = b * mchw4 + (i - b * chw4); }
        vxt[u] = xt4[i];
        vmo[u] = mo4[mi];
        if (CFG) vmu[u] = mou4[mi];
        if (want_nz) vnz[u] = nz4[i];
        if (want_lv) vlv[u] = mo4[mi + chw4];
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const size_t i = base + (size_t)u * stride;
      if (i < units) {
        const float mo[4] = {vmo[u].x, vmo[u].y, vmo[u].z, vmo[u].w}, mu[4] = {vmu[u].x, vmu[u].y, vmu[u].z, vmu[u].w};
        const float xt[4] = {vxt[u].x, vxt[u].y, vxt[u].z, vxt[u].w}, nz[4] = {vnz[u].x, vnz[u].y, vnz[u].z, vnz[u].w};
        const float lv[4] = {vlv[u].x, vlv[u].y, vlv[u].z, vlv[u].w};
        float out[4], mean[4], x0[4], eps[4], var[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          if (CFG) {
            float tmp;
            const float eps_c = predict_eps_t<OBJ, CLIP>(mo[e], xt[e], c.c_rx, c.c_rm1, c.c_sa, c.c_s1ma, tmp);
            const float eps_u = predict_eps_t<OBJ, CLIP>(mu[e], xt[e], c.c_rx, c.c_rm1, c.c_sa, c.c_s1ma, tmp);
            const float mix = __fadd_rn(__fmul_rn(k.gs_u, eps_u), __fmul_rn(k.gs_c, eps_c));
            eps[e] = predict_eps_t<B200_OBJ_EPS, CLIP>(mix, xt[e], c.c_rx, c.c_rm1, c.c_sa, c.c_s1ma, x0[e]);
          } else {
            eps[e] = predict_eps_t<OBJ, CLIP>(mo[e], xt[e], c.c_rx, c.c_rm1, c.c_sa, c.c_s1ma, x0[e]);
          }
          mean[e] = __fadd_rn(__fadd_rn(__fmul_rn(c.c_x0, x0[e]), __fmul_rn(c.c_xt, xt[e])), __fmul_rn(c.c_eps, eps[e]));
          out[e] = mean[e];
          var[e] = 0.0f;
          if (c.add_noise) {
            float sd = c.sd_fixed;
            if (LEARNED) {
              const float frac = __fdiv_rn(__fadd_rn(lv[e], 1.0f), 2.0f);
              const float logvar = __fadd_rn(__fmul_rn(frac, c.max_lv), __fmul_rn(__fsub_rn(1.0f, frac), c.min_lv));
              var[e] = expf(logvar);
              sd = __fsqrt_rn(var[e]);
            }
            out[e] = __fadd_rn(mean[e], __fmul_rn(sd, nz[e]));
          }
        }
        if (want_lv && k.var_out) reinterpret_cast<float4*>(k.var_out)[i] = make_float4(var[0], var[1], var[2], var[3]);
        if (k.sample) reinterpret_cast<float4*>(k.sample)[i] = make_float4(out[0], out[1], out[2], out[3]);
        if (k.mean) reinterpret_cast<float4*>(k.mean)[i] = make_float4(mean[0], mean[1], mean[2], mean[3]);
        if (k.pred_x0) reinterpret_cast<float4*>(k.pred_x0)[i] = make_float4(x0[0], x0[1], x0[2], x0[3]);
        if (k.pred_eps) reinterpret_cast<float4*>(k.pred_eps)[i] = make_float4(eps[0], eps[1], eps[2], eps[3]);
      }
    }
  }
}

template <int U, bool CFG, int OBJ, bool CLIP>
static void launch_vec4_l(const SamplerK& k, int g, cudaStream_t stream) {
  if (k.learned_range) sampler_step_vec4_kernel<U, CFG, OBJ, CLIP, true><<<g, 256, 0, stream>>>(k);
  else sampler_step_vec4_kernel<U, CFG, OBJ, CLIP, false><<<g, 256, 0, stream>>>(k);
}
template <int U, bool CFG, int OBJ>
static void launch_vec4_c(const SamplerK& k, int g, cudaStream_t stream) {
  if (k.clip) launch_vec4_l<U, CFG, OBJ, true>(k, g, stream);
  else launch_vec4_l<U, CFG, OBJ, false>(k, g, stream);
}
template <int U, bool CFG>
static void launch_vec4_o(const SamplerK& k, int g, cudaStream_t stream) {
  if (k.objective == B200_OBJ_EPS) launch_vec4_c<U, CFG, B200_OBJ_EPS>(k, g, stream);
  else if (k.objective == B200_OBJ_X0) launch_vec4_c<U, CFG, B200_OBJ_X0>(k, g, stream);
  else launch_vec4_c<U, CFG, B200_OBJ_V>(k, g, stream);
}
template <int U>
static void launch_vec4(const SamplerK& k, int g, cudaStream_t stream) {
  if (k.mo_u) launch_vec4_o<U, true>(k, g, stream);
  else launch_vec4_o<U, false>(k, g, stream);
}

// Euler / Heun steps in sigma space (diffusions/euler.py:50-66, diffusions/heun.py:56-107; Karras et al. 2022):
//   x0 = predict(model_out, x, t_eval);  bar = sqrt(1 + s_eval^2) x;  d = (bar - x0) / s_eval
//   first order : sample = (bar + d (s_prev - s_t)) / sqrt(1 + s_prev^2)            (x = x_t, s_eval = s_t)
//   second order: d = (d + d1) / 2; bar_t = sqrt(1 + s_t^2) x1; sample = (bar_t + d (s_prev - s_t)) / sqrt(1 + s_prev^2)
//                 (x = the first-order sample, s_eval = s_prev, d1 / x1 = derivative and x_t of the first-order step)
struct OdeK {
  const float* mo; const float* x; const float* d1; const float* x1;
  int C, Cm, HW, objective, clip, second;
  float c_rx, c_rm1, c_sa, c_s1ma, sig_t, sig_prev;
  float* sample; float* pred_x0; float* deriv;
  size_t total;
};

__global__ void __launch_bounds__(256) ode_step_kernel(const OdeK k) {
  const size_t chw = (size_t)k.C * k.HW;
  const float sig_eval = k.second ? k.sig_prev : k.sig_t;
  const float up_eval = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(sig_eval, sig_eval)));
  const float up_t = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(k.sig_t, k.sig_t)));
  const float down = __fsqrt_rn(__fadd_rn(1.0f, __fmul_rn(k.sig_prev, k.sig_prev)));
  const float dsig = __fsub_rn(k.sig_prev, k.sig_t);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < k.total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / chw;
    const size_t mi = b * (size_t)k.Cm * k.HW + (i - b * chw);
    const float x = k.x[i];
    float x0;
    predict_eps(k.mo[mi], x, k.objective, k.clip, k.c_rx, k.c_rm1, k.c_sa, k.c_s1ma, x0);
    const float bar = __fmul_rn(up_eval, x);
    float d = __fdiv_rn(__fsub_rn(bar, x0), sig_eval);
    float base = bar;
    if (k.second) {
      d = __fdiv_rn(__fadd_rn(d, k.d1[i]), 2.0f);
      base = __fmul_rn(up_t, k.x1[i]);
    }
    const float s = __fdiv_rn(__fadd_rn(base, __fmul_rn(d, dsig)), down);
    if (k.sample) k.sample[i] = s;
    if (k.pred_x0) k.pred_x0[i] = x0;
    if (k.deriv) k.deriv[i] = d;
  }
}

// Output stage after the sampling loop (scripts/sample_uncond.py:189-195 + utils/misc.py image_norm_to_float +
// torchvision.utils.save_image's quantisation): fp32 NCHW in [-1, 1] -> uint8 NHWC, one pass on the device:
//   u8 = trunc(clamp(((clamp(x, -1, 1) + 1) / 2) * 255 + 0.5, 0, 255))
// so that only B*H*W*C bytes (a quarter of the fp32 tensor, already in PNG row order) cross PCIe.
__global__ void __launch_bounds__(256) to_uint8_hwc_kernel(const float* __restrict__ x, uint8_t* __restrict__ out, int B, int C,
                                                           int HW) {
  const size_t total = (size_t)B * HW * C;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const size_t pix = i / C;
    const int hw = (int)(pix % HW);
    const size_t b = pix / HW;
    float v = x[(b * C + c) * HW + hw];
    v = fminf(fmaxf(v, -1.0f), 1.0f);
    v = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
    v = __fadd_rn(__fmul_rn(v, 255.0f), 0.5f);
    v = fminf(fmaxf(v, 0.0f), 255.0f);
    out[i] = (uint8_t)v;
  }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_to_uint8_hwc(const float* x, uint8_t* out, int B, int C, int HW, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(x && out && B >= 1 && C >= 1 && HW >= 1, "to_uint8_hwc: bad arguments");
  size_t g = ((size_t)B * C * HW + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  to_uint8_hwc_kernel<<<(int)g, 256, 0, stream>>>(x, out, B, C, HW);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "to_uint8_hwc launch");
}

extern "C" int b200_ode_step(const b200_ode_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d && d->model_out && d->x, "ode_step: null model_out/x");
  B200_REQUIRE(d->Cm >= d->C, "ode_step: model channels %d < C=%d", d->Cm, d->C);
  B200_REQUIRE(d->objective >= 0 && d->objective <= 2, "ode_step: bad objective");
  B200_REQUIRE(!d->second_order || (d->d1 && d->x1), "ode_step: the second-order step needs d1 and x1");
  OdeK k;
  k.mo = d->model_out; k.x = d->x; k.d1 = d->d1; k.x1 = d->x1;
  k.C = d->C; k.Cm = d->Cm; k.HW = d->HW; k.objective = d->objective; k.clip = d->clip; k.second = d->second_order;
  k.c_rx = d->sqrt_recip_ac; k.c_rm1 = d->sqrt_recipm1_ac; k.c_sa = d->sqrt_ac; k.c_s1ma = d->sqrt_1m_ac;
  k.sig_t = d->sigma_t; k.sig_prev = d->sigma_prev;
  k.sample = d->sample; k.pred_x0 = d->pred_x0; k.deriv = d->deriv;
  k.total = (size_t)d->B * d->C * d->HW;
  size_t g = (k.total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g == 0) g = 1;
  ode_step_kernel<<<(int)g, 256, 0, stream>>>(k);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "ode_step launch");
}

extern "C" int b200_sampler_step(const b200_sampler_desc* d, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  B200_REQUIRE(d && d->model_out && d->xt && d->coef, "sampler_step: null model_out/xt/coef");
  B200_REQUIRE(d->Cm == d->C || d->Cm == 2 * d->C, "sampler_step: model channels %d must be C or 2C (C=%d)", d->Cm, d->C);
  B200_REQUIRE(!d->learned_range || d->Cm == 2 * d->C, "sampler_step: learned_range needs 2C model channels");
  B200_REQUIRE(d->objective >= 0 && d->objective <= 2, "sampler_step: bad objective");
  SamplerK k;
  k.mo = d->model_out; k.mo_u = d->model_out_uncond; k.xt = d->xt; k.noise = d->noise; k.coef = d->coef;
  k.C = d->C; k.Cm = d->Cm; k.HW = d->HW; k.objective = d->objective; k.clip = d->clip;
  k.learned_range = d->learned_range;
  k.gs_c = (float)d->guidance_scale;
  k.gs_u = (float)(1.0 - d->guidance_scale);
  k.sample = d->sample; k.mean = d->mean; k.pred_x0 = d->pred_x0; k.pred_eps = d->pred_eps; k.var_out = d->var_out;
  k.total = (size_t)d->B * d->C * d->HW;
  auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; };
  const bool vec = d->HW % 4 == 0 && al16(k.mo) && al16(k.mo_u) && al16(k.xt) && al16(k.noise) && al16(k.sample) &&
                   al16(k.mean) && al16(k.pred_x0) && al16(k.pred_eps) && al16(k.var_out);
  if (vec) {
    // two units per thread per iteration once the tensor gives every thread of a one-wave grid more than one
    const size_t units = k.total / 4;
    const int U = units > (size_t)148 * 4 * 256 ? 2 : 1;
    size_t g = (units + (size_t)256 * U - 1) / ((size_t)256 * U);
    if (g > 148 * 4) g = 148 * 4;
    if (g == 0) g = 1;
    if (U == 2) launch_vec4<2>(k, (int)g, stream);
    else launch_vec4<1>(k, (int)g, stream);
  } else {
    size_t g = (k.total + 255) / 256;
    if (g > 148 * 8) g = 148 * 8;
    if (g == 0) g = 1;
    sampler_step_kernel<<<(int)g, 256, 0, stream>>>(k);
  }
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "sampler_step launch");
}
