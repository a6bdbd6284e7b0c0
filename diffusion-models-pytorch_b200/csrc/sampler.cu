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

__global__ void __launch_bounds__(256) sampler_step_kernel(const SamplerK k) {
  const float c_rx = k.coef[B200_SC_SQRT_RECIP_AC], c_rm1 = k.coef[B200_SC_SQRT_RECIPM1_AC];
  const float c_sa = k.coef[B200_SC_SQRT_AC], c_s1ma = k.coef[B200_SC_SQRT_1M_AC];
  const float c_x0 = k.coef[B200_SC_X0_COEF], c_xt = k.coef[B200_SC_XT_COEF], c_eps = k.coef[B200_SC_EPS_COEF];
  const float var_fixed = k.coef[B200_SC_VAR];
  const float min_lv = k.coef[B200_SC_MIN_LOGVAR], max_lv = k.coef[B200_SC_MAX_LOGVAR];
  const bool add_noise = k.coef[B200_SC_ADD_NOISE] != 0.0f;
  const float sd_fixed = __fsqrt_rn(var_fixed);
  const size_t chw = (size_t)k.C * k.HW;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < k.total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t b = i / chw;
    const size_t r = i - b * chw;  // c*HW + hw
    const size_t mi = b * (size_t)k.Cm * k.HW + r;
    const float xt = k.xt[i];
    float x0, eps;
    if (k.mo_u) {
      float tmp;
      const float eps_c = predict_eps(k.mo[mi], xt, k.objective, k.clip, c_rx, c_rm1, c_sa, c_s1ma, tmp);
      const float eps_u = predict_eps(k.mo_u[mi], xt, k.objective, k.clip, c_rx, c_rm1, c_sa, c_s1ma, tmp);
      const float mix = __fadd_rn(__fmul_rn(k.gs_u, eps_u), __fmul_rn(k.gs_c, eps_c));
      eps = predict_eps(mix, xt, B200_OBJ_EPS, k.clip, c_rx, c_rm1, c_sa, c_s1ma, x0);
    } else {
      eps = predict_eps(k.mo[mi], xt, k.objective, k.clip, c_rx, c_rm1, c_sa, c_s1ma, x0);
    }
    const float mean = __fadd_rn(__fadd_rn(__fmul_rn(c_x0, x0), __fmul_rn(c_xt, xt)), __fmul_rn(c_eps, eps));
    float out = mean;
    if (add_noise) {
      float sd = sd_fixed;
      if (k.learned_range) {
        const float lv = k.mo[mi + chw];
        const float frac = __fdiv_rn(__fadd_rn(lv, 1.0f), 2.0f);
        const float logvar = __fadd_rn(__fmul_rn(frac, max_lv), __fmul_rn(__fsub_rn(1.0f, frac), min_lv));
        const float var = expf(logvar);
        if (k.var_out) k.var_out[i] = var;
        sd = __fsqrt_rn(var);
      }
      const float nz = k.noise ? k.noise[i] : 0.0f;
      out = __fadd_rn(mean, __fmul_rn(sd, nz));
    }
    if (k.sample) k.sample[i] = out;
    if (k.mean) k.mean[i] = mean;
    if (k.pred_x0) k.pred_x0[i] = x0;
    if (k.pred_eps) k.pred_eps[i] = eps;
  }
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
  size_t g = (k.total + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g == 0) g = 1;
  sampler_step_kernel<<<(int)g, 256, 0, stream>>>(k);
  ++g_launch_count;
  return check_cuda(cudaGetLastError(), "sampler_step launch");
}
