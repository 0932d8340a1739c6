// diffusion.cu — the Gaussian-diffusion arithmetic of one sampling / training step, fused.
// The reference spends ~45 ATen launches and ~9 host->device table copies per sampling step
// (SURVEY.md §2.2); here it is one launch reading device-resident f32 tables.
//
// Every arithmetic op is written with explicit round-to-nearest intrinsics (__fmul_rn, ...) in
// the reference's operation order so the compiler cannot contract them into FMAs: apart from
// expf/tanhf/logf (libm vs CUDA differ by <= 2 ulp) results are bit-identical to the
// reference's eager f32 ops.
#include "common.cuh"

namespace ditb200 {

__device__ __forceinline__ float fmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float fadd(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsub(float a, float b) { return __fsub_rn(a, b); }

// model_out element (b, ch, i) after the optional CFG combine (models_original.py:258-266)
__device__ __forceinline__ float fetch_model_out(const float* __restrict__ mo, int b, int ch, int i,
                                                 int C2, int HW, int cfg_half, int n_cfg_ch,
                                                 float cfg_scale) {
  if (cfg_half > 0 && ch < n_cfg_ch) {
    const int bc = (b >= cfg_half) ? b - cfg_half : b;
    const float c = mo[((size_t)bc * C2 + ch) * HW + i];
    const float u = mo[((size_t)(bc + cfg_half) * C2 + ch) * HW + i];
    return fadd(u, fmul(cfg_scale, fsub(c, u)));  // uncond + s * (cond - uncond)
  }
  return mo[((size_t)b * C2 + ch) * HW + i];
}

__global__ void cfg_combine_kernel(const float* __restrict__ raw, float* __restrict__ out,
                                   int n_half, int C2, int HW, int n_cfg_ch, float s) {
  // one thread per element of the FIRST half; it produces both halves' values
  const size_t total = (size_t)n_half * C2 * HW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % HW);
    const int ch = (int)((idx / HW) % C2);
    const size_t other = idx + (size_t)n_half * C2 * HW;
    const float c = raw[idx], u = raw[other];
    if (ch < n_cfg_ch) {
      const float e = fadd(u, fmul(s, fsub(c, u)));
      out[idx] = e;
      out[other] = e;
    } else {
      out[idx] = c;
      out[other] = u;
    }
    (void)i;
  }
}

// ------------------------------------------------------------------ p_sample step
// One thread handles 4 consecutive spatial positions of one (b, c) plane (HW % 4 == 0 fast
// path, scalar otherwise); all loads are 128-bit and coalesced.  Per-sample table scalars are
// fetched once per thread (L1/L2-resident, a few hundred bytes in total).
struct StepCoef {
  float srac, srm1, c1, c2, min_log, max_log, nonzero;
  float ab, ab_prev;  // DDIM
  float ab_next;      // DDIM reverse
};

__device__ __forceinline__ StepCoef load_coef(const ditb200_step_args& a, int b) {
  long long t = a.t[b];
  if (t < 0) t = 0;
  if (t >= a.num_timesteps) t = a.num_timesteps - 1;
  StepCoef k;
  k.srac = a.sqrt_recip_alphas_cumprod[t];
  k.srm1 = a.sqrt_recipm1_alphas_cumprod[t];
  k.c1 = a.posterior_mean_coef1[t];
  k.c2 = a.posterior_mean_coef2[t];
  k.min_log = a.min_log[t];
  k.max_log = (a.var_type == DITB200_VAR_LEARNED_RANGE) ? a.max_log[t] : 0.f;
  k.nonzero = (t != 0) ? 1.f : 0.f;
  k.ab = k.ab_prev = k.ab_next = 1.f;
  if (a.sampler == DITB200_SAMPLER_DDIM) {
    k.ab = a.alphas_cumprod[t];
    k.ab_prev = a.alphas_cumprod_prev[t];
  } else if (a.sampler == DITB200_SAMPLER_DDIM_REVERSE) {
    k.ab_next = a.alphas_cumprod_next[t];
  }
  return k;
}

__device__ __forceinline__ void step_math(const ditb200_step_args& a, const StepCoef& k, float mo,
                                          float v, float x, float noise, const float* mean_in, float& sample,
                                          float& pred, float& mean, float& logvar) {
  // variance (gaussian_diffusion.py:285-308)
  if (a.var_type == DITB200_VAR_LEARNED_RANGE) {
    const float frac = __fdiv_rn(fadd(v, 1.0f), 2.0f);
    logvar = fadd(fmul(frac, k.max_log), fmul(fsub(1.0f, frac), k.min_log));
  } else if (a.var_type == DITB200_VAR_LEARNED) {
    logvar = v;
  } else {
    logvar = k.min_log;
  }
  // x0 prediction (:317-322, :334-339)
  if (a.mean_type == DITB200_MEAN_START_X)
    pred = mo;
  else
    pred = fsub(fmul(k.srac, x), fmul(k.srm1, mo));
  if (a.clip_denoised) pred = fminf(fmaxf(pred, -1.0f), 1.0f);
  // posterior mean (:238-241)
  mean = fadd(fmul(k.c1, pred), fmul(k.c2, x));
  if (a.sampler == DITB200_SAMPLER_DDIM) {
    // ddim_sample (:541-560): eps re-derived from pred_xstart, Eq. 12 of the DDIM paper
    const float eps = __fdiv_rn(fsub(fmul(k.srac, x), pred), k.srm1);
    const float sigma = fmul(fmul(a.eta, sqrtf(__fdiv_rn(fsub(1.0f, k.ab_prev), fsub(1.0f, k.ab)))),
                             sqrtf(fsub(1.0f, __fdiv_rn(k.ab, k.ab_prev))));
    const float mean_pred = fadd(fmul(pred, sqrtf(k.ab_prev)),
                                 fmul(sqrtf(fsub(fsub(1.0f, k.ab_prev), fmul(sigma, sigma))), eps));
    sample = fadd(mean_pred, fmul(fmul(k.nonzero, sigma), noise));
    return;
  }
  if (a.sampler == DITB200_SAMPLER_DDIM_REVERSE) {
    // ddim_reverse_sample (:583-598): the deterministic ODE step towards x_{t+1}
    const float eps = __fdiv_rn(fsub(fmul(k.srac, x), pred), k.srm1);
    sample = fadd(fmul(pred, sqrtf(k.ab_next)), fmul(sqrtf(fsub(1.0f, k.ab_next)), eps));
    return;
  }
  // ancestral update (:410-416): mean + nonzero_mask * exp(0.5 * logvar) * noise; with classifier guidance the
  // mean is the caller's conditioned one (:398-401)
  sample = fadd(mean_in ? *mean_in : mean, fmul(fmul(k.nonzero, expf(fmul(0.5f, logvar))), noise));
}

__global__ void __launch_bounds__(256) p_sample_step_kernel(const ditb200_step_args a) {
  const int C = a.C, HW = a.HW;
  const int C2 = (a.var_type == DITB200_VAR_FIXED) ? C : 2 * C;
  const int hw4 = HW >> 2;
  const size_t total4 = (size_t)a.B * C * hw4;
  const bool has_noise = a.noise != nullptr;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total4;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i4 = (int)(idx % hw4);
    const int c = (int)((idx / hw4) % C);
    const int b = (int)(idx / ((size_t)hw4 * C));
    const StepCoef k = load_coef(a, b);
    const size_t off = ((size_t)b * C + c) * HW + (size_t)i4 * 4;
    const float4 x4 = *reinterpret_cast<const float4*>(a.x + off);
    float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
    if (has_noise) n4 = *reinterpret_cast<const float4*>(a.noise + off);
    float mo[4], vv[4];
    if (a.cfg_half > 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        mo[j] = fetch_model_out(a.model_out, b, c, i4 * 4 + j, C2, HW, a.cfg_half, a.n_cfg_ch,
                                a.cfg_scale);
        vv[j] = (C2 > C) ? fetch_model_out(a.model_out, b, C + c, i4 * 4 + j, C2, HW, a.cfg_half,
                                           a.n_cfg_ch, a.cfg_scale)
                         : 0.f;
      }
    } else {
      const float4 m4 =
          *reinterpret_cast<const float4*>(a.model_out + ((size_t)b * C2 + c) * HW + (size_t)i4 * 4);
      mo[0] = m4.x, mo[1] = m4.y, mo[2] = m4.z, mo[3] = m4.w;
      if (C2 > C) {
        const float4 v4 = *reinterpret_cast<const float4*>(
            a.model_out + ((size_t)b * C2 + C + c) * HW + (size_t)i4 * 4);
        vv[0] = v4.x, vv[1] = v4.y, vv[2] = v4.z, vv[3] = v4.w;
      } else {
        vv[0] = vv[1] = vv[2] = vv[3] = 0.f;
      }
    }
    const float xs[4] = {x4.x, x4.y, x4.z, x4.w};
    const float ns[4] = {n4.x, n4.y, n4.z, n4.w};
    float s[4], p[4], m[4], lv[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)
      step_math(a, k, mo[j], vv[j], xs[j], ns[j], a.mean_override ? a.mean_override + off + j : nullptr, s[j], p[j], m[j], lv[j]);
    *reinterpret_cast<float4*>(a.sample + off) = make_float4(s[0], s[1], s[2], s[3]);
    if (a.pred_xstart) *reinterpret_cast<float4*>(a.pred_xstart + off) = make_float4(p[0], p[1], p[2], p[3]);
    if (a.mean) *reinterpret_cast<float4*>(a.mean + off) = make_float4(m[0], m[1], m[2], m[3]);
    if (a.log_variance)
      *reinterpret_cast<float4*>(a.log_variance + off) = make_float4(lv[0], lv[1], lv[2], lv[3]);
    if (a.variance) {
      if (a.var_table != nullptr) {
        long long tt = a.t[b];
        tt = tt < 0 ? 0 : (tt >= a.num_timesteps ? a.num_timesteps - 1 : tt);
        const float vt = a.var_table[tt];
        *reinterpret_cast<float4*>(a.variance + off) = make_float4(vt, vt, vt, vt);
      } else {
        *reinterpret_cast<float4*>(a.variance + off) = make_float4(expf(lv[0]), expf(lv[1]), expf(lv[2]), expf(lv[3]));
      }
    }
  }
}

// scalar fallback for HW % 4 != 0 or misaligned pointers
__global__ void __launch_bounds__(256) p_sample_step_scalar_kernel(const ditb200_step_args a) {
  const int C = a.C, HW = a.HW;
  const int C2 = (a.var_type == DITB200_VAR_FIXED) ? C : 2 * C;
  const size_t total = (size_t)a.B * C * HW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % HW);
    const int c = (int)((idx / HW) % C);
    const int b = (int)(idx / ((size_t)HW * C));
    const StepCoef k = load_coef(a, b);
    const float mo =
        fetch_model_out(a.model_out, b, c, i, C2, HW, a.cfg_half, a.n_cfg_ch, a.cfg_scale);
    const float v = (C2 > C) ? fetch_model_out(a.model_out, b, C + c, i, C2, HW, a.cfg_half,
                                               a.n_cfg_ch, a.cfg_scale)
                             : 0.f;
    const float nz = a.noise ? a.noise[idx] : 0.f;
    float s, p, m, lv;
    step_math(a, k, mo, v, a.x[idx], nz, a.mean_override ? a.mean_override + idx : nullptr, s, p, m, lv);
    a.sample[idx] = s;
    if (a.pred_xstart) a.pred_xstart[idx] = p;
    if (a.mean) a.mean[idx] = m;
    if (a.log_variance) a.log_variance[idx] = lv;
    if (a.variance) {
      long long tt = a.t[b];
      tt = tt < 0 ? 0 : (tt >= a.num_timesteps ? a.num_timesteps - 1 : tt);
      a.variance[idx] = (a.var_table != nullptr) ? a.var_table[tt] : expf(lv);
    }
  }
}

// ------------------------------------------------------------------------ q_sample
__global__ void q_sample_kernel(const float* __restrict__ x0, const float* __restrict__ noise,
                                const int64_t* __restrict__ t, const float* __restrict__ sac,
                                const float* __restrict__ s1mac, float* __restrict__ x_t, int B,
                                int CHW, int nt) {
  const size_t total = (size_t)B * CHW;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / CHW);
    long long tt = t[b];
    tt = tt < 0 ? 0 : (tt >= nt ? nt - 1 : tt);
    x_t[idx] = fadd(fmul(sac[tt], x0[idx]), fmul(s1mac[tt], noise[idx]));
  }
}

// ------------------------------------------------------------ per-sample affine helper
// out = (ta[t] * a (+|-) tb[t] * b * b2) / td[t]; see include/ditb200.h.  Separate roundings, reference order.
__global__ void __launch_bounds__(256) diffusion_affine_kernel(
    const float* __restrict__ a, const float* __restrict__ b, const float* __restrict__ b2, const int64_t* __restrict__ t,
    const float* __restrict__ ta, const float* __restrict__ tb, const float* __restrict__ td, int subtract,
    float* __restrict__ out, int B, int n, int nt) {
  const size_t total = (size_t)B * n;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
    const int bi = (int)(idx / n);
    long long tt = t[bi];
    tt = tt < 0 ? 0 : (tt >= nt ? nt - 1 : tt);
    float r;
    if (a == nullptr) r = ta[tt];
    else r = ta ? fmul(ta[tt], a[idx]) : a[idx];
    if (b != nullptr) {
      float u = b[idx];
      if (tb) u = fmul(tb[tt], u);
      if (b2) u = fmul(u, b2[idx]);
      r = subtract ? fsub(r, u) : fadd(r, u);
    }
    if (td) r = __fdiv_rn(r, td[tt]);
    out[idx] = r;
  }
}

// _prior_bpd: one CTA per sample
__global__ void __launch_bounds__(256) prior_bpd_kernel(const float* __restrict__ x0, float coef, float lv1,
                                                        float* __restrict__ out, int n) {
  __shared__ float red[8];
  const float* x = x0 + (size_t)blockIdx.x * n;
  const float e1 = expf(lv1);
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const float m = coef * x[i];
    s += 0.5f * (-1.0f - lv1 + e1 + m * m);  // normal_kl(mean1, logvar1, 0, 0)  (diffusion_utils.py:10-36)
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float v = threadIdx.x < 8 ? red[threadIdx.x] : 0.f;
    v = warp_sum(v);
    if (threadIdx.x == 0) out[blockIdx.x] = v / (float)n * 1.4426950408889634f;
  }
}

// ----------------------------------------------------------------- training losses
// One CTA per sample: mse = mean((noise - eps)^2), vb = mean(KL or decoder NLL)/ln 2 with the
// mean prediction detached, loss = mse + vb, plus d loss / d model_out.  Row reductions are
// warp shuffles followed by one shared-memory pass over the 8 warps.
__device__ __forceinline__ float approx_cdf(float x) {
  // 0.5 * (1 + tanh(sqrt(2/pi) * (x + 0.044715 x^3)))   (diffusion_utils.py:38-43)
  const float k0 = 0.7978845608028654f;
  return 0.5f * (1.0f + tanhf(k0 * (x + 0.044715f * x * x * x)));
}
__device__ __forceinline__ float approx_cdf_grad(float x) {
  const float k0 = 0.7978845608028654f;
  const float u = k0 * (x + 0.044715f * x * x * x);
  const float th = tanhf(u);
  return 0.5f * (1.0f - th * th) * k0 * (1.0f + 3.0f * 0.044715f * x * x);
}

__global__ void __launch_bounds__(256) training_losses_kernel(const ditb200_loss_args a) {
  __shared__ float red[4][8];
  const int b = blockIdx.x;
  const int C = a.C, HW = a.HW, N = C * HW;
  const bool has_v = a.var_type != DITB200_VAR_FIXED;
  const int C2 = has_v ? 2 * C : C;
  long long t = a.t[b];
  t = t < 0 ? 0 : (t >= a.num_timesteps ? a.num_timesteps - 1 : t);
  const float srac = a.sqrt_recip_alphas_cumprod[t], srm1 = a.sqrt_recipm1_alphas_cumprod[t];
  const float c1 = a.posterior_mean_coef1[t], c2 = a.posterior_mean_coef2[t];
  const float min_log = a.posterior_log_variance_clipped[t], max_log = a.log_betas[t];
  const float fixed_lv = (a.var_type == DITB200_VAR_FIXED) ? a.fixed_log_var[t] : 0.f;
  const bool t0 = (t == 0);
  const bool start_x = a.mean_type == DITB200_MEAN_START_X;
  const float inv_ln2 = 1.4426950408889634f;
  const float inv_n = 1.0f / (float)N;
  const float* mo_mean = a.model_out + (size_t)b * C2 * HW;
  const float* mo_v = mo_mean + N;
  const float* x0 = a.x0 + (size_t)b * N;
  const float* xt = a.x_t + (size_t)b * N;
  const float* nz = a.noise + (size_t)b * N;
  float* g_mean = a.grad_model_out ? a.grad_model_out + (size_t)b * C2 * HW : nullptr;
  float* g_v = (g_mean && has_v) ? g_mean + N : nullptr;
  float* pred_out = a.pred_xstart ? a.pred_xstart + (size_t)b * N : nullptr;
  const float wm = g_mean ? a.w_mse[b] : 0.f, wv = g_mean ? a.w_vb[b] * a.vb_scale * inv_n * inv_ln2 : 0.f;
  float s_mse = 0.f, s_vb = 0.f, s_x = 0.f, s_e = 0.f;
  for (int i = threadIdx.x; i < N; i += blockDim.x) {
    const float m_out = mo_mean[i], v = has_v ? mo_v[i] : 0.f, xs = x0[i], x = xt[i], n = nz[i];
    const float target = start_x ? xs : n;  // MSE target (gaussian_diffusion.py:771-777)
    const float d = target - m_out;
    s_mse += d * d;
    // q(x_{t-1} | x_t, x_0)
    const float true_mean = c1 * xs + c2 * x;
    const float lv1 = min_log;
    // p(x_{t-1} | x_t)
    float lv2, dlv2_dv;
    if (a.var_type == DITB200_VAR_LEARNED_RANGE) {
      const float frac = (v + 1.0f) * 0.5f;
      lv2 = frac * max_log + (1.0f - frac) * min_log;
      dlv2_dv = 0.5f * (max_log - min_log);
    } else if (a.var_type == DITB200_VAR_LEARNED) {
      lv2 = v, dlv2_dv = 1.0f;
    } else {
      lv2 = fixed_lv, dlv2_dv = 0.f;
    }
    float pred = start_x ? m_out : srac * x - srm1 * m_out;
    float dpred = start_x ? 1.0f : -srm1;
    if (a.clip_denoised) {
      if (pred < -1.0f || pred > 1.0f) dpred = 0.f;  // clamp passes the gradient on [-1, 1] only
      pred = fminf(fmaxf(pred, -1.0f), 1.0f);
    }
    const float mean = c1 * pred + c2 * x;
    float term, dterm_dlv2, dterm_dmean;
    if (!t0) {
      const float e12 = expf(lv1 - lv2), e2 = expf(-lv2);
      const float dm = true_mean - mean;
      term = 0.5f * (-1.0f + lv2 - lv1 + e12 + dm * dm * e2);
      dterm_dlv2 = 0.5f * (1.0f - e12 - dm * dm * e2);
      dterm_dmean = -dm * e2;
    } else {
      // decoder NLL = -discretized_gaussian_log_likelihood(x0, mean, 0.5*lv2)
      const float cx = xs - mean;
      const float inv_std = expf(-0.5f * lv2);
      const float plus_in = inv_std * (cx + 1.0f / 255.0f);
      const float min_in = inv_std * (cx - 1.0f / 255.0f);
      const float cdf_plus = approx_cdf(plus_in), cdf_min = approx_cdf(min_in);
      const float gp = approx_cdf_grad(plus_in), gm = approx_cdf_grad(min_in);
      // d(plus_in)/d(lv2) = -0.5 * plus_in, d(plus_in)/d(mean) = -inv_std; same for min_in
      const float dplus = gp * (-0.5f * plus_in), dmin = gm * (-0.5f * min_in);
      const float mplus = -gp * inv_std, mmin = -gm * inv_std;
      float logp, dlogp, mlogp;
      if (xs < -0.999f) {
        const float cl = fmaxf(cdf_plus, 1e-12f);
        logp = logf(cl);
        dlogp = (cdf_plus > 1e-12f) ? dplus / cl : 0.f;
        mlogp = (cdf_plus > 1e-12f) ? mplus / cl : 0.f;
      } else if (xs > 0.999f) {
        const float om = 1.0f - cdf_min;
        const float cl = fmaxf(om, 1e-12f);
        logp = logf(cl);
        dlogp = (om > 1e-12f) ? -dmin / cl : 0.f;
        mlogp = (om > 1e-12f) ? -mmin / cl : 0.f;
      } else {
        const float dl = cdf_plus - cdf_min;
        const float cl = fmaxf(dl, 1e-12f);
        logp = logf(cl);
        dlogp = (dl > 1e-12f) ? (dplus - dmin) / cl : 0.f;
        mlogp = (dl > 1e-12f) ? (mplus - mmin) / cl : 0.f;
      }
      term = -logp;
      dterm_dlv2 = -dlogp;
      dterm_dmean = -mlogp;
    }
    s_vb += term;
    if (pred_out) pred_out[i] = pred;
    if (a.xstart_mse) {
      const float dx = pred - xs;
      s_x += dx * dx;
    }
    if (a.eps_mse) {
      const float de = (srac * x - pred) / srm1 - n;  // _predict_eps_from_xstart (:341-344)
      s_e += de * de;
    }
    if (g_mean) {
      float gmn = wm * 2.0f * (m_out - target) * inv_n;
      if (a.vb_through_mean) gmn += wv * dterm_dmean * c1 * dpred;
      g_mean[i] = gmn;
      if (g_v) g_v[i] = wv * dterm_dlv2 * dlv2_dv;
    }
  }
  s_mse = warp_sum(s_mse), s_vb = warp_sum(s_vb), s_x = warp_sum(s_x), s_e = warp_sum(s_e);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[0][warp] = s_mse, red[1][warp] = s_vb, red[2][warp] = s_x, red[3][warp] = s_e;
  __syncthreads();
  if (warp == 0) {
    float m = (lane < 8) ? red[0][lane] : 0.f;
    float v = (lane < 8) ? red[1][lane] : 0.f;
    float sx = (lane < 8) ? red[2][lane] : 0.f;
    float se = (lane < 8) ? red[3][lane] : 0.f;
    m = warp_sum(m), v = warp_sum(v), sx = warp_sum(sx), se = warp_sum(se);
    if (lane == 0) {
      const float mse = m * inv_n, vb = v * inv_n * inv_ln2 * a.vb_scale;
      a.mse[b] = mse;
      a.vb[b] = vb;
      a.loss[b] = mse + vb;
      if (a.xstart_mse) a.xstart_mse[b] = sx * inv_n;
      if (a.eps_mse) a.eps_mse[b] = se * inv_n;
    }
  }
}

}  // namespace ditb200

using namespace ditb200;

static unsigned grid_for(size_t work_items, int threads) {
  size_t blocks = (work_items + threads - 1) / threads;
  const size_t cap = (size_t)(num_sms() > 0 ? num_sms() : 148) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  return (unsigned)blocks;
}

extern "C" int ditb200_cfg_combine(const float* raw, float* out, int n_half, int C2, int HW,
                                   int n_cfg_ch, float cfg_scale, void* stream) {
  DITB_REQUIRE(raw && out && n_half > 0 && C2 > 0 && HW > 0 && n_cfg_ch >= 0, DITB200_EINVAL,
               "cfg_combine: bad argument");
  const size_t total = (size_t)n_half * C2 * HW;
  cfg_combine_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
      raw, out, n_half, C2, HW, n_cfg_ch, cfg_scale);
  DITB_LAUNCH_CHECK("cfg_combine");
  return 0;
}

extern "C" int ditb200_p_sample_step(const ditb200_step_args* a, void* stream) {
  DITB_REQUIRE(a != nullptr, DITB200_EINVAL, "p_sample_step: null args");
  DITB_REQUIRE(a->model_out && a->x && a->t && a->sample, DITB200_EINVAL,
               "p_sample_step: null tensor");
  DITB_REQUIRE(a->B > 0 && a->C > 0 && a->HW > 0 && a->num_timesteps > 0, DITB200_EINVAL,
               "p_sample_step: bad shape");
  DITB_REQUIRE(a->posterior_mean_coef1 && a->posterior_mean_coef2 && a->min_log, DITB200_EINVAL,
               "p_sample_step: null table");
  DITB_REQUIRE(a->mean_type == DITB200_MEAN_START_X ||
                   (a->sqrt_recip_alphas_cumprod && a->sqrt_recipm1_alphas_cumprod),
               DITB200_EINVAL, "p_sample_step: eps prediction needs the recip tables");
  DITB_REQUIRE(a->var_type != DITB200_VAR_LEARNED_RANGE || a->max_log, DITB200_EINVAL,
               "p_sample_step: LEARNED_RANGE needs max_log");
  DITB_REQUIRE(a->var_type >= 0 && a->var_type <= 2 && a->mean_type >= 0 && a->mean_type <= 1,
               DITB200_EINVAL, "p_sample_step: bad mean/var type");
  DITB_REQUIRE(a->sampler == DITB200_SAMPLER_ANCESTRAL ||
                   (a->sampler == DITB200_SAMPLER_DDIM && a->alphas_cumprod && a->alphas_cumprod_prev &&
                    a->sqrt_recip_alphas_cumprod && a->sqrt_recipm1_alphas_cumprod) ||
                   (a->sampler == DITB200_SAMPLER_DDIM_REVERSE && a->alphas_cumprod_next &&
                    a->sqrt_recip_alphas_cumprod && a->sqrt_recipm1_alphas_cumprod),
               DITB200_EINVAL, "p_sample_step: DDIM needs the alphas_cumprod (+prev / next) and recip tables");
  DITB_REQUIRE(!a->mean_override || a->sampler == DITB200_SAMPLER_ANCESTRAL, DITB200_EINVAL,
               "p_sample_step: mean_override goes with the ancestral update only");
  DITB_REQUIRE(a->cfg_half == 0 || a->B == 2 * a->cfg_half, DITB200_EINVAL,
               "p_sample_step: cfg_half=%d but B=%d", a->cfg_half, a->B);
  ditb200_step_args k = *a;
  // START_X never reads the recip tables; keep the loads in-bounds anyway
  if (!k.sqrt_recip_alphas_cumprod) k.sqrt_recip_alphas_cumprod = k.posterior_mean_coef1;
  if (!k.sqrt_recipm1_alphas_cumprod) k.sqrt_recipm1_alphas_cumprod = k.posterior_mean_coef1;
  const bool vec = (a->HW % 4 == 0) && aligned16(a->model_out) && aligned16(a->x) &&
                   aligned16(a->sample) && (!a->noise || aligned16(a->noise)) &&
                   (!a->pred_xstart || aligned16(a->pred_xstart)) &&
                   (!a->mean || aligned16(a->mean)) &&
                   (!a->log_variance || aligned16(a->log_variance)) &&
                   (!a->variance || aligned16(a->variance)) && (!a->mean_override || aligned16(a->mean_override));
  const size_t total = (size_t)a->B * a->C * a->HW;
  if (vec)
    p_sample_step_kernel<<<grid_for(total / 4, 256), 256, 0, (cudaStream_t)stream>>>(k);
  else
    p_sample_step_scalar_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(k);
  DITB_LAUNCH_CHECK("p_sample_step");
  return 0;
}

extern "C" int ditb200_q_sample(const float* x0, const float* noise, const int64_t* t,
                                const float* sac, const float* s1mac, float* x_t, int B, int CHW,
                                int num_timesteps, void* stream) {
  DITB_REQUIRE(x0 && noise && t && sac && s1mac && x_t && B > 0 && CHW > 0 && num_timesteps > 0,
               DITB200_EINVAL, "q_sample: bad argument");
  q_sample_kernel<<<grid_for((size_t)B * CHW, 256), 256, 0, (cudaStream_t)stream>>>(
      x0, noise, t, sac, s1mac, x_t, B, CHW, num_timesteps);
  DITB_LAUNCH_CHECK("q_sample");
  return 0;
}

extern "C" int ditb200_diffusion_affine(const float* a, const float* b, const float* b2, const int64_t* t,
                                        const float* ta, const float* tb, const float* td, int subtract, float* out,
                                        int B, int n, int num_timesteps, void* stream) {
  DITB_REQUIRE(t && out && B > 0 && n > 0 && num_timesteps > 0, DITB200_EINVAL, "diffusion_affine: bad argument");
  DITB_REQUIRE(a || ta, DITB200_EINVAL, "diffusion_affine: first term needs a tensor or a table");
  DITB_REQUIRE(b || (!tb && !b2), DITB200_EINVAL, "diffusion_affine: tb / b2 without b");
  diffusion_affine_kernel<<<grid_for((size_t)B * n, 256), 256, 0, (cudaStream_t)stream>>>(a, b, b2, t, ta, tb, td, subtract,
                                                                                       out, B, n, num_timesteps);
  DITB_LAUNCH_CHECK("diffusion_affine");
  return 0;
}

extern "C" int ditb200_prior_bpd(const float* x0, float coef_mean, float log_var, float* out, int B, int n, void* stream) {
  DITB_REQUIRE(x0 && out && B > 0 && n > 0, DITB200_EINVAL, "prior_bpd: bad argument");
  prior_bpd_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(x0, coef_mean, log_var, out, n);
  DITB_LAUNCH_CHECK("prior_bpd");
  return 0;
}

extern "C" int ditb200_training_losses(const ditb200_loss_args* a, void* stream) {
  DITB_REQUIRE(a != nullptr, DITB200_EINVAL, "training_losses: null args");
  DITB_REQUIRE(a->model_out && a->x0 && a->x_t && a->noise && a->t && a->mse && a->vb && a->loss,
               DITB200_EINVAL, "training_losses: null tensor");
  DITB_REQUIRE(a->sqrt_recip_alphas_cumprod && a->sqrt_recipm1_alphas_cumprod &&
                   a->posterior_mean_coef1 && a->posterior_mean_coef2 &&
                   a->posterior_log_variance_clipped && a->log_betas,
               DITB200_EINVAL, "training_losses: null table");
  DITB_REQUIRE(a->B > 0 && a->C > 0 && a->HW > 0 && a->num_timesteps > 0, DITB200_EINVAL,
               "training_losses: bad shape");
  DITB_REQUIRE(!a->grad_model_out || (a->w_mse && a->w_vb), DITB200_EINVAL,
               "training_losses: grad_model_out needs w_mse and w_vb");
  DITB_REQUIRE(a->mean_type == DITB200_MEAN_EPSILON || a->mean_type == DITB200_MEAN_START_X, DITB200_EINVAL,
               "training_losses: bad mean_type %d", a->mean_type);
  DITB_REQUIRE(a->var_type >= 0 && a->var_type <= 2 && (a->var_type != DITB200_VAR_FIXED || a->fixed_log_var),
               DITB200_EINVAL, "training_losses: bad var_type %d (FIXED needs fixed_log_var)", a->var_type);
  training_losses_kernel<<<a->B, 256, 0, (cudaStream_t)stream>>>(*a);
  DITB_LAUNCH_CHECK("training_losses");
  return 0;
}
