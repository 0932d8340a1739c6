// optim.cu — one pass over the flat parameter arena per training step: AdamW update, EMA update and the
// bf16 weight shadow the next forward's tensor-core GEMMs read, fused.
// The reference does this with ~400 per-parameter AdamW launches (train.py:161,206-207) plus a
// per-parameter mul_/add_ loop for the EMA (train.py:41-51), and autocast re-casts every weight to
// bf16 inside the next forward.  Here every byte moves once: read g, p, m, v, ema (20 B), write p, m,
// v, ema, shadow (18 B) = 38 B per parameter, HBM-bound.
#include "common.cuh"

namespace ditb200 {

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, inv_bc1, inv_sqrt_bc2, ema_decay;
};

__device__ __forceinline__ float adam_one(float g, float& p, float& m, float& v, const AdamArgs& a) {
  p *= 1.0f - a.lr * a.weight_decay;  // decoupled weight decay (torch.optim.AdamW)
  m = a.beta1 * m + (1.0f - a.beta1) * g;
  v = a.beta2 * v + (1.0f - a.beta2) * g * g;
  const float denom = sqrtf(v) * a.inv_sqrt_bc2 + a.eps;
  p -= a.lr * a.inv_bc1 * (m / denom);
  return p;
}

__global__ void __launch_bounds__(256) adamw_ema_kernel(float4* __restrict__ p, const float4* __restrict__ g,
                                                        float4* __restrict__ m, float4* __restrict__ v,
                                                        float4* __restrict__ ema, uint2* __restrict__ shadow,
                                                        size_t n4, const AdamArgs a) {
  DITB_PDL_WAIT();
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 gv = ldg_stream_f4(g + i);
    float4 pv = p[i], mv = m[i], vv = v[i];
    adam_one(gv.x, pv.x, mv.x, vv.x, a);
    adam_one(gv.y, pv.y, mv.y, vv.y, a);
    adam_one(gv.z, pv.z, mv.z, vv.z, a);
    adam_one(gv.w, pv.w, mv.w, vv.w, a);
    p[i] = pv, m[i] = mv, v[i] = vv;
    if (ema != nullptr) {  // ema = decay * ema + (1 - decay) * p   (train.py:47-51)
      float4 e = ema[i];
      const float d = a.ema_decay, c = 1.0f - a.ema_decay;
      e.x = d * e.x + c * pv.x, e.y = d * e.y + c * pv.y, e.z = d * e.z + c * pv.z, e.w = d * e.w + c * pv.w;
      ema[i] = e;
    }
    if (shadow != nullptr) {
      uint2 pk;
      pk.x = pack_bf16x2(pv.x, pv.y);
      pk.y = pack_bf16x2(pv.z, pv.w);
      shadow[i] = pk;
    }
  }
}

}  // namespace ditb200

using namespace ditb200;

extern "C" int ditb200_adamw_ema(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema,
                                 void* shadow_bf16, size_t n, float lr, float beta1, float beta2, float eps,
                                 float weight_decay, int step, float ema_decay, void* stream) {
  DITB_REQUIRE(param && grad && exp_avg && exp_avg_sq, DITB200_EINVAL, "adamw_ema: null pointer");
  DITB_REQUIRE(n > 0 && n % 4 == 0 && step >= 1, DITB200_EINVAL, "adamw_ema: n must be a positive multiple of 4, step >= 1");
  DITB_REQUIRE(aligned16(param) && aligned16(grad) && aligned16(exp_avg) && aligned16(exp_avg_sq) &&
                   (!ema || aligned16(ema)) && (!shadow_bf16 || (reinterpret_cast<uintptr_t>(shadow_bf16) & 7u) == 0),
               DITB200_EALIGN, "adamw_ema: misaligned pointer");
  AdamArgs a;
  a.lr = lr, a.beta1 = beta1, a.beta2 = beta2, a.eps = eps, a.weight_decay = weight_decay, a.ema_decay = ema_decay;
  a.inv_bc1 = (float)(1.0 / (1.0 - pow((double)beta1, (double)step)));
  a.inv_sqrt_bc2 = (float)(1.0 / sqrt(1.0 - pow((double)beta2, (double)step)));
  const size_t n4 = n / 4;
  int blocks = num_sms() > 0 ? num_sms() * 8 : 1184;
  if ((size_t)blocks * 256 > n4) blocks = (int)((n4 + 255) / 256);
  DITB_KLAUNCH(adamw_ema_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream,
      reinterpret_cast<float4*>(param), reinterpret_cast<const float4*>(grad), reinterpret_cast<float4*>(exp_avg),
      reinterpret_cast<float4*>(exp_avg_sq), reinterpret_cast<float4*>(ema), reinterpret_cast<uint2*>(shadow_bf16), n4, a);
  DITB_LAUNCH_CHECK("adamw_ema");
  return 0;
}
