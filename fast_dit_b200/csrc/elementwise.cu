// elementwise.cu — the HBM-bound kernels of the denoiser path: embedders, LayerNorm+modulate,
// final layer, bf16 casts.  Each is bounded by the HBM roofline; the
// rules that matter are coalesced 128-bit accesses, one pass over the data, and enough
// CTAs to cover 148 SMs.
#include <stdlib.h>

#include "common.cuh"

namespace ditb200 {

// =============================================================== LayerNorm + modulate
// One warp per token row.  Lane l owns float4 number l + 32*j (j < NV) of the row, so every
// warp-level load/store is a contiguous 512-byte segment.  The row stays in registers between
// the statistics pass and the normalise pass: x is read from HBM exactly once.
template <int NV, bool kOutBf16>
__device__ __forceinline__ void ln_modulate_row(const float* __restrict__ x, const float* __restrict__ shift,
                                                const float* __restrict__ scale, int mod_stride,
                                                void* __restrict__ out, float* __restrict__ stats, int row, int T,
                                                float eps, int lane) {
  constexpr int D = NV * 128;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  float4 v[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = ldg_stream_f4(xr + lane + 32 * j);
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float a = v[j].x - mean, b = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = warp_sum(q) * (1.0f / D);
  const float rstd = 1.0f / sqrtf(var + eps);
  if (stats != nullptr && lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
  const int b = row / T;
  const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)b * mod_stride);
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 h4 = __ldg(sh + lane + 32 * j);
    const float4 c4 = __ldg(sc + lane + 32 * j);
    float4 o;
    o.x = (v[j].x - mean) * rstd * (1.0f + c4.x) + h4.x;
    o.y = (v[j].y - mean) * rstd * (1.0f + c4.y) + h4.y;
    o.z = (v[j].z - mean) * rstd * (1.0f + c4.z) + h4.z;
    o.w = (v[j].w - mean) * rstd * (1.0f + c4.w) + h4.w;
    if constexpr (kOutBf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * D)[lane + 32 * j] = pk;
    } else {
      reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)row * D)[lane + 32 * j] = o;
    }
  }
}

template <int NV, bool kOutBf16>
__global__ void __launch_bounds__(256) ln_modulate_kernel(const float* __restrict__ x,
                                                          const float* __restrict__ shift,
                                                          const float* __restrict__ scale,
                                                          int mod_stride, void* __restrict__ out,
                                                          float* __restrict__ stats, int M, int T,
                                                          float eps, int reverse) {
  DITB_PDL_WAIT();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  ln_modulate_row<NV, kOutBf16>(x, shift, scale, mod_stride, out, stats, reverse ? M - 1 - row : row, T, eps, lane);
}

// Gated residual update + LayerNorm + modulate in one pass (training forward): x_out = x + gate[b] * y is formed
// in registers from the f32 stream and the bf16 branch output, written back once, and normalised / modulated
// from the same registers.
template <int NV, bool kOutBf16>
__device__ __forceinline__ void ln_modulate_resid_row(
    const float* __restrict__ x, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
    const float* __restrict__ shift, const float* __restrict__ scale, int mod_stride, float* __restrict__ x_out,
    void* __restrict__ out, float* __restrict__ stats, int row, int T, float eps, int lane) {
  constexpr int D = NV * 128;
  const int b = row / T;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  const uint2* yr = reinterpret_cast<const uint2*>(y + (size_t)row * D);
  const float4* gr = reinterpret_cast<const float4*>(gate + (size_t)b * mod_stride);
  float4 v[NV];
  uint2 yv[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) v[j] = ldg_stream_f4(xr + lane + 32 * j), yv[j] = yr[lane + 32 * j];
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 g4 = __ldg(gr + lane + 32 * j);
    const float2 y0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv[j].x));
    const float2 y1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&yv[j].y));
    v[j].x = fmaf(g4.x, y0.x, v[j].x), v[j].y = fmaf(g4.y, y0.y, v[j].y);
    v[j].z = fmaf(g4.z, y1.x, v[j].z), v[j].w = fmaf(g4.w, y1.y, v[j].w);
    reinterpret_cast<float4*>(x_out + (size_t)row * D)[lane + 32 * j] = v[j];
    s += (v[j].x + v[j].y) + (v[j].z + v[j].w);
  }
  if (out == nullptr) return;
  const float mean = warp_sum(s) * (1.0f / D);
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    float a = v[j].x - mean, bb = v[j].y - mean, c = v[j].z - mean, d = v[j].w - mean;
    q += (a * a + bb * bb) + (c * c + d * d);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
  if (stats != nullptr && lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
  const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)b * mod_stride);
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride);
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const float4 h4 = __ldg(sh + lane + 32 * j);
    const float4 c4 = __ldg(sc + lane + 32 * j);
    float4 o;
    o.x = (v[j].x - mean) * rstd * (1.0f + c4.x) + h4.x;
    o.y = (v[j].y - mean) * rstd * (1.0f + c4.y) + h4.y;
    o.z = (v[j].z - mean) * rstd * (1.0f + c4.z) + h4.z;
    o.w = (v[j].w - mean) * rstd * (1.0f + c4.w) + h4.w;
    if constexpr (kOutBf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * D)[lane + 32 * j] = pk;
    } else {
      reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)row * D)[lane + 32 * j] = o;
    }
  }
}

template <int NV, bool kOutBf16>
__global__ void __launch_bounds__(128) ln_modulate_resid_kernel(
    const float* __restrict__ x, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
    const float* __restrict__ shift, const float* __restrict__ scale, int mod_stride, float* __restrict__ x_out,
    void* __restrict__ out, float* __restrict__ stats, int M, int T, float eps, int reverse) {
  DITB_PDL_WAIT();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  ln_modulate_resid_row<NV, kOutBf16>(x, y, gate, shift, scale, mod_stride, x_out, out, stats,
                                      reverse ? M - 1 - row : row, T, eps, lane);
}

// Any D % 4 == 0: same mapping, row re-read from L1/L2 instead of held in registers.
template <bool kOutBf16>
__global__ void __launch_bounds__(256) ln_modulate_generic_kernel(
    const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
    int mod_stride, void* __restrict__ out, float* __restrict__ stats, int M, int T, int D,
    float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  if (row >= M) return;
  const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
  const int nv = D >> 2;
  float s = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float4 v = xr[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
  const float mean = warp_sum(s) / (float)D;
  float q = 0.f;
  for (int i = lane; i < nv; i += 32) {
    float4 v = xr[i];
    float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
    q += (a * a + b * b) + (c * c + d * d);
  }
  const float var = warp_sum(q) / (float)D;
  const float rstd = 1.0f / sqrtf(var + eps);
  if (stats != nullptr && lane == 0) {
    stats[2 * row] = mean;
    stats[2 * row + 1] = rstd;
  }
  const int b = row / T;
  const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)b * mod_stride);
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride);
  for (int i = lane; i < nv; i += 32) {
    float4 v = xr[i];
    const float4 h4 = __ldg(sh + i), c4 = __ldg(sc + i);
    float4 o;
    o.x = (v.x - mean) * rstd * (1.0f + c4.x) + h4.x;
    o.y = (v.y - mean) * rstd * (1.0f + c4.y) + h4.y;
    o.z = (v.z - mean) * rstd * (1.0f + c4.z) + h4.z;
    o.w = (v.w - mean) * rstd * (1.0f + c4.w) + h4.w;
    if constexpr (kOutBf16) {
      uint2 pk;
      pk.x = pack_bf16x2(o.x, o.y);
      pk.y = pack_bf16x2(o.z, o.w);
      reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(out) + (size_t)row * D)[i] = pk;
    } else {
      reinterpret_cast<float4*>(reinterpret_cast<float*>(out) + (size_t)row * D)[i] = o;
    }
  }
}

// ================================================================ patch embed + pos
// grid.x = token groups of kTok tokens; threads run over the hidden dimension so every
// store to out[token, :] is coalesced.  Patches of the group are staged in shared memory.
constexpr int kPeTok = 16;
__global__ void __launch_bounds__(256) patch_embed_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ pos, float* __restrict__ out, int B, int C, int H, int W, int p,
    int D, int round_bf16) {
  extern __shared__ float patch[];  // [kPeTok][K]
  const int Hp = H / p, Wp = W / p, T = Hp * Wp, K = C * p * p;
  const int M = B * T;
  const int tok0 = blockIdx.x * kPeTok;
  for (int idx = threadIdx.x; idx < kPeTok * K; idx += blockDim.x) {
    const int tt = idx / K, k = idx - tt * K;
    const int tok = tok0 + tt;
    float v = 0.f;
    if (tok < M) {
      const int b = tok / T, t = tok - b * T;
      const int hp = t / Wp, wp = t - hp * Wp;
      const int c = k / (p * p), r = k - c * p * p;
      const int i = r / p, j = r - i * p;
      v = x[(((size_t)b * C + c) * H + hp * p + i) * W + wp * p + j];
      if (round_bf16) v = bf16_round(v);
    }
    patch[idx] = v;
  }
  __syncthreads();
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float acc[kPeTok];
#pragma unroll
    for (int tt = 0; tt < kPeTok; ++tt) acc[tt] = 0.f;
    const float* wr = w + (size_t)d * K;
    if ((K & 3) == 0) {
      // four k per step: one 128-bit weight load and one 128-bit (broadcast) patch load per token
      for (int k = 0; k < K; k += 4) {
        float4 w4 = __ldg(reinterpret_cast<const float4*>(wr + k));
        if (round_bf16) w4 = make_float4(bf16_round(w4.x), bf16_round(w4.y), bf16_round(w4.z), bf16_round(w4.w));
#pragma unroll
        for (int tt = 0; tt < kPeTok; ++tt) {
          const float4 p4 = *reinterpret_cast<const float4*>(&patch[tt * K + k]);
          acc[tt] = fmaf(w4.x, p4.x, acc[tt]);
          acc[tt] = fmaf(w4.y, p4.y, acc[tt]);
          acc[tt] = fmaf(w4.z, p4.z, acc[tt]);
          acc[tt] = fmaf(w4.w, p4.w, acc[tt]);
        }
      }
    } else {
      for (int k = 0; k < K; ++k) {
        float wv = __ldg(wr + k);
        if (round_bf16) wv = bf16_round(wv);
#pragma unroll
        for (int tt = 0; tt < kPeTok; ++tt) acc[tt] = fmaf(wv, patch[tt * K + k], acc[tt]);
      }
    }
    float bv = bias[d];
    if (round_bf16) bv = bf16_round(bv);
#pragma unroll
    for (int tt = 0; tt < kPeTok; ++tt) {
      const int tok = tok0 + tt;
      if (tok < M) {
        float r = acc[tt] + bv;
        if (round_bf16) r = bf16_round(r);
        const int t = tok % T;
        out[(size_t)tok * D + d] = r + __ldg(pos + (size_t)t * D + d);
      }
    }
  }
}

// K = C*p*p = 16 (every /2 model): a thread keeps the weight rows of 4 consecutive hidden channels in registers
// (64 values) and runs over the group's tokens: 4 broadcast 128-bit shared loads feed 64 FMAs (the kernel above
// spends one such load per 4 FMAs and is bound by the shared-memory pipe), and bias / pos / out move as float4.
// Persistent CTAs: the weights are fetched once, and the 16 pos rows of a token group are requested before the
// patch gather so that a group pays one memory latency, not one per token (ncu: 72 us -> latency-bound on pos).
// Same k order and the same fmaf chain per output as the general kernel: bit-identical results.
__global__ void __launch_bounds__(320, 1) patch_embed_k16_kernel(
    const float* __restrict__ x, const float* __restrict__ w, const float* __restrict__ bias,
    const float* __restrict__ pos, float* __restrict__ out, int B, int C, int H, int W, int p,
    int D, int round_bf16) {
  __shared__ __align__(16) float patch[kPeTok * 16];
  const int Hp = H / p, Wp = W / p, T = Hp * Wp;
  const int M = B * T;
  const int d0 = threadIdx.x * 4;
  const bool active = d0 < D;
  float wq[4][16];
  float bq[4] = {0.f, 0.f, 0.f, 0.f};
  if (active) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
#pragma unroll
      for (int k = 0; k < 16; k += 4) {
        float4 w4 = __ldg(reinterpret_cast<const float4*>(w + (size_t)(d0 + q) * 16 + k));
        if (round_bf16) w4 = make_float4(bf16_round(w4.x), bf16_round(w4.y), bf16_round(w4.z), bf16_round(w4.w));
        wq[q][k] = w4.x, wq[q][k + 1] = w4.y, wq[q][k + 2] = w4.z, wq[q][k + 3] = w4.w;
      }
    }
    float4 b4 = __ldg(reinterpret_cast<const float4*>(bias + d0));
    if (round_bf16) b4 = make_float4(bf16_round(b4.x), bf16_round(b4.y), bf16_round(b4.z), bf16_round(b4.w));
    bq[0] = b4.x, bq[1] = b4.y, bq[2] = b4.z, bq[3] = b4.w;
  }
  const int groups = (M + kPeTok - 1) / kPeTok;
  for (int g = blockIdx.x; g < groups; g += gridDim.x) {
    const int tok0 = g * kPeTok;
    float4 ps[kPeTok];
    if (active) {
#pragma unroll
      for (int tt = 0; tt < kPeTok; ++tt) {
        const int tok = min(tok0 + tt, M - 1);
        ps[tt] = __ldg(reinterpret_cast<const float4*>(pos + (size_t)(tok % T) * D + d0));
      }
    }
    for (int idx = threadIdx.x; idx < kPeTok * 16; idx += blockDim.x) {
      const int tt = idx >> 4, k = idx & 15;
      const int tok = tok0 + tt;
      float v = 0.f;
      if (tok < M) {
        const int b = tok / T, t = tok - b * T;
        const int hp = t / Wp, wp = t - hp * Wp;
        const int c = k / (p * p), r = k - c * p * p;
        const int i = r / p, j = r - i * p;
        v = x[(((size_t)b * C + c) * H + hp * p + i) * W + wp * p + j];
        if (round_bf16) v = bf16_round(v);
      }
      patch[idx] = v;
    }
    __syncthreads();
    if (active) {
#pragma unroll
      for (int tt = 0; tt < kPeTok; ++tt) {
        const int tok = tok0 + tt;
        if (tok < M) {
          float pv[16];
#pragma unroll
          for (int k = 0; k < 16; k += 4) {
            const float4 p4 = *reinterpret_cast<const float4*>(&patch[tt * 16 + k]);
            pv[k] = p4.x, pv[k + 1] = p4.y, pv[k + 2] = p4.z, pv[k + 3] = p4.w;
          }
          float r[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float acc = 0.f;
#pragma unroll
            for (int k = 0; k < 16; ++k) acc = fmaf(wq[q][k], pv[k], acc);
            acc += bq[q];
            r[q] = round_bf16 ? bf16_round(acc) : acc;
          }
          *reinterpret_cast<float4*>(out + (size_t)tok * D + d0) =
              make_float4(r[0] + ps[tt].x, r[1] + ps[tt].y, r[2] + ps[tt].z, r[3] + ps[tt].w);
        }
      }
    }
    __syncthreads();  // the next group's gather overwrites `patch`
  }
}

// ============================================================= timestep sinusoid
__global__ void timestep_embedding_kernel(const void* __restrict__ t, int t_is_float, float* __restrict__ out,
                                          int B, int dim, float neg_log_period) {
  const int half = dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * dim) return;
  const int b = idx / dim, j = idx - b * dim;
  float r = 0.f;
  if (j < 2 * half) {
    const int k = (j < half) ? j : j - half;
    // freqs = exp(-log(P) * arange(half) / half), every op rounded to f32 as torch does
    const float f = expf(__fdiv_rn(__fmul_rn(neg_log_period, (float)k), (float)half));
    const float tv = t_is_float ? reinterpret_cast<const float*>(t)[b] : (float)reinterpret_cast<const int64_t*>(t)[b];
    const float a = __fmul_rn(tv, f);
    r = (j < half) ? cosf(a) : sinf(a);
  }
  out[idx] = r;
}

// ================================================================== label embed
__global__ void label_embed_kernel(const int64_t* __restrict__ y, const float* __restrict__ table,
                                   const float* __restrict__ add, float* __restrict__ out, int B,
                                   int D, int num_rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, d = idx - b * D;
  long long row = y[b];
  if (row < 0) row = 0;
  if (row >= num_rows) row = num_rows - 1;
  float v = table[(size_t)row * D + d];
  if (add != nullptr) v = add[idx] + v;  // c = t + y
  out[idx] = v;
}

// ===================================================================== casts
__global__ void cast_bf16_kernel(const float* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                 size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x * 4;
  for (size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i < n; i += stride) {
    if (i + 4 <= n) {
      float4 v = *reinterpret_cast<const float4*>(in + i);
      uint2 pk;
      pk.x = pack_bf16x2(v.x, v.y);
      pk.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(out + i) = pk;
    } else {
      for (size_t k = i; k < n; ++k) out[k] = __float2bfloat16_rn(in[k]);
    }
  }
}

// SiLU then cast: the A operand of the batched adaLN GEMM (nn.SiLU of MO:114 on c[N, D])
template <bool kOutBf16>
__global__ void silu_cast_kernel(const float* __restrict__ in, void* __restrict__ out, size_t n) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float v = silu_acc(in[i]);
    if constexpr (kOutBf16)
      reinterpret_cast<__nv_bfloat16*>(out)[i] = __float2bfloat16_rn(v);
    else
      reinterpret_cast<float*>(out)[i] = v;
  }
}

// ================================================================= final layer
// CTA = kFlRows token rows.  Phase 1: one warp normalises + modulates two rows into shared
// memory.  Phase 2: a [kFlRows x NO] mini-GEMM against the (transposed, k-chunked) weight,
// thread = (row, output) so h reads are warp broadcasts and weight reads are conflict-free.
// The store does the unpatchify permutation ('nhwpqc->nchpwq') directly into NCHW.
constexpr int kFlRows = 16;
constexpr int kFlKc = 64;
template <int NOW>  // padded outputs / 32
__global__ void __launch_bounds__(256) final_layer_kernel(
    const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
    int mod_stride, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ out, int M, int T, int D, int p, int Cout, float eps, int round_bf16) {
  extern __shared__ float smem[];
  float* h = smem;                         // [kFlRows][D]
  float* wT = smem + (size_t)kFlRows * D;  // [NOW*32][kFlKc + 4]
  constexpr int NOP = NOW * 32;
  const int NO = p * p * Cout;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row0 = blockIdx.x * kFlRows;
  const int nv = D >> 2;
  // ---- phase 1
  for (int rr = warp; rr < kFlRows; rr += 8) {
    const int row = row0 + rr;
    float4* hr = reinterpret_cast<float4*>(h + (size_t)rr * D);
    if (row >= M) {
      for (int i = lane; i < nv; i += 32) hr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      continue;
    }
    const float4* xr = reinterpret_cast<const float4*>(x + (size_t)row * D);
    float s = 0.f;
    for (int i = lane; i < nv; i += 32) {
      float4 v = xr[i];
      s += (v.x + v.y) + (v.z + v.w);
    }
    const float mean = warp_sum(s) / (float)D;
    float q = 0.f;
    for (int i = lane; i < nv; i += 32) {
      float4 v = xr[i];
      float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
      q += (a * a + b * b) + (c * c + d * d);
    }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)D + eps);
    const int b = row / T;
    const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)b * mod_stride);
    const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride);
    for (int i = lane; i < nv; i += 32) {
      float4 v = xr[i];
      const float4 h4 = __ldg(sh + i), c4 = __ldg(sc + i);
      float4 o;
      o.x = (v.x - mean) * rstd * (1.0f + c4.x) + h4.x;
      o.y = (v.y - mean) * rstd * (1.0f + c4.y) + h4.y;
      o.z = (v.z - mean) * rstd * (1.0f + c4.z) + h4.z;
      o.w = (v.w - mean) * rstd * (1.0f + c4.w) + h4.w;
      if (round_bf16) {
        o.x = bf16_round(o.x);
        o.y = bf16_round(o.y);
        o.z = bf16_round(o.z);
        o.w = bf16_round(o.w);
      }
      hr[i] = o;
    }
  }
  // ---- phase 2
  // thread -> OT output columns x RT rows.  All accumulators of a thread reuse its 128-bit weight
  // reads (OT per 4 k) and the broadcast 128-bit h reads (RT per 4 k): OT*RT*4 FMAs per OT+RT loads.
  constexpr int NOPT = NOP < 256 ? NOP : 256;  // output columns covered by one pass of the CTA
  constexpr int OT = NOP / NOPT;
  constexpr int rstep = 256 / NOPT;
  constexpr int RT = kFlRows / rstep;
  constexpr int WLD = kFlKc + 4;  // padded row stride of the weight slab: conflict-free float4 reads
  float acc[OT][RT];
#pragma unroll
  for (int j = 0; j < OT; ++j)
#pragma unroll
    for (int a = 0; a < RT; ++a) acc[j][a] = 0.f;
  const int o0 = threadIdx.x % NOPT;
  const int rr0 = threadIdx.x / NOPT;
  for (int k0 = 0; k0 < D; k0 += kFlKc) {
    __syncthreads();  // h complete (first pass) / previous slab consumed
    for (int idx = threadIdx.x; idx < NOP * kFlKc; idx += 256) {
      const int oo = idx / kFlKc, kk = idx - oo * kFlKc;
      float wv = 0.f;
      if (oo < NO && k0 + kk < D) wv = __ldg(w + (size_t)oo * D + k0 + kk);
      if (round_bf16) wv = bf16_round(wv);
      wT[oo * WLD + kk] = wv;
    }
    __syncthreads();
    const int kmax = min(kFlKc, D - k0);  // D % 4 == 0
    for (int kk = 0; kk < kmax; kk += 4) {
      float4 w4[OT];
#pragma unroll
      for (int j = 0; j < OT; ++j) w4[j] = *reinterpret_cast<const float4*>(&wT[(o0 + j * NOPT) * WLD + kk]);
#pragma unroll
      for (int a = 0; a < RT; ++a) {
        const float4 h4 = *reinterpret_cast<const float4*>(h + (size_t)(rr0 + a * rstep) * D + k0 + kk);
#pragma unroll
        for (int j = 0; j < OT; ++j) {
          acc[j][a] = fmaf(h4.x, w4[j].x, acc[j][a]);
          acc[j][a] = fmaf(h4.y, w4[j].y, acc[j][a]);
          acc[j][a] = fmaf(h4.z, w4[j].z, acc[j][a]);
          acc[j][a] = fmaf(h4.w, w4[j].w, acc[j][a]);
        }
      }
    }
  }
  const int Wp = (int)(sqrtf((float)T) + 0.5f);
  const int Himg = Wp * p;
#pragma unroll
  for (int j = 0; j < OT; ++j) {
    const int o = o0 + j * NOPT;
#pragma unroll
    for (int a = 0; a < RT; ++a) {
      const int row = row0 + rr0 + a * rstep;
      if (row < M && o < NO) {
        float bv = bias[o];
        if (round_bf16) bv = bf16_round(bv);
        float r = acc[j][a] + bv;
        if (round_bf16) r = bf16_round(r);
        const int b = row / T, t = row - b * T;
        const int hp = t / Wp, wp = t - hp * Wp;
        const int c = o % Cout, pq = o / Cout;
        const int pi = pq / p, qj = pq - pi * p;
        out[(((size_t)b * Cout + c) * Himg + hp * p + pi) * Himg + wp * p + qj] = r;
      }
    }
  }
}

}  // namespace ditb200

using namespace ditb200;

// ------------------------------------------------------------------ C entry points
extern "C" int ditb200_ln_modulate(const float* x, const float* shift, const float* scale,
                                   int mod_stride, void* out, int out_dtype, float* stats, int B,
                                   int T, int D, float eps, int reverse, void* stream) {
  DITB_REQUIRE(x && shift && scale && out, DITB200_EINVAL, "ln_modulate: null pointer");
  DITB_REQUIRE(B > 0 && T > 0 && D > 0 && D % 4 == 0, DITB200_EINVAL,
               "ln_modulate: need B,T > 0 and D %% 4 == 0 (D=%d)", D);
  DITB_REQUIRE(mod_stride % 4 == 0, DITB200_EALIGN, "ln_modulate: mod_stride %% 4 != 0");
  DITB_REQUIRE(aligned16(x) && aligned16(shift) && aligned16(scale) && aligned16(out),
               DITB200_EALIGN, "ln_modulate: pointers must be 16-byte aligned");
  DITB_REQUIRE(out_dtype == DITB200_F32 || out_dtype == DITB200_BF16, DITB200_EINVAL,
               "ln_modulate: bad out_dtype %d", out_dtype);
  cudaStream_t st = (cudaStream_t)stream;
  const int M = B * T;
  static const int ln_threads = getenv("DITB200_LN_THREADS") ? atoi(getenv("DITB200_LN_THREADS")) : 128;  // tuning switch (64..512)
  DITB_REQUIRE(ln_threads >= 32 && ln_threads <= 256 && ln_threads % 32 == 0, DITB200_EINVAL,
               "ln_modulate: DITB200_LN_THREADS must be a multiple of 32 in [32, 256]");
  const int wpb = ln_threads / 32;
  const dim3 grid((M + wpb - 1) / wpb), block(ln_threads);
  const bool bf = out_dtype == DITB200_BF16;
#define LN_CASE(NV)                                                                              \
  case NV * 128:                                                                                 \
    if (bf)                                                                                      \
      DITB_KLAUNCH((ln_modulate_kernel<NV, true>), grid, block, 0, st, x, shift, scale, mod_stride, out, stats, M, T, eps, reverse); \
    else                                                                                         \
      DITB_KLAUNCH((ln_modulate_kernel<NV, false>), grid, block, 0, st, x, shift, scale, mod_stride, out, stats, M, T, eps, reverse); \
    break;
  switch (D) {
    LN_CASE(3)
    LN_CASE(6)
    LN_CASE(8)
    LN_CASE(9)
    default:
      if (bf)
        ln_modulate_generic_kernel<true><<<grid, block, 0, st>>>(x, shift, scale, mod_stride, out,
                                                                 stats, M, T, D, eps);
      else
        ln_modulate_generic_kernel<false><<<grid, block, 0, st>>>(x, shift, scale, mod_stride,
                                                                  out, stats, M, T, D, eps);
  }
#undef LN_CASE
  DITB_LAUNCH_CHECK("ln_modulate");
  return 0;
}

extern "C" int ditb200_ln_modulate_resid(const float* x, const void* y, const float* gate, const float* shift,
                                         const float* scale, int mod_stride, float* x_out, void* out, int out_dtype,
                                         float* stats, int B, int T, int D, float eps, int reverse, void* stream) {
  DITB_REQUIRE(x && y && gate && x_out, DITB200_EINVAL, "ln_modulate_resid: null pointer");
  DITB_REQUIRE(out == nullptr || (shift && scale), DITB200_EINVAL, "ln_modulate_resid: out needs shift and scale");
  DITB_REQUIRE(B > 0 && T > 0 && (D == 384 || D == 768 || D == 1024 || D == 1152), DITB200_EINVAL,
               "ln_modulate_resid: bad shape B=%d T=%d D=%d (D in 384, 768, 1024, 1152)", B, T, D);
  DITB_REQUIRE(mod_stride % 4 == 0 && aligned16(x) && aligned16(y) && aligned16(gate) && aligned16(x_out) &&
                   (!out || (aligned16(out) && aligned16(shift) && aligned16(scale))),
               DITB200_EALIGN, "ln_modulate_resid: misaligned pointer or stride");
  DITB_REQUIRE(out_dtype == DITB200_F32 || out_dtype == DITB200_BF16, DITB200_EINVAL, "ln_modulate_resid: bad out_dtype");
  cudaStream_t st = (cudaStream_t)stream;
  const int M = B * T;
  const dim3 grid((M + 3) / 4), block(128);
  const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(y);
  const bool bf = out_dtype == DITB200_BF16;
#define LNR_CASE(NV)                                                                                                  \
  case NV * 128:                                                                                                      \
    if (bf)                                                                                                           \
      DITB_KLAUNCH((ln_modulate_resid_kernel<NV, true>), grid, block, 0, st, x, yb, gate, shift, scale, mod_stride, x_out, out, stats, M, T, eps, reverse); \
    else                                                                                                              \
      DITB_KLAUNCH((ln_modulate_resid_kernel<NV, false>), grid, block, 0, st, x, yb, gate, shift, scale, mod_stride, x_out, out, stats, M, T, eps, reverse); \
    break;
  switch (D) {
    LNR_CASE(3)
    LNR_CASE(6)
    LNR_CASE(8)
    LNR_CASE(9)
  }
#undef LNR_CASE
  DITB_LAUNCH_CHECK("ln_modulate_resid");
  return 0;
}

extern "C" int ditb200_patch_embed(const float* x, const float* w, const float* bias,
                                   const float* pos, float* out, int B, int C, int H, int W, int p,
                                   int D, int round_bf16, void* stream) {
  DITB_REQUIRE(x && w && bias && pos && out, DITB200_EINVAL, "patch_embed: null pointer");
  DITB_REQUIRE(B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0 && D > 0, DITB200_EINVAL,
               "patch_embed: bad shape B=%d C=%d H=%d W=%d p=%d D=%d", B, C, H, W, p, D);
  const int K = C * p * p;
  DITB_REQUIRE(K <= 1024, DITB200_EINVAL, "patch_embed: C*p*p = %d > 1024", K);
  const int M = B * (H / p) * (W / p);
  const size_t smem = (size_t)kPeTok * K * sizeof(float);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(patch_embed_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return check_cuda(e, "patch_embed smem attr");
  }
  if (K == 16 && D % 4 == 0 && D / 4 <= 320 && aligned16(w) && aligned16(bias) && aligned16(pos) && aligned16(out)) {
    const int threads = (D / 4 + 31) / 32 * 32;
    int g = num_sms() > 0 ? num_sms() : 148;
    if (g > (M + kPeTok - 1) / kPeTok) g = (M + kPeTok - 1) / kPeTok;
    patch_embed_k16_kernel<<<g, threads, 0, (cudaStream_t)stream>>>(
        x, w, bias, pos, out, B, C, H, W, p, D, round_bf16);
    DITB_LAUNCH_CHECK("patch_embed");
    return 0;
  }
  patch_embed_kernel<<<(M + kPeTok - 1) / kPeTok, 256, smem, (cudaStream_t)stream>>>(
      x, w, bias, pos, out, B, C, H, W, p, D, round_bf16);
  DITB_LAUNCH_CHECK("patch_embed");
  return 0;
}

extern "C" int ditb200_timestep_embedding(const void* t, int t_is_float, float* out, int B, int dim,
                                          float max_period, void* stream) {
  DITB_REQUIRE(t && out && B > 0 && dim > 0 && max_period > 0.f, DITB200_EINVAL,
               "timestep_embedding: bad argument");
  const int n = B * dim;
  // -math.log(max_period) is a Python double that torch rounds to f32 when it meets the f32 arange
  const float neg_log = (float)(-log((double)max_period));
  timestep_embedding_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, t_is_float, out, B, dim, neg_log);
  DITB_LAUNCH_CHECK("timestep_embedding");
  return 0;
}

extern "C" int ditb200_label_embed(const int64_t* y, const float* table, const float* add,
                                   float* out, int B, int D, int num_rows, void* stream) {
  DITB_REQUIRE(y && table && out && B > 0 && D > 0 && num_rows > 0, DITB200_EINVAL,
               "label_embed: bad argument");
  const int n = B * D;
  label_embed_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(y, table, add, out, B, D,
                                                                         num_rows);
  DITB_LAUNCH_CHECK("label_embed");
  return 0;
}

extern "C" int ditb200_cast_bf16(const float* in, void* out, size_t n, void* stream) {
  DITB_REQUIRE(in && out, DITB200_EINVAL, "cast_bf16: null pointer");
  if (n == 0) return 0;
  DITB_REQUIRE(aligned16(in) && (reinterpret_cast<uintptr_t>(out) & 7u) == 0, DITB200_EALIGN,
               "cast_bf16: misaligned pointer");
  size_t blocks = (n / 4 + 255) / 256;
  const size_t cap = (size_t)(num_sms() > 0 ? num_sms() : 148) * 16;
  if (blocks > cap) blocks = cap;
  if (blocks == 0) blocks = 1;
  cast_bf16_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      in, reinterpret_cast<__nv_bfloat16*>(out), n);
  DITB_LAUNCH_CHECK("cast_bf16");
  return 0;
}

extern "C" int ditb200_silu_cast(const float* in, void* out, int out_dtype, size_t n, void* stream) {
  DITB_REQUIRE(in && out, DITB200_EINVAL, "silu_cast: null pointer");
  DITB_REQUIRE(out_dtype == DITB200_F32 || out_dtype == DITB200_BF16, DITB200_EINVAL,
               "silu_cast: bad out_dtype");
  if (n == 0) return 0;
  size_t blocks = (n + 255) / 256;
  const size_t cap = (size_t)(num_sms() > 0 ? num_sms() : 148) * 8;
  if (blocks > cap) blocks = cap;
  if (out_dtype == DITB200_BF16)
    silu_cast_kernel<true><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n);
  else
    silu_cast_kernel<false><<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(in, out, n);
  DITB_LAUNCH_CHECK("silu_cast");
  return 0;
}

// ---- 32-output variant (p = 2, learn_sigma: p*p*2C = 32; every */2 model): persistent, warp per token row.
// The [32, D] weight lives in shared memory for the lifetime of the CTA; a warp holds its row in registers
// (read once, the next row already in flight), normalises and modulates it there, forms 32 partial dot
// products per lane and reduces the 32 x 32 partials with a 31-shuffle transpose-reduction, after which lane o
// owns output o and stores it straight into its NCHW position (unpatchify).
template <int NV>
__global__ void __launch_bounds__(256) final_layer_rows_kernel(
    const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
    int mod_stride, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ out, int M, int T, int p, int Cout, float eps, int round_bf16) {
  constexpr int D = NV * 128;
  extern __shared__ float ws[];  // [32][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32 * D / 4; i += blockDim.x) {
    float4 v = __ldg(reinterpret_cast<const float4*>(w) + i);
    if (round_bf16) v.x = bf16_round(v.x), v.y = bf16_round(v.y), v.z = bf16_round(v.z), v.w = bf16_round(v.w);
    reinterpret_cast<float4*>(ws)[i] = v;
  }
  __syncthreads();
  const int W = gridDim.x * 8;  // warps in the grid
  const int Wp = (int)(sqrtf((float)T) + 0.5f), Himg = Wp * p;
  const float b_o = bias[lane];
  // Two rows per warp and pass: every 128-bit weight read from shared memory feeds 8 FMAs instead of 4 (the weight
  // re-read per row, 32 x D x 4 bytes, is what bounds this kernel: shared-memory pipe, not HBM).
  for (int row0 = 2 * (blockIdx.x * 8 + warp); row0 < M; row0 += 2 * W) {
    const bool two = row0 + 1 < M;
    if (row0 + 2 * W < M) {  // pull the next pair of rows towards L2 (no registers left for a software prefetch)
      const char* nxt = reinterpret_cast<const char*>(x + (size_t)(row0 + 2 * W) * D);
      const int bytes = (row0 + 2 * W + 1 < M ? 2 : 1) * D * 4;
      for (int o = lane * 128; o < bytes; o += 32 * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(nxt + o));
    }
    float4 v[2][NV];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float4* xr = reinterpret_cast<const float4*>(x + (size_t)(row0 + (two ? r : 0)) * D);
#pragma unroll
      for (int j = 0; j < NV; ++j) v[r][j] = ldg_stream_f4(xr + lane + 32 * j);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float s = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) s += (v[r][j].x + v[r][j].y) + (v[r][j].z + v[r][j].w);
      const float mean = warp_sum(s) * (1.0f / D);
      float q = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        float a = v[r][j].x - mean, b = v[r][j].y - mean, c = v[r][j].z - mean, d = v[r][j].w - mean;
        q += (a * a + b * b) + (c * c + d * d);
      }
      const float rstd = 1.0f / sqrtf(warp_sum(q) * (1.0f / D) + eps);
      const int b = (row0 + (two ? r : 0)) / T;
      const float4* sh = reinterpret_cast<const float4*>(shift + (size_t)b * mod_stride);
      const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride);
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 h4 = __ldg(sh + lane + 32 * j), c4 = __ldg(sc + lane + 32 * j);
        float4& u = v[r][j];
        u.x = (u.x - mean) * rstd * (1.0f + c4.x) + h4.x;
        u.y = (u.y - mean) * rstd * (1.0f + c4.y) + h4.y;
        u.z = (u.z - mean) * rstd * (1.0f + c4.z) + h4.z;
        u.w = (u.w - mean) * rstd * (1.0f + c4.w) + h4.w;
        if (round_bf16) u.x = bf16_round(u.x), u.y = bf16_round(u.y), u.z = bf16_round(u.z), u.w = bf16_round(u.w);
      }
    }
    float acc0[32], acc1[32];
#pragma unroll
    for (int o = 0; o < 32; ++o) {
      const float4* wr = reinterpret_cast<const float4*>(ws + (size_t)o * D);
      float a0 = 0.f, a1 = 0.f;
#pragma unroll
      for (int j = 0; j < NV; ++j) {
        const float4 w4 = wr[lane + 32 * j];
        a0 = fmaf(v[0][j].x, w4.x, a0), a0 = fmaf(v[0][j].y, w4.y, a0), a0 = fmaf(v[0][j].z, w4.z, a0), a0 = fmaf(v[0][j].w, w4.w, a0);
        a1 = fmaf(v[1][j].x, w4.x, a1), a1 = fmaf(v[1][j].y, w4.y, a1), a1 = fmaf(v[1][j].z, w4.z, a1), a1 = fmaf(v[1][j].w, w4.w, a1);
      }
      acc0[o] = a0, acc1[o] = a1;
    }
    // transpose-reduce: after the step with distance s, entry i of a lane holds the sum over its 32/s-lane group of
    // output (i + the bits of `lane` already consumed); 16 + 8 + 4 + 2 + 1 shuffles leave output `lane` in acc[0]
#pragma unroll
    for (int s2 = 16; s2 >= 1; s2 >>= 1) {
      const bool up = (lane & s2) != 0;
#pragma unroll
      for (int i = 0; i < s2; ++i) {
        const float keep0 = up ? acc0[i + s2] : acc0[i], send0 = up ? acc0[i] : acc0[i + s2];
        const float keep1 = up ? acc1[i + s2] : acc1[i], send1 = up ? acc1[i] : acc1[i + s2];
        acc0[i] = keep0 + __shfl_xor_sync(0xffffffffu, send0, s2);
        acc1[i] = keep1 + __shfl_xor_sync(0xffffffffu, send1, s2);
      }
    }
    // output o = lane: (pi, pj, c) = unpatchify order (models_original.py:228-230)
    const int c = lane % Cout, pq = lane / Cout, pi = pq / p, pj = pq - pi * p;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      if (r == 1 && !two) break;
      const int row = row0 + r;
      const int b = row / T, t = row - b * T;
      const int hh = t / Wp, ww = t - hh * Wp;
      out[(((size_t)b * Cout + c) * Himg + hh * p + pi) * Himg + ww * p + pj] = (r ? acc1[0] : acc0[0]) + b_o;
    }
  }
}

// ---- 32-output variant on the tensor cores (round 2).  The SIMT kernel above is bound by re-reading the [32, D]
// weight from shared memory for every row pair (84 us for a 75 MB read at C3).  Here a CTA takes 16 token rows at
// a time and its 8 warps split the D columns: each lane holds its slice of two rows (g and g + 8 of the
// m16n8k8 fragment) in registers — read from HBM once, the next tile's slice already in flight — so mean and
// variance are exact two-pass sums (quad shuffles + one shared exchange each), the normalised, modulated values
// become A fragments directly, and the 16 x 32 x D/8 product per warp runs as mma.sync TF32 x 3 (a = hi + lo, w =
// hi + lo; lo*hi + hi*lo + hi*hi: fp32-accurate, this layer is kept at full precision).  The k index of a fragment
// slot is permuted (slot t / t + 4 of step 2j / 2j + 1 = column 16j + 4t + {0, 1} / {2, 3} of the warp's slice) so that
// both operands are read as 128-bit vectors: x from global memory, w from a [32][D + 16] shared array whose row
// stride (16 mod 32 floats) makes the quarter-warp reads conflict-free.  The 8 partial tiles meet in shared memory
// and leave through 128-byte segments of the NCHW image (unpatchify).
// f = hi + lo with hi = f rounded to TF32 (10 mantissa bits) and lo the exact remainder, which the tensor core reads
// truncated to TF32 again.  Three ALU instructions (cvt.rna.tf32.f32 compiles to ~10 with its NaN handling, and this
// kernel's instruction stream is what bounds it): add half an ulp to the bit pattern, clear the 13 low bits, subtract.
__device__ __forceinline__ void split_tf32(float f, uint32_t& hi, uint32_t& lo) {
  hi = (__float_as_uint(f) + 0x1000u) & 0xffffe000u;
  lo = __float_as_uint(f - __uint_as_float(hi));
}
__device__ __forceinline__ void mma_tf32_16x8x8(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                                uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
constexpr int kFtRows = 16, kFtOStride = 40;
template <int D>
__global__ void __launch_bounds__(256, 1) final_layer_tc_kernel(
    const float* __restrict__ x, const float* __restrict__ shift, const float* __restrict__ scale,
    int mod_stride, const float* __restrict__ w, const float* __restrict__ bias,
    float* __restrict__ out, int M, int T, int p, int Cout, float eps, int chunk) {
  DITB_PDL_WAIT();
  constexpr int KW = D / 8;    // columns per warp
  constexpr int NJ = KW / 16;  // 16-column steps per warp (two k steps of the MMA each)
  constexpr int WS = D + 16;   // padded weight row
  static_assert(KW % 16 == 0 && WS % 32 == 16, "slice / stride assumptions");
  extern __shared__ float ftsm[];
  float* ws = ftsm;                          // [32][WS]
  float* red_s = ws + 32 * WS;               // [2][8 warps][16 rows]
  float* red_o = red_s + 2 * 8 * kFtRows;    // [8 warps][16 rows][kFtOStride]
  float* smod = red_o + 8 * kFtRows * kFtOStride;  // [2 images][shift | scale][D]: the (at most two, T >= 16) images of a tile
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t = lane & 3;
  auto cp16 = [](float* dst, const float* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
  };
  for (int i = tid; i < 32 * (D / 4); i += 256) {  // the weight as asynchronous copies: all 36 per thread in flight at once
    const int n = i / (D / 4), k4 = i - n * (D / 4);
    cp16(ws + n * WS + 4 * k4, w + (size_t)n * D + 4 * k4);
  }
  const int ntiles = (M + kFtRows - 1) / kFtRows;
  const int tile0 = blockIdx.x * chunk, tile1 = min(ntiles, tile0 + chunk);  // a contiguous range: few image changes
  const int kbase = warp * KW + 4 * t;  // this lane's first column
  const int Wp = (int)(sqrtf((float)T) + 0.5f), Himg = Wp * p;
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 xa[NJ], xb[NJ];  // rows g and g + 8 of the current tile, this lane's columns
  auto load_tile = [&](int tile, float4 (&A)[NJ], float4 (&Bv)[NJ]) {
    const int ra = tile * kFtRows + g, rb = ra + 8;
    const float4* pa = reinterpret_cast<const float4*>(x + (size_t)(ra < M ? ra : 0) * D + kbase);
    const float4* pb = reinterpret_cast<const float4*>(x + (size_t)(rb < M ? rb : 0) * D + kbase);
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      A[j] = ra < M ? ldg_stream_f4(pa + 4 * j) : zero4;
      Bv[j] = rb < M ? ldg_stream_f4(pb + 4 * j) : zero4;
    }
  };
  if (tile0 < tile1) load_tile(tile0, xa, xb);
  int img0 = -1, img1 = -1;  // images whose shift / scale rows sit in smod[0] / smod[1]
  for (int tile = tile0; tile < tile1; ++tile) {
    float4 na[NJ], nb[NJ];
    if (tile + 1 < tile1) load_tile(tile + 1, na, nb);  // in flight under this tile's arithmetic
    // ---- the tile's images: shift and scale rows into shared memory when they change (block-uniform decision)
    const int b_lo = (tile * kFtRows) / T, b_hi = min(tile * kFtRows + kFtRows - 1, M - 1) / T;
    if (b_lo != img0 || b_hi != img1) {
      __syncthreads();  // the previous tile's readers of smod are done
      for (int i = tid; i < 4 * (D / 4); i += 256) {
        const int row = i / (D / 4), k4 = i - row * (D / 4);  // row = image slot * 2 + (0 shift | 1 scale)
        const int bimg = row < 2 ? b_lo : b_hi;
        cp16(smod + row * D + 4 * k4, ((row & 1) ? scale : shift) + (size_t)bimg * mod_stride + 4 * k4);
      }
      img0 = b_lo, img1 = b_hi;
    }
    // ---- mean
    float sa = 0.f, sb = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      sa += (xa[j].x + xa[j].y) + (xa[j].z + xa[j].w);
      sb += (xb[j].x + xb[j].y) + (xb[j].z + xb[j].w);
    }
    sa += __shfl_xor_sync(0xffffffffu, sa, 1), sb += __shfl_xor_sync(0xffffffffu, sb, 1);
    sa += __shfl_xor_sync(0xffffffffu, sa, 2), sb += __shfl_xor_sync(0xffffffffu, sb, 2);
    if (t == 0) red_s[warp * kFtRows + g] = sa, red_s[warp * kFtRows + g + 8] = sb;
    asm volatile("cp.async.wait_all;" ::: "memory");  // weight (first tile) and shift / scale copies of this thread
    __syncthreads();
    float ma = 0.f, mb = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) ma += red_s[w8 * kFtRows + g], mb += red_s[w8 * kFtRows + g + 8];
    ma *= 1.0f / D, mb *= 1.0f / D;
    // ---- variance (second pass over the registers)
    float qa = 0.f, qb = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      float d0 = xa[j].x - ma, d1 = xa[j].y - ma, d2 = xa[j].z - ma, d3 = xa[j].w - ma;
      qa += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      d0 = xb[j].x - mb, d1 = xb[j].y - mb, d2 = xb[j].z - mb, d3 = xb[j].w - mb;
      qb += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
    }
    qa += __shfl_xor_sync(0xffffffffu, qa, 1), qb += __shfl_xor_sync(0xffffffffu, qb, 1);
    qa += __shfl_xor_sync(0xffffffffu, qa, 2), qb += __shfl_xor_sync(0xffffffffu, qb, 2);
    float* red_q = red_s + 8 * kFtRows;
    if (t == 0) red_q[warp * kFtRows + g] = qa, red_q[warp * kFtRows + g + 8] = qb;
    __syncthreads();
    float va = 0.f, vb = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; ++w8) va += red_q[w8 * kFtRows + g], vb += red_q[w8 * kFtRows + g + 8];
    const float ra_ = 1.0f / sqrtf(va * (1.0f / D) + eps), rb_ = 1.0f / sqrtf(vb * (1.0f / D) + eps);
    // ---- modulate, split, multiply
    const int row_a = min(tile * kFtRows + g, M - 1), row_b = min(tile * kFtRows + g + 8, M - 1);
    const float* moda = smod + (row_a / T == b_lo ? 0 : 2 * D) + kbase;  // [shift | scale] of row g's image
    const float* modb = smod + (row_b / T == b_lo ? 0 : 2 * D) + kbase;
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) acc[nt][0] = acc[nt][1] = acc[nt][2] = acc[nt][3] = 0.f;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const float4 h4a = *reinterpret_cast<const float4*>(moda + 16 * j);
      const float4 c4a = *reinterpret_cast<const float4*>(moda + D + 16 * j);
      const float4 h4b = *reinterpret_cast<const float4*>(modb + 16 * j);
      const float4 c4b = *reinterpret_cast<const float4*>(modb + D + 16 * j);
      float ha[4], hb[4];
      ha[0] = (xa[j].x - ma) * ra_ * (1.0f + c4a.x) + h4a.x, ha[1] = (xa[j].y - ma) * ra_ * (1.0f + c4a.y) + h4a.y;
      ha[2] = (xa[j].z - ma) * ra_ * (1.0f + c4a.z) + h4a.z, ha[3] = (xa[j].w - ma) * ra_ * (1.0f + c4a.w) + h4a.w;
      hb[0] = (xb[j].x - mb) * rb_ * (1.0f + c4b.x) + h4b.x, hb[1] = (xb[j].y - mb) * rb_ * (1.0f + c4b.y) + h4b.y;
      hb[2] = (xb[j].z - mb) * rb_ * (1.0f + c4b.z) + h4b.z, hb[3] = (xb[j].w - mb) * rb_ * (1.0f + c4b.w) + h4b.w;
      uint32_t ah[4], al[4], bh[4], bl[4];  // hi / lo parts of the two rows' four columns
#pragma unroll
      for (int q = 0; q < 4; ++q) split_tf32(ha[q], ah[q], al[q]), split_tf32(hb[q], bh[q], bl[q]);
      // two output tiles at a time with their MMAs interleaved: a chain of six dependent MMAs per accumulator would
      // leave the tensor pipe waiting on its own latency
#pragma unroll
      for (int np = 0; np < 4; np += 2) {
        uint32_t wh[2][4], wl[2][4];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          const float4 wq = *reinterpret_cast<const float4*>(ws + ((np + u) * 8 + g) * WS + kbase + 16 * j);
          const float wv[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
          for (int q = 0; q < 4; ++q) split_tf32(wv[q], wh[u][q], wl[u][q]);
        }
        // k step 2j: slots t, t + 4 = columns +0, +1;  k step 2j + 1: columns +2, +3.  Small terms first.
#pragma unroll
        for (int u = 0; u < 2; ++u) mma_tf32_16x8x8(acc[np + u], al[0], bl[0], al[1], bl[1], wh[u][0], wh[u][1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) mma_tf32_16x8x8(acc[np + u], ah[0], bh[0], ah[1], bh[1], wl[u][0], wl[u][1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) mma_tf32_16x8x8(acc[np + u], al[2], bl[2], al[3], bl[3], wh[u][2], wh[u][3]);
#pragma unroll
        for (int u = 0; u < 2; ++u) mma_tf32_16x8x8(acc[np + u], ah[2], bh[2], ah[3], bh[3], wl[u][2], wl[u][3]);
#pragma unroll
        for (int u = 0; u < 2; ++u) mma_tf32_16x8x8(acc[np + u], ah[0], bh[0], ah[1], bh[1], wh[u][0], wh[u][1]);
#pragma unroll
        for (int u = 0; u < 2; ++u) mma_tf32_16x8x8(acc[np + u], ah[2], bh[2], ah[3], bh[3], wh[u][2], wh[u][3]);
      }
    }
    // ---- the 8 warps' partial [16 x 32] tiles meet in shared memory
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      *reinterpret_cast<float2*>(red_o + (warp * kFtRows + g) * kFtOStride + nt * 8 + 2 * t) = make_float2(acc[nt][0], acc[nt][1]);
      *reinterpret_cast<float2*>(red_o + (warp * kFtRows + g + 8) * kFtOStride + nt * 8 + 2 * t) = make_float2(acc[nt][2], acc[nt][3]);
    }
    __syncthreads();
    // output element e = ((c * p + pi) * 16 + token) * p + pj: consecutive threads walk along an image row (unpatchify)
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      const int e = tid + 256 * half;
      const int pj = e % p, r1 = e / p, tok = r1 % kFtRows, r2 = r1 / kFtRows, pi = r2 % p, c = r2 / p;
      const int o = (pi * p + pj) * Cout + c;
      const int row = tile * kFtRows + tok;
      if (row < M) {
        float v = bias[o];
#pragma unroll
        for (int w8 = 0; w8 < 8; ++w8) v += red_o[(w8 * kFtRows + tok) * kFtOStride + o];
        const int b = row / T, tt = row - b * T;
        const int hh = tt / Wp, ww = tt - hh * Wp;
        out[(((size_t)b * Cout + c) * Himg + hh * p + pi) * Himg + ww * p + pj] = v;
      }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) xa[j] = na[j], xb[j] = nb[j];
  }
  asm volatile("cp.async.wait_all;" ::: "memory");  // a CTA without tiles still owns its weight copies
}

extern "C" int ditb200_final_layer(const float* x, const float* shift, const float* scale,
                                   int mod_stride, const float* w, const float* bias, float* out,
                                   int B, int T, int D, int p, int Cout, float eps, int round_bf16,
                                   void* stream) {
  DITB_REQUIRE(x && shift && scale && w && bias && out, DITB200_EINVAL, "final_layer: null pointer");
  DITB_REQUIRE(B > 0 && T > 0 && D > 0 && D % 4 == 0 && p > 0 && Cout > 0, DITB200_EINVAL,
               "final_layer: bad shape");
  const int Wp = (int)(sqrt((double)T) + 0.5);
  DITB_REQUIRE(Wp * Wp == T, DITB200_EINVAL, "final_layer: T=%d is not a square grid", T);
  DITB_REQUIRE(mod_stride % 4 == 0 && aligned16(x) && aligned16(shift) && aligned16(scale),
               DITB200_EALIGN, "final_layer: misaligned input");
  const int NO = p * p * Cout;
  const int now = (NO + 31) / 32;
  DITB_REQUIRE(now <= 16, DITB200_EINVAL, "final_layer: p*p*Cout = %d > 512", NO);
  const int M = B * T;
  const dim3 grid((M + kFlRows - 1) / kFlRows), block(256);
  cudaStream_t st = (cudaStream_t)stream;
  static const bool simt_only = getenv("DITB200_FINAL_SIMT") != nullptr;  // measurement switch: the SIMT 32-output kernel
  if (NO == 32 && !simt_only && !round_bf16 && T >= kFtRows && aligned16(w) && (D == 384 || D == 768 || D == 1024 || D == 1152)) {
    const size_t smem = ((size_t)32 * (D + 16) + 2 * 8 * kFtRows + (size_t)8 * kFtRows * kFtOStride + (size_t)4 * D) * sizeof(float);
    const int ntiles = (M + kFtRows - 1) / kFtRows;
    int g = num_sms() > 0 ? num_sms() : 148;
    if (g > ntiles) g = ntiles;
    const int chunk = (ntiles + g - 1) / g;
    g = (ntiles + chunk - 1) / chunk;
#define FTC_CASE(DD)                                                                                           \
  case DD: {                                                                                                   \
    cudaError_t e = cudaFuncSetAttribute(final_layer_tc_kernel<DD>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                           \
    if (e != cudaSuccess) return check_cuda(e, "final_layer smem attr");                                       \
    DITB_KLAUNCH((final_layer_tc_kernel<DD>), dim3(g), dim3(256), smem, st, x, shift, scale, mod_stride, w, bias, out, \
                 M, T, p, Cout, eps, chunk);                                                       \
  } break;
    switch (D) {
      FTC_CASE(384)
      FTC_CASE(768)
      FTC_CASE(1024)
      FTC_CASE(1152)
    }
#undef FTC_CASE
    DITB_LAUNCH_CHECK("final_layer");
    return 0;
  }
  if (NO == 32 && D % 128 == 0 && aligned16(w) && (D / 128 == 3 || D / 128 == 6 || D / 128 == 8 || D / 128 == 9)) {
    const size_t smem = (size_t)32 * D * sizeof(float);
    int g = num_sms() > 0 ? num_sms() : 148;
    if (g > (M + 15) / 16) g = (M + 15) / 16;  // 8 warps x 2 rows per CTA and pass
#define FLR_CASE(NV)                                                                                           \
  case NV: {                                                                                                   \
    cudaError_t e = cudaFuncSetAttribute(final_layer_rows_kernel<NV>, cudaFuncAttributeMaxDynamicSharedMemorySize, \
                                         (int)smem);                                                           \
    if (e != cudaSuccess) return check_cuda(e, "final_layer smem attr");                                       \
    final_layer_rows_kernel<NV><<<g, 256, smem, st>>>(x, shift, scale, mod_stride, w, bias, out, M, T, p, Cout, eps, \
                                                      round_bf16);                                             \
  } break;
    switch (D / 128) {
      FLR_CASE(3)
      FLR_CASE(6)
      FLR_CASE(8)
      FLR_CASE(9)
    }
#undef FLR_CASE
    DITB_LAUNCH_CHECK("final_layer");
    return 0;
  }
#define FL_LAUNCH(NOW)                                                                          \
  {                                                                                             \
    const size_t smem = ((size_t)kFlRows * D + (size_t)(kFlKc + 4) * NOW * 32) * sizeof(float);       \
    DITB_REQUIRE(smem <= 227 * 1024, DITB200_EINVAL, "final_layer: D=%d too large", D);         \
    cudaError_t e = cudaFuncSetAttribute(final_layer_kernel<NOW>,                               \
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
    if (e != cudaSuccess) return check_cuda(e, "final_layer smem attr");                        \
    final_layer_kernel<NOW><<<grid, block, smem, st>>>(x, shift, scale, mod_stride, w, bias, out, \
                                                       M, T, D, p, Cout, eps, round_bf16);      \
  }
  if (now == 1) FL_LAUNCH(1)
  else if (now == 2) FL_LAUNCH(2)
  else if (now <= 4) FL_LAUNCH(4)
  else if (now <= 8) FL_LAUNCH(8)
  else FL_LAUNCH(16)
#undef FL_LAUNCH
  DITB_LAUNCH_CHECK("final_layer");
  return 0;
}
