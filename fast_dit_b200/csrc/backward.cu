// backward.cu — the HBM-bound kernels of the training step's backward pass (a19 in SURVEY.md §8):
// LayerNorm+modulate backward, gated-residual backward, column sums (bias gradients), the
// embedding scatter, patchify / un-unpatchify layout kernels and SiLU'.  The contractions of the
// backward pass (data and weight gradients) run on the tcgen05 GEMM with MN-major operands
// (gemm_tc.cu); attention backward lives in attention.cu.
#include "common.cuh"

namespace ditb200 {

__device__ __forceinline__ float4 load4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 load4(const __nv_bfloat16* p) {
  const uint2 pk = *reinterpret_cast<const uint2*>(p);
  const float2 a = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
  const float2 b = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
  return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void store4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void store4(__nv_bfloat16* p, float4 v) {
  uint2 pk;
  pk.x = pack_bf16x2(v.x, v.y);
  pk.y = pack_bf16x2(v.z, v.w);
  *reinterpret_cast<uint2*>(p) = pk;
}

__device__ __forceinline__ uint2 ldg_stream_u2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
// Asynchronous global -> shared copies (LDGSTS): in flight without holding registers, so ptxas cannot sink them.
__device__ __forceinline__ void cp_async_16B(void* dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_8B(void* dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// ====================================================== LayerNorm + modulate, backward
// h = xhat * (1 + scale[b]) + shift[b],  xhat = (x - mean) * rstd.
//   dx      = rstd * (g - mean(g) - xhat * mean(g * xhat)),  g = dh * (1 + scale[b])
//   dshift[b] += sum_t dh,   dscale[b] += sum_t dh * xhat
// CTA = kLnRows consecutive tokens of ONE image; one warp per row (lane owns float4 l + 32 j);
// per-column partial sums of the CTA are combined in shared memory, then one atomic per column.
constexpr int kLnRows = 32;
template <typename TDh>
__global__ void __launch_bounds__(256) ln_modulate_bwd_kernel(
    const TDh* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ scale, int mod_stride,
    const float* __restrict__ stats, float* __restrict__ dx, int accumulate, float* __restrict__ dshift,
    float* __restrict__ dscale, int dmod_stride, int T, int D) {
  extern __shared__ float red[];  // [2][D]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * kLnRows;
  const int t1 = min(T, t0 + kLnRows);
  const int nv = D >> 2;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const float* sc = scale + (size_t)b * mod_stride;
  const float invD = 1.0f / (float)D;
  for (int t = t0 + warp; t < t1; t += 8) {
    const size_t row = (size_t)b * T + t;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    const float* xr = x + row * D;
    const TDh* dr = dh + row * D;
    float s1 = 0.f, s2 = 0.f;
    for (int i = lane; i < nv; i += 32) {
      const float4 xv = load4(xr + 4 * i), dv = load4(dr + 4 * i), cv = __ldg(reinterpret_cast<const float4*>(sc) + i);
      const float g0 = dv.x * (1.f + cv.x), g1 = dv.y * (1.f + cv.y), g2 = dv.z * (1.f + cv.z), g3 = dv.w * (1.f + cv.w);
      const float h0 = (xv.x - mean) * rstd, h1 = (xv.y - mean) * rstd, h2 = (xv.z - mean) * rstd, h3 = (xv.w - mean) * rstd;
      s1 += (g0 + g1) + (g2 + g3);
      s2 += (g0 * h0 + g1 * h1) + (g2 * h2 + g3 * h3);
      atomicAdd(&red[4 * i + 0], dv.x), atomicAdd(&red[4 * i + 1], dv.y);
      atomicAdd(&red[4 * i + 2], dv.z), atomicAdd(&red[4 * i + 3], dv.w);
      atomicAdd(&red[D + 4 * i + 0], dv.x * h0), atomicAdd(&red[D + 4 * i + 1], dv.y * h1);
      atomicAdd(&red[D + 4 * i + 2], dv.z * h2), atomicAdd(&red[D + 4 * i + 3], dv.w * h3);
    }
    s1 = warp_sum(s1) * invD;
    s2 = warp_sum(s2) * invD;
    float* dxr = dx + row * D;
    for (int i = lane; i < nv; i += 32) {  // second pass: the row is in L1
      const float4 xv = load4(xr + 4 * i), dv = load4(dr + 4 * i), cv = __ldg(reinterpret_cast<const float4*>(sc) + i);
      float4 o;
      o.x = rstd * (dv.x * (1.f + cv.x) - s1 - (xv.x - mean) * rstd * s2);
      o.y = rstd * (dv.y * (1.f + cv.y) - s1 - (xv.y - mean) * rstd * s2);
      o.z = rstd * (dv.z * (1.f + cv.z) - s1 - (xv.z - mean) * rstd * s2);
      o.w = rstd * (dv.w * (1.f + cv.w) - s1 - (xv.w - mean) * rstd * s2);
      if (accumulate) {
        const float4 p = load4(dxr + 4 * i);
        o.x += p.x, o.y += p.y, o.z += p.z, o.w += p.w;
      }
      store4(dxr + 4 * i, o);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dshift + (size_t)b * dmod_stride + i, red[i]);
    atomicAdd(dscale + (size_t)b * dmod_stride + i, red[D + i]);
  }
}

// Fixed-width variant (D = NV * 128): warp per row with the row of x and dh held in registers (each read from
// HBM exactly once, like the forward kernel) and the per-column partial sums of dshift / dscale carried in
// registers across the rows a warp owns; shared memory sees one reduction per CTA instead of eight atomics per
// element.  CTA = 4 warps x kLnWarpRows consecutive tokens of ONE image.
constexpr int kLnWarpRows = 4;
template <int NV, typename TDh>
__global__ void __launch_bounds__(128) ln_modulate_bwd_rows_kernel(
    const TDh* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ scale, int mod_stride,
    const float* __restrict__ stats, float* __restrict__ dx, int accumulate, float* __restrict__ dshift,
    float* __restrict__ dscale, int dmod_stride, int T) {
  DITB_PDL_WAIT();
  constexpr int D = NV * 128;
  __shared__ float red[2 * D];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int t0 = blockIdx.x * (4 * kLnWarpRows) + warp * kLnWarpRows;
  for (int i = threadIdx.x; i < 2 * D; i += blockDim.x) red[i] = 0.f;
  __syncthreads();
  const float4* sc = reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride);
  float4 a_sh[NV], a_sc[NV];
#pragma unroll
  for (int j = 0; j < NV; ++j) a_sh[j] = a_sc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = 0; r < kLnWarpRows; ++r) {
    const int t = t0 + r;
    if (t >= T) break;
    const size_t row = (size_t)b * T + t;
    const float mean = stats[2 * row], rstd = stats[2 * row + 1];
    const float* xr = x + row * D;
    const TDh* dr = dh + row * D;
    float4 xv[NV], dv[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      xv[j] = ldg_stream_f4(reinterpret_cast<const float4*>(xr) + lane + 32 * j);
      dv[j] = load4(dr + 4 * (lane + 32 * j));
    }
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      const float4 cv = __ldg(sc + lane + 32 * j);
      // xv becomes xhat, the column sums take dh and dh * xhat, s1 / s2 the row means of g and g * xhat
      xv[j].x = (xv[j].x - mean) * rstd, xv[j].y = (xv[j].y - mean) * rstd;
      xv[j].z = (xv[j].z - mean) * rstd, xv[j].w = (xv[j].w - mean) * rstd;
      a_sh[j].x += dv[j].x, a_sh[j].y += dv[j].y, a_sh[j].z += dv[j].z, a_sh[j].w += dv[j].w;
      a_sc[j].x += dv[j].x * xv[j].x, a_sc[j].y += dv[j].y * xv[j].y;
      a_sc[j].z += dv[j].z * xv[j].z, a_sc[j].w += dv[j].w * xv[j].w;
      dv[j].x *= 1.f + cv.x, dv[j].y *= 1.f + cv.y, dv[j].z *= 1.f + cv.z, dv[j].w *= 1.f + cv.w;  // g
      s1 += (dv[j].x + dv[j].y) + (dv[j].z + dv[j].w);
      s2 += (dv[j].x * xv[j].x + dv[j].y * xv[j].y) + (dv[j].z * xv[j].z + dv[j].w * xv[j].w);
    }
    s1 = warp_sum(s1) * (1.0f / D);
    s2 = warp_sum(s2) * (1.0f / D);
    float* dxr = dx + row * D;
#pragma unroll
    for (int j = 0; j < NV; ++j) {
      float4 o;
      o.x = rstd * (dv[j].x - s1 - xv[j].x * s2), o.y = rstd * (dv[j].y - s1 - xv[j].y * s2);
      o.z = rstd * (dv[j].z - s1 - xv[j].z * s2), o.w = rstd * (dv[j].w - s1 - xv[j].w * s2);
      float4* dst = reinterpret_cast<float4*>(dxr) + lane + 32 * j;
      if (accumulate) {
        const float4 p = *dst;
        o.x += p.x, o.y += p.y, o.z += p.z, o.w += p.w;
      }
      *dst = o;
    }
  }
#pragma unroll
  for (int j = 0; j < NV; ++j) {
    const int c = 4 * (lane + 32 * j);
    atomicAdd(&red[c], a_sh[j].x), atomicAdd(&red[c + 1], a_sh[j].y), atomicAdd(&red[c + 2], a_sh[j].z), atomicAdd(&red[c + 3], a_sh[j].w);
    atomicAdd(&red[D + c], a_sc[j].x), atomicAdd(&red[D + c + 1], a_sc[j].y), atomicAdd(&red[D + c + 2], a_sc[j].z), atomicAdd(&red[D + c + 3], a_sc[j].w);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < D; i += blockDim.x) {
    atomicAdd(dshift + (size_t)b * dmod_stride + i, red[i]);
    atomicAdd(dscale + (size_t)b * dmod_stride + i, red[D + i]);
  }
}

// ============================ LayerNorm + modulate backward, fused with the NEXT gated-residual backward
// In the backward chain every LayerNorm backward updates the gradient of the residual stream (dx), and the kernel
// that follows it in data order is the gated-residual backward of the branch that joined the stream just before this
// LayerNorm in the forward pass: dy = dx * gate[b], dgate[b] += sum_t dx * y, dbias += sum dy.  Run separately, that
// kernel re-reads the dx this one has just written.  Here both happen in one pass over the rows.
//
// Mapping: a THREAD owns 4 consecutive columns of every row (CTA = D/4 threads = whole rows), so the four per-image
// column sums (dshift, dscale, dgate, dbias) are 16 registers per thread however wide D is, and leave the CTA as
// one 128-bit reduction each.  Rows go kFbRows at a time: all their loads are requested first (48 bytes per thread and
// row in flight) — x and dh into registers, the operands of the second half (the previous dx and y) as asynchronous
// copies into the thread's own shared-memory slots, because ptxas otherwise sinks those loads to their first use
// behind the barrier and the batch pays three memory latencies instead of one.  The 2 * kFbRows row sums cross the
// warp by a transposing shuffle reduction (9 shuffles, not 40) and the CTA through a double-buffered shared array
// (one barrier per batch).  Persistent: each CTA takes one
// contiguous range of row batches (at most a few images: the column sums are flushed when the image changes).
constexpr int kFbRows = 4;
template <int NT, typename TDh, bool kGate>
__global__ void __launch_bounds__(NT, (NT <= 128 ? 4 : 2)) ln_modulate_bwd_cols_kernel(
    const TDh* __restrict__ dh, const float* __restrict__ x, const float* __restrict__ scale, int mod_stride,
    const float* __restrict__ stats, float* __restrict__ dx, int accumulate, float* __restrict__ dshift,
    float* __restrict__ dscale, int dmod_stride, const __nv_bfloat16* __restrict__ y, const float* __restrict__ gate,
    int gate_stride, __nv_bfloat16* __restrict__ dy, float* __restrict__ dgate, int dgate_stride,
    float* __restrict__ dbias, int B, int T, int chunk) {
  DITB_PDL_WAIT();
  constexpr int D = NT * 4, NW = NT / 32;
  __shared__ __align__(16) float red[2][NW][2 * kFbRows];
  __shared__ __align__(16) float4 s_pv[kFbRows][NT];            // previous dx (accumulate != 0), slot [row][thread]
  __shared__ __align__(8) uint2 s_y[kGate ? kFbRows : 1][NT];   // y (bf16 x 4)
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int c = tid * 4;
  const int bpi = T / kFbRows;  // row batches per image (T % kFbRows == 0: a batch never straddles two images)
  const int nb = B * bpi;
  const int q0 = blockIdx.x * chunk, q1 = min(nb, q0 + chunk);
  if (q0 >= q1) return;
  int b_cur = q0 / bpi;
  float4 cv = __ldg(reinterpret_cast<const float4*>(scale + (size_t)b_cur * mod_stride + c));
  float4 gv = make_float4(0.f, 0.f, 0.f, 0.f);
  if constexpr (kGate) gv = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b_cur * gate_stride + c));
  const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
  float4 a_sh = zero4, a_sc = zero4, a_g = zero4, a_b = zero4;
  auto flush = [&](int b) {
    atomicAdd(reinterpret_cast<float4*>(dshift + (size_t)b * dmod_stride + c), a_sh);
    atomicAdd(reinterpret_cast<float4*>(dscale + (size_t)b * dmod_stride + c), a_sc);
    if constexpr (kGate) {
      atomicAdd(reinterpret_cast<float4*>(dgate + (size_t)b * dgate_stride + c), a_g);
      if (dbias != nullptr) atomicAdd(reinterpret_cast<float4*>(dbias + c), a_b);
    }
  };
  int buf = 0;
  for (int q = q0; q < q1; ++q, buf ^= 1) {
    const int b = q / bpi;
    if (b != b_cur) {
      flush(b_cur);
      a_sh = a_sc = a_g = a_b = zero4;
      b_cur = b;
      cv = __ldg(reinterpret_cast<const float4*>(scale + (size_t)b * mod_stride + c));
      if constexpr (kGate) gv = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * gate_stride + c));
    }
    const size_t off0 = (size_t)q * kFbRows * D + c;  // row = q * kFbRows (bpi * kFbRows == T)
    float4 xv[kFbRows], dv[kFbRows];
    float2 st[kFbRows];
#pragma unroll
    for (int u = 0; u < kFbRows; ++u) {
      const size_t off = off0 + (size_t)u * D;
      if (accumulate) cp_async_16B(&s_pv[u][tid], dx + off);
      if constexpr (kGate) cp_async_8B(&s_y[u][tid], y + off);
    }
#pragma unroll
    for (int u = 0; u < kFbRows; ++u) {
      const size_t off = off0 + (size_t)u * D;
      xv[u] = ldg_stream_f4(reinterpret_cast<const float4*>(x + off));
      dv[u] = load4(dh + off);
      st[u] = __ldg(reinterpret_cast<const float2*>(stats) + (size_t)q * kFbRows + u);
    }
    float part[2 * kFbRows];
#pragma unroll
    for (int u = 0; u < kFbRows; ++u) {
      const float mean = st[u].x, rstd = st[u].y;
      // xv becomes xhat, the column sums take dh and dh * xhat, dv becomes g = dh * (1 + scale)
      xv[u].x = (xv[u].x - mean) * rstd, xv[u].y = (xv[u].y - mean) * rstd;
      xv[u].z = (xv[u].z - mean) * rstd, xv[u].w = (xv[u].w - mean) * rstd;
      a_sh.x += dv[u].x, a_sh.y += dv[u].y, a_sh.z += dv[u].z, a_sh.w += dv[u].w;
      a_sc.x += dv[u].x * xv[u].x, a_sc.y += dv[u].y * xv[u].y, a_sc.z += dv[u].z * xv[u].z, a_sc.w += dv[u].w * xv[u].w;
      dv[u].x *= 1.f + cv.x, dv[u].y *= 1.f + cv.y, dv[u].z *= 1.f + cv.z, dv[u].w *= 1.f + cv.w;
      part[u] = (dv[u].x + dv[u].y) + (dv[u].z + dv[u].w);
      part[kFbRows + u] = (dv[u].x * xv[u].x + dv[u].y * xv[u].y) + (dv[u].z * xv[u].z + dv[u].w * xv[u].w);
    }
    // transposing warp reduction of the 8 partial sums: after the three exchange steps lane l holds value
    // number (l >> 2) & 7 summed over the 8 lanes that share its low two bits; two plain steps finish it
    static_assert(kFbRows == 4, "the shuffle network below reduces exactly 8 values");
    float w4[4], w2[2], w1;
    {
      const bool hi = lane & 16;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float send = hi ? part[i] : part[i + 4], keep = hi ? part[i + 4] : part[i];
        w4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
      }
    }
    {
      const bool hi = lane & 8;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const float send = hi ? w4[i] : w4[i + 2], keep = hi ? w4[i + 2] : w4[i];
        w2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
      }
    }
    {
      const bool hi = lane & 4;
      const float send = hi ? w2[0] : w2[1], keep = hi ? w2[1] : w2[0];
      w1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
    }
    w1 += __shfl_xor_sync(0xffffffffu, w1, 2);
    w1 += __shfl_xor_sync(0xffffffffu, w1, 1);
    // lane bits 4,3,2 = bits 2,1,0 of the value's number
    if ((lane & 3) == 0) red[buf][warp][((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = w1;
    __syncthreads();
    float4 t0 = zero4, t1 = zero4;  // row sums of g (rows 0..3) and of g * xhat
#pragma unroll
    for (int w = 0; w < NW; ++w) {
      const float4 p0 = *reinterpret_cast<const float4*>(&red[buf][w][0]);
      const float4 p1 = *reinterpret_cast<const float4*>(&red[buf][w][4]);
      t0.x += p0.x, t0.y += p0.y, t0.z += p0.z, t0.w += p0.w;
      t1.x += p1.x, t1.y += p1.y, t1.z += p1.z, t1.w += p1.w;
    }
    const float s1[kFbRows] = {t0.x * (1.0f / D), t0.y * (1.0f / D), t0.z * (1.0f / D), t0.w * (1.0f / D)};
    const float s2[kFbRows] = {t1.x * (1.0f / D), t1.y * (1.0f / D), t1.z * (1.0f / D), t1.w * (1.0f / D)};
    cp_async_wait_all();  // this thread's own slots: no barrier needed
#pragma unroll
    for (int u = 0; u < kFbRows; ++u) {
      const size_t off = off0 + (size_t)u * D;
      const float rstd = st[u].y;
      float4 o;
      o.x = rstd * (dv[u].x - s1[u] - xv[u].x * s2[u]), o.y = rstd * (dv[u].y - s1[u] - xv[u].y * s2[u]);
      o.z = rstd * (dv[u].z - s1[u] - xv[u].z * s2[u]), o.w = rstd * (dv[u].w - s1[u] - xv[u].w * s2[u]);
      if (accumulate) {
        const float4 pv = s_pv[u][tid];
        o.x += pv.x, o.y += pv.y, o.z += pv.z, o.w += pv.w;
      }
      *reinterpret_cast<float4*>(dx + off) = o;
      if constexpr (kGate) {
        const uint2 yv = s_y[u][tid];
        const float4 yy = make_float4(__uint_as_float(yv.x << 16), __uint_as_float(yv.x & 0xffff0000u),
                                      __uint_as_float(yv.y << 16), __uint_as_float(yv.y & 0xffff0000u));
        const float4 d4 = make_float4(o.x * gv.x, o.y * gv.y, o.z * gv.z, o.w * gv.w);
        store4(dy + off, d4);
        a_g.x += o.x * yy.x, a_g.y += o.y * yy.y, a_g.z += o.z * yy.z, a_g.w += o.w * yy.w;
        a_b.x += d4.x, a_b.y += d4.y, a_b.z += d4.z, a_b.w += d4.w;
      }
    }
  }
  flush(b_cur);
}

// ============================================================ gated residual, backward
// x_out = x + gate[b] * y:  dy = dx_out * gate[b];  dgate[b] += sum_t dx_out * y;  dbias += sum_rows dy.
// Thread = 4 columns, CTA = kGrRows tokens of one image: coalesced 128-bit row accesses, the
// reductions over tokens stay in registers, one atomic per column per CTA.
constexpr int kGrRows = 8;
template <typename TY>
__global__ void gate_resid_bwd_kernel(const float* __restrict__ dxo, const TY* __restrict__ y,
                                      const float* __restrict__ gate, int gate_stride, TY* __restrict__ dy,
                                      float* __restrict__ dgate, int dgate_stride, float* __restrict__ dbias, int T,
                                      int D) {
  DITB_PDL_WAIT();
  const int b = blockIdx.y;
  const int c = (blockIdx.z * blockDim.x + threadIdx.x) * 4;
  if (c >= D) return;
  const int t0 = blockIdx.x * kGrRows, t1 = min(T, t0 + kGrRows);
  const float4 g = __ldg(reinterpret_cast<const float4*>(gate + (size_t)b * gate_stride + c));
  float4 ag = make_float4(0.f, 0.f, 0.f, 0.f), ab = ag;
  for (int tt = t0; tt < t1; tt += 8) {  // 8 rows of loads in flight per thread
    float4 d[8], yy[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (tt + u < t1) {
        const size_t off = ((size_t)b * T + tt + u) * D + c;
        d[u] = load4(dxo + off), yy[u] = load4(y + off);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (tt + u < t1) {
        const size_t off = ((size_t)b * T + tt + u) * D + c;
        const float4 o = make_float4(d[u].x * g.x, d[u].y * g.y, d[u].z * g.z, d[u].w * g.w);
        store4(dy + off, o);
        ag.x += d[u].x * yy[u].x, ag.y += d[u].y * yy[u].y, ag.z += d[u].z * yy[u].z, ag.w += d[u].w * yy[u].w;
        ab.x += o.x, ab.y += o.y, ab.z += o.z, ab.w += o.w;
      }
    }
  }
  atomicAdd(reinterpret_cast<float4*>(dgate + (size_t)b * dgate_stride + c), ag);  // one 128-bit reduction
  if (dbias != nullptr) atomicAdd(reinterpret_cast<float4*>(dbias + c), ab);
}

// ======================================================================== column sums
// out[c] += sum_r in[r, c]: bias gradients of the QKV / fc1 Linears (HBM-bound: every element is read once).
// A block owns kCsCols consecutive columns x kCsRows rows: lane l of every warp owns kCsVec consecutive columns
// (16-byte loads for bf16, 2 x 16 for f32), the block's 8 warps take interleaved rows with 8 loads in flight per
// lane, the warps' partial sums meet in shared memory and the block issues ONE vector atomic per 4 columns.
constexpr int kCsRows = 256, kCsVec = 8, kCsCols = 32 * kCsVec, kCsWarps = 8;
template <typename TIn> struct CsRaw;
template <> struct CsRaw<float> {
  float4 a, b;
  __device__ __forceinline__ void load(const float* p) {
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w) : "l"(p) : "memory");
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w) : "l"(p + 4) : "memory");
  }
  __device__ __forceinline__ void zero() { a = b = make_float4(0.f, 0.f, 0.f, 0.f); }
  __device__ __forceinline__ void add_to(float (&acc)[8]) const {
    acc[0] += a.x, acc[1] += a.y, acc[2] += a.z, acc[3] += a.w, acc[4] += b.x, acc[5] += b.y, acc[6] += b.z, acc[7] += b.w;
  }
};
template <> struct CsRaw<__nv_bfloat16> {
  uint4 pk;
  __device__ __forceinline__ void load(const __nv_bfloat16* p) {
    // pinned (volatile + memory clobber): the batch of loads stays a batch instead of being sunk to its first use
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(pk.x), "=r"(pk.y), "=r"(pk.z), "=r"(pk.w) : "l"(p) : "memory");
  }
  __device__ __forceinline__ void zero() { pk = make_uint4(0u, 0u, 0u, 0u); }
  __device__ __forceinline__ void add_to(float (&acc)[8]) const {
    // bf16 -> f32 is a 16-bit shift: low half = element 0, high half = element 1
    acc[0] += __uint_as_float(pk.x << 16), acc[1] += __uint_as_float(pk.x & 0xffff0000u);
    acc[2] += __uint_as_float(pk.y << 16), acc[3] += __uint_as_float(pk.y & 0xffff0000u);
    acc[4] += __uint_as_float(pk.z << 16), acc[5] += __uint_as_float(pk.z & 0xffff0000u);
    acc[6] += __uint_as_float(pk.w << 16), acc[7] += __uint_as_float(pk.w & 0xffff0000u);
  }
};
template <typename TIn>
__global__ void __launch_bounds__(32 * kCsWarps, 4) colsum_kernel(const TIn* __restrict__ in, float* __restrict__ out, int R, int C) {
  DITB_PDL_WAIT();
  __shared__ float part[kCsWarps][kCsCols];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * kCsCols + lane * kCsVec;
  const int r0 = blockIdx.y * kCsRows, r1 = min(R, r0 + kCsRows);
  float acc[kCsVec];
#pragma unroll
  for (int j = 0; j < kCsVec; ++j) acc[j] = 0.f;
  if (c < C) {  // C % 8 == 0 on this path: a lane's 8 columns are all inside or all outside
    constexpr int kBatch = sizeof(TIn) == 2 ? 8 : 4;  // rows (128 bytes per lane) requested before the first is consumed
    for (int rr = r0 + warp; rr < r1; rr += kBatch * kCsWarps) {
      CsRaw<TIn> raw[kBatch];
#pragma unroll
      for (int u = 0; u < kBatch; ++u) {
        const int r = rr + u * kCsWarps;
        if (r < r1) raw[u].load(in + (size_t)r * C + c); else raw[u].zero();
      }
#pragma unroll
      for (int u = 0; u < kBatch; ++u) raw[u].add_to(acc);
    }
  }
#pragma unroll
  for (int j = 0; j < kCsVec; ++j) part[warp][lane * kCsVec + j] = acc[j];
  __syncthreads();
  const int col = threadIdx.x * 4;  // 64 threads finish the block's 256 columns, 4 each
  if (col < kCsCols && blockIdx.x * kCsCols + col < C) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int w = 0; w < kCsWarps; ++w) {
      const float4 p4 = *reinterpret_cast<const float4*>(&part[w][col]);
      s.x += p4.x, s.y += p4.y, s.z += p4.z, s.w += p4.w;
    }
    atomicAdd(reinterpret_cast<float4*>(out + blockIdx.x * kCsCols + col), s);  // one 128-bit reduction
  }
}

// columns not a multiple of 8 (C % 4 == 0): the 4-columns-per-thread form
constexpr int kCs4Rows = 64;
template <typename TIn>
__global__ void __launch_bounds__(256) colsum4_kernel(const TIn* __restrict__ in, float* __restrict__ out, int R, int C) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= C) return;
  const int r0 = blockIdx.y * kCs4Rows, r1 = min(R, r0 + kCs4Rows);
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int rr = r0; rr < r1; rr += 8) {
    float4 v[8];
#pragma unroll
    for (int u = 0; u < 8; ++u)
      v[u] = (rr + u < r1) ? load4(in + (size_t)(rr + u) * C + c) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) a.x += v[u].x, a.y += v[u].y, a.z += v[u].z, a.w += v[u].w;
  }
  atomicAdd(reinterpret_cast<float4*>(out + c), a);
}

// ===================================================== adaLN modulation Linear, weight + bias gradient
// dW[r, c] = sum_b dmod[b, r] * sc[b, c],  dbias[r] = sum_b dmod[b, r]   (models_original.py:113-116 backward; b = image,
// sc = silu(c) in bf16 as the forward GEMM read it).  The contraction runs over the BATCH only (32 images at C4), so
// as a tensor-core GEMM it is one k block of mostly padding behind a 32 MB f32 store; here it is what it is: an
// outer-product accumulation bound by that store.  CTA = 64 rows x 128 columns, thread = 8 rows x 4 columns, the two
// operand tiles of 32 images at a time in shared memory (per image and thread: two broadcast 128-bit reads of
// dmod, one of sc, 32 FMAs).  dmod is read in f32 straight from the buffer the backward kernels reduce into — no
// bf16 copy — and the first column tile of every row tile also leaves the bias gradient.
constexpr int kAwRows = 64, kAwCols = 128, kAwImgs = 32;
__global__ void __launch_bounds__(256) adaln_wgrad_kernel(const float* __restrict__ dmod, int dmod_stride,
                                                          const __nv_bfloat16* __restrict__ sc, float* __restrict__ dw,
                                                          float* __restrict__ dbias, int N, int R, int D) {
  DITB_PDL_WAIT();
  __shared__ __align__(16) float s_d[kAwImgs][kAwRows];
  __shared__ __align__(16) uint2 s_s[kAwImgs][kAwCols / 4];  // bf16 x 4 per entry
  const int tid = threadIdx.x, tr = tid >> 5, tc = tid & 31;
  const int r0 = blockIdx.x * kAwRows, c0 = blockIdx.y * kAwCols;
  float4 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  float bsum = 0.f;
  for (int b0 = 0; b0 < N; b0 += kAwImgs) {
    const int nb = min(kAwImgs, N - b0);
    if (b0 > 0) __syncthreads();
    // both operand tiles as asynchronous copies (2 x 16 + 4 x 8 bytes per thread, all in flight at once; register
    // loads get serialised load -> store -> load by ptxas); images past nb are zero rows
#pragma unroll
    for (int k = 0; k < 2; ++k) {  // dmod: 16 float4 per image
      const int i = tid + 256 * k, b = i >> 4, j = i & 15;
      if (b < nb) cp_async_16B(&s_d[b][4 * j], dmod + (size_t)(b0 + b) * dmod_stride + r0 + 4 * j);
      else *reinterpret_cast<float4*>(&s_d[b][4 * j]) = make_float4(0.f, 0.f, 0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {  // sc: 32 groups of 4 bf16 per image
      const int i = tid + 256 * k, b = i >> 5, j = i & 31;
      if (b < nb) cp_async_8B(&s_s[b][j], sc + (size_t)(b0 + b) * D + c0 + 4 * j);
      else s_s[b][j] = make_uint2(0u, 0u);
    }
    cp_async_wait_all();
    __syncthreads();
#pragma unroll 4
    for (int b = 0; b < kAwImgs; ++b) {  // images past nb are zero rows
      const float4 d0 = *reinterpret_cast<const float4*>(&s_d[b][tr * 8]);
      const float4 d1 = *reinterpret_cast<const float4*>(&s_d[b][tr * 8 + 4]);
      const uint2 sr = s_s[b][tc];
      const float4 s4 = make_float4(__uint_as_float(sr.x << 16), __uint_as_float(sr.x & 0xffff0000u),
                                    __uint_as_float(sr.y << 16), __uint_as_float(sr.y & 0xffff0000u));
      const float d[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        acc[i].x = fmaf(d[i], s4.x, acc[i].x), acc[i].y = fmaf(d[i], s4.y, acc[i].y);
        acc[i].z = fmaf(d[i], s4.z, acc[i].z), acc[i].w = fmaf(d[i], s4.w, acc[i].w);
      }
    }
    if (blockIdx.y == 0 && tid < kAwRows) {
      for (int b = 0; b < nb; ++b) bsum += s_d[b][tid];
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i)
    *reinterpret_cast<float4*>(dw + (size_t)(r0 + tr * 8 + i) * D + c0 + tc * 4) = acc[i];
  if (dbias != nullptr && blockIdx.y == 0 && tid < kAwRows) dbias[r0 + tid] = bsum;
}

// ================================================================ label embed, backward
__global__ void label_embed_bwd_kernel(const float* __restrict__ dc, const int64_t* __restrict__ y,
                                       float* __restrict__ dtable, int B, int D, int num_rows) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * D) return;
  const int b = idx / D, d = idx - b * D;
  long long row = y[b];
  if (row < 0) row = 0;
  if (row >= num_rows) row = num_rows - 1;
  atomicAdd(dtable + (size_t)row * D + d, dc[idx]);
}

// ==================================================================== layout kernels
// patches[(b,t), (c,i,j)] = x[b, c, hp*p + i, wp*p + j]: the im2col of the patch-embed conv, in the
// flattened conv-weight order, bf16 (B operand of the patch-embed weight gradient).
template <typename TOut>
__global__ void patchify_kernel(const float* __restrict__ x, TOut* __restrict__ out, int B, int C, int H,
                                int W, int p) {
  const int Hp = H / p, Wp = W / p, K = C * p * p;
  const size_t n = (size_t)B * Hp * Wp * K;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int k = (int)(idx % K);
    const size_t tok = idx / K;
    const int t = (int)(tok % (Hp * Wp)), b = (int)(tok / (Hp * Wp));
    const int hp = t / Wp, wp = t - hp * Wp;
    const int c = k / (p * p), r = k - c * p * p;
    const int i = r / p, j = r - i * p;
    const float v = x[(((size_t)b * C + c) * H + hp * p + i) * W + wp * p + j];
    if constexpr (sizeof(TOut) == 2) out[idx] = __float2bfloat16_rn(v); else out[idx] = v;
  }
}

// dz[(b, h, w), (pi, pj, c)] = dout[b, c, h*p + pi, w*p + pj]: inverse of DiT.unpatchify's permutation
// (models_original.py:228-230) applied to the gradient of the model output.
__global__ void unpatchify_bwd_kernel(const float* __restrict__ dout, __nv_bfloat16* __restrict__ dz, int B, int Cout,
                                      int Hp, int p) {
  const int NO = p * p * Cout, Himg = Hp * p;
  const size_t n = (size_t)B * Hp * Hp * NO;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) {
    const int o = (int)(idx % NO);
    const size_t tok = idx / NO;
    const int t = (int)(tok % (Hp * Hp)), b = (int)(tok / (Hp * Hp));
    const int h = t / Hp, w = t - h * Hp;
    const int c = o % Cout, pq = o / Cout;
    const int pi = pq / p, pj = pq - pi * p;
    dz[idx] = __float2bfloat16_rn(dout[(((size_t)b * Cout + c) * Himg + h * p + pi) * Himg + w * p + pj]);
  }
}

// dpre = dact * silu'(pre),  silu'(v) = s (1 + v (1 - s)),  s = sigmoid(v)
__global__ void silu_bwd_kernel(const float* __restrict__ dact, const float* __restrict__ pre, float* __restrict__ out,
                                int accumulate, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const float v = pre[i];
    const float s = 1.0f / (1.0f + expf(-v));
    const float r = dact[i] * s * (1.0f + v * (1.0f - s));
    out[i] = accumulate ? out[i] + r : r;
  }
}

}  // namespace ditb200

using namespace ditb200;

extern "C" int ditb200_ln_modulate_bwd(const void* dh, int dh_dtype, const float* x, const float* scale,
                                       int mod_stride, const float* stats, float* dx, int accumulate,
                                       float* dshift, float* dscale, int dmod_stride, int B, int T, int D,
                                       void* stream) {
  DITB_REQUIRE(dh && x && scale && stats && dx && dshift && dscale, DITB200_EINVAL, "ln_modulate_bwd: null pointer");
  DITB_REQUIRE(B > 0 && T > 0 && D > 0 && D % 4 == 0 && B <= 65535, DITB200_EINVAL, "ln_modulate_bwd: bad shape");
  DITB_REQUIRE(aligned16(dh) && aligned16(x) && aligned16(dx) && aligned16(scale) && mod_stride % 4 == 0,
               DITB200_EALIGN, "ln_modulate_bwd: misaligned pointer or stride");
  cudaStream_t st = (cudaStream_t)stream;
  if (D % 128 == 0 && (D / 128 == 3 || D / 128 == 6 || D / 128 == 8 || D / 128 == 9)) {
    dim3 rgrid((T + 4 * kLnWarpRows - 1) / (4 * kLnWarpRows), B);
#define LNB_CASE(NV)                                                                                             \
  case NV:                                                                                                       \
    if (dh_dtype == DITB200_BF16)                                                                                \
      DITB_KLAUNCH((ln_modulate_bwd_rows_kernel<NV, __nv_bfloat16>), rgrid, dim3(128), 0, st,                    \
          reinterpret_cast<const __nv_bfloat16*>(dh), x, scale, mod_stride, stats, dx, accumulate, dshift, dscale, \
          dmod_stride, T);                                                                                       \
    else                                                                                                         \
      DITB_KLAUNCH((ln_modulate_bwd_rows_kernel<NV, float>), rgrid, dim3(128), 0, st, reinterpret_cast<const float*>(dh), x, scale, \
                                                                   mod_stride, stats, dx, accumulate, dshift,    \
                                                                   dscale, dmod_stride, T);                      \
    break;
    switch (D / 128) {
      LNB_CASE(3)
      LNB_CASE(6)
      LNB_CASE(8)
      LNB_CASE(9)
    }
#undef LNB_CASE
    DITB_LAUNCH_CHECK("ln_modulate_bwd");
    return 0;
  }
  dim3 grid((T + kLnRows - 1) / kLnRows, B);
  const size_t smem = (size_t)2 * D * sizeof(float);
  if (dh_dtype == DITB200_BF16)
    ln_modulate_bwd_kernel<__nv_bfloat16><<<grid, 256, smem, st>>>(reinterpret_cast<const __nv_bfloat16*>(dh), x, scale,
                                                                  mod_stride, stats, dx, accumulate, dshift, dscale,
                                                                  dmod_stride, T, D);
  else
    ln_modulate_bwd_kernel<float><<<grid, 256, smem, st>>>(reinterpret_cast<const float*>(dh), x, scale, mod_stride,
                                                          stats, dx, accumulate, dshift, dscale, dmod_stride, T, D);
  DITB_LAUNCH_CHECK("ln_modulate_bwd");
  return 0;
}

// Fused entry: LayerNorm+modulate backward, then (y != NULL) the gated-residual backward of the branch that joined
// the stream in front of this LayerNorm, on the dx just formed.  See ln_modulate_bwd_cols_kernel.
template <int NT, typename TDh, bool kGate>
static int launch_ln_bwd_cols(const void* dh, const float* x, const float* scale, int mod_stride, const float* stats,
                              float* dx, int accumulate, float* dshift, float* dscale, int dmod_stride, const void* y,
                              const float* gate, int gate_stride, void* dy, float* dgate, int dgate_stride,
                              float* dbias, int B, int T, cudaStream_t st) {
  static int per_sm = 0;  // resident CTAs per SM of this instantiation (a pure function of the kernel: racing writers agree)
  if (per_sm == 0) {
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ln_modulate_bwd_cols_kernel<NT, TDh, kGate>, NT, 0);
    if (e != cudaSuccess) return check_cuda(e, "ln_modulate_bwd_gate occupancy");
    per_sm = n > 0 ? n : 1;
  }
  const int nb = B * (T / kFbRows);
  int grid = num_sms() * per_sm;
  if (grid > nb) grid = nb;
  const int chunk = (nb + grid - 1) / grid;
  grid = (nb + chunk - 1) / chunk;
  DITB_KLAUNCH((ln_modulate_bwd_cols_kernel<NT, TDh, kGate>), dim3(grid), dim3(NT), 0, st,
               reinterpret_cast<const TDh*>(dh), x, scale, mod_stride, stats, dx, accumulate, dshift, dscale, dmod_stride,
               reinterpret_cast<const __nv_bfloat16*>(y), gate, gate_stride, reinterpret_cast<__nv_bfloat16*>(dy), dgate,
               dgate_stride, dbias, B, T, chunk);
  DITB_LAUNCH_CHECK("ln_modulate_bwd_gate");
  return 0;
}

extern "C" int ditb200_ln_modulate_bwd_gate(const void* dh, int dh_dtype, const float* x, const float* scale,
                                            int mod_stride, const float* stats, float* dx, int accumulate,
                                            float* dshift, float* dscale, int dmod_stride, const void* y,
                                            const float* gate, int gate_stride, void* dy, float* dgate,
                                            int dgate_stride, float* dbias, int B, int T, int D, void* stream) {
  DITB_REQUIRE(dh && x && scale && stats && dx && dshift && dscale, DITB200_EINVAL, "ln_modulate_bwd_gate: null pointer");
  DITB_REQUIRE(y == nullptr || (gate && dy && dgate), DITB200_EINVAL, "ln_modulate_bwd_gate: y needs gate, dy and dgate");
  DITB_REQUIRE(B > 0 && T > 0 && T % kFbRows == 0 && (D == 384 || D == 768 || D == 1024 || D == 1152), DITB200_EINVAL,
               "ln_modulate_bwd_gate: bad shape B=%d T=%d D=%d (T %% 4 == 0, D in 384, 768, 1024, 1152)", B, T, D);
  DITB_REQUIRE((long long)B * T * D < (1ll << 40), DITB200_EINVAL, "ln_modulate_bwd_gate: too large");
  DITB_REQUIRE(dh_dtype == DITB200_BF16 || dh_dtype == DITB200_F32, DITB200_EINVAL, "ln_modulate_bwd_gate: bad dh_dtype");
  DITB_REQUIRE(aligned16(dh) && aligned16(x) && aligned16(dx) && aligned16(scale) && mod_stride % 4 == 0 &&
                   aligned16(dshift) && aligned16(dscale) && dmod_stride % 4 == 0 && aligned16(stats),
               DITB200_EALIGN, "ln_modulate_bwd_gate: misaligned pointer or stride");
  if (y != nullptr)
    DITB_REQUIRE(aligned16(y) && aligned16(dy) && aligned16(gate) && gate_stride % 4 == 0 && aligned16(dgate) &&
                     dgate_stride % 4 == 0 && (!dbias || aligned16(dbias)),
                 DITB200_EALIGN, "ln_modulate_bwd_gate: misaligned y / dy / gate / dgate / dbias");
  cudaStream_t st = (cudaStream_t)stream;
#define LNG_ARGS dh, x, scale, mod_stride, stats, dx, accumulate, dshift, dscale, dmod_stride, y, gate, gate_stride, dy, dgate, dgate_stride, dbias, B, T, st
#define LNG_CASE(NT)                                                                                           \
  case NT * 4:                                                                                                 \
    if (dh_dtype == DITB200_BF16)                                                                              \
      return y ? launch_ln_bwd_cols<NT, __nv_bfloat16, true>(LNG_ARGS) : launch_ln_bwd_cols<NT, __nv_bfloat16, false>(LNG_ARGS); \
    return y ? launch_ln_bwd_cols<NT, float, true>(LNG_ARGS) : launch_ln_bwd_cols<NT, float, false>(LNG_ARGS);
  switch (D) {
    LNG_CASE(96)
    LNG_CASE(192)
    LNG_CASE(256)
    LNG_CASE(288)
  }
#undef LNG_CASE
#undef LNG_ARGS
  return DITB200_EINVAL;
}

extern "C" int ditb200_gate_resid_bwd(const float* dx_out, const void* y, int y_dtype, const float* gate,
                                      int gate_stride, void* dy, int dy_dtype, float* dgate, int dgate_stride,
                                      float* dbias, int B, int T, int D, void* stream) {
  DITB_REQUIRE(dx_out && y && gate && dy && dgate, DITB200_EINVAL, "gate_resid_bwd: null pointer");
  DITB_REQUIRE(y_dtype == dy_dtype, DITB200_EINVAL, "gate_resid_bwd: y and dy must share a dtype");
  DITB_REQUIRE(B > 0 && T > 0 && D > 0 && D % 4 == 0 && B <= 65535, DITB200_EINVAL, "gate_resid_bwd: bad shape");
  DITB_REQUIRE(aligned16(dx_out) && aligned16(y) && aligned16(dy) && aligned16(gate) && gate_stride % 4 == 0,
               DITB200_EALIGN, "gate_resid_bwd: misaligned pointer or stride");
  DITB_REQUIRE(aligned16(dgate) && dgate_stride % 4 == 0 && (!dbias || aligned16(dbias)), DITB200_EALIGN,
               "gate_resid_bwd: dgate / dbias must be 16-byte aligned (128-bit reductions)");
  const int threads = 256;
  dim3 grid((T + kGrRows - 1) / kGrRows, B, (D / 4 + threads - 1) / threads);
  cudaStream_t st = (cudaStream_t)stream;
  if (y_dtype == DITB200_BF16)
    DITB_KLAUNCH((gate_resid_bwd_kernel<__nv_bfloat16>), grid, dim3(threads), 0, st,
        dx_out, reinterpret_cast<const __nv_bfloat16*>(y), gate, gate_stride, reinterpret_cast<__nv_bfloat16*>(dy),
        dgate, dgate_stride, dbias, T, D);
  else
    DITB_KLAUNCH((gate_resid_bwd_kernel<float>), grid, dim3(threads), 0, st, dx_out, reinterpret_cast<const float*>(y), gate, gate_stride,
                                                          reinterpret_cast<float*>(dy), dgate, dgate_stride, dbias, T, D);
  DITB_LAUNCH_CHECK("gate_resid_bwd");
  return 0;
}

extern "C" int ditb200_colsum(const void* in, int dtype, float* out, int accumulate, int R, int C, void* stream) {
  DITB_REQUIRE(in && out, DITB200_EINVAL, "colsum: null pointer");
  DITB_REQUIRE(R > 0 && C > 0 && C % 4 == 0, DITB200_EINVAL, "colsum: bad shape R=%d C=%d (C %% 4 == 0)", R, C);
  DITB_REQUIRE(aligned16(in) && aligned16(out), DITB200_EALIGN, "colsum: misaligned input / output");
  cudaStream_t st = (cudaStream_t)stream;
  if (!accumulate) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)C * sizeof(float), st);
    if (e != cudaSuccess) return check_cuda(e, "colsum memset");
  }
  if (C % 8 == 0) {
    dim3 grid((C + kCsCols - 1) / kCsCols, (R + kCsRows - 1) / kCsRows);
    if (dtype == DITB200_BF16)
      DITB_KLAUNCH((colsum_kernel<__nv_bfloat16>), grid, dim3(32 * kCsWarps), 0, st, reinterpret_cast<const __nv_bfloat16*>(in), out, R, C);
    else
      DITB_KLAUNCH((colsum_kernel<float>), grid, dim3(32 * kCsWarps), 0, st, reinterpret_cast<const float*>(in), out, R, C);
  } else {
    dim3 grid((C / 4 + 255) / 256, (R + kCs4Rows - 1) / kCs4Rows);
    if (dtype == DITB200_BF16)
      colsum4_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(in), out, R, C);
    else
      colsum4_kernel<float><<<grid, 256, 0, st>>>(reinterpret_cast<const float*>(in), out, R, C);
  }
  DITB_LAUNCH_CHECK("colsum");
  return 0;
}

extern "C" int ditb200_adaln_wgrad(const float* dmod, int dmod_stride, const void* sc, float* dw, float* dbias, int N,
                                   int R, int D, void* stream) {
  DITB_REQUIRE(dmod && sc && dw, DITB200_EINVAL, "adaln_wgrad: null pointer");
  DITB_REQUIRE(N > 0 && R > 0 && D > 0 && R % kAwRows == 0 && D % kAwCols == 0 && R / kAwRows <= 65535 * 32, DITB200_EINVAL,
               "adaln_wgrad: bad shape N=%d R=%d D=%d (R %% 64 == 0, D %% 128 == 0)", N, R, D);
  DITB_REQUIRE(aligned16(dmod) && dmod_stride % 4 == 0 && aligned16(sc) && aligned16(dw) && (!dbias || aligned16(dbias)),
               DITB200_EALIGN, "adaln_wgrad: misaligned pointer or stride");
  dim3 grid(R / kAwRows, D / kAwCols);
  DITB_KLAUNCH((adaln_wgrad_kernel), grid, dim3(256), 0, (cudaStream_t)stream, dmod, dmod_stride,
               reinterpret_cast<const __nv_bfloat16*>(sc), dw, dbias, N, R, D);
  DITB_LAUNCH_CHECK("adaln_wgrad");
  return 0;
}

extern "C" int ditb200_label_embed_bwd(const float* dc, const int64_t* y, float* dtable, int B, int D, int num_rows,
                                       void* stream) {
  DITB_REQUIRE(dc && y && dtable && B > 0 && D > 0 && num_rows > 0, DITB200_EINVAL, "label_embed_bwd: bad argument");
  const int n = B * D;
  label_embed_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(dc, y, dtable, B, D, num_rows);
  DITB_LAUNCH_CHECK("label_embed_bwd");
  return 0;
}

extern "C" int ditb200_patchify(const float* x, void* patches, int out_dtype, int B, int C, int H, int W, int p,
                                void* stream) {
  DITB_REQUIRE(x && patches && B > 0 && C > 0 && p > 0 && H % p == 0 && W % p == 0, DITB200_EINVAL,
               "patchify: bad argument");
  DITB_REQUIRE(out_dtype == DITB200_BF16 || out_dtype == DITB200_F32, DITB200_EINVAL, "patchify: bad out_dtype %d", out_dtype);
  const size_t n = (size_t)B * C * H * W;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  if (out_dtype == DITB200_BF16)
    patchify_kernel<__nv_bfloat16><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<__nv_bfloat16*>(patches), B, C, H, W, p);
  else
    patchify_kernel<float><<<blocks, 256, 0, (cudaStream_t)stream>>>(x, reinterpret_cast<float*>(patches), B, C, H, W, p);
  DITB_LAUNCH_CHECK("patchify");
  return 0;
}

extern "C" int ditb200_unpatchify_bwd(const float* dout, void* dz, int B, int Cout, int Hp, int p, void* stream) {
  DITB_REQUIRE(dout && dz && B > 0 && Cout > 0 && Hp > 0 && p > 0, DITB200_EINVAL, "unpatchify_bwd: bad argument");
  const size_t n = (size_t)B * Cout * Hp * p * Hp * p;
  const int blocks = (int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096);
  unpatchify_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dout, reinterpret_cast<__nv_bfloat16*>(dz), B, Cout, Hp, p);
  DITB_LAUNCH_CHECK("unpatchify_bwd");
  return 0;
}

extern "C" int ditb200_silu_bwd(const float* dact, const float* pre, float* out, int accumulate, size_t n,
                                void* stream) {
  DITB_REQUIRE(dact && pre && out && n > 0, DITB200_EINVAL, "silu_bwd: bad argument");
  const int blocks = (int)((n + 255) / 256 < 2048 ? (n + 255) / 256 : 2048);
  silu_bwd_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(dact, pre, out, accumulate, n);
  DITB_LAUNCH_CHECK("silu_bwd");
  return 0;
}
