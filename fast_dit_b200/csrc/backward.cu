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
