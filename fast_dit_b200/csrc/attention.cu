// attention.cu — fused multi-head attention forward, softmax(q kᵀ/√hd)·v, non-causal, read
// straight from the token-major qkv matrix the QKV GEMM writes and written straight into the
// token-major matrix the out-projection reads: no permute/reshape copies exist on this path.
//
//   bf16 engine : flash-style, 64 queries x 64 keys per step, mma.sync.m16n8k16 bf16 with f32
//                 accumulation, online softmax in f32 (exp2), K/V double-buffered with cp.async.
//                 hd = 72 (DiT-XL) is zero-padded to 80 in shared memory only.
//   f32 engine  : CUDA-core kernel for the fp32 check mode.
#include <stdlib.h>

#include "common.cuh"

namespace ditb200 {

// ------------------------------------------------------------------ small PTX helpers
__device__ __forceinline__ void cp_async16(void* dst, const void* src, bool valid) {
  const uint32_t d = smem_u32(dst);
  const int sz = valid ? 16 : 0;  // src-size 0 => zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3)
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// =============================================================== bf16 flash forward
constexpr int kBQ = 64, kBKV = 64;

template <int HD>
struct AttnCfg {
  static constexpr int HDP = (HD + 15) / 16 * 16;  // padded head dim (K of QKᵀ, N of PV)
  static constexpr int LDS = HDP + 8;              // smem row stride (elements): conflict-free ldmatrix
  static constexpr int KSTEPS = HDP / 16;
  static constexpr int NT_O = HDP / 8;
  static constexpr int CHUNKS = HD / 8;            // 16-byte chunks of real data per row
  static constexpr int TILE_ELEMS = 64 * LDS;
  static constexpr int SMEM_BYTES = 5 * TILE_ELEMS * 2;  // Q + 2xK + 2xV
};

template <int HD>
__device__ __forceinline__ void load_tile_async(__nv_bfloat16* dst, const __nv_bfloat16* src, int row_stride,
                                                int rows_valid) {
  using C = AttnCfg<HD>;
  // 64 rows x CHUNKS 16-byte chunks, 128 threads
  for (int idx = threadIdx.x; idx < 64 * C::CHUNKS; idx += 128) {
    const int r = idx / C::CHUNKS, c = idx - r * C::CHUNKS;
    const bool ok = r < rows_valid;
    cp_async16(dst + r * C::LDS + c * 8, src + (size_t)(ok ? r : 0) * row_stride + c * 8, ok);
  }
}

template <int HD>
__global__ void __launch_bounds__(128) attn_fwd_bf16_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                            __nv_bfloat16* __restrict__ out,
                                                            float* __restrict__ lse, int T, int H,
                                                            float scale_log2e) {
  using C = AttnCfg<HD>;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sK = sQ + C::TILE_ELEMS;      // 2 stages
  __nv_bfloat16* sV = sK + 2 * C::TILE_ELEMS;  // 2 stages

  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * HD, ld = 3 * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = qb * kBQ;
  const __nv_bfloat16* base = qkv + (size_t)b * T * ld + h * HD;
  const __nv_bfloat16* gQ = base + (size_t)q0 * ld;
  const __nv_bfloat16* gK = base + D;
  const __nv_bfloat16* gV = base + 2 * D;

  // zero the padding columns [HD, HDP) of every buffer once (cp.async never touches them)
  if constexpr (C::HDP > HD) {
    for (int idx = threadIdx.x; idx < 5 * 64; idx += 128) {
      __nv_bfloat16* rowp = sQ + (size_t)idx * C::LDS + HD;
      *reinterpret_cast<uint4*>(rowp) = make_uint4(0, 0, 0, 0);  // HDP - HD == 8 elements
    }
  }
  static_assert(C::HDP - HD == 0 || C::HDP - HD == 8, "padding must be one 16-byte chunk");

  const int n_kv = (T + kBKV - 1) / kBKV;
  load_tile_async<HD>(sQ, gQ, ld, min(kBQ, T - q0));
  load_tile_async<HD>(sK, gK, ld, min(kBKV, T));
  load_tile_async<HD>(sV, gV, ld, min(kBKV, T));
  cp_async_commit();

  uint32_t qf[C::KSTEPS][4];
  float o[C::NT_O][4];
#pragma unroll
  for (int i = 0; i < C::NT_O; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float m_run[2] = {-INFINITY, -INFINITY}, l_run[2] = {0.f, 0.f};

  for (int j = 0; j < n_kv; ++j) {
    const int st = j & 1;
    if (j + 1 < n_kv) {
      const int k1 = (j + 1) * kBKV;
      load_tile_async<HD>(sK + (st ^ 1) * C::TILE_ELEMS, gK + (size_t)k1 * ld, ld, min(kBKV, T - k1));
      load_tile_async<HD>(sV + (st ^ 1) * C::TILE_ELEMS, gV + (size_t)k1 * ld, ld, min(kBKV, T - k1));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    if (j == 0) {
      // Q fragments (A operand, 16 rows of this warp) stay in registers for the whole loop
      const int r = warp * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int cbase = 8 * (lane >> 4);
#pragma unroll
      for (int ks = 0; ks < C::KSTEPS; ++ks)
        ldsm_x4(smem_u32(sQ + r * C::LDS + ks * 16 + cbase), qf[ks][0], qf[ks][1], qf[ks][2], qf[ks][3]);
    }
    const __nv_bfloat16* tK = sK + st * C::TILE_ELEMS;
    const __nv_bfloat16* tV = sV + st * C::TILE_ELEMS;

    // ---- S = Q Kᵀ  (16 x 64 per warp)
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < C::KSTEPS; ++ks) {
#pragma unroll
      for (int np = 0; np < 4; ++np) {  // pairs of 8-key tiles
        uint32_t b0, b1, b2, b3;
        const int kr = np * 16 + (lane & 7) + 8 * (lane >> 4);
        const int kc = ks * 16 + 8 * ((lane >> 3) & 1);
        ldsm_x4(smem_u32(tK + kr * C::LDS + kc), b0, b1, b2, b3);
        mma_bf16_16816(s[2 * np], qf[ks], b0, b1);
        mma_bf16_16816(s[2 * np + 1], qf[ks], b2, b3);
      }
    }
    // ---- mask keys beyond T (last tile only)
    const int kbase = j * kBKV;
    if (kbase + kBKV > T) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int key = kbase + i * 8 + 2 * t4;
        if (key >= T) s[i][0] = -INFINITY, s[i][2] = -INFINITY;
        if (key + 1 >= T) s[i][1] = -INFINITY, s[i][3] = -INFINITY;
      }
    }
    // ---- online softmax (rows g and g+8 of this warp's 16)
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mx[0] = fmaxf(mx[0], fmaxf(s[i][0], s[i][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[i][2], s[i][3]));
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
    }
    float alpha[2], msc[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const float m_new = fmaxf(m_run[r], mx[r]);
      alpha[r] = exp2f((m_run[r] - m_new) * scale_log2e);  // first tile: exp2(-inf) = 0
      m_run[r] = m_new;
      msc[r] = m_new * scale_log2e;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];  // P as bf16 A-fragments: 4 k-steps of 16 keys
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float p0 = exp2f(s[i][0] * scale_log2e - msc[0]);
      const float p1 = exp2f(s[i][1] * scale_log2e - msc[0]);
      const float p2 = exp2f(s[i][2] * scale_log2e - msc[1]);
      const float p3 = exp2f(s[i][3] * scale_log2e - msc[1]);
      rs[0] += p0 + p1;
      rs[1] += p2 + p3;
      pf[i >> 1][(i & 1) * 2 + 0] = pack_bf16x2(p0, p1);
      pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(p2, p3);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) l_run[r] = l_run[r] * alpha[r] + rs[r];
#pragma unroll
    for (int i = 0; i < C::NT_O; ++i) {
      o[i][0] *= alpha[0], o[i][1] *= alpha[0];
      o[i][2] *= alpha[1], o[i][3] *= alpha[1];
    }
    // ---- O += P V
#pragma unroll
    for (int ks = 0; ks < 4; ++ks) {  // 16 keys per step
#pragma unroll
      for (int dp = 0; dp < C::NT_O / 2; ++dp) {  // pairs of 8-wide d tiles
        uint32_t b0, b1, b2, b3;
        const int vr = ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
        const int vc = dp * 16 + 8 * (lane >> 4);
        ldsm_x4_t(smem_u32(tV + vr * C::LDS + vc), b0, b1, b2, b3);
        mma_bf16_16816(o[2 * dp], pf[ks], b0, b1);
        mma_bf16_16816(o[2 * dp + 1], pf[ks], b2, b3);
      }
    }
    __syncthreads();  // everyone done with stage st before it is refilled
  }

  // ---- finalise: O /= l, stage through this warp's rows of sQ, then coalesced 16-byte stores
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 1);
    l_run[r] += __shfl_xor_sync(0xffffffffu, l_run[r], 2);
  }
  const float inv0 = 1.0f / l_run[0], inv1 = 1.0f / l_run[1];
  __nv_bfloat16* sO = sQ + (size_t)warp * 16 * C::LDS;
#pragma unroll
  for (int i = 0; i < C::NT_O; ++i) {
    const int col = i * 8 + 2 * t4;
    *reinterpret_cast<uint32_t*>(sO + g * C::LDS + col) = pack_bf16x2(o[i][0] * inv0, o[i][1] * inv0);
    *reinterpret_cast<uint32_t*>(sO + (g + 8) * C::LDS + col) = pack_bf16x2(o[i][2] * inv1, o[i][3] * inv1);
  }
  __syncwarp();
  __nv_bfloat16* gO = out + ((size_t)b * T + q0 + warp * 16) * D + h * HD;
  for (int idx = lane; idx < 16 * C::CHUNKS; idx += 32) {
    const int r = idx / C::CHUNKS, c = idx - r * C::CHUNKS;
    if (q0 + warp * 16 + r < T)
      *reinterpret_cast<uint4*>(gO + (size_t)r * D + c * 8) = *reinterpret_cast<const uint4*>(sO + r * C::LDS + c * 8);
  }
  if (lse != nullptr && t4 == 0) {
    const int qr = q0 + warp * 16 + g;
    const float ln2 = 0.6931471805599453f;
    float* L = lse + ((size_t)b * H + h) * T;
    if (qr < T) L[qr] = (m_run[0] * scale_log2e + log2f(l_run[0])) * ln2;
    if (qr + 8 < T) L[qr + 8] = (m_run[1] * scale_log2e + log2f(l_run[1])) * ln2;
  }
}

// ================================================================ bf16 flash backward
// Two passes, both re-computing P = exp(S - lse) tile by tile from q, k and the forward's lse, so
// nothing of size T x T ever reaches HBM:
//   dkv kernel : CTA = 64 keys of one (b, h); loops over query tiles; works on the TRANSPOSED
//                score tile (keys = rows) so that P^T and dS^T come out of the MMA already in
//                A-fragment layout:  S^T = K Q^T,  dV += P^T dO,  dP^T = V dO^T,
//                dS^T = P^T o (dP^T - dsum[q]),  dK += dS^T Q.
//   dq kernel  : CTA = 64 queries; loops over key tiles:  S = Q K^T, dP = dO V^T,
//                dS = P o (dP - dsum[q]),  dQ += dS K.
// Splitting avoids atomics on dQ at the price of computing S and dP twice.
// dsum[q] = sum_d dO[q,d] O[q,d] comes from attn_bwd_dsum_kernel.

// acc[16 x 64] += A[16 rows of sA starting at arow][0..HDP) . (64 rows of tB)[0..HDP)^T
template <int HD>
__device__ __forceinline__ void mma_rows_x_tileT(float (&acc)[8][4], const __nv_bfloat16* sA, int arow,
                                                 const __nv_bfloat16* tB, int lane) {
  using C = AttnCfg<HD>;
  const int r = arow + (lane & 7) + 8 * ((lane >> 3) & 1);
  const int cbase = 8 * (lane >> 4);
#pragma unroll
  for (int ks = 0; ks < C::KSTEPS; ++ks) {
    uint32_t a[4];
    ldsm_x4(smem_u32(sA + r * C::LDS + ks * 16 + cbase), a[0], a[1], a[2], a[3]);
#pragma unroll
    for (int np = 0; np < 4; ++np) {
      uint32_t b0, b1, b2, b3;
      const int kr = np * 16 + (lane & 7) + 8 * (lane >> 4);
      const int kc = ks * 16 + 8 * ((lane >> 3) & 1);
      ldsm_x4(smem_u32(tB + kr * C::LDS + kc), b0, b1, b2, b3);
      mma_bf16_16816(acc[2 * np], a, b0, b1);
      mma_bf16_16816(acc[2 * np + 1], a, b2, b3);
    }
  }
}
// acc[16 x HDP] += P[16 x 64] (A fragments in registers) . tB[64 rows][0..HDP)
template <int HD>
__device__ __forceinline__ void mma_frag_x_tile(float (&acc)[AttnCfg<HD>::NT_O][4], const uint32_t (&pf)[4][4],
                                                const __nv_bfloat16* tB, int lane) {
  using C = AttnCfg<HD>;
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) {
#pragma unroll
    for (int dp = 0; dp < C::NT_O / 2; ++dp) {
      uint32_t b0, b1, b2, b3;
      const int vr = ks * 16 + (lane & 7) + 8 * ((lane >> 3) & 1);
      const int vc = dp * 16 + 8 * (lane >> 4);
      ldsm_x4_t(smem_u32(tB + vr * C::LDS + vc), b0, b1, b2, b3);
      mma_bf16_16816(acc[2 * dp], pf[ks], b0, b1);
      mma_bf16_16816(acc[2 * dp + 1], pf[ks], b2, b3);
    }
  }
}
__device__ __forceinline__ void frags_from_acc(uint32_t (&pf)[4][4], const float (&s)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    pf[i >> 1][(i & 1) * 2 + 0] = pack_bf16x2(s[i][0], s[i][1]);
    pf[i >> 1][(i & 1) * 2 + 1] = pack_bf16x2(s[i][2], s[i][3]);
  }
}
// 16 x HD accumulator of this warp -> token-major global rows through a shared-memory staging tile
template <int HD>
__device__ __forceinline__ void store_acc_rows(const float (&acc)[AttnCfg<HD>::NT_O][4], float mul,
                                               __nv_bfloat16* sStage /* this warp's 16 rows */, __nv_bfloat16* gdst,
                                               int ld, int rows_valid, int lane) {
  using C = AttnCfg<HD>;
  const int g = lane >> 2, t4 = lane & 3;
#pragma unroll
  for (int i = 0; i < C::NT_O; ++i) {
    const int col = i * 8 + 2 * t4;
    *reinterpret_cast<uint32_t*>(sStage + g * C::LDS + col) = pack_bf16x2(acc[i][0] * mul, acc[i][1] * mul);
    *reinterpret_cast<uint32_t*>(sStage + (g + 8) * C::LDS + col) = pack_bf16x2(acc[i][2] * mul, acc[i][3] * mul);
  }
  __syncwarp();
  for (int idx = lane; idx < 16 * C::CHUNKS; idx += 32) {
    const int r = idx / C::CHUNKS, c = idx - r * C::CHUNKS;
    if (r < rows_valid)
      *reinterpret_cast<uint4*>(gdst + (size_t)r * ld + c * 8) = *reinterpret_cast<const uint4*>(sStage + r * C::LDS + c * 8);
  }
}

template <int HD>
__global__ void attn_bwd_dsum_kernel(const __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ dout,
                                     float* __restrict__ dsum, int B, int T, int H) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // (b, t, h), h fastest: coalesced over heads
  if (idx >= B * T * H) return;
  const int h = idx % H, bt = idx / H;
  const int b = bt / T, t = bt - b * T;
  const __nv_bfloat16* o = out + (size_t)bt * H * HD + h * HD;
  const __nv_bfloat16* d = dout + (size_t)bt * H * HD + h * HD;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < HD / 8; ++c) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + c * 8), g = *reinterpret_cast<const uint4*>(d + c * 8);
    const __nv_bfloat162* a2 = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&g);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float2 x = __bfloat1622float2(a2[q]), y = __bfloat1622float2(g2[q]);
      s = fmaf(x.x, y.x, s);
      s = fmaf(x.y, y.y, s);
    }
  }
  dsum[((size_t)b * H + h) * T + t] = s;
}

template <int HD>
__global__ void __launch_bounds__(128) attn_bwd_dkv_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                           const __nv_bfloat16* __restrict__ dout,
                                                           const float* __restrict__ lse, const float* __restrict__ dsum,
                                                           __nv_bfloat16* __restrict__ dqkv, int T, int H,
                                                           float scale, float scale_log2e) {
  using C = AttnCfg<HD>;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sK = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sV = sK + C::TILE_ELEMS;
  __nv_bfloat16* sQ = sV + C::TILE_ELEMS;       // 2 stages
  __nv_bfloat16* sdO = sQ + 2 * C::TILE_ELEMS;  // 2 stages
  const int kb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * HD, ld = 3 * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t4 = lane & 3;
  const int k0 = kb * kBKV;
  const __nv_bfloat16* base = qkv + (size_t)b * T * ld + h * HD;
  const __nv_bfloat16* gdO = dout + (size_t)b * T * D + h * HD;
  const float* L = lse + ((size_t)b * H + h) * T;
  const float* Ds = dsum + ((size_t)b * H + h) * T;

  if constexpr (C::HDP > HD) {
    for (int idx = threadIdx.x; idx < 6 * 64; idx += 128)
      *reinterpret_cast<uint4*>(sK + (size_t)idx * C::LDS + HD) = make_uint4(0, 0, 0, 0);
  }
  load_tile_async<HD>(sK, base + D + (size_t)k0 * ld, ld, min(kBKV, T - k0));
  load_tile_async<HD>(sV, base + 2 * D + (size_t)k0 * ld, ld, min(kBKV, T - k0));
  load_tile_async<HD>(sQ, base, ld, min(kBQ, T));
  load_tile_async<HD>(sdO, gdO, D, min(kBQ, T));
  cp_async_commit();

  float dk[C::NT_O][4], dv[C::NT_O][4];
#pragma unroll
  for (int i = 0; i < C::NT_O; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) dk[i][e] = 0.f, dv[i][e] = 0.f;
  const float log2e = 1.4426950408889634f;

  const int n_q = (T + kBQ - 1) / kBQ;
  for (int j = 0; j < n_q; ++j) {
    const int st = j & 1;
    if (j + 1 < n_q) {
      const int q1 = (j + 1) * kBQ;
      load_tile_async<HD>(sQ + (st ^ 1) * C::TILE_ELEMS, base + (size_t)q1 * ld, ld, min(kBQ, T - q1));
      load_tile_async<HD>(sdO + (st ^ 1) * C::TILE_ELEMS, gdO + (size_t)q1 * D, D, min(kBQ, T - q1));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* tQ = sQ + st * C::TILE_ELEMS;
    const __nv_bfloat16* tdO = sdO + st * C::TILE_ELEMS;
    const int q0 = j * kBQ;
    // per-column (= query) statistics of this thread's fragment columns
    float lq[8][2], dq_[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const int q = q0 + i * 8 + 2 * t4 + e;
        const bool ok = q < T;
        lq[i][e] = ok ? __ldg(L + q) * log2e : INFINITY;  // exp2(x - inf) = 0 masks queries beyond T
        dq_[i][e] = ok ? __ldg(Ds + q) : 0.f;
      }
    }
    float s[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
    mma_rows_x_tileT<HD>(s, sK, warp * 16, tQ, lane);  // S^T
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = exp2f(s[i][0] * scale_log2e - lq[i][0]);
      s[i][1] = exp2f(s[i][1] * scale_log2e - lq[i][1]);
      s[i][2] = exp2f(s[i][2] * scale_log2e - lq[i][0]);
      s[i][3] = exp2f(s[i][3] * scale_log2e - lq[i][1]);
    }
    uint32_t pf[4][4];
    frags_from_acc(pf, s);
    mma_frag_x_tile<HD>(dv, pf, tdO, lane);  // dV += P^T dO
    float dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    mma_rows_x_tileT<HD>(dp, sV, warp * 16, tdO, lane);  // dP^T
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] *= dp[i][0] - dq_[i][0];
      s[i][1] *= dp[i][1] - dq_[i][1];
      s[i][2] *= dp[i][2] - dq_[i][0];
      s[i][3] *= dp[i][3] - dq_[i][1];
    }
    frags_from_acc(pf, s);
    mma_frag_x_tile<HD>(dk, pf, tQ, lane);  // dK += dS^T Q
    __syncthreads();
  }
  // dK, dV rows of this warp -> dqkv[:, D + h*HD ...] and [:, 2D + h*HD ...]
  __nv_bfloat16* stage = sQ + (size_t)warp * 16 * C::LDS;
  const int rows_valid = min(16, T - (k0 + warp * 16));
  __nv_bfloat16* gdst = dqkv + ((size_t)b * T + k0 + warp * 16) * ld + h * HD;
  store_acc_rows<HD>(dk, scale, stage, gdst + D, ld, rows_valid, lane);
  __syncwarp();
  store_acc_rows<HD>(dv, 1.0f, stage, gdst + 2 * D, ld, rows_valid, lane);
}

template <int HD>
__global__ void __launch_bounds__(128) attn_bwd_dq_kernel(const __nv_bfloat16* __restrict__ qkv,
                                                          const __nv_bfloat16* __restrict__ dout,
                                                          const float* __restrict__ lse, const float* __restrict__ dsum,
                                                          __nv_bfloat16* __restrict__ dqkv, int T, int H,
                                                          float scale, float scale_log2e) {
  using C = AttnCfg<HD>;
  extern __shared__ __align__(16) uint8_t smem_attn[];
  __nv_bfloat16* sQ = reinterpret_cast<__nv_bfloat16*>(smem_attn);
  __nv_bfloat16* sdO = sQ + C::TILE_ELEMS;
  __nv_bfloat16* sK = sdO + C::TILE_ELEMS;     // 2 stages
  __nv_bfloat16* sV = sK + 2 * C::TILE_ELEMS;  // 2 stages
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * HD, ld = 3 * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int q0 = qb * kBQ;
  const __nv_bfloat16* base = qkv + (size_t)b * T * ld + h * HD;
  const __nv_bfloat16* gK = base + D;
  const __nv_bfloat16* gV = base + 2 * D;
  const __nv_bfloat16* gdO = dout + ((size_t)b * T + q0) * D + h * HD;

  if constexpr (C::HDP > HD) {
    for (int idx = threadIdx.x; idx < 6 * 64; idx += 128)
      *reinterpret_cast<uint4*>(sQ + (size_t)idx * C::LDS + HD) = make_uint4(0, 0, 0, 0);
  }
  load_tile_async<HD>(sQ, base + (size_t)q0 * ld, ld, min(kBQ, T - q0));
  load_tile_async<HD>(sdO, gdO, D, min(kBQ, T - q0));
  load_tile_async<HD>(sK, gK, ld, min(kBKV, T));
  load_tile_async<HD>(sV, gV, ld, min(kBKV, T));
  cp_async_commit();

  const float log2e = 1.4426950408889634f;
  float lrow[2], drow[2];
#pragma unroll
  for (int e = 0; e < 2; ++e) {
    const int q = q0 + warp * 16 + g + 8 * e;
    const bool ok = q < T;
    lrow[e] = ok ? lse[((size_t)b * H + h) * T + q] * log2e : INFINITY;
    drow[e] = ok ? dsum[((size_t)b * H + h) * T + q] : 0.f;
  }
  float dq[C::NT_O][4];
#pragma unroll
  for (int i = 0; i < C::NT_O; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;

  const int n_kv = (T + kBKV - 1) / kBKV;
  for (int j = 0; j < n_kv; ++j) {
    const int st = j & 1;
    if (j + 1 < n_kv) {
      const int k1 = (j + 1) * kBKV;
      load_tile_async<HD>(sK + (st ^ 1) * C::TILE_ELEMS, gK + (size_t)k1 * ld, ld, min(kBKV, T - k1));
      load_tile_async<HD>(sV + (st ^ 1) * C::TILE_ELEMS, gV + (size_t)k1 * ld, ld, min(kBKV, T - k1));
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const __nv_bfloat16* tK = sK + st * C::TILE_ELEMS;
    const __nv_bfloat16* tV = sV + st * C::TILE_ELEMS;
    float s[8][4], dp[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
    mma_rows_x_tileT<HD>(s, sQ, warp * 16, tK, lane);    // S
    mma_rows_x_tileT<HD>(dp, sdO, warp * 16, tV, lane);  // dP
    const int kbase = j * kBKV;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int key = kbase + i * 8 + 2 * t4;
      const float m0 = (key < T) ? 1.f : 0.f, m1 = (key + 1 < T) ? 1.f : 0.f;
      s[i][0] = m0 * exp2f(s[i][0] * scale_log2e - lrow[0]) * (dp[i][0] - drow[0]);
      s[i][1] = m1 * exp2f(s[i][1] * scale_log2e - lrow[0]) * (dp[i][1] - drow[0]);
      s[i][2] = m0 * exp2f(s[i][2] * scale_log2e - lrow[1]) * (dp[i][2] - drow[1]);
      s[i][3] = m1 * exp2f(s[i][3] * scale_log2e - lrow[1]) * (dp[i][3] - drow[1]);
    }
    uint32_t pf[4][4];
    frags_from_acc(pf, s);
    mma_frag_x_tile<HD>(dq, pf, tK, lane);  // dQ += dS K
    __syncthreads();
  }
  __nv_bfloat16* stage = sK + (size_t)warp * 16 * C::LDS;
  const int rows_valid = min(16, T - (q0 + warp * 16));
  __nv_bfloat16* gdst = dqkv + ((size_t)b * T + q0 + warp * 16) * ld + h * HD;
  store_acc_rows<HD>(dq, scale, stage, gdst, ld, rows_valid, lane);
}

// ================================================================= f32 check-mode forward
// CTA = 16 queries of one (b, h); warp w owns queries 4w..4w+3.  Keys/values are staged in
// shared memory 32 at a time; lane j scores key j, lanes own output dims lane, lane+32, ...
constexpr int kFQ = 16, kFKV = 32;
__global__ void __launch_bounds__(128) attn_fwd_f32_kernel(const float* __restrict__ qkv,
                                                           float* __restrict__ out,
                                                           float* __restrict__ lse, int T, int H,
                                                           int HD, float scale) {
  extern __shared__ float smem_f[];
  const int LDF = HD + 1;
  float* sQ = smem_f;                // [16][LDF]
  float* sK = sQ + kFQ * LDF;        // [32][LDF]
  float* sV = sK + kFKV * LDF;       // [32][LDF]
  const int qb = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
  const int D = H * HD, ld = 3 * D;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q0 = qb * kFQ;
  const float* base = qkv + (size_t)b * T * ld + h * HD;
  for (int idx = threadIdx.x; idx < kFQ * HD; idx += 128) {
    const int r = idx / HD, c = idx - r * HD;
    sQ[r * LDF + c] = (q0 + r < T) ? base[(size_t)(q0 + r) * ld + c] : 0.f;
  }
  constexpr int kMaxDPerLane = 4;  // HD <= 128
  float acc[4][kMaxDPerLane];
  float m_run[4], l_run[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    m_run[r] = -INFINITY, l_run[r] = 0.f;
#pragma unroll
    for (int d = 0; d < kMaxDPerLane; ++d) acc[r][d] = 0.f;
  }
  for (int k0 = 0; k0 < T; k0 += kFKV) {
    __syncthreads();
    for (int idx = threadIdx.x; idx < kFKV * HD; idx += 128) {
      const int r = idx / HD, c = idx - r * HD;
      const bool ok = k0 + r < T;
      sK[r * LDF + c] = ok ? base[(size_t)(k0 + r) * ld + D + c] : 0.f;
      sV[r * LDF + c] = ok ? base[(size_t)(k0 + r) * ld + 2 * D + c] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float* qrow = sQ + (warp * 4 + r) * LDF;
      const float* krow = sK + lane * LDF;
      float sc = 0.f;
      for (int c = 0; c < HD; ++c) sc = fmaf(qrow[c], krow[c], sc);
      sc = (k0 + lane < T) ? sc * scale : -INFINITY;
      const float m_new = fmaxf(m_run[r], warp_max(sc));
      const float alpha = expf(m_run[r] - m_new);
      const float p = expf(sc - m_new);
      l_run[r] = l_run[r] * alpha + warp_sum(p);
      m_run[r] = m_new;
#pragma unroll
      for (int d = 0; d < kMaxDPerLane; ++d) acc[r][d] *= alpha;
      for (int jk = 0; jk < kFKV; ++jk) {
        const float pj = __shfl_sync(0xffffffffu, p, jk);
        const float* vrow = sV + jk * LDF;
#pragma unroll
        for (int d = 0; d < kMaxDPerLane; ++d) {
          const int c = lane + 32 * d;
          if (c < HD) acc[r][d] = fmaf(pj, vrow[c], acc[r][d]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int q = q0 + warp * 4 + r;
    if (q >= T) continue;
    const float inv = 1.0f / l_run[r];
#pragma unroll
    for (int d = 0; d < kMaxDPerLane; ++d) {
      const int c = lane + 32 * d;
      if (c < HD) out[((size_t)b * T + q) * D + h * HD + c] = acc[r][d] * inv;
    }
    if (lse != nullptr && lane == 0) lse[((size_t)b * H + h) * T + q] = m_run[r] + logf(l_run[r]);
  }
}

}  // namespace ditb200

using namespace ditb200;

template <int HD>
static int launch_attn_bf16(const void* qkv, void* out, float* lse, int B, int T, int H, cudaStream_t st) {
  using C = AttnCfg<HD>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_bf16_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         C::SMEM_BYTES);
    if (e != cudaSuccess) return check_cuda(e, "attention smem attribute");
    attr_set = true;
  }
  const float scale_log2e = (float)(1.4426950408889634 / sqrt((double)HD));
  dim3 grid((T + kBQ - 1) / kBQ, H, B);
  attn_fwd_bf16_kernel<HD><<<grid, 128, C::SMEM_BYTES, st>>>(reinterpret_cast<const __nv_bfloat16*>(qkv),
                                                             reinterpret_cast<__nv_bfloat16*>(out), lse, T, H,
                                                             scale_log2e);
  DITB_LAUNCH_CHECK("attention_fwd(bf16)");
  return 0;
}

namespace ditb200 {
// attention_tc.cu: tcgen05 / TMEM forward for T in {128, 256}, head dim 64..80
bool attn_fwd_tc_supported(int T, int hd);
int launch_attn_fwd_tc(const void* qkv, void* out, float* lse, int B, int T, int H, int hd, int reverse, cudaStream_t st);
bool attn_bwd_tc_supported(int T, int hd);
int launch_attn_bwd_tc(const void* qkv, const void* dout, const float* lse, const float* dsum, void* dqkv, int B, int T,
                       int H, int hd, cudaStream_t st);
}  // namespace ditb200

// Test hook (host only): which bf16 kernel family serves a (tokens per image, head dim) pair: 2 = tcgen05 with K/V
// streamed in 128-key blocks (forward only), 1 = tcgen05 whole-row kernels, 0 = mma.sync flash kernels, -1 = rejected.
extern "C" int ditb200_debug_attention_path(int T, int hd, int backward) {
  if (T <= 0 || (hd != 64 && hd != 72 && !(hd == 80 && !backward && ditb200::attn_fwd_tc_supported(T, hd)))) return -1;
  if (backward) return ditb200::attn_bwd_tc_supported(T, hd) ? 1 : 0;
  if (!ditb200::attn_fwd_tc_supported(T, hd)) return 0;
  return T > 256 ? 2 : 1;
}

extern "C" int ditb200_attention_fwd(const void* qkv, void* out, float* lse, int dtype, int B, int T, int H,
                                     int hd, int reverse, void* stream) {
  DITB_REQUIRE(qkv && out, DITB200_EINVAL, "attention_fwd: null pointer");
  DITB_REQUIRE(B > 0 && T > 0 && H > 0 && hd > 0, DITB200_EINVAL, "attention_fwd: bad shape");
  DITB_REQUIRE(B <= 65535 && H <= 65535, DITB200_EINVAL, "attention_fwd: B, H must fit a grid dimension");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DITB200_BF16) {
    DITB_REQUIRE(aligned16(qkv) && aligned16(out), DITB200_EALIGN, "attention_fwd: misaligned pointer");
    static const bool legacy = getenv("DITB200_ATTN_MMA_SYNC") != nullptr;  // measurement switch
    if (!legacy && attn_fwd_tc_supported(T, hd)) return launch_attn_fwd_tc(qkv, out, lse, B, T, H, hd, reverse, st);
    if (hd == 64) return launch_attn_bf16<64>(qkv, out, lse, B, T, H, st);
    if (hd == 72) return launch_attn_bf16<72>(qkv, out, lse, B, T, H, st);
    set_error("attention_fwd(bf16): head dim %d not supported (64, 72)", hd);
    return DITB200_EINVAL;
  }
  DITB_REQUIRE(dtype == DITB200_F32, DITB200_EINVAL, "attention_fwd: bad dtype %d", dtype);
  DITB_REQUIRE(hd <= 128, DITB200_EINVAL, "attention_fwd(f32): head dim %d > 128", hd);
  const size_t smem = (size_t)(kFQ + 2 * kFKV) * (hd + 1) * sizeof(float);
  dim3 grid((T + kFQ - 1) / kFQ, H, B);
  attn_fwd_f32_kernel<<<grid, 128, smem, st>>>(reinterpret_cast<const float*>(qkv), reinterpret_cast<float*>(out),
                                               lse, T, H, hd, (float)(1.0 / sqrt((double)hd)));
  DITB_LAUNCH_CHECK("attention_fwd(f32)");
  return 0;
}

template <int HD>
static int launch_attn_bwd_bf16(const void* qkv, const void* out, const void* dout, const float* lse, float* dsum,
                                void* dqkv, int B, int T, int H, cudaStream_t st) {
  using C = AttnCfg<HD>;
  constexpr int kSmem = 6 * C::TILE_ELEMS * 2;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_dkv_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(attn_bwd_dq_kernel<HD>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (e != cudaSuccess) return check_cuda(e, "attention_bwd smem attribute");
    attr_set = true;
  }
  const float scale = (float)(1.0 / sqrt((double)HD));
  const float scale_log2e = (float)(1.4426950408889634 / sqrt((double)HD));
  const __nv_bfloat16* q = reinterpret_cast<const __nv_bfloat16*>(qkv);
  const __nv_bfloat16* d_o = reinterpret_cast<const __nv_bfloat16*>(dout);
  __nv_bfloat16* dq = reinterpret_cast<__nv_bfloat16*>(dqkv);
  const int n = B * T * H;
  attn_bwd_dsum_kernel<HD><<<(n + 255) / 256, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(out), d_o, dsum, B, T, H);
  DITB_LAUNCH_CHECK("attention_bwd(dsum)");
  static const bool legacy = getenv("DITB200_ATTN_MMA_SYNC") != nullptr;  // measurement switch
  if (!legacy && attn_bwd_tc_supported(T, HD)) return launch_attn_bwd_tc(qkv, dout, lse, dsum, dqkv, B, T, H, HD, st);
  dim3 grid((T + 63) / 64, H, B);
  attn_bwd_dkv_kernel<HD><<<grid, 128, kSmem, st>>>(q, d_o, lse, dsum, dq, T, H, scale, scale_log2e);
  DITB_LAUNCH_CHECK("attention_bwd(dkv)");
  attn_bwd_dq_kernel<HD><<<grid, 128, kSmem, st>>>(q, d_o, lse, dsum, dq, T, H, scale, scale_log2e);
  DITB_LAUNCH_CHECK("attention_bwd(dq)");
  return 0;
}

extern "C" int ditb200_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                                     float* dsum, void* dqkv, int dtype, int B, int T, int H, int hd, void* stream) {
  DITB_REQUIRE(qkv && out && dout && lse && dsum && dqkv, DITB200_EINVAL, "attention_bwd: null pointer");
  DITB_REQUIRE(B > 0 && T > 0 && H > 0 && hd > 0 && B <= 65535 && H <= 65535, DITB200_EINVAL, "attention_bwd: bad shape");
  DITB_REQUIRE(dtype == DITB200_BF16, DITB200_EINVAL, "attention_bwd: bf16 only (the fp32 check mode is forward-only)");
  DITB_REQUIRE(aligned16(qkv) && aligned16(out) && aligned16(dout) && aligned16(dqkv), DITB200_EALIGN,
               "attention_bwd: misaligned pointer");
  cudaStream_t st = (cudaStream_t)stream;
  if (hd == 64) return launch_attn_bwd_bf16<64>(qkv, out, dout, lse, dsum, dqkv, B, T, H, st);
  if (hd == 72) return launch_attn_bwd_bf16<72>(qkv, out, dout, lse, dsum, dqkv, B, T, H, st);
  set_error("attention_bwd(bf16): head dim %d not supported (64, 72)", hd);
  return DITB200_EINVAL;
}
