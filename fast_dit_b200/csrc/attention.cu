// attention.cu — fused multi-head attention forward, softmax(q kᵀ/√hd)·v, non-causal, read
// straight from the token-major qkv matrix the QKV GEMM writes and written straight into the
// token-major matrix the out-projection reads: no permute/reshape copies exist on this path.
//
//   bf16 engine : flash-style, 64 queries x 64 keys per step, mma.sync.m16n8k16 bf16 with f32
//                 accumulation, online softmax in f32 (exp2), K/V double-buffered with cp.async.
//                 hd = 72 (DiT-XL) is zero-padded to 80 in shared memory only.
//   f32 engine  : CUDA-core kernel for the fp32 check mode.
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

extern "C" int ditb200_attention_fwd(const void* qkv, void* out, float* lse, int dtype, int B, int T, int H,
                                     int hd, void* stream) {
  DITB_REQUIRE(qkv && out, DITB200_EINVAL, "attention_fwd: null pointer");
  DITB_REQUIRE(B > 0 && T > 0 && H > 0 && hd > 0, DITB200_EINVAL, "attention_fwd: bad shape");
  DITB_REQUIRE(B <= 65535 && H <= 65535, DITB200_EINVAL, "attention_fwd: B, H must fit a grid dimension");
  cudaStream_t st = (cudaStream_t)stream;
  if (dtype == DITB200_BF16) {
    DITB_REQUIRE(aligned16(qkv) && aligned16(out), DITB200_EALIGN, "attention_fwd: misaligned pointer");
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
