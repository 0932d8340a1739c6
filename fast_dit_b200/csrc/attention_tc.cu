// attention_tc.cu — fused multi-head attention forward on the 5th-generation tensor cores, for the DiT
// sequence lengths that fit one score tile: T = 128 or 256 tokens (256 px / patch 2), head dim 64..80.
//
//   out = softmax(q kᵀ / sqrt(hd)) v     per (image b, head h), non-causal, read straight from the token-major
//   qkv matrix of the QKV GEMM and written into the token-major matrix the out-projection reads.
//
// With T <= 256 the whole score row of a query fits tensor memory (128 lanes x 256 f32 columns), so there is
// no online-softmax rescaling: S = Q Kᵀ is produced by tcgen05.mma into TMEM, each softmax thread owns one
// query row (two passes over its TMEM row: max, then exp/sum), P is written back to TMEM as packed bf16 over the
// columns S occupied, and O = P V is a second tcgen05.mma whose A operand is read from TMEM.
//
// One persistent CTA per SM loops over (b, h) work items:
//   warps 0-3  softmax + output of query tile 0 (rows 0..127 of the head)     } ping-pong: while one group
//   warps 4-7  softmax + output of query tile 1 (rows 128..255, if T = 256)   } exponentiates, the other's MMAs run
//   warp 8     TMA producer: Q tiles, K and V of the NEXT item into a 2-stage ring while this one computes
//   warp 9     MMA issuer (one elected lane) + TMEM allocation
// hd = 72 (DiT-XL) is handled without padding copies: the TMA tensor map is 3-D {hd, 3H heads, tokens}, the first
// 64 channels land as a 128-byte-swizzled tile and channels 64..79 as a second, 32-byte-swizzled tile whose
// columns >= hd are zero-filled by TMA (out of bounds in dimension 0); each gets its own tcgen05.mma.
#include <stdlib.h>

#include "common.cuh"

namespace ditb200 {

constexpr int kAtQ = 128;               // queries per tile = TMEM lanes
constexpr int kAtThreads = 320;         // 8 softmax warps + producer + MMA
constexpr int kAtRegion = 256;          // TMEM columns per query tile: S (<=256 f32) / P (<=128) + O (80)
constexpr int kAtOCol = 128;            // O accumulator offset inside the region (P occupies [0,128))

struct AttnSmem {
  static constexpr int kQ0 = kAtQ * 128;        // Q channels 0..63, 128-byte rows
  static constexpr int kQ1 = kAtQ * 32;         // Q channels 64..79, 32-byte rows
  static constexpr int kK0 = 256 * 128, kK1 = 256 * 32, kV0 = 256 * 128, kV1 = 256 * 32;
  static constexpr int kKV = kK0 + kK1 + kV0 + kV1;  // one ring stage (sized for T = 256)
  static constexpr int oQ0 = 0, oQ1 = oQ0 + 2 * kQ0, oKV = oQ1 + 2 * kQ1;
  static constexpr int kStg = 32 * 64;           // per softmax warp: 32 rows x 32 bf16 output staging
  static constexpr int oStg = oKV + 2 * kKV;
  static constexpr int oBars = oStg + 8 * kStg;
  static constexpr int kBytes = oBars + 256 + 1024;
};

// 3-D tile load {c0 = channel, c1 = head slot, c2 = token}
__device__ __forceinline__ void tma_load_3d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc]  and  D[tmem] (+)= A[tmem] * B[smem desc], with the shared-memory
// descriptors passed as (low word, high word): the issue loops advance the low word (start address >> 4) by a
// compile-time constant per k step, so that the one issuing thread spends one add per instruction instead of
// rebuilding the descriptor (attn_fwd_tc_kernel: its P.V product is 32 small MMAs per query tile in a loop whose trip
// count is a run-time value, and was bound by their issue, not by the tensor pipe: profiles/r02_attention_timeline.md)
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr, uint32_t lbo_units) {
  return ((smem_addr & 0x3FFFFu) >> 4) | (lbo_units << 16);
}
__device__ __forceinline__ void umma_bf16_ss_lh(uint32_t tmem_d, uint32_t a_lo, uint32_t b_lo, uint32_t hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 da, {%1, %3};\n\tmov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(a_lo), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_ts_lh(uint32_t tmem_d, uint32_t tmem_a, uint32_t b_lo, uint32_t hi,
                                                uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t.reg .b64 db;\n\tsetp.ne.b32 p, %5, 0;\n\t"
      "mov.b64 db, {%2, %3};\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %4, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "r"(b_lo), "r"(hi), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the same forms with whole 64-bit descriptors (KV-blocked forward and backward: their issue loops have compile-time
// trip counts and ptxas folds the descriptor arithmetic; the low/high split there only costs registers)
__device__ __forceinline__ void umma_bf16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tmem_st_32x16(uint32_t taddr, const uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// two accumulator registers as a packed f32x2 operand (FFMA2 / FADD2 / FMUL2 halve the issue slots of the
// per-element softmax arithmetic; the kernels below are issue- and MUFU-bound, not FMA-bound)
__device__ __forceinline__ float2 f2(uint32_t a, uint32_t b) { return make_float2(__uint_as_float(a), __uint_as_float(b)); }

// shared-memory matrix descriptors (hi word: SBO | version 1 | layout type; lo word: address >> 4 | LBO << 16)
constexpr uint32_t kDescHiSw128 = (1024u >> 4) | (1u << 14) | (2u << 29);  // 8-row groups every 1024 B
constexpr uint32_t kDescHiSw32 = (256u >> 4) | (1u << 14) | (6u << 29);    // 8-row groups every 256 B
__device__ __forceinline__ uint64_t mk_desc(uint32_t hi, uint32_t smem_addr, uint32_t lbo_units) {
  return ((uint64_t)hi << 32) | (uint64_t)(((smem_addr & 0x3FFFFu) >> 4) | (lbo_units << 16));
}

// Timeline probe (tools/attn_trace.py; -DDITB200_ATTN_TRACE builds only, the default library carries none of it):
// CTA 0 and CTA 100 record clock64() at every hand-off of their first 8 work items.
#ifdef DITB200_ATTN_TRACE
__device__ unsigned long long g_attn_trace[2 * 8 * 32];
#define ATR(slot)                                                                                          \
  do {                                                                                                     \
    if ((blockIdx.x == 0 || blockIdx.x == 100) && it < 8)                                                  \
      g_attn_trace[((blockIdx.x != 0) * 8 + it) * 32 + (slot)] = (unsigned long long)clock64();            \
  } while (0)
#else
#define ATR(slot) do { } while (0)
#endif

__global__ void __launch_bounds__(kAtThreads, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_k0,
                   const __grid_constant__ CUtensorMap map_v0, const __grid_constant__ CUtensorMap map_q1,
                   const __grid_constant__ CUtensorMap map_kv1, __nv_bfloat16* __restrict__ out,
                   float* __restrict__ lse, const int B, const int T, const int H, const int hd,
                   const float scale_log2e, const int reverse) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + AttnSmem::oBars);
  uint64_t* q_full = bars;          // [2]  TMA -> MMA      (per query tile)
  uint64_t* q_empty = bars + 2;     // [2]  MMA -> TMA
  uint64_t* kv_full = bars + 4;     // [2]  TMA -> MMA      (ring stage)
  uint64_t* kv_empty = bars + 6;    // [2]  MMA -> TMA
  uint64_t* s_full = bars + 8;      // [2]  MMA -> softmax  (S tile in TMEM)
  uint64_t* p_full = bars + 10;     // [2]  softmax -> MMA  (P written, 128 arrivals)
  uint64_t* o_full = bars + 12;     // [2]  MMA -> softmax  (O tile in TMEM)
  uint64_t* s_free = bars + 14;     // [2]  softmax -> MMA  (region drained, 128 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = T / kAtQ;          // query tiles per head: 1 or 2
  const bool has_c1 = hd > 64;       // second channel chunk (64..79)
  const int D = H * hd;
  const int n_items = B * H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q0), tma_prefetch_desc(&map_k0), tma_prefetch_desc(&map_v0);
    tma_prefetch_desc(&map_q1), tma_prefetch_desc(&map_kv1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1), mbar_init(&q_empty[i], 1), mbar_init(&kv_full[i], 1), mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1), mbar_init(&p_full[i], 128), mbar_init(&o_full[i], 1), mbar_init(&s_free[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<1>(tmem_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  DITB_PDL_TRIGGER();  // PDL build variant only (common.cuh)
  DITB_PDL_WAIT();

  if (warp == 8) {
    // ================================================================== TMA producer
    const uint32_t q_bytes = (uint32_t)(AttnSmem::kQ0 + (has_c1 ? AttnSmem::kQ1 : 0));
    const uint32_t kv_bytes = (uint32_t)(2 * T * 128 + (has_c1 ? 2 * T * 32 : 0));
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int wi = reverse ? n_items - 1 - w : w;  // last-first: start on what the QKV GEMM wrote last (L2)
      const int b = wi / H, h = wi - b * H;
      const int stage = it & 1;
      const uint32_t kvpar = (it >> 1) & 1, par = it & 1;
      const int tok0 = b * T;
      uint8_t* kv = smem + AttnSmem::oKV + stage * AttnSmem::kKV;
      mbar_wait(&kv_empty[stage], kvpar ^ 1u);
      if (lane == 0) ATR(20);
      if (elect_one()) {
        mbar_arrive_expect_tx(&kv_full[stage], kv_bytes);
        tma_load_3d(&map_k0, &kv_full[stage], kv, 0, H + h, tok0);
        if (has_c1) tma_load_3d(&map_kv1, &kv_full[stage], kv + AttnSmem::kK0, 64, H + h, tok0);
      }
      __syncwarp();
      for (int t = 0; t < nqt; ++t) {
        mbar_wait(&q_empty[t], par ^ 1u);
        if (elect_one()) {
          mbar_arrive_expect_tx(&q_full[t], q_bytes);
          tma_load_3d(&map_q0, &q_full[t], smem + AttnSmem::oQ0 + t * AttnSmem::kQ0, 0, h, tok0 + t * kAtQ);
          if (has_c1) tma_load_3d(&map_q1, &q_full[t], smem + AttnSmem::oQ1 + t * AttnSmem::kQ1, 64, h, tok0 + t * kAtQ);
        }
        __syncwarp();
      }
      if (elect_one()) {
        uint8_t* v0 = kv + AttnSmem::kK0 + AttnSmem::kK1;
        for (int kb = 0; kb < T / 64; ++kb)  // V channels 0..63: {64 channels x 64 keys} boxes, 8 KB each
          tma_load_3d(&map_v0, &kv_full[stage], v0 + kb * 8192, 0, 2 * H + h, tok0 + kb * 64);
        if (has_c1) tma_load_3d(&map_kv1, &kv_full[stage], v0 + AttnSmem::kV0, 64, 2 * H + h, tok0);
      }
      if (lane == 0) ATR(21);
      __syncwarp();
    }
  } else if (warp == 9) {
    // ==================================================================== MMA issuer
    const uint32_t idesc_s = umma_idesc_bf16(kAtQ, (uint32_t)T);                 // S: 128 x T, both K-major
    const uint32_t idesc_o64 = umma_idesc_bf16(kAtQ, 64) | (1u << 16);           // O[:, 0:64]: B = V, MN-major
    const uint32_t idesc_o16 = umma_idesc_bf16(kAtQ, 16) | (1u << 16);           // O[:, 64:80]
    // Issue order (per query tile t, items i = 0, 1, ...):  S_t(0) ... then  O_t(i), S_t(i+1)  alternating over t.
    // Putting S_t(i+1) directly behind O_t(i) lets softmax group t start on its next tile while the other group
    // is still exponentiating: the MUFU pipe, which bounds this kernel, never waits for the tensor pipe.
    auto issue_s = [&](int t, int stage) {
      const uint32_t kv = smem_u32(smem + AttnSmem::oKV + stage * AttnSmem::kKV);
      const uint32_t k0 = kv, k1 = kv + AttnSmem::kK0;
      const uint32_t d = tmem_base + (uint32_t)(t * kAtRegion);
      const uint32_t q0 = smem_u32(smem + AttnSmem::oQ0 + t * AttnSmem::kQ0);
      const uint32_t q1 = smem_u32(smem + AttnSmem::oQ1 + t * AttnSmem::kQ1);
      const uint32_t q0lo = desc_lo(q0, 0), k0lo = desc_lo(k0, 0);
#pragma unroll
      for (int j = 0; j < 4; ++j)  // 16 channels per step = 32 bytes inside the 128-byte swizzled row
        umma_bf16_ss_lh(d, q0lo + 2 * j, k0lo + 2 * j, kDescHiSw128, idesc_s, j > 0);
      if (has_c1) umma_bf16_ss_lh(d, desc_lo(q1, 1), desc_lo(k1, 1), kDescHiSw32, idesc_s, 1u);
      umma_commit<1>(&s_full[t]);
      umma_commit<1>(&q_empty[t]);
    };
    auto issue_o = [&](int t, int stage, bool last) {
      const uint32_t kv = smem_u32(smem + AttnSmem::oKV + stage * AttnSmem::kKV);
      const uint32_t v0 = kv + AttnSmem::kK0 + AttnSmem::kK1, v1 = v0 + AttnSmem::kV0;
      const uint32_t p = tmem_base + (uint32_t)(t * kAtRegion);
      const uint32_t d = p + kAtOCol;
      // 16 keys per step: 8 packed TMEM columns of P, 2 KB / 512 B of V = 128 / 32 descriptor units
      const uint32_t v0lo = desc_lo(v0, 0), v1lo = desc_lo(v1, 1);
      if (has_c1) {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          umma_bf16_ts_lh(d, p + 8 * ks, v0lo + 128 * ks, kDescHiSw128, idesc_o64, ks > 0);
          umma_bf16_ts_lh(d + 64, p + 8 * ks, v1lo + 32 * ks, kDescHiSw32, idesc_o16, ks > 0);
        }
        if (T == 256) {
#pragma unroll
          for (int ks = 8; ks < 16; ++ks) {
            umma_bf16_ts_lh(d, p + 8 * ks, v0lo + 128 * ks, kDescHiSw128, idesc_o64, 1u);
            umma_bf16_ts_lh(d + 64, p + 8 * ks, v1lo + 32 * ks, kDescHiSw32, idesc_o16, 1u);
          }
        }
      } else {
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_bf16_ts_lh(d, p + 8 * ks, v0lo + 128 * ks, kDescHiSw128, idesc_o64, ks > 0);
        if (T == 256) {
#pragma unroll
          for (int ks = 8; ks < 16; ++ks) umma_bf16_ts_lh(d, p + 8 * ks, v0lo + 128 * ks, kDescHiSw128, idesc_o64, 1u);
        }
      }
      umma_commit<1>(&o_full[t]);
      if (last) umma_commit<1>(&kv_empty[stage]);
    };
    if ((int)blockIdx.x < n_items) {  // prologue: the score tiles of this CTA's first item
      mbar_wait(&kv_full[0], 0u);
      for (int t = 0; t < nqt; ++t) {
        mbar_wait(&q_full[t], 0u);
        tcgen05_fence_after();
        if (elect_one()) issue_s(t, 0);
        __syncwarp();
      }
    }
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int stage = it & 1;
      const uint32_t par = it & 1;
      const bool has_next = w + (int)gridDim.x < n_items;
      const int nstage = (it + 1) & 1;
      const uint32_t nkvpar = ((it + 1) >> 1) & 1, npar = (it + 1) & 1;
      for (int t = 0; t < nqt; ++t) {
        mbar_wait(&p_full[t], par);
        if (lane == 0) ATR(12 + 4 * t);
        tcgen05_fence_after();
        if (elect_one()) issue_o(t, stage, t == nqt - 1);
        __syncwarp();
        if (lane == 0) ATR(13 + 4 * t);
        if (has_next) {
          if (t == 0) mbar_wait(&kv_full[nstage], nkvpar);
          mbar_wait(&q_full[t], npar);
          mbar_wait(&s_free[t], par);  // group t has pulled O_t(i) out of the region
          if (lane == 0) ATR(14 + 4 * t);
          tcgen05_fence_after();
          if (elect_one()) issue_s(t, nstage);
          __syncwarp();
          if (lane == 0) ATR(15 + 4 * t);
        }
      }
    }
  } else {
    // =================================================== softmax + output, one thread per query row
    const int t = warp >> 2, quarter = warp & 3;
    if (t < nqt) {
      const int row = quarter * 32 + lane;
      const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * kAtRegion);
      const int nch = T / 32;  // 32-key chunks of the score row
      int it = 0;
      for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
        const int wi = reverse ? n_items - 1 - w : w;
        const int b = wi / H, h = wi - b * H;
        const uint32_t par = it & 1;
        mbar_wait(&s_full[t], par);
        if ((threadIdx.x & 127) == 0) ATR(6 * t);
        tcgen05_fence_after();
        // ---- pass 1: row maximum of the raw scores (TMEM loads one chunk ahead of the compare chain)
        uint32_t va[32], vb[32];
        float mx = -INFINITY;
        tmem_ld_32x32(trow, va);
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          tmem_ld_32x32(trow + 32 * (c + 1), vb);
#pragma unroll
          for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(va[j]), __uint_as_float(va[j + 1])));
          tmem_ld_wait();
          tmem_ld_32x32(trow + (c + 2 < nch ? 32 * (c + 2) : 0), va);  // last round: chunk 0 again, for pass 2
#pragma unroll
          for (int j = 0; j < 32; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(vb[j]), __uint_as_float(vb[j + 1])));
        }
        // ---- pass 2: p = exp2((s - max) * scale * log2 e), row sum, P -> TMEM as packed bf16 (in place: the
        //      16 columns chunk c of P lands on were read as part of S chunk c/2 <= c)
        const float msc = mx * scale_log2e;
        if ((threadIdx.x & 127) == 0) ATR(6 * t + 1);
        float2 sum2 = make_float2(0.f, 0.f);
        const float2 sl2 = make_float2(scale_log2e, scale_log2e), nm2 = make_float2(-msc, -msc);
        uint32_t pk[16];
        for (int c = 0; c < nch; c += 2) {
          tmem_ld_wait();
          tmem_ld_32x32(trow + 32 * (c + 1), vb);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 e = __ffma2_rn(f2(va[2 * j], va[2 * j + 1]), sl2, nm2);
            const float2 pp = make_float2(ex2_approx(e.x), ex2_approx(e.y));
            sum2 = __fadd2_rn(sum2, pp);
            pk[j] = pack_bf16x2(pp.x, pp.y);
          }
          tmem_st_32x16(trow + 16 * c, pk);
          tmem_ld_wait();
          if (c + 2 < nch) tmem_ld_32x32(trow + 32 * (c + 2), va);
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float2 e = __ffma2_rn(f2(vb[2 * j], vb[2 * j + 1]), sl2, nm2);
            const float2 pp = make_float2(ex2_approx(e.x), ex2_approx(e.y));
            sum2 = __fadd2_rn(sum2, pp);
            pk[j] = pack_bf16x2(pp.x, pp.y);
          }
          tmem_st_32x16(trow + 16 * (c + 1), pk);
        }
        const float sum = sum2.x + sum2.y;
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(&p_full[t]);
        if ((threadIdx.x & 127) == 0) ATR(6 * t + 2);
        // ---- output: O / sum -> bf16, transposed through this warp's staging buffer so that the global stores
        //      cover whole 64-byte row segments (8 rows per instruction) instead of 32 scattered 16-byte pieces
        mbar_wait(&o_full[t], par);
        if ((threadIdx.x & 127) == 0) ATR(6 * t + 3);
        tcgen05_fence_after();
        uint32_t o2[16];
        tmem_ld_32x32(trow + kAtOCol, va);
        tmem_ld_32x32(trow + kAtOCol + 32, vb);
        if (has_c1) tmem_ld_32x16(trow + kAtOCol + 64, o2);
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(&s_free[t]);  // the accumulators are in registers: the region may take the next score tile
        if ((threadIdx.x & 127) == 0) ATR(6 * t + 4);
        const float inv = 1.0f / sum;
        const uint32_t stg = smem_u32(smem + AttnSmem::oStg) + (uint32_t)(warp * AttnSmem::kStg);
        __nv_bfloat16* obase = out + ((size_t)b * T + t * kAtQ + quarter * 32) * D + h * hd;
        auto flush = [&](const uint32_t* v, int ncol, int col0) {  // ncol in {32, 16} accumulator columns
          const uint32_t my = stg + (uint32_t)lane * 64u;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            if (g * 8 < ncol) {
              uint4 q4;
              q4.x = pack_bf16x2(__uint_as_float(v[8 * g]) * inv, __uint_as_float(v[8 * g + 1]) * inv);
              q4.y = pack_bf16x2(__uint_as_float(v[8 * g + 2]) * inv, __uint_as_float(v[8 * g + 3]) * inv);
              q4.z = pack_bf16x2(__uint_as_float(v[8 * g + 4]) * inv, __uint_as_float(v[8 * g + 5]) * inv);
              q4.w = pack_bf16x2(__uint_as_float(v[8 * g + 6]) * inv, __uint_as_float(v[8 * g + 7]) * inv);
              sts128(my + (uint32_t)((g ^ sw) << 4), q4.x, q4.y, q4.z, q4.w);
            }
          }
          __syncwarp();
          const int g = lane & 3;
          const int col = col0 + g * 8;
          if (g * 8 < ncol && col < hd) {
#pragma unroll
            for (int pss = 0; pss < 4; ++pss) {
              const int r = pss * 8 + (lane >> 2);
              const uint4 q4 = lds128_u(stg + (uint32_t)(r * 64 + ((g ^ ((r >> 1) & 3)) << 4)));
              *reinterpret_cast<uint4*>(obase + (size_t)r * D + col) = q4;
            }
          }
          __syncwarp();
        };
        flush(va, 32, 0);
        flush(vb, 32, 32);
        if (has_c1) flush(o2, 16, 64);
        if (lse != nullptr)
          lse[((size_t)b * H + h) * T + t * kAtQ + row] = (msc + log2f(sum)) * 0.6931471805599453f;
        if ((threadIdx.x & 127) == 0) ATR(6 * t + 5);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}
#undef ATR

// ============================================================================ long sequences (T > 256)
// KV-blocked variant for T = 512, 768, 1024, ... (the 512 px configuration: 1024 tokens).  A work item is one
// (image, head, PAIR of 128-query tiles); K and V stream through a 4-stage ring in blocks of 128 keys, each block
// serving both query tiles.  Per query tile the score block S (128 x 128 f32) and the output accumulator O live
// in TMEM side by side (128 + 80 columns), so the softmax is the online form:
//   block max -> m_new;  when some row of the warp outgrew its reference maximum m by more than 2^8:  alpha =
//   exp2((m - m_new) scale), O *= alpha in TMEM, l *= alpha, m = m_new (otherwise m stays: any reference keeps
//   O / l exact as long as the exponentials stay in range);  P = exp2((S - m) scale) -> TMEM over S;
//   l += sum P;  O += P V_j.
// tcgen05.mma instructions of one CTA execute in issue order, which is what orders  P_j's read by PV_j  before
// S_{j+1} overwrites it, and PV_{j-1}'s write of O before the commit that publishes S_j: no extra barriers.
// Issue order: S(0,0) S(1,0) | PV(0,g) S(0,g+1) PV(1,g) S(1,g+1) | ... over the flattened (item, block) sequence g.
struct AttnKvSmem {
  static constexpr int kQ0 = kAtQ * 128, kQ1 = kAtQ * 32;
  static constexpr int kBlk = 128;                                   // keys per KV block
  static constexpr int kK0 = kBlk * 128, kK1 = kBlk * 32, kV0 = kBlk * 128, kV1 = kBlk * 32;
  static constexpr int kKV = kK0 + kK1 + kV0 + kV1;                  // 40 KB per stage
  static constexpr int kStages = 4;
  static constexpr int oQ0 = 0, oQ1 = oQ0 + 2 * kQ0, oKV = oQ1 + 2 * kQ1;
  static constexpr int kStg = 32 * 64;
  static constexpr int oStg = oKV + kStages * kKV;
  static constexpr int oBars = oStg + 8 * kStg;
  static constexpr int kBytes = oBars + 256 + 1024;
};
constexpr int kKvRegion = 256, kKvOCol = 128;  // TMEM: [S/P 128 | O 80 | spare] per query tile

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,"
      "%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};" ::"r"(taddr),
      "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]), "r"(v[8]), "r"(v[9]),
      "r"(v[10]), "r"(v[11]), "r"(v[12]), "r"(v[13]), "r"(v[14]), "r"(v[15]), "r"(v[16]), "r"(v[17]), "r"(v[18]),
      "r"(v[19]), "r"(v[20]), "r"(v[21]), "r"(v[22]), "r"(v[23]), "r"(v[24]), "r"(v[25]), "r"(v[26]), "r"(v[27]),
      "r"(v[28]), "r"(v[29]), "r"(v[30]), "r"(v[31])
      : "memory");
}

__global__ void __launch_bounds__(kAtThreads, 1)
attn_fwd_tc_kv_kernel(const __grid_constant__ CUtensorMap map_q0, const __grid_constant__ CUtensorMap map_v0,
                      const __grid_constant__ CUtensorMap map_q1, __nv_bfloat16* __restrict__ out,
                      float* __restrict__ lse, const int B, const int T, const int H, const int hd,
                      const float scale_log2e) {
  using SM = AttnKvSmem;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::oBars);
  uint64_t* q_full = bars;         // [2]
  uint64_t* q_empty = bars + 2;    // [2]
  uint64_t* kv_full = bars + 4;    // [4]
  uint64_t* kv_empty = bars + 8;   // [4]
  uint64_t* s_full = bars + 12;    // [2]
  uint64_t* p_full = bars + 14;    // [2]  128 arrivals
  uint64_t* o_full = bars + 16;    // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_c1 = hd > 64;
  const int D = H * hd;
  const int nblk = T / SM::kBlk;     // KV blocks per head
  const int pairs = T / (2 * kAtQ);  // query-tile pairs per head
  const int n_items = B * H * pairs;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_q0), tma_prefetch_desc(&map_v0), tma_prefetch_desc(&map_q1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1), mbar_init(&q_empty[i], 1);
      mbar_init(&s_full[i], 1), mbar_init(&p_full[i], 128), mbar_init(&o_full[i], 1);
    }
    for (int i = 0; i < SM::kStages; ++i) mbar_init(&kv_full[i], 1), mbar_init(&kv_empty[i], 1);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<1>(tmem_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  DITB_PDL_TRIGGER();
  DITB_PDL_WAIT();

  if (warp == 8) {
    // ================================================================== TMA producer
    const uint32_t q_bytes = (uint32_t)(SM::kQ0 + (has_c1 ? SM::kQ1 : 0));
    const uint32_t kv_bytes = (uint32_t)(SM::kK0 + SM::kV0 + (has_c1 ? SM::kK1 + SM::kV1 : 0));
    int it = 0, g = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int bh = w / pairs, pr = w - bh * pairs;
      const int b = bh / H, h = bh - b * H;
      const int tok0 = b * T;
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&q_empty[t], (uint32_t)(it & 1) ^ 1u);
        if (elect_one()) {
          const int q0 = tok0 + (2 * pr + t) * kAtQ;
          mbar_arrive_expect_tx(&q_full[t], q_bytes);
          tma_load_3d(&map_q0, &q_full[t], smem + SM::oQ0 + t * SM::kQ0, 0, h, q0);
          if (has_c1) tma_load_3d(&map_q1, &q_full[t], smem + SM::oQ1 + t * SM::kQ1, 64, h, q0);
        }
        __syncwarp();
      }
      for (int j = 0; j < nblk; ++j, ++g) {
        const int stage = g % SM::kStages;
        mbar_wait(&kv_empty[stage], (uint32_t)((g / SM::kStages) & 1) ^ 1u);
        if (elect_one()) {
          uint8_t* kv = smem + SM::oKV + stage * SM::kKV;
          const int k0 = tok0 + j * SM::kBlk;
          mbar_arrive_expect_tx(&kv_full[stage], kv_bytes);
          tma_load_3d(&map_q0, &kv_full[stage], kv, 0, H + h, k0);  // K block: same {64 ch, 128 tokens} box as Q
          if (has_c1) tma_load_3d(&map_q1, &kv_full[stage], kv + SM::kK0, 64, H + h, k0);
          uint8_t* v0 = kv + SM::kK0 + SM::kK1;
          tma_load_3d(&map_v0, &kv_full[stage], v0, 0, 2 * H + h, k0);
          tma_load_3d(&map_v0, &kv_full[stage], v0 + 8192, 0, 2 * H + h, k0 + 64);
          if (has_c1) tma_load_3d(&map_q1, &kv_full[stage], v0 + SM::kV0, 64, 2 * H + h, k0);
        }
        __syncwarp();
      }
    }
  } else if (warp == 9) {
    // ==================================================================== MMA issuer
    const uint32_t idesc_s = umma_idesc_bf16(kAtQ, SM::kBlk);
    const uint32_t idesc_o64 = umma_idesc_bf16(kAtQ, 64) | (1u << 16);
    const uint32_t idesc_o16 = umma_idesc_bf16(kAtQ, 16) | (1u << 16);
    int my_items = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) ++my_items;
    const int G = my_items * nblk;  // (item, block) steps of this CTA
    auto issue_s = [&](int t, int g) {
      const int stage = g % SM::kStages, blk = g % nblk;
      const uint32_t kv = smem_u32(smem + SM::oKV + stage * SM::kKV);
      const uint32_t k0 = kv, k1 = kv + SM::kK0;
      const uint32_t d = tmem_base + (uint32_t)(t * kKvRegion);
      const uint32_t q0 = smem_u32(smem + SM::oQ0 + t * SM::kQ0), q1 = smem_u32(smem + SM::oQ1 + t * SM::kQ1);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        umma_bf16<1>(d, mk_desc(kDescHiSw128, q0 + 32 * j, 0), mk_desc(kDescHiSw128, k0 + 32 * j, 0), idesc_s, j > 0);
      if (has_c1) umma_bf16<1>(d, mk_desc(kDescHiSw32, q1, 1), mk_desc(kDescHiSw32, k1, 1), idesc_s, 1u);
      umma_commit<1>(&s_full[t]);
      if (blk == nblk - 1) umma_commit<1>(&q_empty[t]);  // last score block of the item: Q tile may be refilled
    };
    auto issue_pv = [&](int t, int g) {
      const int stage = g % SM::kStages, blk = g % nblk;
      const uint32_t kv = smem_u32(smem + SM::oKV + stage * SM::kKV);
      const uint32_t v0 = kv + SM::kK0 + SM::kK1, v1 = v0 + SM::kV0;
      const uint32_t p = tmem_base + (uint32_t)(t * kKvRegion);
      const uint32_t d = p + kKvOCol;
#pragma unroll
      for (int ks = 0; ks < SM::kBlk / 16; ++ks) {
        const uint32_t acc = (blk > 0 || ks > 0) ? 1u : 0u;  // first block of an item overwrites O
        umma_bf16_ts(d, p + 8 * ks, mk_desc(kDescHiSw128, v0 + 2048 * ks, 0), idesc_o64, acc);
        if (has_c1) umma_bf16_ts(d + 64, p + 8 * ks, mk_desc(kDescHiSw32, v1 + 512 * ks, 1), idesc_o16, acc);
      }
      if (blk == nblk - 1) umma_commit<1>(&o_full[t]);
      if (t == 1) umma_commit<1>(&kv_empty[stage]);  // both query tiles are through with this K/V block
    };
    auto wait_inputs_s = [&](int t, int g) {  // K block g (and, for an item's first block, its Q tile)
      if (t == 0) mbar_wait(&kv_full[g % SM::kStages], (uint32_t)((g / SM::kStages) & 1));
      if (g % nblk == 0) mbar_wait(&q_full[t], (uint32_t)((g / nblk) & 1));
    };
    if (G > 0) {
      for (int t = 0; t < 2; ++t) {
        wait_inputs_s(t, 0);
        tcgen05_fence_after();
        if (elect_one()) issue_s(t, 0);
        __syncwarp();
      }
    }
    for (int g = 0; g < G; ++g) {
      for (int t = 0; t < 2; ++t) {
        mbar_wait(&p_full[t], (uint32_t)(g & 1));
        tcgen05_fence_after();
        if (elect_one()) issue_pv(t, g);
        __syncwarp();
        if (g + 1 < G) {
          wait_inputs_s(t, g + 1);
          tcgen05_fence_after();
          if (elect_one()) issue_s(t, g + 1);
          __syncwarp();
        }
      }
    }
  } else {
    // =================================================== online softmax + output, one thread per query row
    const int t = warp >> 2, quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t trow = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(t * kKvRegion);
    int it = 0, g = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int bh = w / pairs, pr = w - bh * pairs;
      const int b = bh / H, h = bh - b * H;
      float m_run = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j, ++g) {
        mbar_wait(&s_full[t], (uint32_t)(g & 1));
        tcgen05_fence_after();
        uint32_t va[32], vb[32];
        // ---- block maximum
        float mx = m_run;
        tmem_ld_32x32(trow, va);
        tmem_ld_32x32(trow + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; q += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(va[q]), __uint_as_float(va[q + 1])));
#pragma unroll
        for (int q = 0; q < 32; q += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(vb[q]), __uint_as_float(vb[q + 1])));
        tmem_ld_32x32(trow + 64, va);
        tmem_ld_32x32(trow + 96, vb);
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 32; q += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(va[q]), __uint_as_float(va[q + 1])));
#pragma unroll
        for (int q = 0; q < 32; q += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(vb[q]), __uint_as_float(vb[q + 1])));
        // ---- rescale the running output — lazily: the reference maximum only has to keep exp2((s - m) scale) in
        //      range, so it is moved (and O, l rescaled: a TMEM round trip of the accumulator on the group's chain)
        //      only when some row of this warp outgrew it by more than 2^8; until then P is taken relative to the
        //      stale maximum (values up to 256, same relative precision in bf16 / f32) and O / l stays exact
        if (j == 0) {
          m_run = mx;  // l_run = 0, and the first P.V overwrites O
        } else if (__any_sync(0xffffffffu, (mx - m_run) * scale_log2e > 8.0f)) {
          const float alpha = ex2_approx((m_run - mx) * scale_log2e);
          uint32_t o2[16];
          tmem_ld_32x32(trow + kKvOCol, va);
          tmem_ld_32x32(trow + kKvOCol + 32, vb);
          if (has_c1) tmem_ld_32x16(trow + kKvOCol + 64, o2);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 32; ++q) va[q] = __float_as_uint(__uint_as_float(va[q]) * alpha);
#pragma unroll
          for (int q = 0; q < 32; ++q) vb[q] = __float_as_uint(__uint_as_float(vb[q]) * alpha);
          tmem_st_32x32(trow + kKvOCol, va);
          tmem_st_32x32(trow + kKvOCol + 32, vb);
          if (has_c1) {
#pragma unroll
            for (int q = 0; q < 16; ++q) o2[q] = __float_as_uint(__uint_as_float(o2[q]) * alpha);
            tmem_st_32x16(trow + kKvOCol + 64, o2);
          }
          l_run *= alpha;
          m_run = mx;
        }
        // ---- P = exp2((S - m) scale), row sum, P -> TMEM (bf16, in place over S)
        const float msc = m_run * scale_log2e;
        float2 sum2 = make_float2(0.f, 0.f);
        const float2 sl2 = make_float2(scale_log2e, scale_log2e), nm2 = make_float2(-msc, -msc);
        uint32_t pk[16];
        tmem_ld_32x32(trow, va);
#pragma unroll
        for (int c = 0; c < 4; c += 2) {
          tmem_ld_wait();
          tmem_ld_32x32(trow + 32 * (c + 1), vb);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float2 e = __ffma2_rn(f2(va[2 * q], va[2 * q + 1]), sl2, nm2);
            const float2 pp = make_float2(ex2_approx(e.x), ex2_approx(e.y));
            sum2 = __fadd2_rn(sum2, pp);
            pk[q] = pack_bf16x2(pp.x, pp.y);
          }
          tmem_st_32x16(trow + 16 * c, pk);
          tmem_ld_wait();
          if (c + 2 < 4) tmem_ld_32x32(trow + 32 * (c + 2), va);
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const float2 e = __ffma2_rn(f2(vb[2 * q], vb[2 * q + 1]), sl2, nm2);
            const float2 pp = make_float2(ex2_approx(e.x), ex2_approx(e.y));
            sum2 = __fadd2_rn(sum2, pp);
            pk[q] = pack_bf16x2(pp.x, pp.y);
          }
          tmem_st_32x16(trow + 16 * (c + 1), pk);
        }
        l_run += sum2.x + sum2.y;
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(&p_full[t]);
      }
      // ---- output of this query tile: O / l -> bf16, staged for coalesced stores
      mbar_wait(&o_full[t], (uint32_t)(it & 1));
      tcgen05_fence_after();
      uint32_t va[32], vb[32], o2[16];
      tmem_ld_32x32(trow + kKvOCol, va);
      tmem_ld_32x32(trow + kKvOCol + 32, vb);
      if (has_c1) tmem_ld_32x16(trow + kKvOCol + 64, o2);
      tmem_ld_wait();
      const float inv = 1.0f / l_run;
      const uint32_t stg = smem_u32(smem + SM::oStg) + (uint32_t)(warp * SM::kStg);
      const int q_tile0 = (2 * pr + t) * kAtQ;
      __nv_bfloat16* obase = out + ((size_t)b * T + q_tile0 + quarter * 32) * D + h * hd;
      auto flush = [&](const uint32_t* v, int ncol, int col0) {
        const uint32_t my = stg + (uint32_t)lane * 64u;
        const int sw = (lane >> 1) & 3;
#pragma unroll
        for (int gq = 0; gq < 4; ++gq) {
          if (gq * 8 < ncol)
            sts128(my + (uint32_t)((gq ^ sw) << 4),
                   pack_bf16x2(__uint_as_float(v[8 * gq]) * inv, __uint_as_float(v[8 * gq + 1]) * inv),
                   pack_bf16x2(__uint_as_float(v[8 * gq + 2]) * inv, __uint_as_float(v[8 * gq + 3]) * inv),
                   pack_bf16x2(__uint_as_float(v[8 * gq + 4]) * inv, __uint_as_float(v[8 * gq + 5]) * inv),
                   pack_bf16x2(__uint_as_float(v[8 * gq + 6]) * inv, __uint_as_float(v[8 * gq + 7]) * inv));
        }
        __syncwarp();
        const int gq = lane & 3;
        const int col = col0 + gq * 8;
        if (gq * 8 < ncol && col < hd) {
#pragma unroll
          for (int pss = 0; pss < 4; ++pss) {
            const int r = pss * 8 + (lane >> 2);
            const uint4 q4 = lds128_u(stg + (uint32_t)(r * 64 + ((gq ^ ((r >> 1) & 3)) << 4)));
            *reinterpret_cast<uint4*>(obase + (size_t)r * D + col) = q4;
          }
        }
        __syncwarp();
      };
      flush(va, 32, 0);
      flush(vb, 32, 32);
      if (has_c1) flush(o2, 16, 64);
      if (lse != nullptr)
        lse[((size_t)b * H + h) * T + q_tile0 + row] = (m_run * scale_log2e + log2f(l_run)) * 0.6931471805599453f;
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}

// ================================================================================ backward (T = 256)
// dQ, dK, dV of one (image, head) per work item, on the tensor cores, in four jobs that share the Q, K, V, dO
// tiles held in shared memory (each [256 tokens x hd], loaded once by TMA; the same bytes serve as a K-major
// operand "rows x channels" and as an MN-major operand "channels x tokens", only the descriptors differ):
//   key tile kt (2 jobs):   S^T = K_kt Q^T,  dP^T = V_kt dO^T          (128 keys x 256 queries each, TMEM)
//                           P^T = exp2(S^T s - lse[q]),  dS^T = P^T (dP^T - dsum[q]) scale      -> bf16 in TMEM
//                           dV_kt = P^T dO,  dK_kt = dS^T Q                                     (A operand from TMEM)
//   query tile qt (2 jobs): S = Q_qt K^T,  dP = dO_qt V^T;  dS = P (dP - dsum) scale;  dQ_qt = dS K
// TMEM: columns [0,256) hold S / S^T, [256,512) dP / dP^T.  The 256 score columns of a job are produced and
// consumed as two halves (warps 0-3 take columns [0,128), warps 4-7 [128,256)), so the second half's MMAs run
// while the first half is exponentiated.  P / dS overwrite the first 64 columns of their own half; when both
// halves are done the four 64-column holes left over take the output accumulators (channels 0..63 and 64..79 are
// separate MMAs anyway), then the epilogue drains them and the next job starts.
struct AttnBwdSmem {
  static constexpr int kT = 256;
  static constexpr int kC0 = kT * 128, kC1 = kT * 32;              // channel chunks 0..63 / 64..79 of one tensor
  static constexpr int oQ0 = 0, oQ1 = oQ0 + kC0, oK0 = oQ1 + kC1, oK1 = oK0 + kC0, oV0 = oK1 + kC1, oV1 = oV0 + kC0,
                       oD0 = oV1 + kC1, oD1 = oD0 + kC0;
  static constexpr int oRow = oD1 + kC1;                             // lse*log2e [256], dsum [256]
  static constexpr int kStg = 32 * 64;
  static constexpr int oStg = oRow + 2 * kT * 4;
  static constexpr int oBars = oStg + 8 * kStg;
  static constexpr int kBytes = oBars + 128 + 1024;
};

__global__ void __launch_bounds__(kAtThreads, 1)
attn_bwd_tc_kernel(const __grid_constant__ CUtensorMap map_qkv0, const __grid_constant__ CUtensorMap map_qkv1,
                   const __grid_constant__ CUtensorMap map_do0, const __grid_constant__ CUtensorMap map_do1,
                   const float* __restrict__ lse, const float* __restrict__ dsum, __nv_bfloat16* __restrict__ dqkv,
                   const int B, const int H, const int hd, const float scale, const float scale_log2e) {
  using SM = AttnBwdSmem;
  constexpr int T = SM::kT;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + SM::oBars);
  uint64_t* ld_full = bars;       // TMA -> MMA (item)
  uint64_t* ld_empty = bars + 1;  // MMA -> TMA
  uint64_t* s_full = bars + 2;    // [2] MMA -> elementwise, per column half
  uint64_t* p_full = bars + 4;    // elementwise -> MMA (256 arrivals)
  uint64_t* o_full = bars + 5;    // MMA -> epilogue
  uint64_t* o_free = bars + 6;    // epilogue -> MMA (256 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 7);
  float* row_l2 = reinterpret_cast<float*>(smem + SM::oRow);
  float* row_ds = row_l2 + T;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool has_c1 = hd > 64;
  const int D = H * hd;
  const int n_items = B * H;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_qkv0), tma_prefetch_desc(&map_qkv1), tma_prefetch_desc(&map_do0), tma_prefetch_desc(&map_do1);
    mbar_init(ld_full, 1), mbar_init(ld_empty, 1), mbar_init(&s_full[0], 1), mbar_init(&s_full[1], 1);
    mbar_init(p_full, 256), mbar_init(o_full, 1), mbar_init(o_free, 256);
    fence_barrier_init();
  }
  if (warp == 9) tmem_alloc<1>(tmem_ptr, 512);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  DITB_PDL_TRIGGER();
  DITB_PDL_WAIT();
  const uint32_t RA = tmem_base, RB = tmem_base + 256;  // S / dP regions

  if (warp == 8) {
    // ================================================================== TMA producer
    const uint32_t bytes = (uint32_t)(4 * SM::kC0 + (has_c1 ? 4 * SM::kC1 : 0));
    int it = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int b = w / H, h = w - b * H;
      const int tok0 = b * T;
      mbar_wait(ld_empty, (uint32_t)(it & 1) ^ 1u);
      if (elect_one()) {
        mbar_arrive_expect_tx(ld_full, bytes);
        tma_load_3d(&map_qkv0, ld_full, smem + SM::oK0, 0, H + h, tok0);
        tma_load_3d(&map_qkv0, ld_full, smem + SM::oQ0, 0, h, tok0);
        tma_load_3d(&map_qkv0, ld_full, smem + SM::oV0, 0, 2 * H + h, tok0);
        tma_load_3d(&map_do0, ld_full, smem + SM::oD0, 0, h, tok0);
        if (has_c1) {
          tma_load_3d(&map_qkv1, ld_full, smem + SM::oK1, 64, H + h, tok0);
          tma_load_3d(&map_qkv1, ld_full, smem + SM::oQ1, 64, h, tok0);
          tma_load_3d(&map_qkv1, ld_full, smem + SM::oV1, 64, 2 * H + h, tok0);
          tma_load_3d(&map_do1, ld_full, smem + SM::oD1, 64, h, tok0);
        }
      }
      __syncwarp();
    }
  } else if (warp == 9) {
    // ==================================================================== MMA issuer
    const uint32_t idesc_s = umma_idesc_bf16(kAtQ, 128);                   // scores: 128 x 128 per half, K-major both
    const uint32_t idesc_o64 = umma_idesc_bf16(kAtQ, 64) | (1u << 16);     // outputs: B MN-major
    const uint32_t idesc_o16 = umma_idesc_bf16(kAtQ, 16) | (1u << 16);
    const uint32_t sQ0 = smem_u32(smem + SM::oQ0), sQ1 = smem_u32(smem + SM::oQ1);
    const uint32_t sK0 = smem_u32(smem + SM::oK0), sK1 = smem_u32(smem + SM::oK1);
    const uint32_t sV0 = smem_u32(smem + SM::oV0), sV1 = smem_u32(smem + SM::oV1);
    const uint32_t sD0 = smem_u32(smem + SM::oD0), sD1 = smem_u32(smem + SM::oD1);
    // D[128 x 128] = A_tile(rows at +tile*128) . B_half(rows at +half*128)^T over hd channels
    auto scores = [&](uint32_t d, uint32_t a0, uint32_t a1, int tile, uint32_t b0, uint32_t b1, int half) {
      const uint32_t ao0 = a0 + (uint32_t)tile * 16384u, ao1 = a1 + (uint32_t)tile * 4096u;
      const uint32_t bo0 = b0 + (uint32_t)half * 16384u, bo1 = b1 + (uint32_t)half * 4096u;
#pragma unroll
      for (int j = 0; j < 4; ++j)
        umma_bf16<1>(d, mk_desc(kDescHiSw128, ao0 + 32 * j, 0), mk_desc(kDescHiSw128, bo0 + 32 * j, 0), idesc_s, j > 0);
      if (has_c1) umma_bf16<1>(d, mk_desc(kDescHiSw32, ao1, 1), mk_desc(kDescHiSw32, bo1, 1), idesc_s, 1u);
    };
    // D[128 x hd] = A[tmem, 128 x 256 packed bf16 in two 64-column pieces] . Bmn (channels x 256 tokens)
    auto outputs = [&](uint32_t region, uint32_t b0, uint32_t b1) {
#pragma unroll
      for (int ks = 0; ks < 16; ++ks) {
        const uint32_t a = region + (uint32_t)(ks < 8 ? 8 * ks : 128 + 8 * (ks - 8));
        umma_bf16_ts(region + 64, a, mk_desc(kDescHiSw128, b0 + 2048 * ks, 0), idesc_o64, ks > 0);
        if (has_c1) umma_bf16_ts(region + 192, a, mk_desc(kDescHiSw32, b1 + 512 * ks, 1), idesc_o16, ks > 0);
      }
    };
    int it = 0, jg = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      mbar_wait(ld_full, (uint32_t)(it & 1));
      for (int j = 0; j < 4; ++j, ++jg) {
        const bool kv_pass = j < 2;
        const int tile = j & 1;
        mbar_wait(o_free, (uint32_t)(jg & 1) ^ 1u);  // the previous job's accumulators have been drained
        tcgen05_fence_after();
        for (int half = 0; half < 2; ++half) {
          if (elect_one()) {
            if (kv_pass) {
              scores(RA + 128 * half, sK0, sK1, tile, sQ0, sQ1, half);   // S^T  = K_t Q^T
              scores(RB + 128 * half, sV0, sV1, tile, sD0, sD1, half);   // dP^T = V_t dO^T
            } else {
              scores(RA + 128 * half, sQ0, sQ1, tile, sK0, sK1, half);   // S  = Q_t K^T
              scores(RB + 128 * half, sD0, sD1, tile, sV0, sV1, half);   // dP = dO_t V^T
            }
            umma_commit<1>(&s_full[half]);
          }
          __syncwarp();
        }
        mbar_wait(p_full, (uint32_t)(jg & 1));
        tcgen05_fence_after();
        if (elect_one()) {
          if (kv_pass) {
            outputs(RA, sD0, sD1);  // dV_t = P^T dO
            outputs(RB, sQ0, sQ1);  // dK_t = dS^T Q
          } else {
            outputs(RB, sK0, sK1);  // dQ_t = dS K   (accumulator in the dP region's holes)
          }
          umma_commit<1>(o_full);
          if (j == 3) umma_commit<1>(ld_empty);
        }
        __syncwarp();
      }
    }
  } else {
    // =========================================== elementwise + output: 8 warps, thread = one row of the job
    const int half = warp >> 2, quarter = warp & 3;
    const int r = quarter * 32 + lane;  // row inside the 128-row tile
    const uint32_t lane_off = (uint32_t)(quarter * 32) << 16;
    const uint32_t stg = smem_u32(smem + SM::oStg) + (uint32_t)(warp * SM::kStg);
    int it = 0, jg = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x, ++it) {
      const int b = w / H, h = w - b * H;
      // per-query row terms of this head: lse * log2(e) and dsum
      asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone is done with the previous item's values
      {
        const size_t base = ((size_t)b * H + h) * T + threadIdx.x;
        row_l2[threadIdx.x] = lse[base] * 1.4426950408889634f;
        row_ds[threadIdx.x] = dsum[base] * scale;  // pre-scaled: dS = P * (dP * scale - dsum * scale)
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      for (int j = 0; j < 4; ++j, ++jg) {
        const bool kv_pass = j < 2;
        const int tile = j & 1;
        const uint32_t ta = RA + lane_off + 128u * half, tb = RB + lane_off + 128u * half;
        float my_l2 = 0.f, my_ds = 0.f;
        if (!kv_pass) my_l2 = row_l2[tile * 128 + r], my_ds = row_ds[tile * 128 + r];
        mbar_wait(&s_full[half], (uint32_t)(jg & 1));
        tcgen05_fence_after();
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t vs[32], vp[32];
          tmem_ld_32x32(ta + 32 * c, vs);
          tmem_ld_32x32(tb + 32 * c, vp);
          tmem_ld_wait();
          uint32_t pk[16], dk[16];
          const float2 sl2 = make_float2(scale_log2e, scale_log2e), sc2 = make_float2(scale, scale);
          const int col0 = half * 128 + 32 * c;  // first column (query in the key pass, key in the query pass)
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {  // four score columns at a time
            float4 l = make_float4(my_l2, my_l2, my_l2, my_l2), d = make_float4(my_ds, my_ds, my_ds, my_ds);
            if (kv_pass) {  // column = query: one broadcast 128-bit read each
              l = *reinterpret_cast<const float4*>(row_l2 + col0 + 4 * q4);
              d = *reinterpret_cast<const float4*>(row_ds + col0 + 4 * q4);
            }
            const float2 e01 = __ffma2_rn(f2(vs[4 * q4], vs[4 * q4 + 1]), sl2, make_float2(-l.x, -l.y));
            const float2 e23 = __ffma2_rn(f2(vs[4 * q4 + 2], vs[4 * q4 + 3]), sl2, make_float2(-l.z, -l.w));
            const float p0 = ex2_approx(e01.x), p1 = ex2_approx(e01.y), p2 = ex2_approx(e23.x), p3 = ex2_approx(e23.y);
            const float2 t01 = __ffma2_rn(f2(vp[4 * q4], vp[4 * q4 + 1]), sc2, make_float2(-d.x, -d.y));
            const float2 t23 = __ffma2_rn(f2(vp[4 * q4 + 2], vp[4 * q4 + 3]), sc2, make_float2(-d.z, -d.w));
            const float2 g01 = __fmul2_rn(make_float2(p0, p1), t01), g23 = __fmul2_rn(make_float2(p2, p3), t23);
            const float g0 = g01.x, g1 = g01.y, g2 = g23.x, g3 = g23.y;
            pk[2 * q4] = pack_bf16x2(p0, p1), pk[2 * q4 + 1] = pack_bf16x2(p2, p3);
            dk[2 * q4] = pack_bf16x2(g0, g1), dk[2 * q4 + 1] = pack_bf16x2(g2, g3);
          }
          if (kv_pass) tmem_st_32x16(ta + 16 * c, pk);  // P^T (the query pass does not need P as an operand)
          tmem_st_32x16(tb + 16 * c, dk);               // dS^T / dS
        }
        tmem_st_wait();
        tcgen05_fence_before();
        mbar_arrive(p_full);
        // ---- outputs: the accumulators sit in the 64-column holes [64,128) (channels 0..63) and [192,208) (64..79)
        mbar_wait(o_full, (uint32_t)(jg & 1));
        tcgen05_fence_after();
        // key pass: warps 0-3 take dV (S region), warps 4-7 dK (dP region); query pass: warps 0-3 take dQ channels
        // 0..63, warps 4-7 channels 64..79 (dP region)
        const uint32_t region = (kv_pass ? (half ? RB : RA) : RB) + lane_off;
        uint32_t va[32], vb[32], o2[16];
        const bool lo_part = kv_pass || half == 0, hi_part = has_c1 && (kv_pass || half == 1);
        if (lo_part) {
          tmem_ld_32x32(region + 64, va);
          tmem_ld_32x32(region + 96, vb);
        }
        if (hi_part) tmem_ld_32x16(region + 192, o2);
        tmem_ld_wait();
        tcgen05_fence_before();
        mbar_arrive(o_free);
        const int slot = kv_pass ? (half ? H + h : 2 * H + h) : h;  // dK | dV | dQ column block of dqkv
        __nv_bfloat16* obase = dqkv + ((size_t)b * T + tile * 128 + quarter * 32) * (3 * D) + (size_t)slot * hd;
        auto flush = [&](const uint32_t* v, int ncol, int c0) {
          const uint32_t my = stg + (uint32_t)lane * 64u;
          const int sw = (lane >> 1) & 3;
#pragma unroll
          for (int gq = 0; gq < 4; ++gq) {
            if (gq * 8 < ncol)
              sts128(my + (uint32_t)((gq ^ sw) << 4),
                     pack_bf16x2(__uint_as_float(v[8 * gq]), __uint_as_float(v[8 * gq + 1])),
                     pack_bf16x2(__uint_as_float(v[8 * gq + 2]), __uint_as_float(v[8 * gq + 3])),
                     pack_bf16x2(__uint_as_float(v[8 * gq + 4]), __uint_as_float(v[8 * gq + 5])),
                     pack_bf16x2(__uint_as_float(v[8 * gq + 6]), __uint_as_float(v[8 * gq + 7])));
          }
          __syncwarp();
          const int gq = lane & 3;
          const int col = c0 + gq * 8;
          if (gq * 8 < ncol && col < hd) {
#pragma unroll
            for (int pss = 0; pss < 4; ++pss) {
              const int rr = pss * 8 + (lane >> 2);
              const uint4 q4 = lds128_u(stg + (uint32_t)(rr * 64 + ((gq ^ ((rr >> 1) & 3)) << 4)));
              *reinterpret_cast<uint4*>(obase + (size_t)rr * (3 * D) + col) = q4;
            }
          }
          __syncwarp();
        };
        if (lo_part) {
          flush(va, 32, 0);
          flush(vb, 32, 32);
        }
        if (hi_part) flush(o2, 16, 64);
      }
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc<1>(tmem_base, 512);
}

// 3-D bf16 tensor map over qkv viewed as {hd channels, 3H head slots, B*T tokens}
static int make_tmap_qkv(CUtensorMap* map, const void* base, int hd, int H, uint64_t tokens, uint32_t box_ch,
                         uint32_t box_tok, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    set_error("attention: ditb200_init() has not been called");
    return DITB200_ENOINIT;
  }
  cuuint64_t dims[3] = {(cuuint64_t)hd, (cuuint64_t)(3 * H), tokens};
  cuuint64_t strides[2] = {(cuuint64_t)hd * 2, (cuuint64_t)3 * H * hd * 2};
  cuuint32_t box[3] = {box_ch, 1, box_tok};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention: cuTensorMapEncodeTiled failed with CUresult %d (hd=%d H=%d box=%ux%u)", (int)r, hd, H,
              box_ch, box_tok);
    return DITB200_EINVAL;
  }
  return 0;
}

bool attn_bwd_tc_supported(int T, int hd) { return T == 256 && hd % 8 == 0 && hd >= 64 && hd <= 80; }

// 3-D bf16 tensor map over dout [B*T, H*hd] viewed as {hd channels, H heads, tokens}
static int make_tmap_heads(CUtensorMap* map, const void* base, int hd, int H, uint64_t tokens, uint32_t box_ch,
                           uint32_t box_tok, CUtensorMapSwizzle swz) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    set_error("attention: ditb200_init() has not been called");
    return DITB200_ENOINIT;
  }
  cuuint64_t dims[3] = {(cuuint64_t)hd, (cuuint64_t)H, tokens};
  cuuint64_t strides[2] = {(cuuint64_t)hd * 2, (cuuint64_t)H * hd * 2};
  cuuint32_t box[3] = {box_ch, 1, box_tok};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("attention_bwd: cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return DITB200_EINVAL;
  }
  return 0;
}

// dqkv from qkv, dout, lse and dsum (= rowsum(dout * out), computed by the caller) for T = 256
int launch_attn_bwd_tc(const void* qkv, const void* dout, const float* lse, const float* dsum, void* dqkv, int B, int T,
                       int H, int hd, cudaStream_t st) {
  DITB_REQUIRE(is_initialised(), DITB200_ENOINIT, "attention: ditb200_init() has not been called");
  CUtensorMap mq0, mq1, md0, md1;
  const uint64_t tokens = (uint64_t)B * T;
  int rc;
  if ((rc = make_tmap_qkv(&mq0, qkv, hd, H, tokens, 64, 256, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_qkv(&mq1, qkv, hd, H, tokens, 16, 256, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = make_tmap_heads(&md0, dout, hd, H, tokens, 64, 256, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_heads(&md1, dout, hd, H, tokens, 16, 256, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnBwdSmem::kBytes);
    if (e != cudaSuccess) return check_cuda(e, "attention_bwd(tcgen05) smem attribute");
    attr_set = true;
  }
  int grid = num_sms();
  if (grid > B * H) grid = B * H;
  const float scale = (float)(1.0 / sqrt((double)hd));
  DITB_KLAUNCH(attn_bwd_tc_kernel, dim3(grid), dim3(kAtThreads), AttnBwdSmem::kBytes, st, mq0, mq1, md0, md1, lse, dsum,
                                                                   reinterpret_cast<__nv_bfloat16*>(dqkv), B, H, hd, scale,
                                                                   scale * 1.4426950408889634f);
  DITB_LAUNCH_CHECK("attention_bwd(tcgen05)");
  return 0;
}

bool attn_fwd_tc_supported(int T, int hd) {
  return (T == 128 || T == 256 || (T > 256 && T % 256 == 0)) && hd % 8 == 0 && hd >= 64 && hd <= 80;
}

int launch_attn_fwd_tc(const void* qkv, void* out, float* lse, int B, int T, int H, int hd, int reverse, cudaStream_t st) {
  DITB_REQUIRE(is_initialised(), DITB200_ENOINIT, "attention: ditb200_init() has not been called");
  CUtensorMap mq0, mk0, mv0, mq1, mkv1;
  const uint64_t tokens = (uint64_t)B * T;
  int rc;
  if ((rc = make_tmap_qkv(&mq0, qkv, hd, H, tokens, 64, kAtQ, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  static const bool force_kv = getenv("DITB200_ATTN_KV") != nullptr;  // measurement switch: KV-blocked kernel at T = 256
  const bool use_kv = T > 256 || (force_kv && T == 256);
  const uint32_t kbox = use_kv ? (uint32_t)kAtQ : (uint32_t)T;  // the KV-blocked kernel loads K like Q (128 tokens)
  if ((rc = make_tmap_qkv(&mk0, qkv, hd, H, tokens, 64, kbox, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_qkv(&mv0, qkv, hd, H, tokens, 64, 64, CU_TENSOR_MAP_SWIZZLE_128B))) return rc;
  if ((rc = make_tmap_qkv(&mq1, qkv, hd, H, tokens, 16, kAtQ, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  if ((rc = make_tmap_qkv(&mkv1, qkv, hd, H, tokens, 16, kbox, CU_TENSOR_MAP_SWIZZLE_32B))) return rc;
  const float scale_log2e = (float)(1.4426950408889634 / sqrt((double)hd));
  if (use_kv) {  // KV-blocked online-softmax kernel
    static bool kv_attr_set = false;
    if (!kv_attr_set) {
      cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           AttnKvSmem::kBytes);
      if (e != cudaSuccess) return check_cuda(e, "attention_fwd(tcgen05, kv) smem attribute");
      kv_attr_set = true;
    }
    int items = B * H * (T / 256), g = num_sms();
    if (g > items) g = items;
    DITB_KLAUNCH(attn_fwd_tc_kv_kernel, dim3(g), dim3(kAtThreads), AttnKvSmem::kBytes, st, mq0, mv0, mq1, reinterpret_cast<__nv_bfloat16*>(out),
                                                                    lse, B, T, H, hd, scale_log2e);
    DITB_LAUNCH_CHECK("attention_fwd(tcgen05, kv)");
    return 0;
  }
  int grid = num_sms();
  if (grid > B * H) grid = B * H;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(attn_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnSmem::kBytes);
    if (e != cudaSuccess) return check_cuda(e, "attention_fwd(tcgen05) smem attribute");
    attr_set = true;
  }
  DITB_KLAUNCH(attn_fwd_tc_kernel, grid, kAtThreads, AttnSmem::kBytes, st, mq0, mk0, mv0, mq1, mkv1,
               reinterpret_cast<__nv_bfloat16*>(out), lse, B, T, H, hd, scale_log2e, reverse ? 1 : 0);
  DITB_LAUNCH_CHECK("attention_fwd(tcgen05)");
  return 0;
}

}  // namespace ditb200

#ifdef DITB200_ATTN_TRACE
// probe builds only (not declared in include/ditb200.h): copies the timeline of the last forward launch
extern "C" int ditb200_attn_trace_read(unsigned long long* dst, int n) {
  if (n > 2 * 8 * 32) n = 2 * 8 * 32;
  cudaDeviceSynchronize();
  return (int)cudaMemcpyFromSymbol(dst, ditb200::g_attn_trace, sizeof(unsigned long long) * n);
}
#endif
