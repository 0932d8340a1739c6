// gemm_tc.cu — the tensor-core GEMM of the DiT block, forward and backward:
//   out = epilogue(op(A) · op(W)ᵀ), bf16 operands, f32 accumulation in tensor memory (TMEM),
//   operands streamed by TMA into 128-byte-swizzled shared-memory tiles.
//
// Operand layouts.  The forward GEMMs and the data-gradient's A operand are K-major
// (A[M,K], W[N,K] row-major: the contraction index is contiguous).  The backward pass also needs
//   data gradient    dX[M,Kin]  = dY[M,Nout] · W[Nout,Kin]        -> W is read "[K,N]" (trans_w)
//   weight gradient  dW[Nout,Kin] = dYᵀ · X,  contraction = tokens -> both read "[K,M]/[K,N]"
// Those are MN-major operands for tcgen05.mma: the same TMA engine fetches {64 MN-elements x 64 k}
// boxes out of the row-major matrix and the shared-memory descriptor (LBO = distance between
// 64-element MN chunks, SBO = distance between 8-k groups) plus the a_major/b_major bits of the
// instruction descriptor tell the tensor core to read them transposed.  No transposed copy of an
// activation or a weight is ever written to HBM.
//
// Structure (one persistent CTA per SM, or one CTA pair per two SMs with cta_group::2):
//   warp 0      TMA producer: fills a ring of kStages {A 128x64, B (BN/kCG)x64} bf16 tiles,
//               signalling full[stage] with complete_tx bytes.
//   warp 1      MMA issuer (one elected lane, leader CTA only): tcgen05.mma.kind::f16 128xBNx16 (or
//               256xBNx16 across the pair) into one of two TMEM accumulator stages; tcgen05.commit
//               releases smem slots / publishes accumulators.
//   warps 2..9  epilogue (two warps per TMEM lane quarter, each taking half of the tile's columns):
//               tcgen05.ld the accumulator (thread = output row, 32 columns per load), apply
//               bias / GELU-tanh / adaLN gate + residual / GELU', store to HBM, hand the TMEM stage
//               back.  Runs concurrently with the next tile's main loop.
// Work units are (tile, k-split); tiles are visited n-fastest so the CTAs running at the same time
// share A row-panels in L2.  With split_k > 1 partial tiles are combined with f32 vector atomics.
#include "common.cuh"

namespace ditb200 {

constexpr int kBM = 128;       // rows per CTA tile = TMEM lanes
constexpr int kBK = 64;        // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;     // K per tcgen05.mma for 16-bit inputs
constexpr int kEpiWarps = 8;
constexpr int kNumThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBudget = 200 * 1024;
constexpr int kMnChunkBytes = 64 * kBK * 2;  // one {64 MN x 64 k} TMA box of an MN-major operand

struct EpiParams {
  const float* bias;
  void* out;
  const float* resid;
  const float* gate;
  __nv_bfloat16* aux_out;
  const __nv_bfloat16* aux_in;
  int gate_stride, rows_per_gate;
  int epilogue, out_bf16;
  int atomic;  // out += (f32 vector atomics): split-K partials and gradient accumulation
};

template <int kCG, int BN>
struct TcCfg {
  static constexpr int kBRows = BN / kCG;  // B rows held by each CTA
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "UMMA N / epilogue column split");
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
  static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "swizzle-128B tiles need 1024-B alignment");
};

// d/du gelu_tanh(u)
__device__ __forceinline__ float dgelu_tanh_f(float u) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u2 = u * u;
  const float z = k0 * (u + k1 * u2 * u);
  const float e = __expf(2.0f * z);
  const float t = 1.0f - __fdividef(2.0f, e + 1.0f);  // tanh(z)
  return 0.5f * (1.0f + t) + 0.5f * u * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * u2);
}

template <int kCG, int BN>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const EpiParams ep, const int M, const int N, const int K, const int a_mn, const int b_mn,
               const int split_k) {
  using Cfg = TcCfg<kCG, BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms (same offset in both CTAs of a pair)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::kStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cta_rank = (kCG == 2) ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);    // one arrive.expect_tx (leader CTA) covering every CTA's bytes
      mbar_init(&empty[s], 1);   // one tcgen05.commit per phase
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps * kCG);  // one arrive per epilogue warp of every CTA
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<kCG>(tmem_ptr, Cfg::kTmemCols);
  tcgen05_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  const int tile_m = kBM * kCG;
  const int m_tiles = (M + tile_m - 1) / tile_m;
  const int n_tiles = (N + BN - 1) / BN;
  const int k_blocks = (K + kBK - 1) / kBK;
  const int kb_per = (k_blocks + split_k - 1) / split_k;
  const int num_units = m_tiles * n_tiles * split_k;  // work units = (tile, k-split)
  const int cta_unit = blockIdx.x / kCG;              // CTA (pair) index
  const int num_ctas = gridDim.x / kCG;

  // Producer and MMA roles run warp-uniform loops (every lane waits on the barriers) and elect one
  // lane only around the asynchronous issues.  Keeping the control flow uniform lets the compiler
  // hold descriptors / barrier addresses in uniform registers; a divergent `if (lane == 0)` loop
  // costs a vector->uniform register move in front of every UTCHMMA and made the issue thread the
  // bottleneck (profiles/r01_gemm_notes.md).
  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    for (int u = cta_unit; u < num_units; u += num_ctas) {
      const int tile = u / split_k, split = u - tile * split_k;
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      const int row_a = m_blk * tile_m + (int)cta_rank * kBM;
      const int row_b = n_blk * BN + (int)cta_rank * Cfg::kBRows;
      const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&empty[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* dst_a = smem_a + stage * Cfg::kABytes;
          uint8_t* dst_b = smem_b + stage * Cfg::kBBytes;
          // Both CTAs' bytes complete on the leader's barrier, which only the leader arms.  The
          // peer may run ahead of the arming: its empty[stage] wait guarantees the leader's
          // barrier is already in the matching phase, and a transiently negative tx-count
          // cannot complete the phase while the leader's arrival is still pending.
          if (leader) mbar_arrive_expect_tx(&full[stage], kCG * Cfg::kStageBytes);
          if (!a_mn) {
            if constexpr (kCG == 1) tma_load_2d(&tma_a, &full[stage], dst_a, kb * kBK, row_a);
            else tma_load_2d_pair(&tma_a, &full[stage], dst_a, kb * kBK, row_a);
          } else {
#pragma unroll
            for (int c = 0; c < kBM / 64; ++c) {
              if constexpr (kCG == 1) tma_load_2d(&tma_a, &full[stage], dst_a + c * kMnChunkBytes, row_a + 64 * c, kb * kBK);
              else tma_load_2d_pair(&tma_a, &full[stage], dst_a + c * kMnChunkBytes, row_a + 64 * c, kb * kBK);
            }
          }
          if (!b_mn) {
            if constexpr (kCG == 1) tma_load_2d(&tma_b, &full[stage], dst_b, kb * kBK, row_b);
            else tma_load_2d_pair(&tma_b, &full[stage], dst_b, kb * kBK, row_b);
          } else {
#pragma unroll
            for (int c = 0; c < Cfg::kBRows / 64; ++c) {
              if constexpr (kCG == 1) tma_load_2d(&tma_b, &full[stage], dst_b + c * kMnChunkBytes, row_b + 64 * c, kb * kBK);
              else tma_load_2d_pair(&tma_b, &full[stage], dst_b + c * kMnChunkBytes, row_b + 64 * c, kb * kBK);
            }
          }
        }
        __syncwarp();
        if (++stage == kStages) stage = 0, phase ^= 1u;
      }
    }
  } else if (warp == 1) {
    // ======================================================================= MMA issuer
    if (leader) {
      const uint32_t idesc = umma_idesc_bf16(kBM * kCG, BN) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
      // descriptor hi word: SBO 1024 B (next 8-row group) | version 1 | SWIZZLE_128B
      constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
      // descriptor lo word: start address >> 4 | LBO << 16.  K-major: LBO unused, +32 B per 16-k step.
      // MN-major: LBO = 8 KB between 64-element MN chunks, +2 KB (16 k-rows of 128 B) per step.
      const uint32_t a_lo0 = ((smem_u32(smem_a) & 0x3FFFFu) >> 4) | (a_mn ? ((uint32_t)(kMnChunkBytes >> 4) << 16) : 0u);
      const uint32_t b_lo0 = ((smem_u32(smem_b) & 0x3FFFFu) >> 4) | (b_mn ? ((uint32_t)(kMnChunkBytes >> 4) << 16) : 0u);
      const uint32_t a_kstep = a_mn ? (2048u >> 4) : 2u;
      const uint32_t b_kstep = b_mn ? (2048u >> 4) : 2u;
      int stage = 0;
      uint32_t phase = 0;
      int iter = 0;
      for (int u = cta_unit; u < num_units; u += num_ctas, ++iter) {
        const int split = u % split_k;
        const int kb0 = split * kb_per, kb1 = min(k_blocks, kb0 + kb_per);
        const int acc = iter & 1;
        const uint32_t acc_phase = (iter >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1u);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(acc * BN);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full[stage], phase);
          tcgen05_fence_after();
          if (elect_one()) {
            const uint32_t a_lo = a_lo0 + (uint32_t)stage * (Cfg::kABytes >> 4);
            const uint32_t b_lo = b_lo0 + (uint32_t)stage * (Cfg::kBBytes >> 4);
#pragma unroll
            for (int k = 0; k < kBK / kUmmaK; ++k) {
              const uint64_t da = ((uint64_t)desc_hi << 32) | (uint64_t)(a_lo + a_kstep * k);
              const uint64_t db = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + b_kstep * k);
              umma_bf16<kCG>(d_tmem, da, db, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            umma_commit<kCG>(&empty[stage]);  // frees this smem slot (in both CTAs) when the MMAs retire
            if (kb == kb1 - 1) umma_commit<kCG>(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == kStages) stage = 0, phase ^= 1u;
        }
      }
    }
  } else {
    // ========================================================================= epilogue
    const int quarter = warp & 3;         // TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;     // which half of the tile's columns
    const int row_in_tile = quarter * 32 + lane;
    constexpr int kChunks = BN / 64;      // 32-column chunks per warp
    int iter = 0;
    for (int u = cta_unit; u < num_units; u += num_ctas, ++iter) {
      const int tile = u / split_k, split = u - tile * split_k;
      const int m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int row = m_blk * tile_m + (int)cta_rank * kBM + row_in_tile;
      const int col0 = n_blk * BN + half * (BN / 2);
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const bool row_ok = row < M;
      const bool add_bias = ep.bias != nullptr && split == 0;
      const float* gate_row = nullptr;
      if (ep.epilogue == DITB200_EPI_BIAS_GATE_RESID && row_ok)
        gate_row = ep.gate + (size_t)(row / ep.rows_per_gate) * ep.gate_stride;
#pragma unroll 1
      for (int c = 0; c < kChunks; ++c) {
        uint32_t v[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) +
                               (uint32_t)(acc * BN + half * (BN / 2) + c * 32);
        tmem_ld_32x32(taddr, v);
        tmem_ld_wait();
        const int col = col0 + c * 32;
        if (!row_ok || col >= N) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (add_bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col + j < N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col + j));
              f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
            }
          }
        }
        if (ep.aux_out != nullptr) {  // the pre-activation / un-gated branch value, kept for backward
          __nv_bfloat16* arow = ep.aux_out + (size_t)row * N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col + j < N) {
              uint4 pk;
              pk.x = pack_bf16x2(f[j], f[j + 1]), pk.y = pack_bf16x2(f[j + 2], f[j + 3]);
              pk.z = pack_bf16x2(f[j + 4], f[j + 5]), pk.w = pack_bf16x2(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(arow + j) = pk;
            }
          }
        }
        if (ep.epilogue == DITB200_EPI_BIAS_GELU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_tanh_fast(f[j]);
        } else if (ep.epilogue == DITB200_EPI_BIAS_SILU) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = silu_f(f[j]);
        } else if (ep.epilogue == DITB200_EPI_BIAS_GATE_RESID) {
          const float* rrow = ep.resid + (size_t)row * N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col + j < N) {
              const float4 g4 = __ldg(reinterpret_cast<const float4*>(gate_row + col + j));
              const float4 r4 = *reinterpret_cast<const float4*>(rrow + j);
              f[j] = fmaf(g4.x, f[j], r4.x);
              f[j + 1] = fmaf(g4.y, f[j + 1], r4.y);
              f[j + 2] = fmaf(g4.z, f[j + 2], r4.z);
              f[j + 3] = fmaf(g4.w, f[j + 3], r4.w);
            }
          }
        } else if (ep.epilogue == DITB200_EPI_MUL_DGELU) {
          const __nv_bfloat16* urow = ep.aux_in + (size_t)row * N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col + j < N) {
              const uint4 pk = *reinterpret_cast<const uint4*>(urow + j);
              const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&pk);
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const float2 uu = __bfloat1622float2(h2[q]);
                f[j + 2 * q] *= dgelu_tanh_f(uu.x);
                f[j + 2 * q + 1] *= dgelu_tanh_f(uu.y);
              }
            }
          }
        }
        if (ep.out_bf16) {
          __nv_bfloat16* orow = reinterpret_cast<__nv_bfloat16*>(ep.out) + (size_t)row * N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            if (col + j < N) {
              uint4 pk;
              pk.x = pack_bf16x2(f[j], f[j + 1]);
              pk.y = pack_bf16x2(f[j + 2], f[j + 3]);
              pk.z = pack_bf16x2(f[j + 4], f[j + 5]);
              pk.w = pack_bf16x2(f[j + 6], f[j + 7]);
              *reinterpret_cast<uint4*>(orow + j) = pk;
            }
          }
        } else if (ep.atomic) {
          float* orow = reinterpret_cast<float*>(ep.out) + (size_t)row * N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col + j < N)
              atomicAdd(reinterpret_cast<float4*>(orow + j), make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]));
          }
        } else {
          float* orow = reinterpret_cast<float*>(ep.out) + (size_t)row * N + col;
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (col + j < N)
              *reinterpret_cast<float4*>(orow + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          }
        }
      }
      // all tcgen05.ld of this stage are complete (wait::ld above): give the stage back
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCG == 1) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_leader(&tmem_empty[acc]);
      }
    }
  }

  tcgen05_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc<kCG>(tmem_base, Cfg::kTmemCols);
}

// ----------------------------------------------------------------------------- host side
// 2-D bf16 tensor map over a row-major [rows, cols] matrix with a {box_rows x box_cols} box.
static int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint32_t box_rows, uint32_t box_cols) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) {
    set_error("gemm: ditb200_init() has not been called");
    return DITB200_ENOINIT;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};  // bytes, dim 1
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm: cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box=%ux%u)",
              (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
    return DITB200_EINVAL;
  }
  return 0;
}

template <int kCG, int BN>
static int launch_cfg(const ditb200_gemm_args* a, int split_k, cudaStream_t st) {
  using Cfg = TcCfg<kCG, BN>;
  CUtensorMap ta, tb;
  int rc;
  // K-major operand [MN, K]: box {128 (or B rows) x 64 k}.  MN-major operand stored [K, MN]: box {64 k x 64 MN}.
  if (!a->trans_a) rc = make_tmap_2d(&ta, a->a, (uint64_t)a->M, (uint64_t)a->K, kBM, kBK);
  else rc = make_tmap_2d(&ta, a->a, (uint64_t)a->K, (uint64_t)a->M, kBK, 64);
  if (rc) return rc;
  if (!a->trans_w) rc = make_tmap_2d(&tb, a->w, (uint64_t)a->N, (uint64_t)a->K, Cfg::kBRows, kBK);
  else rc = make_tmap_2d(&tb, a->w, (uint64_t)a->K, (uint64_t)a->N, kBK, 64);
  if (rc) return rc;
  EpiParams ep;
  ep.bias = a->bias, ep.out = a->out, ep.resid = a->resid, ep.gate = a->gate;
  ep.aux_out = reinterpret_cast<__nv_bfloat16*>(a->aux_out);
  ep.aux_in = reinterpret_cast<const __nv_bfloat16*>(a->aux_in);
  ep.gate_stride = a->gate_stride, ep.rows_per_gate = a->rows_per_gate;
  ep.epilogue = a->epilogue, ep.out_bf16 = (a->out_dtype == DITB200_BF16);
  ep.atomic = (split_k > 1 || a->accumulate) ? 1 : 0;
  static bool attr_set = false;  // per instantiation; benign race (idempotent)
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<kCG, BN>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "gemm_tc smem attribute");
    attr_set = true;
  }
  if (split_k > 1 && !a->accumulate) {  // partial tiles are summed into out: start from zero
    cudaError_t e = cudaMemsetAsync(a->out, 0, (size_t)a->M * a->N * sizeof(float), st);
    if (e != cudaSuccess) return check_cuda(e, "gemm_tc split-K memset");
  }
  const int tile_m = kBM * kCG;
  const int units = ((a->M + tile_m - 1) / tile_m) * ((a->N + BN - 1) / BN) * split_k;
  int ctas = num_sms() / kCG;
  if (ctas > units) ctas = units;
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)(ctas * kCG));
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<kCG, BN>, ta, tb, ep, a->M, a->N, a->K,
                                     a->trans_a ? 1 : 0, a->trans_w ? 1 : 0, split_k);
  if (e != cudaSuccess) return check_cuda(e, "gemm_tc launch");
  return 0;
}

// Tile choice, from measurements on B200 (profiles/r01_gemm_notes.md): the widest tile wins for every
// DiT shape because the per-k-block issue/barrier cost is amortised over twice the MMA work, and the
// CTA pair halves each SM's B traffic.  Narrow tiles only when N itself is narrow; single CTAs when M
// fits one 128-row tile (adaLN, M = batch).  An MN-major B operand is fetched in 64-row boxes, so its
// per-CTA share must be a multiple of 64 (no 192-wide pair tile).
static void choose_tile(int M, int N, int trans_w, int* cg_out, int* bn_out) {
  *cg_out = (M > kBM) ? 2 : 1;
  int bn = (N > 192) ? 256 : (N > 128) ? 192 : 128;
  if (trans_w && bn == 192) bn = 256;
  *bn_out = bn;
}

int launch_gemm_tcgen05(const ditb200_gemm_args* a, cudaStream_t st) {
  DITB_REQUIRE(is_initialised(), DITB200_ENOINIT, "gemm: ditb200_init() has not been called");
  // a K-major operand has K as its contiguous dimension: TMA needs 16-byte row pitches
  DITB_REQUIRE((a->trans_a && a->trans_w) || a->K % 8 == 0, DITB200_EINVAL,
               "gemm(tcgen05): K=%d must be a multiple of 8", a->K);
  DITB_REQUIRE(a->N % 8 == 0, DITB200_EINVAL, "gemm(tcgen05): N=%d must be a multiple of 8", a->N);
  DITB_REQUIRE(!a->trans_a || a->M % 8 == 0, DITB200_EINVAL, "gemm(tcgen05): trans_a needs M %% 8 == 0 (M=%d)", a->M);
  DITB_REQUIRE(aligned16(a->a) && aligned16(a->w) && aligned16(a->out), DITB200_EALIGN,
               "gemm(tcgen05): a, w, out must be 16-byte aligned");
  DITB_REQUIRE(!a->bias || aligned16(a->bias), DITB200_EALIGN, "gemm(tcgen05): bias misaligned");
  if (a->epilogue == DITB200_EPI_BIAS_GATE_RESID)
    DITB_REQUIRE(aligned16(a->resid) && aligned16(a->gate) && a->gate_stride % 4 == 0,
                 DITB200_EALIGN, "gemm(tcgen05): resid/gate misaligned");
  if (a->epilogue == DITB200_EPI_MUL_DGELU)
    DITB_REQUIRE(a->aux_in && aligned16(a->aux_in) && a->aux_dtype == DITB200_BF16, DITB200_EINVAL,
                 "gemm(tcgen05): MUL_DGELU needs a 16-byte-aligned bf16 aux_in");
  if (a->aux_out)
    DITB_REQUIRE(aligned16(a->aux_out) && a->aux_dtype == DITB200_BF16, DITB200_EINVAL,
                 "gemm(tcgen05): aux_out must be 16-byte-aligned bf16");
  int split_k = a->split_k > 1 ? a->split_k : 1;
  const int k_blocks = (a->K + kBK - 1) / kBK;
  if (split_k > k_blocks) split_k = k_blocks;
  if (split_k > 1) {  // every split must own at least one k-block
    const int kb_per = (k_blocks + split_k - 1) / split_k;
    split_k = (k_blocks + kb_per - 1) / kb_per;
  }
  if (split_k > 1 || a->accumulate) {
    DITB_REQUIRE(a->out_dtype == DITB200_F32 && a->epilogue == DITB200_EPI_BIAS && !a->aux_out, DITB200_EINVAL,
                 "gemm(tcgen05): split_k / accumulate need an f32 output and the plain bias epilogue");
  }
  int cg = a->cta_group, bn = a->tile_n;
  if (cg == 0 || bn == 0) {
    int acg, abn;
    choose_tile(a->M, a->N, a->trans_w, &acg, &abn);
    if (cg == 0) cg = acg;
    if (bn == 0) bn = abn;
  }
  DITB_REQUIRE(!a->trans_w || (bn / cg) % 64 == 0, DITB200_EINVAL,
               "gemm(tcgen05): trans_w needs tile_n / cta_group to be a multiple of 64 (got %d / %d)", bn, cg);
#define TC_CASE(CG, BN_)                   \
  if (cg == CG && bn == BN_) return launch_cfg<CG, BN_>(a, split_k, st);
  TC_CASE(1, 128)
  TC_CASE(1, 192)
  TC_CASE(1, 256)
  TC_CASE(2, 128)
  TC_CASE(2, 192)
  TC_CASE(2, 256)
#undef TC_CASE
  set_error("gemm(tcgen05): unsupported tile cta_group=%d tile_n=%d", cg, bn);
  return DITB200_EINVAL;
}

}  // namespace ditb200
