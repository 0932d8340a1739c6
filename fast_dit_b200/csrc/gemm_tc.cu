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
#include <stdlib.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace ditb200 {

constexpr int kBM = 128;       // rows per CTA tile = TMEM lanes
constexpr int kBK = 64;        // 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;     // K per tcgen05.mma for 16-bit inputs
constexpr int kEpiWarps = 8;
constexpr int kNumThreads = 64 + 32 * kEpiWarps;
constexpr int kEpiStageBytes = 2 * 32 * 128;             // per epilogue warp: 32 rows x 32 f32 transpose buffer, or two
                                                         // 32 x 64 bf16 TMA-store tiles
constexpr int kBarBytes = 512;                           // mbarriers, TMEM address, CLC responses
constexpr int kClcSlots = 6;                             // cluster-launch-control responses in flight / unread
constexpr int kClcAhead = 2;                             // queries kept in flight (their latency is about one tile)
constexpr int kSmemBudget = 232448 - 1024 - kBarBytes - kEpiWarps * kEpiStageBytes;  // what is left for the TMA ring
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
  int tma_store;  // bf16 output through smem staging + TMA store (tma_out is valid)
};

// Explicit tile schedule (kernel parameter, read through the constant bank): for every CTA pair the list of tiles it
// processes, in order.  Used where mixing tile widths balances better than any uniform cover — N = 1152 on 74 pairs is
// 5.19 rounds of 192-wide tiles, but 56 row panels cut 256|256|256|192|192 and 8 cut 6 x 192 pack into 42 pairs x
// four 256-wide + 32 pairs x five 192-wide tiles with nothing left over (launch_cfg: plan_table).
// entry = m_blk << 16 | (first column / 16) << 5 | (columns / 16)
constexpr int kTabCap = 1024, kTabPairs = 160;
struct SchedTable {
  int use;
  uint16_t start[kTabPairs + 1];
  uint32_t e[kTabCap];
};

template <int kCG, int BN>
struct TcCfg {
  static constexpr int kBRows = BN / kCG;  // B rows held by each CTA
  static constexpr int kABytes = kBM * kBK * 2;
  static constexpr int kBBytes = kBRows * kBK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (kSmemBudget / kStageBytes) > 8 ? 8 : (kSmemBudget / kStageBytes);
  static constexpr int kTmemCols = (2 * BN <= 32) ? 32 : (2 * BN <= 64) ? 64 : (2 * BN <= 128) ? 128 : (2 * BN <= 256) ? 256 : 512;
  static constexpr int kSmemBytes = kStages * kStageBytes + kEpiWarps * kEpiStageBytes + 1024 /*align slack*/ + kBarBytes;
  static_assert(BN % 64 == 0 && BN >= 64 && BN <= 256, "UMMA N / epilogue column split (32-column chunks per half)");
  // epilogue column split between the two warps of a lane quarter: [0, kHalf0) and [kHalf0, BN)
  static constexpr int kHalf0 = BN / 2;
  static_assert(2 * BN <= 512, "two accumulator stages must fit TMEM");
  static_assert(kABytes % 1024 == 0 && kBBytes % 1024 == 0, "swizzle-128B tiles need 1024-B alignment");
};

// d/du gelu_tanh(u)
__device__ __forceinline__ float dgelu_tanh_f(float u) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u2 = u * u;
  const float z = k0 * (u + k1 * u2 * u);
  const float t = tanh_approx(z);
  return 0.5f * (1.0f + t) + 0.5f * u * (1.0f - t * t) * k0 * (1.0f + 3.0f * k1 * u2);
}

// gelu_tanh(u) and its derivative from one tanh evaluation
__device__ __forceinline__ void gelu_and_dgelu(float u, float& g, float& d) {
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  const float u2 = u * u;
  const float t = tanh_approx(k0 * u * fmaf(k1, u2, 1.0f));
  const float hp = 0.5f * (1.0f + t);
  g = u * hp;
  d = fmaf(0.5f * u * (1.0f - t * t), k0 * fmaf(3.0f * k1, u2, 1.0f), hp);
}

// ------------------------------------------------------------------------------ epilogue
// One epilogue warp owns 32 accumulator rows (its TMEM lane quarter) x NCH 32-column chunks of the tile.
// tcgen05.ld hands every thread one ROW (32 consecutive columns); storing that way would touch 32 different
// 128-byte lines per instruction.  The warp therefore transposes each chunk through its private 4 KB staging
// buffer (16-byte slots XOR-swizzled by row: conflict-free in both directions) and does all global traffic with
// lane = 4 consecutive columns: 8 lanes cover one 128-byte row segment, one instruction covers 4 full rows.
// EPI / OUT / AUX are compile-time (-1 = decide at run time): the kernel dispatches once per tile to a
// straight-line instantiation instead of re-deciding the epilogue flavour for every row.
enum { OUT_BF16 = 0, OUT_F32 = 1, OUT_ATOMIC = 2 };

template <int EPI, int OUT, int AUX, int NCH>
__device__ __forceinline__ void epi_tile(const EpiParams& ep, const uint32_t taddr0, const uint32_t stg, const int lane,
                                         const int row0, const int colw, const int nch, const int M, const int N,
                                         const bool add_bias) {
  const int epi = EPI >= 0 ? EPI : ep.epilogue;
  const int omode = OUT >= 0 ? OUT : (ep.out_bf16 ? OUT_BF16 : (ep.atomic ? OUT_ATOMIC : OUT_F32));
  const bool aux = AUX >= 0 ? (AUX != 0) : (ep.aux_out != nullptr);
  const int sub = lane >> 3, grp = lane & 7;
  const int rows_valid = M - row0;  // rows [0, rows_valid) of this warp's 32 exist (may be <= 0 or >= 32)
  const uint32_t my_row = stg + (uint32_t)lane * 128u;  // shared-window addresses
  const int swz_w = lane & 7;
  // adaLN gate: one [N] vector per image.  Fast path: all 32 rows belong to the same image.
  const float* gate_base = nullptr;
  bool gate_uniform = true;
  if (epi == DITB200_EPI_BIAS_GATE_RESID) {
    const int last = min(row0 + 31, M - 1);
    const int g_lo = row0 / ep.rows_per_gate, g_hi = max(last, row0) / ep.rows_per_gate;
    gate_uniform = g_lo == g_hi;
    gate_base = ep.gate + (size_t)(rows_valid > 0 ? g_lo : 0) * ep.gate_stride;
  }
  float4 r4[8];  // residual values of the chunk, requested before the accumulators are waited for (GATE_RESID only)
  uint2 u2[8];   // fc1 pre-activations of the chunk (MUL_DGELU only), same idea
  auto load_resid = [&](int c, float4 (&dst)[8]) {
    const int col = colw + c * 32 + grp * 4;
    const float* rp = ep.resid + (size_t)row0 * N + col;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int r = i * 4 + sub;
      if (col < N && r < rows_valid) dst[i] = ldg_pinned_f4(rp + (size_t)r * N);
    }
  };
#pragma unroll 1
  for (int c = 0; c < NCH; ++c) {
    if (c >= nch || colw + c * 32 >= N) break;  // warp-uniform: narrow last tile column / columns past the matrix
    uint32_t v[32];
    tmem_ld_32x32(taddr0 + (uint32_t)(c * 32), v);
    if (epi == DITB200_EPI_BIAS_GATE_RESID) load_resid(c, r4);
    if (epi == DITB200_EPI_MUL_DGELU || epi == DITB200_EPI_MUL_AUX) {
      const int colp = colw + c * 32 + grp * 4;
      const __nv_bfloat16* up = ep.aux_in + (size_t)row0 * N + colp;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub;
        if (colp < N && r < rows_valid) u2[i] = ldg_pinned_u2(up + (size_t)r * N);
      }
    }
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 8; ++g)
      sts128(my_row + (uint32_t)((g ^ swz_w) << 4), v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    __syncwarp();
    const int col = colw + c * 32 + grp * 4;
    if (col < N) {
      float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (add_bias) b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col));
      float4 g4 = make_float4(0.f, 0.f, 0.f, 0.f);
      if (epi == DITB200_EPI_BIAS_GATE_RESID && gate_uniform) g4 = __ldg(reinterpret_cast<const float4*>(gate_base + col));
      const size_t off0 = (size_t)row0 * N + col;
      float4 fr[8];  // all eight staged rows first: the shared loads overlap instead of one per row iteration
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub;
        fr[i] = lds128_f(stg + (uint32_t)(r * 128 + ((grp ^ (r & 7)) << 4)));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = i * 4 + sub;
        if (r < rows_valid) {
          float4 f = fr[i];
          f.x += b4.x, f.y += b4.y, f.z += b4.z, f.w += b4.w;
          const size_t off = off0 + (size_t)r * N;
          if (aux && epi != DITB200_EPI_BIAS_GELU_DAUX) {  // the pre-activation / un-gated branch value, kept for backward
            uint2 pk;
            pk.x = pack_bf16x2(f.x, f.y), pk.y = pack_bf16x2(f.z, f.w);
            *reinterpret_cast<uint2*>(ep.aux_out + off) = pk;
          }
          if (epi == DITB200_EPI_BIAS_GELU_DAUX) {
            float4 d;
            gelu_and_dgelu(f.x, f.x, d.x), gelu_and_dgelu(f.y, f.y, d.y);
            gelu_and_dgelu(f.z, f.z, d.z), gelu_and_dgelu(f.w, f.w, d.w);
            uint2 pk;
            pk.x = pack_bf16x2(d.x, d.y), pk.y = pack_bf16x2(d.z, d.w);
            *reinterpret_cast<uint2*>(ep.aux_out + off) = pk;
          } else if (epi == DITB200_EPI_MUL_AUX) {
            const uint2 pk = u2[i];
            const float2 u0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
            const float2 u1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
            f.x *= u0.x, f.y *= u0.y, f.z *= u1.x, f.w *= u1.y;
          } else if (epi == DITB200_EPI_BIAS_GELU) {
            f.x = gelu_tanh_fast(f.x), f.y = gelu_tanh_fast(f.y), f.z = gelu_tanh_fast(f.z), f.w = gelu_tanh_fast(f.w);
          } else if (epi == DITB200_EPI_BIAS_SILU) {
            f.x = silu_f(f.x), f.y = silu_f(f.y), f.z = silu_f(f.z), f.w = silu_f(f.w);
          } else if (epi == DITB200_EPI_BIAS_GATE_RESID) {
            if (!gate_uniform)
              g4 = __ldg(reinterpret_cast<const float4*>(ep.gate + (size_t)((row0 + r) / ep.rows_per_gate) * ep.gate_stride + col));
            const float4 rr = r4[i];
            f.x = fmaf(g4.x, f.x, rr.x), f.y = fmaf(g4.y, f.y, rr.y), f.z = fmaf(g4.z, f.z, rr.z), f.w = fmaf(g4.w, f.w, rr.w);
          } else if (epi == DITB200_EPI_MUL_DGELU) {
            const uint2 pk = u2[i];
            const float2 u0 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.x));
            const float2 u1 = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&pk.y));
            f.x *= dgelu_tanh_f(u0.x), f.y *= dgelu_tanh_f(u0.y), f.z *= dgelu_tanh_f(u1.x), f.w *= dgelu_tanh_f(u1.y);
          }
          if (omode == OUT_BF16) {
            uint2 pk;
            pk.x = pack_bf16x2(f.x, f.y), pk.y = pack_bf16x2(f.z, f.w);
            *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(ep.out) + off) = pk;
          } else if (omode == OUT_ATOMIC) {
            atomicAdd(reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + off), f);
          } else {
            *reinterpret_cast<float4*>(reinterpret_cast<float*>(ep.out) + off) = f;
          }
        }
      }
    }
    __syncwarp();  // the staging buffer is rewritten by the next chunk
  }
}

// ------------------------------------------------------------------------------ tile schedule
// Static schedule shared by the three warp roles of a CTA (pair).  Work units are output tiles (x k-splits).
// When N is not a multiple of BN the last tile column is NARROW: its tcgen05.mma runs with N = part_cols
// (the instruction descriptor is a run-time value) instead of multiplying zero-filled rows, and the schedule
// hands those cheap tiles to the CTAs that received one full tile less (longest-processing-time order):
//   phase 0  full tiles f = p, p+P, ...            (n-fastest, so concurrent CTAs share A row panels in L2)
//   phase 1  CTAs p >= r (r = F mod P, the "light" ones) take q narrow tiles each
//   phase 2  whatever narrow tiles are left, round-robin
// With split_k > 1 the plain round-robin over (tile, split) units is kept.
struct TileSched {
  int P, p;                 // CTAs (pairs) in the grid, this CTA's index
  int n_full, part_cols;    // full tile columns, width of the narrow last column (0 = none)
  int F, H, r, q;           // full tiles, narrow tiles, light-CTA threshold, narrow tiles per light CTA
  int split_k, n_tiles, num_units;
  int m_last;               // >= 0: tile rows are visited last-first (m_blk -> m_last - m_blk); -1: first-first
  int col0;                 // first output column of the current tile (n_blk * bn, or the table's)
  const uint32_t* tab;      // explicit schedule (SchedTable::e) or nullptr
  int ti, tend;             // this pair's range of table entries
  int phase, cur, end_a;
  // dynamic mode (cluster launch control): the grid has one cluster per work unit; a running cluster finishes its
  // own unit and then cancels clusters that have not been launched yet and does their units.  SMs that become
  // free late (another kernel, e.g. an NCCL collective, held them) simply never receive work instead of owning a
  // fixed share of the schedule.  Responses land in a ring of kClcSlots 16-byte slots in every CTA of the pair.
  int dyn, role, it;        // role: 0 = producer of the leader CTA (issues the queries), 1 = producer of the peer CTA, 2 = reader
  uint32_t clc_resp;        // shared address of the response ring
  uint64_t *clc_full, *clc_empty;

  __host__ __device__ __forceinline__ void init(int M, int N, int tile_m, int bn, int part, int split, int pairs, int pair,
                                               int reverse_m = 0) {
    P = pairs, p = pair, split_k = split;
    const int m_tiles = (M + tile_m - 1) / tile_m;
    m_last = reverse_m ? m_tiles - 1 : -1;
    col0 = 0, tab = nullptr, ti = tend = 0;
    n_tiles = (N + bn - 1) / bn;
    num_units = m_tiles * n_tiles * split_k;
    part_cols = part;
    n_full = part ? n_tiles - 1 : n_tiles;
    F = m_tiles * n_full, H = part ? m_tiles : 0;
    r = F % P;
    q = part ? max(1, bn / max(part, bn / 2)) : 1;  // a narrow tile still loads the whole A tile: never cheaper than half a full one
    phase = 0, cur = p, end_a = 0;
    dyn = 0, role = 2, it = 0;
  }
  template <int kCG>
  __device__ __forceinline__ bool next_dyn(int bn, int& m_blk, int& n_blk, int& ncols, int& split) {
    int u = p;
    if (it > 0) {  // the unit of iteration `it` is the answer to query `it`
      const int s = (it - 1) % kClcSlots;
      mbar_wait(&clc_full[s], (uint32_t)((it - 1) / kClcSlots) & 1u);
      uint32_t valid, x;
      asm volatile(
          "{\n\t.reg .pred p1;\n\t.reg .b128 r;\n\t"
          "ld.shared.b128 r, [%2];\n\t"
          "clusterlaunchcontrol.query_cancel.is_canceled.pred.b128 p1, r;\n\t"
          "selp.u32 %1, 1, 0, p1;\n\t"
          "mov.u32 %0, 0;\n\t"
          "@p1 clusterlaunchcontrol.query_cancel.get_first_ctaid.v4.b32.b128 {%0, _, _, _}, r;\n\t}"
          : "=r"(x), "=r"(valid)
          : "r"(clc_resp + 16u * (uint32_t)s)
          : "memory");
      fence_proxy_async();  // the slot's next writer is the asynchronous proxy
      __syncwarp();
      if ((threadIdx.x & 31) == 0) {
        if (role <= 1) mbar_arrive_expect_tx(&clc_full[s], 16);  // re-arm this CTA's slot for query it + kClcSlots
        if constexpr (kCG == 1) mbar_arrive(&clc_empty[s]); else mbar_arrive_leader(&clc_empty[s], 0);
      }
      if (!valid) {
        // queries it+1 .. it+kClcAhead-1 are still in flight (they fail too): their answers must have landed in
        // this CTA's shared memory before the CTA may exit
        if (role <= 1)
          for (int j = it + 1; j < it + kClcAhead; ++j) mbar_wait(&clc_full[(j - 1) % kClcSlots], (uint32_t)((j - 1) / kClcSlots) & 1u);
        return false;
      }
      u = (int)x / kCG;
    }
    if (role == 0) {
      // keep kClcAhead queries in flight; query j goes into slot (j - 1) % kClcSlots once every reader is done with
      // the slot's previous answer.  No query is issued after a failed answer has been OBSERVED (undefined by the
      // PTX model): at this point answers up to `it` were valid.
      for (int j = (it == 0) ? 1 : it + kClcAhead; j <= it + kClcAhead; ++j) {
        const int s = (j - 1) % kClcSlots;
        mbar_wait(&clc_empty[s], ((uint32_t)((j - 1) / kClcSlots) & 1u) ^ 1u);
        if (elect_one()) {
          if constexpr (kCG == 1)
            asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.b128 [%0], [%1];"
                         ::"r"(clc_resp + 16u * (uint32_t)s), "r"(smem_u32(&clc_full[s])) : "memory");
          else
            asm volatile("clusterlaunchcontrol.try_cancel.async.shared::cta.mbarrier::complete_tx::bytes.multicast::cluster::all.b128 [%0], [%1];"
                         ::"r"(clc_resp + 16u * (uint32_t)s), "r"(smem_u32(&clc_full[s])) : "memory");
        }
        __syncwarp();
      }
    }
    ++it;
    if (split_k > 1 || part_cols == 0) {
      const int tile = u / split_k;
      split = u - tile * split_k;
      m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      ncols = (part_cols && n_blk == n_tiles - 1) ? part_cols : bn;
    } else {  // full tiles first, the narrow column last
      split = 0;
      if (u < F) m_blk = u / n_full, n_blk = u - m_blk * n_full, ncols = bn;
      else m_blk = u - F, n_blk = n_full, ncols = part_cols;
    }
    return true;
  }
  // next unit of this CTA: tile coordinates, tile width in columns, k-split index.  With m_last >= 0 the same
  // schedule runs over the tile rows mirrored: the kernel then starts with the rows its producer kernel wrote LAST
  // (still in L2) and finishes with the rows the next kernel, traversing the other way, reads first.
  template <int kCG>
  __host__ __device__ __forceinline__ bool next(int bn, int& m_blk, int& n_blk, int& ncols, int& split) {
    bool ok;
    if (tab != nullptr) {
      ok = ti < tend;
      if (ok) {
        const uint32_t e = tab[ti++];
        m_blk = (int)(e >> 16), col0 = (int)((e >> 5) & 0x7ffu) << 4, ncols = (int)(e & 31u) << 4;
        n_blk = col0 / bn, split = 0;
      }
    } else {
      ok = next_unmirrored<kCG>(bn, m_blk, n_blk, ncols, split);
      col0 = n_blk * bn;
    }
    if (ok && m_last >= 0) m_blk = m_last - m_blk;
    return ok;
  }
  template <int kCG>
  __host__ __device__ __forceinline__ bool next_unmirrored(int bn, int& m_blk, int& n_blk, int& ncols, int& split) {
#ifdef __CUDA_ARCH__
    if (dyn) return next_dyn<kCG>(bn, m_blk, n_blk, ncols, split);
#endif
    if (split_k > 1 || part_cols == 0) {  // plain round-robin
      if (cur >= num_units) return false;
      const int tile = cur / split_k;
      split = cur - tile * split_k;
      m_blk = tile / n_tiles, n_blk = tile - m_blk * n_tiles;
      ncols = (part_cols && n_blk == n_tiles - 1) ? part_cols : bn;
      cur += P;
      return true;
    }
    split = 0;
    if (phase == 0) {
      if (cur < F) {
        m_blk = cur / n_full, n_blk = cur - m_blk * n_full, ncols = bn;
        cur += P;
        return true;
      }
      phase = 1;
      cur = (p >= r) ? (p - r) * q : H;
      end_a = min(H, cur + q);
    }
    if (phase == 1) {
      if (cur < end_a) {
        m_blk = cur, n_blk = n_full, ncols = part_cols;
        ++cur;
        return true;
      }
      phase = 2;
      cur = (P - r) * q + p;
    }
    if (cur < H) {
      m_blk = cur, n_blk = n_full, ncols = part_cols;
      cur += P;
      return true;
    }
    return false;
  }
};

// bf16 outputs without a second operand (bias / bias+GELU): the accumulator rows a thread gets from tcgen05.ld
// are finished in place, packed, written to one of the warp's two 4 KB staging tiles in the 128-byte-swizzled
// layout of the output tensor map, and a single elected lane hands the 32-row x 64-column tile to TMA: whole
// 128-byte lines per row, no transposing read-back, no per-row address arithmetic, no LSU store wavefronts;
// rows / columns past the matrix are clipped by the tensor map.  A staging tile is reused once the bulk group
// that read it has drained (two in flight).
// MODE 0 bias, 1 bias+GELU, 2 bias+GELU' (second pass of GELU_DAUX over the same accumulators, into aux_out),
// 3 (acc + bias) * aux_in with the aux_in tile fetched by TMA one round ahead, 5 GELU_DAUX in ONE pass: gelu into
// staging tile 0 -> out, gelu' into tile 1 -> aux_out (tma_ld), one tanh per element, both stores per round.
template <int MODE, int NCH>
__device__ __forceinline__ void epi_tile_tma(const EpiParams& ep, const CUtensorMap* tma_st, const CUtensorMap* tma_ld,
                                             uint64_t* ebar, uint32_t& ephase, const uint32_t taddr0,
                                             const uint32_t stg, const int lane, const int row0, const int colw,
                                             const int nch, const int N, const bool add_bias, int& buf) {
  const uint32_t my_off = (uint32_t)lane * 128u;
  const int sw = lane & 7;
  if (MODE == 3) {  // tile 0 receives aux_in, tile 1 is the store source
    if (lane == 0 && nch > 0 && colw < N) {
      mbar_arrive_expect_tx(ebar, 32 * 128);
      tma_load_2d(tma_ld, ebar, reinterpret_cast<void*>(__cvta_shared_to_generic((size_t)stg)), colw, row0);
    }
  }
#pragma unroll 1
  for (int c = 0; c < NCH; c += 2) {  // 64 columns per store
    const int col = colw + c * 32;
    if (c >= nch || col >= N) break;  // warp-uniform
    const uint32_t tile = MODE == 3 ? stg + 4096u : (MODE == 5 ? stg : stg + (uint32_t)buf * 4096u);
    if (lane == 0) {  // the store that last used this staging tile has read it
      if (MODE == 3 || MODE == 5) bulk_wait_group_read<0>(); else bulk_wait_group_read<1>();
    }
    __syncwarp();
    if (MODE == 3) mbar_wait(ebar, ephase);
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      const int colh = col + hh * 32;
      uint32_t v[32];
      if (c + hh < nch && colh < N) {  // warp-uniform; a missing half lies past the matrix (the split above) and is clipped
        tmem_ld_32x32(taddr0 + (uint32_t)((c + hh) * 32), v);
        uint4 ax[4];
        if (MODE == 3) {
#pragma unroll
          for (int g = 0; g < 4; ++g) ax[g] = lds128_u(stg + my_off + (uint32_t)(((hh * 4 + g) ^ sw) << 4));
        }
        tmem_ld_wait();
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (add_bias) {
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (colh + j < N) {
              const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + colh + j));
              f[j] += b4.x, f[j + 1] += b4.y, f[j + 2] += b4.z, f[j + 3] += b4.w;
            }
          }
        }
        if (MODE == 5) {  // f <- gelu, second tile <- gelu'
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float d[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) gelu_and_dgelu(f[8 * g + j], f[8 * g + j], d[j]);
            sts128(stg + 4096u + my_off + (uint32_t)(((hh * 4 + g) ^ sw) << 4), pack_bf16x2(d[0], d[1]),
                   pack_bf16x2(d[2], d[3]), pack_bf16x2(d[4], d[5]), pack_bf16x2(d[6], d[7]));
          }
        } else if (MODE == 1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = gelu_tanh_fast(f[j]);
        } else if (MODE == 2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float gdummy;
            gelu_and_dgelu(f[j], gdummy, f[j]);
          }
        } else if (MODE == 3) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&ax[g]);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const float2 m = __bfloat1622float2(h2[q]);
              f[8 * g + 2 * q] *= m.x, f[8 * g + 2 * q + 1] *= m.y;
            }
          }
        }
#pragma unroll
        for (int g = 0; g < 4; ++g)
          sts128(tile + my_off + (uint32_t)(((hh * 4 + g) ^ sw) << 4), pack_bf16x2(f[8 * g], f[8 * g + 1]),
                 pack_bf16x2(f[8 * g + 2], f[8 * g + 3]), pack_bf16x2(f[8 * g + 4], f[8 * g + 5]),
                 pack_bf16x2(f[8 * g + 6], f[8 * g + 7]));
      }
    }
    fence_proxy_async();  // make this thread's shared-memory writes visible to the async (TMA) proxy
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tma_st, tile, col, row0);
      if (MODE == 5) tma_store_2d(tma_ld, stg + 4096u, col, row0);  // aux_out tile
      bulk_commit_group();
      if (MODE == 3 && c + 2 < nch && col + 64 < N) {  // next round's aux_in tile (every lane has read this one)
        mbar_arrive_expect_tx(ebar, 32 * 128);
        tma_load_2d(tma_ld, ebar, reinterpret_cast<void*>(__cvta_shared_to_generic((size_t)stg)), col + 64, row0);
      }
    }
    if (MODE == 3) ephase ^= 1u; else if (MODE != 5) buf ^= 1;
  }
}

// kMC = CTA pairs per cluster.  With kMC == 2 (cluster of 4 CTAs) the two pairs work on vertically adjacent
// 256-row tiles of the same tile column and SHARE the B tile: every CTA fetches one quarter of it and TMA
// multicasts that quarter to the CTA holding the same B half in the other pair, so a 512 x 256 super-tile moves
// 96 KB per k-block out of L2 instead of 128 KB.  The L2 -> SM fabric, not the tensor pipe, bounds these GEMMs.
template <int kCG, int BN, int kMC>
__global__ void __launch_bounds__(kNumThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
               const __grid_constant__ CUtensorMap tma_out, const __grid_constant__ CUtensorMap tma_aux,
               const __grid_constant__ EpiParams ep, const int M, const int N, const int K, const int a_mn, const int b_mn,
               const int split_k, const int part_cols, const int dyn, const int reverse_m,
               const __grid_constant__ SchedTable stab) {
  using Cfg = TcCfg<kCG, BN>;
  constexpr int kStages = Cfg::kStages;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment for the 128B-swizzle atoms (same offset in both CTAs of a pair)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::kABytes;
  uint8_t* smem_epi = smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + kEpiWarps * kEpiStageBytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + kStages;
  uint64_t* tmem_full = bars + 2 * kStages;
  uint64_t* tmem_empty = bars + 2 * kStages + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);
  uint64_t* epi_bar = bars + 2 * kStages + 5;  // [kEpiWarps] aux_in tiles landing in the epilogue staging (MUL_AUX)
  uint64_t* clc_full = bars + 2 * kStages + 5 + kEpiWarps;  // [kClcSlots] a launch-control response has landed
  uint64_t* clc_empty = clc_full + kClcSlots;               // [kClcSlots] (leader CTA) every reader is done with it
  uint8_t* clc_resp = reinterpret_cast<uint8_t*>(bars) + kBarBytes - 16 * kClcSlots;
  static_assert((2 * 8 + 5 + kEpiWarps + 2 * kClcSlots) * 8 + 16 + 16 * kClcSlots <= kBarBytes, "barrier block");

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t cl_rank = (kCG == 2) ? cluster_ctarank() : 0u;  // rank in the cluster: pair = rank / 2
  const uint32_t cta_rank = cl_rank & 1u;                        // rank inside the CTA pair
  const uint32_t pair = cl_rank >> 1, lead_rank = cl_rank & ~1u;
  const bool leader = cta_rank == 0;
  static_assert(kMC == 1 || kCG == 2, "multicast clusters are built from CTA pairs");

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full[s], 1);    // one arrive.expect_tx (leader CTA) covering every CTA's bytes
      mbar_init(&empty[s], kMC);  // one tcgen05.commit per phase from every pair that reads (a copy of) the slot
    }
    for (int s = 0; s < kEpiWarps; ++s) mbar_init(&epi_bar[s], 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], kEpiWarps * kCG);  // one arrive per epilogue warp of every CTA
    }
    for (int s = 0; s < kClcSlots; ++s) {
      mbar_init(&clc_full[s], 1);                             // this CTA's producer arms it for 16 response bytes
      mbar_init(&clc_empty[s], kCG * (1 + kEpiWarps) + 1);  // producers and epilogue warps of the pair + the MMA warp
    }
    fence_barrier_init();
    if (dyn) for (int s = 0; s < kClcSlots; ++s) mbar_arrive_expect_tx(&clc_full[s], 16);
  }
  if (warp == 1) tmem_alloc<kCG>(tmem_ptr, Cfg::kTmemCols);
  tcgen05_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  // (PDL build variant only) barriers, TMEM and tensor maps are set up: let the next kernel's CTAs queue for this SM,
  // then wait for the previous kernel's results before the first operand load / residual read / output write
  DITB_PDL_TRIGGER();
  DITB_PDL_WAIT();

  const int tile_m = kBM * kCG;
  const int k_blocks = (K + kBK - 1) / kBK;
  const int kb_per = (k_blocks + split_k - 1) / split_k;
  TileSched sched;
  sched.init(M, N, tile_m * kMC, BN, part_cols, split_k, (int)gridDim.x / (kCG * kMC), (int)blockIdx.x / (kCG * kMC), reverse_m);
  if (kMC == 1 && stab.use) {
    const int pr = (int)blockIdx.x / kCG;
    sched.tab = stab.e, sched.ti = stab.start[pr], sched.tend = stab.start[pr + 1];
  }
  if (kMC == 1 && dyn) {
    sched.dyn = 1, sched.role = (warp == 0) ? (leader ? 0 : 1) : 2;
    sched.clc_resp = smem_u32(clc_resp), sched.clc_full = clc_full, sched.clc_empty = clc_empty;
  }
  int m_blk, n_blk, ncols, split;

  // Producer and MMA roles run warp-uniform loops (every lane waits on the barriers) and elect one
  // lane only around the asynchronous issues.  Keeping the control flow uniform lets the compiler
  // hold descriptors / barrier addresses in uniform registers; a divergent `if (lane == 0)` loop
  // costs a vector->uniform register move in front of every UTCHMMA and made the issue thread the
  // bottleneck (profiles/r01_gemm_notes.md).
  if (warp == 0) {
    // ===================================================================== TMA producer
    int stage = 0;
    uint32_t phase = 0;
    while (sched.next<kCG>(BN, m_blk, n_blk, ncols, split)) {
      const int row_a = (m_blk * kMC + (int)pair) * tile_m + (int)cta_rank * kBM;
      const int row_b = sched.col0 + (int)cta_rank * (ncols / kCG);  // each CTA of a pair holds half of the tile's B rows
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
            else tma_load_2d_pair(&tma_a, &full[stage], dst_a, kb * kBK, row_a, lead_rank);
          } else {
#pragma unroll
            for (int c = 0; c < kBM / 64; ++c) {
              if constexpr (kCG == 1) tma_load_2d(&tma_a, &full[stage], dst_a + c * kMnChunkBytes, row_a + 64 * c, kb * kBK);
              else tma_load_2d_pair(&tma_a, &full[stage], dst_a + c * kMnChunkBytes, row_a + 64 * c, kb * kBK, lead_rank);
            }
          }
          if constexpr (kMC == 2) {
            // this CTA's quarter of the B tile (box = kBRows / 2 rows), delivered to both CTAs holding this half
            constexpr int kQRows = Cfg::kBRows / 2;
            tma_load_2d_pair_mc(&tma_b, &full[stage], dst_b + (int)pair * (Cfg::kBBytes / 2), kb * kBK,
                                row_b + (int)pair * kQRows, (uint16_t)(0x5u << cta_rank));
          } else if (!b_mn) {
            if constexpr (kCG == 1) tma_load_2d(&tma_b, &full[stage], dst_b, kb * kBK, row_b);
            else tma_load_2d_pair(&tma_b, &full[stage], dst_b, kb * kBK, row_b, lead_rank);
          } else {
#pragma unroll
            for (int c = 0; c < Cfg::kBRows / 64; ++c) {
              if constexpr (kCG == 1) tma_load_2d(&tma_b, &full[stage], dst_b + c * kMnChunkBytes, row_b + 64 * c, kb * kBK);
              else tma_load_2d_pair(&tma_b, &full[stage], dst_b + c * kMnChunkBytes, row_b + 64 * c, kb * kBK, lead_rank);
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
      const uint32_t idesc0 = umma_idesc_bf16(kBM * kCG, 0) | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
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
      for (; sched.next<kCG>(BN, m_blk, n_blk, ncols, split); ++iter) {
        const uint32_t idesc = idesc0 | ((uint32_t)(ncols >> 3) << 17);
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
            // frees this smem slot when the MMAs retire: in both CTAs of the pair, and (kMC == 2) in the other
            // pair too, whose multicast loads write into this pair's copy of the B tile
            umma_commit<kCG>(&empty[stage], (uint16_t)(kMC == 2 ? 0xFu : 0x3u));
            if (kb == kb1 - 1) umma_commit<kCG>(&tmem_full[acc], (uint16_t)(0x3u << (2 * pair)));
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
    const uint32_t stg = smem_u32(smem_epi) + (uint32_t)((warp - 2) * kEpiStageBytes);
    int tma_buf = 0;
    uint32_t ephase = 0;
    uint64_t* ebar = &epi_bar[warp - 2];
    int iter = 0;
    for (; sched.next<kCG>(BN, m_blk, n_blk, ncols, split); ++iter) {
      const int acc = iter & 1;
      const uint32_t acc_phase = (iter >> 1) & 1;
      const int row0 = (m_blk * kMC + (int)pair) * tile_m + (int)cta_rank * kBM + quarter * 32;  // first row of this warp's lane quarter
      if (ep.epilogue == DITB200_EPI_BIAS_GATE_RESID) {
        // The residual tile does not depend on the MMAs: pull this warp's 32 x (ncols/2) block towards L2
        // while the accumulators are still being produced, so the epilogue loads below do not pay HBM latency.
        const int chunks_ = (ncols + 31) >> 5, c0_ = (chunks_ + 1) >> 1;
        const int pc0 = sched.col0 + (half ? c0_ * 32 : 0), pn = half ? chunks_ - c0_ : c0_;
        if (row0 + lane < M) {
          const float* rp = ep.resid + (size_t)(row0 + lane) * N + pc0;
          for (int c = 0; c < pn; ++c)
            if (pc0 + c * 32 < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(rp + c * 32));
        }
      }
      if (ep.epilogue == DITB200_EPI_MUL_DGELU || ep.epilogue == DITB200_EPI_MUL_AUX) {  // same for the saved fc1 values (bf16: 64 columns per line)
        const int chunks_ = (ncols + 31) >> 5, c0_ = (chunks_ + 1) >> 1;
        const int pc0 = sched.col0 + (half ? c0_ * 32 : 0), pn = half ? chunks_ - c0_ : c0_;
        if (row0 + lane < M) {
          const __nv_bfloat16* up = ep.aux_in + (size_t)(row0 + lane) * N + pc0;
          for (int c = 0; c < pn; c += 2)
            if (pc0 + c * 32 < N) asm volatile("prefetch.global.L2 [%0];" ::"l"(up + c * 32));
        }
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tcgen05_fence_after();
      const bool add_bias = ep.bias != nullptr && split == 0;
      // the tile's 32-column chunks are split between the two warps of the lane quarter (narrow tiles too)
      // (TMA-store path: 64-column stores, so the first warp takes an even number of chunks)
      const int chunks = (ncols + 31) >> 5;
      const int chunks0 = ep.tma_store ? min(chunks, (((chunks + 1) >> 1) + 1) & ~1) : (chunks + 1) >> 1;
      const int col_off = half ? chunks0 * 32 : 0;  // first tile column of this warp
      const int nch = half ? chunks - chunks0 : chunks0;
      const uint32_t taddr0 = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * BN + col_off);
      const int colw = sched.col0 + col_off;
      constexpr int NCH = BN / 64;
#define EPI_CASE(E, O, A) epi_tile<E, O, A, NCH>(ep, taddr0, stg, lane, row0, colw, nch, M, N, add_bias)
      const bool has_aux = ep.aux_out != nullptr;
      const int omode = ep.out_bf16 ? OUT_BF16 : (ep.atomic ? OUT_ATOMIC : OUT_F32);
#define EPI_TMA(MODE, ST) \
  epi_tile_tma<MODE, NCH>(ep, ST, &tma_aux, ebar, ephase, taddr0, stg, lane, row0, colw, nch, N, add_bias, tma_buf)
      if (ep.tma_store) {  // host: bf16 outputs finished in the accumulator's row layout and stored by TMA
        if (ep.epilogue == DITB200_EPI_BIAS_GELU) EPI_TMA(1, &tma_out);
        else if (ep.epilogue == DITB200_EPI_BIAS_GELU_DAUX) {
          EPI_TMA(5, &tma_out);   // out = gelu(acc + bias), aux_out = gelu'(acc + bias): one pass, two stores per round
        } else if (ep.epilogue == DITB200_EPI_MUL_AUX) EPI_TMA(3, &tma_out);
        else EPI_TMA(0, &tma_out);
      } else if (ep.epilogue == DITB200_EPI_BIAS && !has_aux) {
        if (omode == OUT_BF16) EPI_CASE(DITB200_EPI_BIAS, OUT_BF16, 0);
        else if (omode == OUT_F32) EPI_CASE(DITB200_EPI_BIAS, OUT_F32, 0);
        else EPI_CASE(DITB200_EPI_BIAS, OUT_ATOMIC, 0);
      } else if (ep.epilogue == DITB200_EPI_BIAS_GELU && omode == OUT_BF16) {
        if (has_aux) EPI_CASE(DITB200_EPI_BIAS_GELU, OUT_BF16, 1); else EPI_CASE(DITB200_EPI_BIAS_GELU, OUT_BF16, 0);
      } else if (ep.epilogue == DITB200_EPI_BIAS_GATE_RESID && omode == OUT_F32) {
        if (has_aux) EPI_CASE(DITB200_EPI_BIAS_GATE_RESID, OUT_F32, 1); else EPI_CASE(DITB200_EPI_BIAS_GATE_RESID, OUT_F32, 0);
      } else if (ep.epilogue == DITB200_EPI_MUL_DGELU && omode == OUT_BF16 && !has_aux) {
        EPI_CASE(DITB200_EPI_MUL_DGELU, OUT_BF16, 0);
      } else if (ep.epilogue == DITB200_EPI_MUL_AUX && omode == OUT_BF16 && !has_aux) {
        EPI_CASE(DITB200_EPI_MUL_AUX, OUT_BF16, 0);
      } else if (ep.epilogue == DITB200_EPI_BIAS_GELU_DAUX && omode == OUT_BF16 && has_aux) {
        EPI_CASE(DITB200_EPI_BIAS_GELU_DAUX, OUT_BF16, 1);
      } else {
        EPI_CASE(-1, -1, -1);  // rare combinations: everything decided at run time
      }
#undef EPI_CASE
#undef EPI_TMA
      // all tcgen05.ld of this stage are complete (wait::ld above): give the stage back
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (kCG == 1) mbar_arrive(&tmem_empty[acc]); else mbar_arrive_leader(&tmem_empty[acc], lead_rank);
      }
    }
  }

  if (warp >= 2 && lane == 0) bulk_wait_group_read<0>();  // staging tiles stay alive until TMA has read them
  tcgen05_fence_before();
  if constexpr (kCG == 2) cluster_sync_all(); else __syncthreads();
  if (warp == 1) tmem_dealloc<kCG>(tmem_base, Cfg::kTmemCols);
}

// ----------------------------------------------------------------------------- host side
// 2-D bf16 tensor map over a row-major [rows, cols] matrix with a {box_rows x box_cols} box.
static int make_tmap_2d(CUtensorMap* map, const void* base, uint64_t rows, uint64_t cols,
                        uint32_t box_rows, uint32_t box_cols, CUtensorMapSwizzle swz = CU_TENSOR_MAP_SWIZZLE_128B) {
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
                   box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("gemm: cuTensorMapEncodeTiled failed with CUresult %d (rows=%llu cols=%llu box=%ux%u)",
              (int)r, (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
    return DITB200_EINVAL;
  }
  return 0;
}

// Width of the narrow last tile column (0 = N is a multiple of bn, or narrowing is not possible).  tcgen05.mma
// needs N % 16 == 0; an MN-major B operand is fetched in 64-column boxes per CTA.
static int narrow_cols(int N, int bn, int cg, int trans_w) {
  const int rem = N % bn;
  if (rem == 0) return 0;
  static const bool off = getenv("DITB200_NO_NARROW") != nullptr;  // measurement switch
  if (off) return 0;
  const int np = (rem + 15) / 16 * 16;
  if (np >= bn) return 0;
  if (trans_w && (np / cg) % 64 != 0) return 0;
  return np;
}

// Dynamic tile scheduling through cluster launch control (see TileSched) is requested per call
// (ditb200_gemm_args::dynamic_sched).  Off by default: with the GPU to itself the static longest-first schedule is
// as good and has no query latency.

// Host replay of the static tile schedules (the very TileSched the kernel runs): for every CTA pair p the units it
// would process, in order, as rows {p, m_blk, n_blk, ncols, k_part}.  Test hook: coverage and balance of the
// schedule can be checked without a GPU (tests/test_host_logic.py).  Returns the number of rows (or -needed).
extern "C" int ditb200_debug_tile_schedule(int M, int N, int tile_m, int bn, int part_cols, int split_k, int pairs,
                                           int* rows, int cap) {
  if (M <= 0 || N <= 0 || tile_m <= 0 || bn <= 0 || split_k <= 0 || pairs <= 0 || part_cols < 0 || part_cols >= bn) return 0;
  int n = 0;
  for (int p = 0; p < pairs; ++p) {
    TileSched sc;
    sc.init(M, N, tile_m, bn, part_cols, split_k, pairs, p);
    int m_blk = 0, n_blk = 0, ncols = 0, split = 0;
    while (sc.next<1>(bn, m_blk, n_blk, ncols, split)) {
      if (rows && n < cap) rows[5 * n] = p, rows[5 * n + 1] = m_blk, rows[5 * n + 2] = n_blk, rows[5 * n + 3] = ncols, rows[5 * n + 4] = split;
      ++n;
    }
  }
  return (rows && n > cap) ? -n : n;
}

static void choose_tile_fwd(int M, int N, int K, int sms, int* cg_out, int* bn_out);  // choose_tile for a K-major weight

// ------------------------------------------------------------------ explicit schedules (SchedTable)
// Cost of one pair-tile of w columns, in column units: every tile streams the whole 256-row A panel, which costs
// about as much as 104 further columns of MMA work (in-model, M = 16384, K = 4608: 192 columns 27.8 us, 256 columns
// 33.8 us; profiles/r02_bench_c3_v1.json).
static inline int tile_cost(int w) { return w + 104; }

// Max load over the pairs of the formula schedule TileSched would run for width bn (host replay, same cost model).
static int formula_load(int M, int N, int cg, int bn, int trans_w, int P) {
  const int tile_m = kBM * cg;
  const int part = narrow_cols(N, bn, cg, trans_w);
  int mx = 0;
  for (int p = 0; p < P; ++p) {
    TileSched sc;
    sc.init(M, N, tile_m, bn, part, 1, P, p);
    int m_blk = 0, n_blk = 0, ncols = 0, split = 0, load = 0;
    while (sc.next<1>(bn, m_blk, n_blk, ncols, split)) load += tile_cost(ncols);
    mx = load > mx ? load : mx;
  }
  return mx;
}

// Two tile widths, two classes of pairs.  N = f * 256 + rem with (256 + rem) / 2 = wt a legal MMA width: a row panel is
// cut either "mixed" — (f - 1) tiles of 256 and two of wt — or, where wt divides N, "uniform" into N / wt tiles of wt.
// `a` panels are cut mixed; the 256-wide tiles go round-robin to the first nA pairs, the wt-wide ones to the others.
// (a, nA) minimising the longest pair is found by enumeration; the table is used only if it beats the formula
// schedule of the width the chooser would otherwise take by 2 %.  Returns the max load (0: no table).
static int plan_table(int M, int N, int P, int formula, SchedTable* t) {
  const int bn = 256, m_tiles = (M + 255) / 256, f = N / bn, rem = N % bn;
  if (P > kTabPairs || P < 2 || rem == 0 || f < 1 || (bn + rem) % 32 != 0 || m_tiles > 65535 || N > 16 * 0x7ff) return 0;
  const int wt = (bn + rem) / 2;
  const bool uni = N % wt == 0;
  const int cb = tile_cost(bn), cm = tile_cost(wt);
  long best = -1;
  int best_a = 0, best_na = 0;
  for (int a = uni ? 0 : m_tiles; a <= m_tiles; ++a) {
    const long B = (long)a * (f - 1), Mi = (long)a * 2 + (long)(m_tiles - a) * (uni ? N / wt : 0);
    if (B + Mi > kTabCap) continue;
    for (int nA = 0; nA <= P; ++nA) {
      if ((nA == 0) != (B == 0) || (nA == P && Mi > 0)) continue;
      const long la = nA ? (B + nA - 1) / nA * cb : 0, lb = Mi ? (Mi + (P - nA) - 1) / (P - nA) * cm : 0;
      const long load = la > lb ? la : lb;
      if (best < 0 || load < best) best = load, best_a = a, best_na = nA;
    }
  }
  if (best < 0 || best * 100 > (long)formula * 98) return 0;
  // build: big tiles in (panel, slot) order to pairs [0, nA) round-robin (concurrent pairs share A panels in L2);
  // mid tiles likewise to pairs [nA, P)
  const int a = best_a, nA = best_na, nB = P - nA;
  auto entry = [](int m_blk, int col0, int w) { return ((uint32_t)m_blk << 16) | ((uint32_t)(col0 / 16) << 5) | (uint32_t)(w / 16); };
  const long B = (long)a * (f - 1), Mi = (long)a * 2 + (long)(m_tiles - a) * (uni ? N / wt : 0);
  int idx = 0;
  for (int p = 0; p < P; ++p) {
    t->start[p] = (uint16_t)idx;
    if (p < nA) {
      for (long i = p; i < B; i += nA) t->e[idx++] = entry((int)(i / (f - 1)), (int)(i % (f - 1)) * bn, bn);
    } else {
      for (long j = p - nA; j < Mi; j += nB) {
        if (j < 2L * a) t->e[idx++] = entry((int)(j / 2), (f - 1) * bn + (int)(j % 2) * wt, wt);
        else {
          const long u = j - 2L * a;
          const int per = N / wt;
          t->e[idx++] = entry(a + (int)(u / per), (int)(u % per) * wt, wt);
        }
      }
    }
  }
  t->start[P] = (uint16_t)idx;
  for (int p = P + 1; p <= kTabPairs; ++p) t->start[p] = (uint16_t)idx;
  t->use = 1;
  return (int)best;
}

// Plans are cached per shape: the search runs once.
struct TablePlan {
  int M, N, P, load;
  SchedTable tab;
};
static const SchedTable* table_for(int M, int N, int P, int bn) {
  static std::mutex mu;
  static TablePlan* plans[16];
  static int n_plans = 0;
  static const bool off = getenv("DITB200_GEMM_NO_TABLE") != nullptr;  // measurement switch
  if (off) return nullptr;
  std::lock_guard<std::mutex> lk(mu);
  for (int i = 0; i < n_plans; ++i)
    if (plans[i]->M == M && plans[i]->N == N && plans[i]->P == P) return plans[i]->load ? &plans[i]->tab : nullptr;
  if (n_plans == 16) return nullptr;
  TablePlan* pl = new TablePlan();
  memset(pl, 0, sizeof(*pl));
  pl->M = M, pl->N = N, pl->P = P;
  pl->load = plan_table(M, N, P, formula_load(M, N, 2, bn, 0, P), &pl->tab);  // host replay: once per shape
  plans[n_plans++] = pl;
  return pl->load ? &pl->tab : nullptr;
}

// Test hook (host only): the explicit schedule plan_table builds for an M x N forward GEMM on `pairs` CTA pairs, as rows
// {pair, m_blk, first column, columns}; returns the number of rows, 0 when the formula schedule is kept, -needed when
// cap is too small.  out_loads[2] = {formula max load, table max load} in tile_cost units.
extern "C" int ditb200_debug_gemm_table(int M, int N, int K, int pairs, int* rows, int cap, int* out_loads) {
  if (M <= 0 || N <= 0 || K <= 0 || pairs < 2 || pairs > kTabPairs) return 0;
  int cg = 0, bn = 0;
  choose_tile_fwd(M, N, K, pairs * 2, &cg, &bn);
  if (cg != 2) return 0;
  const int formula = formula_load(M, N, 2, bn, 0, pairs);
  SchedTable* t = new SchedTable();
  memset(t, 0, sizeof(*t));
  const int load = plan_table(M, N, pairs, formula, t);
  if (out_loads) out_loads[0] = formula, out_loads[1] = load;
  int n = 0;
  if (load) {
    for (int p = 0; p < pairs; ++p)
      for (int i = t->start[p]; i < t->start[p + 1]; ++i, ++n)
        if (rows && n < cap) {
          const uint32_t e = t->e[i];
          rows[4 * n] = p, rows[4 * n + 1] = (int)(e >> 16), rows[4 * n + 2] = (int)((e >> 5) & 0x7ffu) * 16, rows[4 * n + 3] = (int)(e & 31u) * 16;
        }
  }
  delete t;
  return (rows && n > cap) ? -n : n;
}

template <int kCG, int BN, int kMC = 1>
static int launch_cfg(const ditb200_gemm_args* a, int split_k, cudaStream_t st, const SchedTable* table = nullptr) {
  using Cfg = TcCfg<kCG, BN>;
  CUtensorMap ta, tb;
  int rc;
  // K-major operand [MN, K]: box {128 (or B rows) x 64 k}.  MN-major operand stored [K, MN]: box {64 k x 64 MN}.
  if (!a->trans_a) rc = make_tmap_2d(&ta, a->a, (uint64_t)a->M, (uint64_t)a->K, kBM, kBK);
  else rc = make_tmap_2d(&ta, a->a, (uint64_t)a->K, (uint64_t)a->M, kBK, 64);
  if (rc) return rc;
  if (!a->trans_w) rc = make_tmap_2d(&tb, a->w, (uint64_t)a->N, (uint64_t)a->K, Cfg::kBRows / kMC, kBK);
  else rc = make_tmap_2d(&tb, a->w, (uint64_t)a->K, (uint64_t)a->N, kBK, 64);
  if (rc) return rc;
  EpiParams ep;
  ep.bias = a->bias, ep.out = a->out, ep.resid = a->resid, ep.gate = a->gate;
  ep.aux_out = reinterpret_cast<__nv_bfloat16*>(a->aux_out);
  ep.aux_in = reinterpret_cast<const __nv_bfloat16*>(a->aux_in);
  ep.gate_stride = a->gate_stride, ep.rows_per_gate = a->rows_per_gate;
  ep.epilogue = a->epilogue, ep.out_bf16 = (a->out_dtype == DITB200_BF16);
  ep.atomic = (split_k > 1 || a->accumulate) ? 1 : 0;
  CUtensorMap tout;
  memset(&tout, 0, sizeof(tout));
  static const bool no_tma_store = getenv("DITB200_NO_TMA_STORE") != nullptr;  // measurement switch
  CUtensorMap taux;
  memset(&taux, 0, sizeof(taux));
  const bool plain_bf16 = (a->epilogue == DITB200_EPI_BIAS || a->epilogue == DITB200_EPI_BIAS_GELU) && !a->aux_out && !a->aux_in;
  const bool daux = a->epilogue == DITB200_EPI_BIAS_GELU_DAUX && a->aux_out && a->aux_dtype == DITB200_BF16;
  const bool mulaux = a->epilogue == DITB200_EPI_MUL_AUX && a->aux_in && !a->aux_out;
  // (a TMA load + store version of the f32 gated residual was measured 15 % slower than the register path: with
  // 8 KB of staging per warp only one residual tile can be in flight ahead of the one being combined)
  const bool tma_store = !no_tma_store && a->out_dtype == DITB200_BF16 && (plain_bf16 || daux || mulaux) &&
                         split_k == 1 && !a->accumulate;
  if (tma_store) {
    rc = make_tmap_2d(&tout, a->out, (uint64_t)a->M, (uint64_t)a->N, 32, 64, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    if (daux || mulaux) {
      rc = make_tmap_2d(&taux, daux ? a->aux_out : a->aux_in, (uint64_t)a->M, (uint64_t)a->N, 32, 64,
                        CU_TENSOR_MAP_SWIZZLE_128B);
      if (rc) return rc;
    }
  }
  ep.tma_store = tma_store ? 1 : 0;
  static bool attr_set = false;  // per instantiation; benign race (idempotent)
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tc_kernel<kCG, BN, kMC>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    if (e != cudaSuccess) return check_cuda(e, "gemm_tc smem attribute");
    attr_set = true;
  }
  if (split_k > 1 && !a->accumulate) {  // partial tiles are summed into out: start from zero
    cudaError_t e = cudaMemsetAsync(a->out, 0, (size_t)a->M * a->N * sizeof(float), st);
    if (e != cudaSuccess) return check_cuda(e, "gemm_tc split-K memset");
  }
  const int tile_m = kBM * kCG * kMC;  // rows per cluster work unit
  const int units = ((a->M + tile_m - 1) / tile_m) * ((a->N + BN - 1) / BN) * split_k;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kNumThreads);
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kCG * kMC;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#ifdef DITB200_PDL
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.numAttrs = 2;
#endif
  int clusters = num_sms() / (kCG * kMC);
  if constexpr (kMC > 1) {
    // clusters of 4 cannot tile every GPC: ask how many are co-resident (a persistent kernel must not queue any)
    static int max_clusters = 0;
    if (max_clusters == 0) {
      cfg.gridDim = dim3((unsigned)(clusters * kCG * kMC));
      int n = 0;
      cudaError_t e = cudaOccupancyMaxActiveClusters(&n, gemm_tc_kernel<kCG, BN, kMC>, &cfg);
      if (e != cudaSuccess) return check_cuda(e, "gemm_tc cluster occupancy");
      max_clusters = n > 0 ? n : 1;
    }
    if (clusters > max_clusters) clusters = max_clusters;
  }
  if (clusters > units) clusters = units;
  const int dyn = (kMC == 1 && a->dynamic_sched && units > clusters) ? 1 : 0;
  const int part = kMC > 1 ? 0 : narrow_cols(a->N, BN, kCG, a->trans_w);
  if (dyn) clusters = units;  // one cluster per unit; the running ones cancel and absorb the rest
  static const SchedTable no_table = {};
  if (table != nullptr && (dyn || kMC > 1 || split_k > 1 || clusters != num_sms() / kCG)) table = nullptr;
  cfg.gridDim = dim3((unsigned)(clusters * kCG * kMC));
  cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tc_kernel<kCG, BN, kMC>, ta, tb, tout, taux, ep, a->M, a->N, a->K,
                                     a->trans_a ? 1 : 0, a->trans_w ? 1 : 0, split_k, part, dyn, a->reverse_m ? 1 : 0,
                                     table ? *table : no_table);
  if (e != cudaSuccess) return check_cuda(e, "gemm_tc launch");
  return 0;
}

// Tile choice.  Measured on B200 (profiles/r01_gemm_notes.md): per flop the 256-wide pair tile is the most
// efficient (the 12 TB/s L2->SM fabric, not the tensor pipe, bounds the main loop, and arithmetic intensity
// grows with the tile), 192 is within a few per cent, 128 costs ~30 %.  What decides between them for a given
// shape is the length of the schedule: the chooser replays TileSched on the host for every legal width and takes
// the smallest (max columns assigned to any CTA) / efficiency.
static double sched_cost(int M, int N, int cg, int bn, int trans_w, int split_k, int sms) {
  const int tile_m = kBM * cg;
  const int m_tiles = (M + tile_m - 1) / tile_m;
  const int part = narrow_cols(N, bn, cg, trans_w);
  const int n_tiles = (N + bn - 1) / bn;
  const double eff = bn >= 256 ? 1.0 : (bn >= 192 ? 0.97 : 0.72);
  int P = sms / cg;
  const int units = m_tiles * n_tiles * split_k;
  if (P > units) P = units;
  if (split_k > 1 || part == 0) {
    const int per = (units + P - 1) / P;
    return (double)per * bn / split_k / eff;
  }
  // replay of TileSched (split_k == 1): columns assigned to every CTA
  const int F = m_tiles * (n_tiles - 1), H = m_tiles;
  const int pc = part > bn / 2 ? part : bn / 2;  // cost of a narrow tile in columns (same floor as TileSched::init)
  const int r = F % P, q = bn / pc > 1 ? bn / pc : 1;
  int load[160];
  for (int p = 0; p < P; ++p) load[p] = (F / P + (p < r ? 1 : 0)) * bn;
  int j = 0;
  for (int p = r; p < P && j < H; ++p)
    for (int t = 0; t < q && j < H; ++t, ++j) load[p] += pc;
  for (int p = 0; j < H; ++j, p = (p + 1 == P ? 0 : p + 1)) load[p] += pc;
  int mx = 0;
  for (int p = 0; p < P; ++p) mx = load[p] > mx ? load[p] : mx;
  return (double)mx / eff;
}

static void choose_tile(int M, int N, int K, int trans_w, int split_k, int sms, int* cg_out, int* bn_out) {
  const int cg = (M > kBM) ? 2 : 1;
  *cg_out = cg;
  const int cand[3] = {256, 192, 128};
  double best = 1e30;
  int best_bn = 128;
  for (int i = 0; i < 3; ++i) {
    const int bn = cand[i];
    if (trans_w && (bn / cg) % 64 != 0) continue;   // MN-major B: 64-column boxes per CTA
    if (bn > 128 && N <= bn - 64) continue;          // a tile wider than the matrix buys nothing
    double c = sched_cost(M, N, cg, bn, trans_w, split_k, sms);
    // measured (N = 1152, M = 16384): with a long k loop an exact 192-wide cover beats 256 + a narrow column
    // (K = 4608: 134.1 vs 136.4 us); with a short one the shorter schedule of 256 + narrow wins (K = 1152: 39.5 vs 41.4)
    static const bool no192 = getenv("DITB200_GEMM_NO192") != nullptr;  // measurement switch
    if (!no192 && bn == 192 && N % 192 == 0 && N % 256 != 0 && N < 2048 && K > 2048) c *= 0.8;
    if (c < best) best = c, best_bn = bn;
  }
  *bn_out = best_bn;
}

static void choose_tile_fwd(int M, int N, int K, int sms, int* cg_out, int* bn_out) { choose_tile(M, N, K, 0, 1, sms, cg_out, bn_out); }

// Test hook: the tile shape the automatic chooser picks for a GEMM on a GPU with `sms` SMs: out = {cta_group,
// tile_n, narrow last column}.  Host arithmetic only.
extern "C" int ditb200_debug_gemm_plan(int M, int N, int K, int trans_w, int split_k, int sms, int* out) {
  if (M <= 0 || N <= 0 || K <= 0 || split_k <= 0 || sms < 2 || !out) return DITB200_EINVAL;
  int cg = 0, bn = 0;
  choose_tile(M, N, K, trans_w, split_k, sms, &cg, &bn);
  out[0] = cg, out[1] = bn, out[2] = narrow_cols(N, bn, cg, trans_w);
  return 0;
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
  if (a->epilogue == DITB200_EPI_MUL_DGELU || a->epilogue == DITB200_EPI_MUL_AUX)
    DITB_REQUIRE(a->aux_in && aligned16(a->aux_in) && a->aux_dtype == DITB200_BF16, DITB200_EINVAL,
                 "gemm(tcgen05): MUL_DGELU / MUL_AUX need a 16-byte-aligned bf16 aux_in");
  if (a->epilogue == DITB200_EPI_BIAS_GELU_DAUX)
    DITB_REQUIRE(a->aux_out != nullptr, DITB200_EINVAL, "gemm(tcgen05): BIAS_GELU_DAUX needs aux_out");
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
    choose_tile(a->M, a->N, a->K, a->trans_w, split_k, num_sms(), &acg, &abn);
    if (cg == 0) cg = acg;
    if (bn == 0) bn = abn;
  }
  // Explicit two-width schedule where it beats every uniform cover (forward GEMMs with a K-major weight)
  if (a->cta_group == 0 && a->tile_n == 0 && cg == 2 && !a->trans_w && split_k == 1 && !a->accumulate && !a->dynamic_sched) {
    const int P = num_sms() / 2;
    const SchedTable* tab = table_for(a->M, a->N, P, bn);
    if (tab != nullptr) return launch_cfg<2, 256>(a, split_k, st, tab);
  }
  static const bool mc_auto = getenv("DITB200_GEMM_MC") != nullptr;  // measurement switch: multicast clusters where they fit exactly
  if (mc_auto && a->cta_group == 0 && cg == 2 && bn == 256 && !a->trans_w && a->N % 256 == 0 && a->M % 512 == 0) cg = 4;
  DITB_REQUIRE(!a->trans_w || (bn / cg) % 64 == 0, DITB200_EINVAL,
               "gemm(tcgen05): trans_w needs tile_n / cta_group to be a multiple of 64 (got %d / %d)", bn, cg);
  if (cg == 4) {  // two CTA pairs per cluster sharing the B tile by TMA multicast (K-major B only)
    DITB_REQUIRE(!a->trans_w && a->M > kBM, DITB200_EINVAL, "gemm(tcgen05): cta_group 4 needs a K-major w and M > 128");
    if (bn == 256) return launch_cfg<2, 256, 2>(a, split_k, st);
    if (bn == 128) return launch_cfg<2, 128, 2>(a, split_k, st);
    set_error("gemm(tcgen05): cta_group 4 supports tile_n 128 / 256 (got %d)", bn);
    return DITB200_EINVAL;
  }
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
