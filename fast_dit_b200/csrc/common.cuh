// common.cuh — shared device helpers and the host-side error convention of libditb200.
// sm_100a only.  All PTX wrappers here are hand-written against the PTX ISA
// (mbarrier, cp.async.bulk.tensor, tcgen05.*); nothing is pulled from CUTLASS.
#pragma once

#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/ditb200.h"

namespace ditb200 {

// ------------------------------------------------------------------ host side
void set_error(const char* fmt, ...);
int check_cuda(cudaError_t e, const char* what);
int num_sms();
bool is_initialised();
// cuTensorMapEncodeTiled resolved through cudaGetDriverEntryPoint (no libcuda link dependency).
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled_fn();

#define DITB_REQUIRE(cond, code, ...)      \
  do {                                     \
    if (!(cond)) {                         \
      ::ditb200::set_error(__VA_ARGS__);   \
      return (code);                       \
    }                                      \
  } while (0)

#define DITB_LAUNCH_CHECK(what)                                   \
  do {                                                            \
    cudaError_t e__ = cudaGetLastError();                         \
    if (e__ != cudaSuccess) return ::ditb200::check_cuda(e__, what); \
  } while (0)

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Programmatic dependent launch (on by default since round 2: +1.6 % on the C3 sampling step, same-box A/B;
// -DDITB200_NO_PDL builds the variant without it: the macros vanish and the launches below are plain <<<>>>).
// A kernel launched with the stream-serialization attribute may become resident
// while its predecessor in the stream is still running; DITB_PDL_WAIT() blocks until that predecessor has completed
// and its writes are visible, so it precedes every global access of the kernel (weights could go before it; kept
// simple).  DITB_PDL_TRIGGER() lets the successor start launching; persistent
// one-CTA-per-SM kernels call it at once (a successor cannot co-reside with them anyway and takes each SM as it is
// vacated), multi-wave kernels never do (their exit is the trigger), so that an early successor cannot take their
// registers.
#if !defined(DITB200_NO_PDL) && !defined(DITB200_PDL)
#define DITB200_PDL 1
#endif
#ifdef DITB200_PDL
#define DITB_PDL_WAIT() asm volatile("griddepcontrol.wait;" ::: "memory")
#define DITB_PDL_TRIGGER() asm volatile("griddepcontrol.launch_dependents;" ::: "memory")
#else
#define DITB_PDL_WAIT() ((void)0)
#define DITB_PDL_TRIGGER() ((void)0)
#endif

// DITB_KLAUNCH((kernel<...>), grid, block, smem, stream, args...): plain launch in the default build; with the
// stream-serialization attribute in the PDL variant (only for kernels that execute DITB_PDL_WAIT()).
#ifdef DITB200_PDL
template <typename... KArgs, typename... Args>
static inline void launch_dependent(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at, cfg.numAttrs = 1;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);  // errors surface in DITB_LAUNCH_CHECK
}
#define DITB_KLAUNCH(kern, g, b, s, st, ...) launch_dependent(kern, g, b, s, st, __VA_ARGS__)
#else
#define DITB_KLAUNCH(kern, g, b, s, st, ...) kern<<<g, b, s, st>>>(__VA_ARGS__)
#endif

// ---------------------------------------------------------------- device side
#ifdef __CUDACC__

__device__ __forceinline__ float silu_f(float v) { return v / (1.0f + __expf(-v)); }
// SiLU with the accurate expf (used where we must track torch closely)
__device__ __forceinline__ float silu_acc(float v) { return v / (1.0f + expf(-v)); }

__device__ __forceinline__ float gelu_tanh_f(float u) {
  // gelu_tanh(u) = 0.5 u (1 + tanh(sqrt(2/pi) (u + 0.044715 u^3)))   (nn.GELU(approximate="tanh"))
  const float k0 = 0.7978845608028654f, k1 = 0.044715f;
  float inner = k0 * (u + k1 * u * u * u);
  return 0.5f * u * (1.0f + tanhf(inner));
}
// fast variant for the tensor-core epilogue: the hardware tanh (one MUFU op, |rel err| ~ 2^-11, below the
// bf16 rounding of the stored activation) and five FMA-pipe ops
__device__ __forceinline__ float tanh_approx(float z) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(z));
  return t;
}
__device__ __forceinline__ float gelu_tanh_fast(float u) {
  const float k0 = 0.7978845608028654f, k0k1 = 0.7978845608028654f * 0.044715f;
  const float z = u * fmaf(k0k1, u * u, k0);
  const float hu = 0.5f * u;
  return fmaf(hu, tanh_approx(z), hu);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// streaming (read-once) 128-bit global load / store
__device__ __forceinline__ float4 ldg_stream_f4(const float4* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}

// Pinned loads: `asm volatile` keeps them where they are written relative to the other volatile asm (TMEM
// loads / waits), so a batch issued ahead of a wait stays a batch instead of being sunk to its first use.
__device__ __forceinline__ float4 ldg_pinned_f4(const float* p) {
  float4 r;
  asm volatile("ld.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}
__device__ __forceinline__ uint2 ldg_pinned_u2(const void* p) {
  uint2 r;
  asm volatile("ld.global.nc.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p) : "memory");
  return r;
}

// 128-bit shared-memory accesses by 32-bit shared-window address.  Pointers derived from the manually aligned
// dynamic-smem base lose their address space (the compiler emits generic LD/ST, which queue behind global
// traffic on the long scoreboard); these stay LDS/STS.
__device__ __forceinline__ void sts128(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
__device__ __forceinline__ float4 lds128_f(uint32_t addr) {
  float4 r;
  asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "r"(addr) : "memory");
  return r;
}
__device__ __forceinline__ uint4 lds128_u(uint32_t addr) {
  uint4 r;
  asm volatile("ld.shared.v4.b32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(addr) : "memory");
  return r;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// shared::cluster address of `p` (an address in this CTA's shared memory) as seen in CTA `rank`
__device__ __forceinline__ uint32_t mapa_u32(const void* p, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(p)), "r"(rank));
  return r;
}
// arrive on the barrier at the same smem offset in CTA `leader_rank` (the pair's leader) of the cluster
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar, uint32_t leader_rank = 0) {
  // default semantics on purpose: an explicit .release.cluster costs a MEMBAR.ALL.CTA + ERRBAR
  // (waits for every outstanding global store of the thread) in front of each arrive
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(mapa_u32(bar, leader_rank)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

// ------------------------------------------------------------------ clusters
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::
                   : "memory");
}

// ----------------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2-D tile load, signals `bar` (own CTA) with complete_tx
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, "
      "%4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 2-D tile load issued by either CTA of a pair; completion bytes go to the LEADER CTA's barrier
__device__ __forceinline__ void tma_load_2d_pair(const CUtensorMap* m, uint64_t* bar, void* dst,
                                                 int c0, int c1, uint32_t leader_rank = 0) {
  const uint32_t bar_addr = mapa_u32(bar, leader_rank);
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1)
      : "memory");
}

// 2-D tile load multicast to the CTAs in `mask` (same smem offset in each); in every destination CTA the
// completion bytes go to the barrier at `bar`'s offset in that CTA's PAIR LEADER (peer bit of the shared-window
// address cleared, as the cta_group::2 form defines it)
__device__ __forceinline__ void tma_load_2d_pair_mc(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                                    uint16_t mask) {
  const uint32_t bar_addr = smem_u32(bar) & 0xFEFFFFFFu;
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// 2-D tile store shared -> global (bulk async-group completion; rows / columns outside the tensor are clipped)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* m, uint32_t smem_src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(smem_src), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_group_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}

// ------------------------------------------------------------------- tcgen05
template <int kCtaGroup>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  if constexpr (kCtaGroup == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(smem_dst)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int kCtaGroup>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}

// D[tmem] (+)= A[smem desc] * B[smem desc], bf16 inputs, f32 accumulate
template <int kCtaGroup>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  if constexpr (kCtaGroup == 1)
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all previously issued MMAs of this thread arrive on `bar` when done (implies fence::before_thread_sync)
template <int kCtaGroup>
__device__ __forceinline__ void umma_commit(uint64_t* bar, uint16_t mask = 3) {
  if constexpr (kCtaGroup == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
  else
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 "
        "[%0], %1;" ::"r"(smem_u32(bar)),
        "h"(mask)
        : "memory");
}

// 32 lanes x 32 columns of f32 accumulators: thread i of the warp gets TMEM lane (base_lane + i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]),
        "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
        "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
        "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
        "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major operand tile in shared memory, rows of 64 bf16 (128 B) laid out by TMA with
// CU_TENSOR_MAP_SWIZZLE_128B: 8-row x 128-byte swizzle atoms stacked every 1024 B.
// Descriptor: start address >> 4 in [0,14), LBO (unused for one atom along K) in [16,30),
// SBO = 1024 B >> 4 in [32,46), descriptor version 1 in [46,48), layout SWIZZLE_128B (=2) in [61,64).
__device__ __forceinline__ uint64_t umma_smem_desc_k128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(1024u >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// instruction descriptor for kind::f16: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1),
// both operands K-major (bits 15,16 = 0), N>>3 in [17,23), M>>4 in [24,29).
__host__ __device__ constexpr uint32_t umma_idesc_bf16(uint32_t M, uint32_t N) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((N >> 3) << 17) | ((M >> 4) << 24);
}

#endif  // __CUDACC__

}  // namespace ditb200
