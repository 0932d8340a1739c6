// gemm_simt.cu — f32 CUDA-core GEMM: the "fp32 check mode" engine and the small-M linears
// (timestep MLP, adaLN modulation).  out = epilogue(act_in(A)[M,K] · W[N,K]ᵀ).
//
// 64x64 output tile per 256-thread CTA, 4x4 micro-tile per thread, K staged through shared
// memory in slabs of 16 (stored k-major so the inner product reads are conflict-free
// broadcasts).  Partial sums are flushed into the running accumulator every 32 k so the
// rounding error grows like sqrt(K/32) instead of sqrt(K): the check mode has to sit within
// 1e-5 of an fp32 reference through 28 blocks.
#include "common.cuh"

namespace ditb200 {

constexpr int kTM = 64, kTN = 64, kTK = 16;

struct SimtArgs {
  const float* a;
  int lda;
  const void* w;  // [N, K] f32 or bf16
  int w_bf16;
  const float* bias;
  float* out_f32;
  __nv_bfloat16* out_bf16;
  int ldo;
  const float* add;  // small_linear: out += add[m, n]
  int ldadd;
  const float* resid;  // GATE_RESID
  const float* gate;
  int gate_stride, rows_per_gate;
  int M, N, K;
  int epilogue;
  int silu_in;
};

__device__ __forceinline__ float load_w(const void* w, int w_bf16, size_t idx) {
  if (w_bf16) return __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(w)[idx]);
  return reinterpret_cast<const float*>(w)[idx];
}

__global__ void __launch_bounds__(256) gemm_simt_kernel(const SimtArgs p) {
  __shared__ float As[kTK][kTM + 4];
  __shared__ float Ws[kTK][kTN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 4x4
  const int m0 = blockIdx.y * kTM, n0 = blockIdx.x * kTN;
  float acc[4][4], part[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f, part[i][j] = 0.f;

  // loader mapping: 256 threads load a 64 x 16 slab: row = tid / 4, k-group = (tid % 4) * 4
  const int lrow = tid >> 2, lk = (tid & 3) * 4;
  for (int k0 = 0; k0 < p.K; k0 += kTK) {
    {
      const int m = m0 + lrow;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + lk + j;
        float v = 0.f;
        if (m < p.M && k < p.K) {
          v = p.a[(size_t)m * p.lda + k];
          if (p.silu_in) v = silu_acc(v);
        }
        As[lk + j][lrow] = v;
      }
      const int n = n0 + lrow;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = k0 + lk + j;
        float v = 0.f;
        if (n < p.N && k < p.K) v = load_w(p.w, p.w_bf16, (size_t)n * p.K + k);
        Ws[lk + j][lrow] = v;
      }
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < kTK; ++kk) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[kk][tx * 4]);
      const float av[4] = {a4.x, a4.y, a4.z, a4.w};
      const float wv[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) part[i][j] = fmaf(av[i], wv[j], part[i][j]);
    }
    __syncthreads();
    if (((k0 / kTK) & 1) == 1 || k0 + kTK >= p.K) {
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] += part[i][j], part[i][j] = 0.f;
    }
  }

#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= p.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= p.N) continue;
      float v = acc[i][j];
      if (p.bias) v += p.bias[n];
      switch (p.epilogue) {
        case DITB200_EPI_BIAS_GELU:
          v = gelu_tanh_f(v);
          break;
        case DITB200_EPI_BIAS_SILU:
          v = silu_acc(v);
          break;
        case DITB200_EPI_BIAS_GATE_RESID: {
          const float g = p.gate[(size_t)(m / p.rows_per_gate) * p.gate_stride + n];
          v = p.resid[(size_t)m * p.ldo + n] + g * v;
        } break;
        default:
          break;
      }
      if (p.add) v += p.add[(size_t)m * p.ldadd + n];
      if (p.out_bf16)
        p.out_bf16[(size_t)m * p.ldo + n] = __float2bfloat16_rn(v);
      else
        p.out_f32[(size_t)m * p.ldo + n] = v;
    }
  }
}

// Small-M variant (M <= 64: the timestep MLP, adaLN in the fp32 check mode): the 64 x 64 tiling above leaves
// N / 64 CTAs on a 148-SM machine.  Here a warp owns ONE output column for all M rows (lane = rows l, l + 32),
// a CTA of 8 warps owns 8 columns, and A is staged through shared memory in 128-wide k slabs (row stride 129:
// conflict-free column reads).  f32 FMAs in the same k order as above (partials flushed every 32 k).
constexpr int kSmK = 128;
__global__ void __launch_bounds__(256) gemm_simt_small_m_kernel(const SimtArgs p) {
  __shared__ float As[64][kSmK + 1];
  __shared__ float Ws[8][kSmK];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n = blockIdx.x * 8 + warp;
  float acc0 = 0.f, acc1 = 0.f;
  // The next k slab travels global -> registers while the current one is multiplied out of shared memory: the
  // kernel is a chain of K / 128 short slabs, and without the overlap each one pays a full memory latency.
  float ra[64 * kSmK / 256], rw[8 * kSmK / 256];
  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 64 * kSmK / 256; ++i) {
      const int idx = threadIdx.x + 256 * i;
      const int r = idx / kSmK, k = idx - r * kSmK;
      ra[i] = (r < p.M && k0 + k < p.K) ? p.a[(size_t)r * p.lda + k0 + k] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8 * kSmK / 256; ++i) {
      const int idx = threadIdx.x + 256 * i;
      const int c = idx / kSmK, k = idx - c * kSmK;
      const int nn = blockIdx.x * 8 + c;
      rw[i] = (nn < p.N && k0 + k < p.K) ? load_w(p.w, p.w_bf16, (size_t)nn * p.K + k0 + k) : 0.f;
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < p.K; k0 += kSmK) {
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 64 * kSmK / 256; ++i) {
      const int idx = threadIdx.x + 256 * i;
      const int r = idx / kSmK, k = idx - r * kSmK;
      As[r][k] = (p.silu_in && r < p.M && k0 + k < p.K) ? silu_acc(ra[i]) : ra[i];
    }
#pragma unroll
    for (int i = 0; i < 8 * kSmK / 256; ++i) {
      const int idx = threadIdx.x + 256 * i;
      Ws[idx / kSmK][idx % kSmK] = rw[i];
    }
    __syncthreads();
    if (k0 + kSmK < p.K) fetch(k0 + kSmK);
#pragma unroll
    for (int kc = 0; kc < kSmK; kc += 32) {
      float p0 = 0.f, p1 = 0.f;
#pragma unroll
      for (int kk = 0; kk < 32; ++kk) {
        const float wv = Ws[warp][kc + kk];
        p0 = fmaf(As[lane][kc + kk], wv, p0);
        p1 = fmaf(As[lane + 32][kc + kk], wv, p1);
      }
      acc0 += p0, acc1 += p1;
    }
  }
  if (n >= p.N) return;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int m = lane + 32 * h;
    if (m >= p.M) continue;
    float v = h ? acc1 : acc0;
    if (p.bias) v += p.bias[n];
    if (p.epilogue == DITB200_EPI_BIAS_GELU) v = gelu_tanh_f(v);
    else if (p.epilogue == DITB200_EPI_BIAS_SILU) v = silu_acc(v);
    else if (p.epilogue == DITB200_EPI_BIAS_GATE_RESID)
      v = p.resid[(size_t)m * p.ldo + n] + p.gate[(size_t)(m / p.rows_per_gate) * p.gate_stride + n] * v;
    if (p.add) v += p.add[(size_t)m * p.ldadd + n];
    if (p.out_bf16) p.out_bf16[(size_t)m * p.ldo + n] = __float2bfloat16_rn(v);
    else p.out_f32[(size_t)m * p.ldo + n] = v;
  }
}

int launch_gemm_simt(const SimtArgs& p, cudaStream_t st) {
  if (p.M <= 64) {
    gemm_simt_small_m_kernel<<<(p.N + 7) / 8, 256, 0, st>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return check_cuda(e, "gemm_simt(small M)");
    return 0;
  }
  dim3 grid((p.N + kTN - 1) / kTN, (p.M + kTM - 1) / kTM);
  gemm_simt_kernel<<<grid, 256, 0, st>>>(p);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return check_cuda(e, "gemm_simt");
  return 0;
}

// defined in gemm_tc.cu
int launch_gemm_tcgen05(const ditb200_gemm_args* a, cudaStream_t st);

}  // namespace ditb200

using namespace ditb200;

extern "C" int ditb200_small_linear(const float* a, int lda, const void* w, int w_dtype,
                                    const float* bias, const float* add, int ldadd, float* out,
                                    int ldo, int M, int N, int K, int silu_in, int silu_out,
                                    void* stream) {
  DITB_REQUIRE(a && w && out, DITB200_EINVAL, "small_linear: null pointer");
  DITB_REQUIRE(M > 0 && M <= 1024 && N > 0 && K > 0, DITB200_EINVAL,
               "small_linear: bad shape M=%d N=%d K=%d (M <= 1024)", M, N, K);
  DITB_REQUIRE(lda >= K && ldo >= N && (!add || ldadd >= N), DITB200_EINVAL,
               "small_linear: bad leading dimension");
  DITB_REQUIRE(w_dtype == DITB200_F32 || w_dtype == DITB200_BF16, DITB200_EINVAL,
               "small_linear: bad w_dtype");
  SimtArgs p{};
  p.a = a, p.lda = lda, p.w = w, p.w_bf16 = (w_dtype == DITB200_BF16), p.bias = bias;
  p.out_f32 = out, p.out_bf16 = nullptr, p.ldo = ldo, p.add = add, p.ldadd = ldadd;
  p.M = M, p.N = N, p.K = K;
  p.epilogue = silu_out ? DITB200_EPI_BIAS_SILU : DITB200_EPI_BIAS;
  p.silu_in = silu_in;
  return launch_gemm_simt(p, (cudaStream_t)stream);
}

extern "C" int ditb200_gemm(const ditb200_gemm_args* a, void* stream) {
  DITB_REQUIRE(a != nullptr, DITB200_EINVAL, "gemm: null args");
  DITB_REQUIRE(a->a && a->w && a->out, DITB200_EINVAL, "gemm: null tensor");
  DITB_REQUIRE(a->M > 0 && a->N > 0 && a->K > 0, DITB200_EINVAL, "gemm: bad shape %d %d %d", a->M,
               a->N, a->K);
  DITB_REQUIRE(a->epilogue >= 0 && a->epilogue <= DITB200_EPI_MUL_AUX, DITB200_EINVAL,
               "gemm: bad epilogue %d", a->epilogue);
  DITB_REQUIRE(a->out_dtype == DITB200_F32 || a->out_dtype == DITB200_BF16, DITB200_EINVAL,
               "gemm: bad out_dtype");
  if (a->epilogue == DITB200_EPI_BIAS_GATE_RESID) {
    DITB_REQUIRE(a->resid && a->gate && a->rows_per_gate > 0 && a->gate_stride >= a->N,
                 DITB200_EINVAL, "gemm: GATE_RESID needs resid, gate, rows_per_gate, gate_stride");
    DITB_REQUIRE(a->out_dtype == DITB200_F32, DITB200_EINVAL,
                 "gemm: GATE_RESID writes the f32 residual stream");
  }
  if (a->engine == DITB200_GEMM_TCGEN05) return launch_gemm_tcgen05(a, (cudaStream_t)stream);
  DITB_REQUIRE(a->engine == DITB200_GEMM_FP32, DITB200_EINVAL, "gemm: bad engine %d", a->engine);
  DITB_REQUIRE(!a->trans_a && !a->trans_w && !a->aux_out && !a->aux_in && !a->accumulate && a->split_k <= 1 &&
                   a->epilogue < DITB200_EPI_MUL_DGELU,
               DITB200_EINVAL, "gemm: the fp32 check engine runs the forward GEMMs only (no trans/aux/split_k)");
  SimtArgs p{};
  p.a = reinterpret_cast<const float*>(a->a), p.lda = a->K;
  p.w = a->w, p.w_bf16 = 0, p.bias = a->bias;
  if (a->out_dtype == DITB200_BF16)
    p.out_bf16 = reinterpret_cast<__nv_bfloat16*>(a->out);
  else
    p.out_f32 = reinterpret_cast<float*>(a->out);
  p.ldo = a->N;
  p.resid = a->resid, p.gate = a->gate, p.gate_stride = a->gate_stride;
  p.rows_per_gate = a->rows_per_gate;
  p.M = a->M, p.N = a->N, p.K = a->K, p.epilogue = a->epilogue, p.silu_in = 0;
  return launch_gemm_simt(p, (cudaStream_t)stream);
}
