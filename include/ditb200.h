/*
 * ditb200.h — C-ABI of libditb200.so, the B200 (sm_100a) denoiser hot path.
 *
 * The reference (alexandor91/fast-DiT) has no FFI: its boundary is Python duck
 * typing over torch ops.  Each entry point below replaces the torch/timm
 * library call(s) named in its comment (file:line relative to the reference
 * tree).  The Python mirror of the reference interface lives in
 * fast_dit_b200/{models.py,diffusion/}; INTEGRATION.md shows the ctypes stub a
 * reference maintainer would add.
 *
 * Conventions
 *   - Every pointer is a DEVICE pointer owned by the caller (PyTorch's caching
 *     allocator).  The library allocates nothing after ditb200_init() and keeps
 *     no pointer past the call.
 *   - Every entry point is asynchronous on `stream` (a cudaStream_t passed as
 *     void*), re-entrant, and safe to call from PyTorch's autograd thread.
 *   - Return value: 0 = OK, negative = argument/shape error (DITB200_E*),
 *     positive = cudaError_t.  ditb200_last_error() returns the message of the
 *     last failure on the calling thread.
 *   - dtype codes: DITB200_F32 = 0, DITB200_BF16 = 1.
 *   - "tokens" are rows of the [M = B*T, D] activation matrix, row-major.
 */
#ifndef DITB200_H_
#define DITB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DITB200_ABI_VERSION 3

#define DITB200_F32 0
#define DITB200_BF16 1

#define DITB200_EINVAL (-1)   /* bad argument / unsupported shape */
#define DITB200_ENOINIT (-2)  /* ditb200_init() not called for this device */
#define DITB200_EALIGN (-3)   /* pointer or stride not aligned as required */

/* GEMM epilogues (what is fused after acc = A·Wᵀ) */
#define DITB200_EPI_BIAS 0            /* out = acc + bias                              */
#define DITB200_EPI_BIAS_GELU 1       /* out = gelu_tanh(acc + bias)                   */
#define DITB200_EPI_BIAS_GATE_RESID 2 /* out_f32 = resid + gate[row/T] * (acc + bias)  */
#define DITB200_EPI_BIAS_SILU 3       /* out = silu(acc + bias)                        */
#define DITB200_EPI_MUL_DGELU 4       /* out = (acc + bias) * gelu_tanh'(aux_in)  (backward of fc1's GELU) */
#define DITB200_EPI_BIAS_GELU_DAUX 5  /* out = gelu_tanh(acc + bias) and aux_out = gelu_tanh'(acc + bias): the forward
                                         leaves the derivative behind, so the backward multiplies instead of
                                         re-evaluating tanh in its epilogue */
#define DITB200_EPI_MUL_AUX 6         /* out = (acc + bias) * aux_in  (fc2 data gradient times the saved gelu') */

/* GEMM engines */
#define DITB200_GEMM_TCGEN05 0 /* bf16 operands, tcgen05.mma + TMEM accumulators, TMA-fed */
#define DITB200_GEMM_FP32 1    /* fp32 operands, CUDA-core FFMA (the "fp32 check mode")    */

/* diffusion model-output conventions (gaussian_diffusion.py:25-45) */
#define DITB200_MEAN_EPSILON 0
#define DITB200_MEAN_START_X 1
#define DITB200_VAR_LEARNED_RANGE 0
#define DITB200_VAR_LEARNED 1
#define DITB200_VAR_FIXED 2 /* FIXED_SMALL / FIXED_LARGE: log-variance read from a table */

/* update rule of ditb200_p_sample_step */
#define DITB200_SAMPLER_ANCESTRAL 0 /* p_sample  (gaussian_diffusion.py:376-417) */
#define DITB200_SAMPLER_DDIM 1      /* ddim_sample (gaussian_diffusion.py:513-560) */
#define DITB200_SAMPLER_DDIM_REVERSE 2 /* ddim_reverse_sample, x_t -> x_{t+1} (gaussian_diffusion.py:562-598) */

int ditb200_abi_version(void);
/* One-off per process+device: resolves the driver's tensor-map encoder, reads
 * the SM count, raises dynamic shared-memory limits.  Idempotent. */
int ditb200_init(int device);
const char* ditb200_last_error(void);
int ditb200_sm_count(void);
/* Test hook (host only, no GPU needed): replays the GEMM kernel's static tile schedule for `pairs` CTA pairs and
 * writes, for every pair in order, the units it would process as rows {pair, m_blk, n_blk, ncols, k_part} into
 * rows[5 * cap].  part_cols = width of the narrow last tile column (0: none).  Returns the number of rows, or -needed when cap
 * is too small. */
int ditb200_debug_tile_schedule(int M, int N, int tile_m, int bn, int part_cols, int split_k, int pairs,
                                int* rows, int cap);
/* Test hook (host only): the tile the automatic chooser picks for an M x N x K GEMM on a GPU with `sms` SMs:
 * out[3] = {cta_group, tile_n, width of the narrow last tile column (0: none)}. */
int ditb200_debug_gemm_plan(int M, int N, int K, int trans_w, int split_k, int sms, int* out);
/* Test hook (host only): the explicit two-width tile schedule the GEMM uses where it balances better than a uniform
 * cover (N = 1152 on 74 CTA pairs: 256|256|256|192|192 panels).  rows[4 * cap] receives {pair, m_blk, first column,
 * columns} per tile in execution order; out_loads[2] = {longest pair under the formula schedule, under the table}
 * in cost units.  Returns the number of tiles, 0 when the formula schedule is kept, -needed when cap is too small. */
int ditb200_debug_gemm_table(int M, int N, int K, int pairs, int* rows, int cap, int* out_loads);
/* Test hook (host only): kernel family that serves bf16 attention for T tokens per image and head dim hd:
 * 2 = tcgen05, K/V streamed in 128-key blocks (forward, T = 512 / 768 / 1024 ...), 1 = tcgen05 whole-row kernels
 * (T = 128 / 256 forward, T = 256 backward), 0 = mma.sync flash kernels (every other T), -1 = unsupported head dim. */
int ditb200_debug_attention_path(int T, int hd, int backward);

/* ---------------------------------------------------------------- embedders */

/* PatchEmbed + pos_embed: Conv2d(C→D, k=s=p) ≡ per-patch GEMM, flatten NCHW →
 * [B,T,D], add the frozen sin-cos table.
 * Replaces timm PatchEmbed.forward + `+ self.pos_embed`
 * (train_options/models_original.py:169,240).
 * x[B,C,H,W] f32; w[D, C*p*p] f32 (conv weight flattened); bias[D]; pos[T,D];
 * out[B*T, D] f32.  round_bf16 != 0 rounds operands and the conv result to
 * bf16 first (what torch.autocast does to conv2d). */
int ditb200_patch_embed(const float* x, const float* w, const float* bias, const float* pos,
                        float* out, int B, int C, int H, int W, int p, int D, int round_bf16,
                        void* stream);

/* Sinusoidal timestep features, cos half first.
 * Replaces TimestepEmbedder.timestep_embedding (models_original.py:40-59).
 * t[B] int64 (t_is_float == 0) or f32 (t_is_float != 0: the reference embeds fractional timesteps too, it
 * computes t[:, None].float() * freqs) → out[B, dim] f32. */
int ditb200_timestep_embedding(const void* t, int t_is_float, float* out, int B, int dim, float max_period,
                               void* stream);

/* Small-M linear: out[m, n] = act_out( sum_k act_in(a[m,k]) * w[n,k] + bias[n] ) (+ add[m,n]).
 * Replaces nn.Linear / nn.SiLU in TimestepEmbedder.mlp (models_original.py:33-37),
 * DiTBlock.adaLN_modulation (:113-116) and FinalLayer.adaLN_modulation (:134-137).
 * a[M,K] f32 (row stride lda); w[N,K] f32 or bf16 (w_dtype); bias[N] f32 or NULL;
 * add[M,N] f32 or NULL (row stride ldadd); out[M,N] f32 (row stride ldo).
 * silu_in / silu_out: apply SiLU to the input / to the result.  M <= 1024. */
int ditb200_small_linear(const float* a, int lda, const void* w, int w_dtype, const float* bias,
                         const float* add, int ldadd, float* out, int ldo, int M, int N, int K,
                         int silu_in, int silu_out, void* stream);

/* Label embedding gather: out[b,:] = table[y[b],:] (+ add[b,:]).
 * Replaces LabelEmbedder.forward's nn.Embedding and `c = t + y`
 * (models_original.py:93,243).  Label dropout (token_drop, :79-87) stays on
 * the host because it consumes torch's RNG stream. */
int ditb200_label_embed(const int64_t* y, const float* table, const float* add, float* out, int B,
                        int D, int num_rows, void* stream);

/* ------------------------------------------------------- LayerNorm + modulate */

/* out[b,t,:] = LN(x[b,t,:]) * (1 + scale[b,:]) + shift[b,:], LN without affine,
 * biased variance, statistics in f32.
 * Replaces nn.LayerNorm(elementwise_affine=False, eps=1e-6) + modulate()
 * (models_original.py:19-20,107,109,120-121).
 * x[B*T, D] f32; shift/scale point at [B, D] slices with row stride mod_stride
 * (floats); out[B*T, D] of out_dtype.  If stats != NULL, writes mean and rstd
 * per row to stats[2*row], stats[2*row+1] (saved for backward).  reverse != 0: rows are visited last-first
 * (L2 reuse between neighbouring kernels of a chain, see ditb200_gemm_args.reverse_m); same results. */
int ditb200_ln_modulate(const float* x, const float* shift, const float* scale, int mod_stride,
                        void* out, int out_dtype, float* stats, int B, int T, int D, float eps,
                        int reverse, void* stream);

/* Gated residual update fused in front of ditb200_ln_modulate (the training forward keeps the branch output y
 * for backward, so the GEMM that produces it needs no residual traffic of its own):
 *   x_out[b,t,:] = x[b,t,:] + gate[b,:] * y[b,t,:]                      (models_original.py:120-121)
 *   out[b,t,:]   = LN(x_out[b,t,:]) * (1 + scale[b,:]) + shift[b,:]     (out == NULL: only x_out)
 * x, x_out f32 [B*T, D] (may alias); y bf16 [B*T, D]; gate/shift/scale [B, D] slices with row stride mod_stride;
 * out bf16 or f32; stats as in ditb200_ln_modulate.  D must be 384, 768, 1024 or 1152. */
int ditb200_ln_modulate_resid(const float* x, const void* y, const float* gate, const float* shift,
                              const float* scale, int mod_stride, float* x_out, void* out, int out_dtype,
                              float* stats, int B, int T, int D, float eps, int reverse, void* stream);

/* Backward of ditb200_ln_modulate with respect to x, shift and scale.
 * dh[B*T, D] (dh_dtype) is the gradient of the modulated output; x, scale, stats as in the forward.
 * dx[B*T, D] f32 receives the gradient wrt x (added to dx when accumulate != 0: the residual stream's
 * gradient flows through both the branch and the skip connection).  dshift/dscale [B, D] f32 (row stride
 * dmod_stride) are ACCUMULATED with atomics: the caller zeroes them once per step. */
int ditb200_ln_modulate_bwd(const void* dh, int dh_dtype, const float* x, const float* scale, int mod_stride,
                            const float* stats, float* dx, int accumulate, float* dshift, float* dscale,
                            int dmod_stride, int B, int T, int D, void* stream);

/* ditb200_ln_modulate_bwd followed, in the same pass over the rows, by ditb200_gate_resid_bwd of the branch that
 * joined the residual stream in front of this LayerNorm (models_original.py:120-121 read backwards: the gradient
 * of the stream this call completes, dx, is what that branch's backward consumes):
 *   dx (+)= LayerNorm+modulate backward as above;  dshift / dscale += ...
 *   y != NULL:  dy[B*T, D] (bf16) = dx * gate[b];  dgate[b, :] += sum_t dx[b,t,:] * y[b,t,:];  dbias[:] += sum dy
 * y == NULL: only the LayerNorm part.  y, dy bf16 [B*T, D]; the reductions are atomics onto caller-zeroed f32 rows.
 * T % 4 == 0 and D in {384, 768, 1024, 1152}; other shapes: call the two entries above / below one after the other. */
int ditb200_ln_modulate_bwd_gate(const void* dh, int dh_dtype, const float* x, const float* scale, int mod_stride,
                                 const float* stats, float* dx, int accumulate, float* dshift, float* dscale,
                                 int dmod_stride, const void* y, const float* gate, int gate_stride, void* dy,
                                 float* dgate, int dgate_stride, float* dbias, int B, int T, int D, void* stream);

/* Backward of the gated-residual epilogue x_out = x + gate[b] * y (models_original.py:120-121):
 * dy[B*T, D] (dy_dtype == y_dtype) = dx_out * gate[b];  dgate[b, :] += sum_t dx_out[b,t,:] * y[b,t,:];
 * dbias[:] += sum over all tokens of dy (the bias gradient of the Linear that produced y; may be NULL).
 * dgate / dbias are ACCUMULATED with atomics: the caller zeroes them once per step.
 * dx_out itself is also the gradient of the skip path and is left untouched. */
int ditb200_gate_resid_bwd(const float* dx_out, const void* y, int y_dtype, const float* gate, int gate_stride,
                           void* dy, int dy_dtype, float* dgate, int dgate_stride, float* dbias, int B, int T,
                           int D, void* stream);

/* out[C] f32 (+)= column sums of in[R, C] (dtype): bias gradients of the QKV / fc1 / embedding Linears.
 * C % 4 == 0.  accumulate == 0 zeroes out first. */
int ditb200_colsum(const void* in, int dtype, float* out, int accumulate, int R, int C, void* stream);

/* Weight and bias gradient of an adaLN modulation Linear (models_original.py:113-116, 128-131 backward):
 *   dw[r, c] = sum_b dmod[b, r] * sc[b, c];   dbias[r] = sum_b dmod[b, r]   (dbias may be NULL)
 * dmod f32 [N, R] with row stride dmod_stride (the buffer the LayerNorm / gate backward kernels reduce into),
 * sc bf16 [N, D] = silu(c) as the forward read it, dw f32 [R, D] and dbias f32 [R] are OVERWRITTEN.
 * The contraction runs over the batch only: an outer-product kernel bound by the store of dw.
 * R % 64 == 0 and D % 128 == 0. */
int ditb200_adaln_wgrad(const float* dmod, int dmod_stride, const void* sc, float* dw, float* dbias, int N, int R,
                        int D, void* stream);

/* dtable[y[b], :] += dc[b, :] (f32 atomics): backward of the label-embedding gather. */
int ditb200_label_embed_bwd(const float* dc, const int64_t* y, float* dtable, int B, int D, int num_rows,
                            void* stream);

/* patches[B*T, C*p*p] (out_dtype: bf16 or f32) = im2col of x[B,C,H,W] f32 in the conv-weight order (c, i, j): the
 * token-major operand of the patch-embed weight gradient (backward of timm PatchEmbed.proj), and the A operand
 * of a wide patch embedding run as a GEMM (the fork's dino_embedder, C = 768: /root/reference/models.py:652,744). */
int ditb200_patchify(const float* x, void* patches, int out_dtype, int B, int C, int H, int W, int p, void* stream);

/* dz[B*T, p*p*Cout] bf16 = inverse of DiT.unpatchify (models_original.py:218-231) applied to the
 * gradient dout[B, Cout, Hp*p, Hp*p] f32 of the model output; T = Hp*Hp. */
int ditb200_unpatchify_bwd(const float* dout, void* dz, int B, int Cout, int Hp, int p, void* stream);

/* out (+)= dact * silu'(pre), elementwise f32: backward of the nn.SiLU in TimestepEmbedder.mlp and in
 * front of every adaLN_modulation Linear (models_original.py:35,114,135). */
int ditb200_silu_bwd(const float* dact, const float* pre, float* out, int accumulate, size_t n, void* stream);

/* ----------------------------------------------------------------- GEMMs */

typedef struct ditb200_gemm_args {
  const void* a;      /* [M, K] row-major, bf16 (TCGEN05) or f32 (FP32)                  */
  const void* w;      /* [N, K] row-major (nn.Linear layout), same dtype as a            */
  const float* bias;  /* [N] f32 or NULL                                                 */
  void* out;          /* [M, N] row-major; dtype out_dtype (f32 required for GATE_RESID) */
  const float* resid; /* GATE_RESID: [M, N] f32 (may alias out)                          */
  const float* gate;  /* GATE_RESID: gate[(row / rows_per_gate) * gate_stride + col]     */
  int gate_stride;
  int rows_per_gate; /* T: tokens per image */
  int M, N, K;
  int epilogue;  /* DITB200_EPI_*  */
  int out_dtype; /* DITB200_F32 / DITB200_BF16 */
  int engine;    /* DITB200_GEMM_* */
  int tile_n;    /* TCGEN05 only: output-tile width 128/192/256, 0 = choose from the shape */
  int cta_group; /* TCGEN05 only: 1 = one CTA per tile, 2 = CTA pair (256-row tile), 0 = choose */
  /* --- training extras (all optional, zero = off) --- */
  void* aux_out;       /* [M, N], dtype aux_dtype: the pre-epilogue value acc + bias.  With EPI_BIAS_GELU this
                          is fc1's pre-activation, with EPI_BIAS_GATE_RESID the un-gated branch output; both
                          are what the backward pass needs and the fused forward would otherwise discard */
  const void* aux_in;  /* [M, N], dtype aux_dtype: EPI_MUL_DGELU's pre-activation */
  int aux_dtype;
  int accumulate;      /* out += result instead of out = result (f32 out, plain EPI_BIAS; f32 vector atomics) */
  int split_k;         /* TCGEN05: split the K loop over this many work units per tile, combined with f32 atomics
                          (f32 out, plain EPI_BIAS; out is zeroed first unless accumulate != 0); 0/1 = off */
  int trans_a;         /* TCGEN05: a is stored [K, M] row-major and read as an MN-major operand */
  int trans_w;         /* TCGEN05: w is stored [K, N] row-major and read as an MN-major operand.
                          data gradient   dX[M,Kin]    = dY[M,Nout] . W[Nout,Kin]      -> trans_w
                          weight gradient dW[Nout,Kin] = dY[tokens,Nout]^T . X[tokens,Kin] -> trans_a + trans_w */
  int dynamic_sched;   /* TCGEN05 tile scheduling.  0: every persistent CTA pair owns a fixed, longest-first share of
                          the tiles.  1: cluster launch control — the grid has one cluster per tile and running
                          clusters cancel and absorb the ones not yet launched, so SMs held by another kernel (the
                          overlapped NCCL all-reduce of a data-parallel backward, the role of torch DDP in
                          train_options/train_original.py:149) never own tiles.  Same results bit for bit. */
  int reverse_m;       /* TCGEN05: visit the tile rows last-first.  Kernels of a chain alternate their direction so
                          that each starts on the rows its producer wrote last (still in the 126 MB L2).  Same
                          results bit for bit. */
} ditb200_gemm_args;

/* out = epilogue(a · wᵀ).  Replaces, per DiTBlock: timm Attention.qkv (EPI_BIAS),
 * Attention.proj + `x + gate_msa * (.)` (EPI_BIAS_GATE_RESID; models_original.py:120),
 * Mlp.fc1 + GELU(tanh) (EPI_BIAS_GELU; :110-112), Mlp.fc2 + `x + gate_mlp * (.)` (:121).
 * The backward pass runs its data and weight gradients through the same entry point with trans_a /
 * trans_w (autograd's addmm backward in the reference).
 * TCGEN05 engine: N % 8 == 0; K % 8 == 0 unless both operands are transposed; pointers 16-byte aligned. */
int ditb200_gemm(const ditb200_gemm_args* args, void* stream);

/* f32 → bf16 cast of a contiguous array (weight shadows). n elements. */
int ditb200_cast_bf16(const float* in, void* out, size_t n, void* stream);

/* out[i] = silu(in[i]) cast to out_dtype: the nn.SiLU in front of every adaLN_modulation
 * Linear (models_original.py:114,135), producing the A operand of the batched adaLN GEMM. */
int ditb200_silu_cast(const float* in, void* out, int out_dtype, size_t n, void* stream);

/* ------------------------------------------------------------- attention */

/* Fused multi-head attention forward, softmax(q kᵀ / sqrt(hd)) v, non-causal.
 * Replaces timm Attention.forward's reshape/permute + F.scaled_dot_product_attention
 * + transpose/reshape (see performance/A100/train_original.out:36-37 for the layout).
 * qkv[B*T, 3*H*hd] token-major as the qkv Linear writes it (cols: Q heads | K heads |
 * V heads); out[B*T, H*hd] token-major (what proj consumes).  dtype of both =
 * dtype.  lse (optional, f32 [B,H,T]) receives log-sum-exp rows for backward.  reverse != 0: the (image, head)
 * items are visited last-first (tcgen05 kernels; L2 reuse, see ditb200_gemm_args.reverse_m); same results. */
int ditb200_attention_fwd(const void* qkv, void* out, float* lse, int dtype, int B, int T, int H,
                          int hd, int reverse, void* stream);

/* Backward of ditb200_attention_fwd.  qkv/out/dout/dqkv in `dtype`; lse f32 [B,H,T] from the forward;
 * dsum f32 [B,H,T] is caller-owned scratch (row sums of dout*out).  dqkv[B*T, 3*H*hd] is written in
 * full (dQ | dK | dV in the forward's column layout), ready to be the dY of the QKV GEMM's backward. */
int ditb200_attention_bwd(const void* qkv, const void* out, const void* dout, const float* lse,
                          float* dsum, void* dqkv, int dtype, int B, int T, int H, int hd, void* stream);

/* ----------------------------------------------------------- final layer */

/* FinalLayer after its adaLN: LN + modulate + Linear(D → p*p*Cout) + unpatchify
 * into NCHW.  Replaces FinalLayer.forward's norm/modulate/linear
 * (models_original.py:138-142) and DiT.unpatchify (:218-231).
 * x[B*T, D] f32; shift/scale [B,D] slices (row stride mod_stride); w[p*p*Cout, D]
 * f32; bias[p*p*Cout]; out[B, Cout, Hp*p, Wp*p] f32 with T = Hp*Wp, Hp == Wp.
 * round_bf16 != 0 rounds the modulated activations and weights to bf16 first. */
int ditb200_final_layer(const float* x, const float* shift, const float* scale, int mod_stride,
                        const float* w, const float* bias, float* out, int B, int T, int D, int p,
                        int Cout, float eps, int round_bf16, void* stream);

/* ------------------------------------------------------------- optimizer */

/* One fused pass over flat f32 arrays of n elements (n % 4 == 0): torch.optim.AdamW's update
 * (decoupled weight decay, bias correction with `step` >= 1), the EMA update
 * ema = ema_decay*ema + (1-ema_decay)*param (train.py:41-51; ema may be NULL) and the bf16 copy of the
 * updated parameters that the next forward's tensor-core GEMMs read (shadow_bf16 may be NULL).
 * Replaces opt.step() + update_ema() (train.py:161,206-207; train_options/train_original.py:210-211). */
int ditb200_adamw_ema(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, float* ema,
                      void* shadow_bf16, size_t n, float lr, float beta1, float beta2, float eps,
                      float weight_decay, int step, float ema_decay, void* stream);

/* ------------------------------------------------------------- diffusion */

/* Classifier-free-guidance combine on the model output.
 * Replaces DiT.forward_with_cfg's slicing/cat arithmetic (models_original.py:258-266):
 * for the first n_cfg_ch channels, both halves := u + s*(c - u); other channels
 * pass through.  raw[2n, C2, HW] f32 → out[2n, C2, HW] f32 (may alias raw). */
int ditb200_cfg_combine(const float* raw, float* out, int n_half, int C2, int HW, int n_cfg_ch,
                        float cfg_scale, void* stream);

typedef struct ditb200_step_args {
  const float* model_out; /* [B, C2, HW] f32; C2 = 2C if learned variance else C */
  const float* x;         /* [B, C, HW] f32: x_t */
  const float* noise;     /* [B, C, HW] f32 or NULL (then sample = mean) */
  const int64_t* t;       /* [B] int64 indices into the tables */
  /* f32 tables of length num_timesteps (fp64 tables of GaussianDiffusion.__init__,
   * gaussian_diffusion.py:153-201, rounded once to f32 — _extract_into_tensor casts
   * after the gather, :870, which is the same value) */
  const float* sqrt_recip_alphas_cumprod;
  const float* sqrt_recipm1_alphas_cumprod;
  const float* posterior_mean_coef1;
  const float* posterior_mean_coef2;
  const float* min_log;  /* LEARNED_RANGE: posterior_log_variance_clipped; FIXED: the log-variance table */
  const float* max_log;  /* LEARNED_RANGE: log(betas) */
  float* sample;         /* [B, C, HW] out */
  float* pred_xstart;    /* [B, C, HW] out or NULL */
  float* mean;           /* optional out */
  float* log_variance;   /* optional out */
  float* variance;       /* optional out: var_table[t] if var_table != NULL, else exp(log_variance) */
  const float* var_table; /* FIXED types: the variance table that goes with min_log (gaussian_diffusion.py:295-308) */
  /* DDIM only (sampler == DITB200_SAMPLER_DDIM): alphas_cumprod / alphas_cumprod_prev tables */
  const float* alphas_cumprod;
  const float* alphas_cumprod_prev;
  float eta;
  int sampler;           /* DITB200_SAMPLER_* */
  int B, C, HW;
  int num_timesteps;
  int mean_type;     /* DITB200_MEAN_* */
  int var_type;      /* DITB200_VAR_*  */
  int clip_denoised; /* clamp pred_xstart to [-1, 1] */
  int cfg_half;      /* > 0: model_out is the raw two-half CFG batch (n = cfg_half); the
                        combine of ditb200_cfg_combine is applied on the fly */
  int n_cfg_ch;
  float cfg_scale;
  /* DDIM_REVERSE only: alphas_cumprod_next table (gaussian_diffusion.py:171) */
  const float* alphas_cumprod_next;
  /* ANCESTRAL only, optional [B, C, HW]: use this mean instead of the posterior mean in the update — the
   * classifier-guided mean of condition_mean (gaussian_diffusion.py:346-357, 398-401); the `mean` output
   * still reports the unconditioned one */
  const float* mean_override;
} ditb200_step_args;

/* One sampling step x_t → x_{t-1}, fully fused.
 * Replaces GaussianDiffusion.p_mean_variance + _predict_xstart_from_eps +
 * q_posterior_mean_variance + the update in p_sample
 * (diffusion/gaussian_diffusion.py:285-293, 320-323, 334-339, 238-241, 410-416),
 * or the DDIM update of ddim_sample (:541-560) when sampler == DITB200_SAMPLER_DDIM,
 * or the reversed DDIM ODE step of ddim_reverse_sample (:583-598) when sampler == DITB200_SAMPLER_DDIM_REVERSE,
 * and, with cfg_half > 0, forward_with_cfg's combine (models_original.py:258-266). */
int ditb200_p_sample_step(const ditb200_step_args* args, void* stream);

/* q_sample: x_t = sqrt_ac[t]*x0 + sqrt_1mac[t]*noise (gaussian_diffusion.py:215-230). */
int ditb200_q_sample(const float* x0, const float* noise, const int64_t* t,
                     const float* sqrt_alphas_cumprod, const float* sqrt_one_minus_alphas_cumprod,
                     float* x_t, int B, int CHW, int num_timesteps, void* stream);

/* Per-sample affine combination with coefficients gathered from f32 schedule tables:
 *     out[b, i] = ( ta[t_b] * a[b, i]  (+|-)  tb[t_b] * b[b, i] * b2[b, i] ) / td[t_b]
 * Any factor may be absent: ta == NULL -> coefficient 1, a == NULL -> the first term is ta[t_b] itself (a
 * broadcast of the table: _extract_into_tensor, gaussian_diffusion.py:861-873); b == NULL -> no second term;
 * tb == NULL -> coefficient 1; b2 == NULL -> factor 1; td == NULL -> no division; subtract != 0 -> minus.
 * Every product / sum is rounded separately in the order written (no FMA contraction), so the results are the
 * reference's eager f32 values bit for bit.  Serves the standalone helpers of GaussianDiffusion:
 * q_mean_variance (:203-213), q_posterior_mean_variance (:232-252), _predict_xstart_from_eps (:334-339),
 * _predict_eps_from_xstart (:341-344), condition_mean (:346-357) and the eps update of condition_score (:367-368).
 * t int64 [B]; tables f32 [num_timesteps]; tensors f32 [B, n]. */
int ditb200_diffusion_affine(const float* a, const float* b, const float* b2, const int64_t* t, const float* ta,
                             const float* tb, const float* td, int subtract, float* out, int B, int n,
                             int num_timesteps, void* stream);

/* _prior_bpd (gaussian_diffusion.py:789-803): per sample, mean over elements of
 * KL( N(sqrt_ac[T-1] * x0, 1 - ac[T-1]) || N(0, 1) ) in bits.  coef_mean = sqrt_alphas_cumprod[T-1],
 * log_var = log_one_minus_alphas_cumprod[T-1] (f32 values of the tables).  x0 [B, n] -> out [B]. */
int ditb200_prior_bpd(const float* x0, float coef_mean, float log_var, float* out, int B, int n, void* stream);

typedef struct ditb200_loss_args {
  const float* model_out; /* [B, 2C, HW] f32 ([B, C, HW] with DITB200_VAR_FIXED) */
  const float* x0;        /* [B, C, HW] */
  const float* x_t;       /* [B, C, HW] */
  const float* noise;     /* [B, C, HW] */
  const int64_t* t;
  const float* sqrt_recip_alphas_cumprod;
  const float* sqrt_recipm1_alphas_cumprod;
  const float* posterior_mean_coef1;
  const float* posterior_mean_coef2;
  const float* posterior_log_variance_clipped;
  const float* log_betas;
  float* mse;             /* [B] out */
  float* vb;              /* [B] out */
  float* loss;            /* [B] out: mse + vb */
  float* grad_model_out;  /* [B, 2C, HW] out or NULL: d(sum_b w_mse[b]*mse[b] + w_vb[b]*vb[b]) / d model_out */
  const float* w_mse;     /* [B] upstream gradients of mse (+loss); required with grad_model_out */
  const float* w_vb;      /* [B] upstream gradients of vb (+loss) */
  float vb_scale;         /* 1, or num_timesteps/1000 for RESCALED_MSE (gaussian_diffusion.py:766-769);
                             num_timesteps for RESCALED_KL (:745-746) */
  int B, C, HW, num_timesteps;
  /* --- everything below is optional; all-zero reproduces the MSE + LEARNED_RANGE + EPSILON loss --- */
  int mean_type;          /* DITB200_MEAN_*: what model_out's first C channels predict */
  int var_type;           /* DITB200_VAR_*; FIXED reads fixed_log_var */
  const float* fixed_log_var; /* VAR_FIXED: the model log-variance table (gaussian_diffusion.py:295-308) */
  int clip_denoised;      /* clamp the x0 prediction to [-1, 1] before the posterior mean (_vb_terms_bpd's
                             clip_denoised; training passes False, calc_bpd_loop True) */
  int vb_through_mean;    /* 0: the VLB term sees a detached mean (MSE family, :755-758).  1: its gradient also
                             flows into the mean channels — the KL / RESCALED_KL loss (:735-746) */
  float* pred_xstart;     /* [B, C, HW] out or NULL: the (clipped) x0 prediction (_vb_terms_bpd's 'pred_xstart') */
  float* xstart_mse;      /* [B] out or NULL: mean((pred_xstart - x0)^2)                     (calc_bpd_loop :842) */
  float* eps_mse;         /* [B] out or NULL: mean((eps(pred_xstart) - noise)^2)             (calc_bpd_loop :843-844) */
} ditb200_loss_args;

/* training_losses / _vb_terms_bpd: mse, vb (KL or decoder NLL at t == 0, in bits), loss = mse + vb, and the
 * gradient wrt the model output, one kernel with warp-shuffle row reductions.  The default (zeroed optional
 * fields) is MSE + LEARNED_RANGE + EPSILON, what create_diffusion("") builds.
 * Replaces gaussian_diffusion.py:735-781, 682-713, 836-844 and diffusion_utils.py:10-36,62-88. */
int ditb200_training_losses(const ditb200_loss_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DITB200_H_ */
