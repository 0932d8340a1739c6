"""Thin torch-tensor wrappers over the libditb200 C-ABI.

PyTorch is plumbing here: it owns device memory and the stream.  Every function
takes CUDA tensors, passes raw pointers + the current stream to the library and
returns the output tensor.  Nothing in this file computes on the host, and
nothing falls back to torch ops.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib as L

_DT = {torch.float32: L.F32, torch.bfloat16: L.BF16}


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _lib_for(t: torch.Tensor):
    if not t.is_cuda:
        raise L.Ditb200Error("libditb200 ops need CUDA tensors (there is no CPU path)")
    idx = t.device.index if t.device.index is not None else torch.cuda.current_device()
    if idx != torch.cuda.current_device():
        # kernels are launched on torch's CURRENT stream, which belongs to the current device
        raise L.Ditb200Error(f"tensor lives on cuda:{idx} but the current device is cuda:{torch.cuda.current_device()}; "
                             "call torch.cuda.set_device() first (one process per GPU)")
    return L.ensure_init(idx)


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


LAUNCHES = 0          # kernels launched through this module since import (bench.py reads deltas)
_PROFILE = None       # list of (name, start_event, end_event, meta) while profile() is active


class profile:
    """Context manager: brackets every kernel launch with CUDA events on the launching stream.

        with ops.profile() as p: model(x, t, y)
        p.summary() -> {kernel name: (launches, total ms)}
    """

    def __enter__(self):
        global _PROFILE
        self.records = []
        _PROFILE = self.records
        return self

    def __exit__(self, *exc):
        global _PROFILE
        _PROFILE = None
        torch.cuda.synchronize()
        return False

    def summary(self):
        out = {}
        for name, e0, e1, meta, _tag in self.records:
            n, ms, flops = out.get(name, (0, 0.0, 0.0))
            out[name] = (n + 1, ms + e0.elapsed_time(e1), flops + (meta or 0.0))
        return out

    def summary_by_tag(self):
        """{(kernel name, tag): (launches, total ms, flops)}; the GEMM tags itself with its shape and flags."""
        out = {}
        for name, e0, e1, meta, tag in self.records:
            n, ms, flops = out.get((name, tag), (0, 0.0, 0.0))
            out[(name, tag)] = (n + 1, ms + e0.elapsed_time(e1), flops + (meta or 0.0))
        return out


def _call(name, fn, *args, meta=None, tag=None):
    """Invoke one C-ABI entry point (= one kernel launch) and check its return code."""
    global LAUNCHES
    LAUNCHES += 1
    if _PROFILE is None:
        L.check(fn(*args), name)
        return
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    L.check(fn(*args), name)
    e1.record()
    _PROFILE.append((name, e0, e1, meta, tag))


def _chk_contig(*ts):
    for t in ts:
        if t is not None and not t.is_contiguous():
            raise L.Ditb200Error("non-contiguous tensor passed to libditb200")


# ------------------------------------------------------------------ embedders
def patch_embed(x, w, bias, pos, p: int, round_bf16: bool = False):
    """x[B,C,H,W] f32, w[D,C,p,p], bias[D], pos[T,D] (or [1,T,D]) -> [B*T, D] f32."""
    lib = _lib_for(x)
    _chk_contig(x, w, bias, pos)
    B, Cc, H, W = x.shape
    D = w.shape[0]
    T = (H // p) * (W // p)
    out = torch.empty((B * T, D), device=x.device, dtype=torch.float32)
    _call("patch_embed", lib.ditb200_patch_embed, _p(x), _p(w), _p(bias), _p(pos), _p(out), B, Cc, H, W, p, D,
                                    int(round_bf16), _stream())
    return out


def timestep_embedding(t, dim: int, max_period: float = 10000.0):
    """t[B] -> [B, dim] sinusoidal features.  Integer timesteps travel as int64; floating ones (the reference's
    embedder accepts fractional t: models_original.py:40-59 does t[:, None].float()) as f32."""
    lib = _lib_for(t)
    if t.is_floating_point():
        t, is_float = t.to(torch.float32).contiguous(), 1
    else:
        t, is_float = t.to(torch.int64).contiguous(), 0
    out = torch.empty((t.shape[0], dim), device=t.device, dtype=torch.float32)
    _call("timestep_embedding", lib.ditb200_timestep_embedding, _p(t), is_float, _p(out), t.shape[0], dim,
          float(max_period), _stream())
    return out


def small_linear(a, w, bias=None, add=None, silu_in=False, silu_out=False, out=None):
    """out[M,N] = act_out(act_in(a) @ w.T + bias) (+ add); a f32 [M,K]; w f32/bf16 [N,K]."""
    lib = _lib_for(a)
    M, K = a.shape
    N = w.shape[0]
    assert w.shape[1] == K and a.stride(1) == 1 and w.is_contiguous()
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=torch.float32)
    assert out.stride(1) == 1
    _call("small_linear", lib.ditb200_small_linear, _p(a), a.stride(0), _p(w), _DT[w.dtype], _p(bias), _p(add),
                                     add.stride(0) if add is not None else 0, _p(out), out.stride(0),
                                     M, N, K, int(silu_in), int(silu_out), _stream())
    return out


def label_embed(y, table, add=None):
    lib = _lib_for(table)
    y = y.to(torch.int64).contiguous()
    _chk_contig(table, add)
    B, D = y.shape[0], table.shape[1]
    out = torch.empty((B, D), device=table.device, dtype=torch.float32)
    _call("label_embed", lib.ditb200_label_embed, _p(y), _p(table), _p(add), _p(out), B, D, table.shape[0], _stream())
    return out


# ------------------------------------------------------- LayerNorm + modulate
def ln_modulate(x, shift, scale, T: int, out_dtype=torch.bfloat16, eps: float = 1e-6, stats=None, out=None,
                reverse: bool = False):
    """x[B*T, D] f32; shift/scale: [B, D] views (unit inner stride, common row stride)."""
    lib = _lib_for(x)
    M, D = x.shape
    B = M // T
    assert x.is_contiguous() and shift.stride(1) == 1 and scale.stride(1) == 1
    assert shift.stride(0) == scale.stride(0)
    if out is None:
        out = torch.empty((M, D), device=x.device, dtype=out_dtype)
    _call("ln_modulate", lib.ditb200_ln_modulate, _p(x), _p(shift), _p(scale), shift.stride(0), _p(out), _DT[out.dtype],
                                    _p(stats), B, T, D, float(eps), int(reverse), _stream())
    return out


def ln_modulate_resid(x, y, gate, shift, scale, T: int, out_dtype=torch.bfloat16, eps: float = 1e-6, stats=None,
                      x_out=None, want_out: bool = True, reverse: bool = False):
    """x_out = x + gate[b] * y (y bf16); out = LN(x_out) * (1 + scale[b]) + shift[b].  Returns (x_out, out)."""
    lib = _lib_for(x)
    M, D = x.shape
    B = M // T
    _chk_contig(x, y)
    assert y.dtype == torch.bfloat16 and y.shape == x.shape
    assert gate.stride(1) == 1
    if want_out:
        assert shift.stride(1) == 1 and scale.stride(1) == 1 and gate.stride(0) == shift.stride(0) == scale.stride(0)
    if x_out is None:
        x_out = torch.empty_like(x)
    out = torch.empty((M, D), device=x.device, dtype=out_dtype) if want_out else None
    _call("ln_modulate_resid", lib.ditb200_ln_modulate_resid, _p(x), _p(y), _p(gate), _p(shift if want_out else None),
          _p(scale if want_out else None), gate.stride(0), _p(x_out), _p(out), _DT[out_dtype], _p(stats), B, T, D,
          float(eps), int(reverse), _stream())
    return x_out, out


# ----------------------------------------------------------------------- GEMM
def gemm(a, w, bias=None, *, epilogue=L.EPI_BIAS, out_dtype=None, out=None, resid=None, gate=None,
         rows_per_gate: int = 0, engine: Optional[int] = None, tile_n: int = 0, cta_group: int = 0,
         aux_out=None, aux_in=None, accumulate: bool = False, split_k: int = 0, trans_a: bool = False,
         trans_w: bool = False, reverse_m: bool = False):
    """out = epilogue(op(a) @ op(w).T).  a[M,K] (or [K,M] with trans_a), w[N,K] (or [K,N] with trans_w), both
    bf16 (tcgen05) or both f32 (check mode, forward only)."""
    lib = _lib_for(a)
    _chk_contig(a, w, bias, resid, aux_out, aux_in)
    K, M = a.shape if trans_a else a.shape[::-1]
    Kw, N = w.shape if trans_w else w.shape[::-1]
    assert Kw == K and a.dtype == w.dtype, (a.shape, w.shape, trans_a, trans_w)
    if engine is None:
        engine = L.GEMM_TCGEN05 if a.dtype == torch.bfloat16 else L.GEMM_FP32
    if epilogue == L.EPI_BIAS_GATE_RESID:
        out_dtype = torch.float32
        if out is None:
            out = resid  # in place on the residual stream
        assert gate is not None and gate.stride(1) == 1
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype or a.dtype)
    assert out.shape == (M, N) and out.is_contiguous()
    aux = aux_out if aux_out is not None else aux_in
    args = L.GemmArgs(_p(a), _p(w), _p(bias), _p(out), _p(resid), _p(gate),
                      gate.stride(0) if gate is not None else 0, rows_per_gate, M, N, K, epilogue,
                      _DT[out.dtype], engine, tile_n, cta_group,
                      _p(aux_out), _p(aux_in), _DT[aux.dtype] if aux is not None else 0, int(accumulate), int(split_k),
                      int(trans_a), int(trans_w), int(_GEMM_DYNAMIC), int(reverse_m))
    _call("gemm_tc" if engine == L.GEMM_TCGEN05 else "gemm_fp32", lib.ditb200_gemm, C.byref(args), _stream(),
          meta=2.0 * M * N * K,
          tag=None if _PROFILE is None else
          f"M{M} N{N} K{K} epi{epilogue}{' tA' if trans_a else ''}{' tW' if trans_w else ''}"
          f"{' sk' + str(split_k) if split_k > 1 else ''}{' acc' if accumulate else ''}{' aux' if aux_out is not None else ''}"
          f" {'f32' if out.dtype == torch.float32 else 'bf16'}")
    return out


_GEMM_DYNAMIC = os.environ.get("DITB200_GEMM_DYNAMIC") is not None


def set_gemm_dynamic(on: bool) -> bool:
    """Tile scheduler requested by the following gemm() calls (ditb200_gemm_args.dynamic_sched: cluster launch
    control instead of the static longest-first schedule); returns the previous setting.  Host-side default of this
    module only: the library itself keeps no such state."""
    global _GEMM_DYNAMIC
    prev, _GEMM_DYNAMIC = _GEMM_DYNAMIC, bool(on)
    return prev


def cast_bf16(x, out=None):
    lib = _lib_for(x)
    _chk_contig(x)
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    _call("cast_bf16", lib.ditb200_cast_bf16, _p(x), _p(out), x.numel(), _stream())
    return out


def silu_cast(x, out_dtype=torch.bfloat16):
    lib = _lib_for(x)
    _chk_contig(x)
    out = torch.empty(x.shape, device=x.device, dtype=out_dtype)
    _call("silu_cast", lib.ditb200_silu_cast, _p(x), _p(out), _DT[out_dtype], x.numel(), _stream())
    return out


# ------------------------------------------------------------------ attention
def attention(qkv, B: int, T: int, H: int, hd: int, lse=None, out=None, reverse: bool = False):
    """qkv[B*T, 3*H*hd] -> out[B*T, H*hd], same dtype (bf16 or f32)."""
    lib = _lib_for(qkv)
    _chk_contig(qkv)
    if out is None:
        out = torch.empty((B * T, H * hd), device=qkv.device, dtype=qkv.dtype)
    _call("attention_fwd", lib.ditb200_attention_fwd, _p(qkv), _p(out), _p(lse), _DT[qkv.dtype], B, T, H, hd,
          int(reverse), _stream(),
          meta=4.0 * B * H * T * T * hd)
    return out


# ---------------------------------------------------------------- final layer
def final_layer(x, shift, scale, w, bias, T: int, p: int, c_out: int, eps: float = 1e-6,
                round_bf16: bool = False):
    """x[B*T, D] f32 -> [B, c_out, H, W] f32 (LN + modulate + linear + unpatchify)."""
    lib = _lib_for(x)
    _chk_contig(x, w, bias)
    M, D = x.shape
    B = M // T
    hp = int(round(T ** 0.5))
    out = torch.empty((B, c_out, hp * p, hp * p), device=x.device, dtype=torch.float32)
    assert shift.stride(1) == 1 and shift.stride(0) == scale.stride(0)
    _call("final_layer", lib.ditb200_final_layer, _p(x), _p(shift), _p(scale), shift.stride(0), _p(w), _p(bias), _p(out),
                                    B, T, D, p, c_out, float(eps), int(round_bf16), _stream())
    return out


# ------------------------------------------------------------------ diffusion
def cfg_combine(raw, n_cfg_ch: int, cfg_scale: float, out=None):
    lib = _lib_for(raw)
    _chk_contig(raw)
    B, C2 = raw.shape[0], raw.shape[1]
    HW = raw[0, 0].numel()
    if out is None:
        out = torch.empty_like(raw)
    _call("cfg_combine", lib.ditb200_cfg_combine, _p(raw), _p(out), B // 2, C2, HW, n_cfg_ch, float(cfg_scale), _stream())
    return out


def p_sample_step(model_out, x, noise, t, tables, *, mean_type, var_type, clip_denoised,
                  cfg_half=0, n_cfg_ch=0, cfg_scale=1.0, want=("sample", "pred_xstart"),
                  sampler=L.SAMPLER_ANCESTRAL, eta=0.0, mean_override=None):
    """One fused sampling step (ancestral, DDIM or reversed DDIM).  tables: dict of f32 device tensors (see
    diffusion/); mean_override: the classifier-guided mean of condition_mean (ancestral only)."""
    lib = _lib_for(x)
    _chk_contig(model_out, x, noise, t, mean_override)
    B, Cc = x.shape[0], x.shape[1]
    HW = x[0, 0].numel()
    outs = {k: torch.empty_like(x) for k in want}
    if "sample" not in outs:
        outs["sample"] = torch.empty_like(x)
    args = L.StepArgs(
        _p(model_out), _p(x), _p(noise), _p(t),
        _p(tables.get("sqrt_recip_alphas_cumprod")), _p(tables.get("sqrt_recipm1_alphas_cumprod")),
        _p(tables["posterior_mean_coef1"]), _p(tables["posterior_mean_coef2"]),
        _p(tables["min_log"]), _p(tables.get("max_log")),
        _p(outs["sample"]), _p(outs.get("pred_xstart")), _p(outs.get("mean")), _p(outs.get("log_variance")),
        _p(outs.get("variance")), _p(tables.get("var_table")), _p(tables.get("alphas_cumprod")), _p(tables.get("alphas_cumprod_prev")),
        float(eta), int(sampler),
        B, Cc, HW, int(tables["posterior_mean_coef1"].numel()),
        mean_type, var_type, int(bool(clip_denoised)), int(cfg_half), int(n_cfg_ch), float(cfg_scale),
        _p(tables.get("alphas_cumprod_next")), _p(mean_override))
    _call("p_sample_step", lib.ditb200_p_sample_step, C.byref(args), _stream())
    return outs


def q_sample(x0, noise, t, sqrt_ac, sqrt_1mac):
    lib = _lib_for(x0)
    _chk_contig(x0, noise, t)
    out = torch.empty_like(x0)
    B = x0.shape[0]
    _call("q_sample", lib.ditb200_q_sample, _p(x0), _p(noise), _p(t), _p(sqrt_ac), _p(sqrt_1mac), _p(out), B,
                                 x0[0].numel(), int(sqrt_ac.numel()), _stream())
    return out


def training_losses(model_out, x0, x_t, noise, t, tables, vb_scale: float = 1.0, w_mse=None, w_vb=None, *,
                    mean_type=L.MEAN_EPSILON, var_type=L.VAR_LEARNED_RANGE, clip_denoised=False, vb_through_mean=False,
                    want_pred=False, want_bpd_terms=False):
    """mse / vb / loss per sample; with w_mse and w_vb ([B] upstream gradients) also the gradient
    with respect to model_out.  vb_through_mean: the VLB term's gradient reaches the mean channels (KL losses).
    want_pred / want_bpd_terms: also return the x0 prediction / calc_bpd_loop's xstart_mse and eps mse."""
    lib = _lib_for(x0)
    _chk_contig(model_out, x0, x_t, noise, t, w_mse, w_vb)
    B, Cc = x0.shape[0], x0.shape[1]
    HW = x0[0, 0].numel()
    dev = x0.device
    mse = torch.empty(B, device=dev, dtype=torch.float32)
    vb = torch.empty_like(mse)
    loss = torch.empty_like(mse)
    grad = torch.empty_like(model_out) if w_mse is not None else None
    pred = torch.empty_like(x0) if want_pred else None
    xs_mse = torch.empty_like(mse) if want_bpd_terms else None
    eps_mse = torch.empty_like(mse) if want_bpd_terms else None
    args = L.LossArgs(
        _p(model_out), _p(x0), _p(x_t), _p(noise), _p(t),
        _p(tables["sqrt_recip_alphas_cumprod"]), _p(tables["sqrt_recipm1_alphas_cumprod"]),
        _p(tables["posterior_mean_coef1"]), _p(tables["posterior_mean_coef2"]),
        _p(tables["posterior_log_variance_clipped"]), _p(tables["log_betas"]),
        _p(mse), _p(vb), _p(loss), _p(grad), _p(w_mse), _p(w_vb), float(vb_scale), B, Cc, HW,
        int(tables["posterior_mean_coef1"].numel()),
        int(mean_type), int(var_type), _p(tables.get("min_log") if var_type == L.VAR_FIXED else None),
        int(bool(clip_denoised)), int(bool(vb_through_mean)), _p(pred), _p(xs_mse), _p(eps_mse))
    _call("training_losses", lib.ditb200_training_losses, C.byref(args), _stream())
    return {"mse": mse, "vb": vb, "loss": loss, "grad_model_out": grad, "pred_xstart": pred, "xstart_mse": xs_mse,
            "eps_mse": eps_mse}


def diffusion_affine(t, *, a=None, ta=None, b=None, tb=None, b2=None, td=None, subtract=False, like=None):
    """out = (ta[t] * a (+|-) tb[t] * b * b2) / td[t] per sample (include/ditb200.h: ditb200_diffusion_affine)."""
    ref = a if a is not None else (b if b is not None else like)
    lib = _lib_for(ref)
    _chk_contig(a, b, b2, t)
    B = ref.shape[0]
    n = ref[0].numel()
    out = torch.empty(ref.shape, device=ref.device, dtype=torch.float32)
    table = ta if ta is not None else (tb if tb is not None else td)
    _call("diffusion_affine", lib.ditb200_diffusion_affine, _p(a), _p(b), _p(b2), _p(t), _p(ta), _p(tb), _p(td),
          int(bool(subtract)), _p(out), B, n, int(table.numel()) if table is not None else 1, _stream())
    return out


def prior_bpd(x0, coef_mean: float, log_var: float):
    lib = _lib_for(x0)
    _chk_contig(x0)
    out = torch.empty(x0.shape[0], device=x0.device, dtype=torch.float32)
    _call("prior_bpd", lib.ditb200_prior_bpd, _p(x0), float(coef_mean), float(log_var), _p(out), x0.shape[0],
          x0[0].numel(), _stream())
    return out


# ------------------------------------------------------------------- backward pass
def ln_modulate_bwd(dh, x, scale, stats, T: int, dx, accumulate: bool, dshift, dscale):
    """Backward of ln_modulate: dx (+)= d/dx, dshift/dscale ([B, D] views, common row stride) += sums over tokens."""
    lib = _lib_for(x)
    _chk_contig(dh, x, stats, dx)
    M, D = x.shape
    assert scale.stride(1) == 1 and dshift.stride(1) == 1 and dscale.stride(1) == 1 and dshift.stride(0) == dscale.stride(0)
    _call("ln_modulate_bwd", lib.ditb200_ln_modulate_bwd, _p(dh), _DT[dh.dtype], _p(x), _p(scale), scale.stride(0),
          _p(stats), _p(dx), int(accumulate), _p(dshift), _p(dscale), dshift.stride(0), M // T, T, D, _stream())
    return dx


def ln_modulate_bwd_gate_ok(T: int, D: int) -> bool:
    """Shapes the fused LayerNorm-backward + gated-residual-backward kernel serves."""
    return T % 4 == 0 and D in (384, 768, 1024, 1152)


def ln_modulate_bwd_gate(dh, x, scale, stats, T: int, dx, accumulate: bool, dshift, dscale, y=None, gate=None,
                         dgate=None, dbias=None, dy=None):
    """ln_modulate_bwd, then — on the dx just formed, in the same pass — gate_resid_bwd of the branch y (bf16) that
    joined the stream in front of this LayerNorm: dy = dx * gate[b]; dgate += sum_t dx * y; dbias += colsum(dy).
    Returns (dx, dy); y=None: only the LayerNorm part."""
    lib = _lib_for(x)
    _chk_contig(dh, x, stats, dx, y, dy, dbias)
    M, D = x.shape
    assert scale.stride(1) == 1 and dshift.stride(1) == 1 and dscale.stride(1) == 1 and dshift.stride(0) == dscale.stride(0)
    if y is not None:
        assert y.dtype == torch.bfloat16 and gate.stride(1) == 1 and dgate.stride(1) == 1
        if dy is None:
            dy = torch.empty_like(y)
    _call("ln_modulate_bwd_gate", lib.ditb200_ln_modulate_bwd_gate, _p(dh), _DT[dh.dtype], _p(x), _p(scale),
          scale.stride(0), _p(stats), _p(dx), int(accumulate), _p(dshift), _p(dscale), dshift.stride(0), _p(y),
          _p(gate), gate.stride(0) if gate is not None else 0, _p(dy), _p(dgate),
          dgate.stride(0) if dgate is not None else 0, _p(dbias), M // T, T, D, _stream())
    return dx, dy


def gate_resid_bwd(dx_out, y, gate, T: int, dgate, dbias=None, dy=None):
    """dy = dx_out * gate[b]; dgate += sum_t dx_out * y; dbias += colsum(dy)."""
    lib = _lib_for(dx_out)
    _chk_contig(dx_out, y, dy, dbias)
    M, D = dx_out.shape
    if dy is None:
        dy = torch.empty_like(y)
    assert gate.stride(1) == 1 and dgate.stride(1) == 1
    _call("gate_resid_bwd", lib.ditb200_gate_resid_bwd, _p(dx_out), _p(y), _DT[y.dtype], _p(gate), gate.stride(0),
          _p(dy), _DT[dy.dtype], _p(dgate), dgate.stride(0), _p(dbias), M // T, T, D, _stream())
    return dy


def colsum(x, out=None, accumulate: bool = False):
    lib = _lib_for(x)
    _chk_contig(x, out)
    R, Cc = x.shape
    if out is None:
        out = torch.empty(Cc, device=x.device, dtype=torch.float32)
    _call("colsum", lib.ditb200_colsum, _p(x), _DT[x.dtype], _p(out), int(accumulate), R, Cc, _stream())
    return out


def adaln_wgrad_ok(N: int, R: int, D: int) -> bool:
    """Shapes (and batch sizes: the kernel's cost grows with N, the tensor-core route's does not) the outer-product
    weight-gradient kernel of the adaLN Linears serves."""
    return N <= 256 and R % 64 == 0 and D % 128 == 0


def adaln_wgrad_preferred(N: int, R: int, D: int) -> bool:
    """Whether training routes an adaLN Linear's weight gradient through the outer-product kernel rather than cast +
    tcgen05 GEMM (K = N images) + column sum.  The kernel's time is about (0.84 + 0.064 N) ps per output element
    (23 us at N = 32 for XL/2's 6912 x 1152, 61 us at N = 256 for B/4's 4608 x 768), the GEMM route's about 4.5 ps
    whatever N: the crossover is near 57 images.  DITB200_ADALN_SIMT_MAX_N overrides the limit (measurement)."""
    limit = int(os.environ.get("DITB200_ADALN_SIMT_MAX_N", "64"))
    return adaln_wgrad_ok(N, R, D) and N <= limit


def adaln_wgrad(dmod, sc, dw, dbias=None):
    """dw[r, c] = sum_b dmod[b, r] * sc[b, c]; dbias[r] = sum_b dmod[b, r].  dmod f32 [N, R] (a column slice of a wider
    buffer is fine), sc bf16 [N, D]; dw [R, D] and dbias [R] are overwritten."""
    lib = _lib_for(dmod)
    _chk_contig(sc, dw, dbias)
    N, R = dmod.shape
    D = sc.shape[1]
    assert dmod.stride(1) == 1 and dmod.dtype == torch.float32 and sc.dtype == torch.bfloat16 and dw.shape == (R, D)
    _call("adaln_wgrad", lib.ditb200_adaln_wgrad, _p(dmod), dmod.stride(0), _p(sc), _p(dw), _p(dbias), N, R, D,
          _stream())
    return dw


def label_embed_bwd(dc, y, dtable):
    lib = _lib_for(dc)
    y = y.to(torch.int64).contiguous()
    _chk_contig(dc, dtable)
    _call("label_embed_bwd", lib.ditb200_label_embed_bwd, _p(dc), _p(y), _p(dtable), dc.shape[0], dc.shape[1],
          dtable.shape[0], _stream())
    return dtable


def patchify(x, p: int, out_dtype=torch.bfloat16):
    lib = _lib_for(x)
    _chk_contig(x)
    B, Cc, H, W = x.shape
    out = torch.empty((B * (H // p) * (W // p), Cc * p * p), device=x.device, dtype=out_dtype)
    _call("patchify", lib.ditb200_patchify, _p(x), _p(out), _DT[out_dtype], B, Cc, H, W, p, _stream())
    return out


def unpatchify_bwd(dout, p: int):
    lib = _lib_for(dout)
    _chk_contig(dout)
    B, Cout, Himg, _ = dout.shape
    hp = Himg // p
    dz = torch.empty((B * hp * hp, p * p * Cout), device=dout.device, dtype=torch.bfloat16)
    _call("unpatchify_bwd", lib.ditb200_unpatchify_bwd, _p(dout), _p(dz), B, Cout, hp, p, _stream())
    return dz


def silu_bwd(dact, pre, out=None, accumulate: bool = False):
    lib = _lib_for(dact)
    _chk_contig(dact, pre, out)
    if out is None:
        out = torch.empty_like(pre)
    _call("silu_bwd", lib.ditb200_silu_bwd, _p(dact), _p(pre), _p(out), int(accumulate), pre.numel(), _stream())
    return out


def attention_bwd(qkv, out, dout, lse, B: int, T: int, H: int, hd: int):
    """dqkv[B*T, 3*H*hd] from the forward's qkv, out, lse and the gradient of out."""
    lib = _lib_for(qkv)
    _chk_contig(qkv, out, dout, lse)
    dsum = torch.empty((B, H, T), device=qkv.device, dtype=torch.float32)
    dqkv = torch.empty_like(qkv)
    _call("attention_bwd", lib.ditb200_attention_bwd, _p(qkv), _p(out), _p(dout), _p(lse), _p(dsum), _p(dqkv),
          _DT[qkv.dtype], B, T, H, hd, _stream(), meta=10.0 * B * H * T * T * hd)
    return dqkv


def adamw_ema(param, grad, exp_avg, exp_avg_sq, ema, shadow, *, lr, beta1, beta2, eps, weight_decay, step, ema_decay):
    """One fused optimizer + EMA + bf16-shadow pass over flat arrays (csrc/optim.cu)."""
    lib = _lib_for(param)
    _chk_contig(param, grad, exp_avg, exp_avg_sq, ema, shadow)
    _call("adamw_ema", lib.ditb200_adamw_ema, _p(param), _p(grad), _p(exp_avg), _p(exp_avg_sq), _p(ema), _p(shadow),
          param.numel(), float(lr), float(beta1), float(beta2), float(eps), float(weight_decay), int(step),
          float(ema_decay), _stream())
