"""ctypes binding of libditb200.so (include/ditb200.h).

There is no fallback: if the shared library is missing or an entry point fails,
the caller gets an exception.  The CUDA path is the only path.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "lib" / "libditb200.so"

F32, BF16 = 0, 1
ABI_VERSION = 3
EPI_BIAS, EPI_BIAS_GELU, EPI_BIAS_GATE_RESID, EPI_BIAS_SILU, EPI_MUL_DGELU = 0, 1, 2, 3, 4
EPI_BIAS_GELU_DAUX, EPI_MUL_AUX = 5, 6
GEMM_TCGEN05, GEMM_FP32 = 0, 1
MEAN_EPSILON, MEAN_START_X = 0, 1
VAR_LEARNED_RANGE, VAR_LEARNED, VAR_FIXED = 0, 1, 2
SAMPLER_ANCESTRAL, SAMPLER_DDIM, SAMPLER_DDIM_REVERSE = 0, 1, 2

_vp, _i, _f, _sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t


class GemmArgs(C.Structure):
    _fields_ = [
        ("a", _vp), ("w", _vp), ("bias", _vp), ("out", _vp), ("resid", _vp), ("gate", _vp),
        ("gate_stride", _i), ("rows_per_gate", _i),
        ("M", _i), ("N", _i), ("K", _i),
        ("epilogue", _i), ("out_dtype", _i), ("engine", _i), ("tile_n", _i), ("cta_group", _i),
        ("aux_out", _vp), ("aux_in", _vp), ("aux_dtype", _i), ("accumulate", _i), ("split_k", _i),
        ("trans_a", _i), ("trans_w", _i), ("dynamic_sched", _i), ("reverse_m", _i),
    ]


class StepArgs(C.Structure):
    _fields_ = [
        ("model_out", _vp), ("x", _vp), ("noise", _vp), ("t", _vp),
        ("sqrt_recip_alphas_cumprod", _vp), ("sqrt_recipm1_alphas_cumprod", _vp),
        ("posterior_mean_coef1", _vp), ("posterior_mean_coef2", _vp),
        ("min_log", _vp), ("max_log", _vp),
        ("sample", _vp), ("pred_xstart", _vp), ("mean", _vp), ("log_variance", _vp), ("variance", _vp),
        ("var_table", _vp), ("alphas_cumprod", _vp), ("alphas_cumprod_prev", _vp), ("eta", _f), ("sampler", _i),
        ("B", _i), ("C", _i), ("HW", _i), ("num_timesteps", _i),
        ("mean_type", _i), ("var_type", _i), ("clip_denoised", _i),
        ("cfg_half", _i), ("n_cfg_ch", _i), ("cfg_scale", _f),
        ("alphas_cumprod_next", _vp), ("mean_override", _vp),
    ]


class LossArgs(C.Structure):
    _fields_ = [
        ("model_out", _vp), ("x0", _vp), ("x_t", _vp), ("noise", _vp), ("t", _vp),
        ("sqrt_recip_alphas_cumprod", _vp), ("sqrt_recipm1_alphas_cumprod", _vp),
        ("posterior_mean_coef1", _vp), ("posterior_mean_coef2", _vp),
        ("posterior_log_variance_clipped", _vp), ("log_betas", _vp),
        ("mse", _vp), ("vb", _vp), ("loss", _vp), ("grad_model_out", _vp),
        ("w_mse", _vp), ("w_vb", _vp), ("vb_scale", _f),
        ("B", _i), ("C", _i), ("HW", _i), ("num_timesteps", _i),
        ("mean_type", _i), ("var_type", _i), ("fixed_log_var", _vp), ("clip_denoised", _i), ("vb_through_mean", _i),
        ("pred_xstart", _vp), ("xstart_mse", _vp), ("eps_mse", _vp),
    ]


# name -> (restype, argtypes); must list every symbol include/ditb200.h declares
SIGNATURES = {
    "ditb200_abi_version": (_i, []),
    "ditb200_init": (_i, [_i]),
    "ditb200_last_error": (C.c_char_p, []),
    "ditb200_debug_tile_schedule": (_i, [_i] * 7 + [C.c_void_p, _i]),
    "ditb200_debug_gemm_plan": (_i, [_i] * 6 + [C.c_void_p]),
    "ditb200_debug_gemm_table": (_i, [_i] * 4 + [C.c_void_p, _i, C.c_void_p]),
    "ditb200_debug_attention_path": (_i, [_i] * 3),
    "ditb200_sm_count": (_i, []),
    "ditb200_patch_embed": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _i, _vp]),
    "ditb200_timestep_embedding": (_i, [_vp, _i, _vp, _i, _i, _f, _vp]),
    "ditb200_small_linear": (_i, [_vp, _i, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ditb200_label_embed": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ditb200_ln_modulate": (_i, [_vp, _vp, _vp, _i, _vp, _i, _vp, _i, _i, _i, _f, _i, _vp]),
    "ditb200_ln_modulate_resid": (_i, [_vp, _vp, _vp, _vp, _vp, _i, _vp, _vp, _i, _vp, _i, _i, _i, _f, _i, _vp]),
    "ditb200_ln_modulate_bwd": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _i, _i, _vp]),
    "ditb200_ln_modulate_bwd_gate": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i, _vp, _vp, _i,
                                          _vp, _i, _i, _i, _vp]),
    "ditb200_gate_resid_bwd": (_i, [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _i, _i, _vp]),
    "ditb200_colsum": (_i, [_vp, _i, _vp, _i, _i, _i, _vp]),
    "ditb200_adaln_wgrad": (_i, [_vp, _i, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ditb200_label_embed_bwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _vp]),
    "ditb200_patchify": (_i, [_vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ditb200_unpatchify_bwd": (_i, [_vp, _vp, _i, _i, _i, _i, _vp]),
    "ditb200_silu_bwd": (_i, [_vp, _vp, _vp, _i, _sz, _vp]),
    "ditb200_gemm": (_i, [C.POINTER(GemmArgs), _vp]),
    "ditb200_cast_bf16": (_i, [_vp, _vp, _sz, _vp]),
    "ditb200_silu_cast": (_i, [_vp, _vp, _i, _sz, _vp]),
    "ditb200_attention_fwd": (_i, [_vp, _vp, _vp, _i, _i, _i, _i, _i, _i, _vp]),
    "ditb200_attention_bwd": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _i, _i, _vp]),
    "ditb200_final_layer": (_i, [_vp, _vp, _vp, _i, _vp, _vp, _vp, _i, _i, _i, _i, _i, _f, _i, _vp]),
    "ditb200_adamw_ema": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _f, _f, _f, _f, _f, _i, _f, _vp]),
    "ditb200_cfg_combine": (_i, [_vp, _vp, _i, _i, _i, _i, _f, _vp]),
    "ditb200_p_sample_step": (_i, [C.POINTER(StepArgs), _vp]),
    "ditb200_q_sample": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _i, _i, _i, _vp]),
    "ditb200_training_losses": (_i, [C.POINTER(LossArgs), _vp]),
    "ditb200_diffusion_affine": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _vp, _i, _i, _i, _vp]),
    "ditb200_prior_bpd": (_i, [_vp, _f, _f, _vp, _i, _i, _vp]),
}

_lock = threading.Lock()
_lib = None
_inited_devices: set[int] = set()


class Ditb200Error(RuntimeError):
    pass


def load() -> C.CDLL:
    """Load libditb200.so and bind every declared entry point.  Raises if absent."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not LIB_PATH.exists():
            raise Ditb200Error(
                f"{LIB_PATH} is missing: build it with `python -m fast_dit_b200.build` "
                "(there is no CPU or PyTorch fallback for the denoiser path)"
            )
        lib = C.CDLL(os.fspath(LIB_PATH))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if lib.ditb200_abi_version() != ABI_VERSION:
            raise Ditb200Error("libditb200 ABI version mismatch; rebuild")
        _lib = lib
    return _lib


def ensure_init(device_index: int) -> C.CDLL:
    lib = load()
    if device_index not in _inited_devices:
        with _lock:
            rc = lib.ditb200_init(int(device_index))
            if rc != 0:
                raise Ditb200Error(f"ditb200_init({device_index}) failed: rc={rc}: "
                                   f"{lib.ditb200_last_error().decode()}")
            _inited_devices.add(device_index)
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().ditb200_last_error().decode(errors="replace")
        raise Ditb200Error(f"{what} failed (rc={rc}): {msg}")
