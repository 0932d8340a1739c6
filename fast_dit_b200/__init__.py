"""fast_dit_b200 — the fast-DiT denoiser hot path on B200 (sm_100a).

Same Python surface as the reference (alexandor91/fast-DiT, upstream-form model):

    from fast_dit_b200 import DiT_models, create_diffusion
    model = DiT_models["DiT-XL/2"](input_size=32, num_classes=1000).cuda().eval()
    diffusion = create_diffusion("250")
    samples = diffusion.p_sample_loop(model.forward_with_cfg, z.shape, z, clip_denoised=False,
                                      model_kwargs=dict(y=y, cfg_scale=4.0), device="cuda")

All arithmetic on that path runs in libditb200.so (hand-written CUDA behind a C ABI,
include/ditb200.h).  There is no PyTorch or CPU fallback: without the library, or without an
sm_100 GPU, calls raise.
"""
from .models import DiT, DiT_models  # noqa: F401
from .diffusion import create_diffusion  # noqa: F401

__all__ = ["DiT", "DiT_models", "create_diffusion"]
__version__ = "0.1.0"
