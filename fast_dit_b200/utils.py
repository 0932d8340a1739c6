"""Small host-side helpers shared by bench.py, smoke() and the examples."""
from __future__ import annotations

import torch


def rerandomise_zero_params(model, seed: int = 1234, std: float = 0.02) -> int:
    """Benchmark/parity weight protocol: adaLN-Zero initialises every gate and the
    output layer to zero, so a freshly constructed DiT returns exactly 0 for any input.  Re-draw
    every all-zero parameter from N(0, std^2) with a dedicated CPU generator, in
    named_parameters() order, so random-init runs exercise the whole network."""
    g = torch.Generator().manual_seed(seed)
    n = 0
    with torch.no_grad():
        for _, p in model.named_parameters():
            if p.numel() and float(p.detach().abs().max()) == 0.0:
                p.copy_(torch.randn(p.shape, generator=g, dtype=torch.float32).to(p.device) * std)
                n += 1
    return n


def forward_flops_per_image(model) -> float:
    """Algorithmic forward FLOPs per image (2 per MAC, contractions only)."""
    L, D, T = model.depth, model.hidden_size, model.x_embedder.num_patches
    C, p, co = model.in_channels, model.patch_size, model.out_channels
    return (L * (24 * T * D * D + 4 * T * T * D + 12 * D * D) + 2 * T * C * p * p * D
            + (2 * 256 * D + 2 * D * D) + (4 * D * D + 2 * T * D * p * p * co))
