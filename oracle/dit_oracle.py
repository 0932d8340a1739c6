"""CPU oracle for the DiT denoiser — TEST INFRASTRUCTURE, not a product path.

Functional restatement of /root/reference/train_options/models_original.py (cited
as MO:line) and of timm==0.9.16's PatchEmbed / Attention / Mlp (environment.yml:139,
semantics per SURVEY.md §8c; the layout is evidenced by
performance/A100/train_original.out:36-37).  It works on a plain state_dict with the
reference's key names, so it can be driven by the reference's
own module, by the product module, or by a checkpoint.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn.functional as F


@dataclass(frozen=True)
class DiTConfig:
    input_size: int = 32
    patch_size: int = 2
    in_channels: int = 4
    hidden_size: int = 1152
    depth: int = 28
    num_heads: int = 16
    mlp_ratio: float = 4.0
    class_dropout_prob: float = 0.1
    num_classes: int = 1000
    learn_sigma: bool = True

    @property
    def out_channels(self) -> int:  # MO:165
        return self.in_channels * 2 if self.learn_sigma else self.in_channels

    @property
    def grid(self) -> int:
        return self.input_size // self.patch_size

    @property
    def num_patches(self) -> int:
        return self.grid * self.grid


# name -> (depth, hidden, patch, heads)   MO:328-370
CONFIGS = {
    f"DiT-{n}/{p}": (d, h, p, nh)
    for n, (d, h, nh) in {"XL": (28, 1152, 16), "L": (24, 1024, 16), "B": (12, 768, 12), "S": (12, 384, 6)}.items()
    for p in (2, 4, 8)
}


def config_for(name: str, **kw) -> DiTConfig:
    d, h, p, nh = CONFIGS[name]
    return DiTConfig(depth=d, hidden_size=h, patch_size=p, num_heads=nh, **kw)


# ------------------------------------------------------------------ embeddings
def sincos_1d(dim: int, pos: np.ndarray) -> np.ndarray:
    """MO:302-321: [sin | cos] of pos * 10000^(-k/(dim/2)), fp64."""
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    ang = pos.reshape(-1)[:, None] * omega[None, :]
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def sincos_pos_embed(dim: int, grid: int) -> np.ndarray:
    """MO:274-299.  meshgrid(w, h) puts the column coordinate first, so dims [0, dim/2)
    encode the column and [dim/2, dim) the row; token index = row*grid + col."""
    ar = np.arange(grid, dtype=np.float32)
    col, row = np.meshgrid(ar, ar)  # col[i, j] = j, row[i, j] = i
    return np.concatenate([sincos_1d(dim // 2, col), sincos_1d(dim // 2, row)], axis=1)


def timestep_embedding(t: torch.Tensor, dim: int = 256, max_period: float = 10000.0) -> torch.Tensor:
    """MO:40-59: cos half first, fp32 throughout."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    ang = t[:, None].float() * freqs[None]
    emb = torch.cat([ang.cos(), ang.sin()], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


# --------------------------------------------------------------------- forward
def _modulate(x, shift, scale):  # MO:19-20
    return x * (1 + scale[:, None, :]) + shift[:, None, :]


def _ln(x):  # nn.LayerNorm(elementwise_affine=False, eps=1e-6)   MO:107,109,131
    return F.layer_norm(x, (x.shape[-1],), eps=1e-6)


def dit_forward(sd: dict, cfg: DiTConfig, x: torch.Tensor, t: torch.Tensor, y: torch.Tensor,
                drop_ids: torch.Tensor | None = None) -> torch.Tensor:
    """DiT.forward (MO:233-248).  `drop_ids` (bool [N]) plays the role of the training-mode
    label dropout mask (MO:79-87); None = eval mode."""
    p, D, H = cfg.patch_size, cfg.hidden_size, cfg.num_heads
    hd = D // H
    # PatchEmbed: Conv2d(k=s=p) -> flatten(2).transpose(1, 2); + pos_embed      MO:169,240
    h = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p)
    h = h.flatten(2).transpose(1, 2) + sd["pos_embed"]
    N, T, _ = h.shape
    # TimestepEmbedder: Linear -> SiLU -> Linear                                  MO:33-37,61-64
    te = timestep_embedding(t, 256)
    te = F.linear(te, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    te = F.linear(F.silu(te), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    # LabelEmbedder                                                               MO:79-94
    labels = y
    if drop_ids is not None:
        labels = torch.where(drop_ids, torch.full_like(y, cfg.num_classes), y)
    c = te + sd["y_embedder.embedding_table.weight"][labels]                     # MO:243
    sc = F.silu(c)
    for i in range(cfg.depth):
        pre = f"blocks.{i}."
        mod = F.linear(sc, sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
        sh1, s1, g1, sh2, s2, g2 = mod.chunk(6, dim=1)                            # MO:119
        # timm Attention: fused qkv, [B,N,3,H,hd] -> q,k,v [B,H,N,hd], SDPA, proj
        a = _modulate(_ln(h), sh1, s1)
        qkv = F.linear(a, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
        q, k, v = qkv.reshape(N, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(N, T, D)
        a = F.linear(a, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
        h = h + g1[:, None, :] * a                                                # MO:120
        # timm Mlp: fc1 -> GELU(tanh) -> fc2
        m = _modulate(_ln(h), sh2, s2)
        m = F.gelu(F.linear(m, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]), approximate="tanh")
        m = F.linear(m, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
        h = h + g2[:, None, :] * m                                                # MO:121
    # FinalLayer                                                                   MO:138-142
    mod = F.linear(sc, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    sh, s = mod.chunk(2, dim=1)
    h = F.linear(_modulate(_ln(h), sh, s), sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    # unpatchify: (n, h, w, p, q, c) -> (n, c, h*p, w*q)                          MO:218-231
    g, co = cfg.grid, cfg.out_channels
    h = h.reshape(N, g, g, p, p, co).permute(0, 5, 1, 3, 2, 4)
    return h.reshape(N, co, g * p, g * p)


def dit_forward_with_cfg(sd, cfg, x, t, y, cfg_scale: float) -> torch.Tensor:
    """MO:250-266: first half of x duplicated; guidance on eps channels 0..2 only."""
    n = x.shape[0] // 2
    half = x[:n]
    out = dit_forward(sd, cfg, torch.cat([half, half], 0), t, y)
    eps, rest = out[:, :3], out[:, 3:]
    cond, uncond = eps[:n], eps[n:]
    g = uncond + cfg_scale * (cond - uncond)
    return torch.cat([torch.cat([g, g], 0), rest], dim=1)


# ------------------------------------------------------------ weight protocol
def rerandomise_zero_params(named_params, seed: int = 1234, std: float = 0.02) -> int:
    """SURVEY.md §0.5/§8c protocol: adaLN-Zero leaves every gate and the final linear at 0, so a
    fresh DiT outputs exactly 0.  Re-draw every all-zero parameter from N(0, std^2) with a
    dedicated CPU generator, in named_parameters() order.  Returns how many were re-drawn."""
    g = torch.Generator().manual_seed(seed)
    n = 0
    with torch.no_grad():
        for _, prm in named_params:
            if prm.numel() and float(prm.detach().abs().max()) == 0.0:
                prm.copy_(torch.randn(prm.shape, generator=g, dtype=torch.float32) * std)
                n += 1
    return n


def flops_per_image(cfg: DiTConfig) -> float:
    """Algorithmic forward FLOPs (2 per MAC), contractions only — BASELINE.md §2."""
    L, D, T = cfg.depth, cfg.hidden_size, cfg.num_patches
    C, p = cfg.in_channels, cfg.patch_size
    return (L * (24 * T * D * D + 4 * T * T * D + 12 * D * D) + 2 * T * C * p * p * D
            + (2 * 256 * D + 2 * D * D) + (4 * D * D + 2 * T * D * p * p * cfg.out_channels))
