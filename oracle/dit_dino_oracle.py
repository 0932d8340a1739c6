"""oracle/dit_dino_oracle.py — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

CPU restatement (functional, torch fp32) of the FORK's denoiser variant, /root/reference/models.py (cited FK:line):
the DiT of train_options/models_original.py plus a DINO-feature cross-attention branch.

    DiT.forward(x, t, dino_feat, y)                                   FK:733-754
      c = t_embedder(t)            (the label embedding is computed but NOT added: FK:743)
      dino_tokens = dino_embedder(dino_feat)   PatchEmbed(dino_feat_size -> D), no pos_embed          FK:652,744
      blocks: 9-chunk adaLN (shift/scale/gate for msa, mca, mlp)                                     FK:585,593
              x += gate_msa * attn(modulate(norm1(x), ...))                                          FK:594
              x += gate_mca * cross_atten(modulate(norm3(x), ...), dino_tokens)   only in the 14th and 16th
                                                                                  block (counter, FK:746-750)
              x += gate_mlp * mlp(modulate(norm2(x), ...))                                           FK:596
    CrossAttention.forward(x, context)                                                               FK:534-567
      q = LayerNorm_affine(x)  (no q projection); k, v = split(Linear(D -> 2D, bias=False)(context));
      k = LayerNorm_affine(k); heads = channel groups; softmax(q k^T * hd^-0.5) v; Linear(D -> D) + bias;
      both dropouts (p = 0.2) are inactive in eval mode, which is what this oracle states.

Parity pin: tests/golden/dit_fork_*.npz, written by tools/gen_golden.py from the UNMODIFIED fork module (imported
with empty stand-ins for umap / cv2 / matplotlib and oracle/timm_standin for timm); tests/test_oracle_golden.py
checks this file against them.  The fork's own forward_with_cfg (FK:756-772) cannot run — it calls
forward(combined, t, y) without dino_feat — so only forward() is pinned.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from .dit_oracle import DiTConfig, _ln, _modulate, timestep_embedding

CROSS_BLOCKS = (13, 15)  # 0-based indices of the 14th and 16th block (FK:746-747)


def _ln_affine(x, w, b):  # nn.LayerNorm(dim, eps=1e-6) with weight and bias   FK:517-518
    return F.layer_norm(x, (x.shape[-1],), w, b, eps=1e-6)


def cross_attention(sd: dict, pre: str, x: torch.Tensor, context: torch.Tensor, heads: int) -> torch.Tensor:
    """CrossAttention.forward (FK:534-567), eval mode."""
    N, T, D = x.shape
    hd = D // heads
    q = _ln_affine(x, sd[pre + "norm_q.weight"], sd[pre + "norm_q.bias"])                      # FK:541
    k, v = F.linear(context, sd[pre + "linear.weight"]).chunk(2, dim=2)                        # FK:543-546
    k = _ln_affine(k, sd[pre + "norm_k.weight"], sd[pre + "norm_k.bias"])                      # FK:547
    q, k, v = (z.reshape(N, -1, heads, hd).transpose(1, 2) for z in (q, k, v))                 # FK:550
    attn = (torch.einsum("bhid,bhjd->bhij", q, k) * hd ** -0.5).softmax(dim=-1)                # FK:553-560
    out = torch.einsum("bhij,bhjd->bhid", attn, v).transpose(1, 2).reshape(N, T, D)            # FK:566-567
    return F.linear(out, sd[pre + "to_out.0.weight"], sd[pre + "to_out.0.bias"])               # FK:569


def dit_dino_forward(sd: dict, cfg: DiTConfig, x: torch.Tensor, t: torch.Tensor, dino_feat: torch.Tensor,
                     y: torch.Tensor | None = None) -> torch.Tensor:
    """DiT.forward(x, t, dino_feat, y) of the fork (FK:733-754); y does not influence the output (FK:743)."""
    p, D, H = cfg.patch_size, cfg.hidden_size, cfg.num_heads
    hd = D // H
    h = F.conv2d(x, sd["x_embedder.proj.weight"], sd["x_embedder.proj.bias"], stride=p)
    h = h.flatten(2).transpose(1, 2) + sd["pos_embed"]                                         # FK:740
    N, T, _ = h.shape
    te = timestep_embedding(t, 256)
    te = F.linear(te, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    c = F.linear(F.silu(te), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])      # FK:741,743
    dino = F.conv2d(dino_feat, sd["dino_embedder.proj.weight"], sd["dino_embedder.proj.bias"], stride=p)
    dino = dino.flatten(2).transpose(1, 2)                                                     # FK:744
    sc = F.silu(c)
    for i in range(cfg.depth):
        pre = f"blocks.{i}."
        mod = F.linear(sc, sd[pre + "adaLN_modulation.1.weight"], sd[pre + "adaLN_modulation.1.bias"])
        sh_a, s_a, g_a, sh_c, s_c, g_c, sh_m, s_m, g_m = mod.chunk(9, dim=1)                   # FK:585,593
        a = _modulate(_ln(h), sh_a, s_a)
        qkv = F.linear(a, sd[pre + "attn.qkv.weight"], sd[pre + "attn.qkv.bias"])
        q, k, v = qkv.reshape(N, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
        a = F.scaled_dot_product_attention(q, k, v).transpose(1, 2).reshape(N, T, D)
        a = F.linear(a, sd[pre + "attn.proj.weight"], sd[pre + "attn.proj.bias"])
        h = h + g_a[:, None, :] * a                                                            # FK:594
        if i in CROSS_BLOCKS:
            ca = cross_attention(sd, pre + "cross_atten.", _modulate(_ln(h), sh_c, s_c), dino, H)
            h = h + g_c[:, None, :] * ca                                                       # FK:595
        m = _modulate(_ln(h), sh_m, s_m)
        m = F.gelu(F.linear(m, sd[pre + "mlp.fc1.weight"], sd[pre + "mlp.fc1.bias"]), approximate="tanh")
        m = F.linear(m, sd[pre + "mlp.fc2.weight"], sd[pre + "mlp.fc2.bias"])
        h = h + g_m[:, None, :] * m                                                            # FK:596
    mod = F.linear(sc, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    sh, s = mod.chunk(2, dim=1)
    h = F.linear(_modulate(_ln(h), sh, s), sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])
    g, co = cfg.grid, cfg.out_channels
    h = h.reshape(N, g, g, p, p, co).permute(0, 5, 1, 3, 2, 4)
    return h.reshape(N, co, g * p, g * p)
