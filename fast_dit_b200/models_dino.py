"""The fork's DINO cross-attention DiT on libditb200 — inference.

Drop-in for the denoiser in /root/reference/models.py (cited FK:line): the upstream DiT plus
    * `dino_embedder`, a second PatchEmbed that turns DINO feature maps [N, dino_feat_size, H, W] into tokens (FK:652);
    * a 9-chunk adaLN per block (shift / scale / gate for self-attention, cross-attention and the MLP: FK:585,593);
    * `cross_atten` (FK:506-567) — LayerNorm(affine) on the modulated tokens as the query (no q projection), one
      bias-free Linear(D -> 2D) producing k | v from the DINO tokens, LayerNorm(affine) on k, per-head softmax
      attention, Linear(D -> D) — gated into the stream in the 14th and 16th block only (FK:746-750);
    * the conditioning vector is the timestep embedding alone: the label embedding is evaluated and dropped (FK:743).
Same registry keys, constructor kwargs (+ `dino_feat_size`), `state_dict` layout (286 keys for 16 blocks) and — built
under the same torch seed — bit-identical initial weights as the fork's `DiT`.

`forward(x, t, dino_feat, y)` keeps the fork's positional signature.  The fork's own `forward_with_cfg` cannot run
(FK:763 calls `self.forward(combined, t, y)` without dino_feat); here it takes the features as a keyword,
`forward_with_cfg(x, t, y, cfg_scale, dino_feat=...)`, so `p_sample_loop(model.forward_with_cfg, ...,
model_kwargs=dict(y=y, cfg_scale=s, dino_feat=f))` works with the reference's calling convention.

Every kernel is one the upstream path already uses: the DINO patch-embed is im2col + the tcgen05 GEMM, the two affine
LayerNorms are `ditb200_ln_modulate` with (weight - 1, bias) as the modulation, cross-attention runs through
`ditb200_attention_fwd` on a packed [q | k | v] matrix, and its output projection uses the gated-residual epilogue.
Training through this variant (dropout 0.2 inside the branch, gradient checkpointing of every block: FK:748,750) is
not built: `forward` with gradients enabled raises.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .models import (DiT as _BaseDiT, _AttentionParams, _FinalLayerParams, _LabelEmbedderParams, _MlpParams,
                     _PatchEmbedParams, _TimestepEmbedderParams, get_2d_sincos_pos_embed)

CROSS_BLOCKS = (13, 15)  # 0-based: "counter == 14 or counter == 16" (FK:746-747)


class _CrossAttentionParams(nn.Module):
    """Parameter holder with the registration order of the fork's CrossAttention (FK:507-531)."""

    def __init__(self, dim, heads):
        super().__init__()
        self.heads = heads
        self.norm_q = nn.LayerNorm(dim, eps=1e-6)
        self.norm_k = nn.LayerNorm(dim, eps=1e-6)
        self.linear = nn.Linear(dim, dim * 2, bias=False)
        self.attn_drop = nn.Dropout(0.2)
        self.to_out = nn.Sequential(nn.Linear(dim, dim), nn.Dropout(0.2))


class _DinoBlockParams(nn.Module):
    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0):  # FK:574-589
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.attn = _AttentionParams(hidden_size, num_heads)
        self.norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.norm3 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp = _MlpParams(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 9 * hidden_size, bias=True))
        self.cross_atten = _CrossAttentionParams(hidden_size, num_heads)


class DiT(nn.Module):
    """Diffusion Transformer with a DINO cross-attention branch (FK:624-772) on libditb200."""

    def __init__(self, input_size=32, patch_size=2, in_channels=4, hidden_size=1152, dino_feat_size=768, depth=28,
                 num_heads=16, mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000, learn_sigma=True,
                 precision="bf16"):
        super().__init__()
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.hidden_size = hidden_size
        self.depth = depth
        self.input_size = input_size
        self.precision = precision
        self.counter = 0  # attribute of the fork's module (FK:650); the block choice here does not mutate state
        # registration order of FK:653-666: construction draws from the global RNG in this order
        self.x_embedder = _PatchEmbedParams(input_size, patch_size, in_channels, hidden_size, bias=True)
        self.dino_embedder = _PatchEmbedParams(input_size, patch_size, dino_feat_size, hidden_size, bias=True)
        self.t_embedder = _TimestepEmbedderParams(hidden_size)
        self.y_embedder = _LabelEmbedderParams(num_classes, hidden_size, class_dropout_prob)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.x_embedder.num_patches, hidden_size), requires_grad=False)
        self.blocks = nn.ModuleList([_DinoBlockParams(hidden_size, num_heads, mlp_ratio) for _ in range(depth)])
        self.final_layer = _FinalLayerParams(hidden_size, patch_size, self.out_channels)
        self.initialize_weights()
        self._shadow = {}

    def initialize_weights(self):
        """Same draws, in the same order, as FK:675-712."""
        def basic(m):
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self.apply(basic)
        grid = int(self.x_embedder.num_patches ** 0.5)
        self.pos_embed.data.copy_(torch.from_numpy(get_2d_sincos_pos_embed(self.pos_embed.shape[-1], grid)).float().unsqueeze(0))
        for emb in (self.x_embedder, self.dino_embedder):
            w = emb.proj.weight.data
            nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
            nn.init.constant_(emb.proj.bias, 0)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for blk in self.blocks:
            nn.init.constant_(blk.adaLN_modulation[-1].weight, 0)
            nn.init.constant_(blk.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].weight, 0)
        nn.init.constant_(self.final_layer.adaLN_modulation[-1].bias, 0)
        nn.init.constant_(self.final_layer.linear.weight, 0)
        nn.init.constant_(self.final_layer.linear.bias, 0)

    unpatchify = _BaseDiT.unpatchify

    # ------------------------------------------------------------------ weight shadows
    def _cross_blocks(self):
        return [i for i in CROSS_BLOCKS if i < self.depth]

    def _shadows(self):
        """bf16 copies of every GEMM weight + the concatenated adaLN matrix, rebuilt when a parameter changes."""
        ada_w = [b.adaLN_modulation[1].weight for b in self.blocks] + [self.final_layer.adaLN_modulation[1].weight]
        ada_b = [b.adaLN_modulation[1].bias for b in self.blocks] + [self.final_layer.adaLN_modulation[1].bias]
        gw = []
        for b in self.blocks:
            gw += [b.attn.qkv.weight, b.attn.proj.weight, b.mlp.fc1.weight, b.mlp.fc2.weight]
        cw = []
        for i in self._cross_blocks():
            ca = self.blocks[i].cross_atten
            cw += [ca.linear.weight, ca.to_out[0].weight, ca.norm_q.weight, ca.norm_k.weight]
        every = gw + ada_w + ada_b + cw + [self.dino_embedder.proj.weight]
        key = (self.precision, ada_w[0].device, tuple(p._version for p in every), tuple(p.data_ptr() for p in every[:3]))
        if self._shadow.get("key") == key:
            return self._shadow
        bf16 = self.precision == "bf16"
        cast = (lambda w: ops.cast_bf16(w.detach().contiguous())) if bf16 else (lambda w: w.detach().contiguous())
        with torch.no_grad():
            sh = {"key": key, "ada_b": torch.cat([b.detach() for b in ada_b]).contiguous(),
                  "ada_w": cast(torch.cat([w.detach() for w in ada_w], dim=0)), "w": [cast(w) for w in gw], "cross": {}}
            D = self.hidden_size
            sh["dino_w"] = cast(self.dino_embedder.proj.weight.detach().reshape(D, -1))
            for i in self._cross_blocks():
                ca = self.blocks[i].cross_atten
                wkv = ca.linear.weight.detach()
                sh["cross"][i] = {"wk": cast(wkv[:D]), "wv": cast(wkv[D:]), "wo": cast(ca.to_out[0].weight),
                                  # LayerNorm(x) * w + b  ==  LN-modulate with scale = w - 1, shift = b
                                  "q_scale": (ca.norm_q.weight.detach() - 1.0).contiguous(),
                                  "k_scale": (ca.norm_k.weight.detach() - 1.0).contiguous()}
        self._shadow = sh
        return sh

    def _apply(self, fn, *a, **k):
        self._shadow = {}
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------------ forward
    def conditioning(self, t):
        """c = t_embedder(t) (FK:741,743): no label term."""
        te = self.t_embedder
        h = ops.small_linear(ops.timestep_embedding(t, te.frequency_embedding_size), te.mlp[0].weight, te.mlp[0].bias,
                             silu_out=True)
        return ops.small_linear(h, te.mlp[2].weight, te.mlp[2].bias)

    def dino_tokens(self, dino_feat, sh):
        """dino_embedder(dino_feat) (FK:744): Conv2d(k = s = p) as im2col + GEMM; no positional table is added."""
        p, D = self.patch_size, self.hidden_size
        bf16 = self.precision == "bf16"
        cols = ops.patchify(dino_feat.float().contiguous(), p, out_dtype=torch.bfloat16 if bf16 else torch.float32)
        return ops.gemm(cols, sh["dino_w"], self.dino_embedder.proj.bias)

    def forward(self, x, t, dino_feat, y=None):
        """x [N, C, H, W], t [N], dino_feat [N, dino_feat_size, H, W], y [N] (ignored by the arithmetic, FK:743)."""
        if not x.is_cuda:
            raise L.Ditb200Error("fast_dit_b200 runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise L.Ditb200Error("the DINO cross-attention variant is inference-only here: call it under torch.no_grad() "
                                 "(training through the fork's checkpointed blocks is not built)")
        return self._forward_inference(x, t, dino_feat)

    @torch.no_grad()
    def _forward_inference(self, x, t, dino_feat):
        sh = self._shadows()
        bf16 = self.precision == "bf16"
        act = torch.bfloat16 if bf16 else torch.float32
        D, Hh, p = self.hidden_size, self.num_heads, self.patch_size
        hd = D // Hh
        x = x.float().contiguous()
        N = x.shape[0]
        T = (x.shape[2] // p) * (x.shape[3] // p)
        if self.pos_embed.shape[1] != T:
            raise L.Ditb200Error(f"input grid gives {T} tokens but pos_embed has {self.pos_embed.shape[1]}")
        if dino_feat.shape[0] != N or tuple(dino_feat.shape[2:]) != tuple(x.shape[2:]):
            raise L.Ditb200Error("dino_feat must be [N, dino_feat_size, H, W] on the latent's grid (the fork embeds it with "
                                 "the same input_size / patch_size, FK:652)")
        tok = ops.patch_embed(x, self.x_embedder.proj.weight, self.x_embedder.proj.bias, self.pos_embed, p)
        c = self.conditioning(t)
        if bf16:
            mod = ops.gemm(ops.silu_cast(c, torch.bfloat16), sh["ada_w"], sh["ada_b"], out_dtype=torch.float32)
        else:
            mod = ops.small_linear(c, sh["ada_w"], sh["ada_b"], silu_in=True)
        cross = self._cross_blocks()
        dino = self.dino_tokens(dino_feat, sh) if cross else None
        w = sh["w"]
        for i, blk in enumerate(self.blocks):
            m = mod[:, i * 9 * D:(i + 1) * 9 * D]
            sh_a, sc_a, g_a, sh_c, sc_c, g_c, sh_m, sc_m, g_m = (m[:, j * D:(j + 1) * D] for j in range(9))
            h = ops.ln_modulate(tok, sh_a, sc_a, T, out_dtype=act)
            qkv = ops.gemm(h, w[4 * i], blk.attn.qkv.bias)
            o = ops.attention(qkv, N, T, Hh, hd)
            ops.gemm(o, w[4 * i + 1], blk.attn.proj.bias, epilogue=L.EPI_BIAS_GATE_RESID, resid=tok, gate=g_a, rows_per_gate=T)
            if i in cross:
                self._cross_attention(tok, dino, blk.cross_atten, sh["cross"][i], sh_c, sc_c, g_c, N, T, act)
            h = ops.ln_modulate(tok, sh_m, sc_m, T, out_dtype=act)
            u = ops.gemm(h, w[4 * i + 2], blk.mlp.fc1.bias, epilogue=L.EPI_BIAS_GELU)
            ops.gemm(u, w[4 * i + 3], blk.mlp.fc2.bias, epilogue=L.EPI_BIAS_GATE_RESID, resid=tok, gate=g_m, rows_per_gate=T)
        mf = mod[:, self.depth * 9 * D:]
        fl = self.final_layer
        return ops.final_layer(tok, mf[:, :D], mf[:, D:], fl.linear.weight, fl.linear.bias, T, p, self.out_channels)

    def _cross_attention(self, tok, dino, ca, shc, shift, scale, gate, N, T, act):
        """tok += gate_mca * cross_atten(modulate(norm3(tok), shift_mca, scale_mca), dino_tokens)   (FK:595, 534-569)."""
        D, Hh = self.hidden_size, self.num_heads
        M = N * T
        xm = ops.ln_modulate(tok, shift, scale, T, out_dtype=torch.float32)            # modulate(norm3(x))
        bq = ca.norm_q.bias.view(1, D).expand(N, D)                                    # one [D] vector for every image
        q = ops.ln_modulate(xm, bq, shc["q_scale"].view(1, D).expand(N, D), T, out_dtype=act)   # norm_q (affine)
        k32 = ops.gemm(dino, shc["wk"], None, out_dtype=torch.float32)                 # linear(context)[..., :D]
        bk = ca.norm_k.bias.view(1, D).expand(N, D)
        k = ops.ln_modulate(k32, bk, shc["k_scale"].view(1, D).expand(N, D), T, out_dtype=act)  # norm_k (affine)
        v = ops.gemm(dino, shc["wv"], None)                                            # linear(context)[..., D:]
        qkv = torch.empty((M, 3 * D), device=tok.device, dtype=act)                    # the packed layout the kernel reads
        qkv[:, :D].copy_(q)
        qkv[:, D:2 * D].copy_(k)
        qkv[:, 2 * D:].copy_(v)
        o = ops.attention(qkv, N, T, Hh, D // Hh)
        ops.gemm(o, shc["wo"], ca.to_out[0].bias, epilogue=L.EPI_BIAS_GATE_RESID, resid=tok, gate=gate, rows_per_gate=T)

    def forward_with_cfg(self, x, t, y, cfg_scale, dino_feat=None):
        """Classifier-free guidance forward (FK:756-772) with the features the fork's version forgot to pass."""
        if dino_feat is None:
            raise TypeError("forward_with_cfg needs dino_feat= (the fork's own signature cannot run: models.py:763)")
        n = x.shape[0] // 2
        half = x[:n]
        raw = self.forward(torch.cat([half, half], dim=0), t, dino_feat, y)
        return ops.cfg_combine(raw.contiguous(), 3, float(cfg_scale))


def _factory(depth, hidden_size, patch_size, num_heads):
    def make(**kwargs):
        return DiT(depth=depth, hidden_size=hidden_size, patch_size=patch_size, num_heads=num_heads, **kwargs)
    return make


_SIZES = {"XL": (28, 1152, 16), "L": (24, 1024, 16), "B": (12, 768, 12), "S": (12, 384, 6)}
DiT_models = {f"DiT-{n}/{p}": _factory(d, h, p, nh) for n, (d, h, nh) in _SIZES.items() for p in (2, 4, 8)}  # FK:871-876
