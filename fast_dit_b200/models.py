"""DiT denoiser with the reference's Python surface, executed by libditb200 (sm_100a CUDA).

Drop-in for /root/reference/train_options/models_original.py (cited as MO:line): the same
`DiT_models` registry and factory kwargs (MO:328-370), `DiT.forward(x, t, y)` (MO:233),
`DiT.forward_with_cfg(x, t, y, cfg_scale)` (MO:250), the same attributes
(`in_channels`, `out_channels`, `patch_size`, `num_heads`, `learn_sigma`,
`x_embedder.num_patches`, `y_embedder.num_classes`) and the same `state_dict` keys and
shapes, so released checkpoints and train.py checkpoints load.

The nn.Module tree below only *holds parameters* (in the reference's registration order, so
that constructing under the same torch seed yields bit-identical weights); no submodule has a
forward of its own.  All arithmetic happens in `DiT.forward`, which sequences C-ABI kernels:
patch-embed, timestep/label embedders, one batched adaLN GEMM for all blocks, and per block
LayerNorm+modulate -> QKV GEMM -> fused attention -> out-proj GEMM with gated-residual
epilogue -> LayerNorm+modulate -> fc1 GEMM with GELU epilogue -> fc2 GEMM with gated-residual
epilogue; then the fused final layer (+unpatchify).  There is no PyTorch fallback.
"""
from __future__ import annotations

import math
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib as L
from . import ops

PRECISIONS = ("bf16", "fp32")
_BRANCH_MODE = int(os.environ.get("DITB200_INFER_BRANCH", "2"))
_ZIGZAG = os.environ.get("DITB200_ZIGZAG", "1") != "0"


# --------------------------------------------------------------- parameter holders
class _PatchEmbedParams(nn.Module):
    """Holds `proj` (Conv2d weight layout [D, C, p, p]) like timm's PatchEmbed (MO:169)."""

    def __init__(self, img_size, patch_size, in_chans, embed_dim, bias=True):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.patch_size = (patch_size, patch_size)
        self.grid_size = (img_size // patch_size, img_size // patch_size)
        self.num_patches = self.grid_size[0] * self.grid_size[1]
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=patch_size, bias=bias)


class _TimestepEmbedderParams(nn.Module):
    def __init__(self, hidden_size, frequency_embedding_size=256):  # MO:31-38
        super().__init__()
        self.mlp = nn.Sequential(
            nn.Linear(frequency_embedding_size, hidden_size, bias=True),
            nn.SiLU(),
            nn.Linear(hidden_size, hidden_size, bias=True),
        )
        self.frequency_embedding_size = frequency_embedding_size


class _LabelEmbedderParams(nn.Module):
    def __init__(self, num_classes, hidden_size, dropout_prob):  # MO:71-77
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + int(dropout_prob > 0), hidden_size)
        self.num_classes = num_classes
        self.dropout_prob = dropout_prob

    def token_drop(self, labels, force_drop_ids=None):
        """Label dropout for classifier-free guidance (MO:79-87).  Stays on the host side of the
        boundary because it draws from torch's RNG stream, which parity requires to be shared."""
        if force_drop_ids is None:
            drop = torch.rand(labels.shape[0], device=labels.device) < self.dropout_prob
        else:
            drop = force_drop_ids == 1
        return torch.where(drop, self.num_classes, labels)


class _AttentionParams(nn.Module):
    def __init__(self, dim, num_heads):  # timm Attention(dim, num_heads, qkv_bias=True)
        super().__init__()
        self.num_heads = num_heads
        self.head_dim = dim // num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _MlpParams(nn.Module):
    def __init__(self, dim, hidden):  # timm Mlp(in, hidden, act=GELU(tanh))
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _BlockParams(nn.Module):
    def __init__(self, hidden_size, num_heads, mlp_ratio=4.0):  # MO:105-116
        super().__init__()
        self.norm1 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.attn = _AttentionParams(hidden_size, num_heads)
        self.norm2 = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.mlp = _MlpParams(hidden_size, int(hidden_size * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 6 * hidden_size, bias=True))


class _FinalLayerParams(nn.Module):
    def __init__(self, hidden_size, patch_size, out_channels):  # MO:129-136
        super().__init__()
        self.norm_final = nn.LayerNorm(hidden_size, elementwise_affine=False, eps=1e-6)
        self.linear = nn.Linear(hidden_size, patch_size * patch_size * out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden_size, 2 * hidden_size, bias=True))


def _sincos_1d(dim, pos):
    omega = 1.0 / 10000 ** (np.arange(dim // 2, dtype=np.float64) / (dim / 2.0))
    ang = np.outer(pos.reshape(-1), omega)
    return np.concatenate([np.sin(ang), np.cos(ang)], axis=1)


def get_2d_sincos_pos_embed(embed_dim, grid_size):
    """Frozen 2-D sin-cos table (MO:274-321): first half of the channels encodes the column,
    second half the row; each half is [sin | cos]; fp64."""
    ar = np.arange(grid_size, dtype=np.float32)
    col, row = np.meshgrid(ar, ar)
    return np.concatenate([_sincos_1d(embed_dim // 2, col), _sincos_1d(embed_dim // 2, row)], axis=1)


# ----------------------------------------------------------------------- the model
class DiT(nn.Module):
    """Diffusion Transformer denoiser (MO:145-266) on libditb200."""

    def __init__(self, input_size=32, patch_size=2, in_channels=4, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000, learn_sigma=True, precision="bf16"):
        super().__init__()
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {PRECISIONS}")
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.out_channels = in_channels * 2 if learn_sigma else in_channels
        self.patch_size = patch_size
        self.num_heads = num_heads
        self.hidden_size = hidden_size
        self.depth = depth
        self.input_size = input_size
        self.precision = precision

        self.x_embedder = _PatchEmbedParams(input_size, patch_size, in_channels, hidden_size, bias=True)
        self.t_embedder = _TimestepEmbedderParams(hidden_size)
        self.y_embedder = _LabelEmbedderParams(num_classes, hidden_size, class_dropout_prob)
        self.pos_embed = nn.Parameter(torch.zeros(1, self.x_embedder.num_patches, hidden_size), requires_grad=False)
        self.blocks = nn.ModuleList([_BlockParams(hidden_size, num_heads, mlp_ratio) for _ in range(depth)])
        self.final_layer = _FinalLayerParams(hidden_size, patch_size, self.out_channels)
        self.initialize_weights()
        self._shadow = {}  # bf16 / concatenated copies of GEMM weights, keyed by parameter versions

    def initialize_weights(self):
        """Same draws, in the same order, as MO:182-216: xavier on every Linear, sin-cos pos table,
        xavier on the flattened patch-embed kernel, N(0, 0.02) embeddings, zeroed adaLN/output."""
        def basic(m):
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
        self.apply(basic)
        grid = int(self.x_embedder.num_patches ** 0.5)
        pe = get_2d_sincos_pos_embed(self.pos_embed.shape[-1], grid)
        self.pos_embed.data.copy_(torch.from_numpy(pe).float().unsqueeze(0))
        w = self.x_embedder.proj.weight.data
        nn.init.xavier_uniform_(w.view([w.shape[0], -1]))
        nn.init.constant_(self.x_embedder.proj.bias, 0)
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

    # ------------------------------------------------------------------ weight shadows
    def _ada_params(self):
        ws = [b.adaLN_modulation[1].weight for b in self.blocks] + [self.final_layer.adaLN_modulation[1].weight]
        bs = [b.adaLN_modulation[1].bias for b in self.blocks] + [self.final_layer.adaLN_modulation[1].bias]
        return ws, bs

    def _gemm_weights(self):
        out = []
        for b in self.blocks:
            out += [b.attn.qkv.weight, b.attn.proj.weight, b.mlp.fc1.weight, b.mlp.fc2.weight]
        return out

    def _shadows(self):
        """bf16 copies of the GEMM weights + the concatenated adaLN matrix, rebuilt whenever a
        parameter changes (optimizer step, load_state_dict, .to())."""
        flat = getattr(self, "_flat", None)
        if flat is not None:  # optim.FusedAdamWEMA owns the parameters and keeps the shadows in step
            return flat.shadows()
        ada_w, ada_b = self._ada_params()
        gw = self._gemm_weights()
        key = (self.precision, ada_w[0].device, tuple(p._version for p in gw + ada_w + ada_b),
               tuple(p.data_ptr() for p in gw[:2] + ada_w[:1]))
        if self._shadow.get("key") == key:
            return self._shadow
        with torch.no_grad():
            sh = {"key": key}
            ada_w_all = torch.cat([w.detach() for w in ada_w], dim=0).contiguous()
            sh["ada_b"] = torch.cat([b.detach() for b in ada_b], dim=0).contiguous()
            if self.precision == "bf16":
                sh["ada_w"] = ops.cast_bf16(ada_w_all)
                sh["w"] = [ops.cast_bf16(w.detach().contiguous()) for w in gw]
            else:
                sh["ada_w"] = ada_w_all
                sh["w"] = [w.detach() for w in gw]
        self._shadow = sh
        return sh

    def _apply(self, fn, *a, **k):  # .to()/.cuda() move parameters: drop stale shadows and arenas
        self._shadow = {}
        self._flat = self._grad_arena = self._layout = None
        return super()._apply(fn, *a, **k)

    # ------------------------------------------------------------------------ forward
    def unpatchify(self, x):
        """[N, T, p*p*C] -> [N, C, H, W] (MO:218-231).  The fused final-layer kernel writes NCHW
        directly; this host-side view exists for API compatibility with reference callers."""
        c, p = self.out_channels, self.patch_size
        h = w = int(x.shape[1] ** 0.5)
        x = x.reshape(x.shape[0], h, w, p, p, c).permute(0, 5, 1, 3, 2, 4)
        return x.reshape(x.shape[0], c, h * p, w * p)

    def forward(self, x, t, y):
        """x: [N, C, H, W] latents, t: [N] timesteps, y: [N] labels -> [N, out_channels, H, W]."""
        if not x.is_cuda:
            raise L.Ditb200Error("fast_dit_b200.DiT runs on CUDA (sm_100a) only; move the model and inputs to the GPU")
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            from .autograd import dit_forward_autograd
            return dit_forward_autograd(self, x, t, y)
        return self._forward_inference(x, t, y)

    def _check_labels(self, y):
        """nn.Embedding raises on an out-of-range index (MO:93); the gather kernel clamps instead.  The one
        configuration where a caller can plausibly pass one is class_dropout_prob == 0: the table then has NO null
        row, and forward_with_cfg's y = num_classes would silently read class num_classes - 1.  Checked on the host
        there (one device->host read; never inside a CUDA-graph capture)."""
        ye = self.y_embedder
        if ye.dropout_prob > 0 or torch.cuda.is_current_stream_capturing():
            return
        rows = ye.embedding_table.weight.shape[0]
        if y.numel() and (int(y.max()) >= rows or int(y.min()) < 0):
            raise IndexError(f"label out of range for an embedding table of {rows} rows (class_dropout_prob == 0: "
                             "there is no null-class row for classifier-free guidance)")

    def conditioning(self, t, y, force_drop_ids=None):
        """c = t_embedder(t) + y_embedder(y) (MO:241-243), f32 [N, D]."""
        te = self.t_embedder
        t_freq = ops.timestep_embedding(t, te.frequency_embedding_size)
        h = ops.small_linear(t_freq, te.mlp[0].weight, te.mlp[0].bias, silu_out=True)
        ye = self.y_embedder
        if (self.training and ye.dropout_prob > 0) or force_drop_ids is not None:
            y = ye.token_drop(y, force_drop_ids)
        self._check_labels(y)
        y_emb = ops.label_embed(y, ye.embedding_table.weight)
        return ops.small_linear(h, te.mlp[2].weight, te.mlp[2].bias, add=y_emb)

    @torch.no_grad()
    def _forward_inference(self, x, t, y):
        sh = self._shadows()
        bf16 = self.precision == "bf16"
        act = torch.bfloat16 if bf16 else torch.float32
        D, Hh, p = self.hidden_size, self.num_heads, self.patch_size
        hd = D // Hh
        x = x.float().contiguous()
        N = x.shape[0]
        T = (x.shape[2] // p) * (x.shape[3] // p)
        pos = self.pos_embed
        if pos.shape[1] != T:
            raise L.Ditb200Error(f"input grid gives {T} tokens but pos_embed has {pos.shape[1]}")
        tok = ops.patch_embed(x, self.x_embedder.proj.weight, self.x_embedder.proj.bias, pos, p)
        c = self.conditioning(t, y)
        # every block's adaLN (and the final layer's) in one GEMM: c does not depend on x
        if bf16:
            mod = ops.gemm(ops.silu_cast(c, torch.bfloat16), sh["ada_w"], sh["ada_b"], out_dtype=torch.float32)
        else:
            mod = ops.small_linear(c, sh["ada_w"], sh["ada_b"], silu_in=True)
        w = sh["w"]
        # Where the gated residual update `x + gate * branch` (MO:120-121) runs.  0: in the epilogue of the GEMM that
        # produces the branch (f32 stream read + written there).  1 / 2: the GEMM stores the bf16 branch by TMA and
        # the update rides in front of the next LayerNorm (ln_modulate_resid), for proj + fc2 / for proj only
        # (proj's k loop is too short to hide the f32 epilogue).  Chosen per build by measurement.
        branch = _BRANCH_MODE if (bf16 and D in (384, 768, 1024, 1152)) else 0  # widths ln_modulate_resid serves
        pend = None  # (branch output, gate) not yet folded into tok
        # Traversal direction: every kernel of the chain walks the token rows the opposite way to its producer, so it
        # starts on the rows that were written last and are still in the 126 MB L2 (results do not depend on it).
        zig = _ZIGZAG and bf16
        rv = [False]

        def nxt():
            rv[0] = zig and not rv[0]
            return rv[0]

        for i, blk in enumerate(self.blocks):
            m = mod[:, i * 6 * D:(i + 1) * 6 * D]
            sh1, sc1, g1, sh2, sc2, g2 = (m[:, j * D:(j + 1) * D] for j in range(6))
            if pend is None:
                h = ops.ln_modulate(tok, sh1, sc1, T, out_dtype=act, reverse=nxt())
            else:
                _, h = ops.ln_modulate_resid(tok, pend[0], pend[1], sh1, sc1, T, out_dtype=act, x_out=tok, reverse=nxt())
                pend = None
            qkv = ops.gemm(h, w[4 * i], blk.attn.qkv.bias, reverse_m=nxt())
            o = ops.attention(qkv, N, T, Hh, hd, reverse=nxt())
            if branch:
                yb = ops.gemm(o, w[4 * i + 1], blk.attn.proj.bias, reverse_m=nxt())
                _, h = ops.ln_modulate_resid(tok, yb, g1, sh2, sc2, T, out_dtype=act, x_out=tok, reverse=nxt())
            else:
                ops.gemm(o, w[4 * i + 1], blk.attn.proj.bias, epilogue=L.EPI_BIAS_GATE_RESID, resid=tok, gate=g1,
                         rows_per_gate=T, reverse_m=nxt())
                h = ops.ln_modulate(tok, sh2, sc2, T, out_dtype=act, reverse=nxt())
            u = ops.gemm(h, w[4 * i + 2], blk.mlp.fc1.bias, epilogue=L.EPI_BIAS_GELU, reverse_m=nxt())
            if branch == 1:
                pend = (ops.gemm(u, w[4 * i + 3], blk.mlp.fc2.bias, reverse_m=nxt()), g2)
            else:
                ops.gemm(u, w[4 * i + 3], blk.mlp.fc2.bias, epilogue=L.EPI_BIAS_GATE_RESID, resid=tok, gate=g2,
                         rows_per_gate=T, reverse_m=nxt())
        if pend is not None:  # last fc2 branch: plain update, the final layer normalises by itself
            ops.ln_modulate_resid(tok, pend[0], pend[1], None, None, T, x_out=tok, want_out=False)
        mf = mod[:, self.depth * 6 * D:]
        fl = self.final_layer
        return ops.final_layer(tok, mf[:, :D], mf[:, D:], fl.linear.weight, fl.linear.bias, T, p, self.out_channels)

    def forward_raw_cfg(self, x, t, y):
        """The model pass of forward_with_cfg without the guidance combine: returns the raw
        two-half output so the diffusion step kernel can fuse the combine (MO:255-257)."""
        n = x.shape[0] // 2
        half = x[:n]
        return self.forward(torch.cat([half, half], dim=0), t, y)

    def forward_with_cfg(self, x, t, y, cfg_scale):
        """Classifier-free guidance forward (MO:250-266): guidance on the first three eps
        channels only, exactly as the reference does 'for exact reproducibility'."""
        raw = self.forward_raw_cfg(x, t, y)
        return ops.cfg_combine(raw.contiguous(), 3, float(cfg_scale))


def _factory(depth, hidden_size, patch_size, num_heads):
    def make(**kwargs):
        return DiT(depth=depth, hidden_size=hidden_size, patch_size=patch_size, num_heads=num_heads, **kwargs)
    return make


_SIZES = {"XL": (28, 1152, 16), "L": (24, 1024, 16), "B": (12, 768, 12), "S": (12, 384, 6)}

# registry with the reference's twelve keys (MO:365-370)
DiT_models = {f"DiT-{n}/{p}": _factory(d, h, p, nh) for n, (d, h, nh) in _SIZES.items() for p in (2, 4, 8)}

# module-level factory names the reference exports (MO:328-362)
for _n, (_d, _h, _nh) in _SIZES.items():
    for _p in (2, 4, 8):
        globals()[f"DiT_{_n}_{_p}"] = DiT_models[f"DiT-{_n}/{_p}"]
