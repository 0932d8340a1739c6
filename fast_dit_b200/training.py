"""Training step of the DiT denoiser on libditb200: forward that keeps what backward needs, and a
hand-sequenced backward whose every kernel is a C-ABI entry point.

The reference gets this path from torch.autograd over ~600 ATen/cuBLAS kernels per forward
(train_options/train_original.py:206-209 -> models_original.py:233-248).  Here one
torch.autograd.Function spans the whole model: PyTorch only links it into the graph between the
diffusion loss (autograd.py) and the parameters; the arithmetic is

  forward   the inference kernel sequence, with the GEMM epilogues also writing the pre-GELU
            activation and the un-gated branch outputs (aux_out), LayerNorm statistics and the
            attention log-sum-exp;
  backward  per block, in reverse: gated-residual backward -> weight gradient (tokens contracted on
            the tensor cores straight from the row-major activations, MN-major operands) -> data
            gradient (with GELU' fused into the epilogue) -> LayerNorm+modulate backward -> flash
            attention backward -> adaLN backward; bias gradients by column sums.

Gradients are written into one flat f32 arena laid out in `parameters()` order, so a data-parallel
wrapper (parallel.DataParallel) can all-reduce each block's contiguous slice as soon as that block's
backward is done, overlapping NCCL with the rest of backward (train_original.py:149's DDP).
"""
from __future__ import annotations

import math

import torch

from . import _lib as L
from . import ops


# ------------------------------------------------------------------ arenas
class ArenaLayout:
    """Where every trainable parameter lives in a flat array.  One layout is shared by the gradient arena,
    and — when optim.FusedAdamWEMA owns the parameters — by the f32 master weights, the Adam moments, the
    EMA weights and the bf16 weight shadows, so the optimizer is one pass over [0, total).

    Order: embedders | per block {qkv, proj, fc1, fc2} weights+biases | final linear | every
    adaLN_modulation weight (blocks 0..L-1, final layer) | every adaLN bias.  The adaLN weights are
    contiguous and unpadded on purpose: that region IS the [(6L+2)·D, D] matrix of the batched adaLN GEMM.
    Buckets (what a data-parallel step all-reduces as soon as it is complete): 'final_layer',
    'blocks.i' for i = L-1..0, 'embed' (which also carries the small adaLN-bias region)."""

    ALIGN = 64  # floats: gradients start on 256-byte boundaries (vector atomics, 128-bit stores)
    SHARD_ALIGN = 512  # floats: a region an optimizer shard is cut from divides evenly among <= 8 ranks, 64-float parts

    def __init__(self, model):
        ada_w, ada_b = model._ada_params()
        ada_ids = {id(p) for p in ada_w + ada_b}
        names = {id(p): n for n, p in model.named_parameters()}
        groups = {"embed": [], "final_layer": []}
        for i in range(model.depth):
            groups[f"blocks.{i}"] = []
        for p in model.parameters():
            if not p.requires_grad or id(p) in ada_ids:
                continue
            n = names[id(p)]
            head = n.split(".")[0]
            key = "embed" if head in ("x_embedder", "t_embedder", "y_embedder") else \
                  ".".join(n.split(".")[:2]) if head == "blocks" else head
            groups[key].append(p)
        self.offsets, self.ranges = {}, {}
        off = 0

        def place(plist, pad, align=None):
            nonlocal off
            align = align or self.ALIGN
            lo = off
            for p in plist:
                self.offsets[id(p)] = (off, p.numel(), tuple(p.shape))
                off += p.numel()
                if pad:
                    off = (off + self.ALIGN - 1) // self.ALIGN * self.ALIGN
            off = (off + align - 1) // align * align
            return lo, off

        # Inside a block the four GEMM weights come first and the biases after them: the weight part (99.9 % of the
        # block) is what a sharded optimizer reduce-scatters and updates 1/W of per rank (parallel.DataParallel
        # (shard_optimizer=True)); the small tail is all-reduced and updated everywhere.
        self.big, self.small = {}, {}
        for key in ["embed"] + [f"blocks.{i}" for i in range(model.depth)] + ["final_layer"]:
            if key.startswith("blocks."):
                off = (off + self.SHARD_ALIGN - 1) // self.SHARD_ALIGN * self.SHARD_ALIGN
                lo, mid = place([p for p in groups[key] if p.dim() >= 2], True, self.SHARD_ALIGN)
                _, hi = place([p for p in groups[key] if p.dim() < 2], True)
                self.ranges[key], self.big[key], self.small[key] = (lo, hi), [(lo, mid)], [(mid, hi)]
            else:
                self.ranges[key] = place(groups[key], True)
                self.big[key], self.small[key] = [], [self.ranges[key]]
        assert all(p.numel() % 4 == 0 for p in ada_w + ada_b)
        self.ranges["ada_w"] = place([p for p in ada_w if p.requires_grad], False)
        self.ranges["ada_b"] = place([p for p in ada_b if p.requires_grad], False)
        self.total = off
        self.ada_rows = sum(p.shape[0] for p in ada_w)
        self.buckets = {"embed": [self.ranges["embed"], self.ranges["ada_b"]]}
        self.small["embed"] = [self.ranges["embed"], self.ranges["ada_b"]]
        for i, p in enumerate(ada_w):
            key = f"blocks.{i}" if i < model.depth else "final_layer"
            o, n, _ = self.offsets[id(p)]
            self.buckets[key] = [self.ranges[key], (o, o + n)]
            if n % self.SHARD_ALIGN == 0:
                self.big[key] = self.big[key] + [(o, o + n)]
            else:  # odd hidden sizes: the adaLN slice stays replicated
                self.small[key] = self.small[key] + [(o, o + n)]

    def view(self, flat, p):
        off, n, shape = self.offsets[id(p)]
        return flat[off:off + n].view(shape)


def layout_for(model) -> ArenaLayout:
    lay = getattr(model, "_layout", None)
    if lay is None:
        lay = ArenaLayout(model)
        model._layout = lay
    return lay


class GradArena:
    """One flat f32 buffer holding every trainable parameter's gradient (layout: ArenaLayout)."""

    def __init__(self, model):
        self.layout = layout_for(model)
        self.flat = torch.zeros(self.layout.total, device=next(model.parameters()).device, dtype=torch.float32)

    def view(self, p):
        return self.layout.view(self.flat, p)

    def bucket(self, key):
        """The contiguous slices that make up one data-parallel bucket."""
        return [self.flat[lo:hi] for lo, hi in self.layout.buckets[key]]

    def aliases(self, p):
        g = p.grad
        return g is not None and g.data_ptr() == self.flat.data_ptr() + 4 * self.layout.offsets[id(p)][0]


def _arena_for(model):
    """Reuse the model's arena unless a parameter's .grad still lives in it (the caller is accumulating
    gradients across backward calls): then this backward gets a scratch arena and autograd adds."""
    dev = next(model.parameters()).device
    ar = getattr(model, "_grad_arena", None)
    if ar is None or ar.flat.device != dev:
        ar = GradArena(model)
        model._grad_arena = ar
        return ar
    if any(ar.aliases(p) for p in model.parameters() if p.requires_grad):
        if getattr(model, "_grad_sync", None) is not None:
            raise L.Ditb200Error("DataParallel needs zero_grad(set_to_none=True) between steps: gradient "
                                 "accumulation across backward calls is not supported under data parallelism")
        return GradArena(model)
    return ar


def _split_k(M, N, K, trans_w):
    """Work units per tile for a long-K GEMM: fill the machine when there are few output tiles (weight
    gradients contract over all tokens; the adaLN data gradient has M = batch)."""
    cg = 2 if M > 128 else 1
    bn = 256 if N > 192 else (256 if (N > 128 and trans_w) else 192 if N > 128 else 128)
    tiles = math.ceil(M / (128 * cg)) * math.ceil(N / bn)
    ctas = 148 // cg
    kb = math.ceil(K / 64)
    best, best_cost = 1, None
    # measured on B200 (tools/tc_probe.py sweep, K = 8192): even when one wave holds every tile, two k-splits
    # per tile run 10-15 % faster than one (3456 x 1152: 60.7 vs 70.8 us)
    s_min = 2 if kb >= 64 else 1
    for s in range(s_min, min(32, kb) + 1):
        waves = math.ceil(tiles * s / ctas)
        cost = waves / s * (1.0 + 0.03 * (s - 1))  # atomics + shorter main loops are not free
        if best_cost is None or cost < best_cost - 1e-9:
            best, best_cost = s, cost
    return best


def _wgrad(dy, x, out):
    """out[Nout, Kin] = dy[tokens, Nout]^T @ x[tokens, Kin] on the tensor cores (no transposed copies)."""
    Mo, No, K = dy.shape[1], x.shape[1], dy.shape[0]
    return ops.gemm(dy, x, None, out=out, trans_a=True, trans_w=True, split_k=_split_k(Mo, No, K, True))


def _dgrad(dy, w, **kw):
    """dx[tokens, Kin] = dy[tokens, Nout] @ w[Nout, Kin]."""
    return ops.gemm(dy, w, None, trans_w=True, **kw)


class _DiTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, model, x, t, y, *params):
        if model.precision != "bf16":
            raise L.Ditb200Error("training runs in bf16 (tcgen05) only; the fp32 check mode is forward-only")
        sh = model._shadows()
        D, Hh, p, Ld = model.hidden_size, model.num_heads, model.patch_size, model.depth
        hd = D // Hh
        x = x.float().contiguous()
        N = x.shape[0]
        T = (x.shape[2] // p) * (x.shape[3] // p)
        M = N * T
        dev = x.device
        if model.pos_embed.shape[1] != T:
            raise L.Ditb200Error(f"input grid gives {T} tokens but pos_embed has {model.pos_embed.shape[1]}")
        tok = ops.patch_embed(x, model.x_embedder.proj.weight, model.x_embedder.proj.bias, model.pos_embed, p)
        # conditioning (models_original.py:241-243), keeping the SiLU pre-activations
        te, ye = model.t_embedder, model.y_embedder
        t_freq = ops.timestep_embedding(t, te.frequency_embedding_size)
        pre1 = ops.small_linear(t_freq, te.mlp[0].weight, te.mlp[0].bias)
        h1 = ops.silu_cast(pre1, torch.float32)
        y_idx = ye.token_drop(y) if (model.training and ye.dropout_prob > 0) else y
        model._check_labels(y_idx)
        y_emb = ops.label_embed(y_idx, ye.embedding_table.weight)
        c = ops.small_linear(h1, te.mlp[2].weight, te.mlp[2].bias, add=y_emb)
        sc = ops.silu_cast(c, torch.bfloat16)
        mod = ops.gemm(sc, sh["ada_w"], sh["ada_b"], out_dtype=torch.float32)
        w = sh["w"]
        saved = []
        bf = torch.bfloat16
        # The branch outputs y1 / y2 are kept (bf16) for backward anyway, so the proj / fc2 GEMMs write only them
        # (plain bias epilogue through TMA) and the gated residual update x + gate * y rides in front of the NEXT
        # LayerNorm+modulate kernel, which reads the stream regardless (ops.ln_modulate_resid).
        fused = D in (384, 768, 1024, 1152)
        pend = None  # (y2, g2) of the previous block, not yet added to the stream `tok`
        for i, blk in enumerate(model.blocks):
            m = mod[:, i * 6 * D:(i + 1) * 6 * D]
            sh1, sc1, g1, sh2, sc2, g2 = (m[:, j * D:(j + 1) * D] for j in range(6))
            st1 = torch.empty((M, 2), device=dev, dtype=torch.float32)
            if pend is None:
                hA = ops.ln_modulate(tok, sh1, sc1, T, out_dtype=bf, stats=st1)
            else:
                tok, hA = ops.ln_modulate_resid(tok, pend[0], pend[1], sh1, sc1, T, out_dtype=bf, stats=st1)
            qkv = ops.gemm(hA, w[4 * i], blk.attn.qkv.bias)
            lse = torch.empty((N, Hh, T), device=dev, dtype=torch.float32)
            o = ops.attention(qkv, N, T, Hh, hd, lse=lse)
            st2 = torch.empty((M, 2), device=dev, dtype=torch.float32)
            if fused:
                y1 = ops.gemm(o, w[4 * i + 1], blk.attn.proj.bias)
                x_mid, hB = ops.ln_modulate_resid(tok, y1, g1, sh2, sc2, T, out_dtype=bf, stats=st2)
            else:
                y1 = torch.empty((M, D), device=dev, dtype=bf)
                x_mid = torch.empty_like(tok)
                ops.gemm(o, w[4 * i + 1], blk.attn.proj.bias, epilogue=L.EPI_BIAS_GATE_RESID, resid=tok, gate=g1,
                         rows_per_gate=T, out=x_mid, aux_out=y1)
                hB = ops.ln_modulate(x_mid, sh2, sc2, T, out_dtype=bf, stats=st2)
            a1 = torch.empty((M, w[4 * i + 2].shape[0]), device=dev, dtype=bf)
            # a1 receives gelu'(fc1 pre-activation): the backward epilogue multiplies instead of re-evaluating tanh
            u = ops.gemm(hB, w[4 * i + 2], blk.mlp.fc1.bias, epilogue=L.EPI_BIAS_GELU_DAUX, aux_out=a1)
            if fused:
                y2 = ops.gemm(u, w[4 * i + 3], blk.mlp.fc2.bias)
                saved.append((tok, st1, hA, qkv, lse, o, y1, x_mid, st2, hB, a1, u, y2))
                tok, pend = x_mid, (y2, g2)
            else:
                y2 = torch.empty((M, D), device=dev, dtype=bf)
                x_out = torch.empty_like(tok)
                ops.gemm(u, w[4 * i + 3], blk.mlp.fc2.bias, epilogue=L.EPI_BIAS_GATE_RESID, resid=x_mid, gate=g2,
                         rows_per_gate=T, out=x_out, aux_out=y2)
                saved.append((tok, st1, hA, qkv, lse, o, y1, x_mid, st2, hB, a1, u, y2))
                tok = x_out
        if pend is not None:  # the last block's MLP branch: only the stream update (the final layer normalises itself)
            tok, _ = ops.ln_modulate_resid(tok, pend[0], pend[1], pend[1], pend[1], T, want_out=False)
        mf = mod[:, Ld * 6 * D:]
        fl = model.final_layer
        out = ops.final_layer(tok, mf[:, :D], mf[:, D:], fl.linear.weight, fl.linear.bias, T, p, model.out_channels)
        ctx.model = model
        ctx.saved = saved
        ctx.misc = (x, t_freq, pre1, h1, y_idx, c, sc, mod, tok, N, T)
        ctx.n_params = len(params)
        return out

    @staticmethod
    def backward(ctx, dout):
        model = ctx.model
        sh = model._shadows()
        x, t_freq, pre1, h1, y_idx, c, sc, mod, tok_final, N, T = ctx.misc
        D, Hh, p, Ld = model.hidden_size, model.num_heads, model.patch_size, model.depth
        hd = D // Hh
        dev = x.device
        bf = torch.bfloat16
        arena = _arena_for(model)
        G = arena.view
        def sync(key, arena):
            """A bucket's gradients are final: hand it to the data-parallel wrapper (all-reduce issued at once,
            parallel.DataParallel) and to an optimizer that updates during backward (optim.FusedAdamWEMA with
            overlap_backward=True; it runs after the bucket's collective)."""
            reduce = getattr(model, "_grad_sync", None)
            after = reduce(key, arena) if reduce is not None else None
            plan = None
            if isinstance(after, tuple):  # sharded optimizer: (what to wait for, which parts this rank updates)
                after, plan = after
            ready = getattr(model, "_bucket_ready", None)
            if ready is not None:
                ready(key, arena, after, plan)

        w = sh["w"]
        ada_w = sh["ada_w"]
        M = N * T
        # Gradient of every adaLN output, one row per image: [blocks 0..L-1: 6 chunks each | final layer: 2 chunks].
        # The LayerNorm / gate backward kernels reduce into its slices; the weight gradients are taken block by block
        # (so every bucket is complete when its block is), the data gradient d silu(c) once, over all of it, at the
        # end — unless an optimizer updates each bucket during backward (FusedAdamWEMA(overlap_backward=True), the
        # sharded optimizer): the adaLN weights of a finished block are then no longer the ones the forward used, so
        # every block contributes its share of d silu(c) before it hands its bucket over.
        early_update = getattr(model, "_bucket_ready", None) is not None
        widths = [6 * D] * Ld + [2 * D]
        if early_update:
            dmods = [torch.zeros((N, wd), device=dev, dtype=torch.float32) for wd in widths]
            dsc = torch.zeros((N, D), device=dev, dtype=torch.float32)
        else:
            dmod_all = torch.zeros((N, sum(widths)), device=dev, dtype=torch.float32)
            dmods = [dmod_all[:, 6 * D * i:6 * D * i + wd] for i, wd in enumerate(widths)]
        simt_ada = ops.adaln_wgrad_preferred(N, 2 * D, D)  # 6 * D and 2 * D divide alike

        def ada_bwd(dmod, w_rows, lin):
            """adaLN_modulation[1] backward: weight and bias gradients of one Linear (and, in early-update mode, its
            contribution to d silu(c))."""
            dmod_bf = None
            if simt_ada:
                ops.adaln_wgrad(dmod, sc, G(lin.weight), G(lin.bias))
            else:
                dmod = dmod.contiguous()
                dmod_bf = ops.cast_bf16(dmod)
                ops.gemm(dmod_bf, sc, None, out=G(lin.weight), trans_a=True, trans_w=True)
                ops.colsum(dmod, out=G(lin.bias))
            if early_update:
                ops.gemm(dmod_bf if dmod_bf is not None else ops.cast_bf16(dmod), w_rows, None, out=dsc, trans_w=True,
                         accumulate=True, split_k=_split_k(N, D, dmod.shape[1], True))

        # Every LayerNorm backward completes the gradient of the residual stream at its point of the chain, and the
        # next consumer of that gradient is the gated-residual backward of the branch that joined the stream just
        # before that LayerNorm (the previous block's MLP branch, or this block's attention branch): one fused pass
        # where the shapes allow it (ops.ln_modulate_bwd_gate), the two kernels otherwise.
        fuse_gate = ops.ln_modulate_bwd_gate_ok(T, D)
        dtok = torch.empty((M, D), device=dev, dtype=torch.float32)

        def ln_bwd(dh, x_ln, scale, stats, accumulate, dshift, dscale, branch):
            """branch = (y, gate, dgate, bias parameter) of the gated residual whose backward comes next, or None.
            Returns that branch's dy."""
            if branch is None:
                ops.ln_modulate_bwd(dh, x_ln, scale, stats, T, dtok, accumulate, dshift, dscale)
                return None
            yb, gate, dgate, bias = branch
            gb = G(bias).zero_()
            if fuse_gate:
                return ops.ln_modulate_bwd_gate(dh, x_ln, scale, stats, T, dtok, accumulate, dshift, dscale,
                                                y=yb, gate=gate, dgate=dgate, dbias=gb)[1]
            ops.ln_modulate_bwd(dh, x_ln, scale, stats, T, dtok, accumulate, dshift, dscale)
            return ops.gate_resid_bwd(dtok, yb, gate, T, dgate, dbias=gb)

        def mlp_branch(i):
            """The MLP branch of block i as the gate backward needs it."""
            m = mod[:, i * 6 * D:(i + 1) * 6 * D]
            return (ctx.saved[i][12], m[:, 5 * D:6 * D], dmods[i][:, 5 * D:6 * D], model.blocks[i].mlp.fc2.bias)

        # ---------------- final layer (models_original.py:138-142, 218-231)
        fl = model.final_layer
        mf = mod[:, Ld * 6 * D:]
        dz = ops.unpatchify_bwd(dout.float().contiguous(), p)
        stf = torch.empty((M, 2), device=dev, dtype=torch.float32)
        hf = ops.ln_modulate(tok_final, mf[:, :D], mf[:, D:], T, out_dtype=bf, stats=stf)
        _wgrad(dz, hf, G(fl.linear.weight))
        ops.colsum(dz, out=G(fl.linear.bias))
        dhf = _dgrad(dz, ops.cast_bf16(fl.linear.weight.detach()))
        dmod_f = dmods[Ld]
        dy2 = ln_bwd(dhf, tok_final, mf[:, D:], stf, False, dmod_f[:, :D], dmod_f[:, D:], mlp_branch(Ld - 1))
        ada_bwd(dmod_f, ada_w[Ld * 6 * D:], fl.adaLN_modulation[1])
        del dz, hf, dhf
        sync("final_layer", arena)

        # ---------------- blocks, last to first (models_original.py:118-122)
        for i in range(Ld - 1, -1, -1):
            blk = model.blocks[i]
            x_in, st1, hA, qkv, lse, o, y1, x_mid, st2, hB, a1, u, y2 = ctx.saved[i]
            m = mod[:, i * 6 * D:(i + 1) * 6 * D]
            sc1, g1, sc2 = m[:, D:2 * D], m[:, 2 * D:3 * D], m[:, 4 * D:5 * D]
            dmod = dmods[i]
            dm = [dmod[:, j * D:(j + 1) * D] for j in range(6)]
            # x_out = x_mid + g2 * fc2(gelu(fc1(modulate(LN(x_mid))))); dy2 = g2 * d x_out came with the previous kernel
            _wgrad(dy2, u, G(blk.mlp.fc2.weight))
            da1 = _dgrad(dy2, w[4 * i + 3], epilogue=L.EPI_MUL_AUX, aux_in=a1)
            _wgrad(da1, hB, G(blk.mlp.fc1.weight))
            ops.colsum(da1, out=G(blk.mlp.fc1.bias))
            dh = _dgrad(da1, w[4 * i + 2])
            dy1 = ln_bwd(dh, x_mid, sc2, st2, True, dm[3], dm[4], (y1, g1, dm[2], blk.attn.proj.bias))
            del dy2, da1, dh, y2, u, a1, hB
            # x_mid = x_in + g1 * proj(attention(qkv(modulate(LN(x_in)))))
            _wgrad(dy1, o, G(blk.attn.proj.weight))
            do = _dgrad(dy1, w[4 * i + 1])
            dqkv = ops.attention_bwd(qkv, o, do, lse, N, T, Hh, hd)
            _wgrad(dqkv, hA, G(blk.attn.qkv.weight))
            ops.colsum(dqkv, out=G(blk.attn.qkv.bias))
            dh = _dgrad(dqkv, w[4 * i])
            dy2 = ln_bwd(dh, x_in, sc1, st1, True, dm[0], dm[1], mlp_branch(i - 1) if i > 0 else None)
            ctx.saved[i] = None
            ada_bwd(dmod, ada_w[i * 6 * D:(i + 1) * 6 * D], blk.adaLN_modulation[1])
            del dy1, do, dqkv, dh
            sync(f"blocks.{i}", arena)

        if not early_update:  # d silu(c), summed over every adaLN layer: one long-K GEMM over the whole modulation gradient
            dsc = ops.gemm(ops.cast_bf16(dmod_all), ada_w, None, out_dtype=torch.float32, trans_w=True,
                           split_k=_split_k(N, D, dmod_all.shape[1], True))

        # ---------------- embedders (models_original.py:240-243)
        pe = model.x_embedder.proj
        _wgrad(ops.cast_bf16(dtok), ops.patchify(x, p), G(pe.weight).view(D, -1))
        ops.colsum(dtok, out=G(pe.bias))
        dc = ops.silu_bwd(dsc, c)
        te, ye = model.t_embedder, model.y_embedder
        ops.label_embed_bwd(dc, y_idx, G(ye.embedding_table.weight).zero_())
        dc_bf = ops.cast_bf16(dc)
        ops.gemm(dc_bf, ops.cast_bf16(h1), None, out=G(te.mlp[2].weight), trans_a=True, trans_w=True)
        ops.colsum(dc, out=G(te.mlp[2].bias))
        dh1 = _dgrad(dc_bf, ops.cast_bf16(te.mlp[2].weight.detach()), out_dtype=torch.float32)
        dpre = ops.silu_bwd(dh1, pre1)
        ops.gemm(ops.cast_bf16(dpre), ops.cast_bf16(t_freq), None, out=G(te.mlp[0].weight), trans_a=True, trans_w=True)
        ops.colsum(dpre, out=G(te.mlp[0].bias))
        sync("embed", arena)
        sync(None, arena)  # all buckets issued: make the compute stream wait for the collectives / the early updates

        grads = [arena.view(q) if q.requires_grad else None for q in model.parameters()]
        assert len(grads) == ctx.n_params
        return (None, None, None, None, *grads)


def dit_forward_train(model, x, t, y):
    """DiT.forward with autograd: returns the model output wired to every parameter's gradient."""
    return _DiTFunction.apply(model, x, t, y, *model.parameters())
