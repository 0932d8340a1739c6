"""Backward pass on the GPU: each backward kernel against torch autograd of the same fp32 op, then the
whole training step (model gradients, diffusion loss) against the CPU oracle's autograd."""
import math

import pytest
import torch
import torch.nn.functional as F

from util import rel_l2

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from fast_dit_b200 import ops as o

    return o


def _g(seed=0):
    return torch.Generator(device="cuda").manual_seed(seed)


# ------------------------------------------------------------------------------- GEMM variants
@pytest.mark.parametrize("ta,tw", [(False, True), (True, True), (True, False)])
@pytest.mark.parametrize("M,N,K", [(512, 384, 256), (1152, 1152, 2048), (32, 1152, 1024), (384, 16, 1024),
                                   (2304, 384, 32), (300 * 8, 520, 200)])
def test_gemm_transposed_operands(ops, dev, ta, tw, M, N, K):
    if not (ta and tw) and K % 8:
        pytest.skip("K-major operands need K % 8 == 0")
    g = _g(1)
    A = torch.randn(M, K, device=dev, generator=g)
    W = torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)
    a = (A.t().contiguous() if ta else A).bfloat16()
    w = (W.t().contiguous() if tw else W).bfloat16()
    ref = (a.double().t() if ta else a.double()) @ (w.double() if tw else w.double().t())
    y = ops.gemm(a, w, None, out_dtype=torch.float32, trans_a=ta, trans_w=tw)
    assert rel_l2(y, ref) < 2e-5
    # split-K and accumulate on top of an existing value
    base = torch.randn(M, N, device=dev, generator=g)
    y2 = ops.gemm(a, w, None, out=base.clone(), trans_a=ta, trans_w=tw, split_k=3, accumulate=True)
    assert rel_l2(y2, ref + base.double()) < 2e-5
    y3 = ops.gemm(a, w, None, out_dtype=torch.float32, trans_a=ta, trans_w=tw, split_k=4)
    assert rel_l2(y3, ref) < 2e-5


@pytest.mark.parametrize("cg,bn", [(1, 128), (1, 256), (2, 128), (2, 256), (1, 192)])
def test_gemm_transposed_tiles(ops, dev, cg, bn):
    g = _g(2)
    M, N, K = 768, 640, 512
    a = torch.randn(K, M, device=dev, generator=g).bfloat16()
    w = (torch.randn(K, N, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    ref = a.double().t() @ w.double()
    y = ops.gemm(a, w, None, out_dtype=torch.float32, trans_a=True, trans_w=True, tile_n=bn, cta_group=cg)
    assert rel_l2(y, ref) < 2e-5


def test_gemm_aux_out_and_dgelu(ops, dev):
    from fast_dit_b200 import _lib as L

    g = _g(3)
    M, N, K = 1024, 1536, 384
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, device=dev, generator=g)
    pre_ref = a.double() @ w.double().t() + b.double()
    aux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    u = ops.gemm(a, w, b, epilogue=L.EPI_BIAS_GELU, aux_out=aux)
    assert rel_l2(aux.float(), pre_ref) < 4e-3
    assert rel_l2(u.float(), F.gelu(pre_ref, approximate="tanh")) < 4e-3
    # gated residual keeps the un-gated branch
    T = 64
    resid = torch.randn(M, N, device=dev, generator=g)
    gate = torch.randn(M // T, N, device=dev, generator=g)
    aux2 = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    out = torch.empty_like(resid)
    ops.gemm(a, w, b, epilogue=L.EPI_BIAS_GATE_RESID, resid=resid, gate=gate, rows_per_gate=T, out=out, aux_out=aux2)
    assert rel_l2(aux2.float(), pre_ref) < 4e-3
    assert rel_l2(out, resid.double() + gate.double().repeat_interleave(T, 0) * pre_ref) < 1e-5
    # data gradient through GELU: dpre = (dy @ W2) * gelu'(pre)
    dy = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w2 = (torch.randn(K, N, device=dev, generator=g) / math.sqrt(K)).bfloat16()  # fc2.weight [out=K, in=N]
    pre = aux.float().double().requires_grad_(True)
    F.gelu(pre, approximate="tanh").backward(dy.double() @ w2.double())
    got = ops.gemm(dy, w2, None, trans_w=True, epilogue=L.EPI_MUL_DGELU, aux_in=aux)
    assert rel_l2(got.float(), pre.grad) < 5e-3
    # the pair the training step uses: forward leaves gelu'(pre) behind, backward multiplies by it
    daux = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    u2 = ops.gemm(a, w, b, epilogue=L.EPI_BIAS_GELU_DAUX, aux_out=daux)
    assert rel_l2(u2.float(), F.gelu(pre_ref, approximate="tanh")) < 4e-3
    pre_d = pre_ref.clone().requires_grad_(True)
    F.gelu(pre_d, approximate="tanh").sum().backward()
    assert rel_l2(daux.float(), pre_d.grad) < 4e-3
    got2 = ops.gemm(dy, w2, None, trans_w=True, epilogue=L.EPI_MUL_AUX, aux_in=daux)
    assert rel_l2(got2.float(), (dy.double() @ w2.double()) * daux.double()) < 4e-3
    assert rel_l2(got2.float(), (dy.double() @ w2.double()) * pre_d.grad) < 8e-3


# ---------------------------------------------------------------------- elementwise backward
@pytest.mark.parametrize("D,T", [(384, 37), (1152, 256), (768, 64)])
@pytest.mark.parametrize("dh_dtype", [torch.float32, torch.bfloat16])
def test_ln_modulate_bwd(ops, dev, D, T, dh_dtype):
    g = _g(4)
    B = 3
    x = (torch.randn(B * T, D, device=dev, generator=g) * 2 + 0.3)
    mod = torch.randn(B, 6 * D, device=dev, generator=g) * 0.5
    shift, scale = mod[:, :D], mod[:, D:2 * D]
    dh = torch.randn(B * T, D, device=dev, generator=g).to(dh_dtype)
    stats = torch.empty(B * T, 2, device=dev)
    ops.ln_modulate(x, shift, scale, T, out_dtype=torch.float32, stats=stats)
    xd = x.double().requires_grad_(True)
    sd = shift.double().clone().requires_grad_(True)
    cd = scale.double().clone().requires_grad_(True)
    h = F.layer_norm(xd, (D,), eps=1e-6).view(B, T, D) * (1 + cd[:, None]) + sd[:, None]
    h.backward(dh.double().view(B, T, D))
    prev = torch.randn(B * T, D, device=dev, generator=g)
    dmod = torch.zeros(B, 6 * D, device=dev)
    dx = ops.ln_modulate_bwd(dh, x, scale, stats, T, prev.clone(), True, dmod[:, :D], dmod[:, D:2 * D])
    assert rel_l2(dx, xd.grad + prev.double()) < 2e-5
    assert rel_l2(dmod[:, :D], sd.grad) < 2e-5
    assert rel_l2(dmod[:, D:2 * D], cd.grad) < 2e-5
    dx2 = ops.ln_modulate_bwd(dh, x, scale, stats, T, torch.empty_like(x), False, dmod[:, :D], dmod[:, D:2 * D])
    assert rel_l2(dx2, xd.grad) < 2e-5


@pytest.mark.parametrize("D,T,B", [(384, 36, 3), (1152, 256, 5), (768, 64, 40), (1024, 8, 2), (1152, 4, 1)])
@pytest.mark.parametrize("dh_dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("with_gate", [True, False])
def test_ln_modulate_bwd_gate_fused(ops, dev, D, T, B, dh_dtype, with_gate):
    """The fused LayerNorm-backward + next gated-residual-backward kernel against the fp64 formulas, and against the two
    separate kernels it replaces (training.py's backward chain); B * T / 4 row batches over a persistent grid, so
    (768, 64, 40) makes CTAs cross image boundaries and flush their column sums in between."""
    assert ops.ln_modulate_bwd_gate_ok(T, D)
    g = _g(14)
    x = (torch.randn(B * T, D, device=dev, generator=g) * 2 + 0.3)
    mod = torch.randn(B, 6 * D, device=dev, generator=g) * 0.5
    shift, scale, gate = mod[:, :D], mod[:, D:2 * D], mod[:, 2 * D:3 * D]
    dh = torch.randn(B * T, D, device=dev, generator=g).to(dh_dtype)
    y = torch.randn(B * T, D, device=dev, generator=g).bfloat16()
    stats = torch.empty(B * T, 2, device=dev)
    ops.ln_modulate(x, shift, scale, T, out_dtype=torch.float32, stats=stats)
    xd = x.double().requires_grad_(True)
    sd = shift.double().clone().requires_grad_(True)
    cd = scale.double().clone().requires_grad_(True)
    h = F.layer_norm(xd, (D,), eps=1e-6).view(B, T, D) * (1 + cd[:, None]) + sd[:, None]
    h.backward(dh.double().view(B, T, D))
    prev = torch.randn(B * T, D, device=dev, generator=g)
    for accumulate in (True, False):
        want_dx = xd.grad + (prev.double() if accumulate else 0)
        dmod = torch.zeros(B, 6 * D, device=dev)
        dbias = torch.zeros(D, device=dev)
        dx0 = prev.clone() if accumulate else torch.full_like(prev, float("nan"))
        kw = dict(y=y, gate=gate, dgate=dmod[:, 2 * D:3 * D], dbias=dbias) if with_gate else {}
        dx, dy = ops.ln_modulate_bwd_gate(dh, x, scale, stats, T, dx0, accumulate, dmod[:, :D], dmod[:, D:2 * D], **kw)
        assert rel_l2(dx, want_dx) < 2e-5
        assert rel_l2(dmod[:, :D], sd.grad) < 2e-5
        assert rel_l2(dmod[:, D:2 * D], cd.grad) < 2e-5
        if not with_gate:
            assert dy is None and float(dmod[:, 2 * D:].abs().max()) == 0.0
            continue
        want_dy = want_dx.view(B, T, D) * gate.double()[:, None]
        assert rel_l2(dy.float(), want_dy.view(B * T, D)) < 4e-3
        assert rel_l2(dmod[:, 2 * D:3 * D], (want_dx * y.double()).view(B, T, D).sum(1)) < 2e-5
        assert rel_l2(dbias, want_dy.sum((0, 1))) < 2e-5
        assert float(dmod[:, 3 * D:].abs().max()) == 0.0
        # the two kernels it replaces, on the same inputs: same dx to rounding, the same bf16 dy almost everywhere
        dmod2 = torch.zeros(B, 6 * D, device=dev)
        dbias2 = torch.zeros(D, device=dev)
        dx2 = ops.ln_modulate_bwd(dh, x, scale, stats, T, prev.clone() if accumulate else torch.empty_like(prev),
                                  accumulate, dmod2[:, :D], dmod2[:, D:2 * D])
        dy2 = ops.gate_resid_bwd(dx2, y, gate, T, dmod2[:, 2 * D:3 * D], dbias=dbias2)
        assert rel_l2(dx, dx2) < 1e-6 and rel_l2(dy.float(), dy2.float()) < 1e-3
        assert rel_l2(dmod, dmod2) < 1e-5 and rel_l2(dbias, dbias2) < 1e-5


@pytest.mark.parametrize("N,D,chunks", [(32, 1152, 6), (5, 384, 2), (70, 768, 6), (256, 128, 1)])
def test_adaln_wgrad(ops, dev, N, D, chunks):
    """Outer-product weight / bias gradient of an adaLN Linear, reading a column slice of a wider f32 buffer."""
    g = _g(15)
    R = chunks * D
    assert ops.adaln_wgrad_ok(N, R, D)
    wide = torch.randn(N, R + 2 * D, device=dev, generator=g)
    dmod = wide[:, D:D + R]
    sc = torch.randn(N, D, device=dev, generator=g).bfloat16()
    dw = torch.full((R, D), float("nan"), device=dev)
    db = torch.full((R,), float("nan"), device=dev)
    ops.adaln_wgrad(dmod, sc, dw, db)
    assert rel_l2(dw, dmod.double().t() @ sc.double()) < 1e-6
    assert rel_l2(db, dmod.double().sum(0)) < 1e-6
    dw2 = torch.full((R, D), float("nan"), device=dev)
    ops.adaln_wgrad(dmod, sc, dw2)
    assert torch.equal(dw, dw2)


@pytest.mark.parametrize("N,D,chunks", [(256, 768, 6), (96, 1152, 6), (256, 384, 2)])
def test_adaln_weight_gradient_gemm_route(ops, dev, N, D, chunks):
    """Above ~64 images training.py takes the adaLN weight gradient through cast + tcgen05 GEMM over the batch (both
    operands MN-major) + column sum instead of the outer-product kernel (ops.adaln_wgrad_preferred): same
    statement as test_adaln_wgrad, through that route, and the two routes against each other."""
    g = _g(16)
    R = chunks * D
    assert ops.adaln_wgrad_ok(N, R, D) and not ops.adaln_wgrad_preferred(N, R, D)
    assert ops.adaln_wgrad_preferred(32, R, D)
    dmod = torch.randn(N, R, device=dev, generator=g)
    sc = torch.randn(N, D, device=dev, generator=g).bfloat16()
    dw = torch.full((R, D), float("nan"), device=dev)
    db = torch.full((R,), float("nan"), device=dev)
    dmod_bf = ops.cast_bf16(dmod)
    ops.gemm(dmod_bf, sc, None, out=dw, trans_a=True, trans_w=True)
    ops.colsum(dmod, out=db)
    assert rel_l2(dw, dmod_bf.double().t() @ sc.double()) < 1e-5   # exact products of the bf16 operands, f32 sums
    assert rel_l2(dw, dmod.double().t() @ sc.double()) < 4e-3      # the cast of dmod is the route's only rounding
    assert rel_l2(db, dmod.double().sum(0)) < 1e-6
    dw2 = torch.full((R, D), float("nan"), device=dev)
    ops.adaln_wgrad(dmod, sc, dw2)
    assert rel_l2(dw, dw2) < 4e-3


def test_model_gradients_with_adaln_gemm_route(dev, monkeypatch):
    """The whole backward with the adaLN weight gradients forced through the GEMM route (what batches above 64
    images take) against the outer-product route on the same inputs."""
    from util import build_product_model

    from fast_dit_b200 import ops as _ops

    m = build_product_model("DiT-S/2", input_size=32, num_classes=1000, precision="bf16").cuda()
    g = torch.Generator().manual_seed(12)
    B = 8
    x = torch.randn(B, 4, 32, 32, generator=g).cuda()
    t = torch.randint(0, 1000, (B,), generator=g).cuda()
    y = torch.randint(0, 1001, (B,), generator=g).cuda()
    dout = torch.randn(B, 8, 32, 32, generator=g).cuda()
    m(x, t, y).backward(dout)
    simt = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    monkeypatch.setattr(_ops, "adaln_wgrad_preferred", lambda N, R, D: False)
    m(x, t, y).backward(dout)
    ada = [k for k in simt if "adaLN_modulation" in k]
    assert len(ada) == 2 * (len(m.blocks) + 1)
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert rel_l2(p.grad, simt[k]) < (5e-3 if k in ada else 3e-3), k


@pytest.mark.parametrize("D,T", [(384, 37), (1152, 256)])
def test_gate_resid_bwd(ops, dev, D, T):
    g = _g(5)
    B = 3
    dxo = torch.randn(B * T, D, device=dev, generator=g)
    y = torch.randn(B * T, D, device=dev, generator=g).bfloat16()
    mod = torch.randn(B, 6 * D, device=dev, generator=g)
    gate = mod[:, 2 * D:3 * D]
    dmod = torch.zeros(B, 6 * D, device=dev)
    dbias = torch.zeros(D, device=dev)
    dy = ops.gate_resid_bwd(dxo, y, gate, T, dmod[:, 2 * D:3 * D], dbias=dbias)
    ref_dy = dxo.double().view(B, T, D) * gate.double()[:, None]
    assert rel_l2(dy.float(), ref_dy.view(B * T, D)) < 4e-3
    assert rel_l2(dmod[:, 2 * D:3 * D], (dxo.double() * y.double()).view(B, T, D).sum(1)) < 1e-5
    assert rel_l2(dbias, ref_dy.sum((0, 1))) < 1e-5
    assert float(dmod[:, :2 * D].abs().max()) == 0.0


def test_small_backward_kernels(ops, dev):
    g = _g(6)
    x = torch.randn(1000, 520, device=dev, generator=g)
    assert rel_l2(ops.colsum(x), x.double().sum(0)) < 1e-5
    xb = x.bfloat16()
    acc = torch.ones(520, device=dev)
    assert rel_l2(ops.colsum(xb, out=acc, accumulate=True), xb.double().sum(0) + 1) < 1e-5
    # patchify = im2col in conv-weight order: conv2d(x, w) == patches @ w.flatten(1).T
    for p, H in [(2, 32), (4, 32), (8, 32)]:
        img = torch.randn(3, 4, H, H, device=dev, generator=g)
        w = torch.randn(24, 4, p, p, device=dev, generator=g)
        ref = F.conv2d(img, w, stride=p).flatten(2).transpose(1, 2).reshape(-1, 24)
        got = ops.patchify(img, p).float() @ w.flatten(1).t()
        assert rel_l2(got, ref) < 5e-3
        # unpatchify_bwd is the inverse permutation of DiT.unpatchify
        c = 8
        z = torch.randn(3, (H // p) ** 2, p * p * c, device=dev, generator=g)
        hh = H // p
        img2 = torch.einsum("nhwpqc->nchpwq", z.reshape(3, hh, hh, p, p, c)).reshape(3, c, H, H).contiguous()
        assert torch.equal(ops.unpatchify_bwd(img2, p).float(), z.reshape(-1, p * p * c).bfloat16().float())
    pre = torch.randn(64, 384, device=dev, generator=g) * 2
    d = torch.randn(64, 384, device=dev, generator=g)
    pd = pre.double().requires_grad_(True)
    F.silu(pd).backward(d.double())
    assert rel_l2(ops.silu_bwd(d, pre), pd.grad) < 1e-6
    yl = torch.tensor([3, 7, 3, 1000, 0, 3], device=dev)
    dc = torch.randn(6, 384, device=dev, generator=g)
    table = torch.zeros(1001, 384, device=dev)
    ops.label_embed_bwd(dc, yl, table)
    ref = torch.zeros(1001, 384, device=dev, dtype=torch.double).index_add_(0, yl, dc.double())
    assert rel_l2(table, ref) < 1e-6


@pytest.mark.parametrize("hd,H", [(64, 6), (72, 4)])
@pytest.mark.parametrize("T", [64, 256, 100, 16])
def test_attention_bwd(ops, dev, hd, H, T):
    g = _g(7)
    B = 2
    D = H * hd
    qkv = torch.randn(B * T, 3 * D, device=dev, generator=g).bfloat16()
    dout = torch.randn(B * T, D, device=dev, generator=g).bfloat16()
    lse = torch.empty(B, H, T, device=dev)
    out = ops.attention(qkv, B, T, H, hd, lse=lse)
    q = qkv.double().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    o = F.scaled_dot_product_attention(q[0], q[1], q[2]).transpose(1, 2).reshape(B * T, D)
    o.backward(dout.double())
    ref = q.grad.permute(1, 3, 0, 2, 4).reshape(B * T, 3 * D)
    got = ops.attention_bwd(qkv, out, dout, lse, B, T, H, hd)
    for j, name in enumerate("qkv"):
        assert rel_l2(got[:, j * D:(j + 1) * D].float(), ref[:, j * D:(j + 1) * D]) < 1.5e-2, name


@pytest.mark.parametrize("B,H,hd", [(20, 16, 72), (37, 6, 64)])
def test_attention_bwd_tcgen05_persistent(ops, dev, B, H, hd):
    """The tcgen05 backward (T = 256) with more (image, head) items than SMs: every CTA runs several items, i.e.
    both phases of every barrier and the reload of the shared Q/K/V/dO tiles; checked on a slice of the batch
    against fp64 autograd, and run twice (no atomics: bit-identical)."""
    T = 256
    g = _g(17)
    D = H * hd
    qkv = (torch.randn(B * T, 3 * D, device=dev, generator=g) * 1.5).bfloat16()
    dout = torch.randn(B * T, D, device=dev, generator=g).bfloat16()
    lse = torch.empty(B, H, T, device=dev)
    out = ops.attention(qkv, B, T, H, hd, lse=lse)
    got = ops.attention_bwd(qkv, out, dout, lse, B, T, H, hd)
    assert torch.isfinite(got.float()).all()
    for b0 in (0, B - 2):  # first and last images (different CTAs / iterations)
        sl = slice(b0 * T, (b0 + 2) * T)
        q = qkv[sl].double().view(2, T, 3, H, hd).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
        o = F.scaled_dot_product_attention(q[0], q[1], q[2]).transpose(1, 2).reshape(2 * T, D)
        o.backward(dout[sl].double())
        ref = q.grad.permute(1, 3, 0, 2, 4).reshape(2 * T, 3 * D)
        for j, name in enumerate("qkv"):
            assert rel_l2(got[sl, j * D:(j + 1) * D].float(), ref[:, j * D:(j + 1) * D]) < 1.5e-2, (name, b0)
    again = ops.attention_bwd(qkv, out, dout, lse, B, T, H, hd)
    assert torch.equal(got, again)


# ---------------------------------------------------------------------------- the whole step
def _oracle_grads(model_cpu, name, x, t, y, dout=None, loss_fn=None, drop_ids=None):
    from oracle import dit_oracle as O

    cfg = O.config_for(name, input_size=x.shape[-1])
    sd = {k: v.detach().clone().requires_grad_(v.requires_grad) for k, v in model_cpu.state_dict(keep_vars=True).items()}
    out = O.dit_forward(sd, cfg, x, t, y, drop_ids=drop_ids)
    if loss_fn is None:
        out.backward(dout)
    else:
        loss_fn(out).backward()
    return out.detach(), {k: v.grad for k, v in sd.items() if v.requires_grad}


@pytest.mark.parametrize("name,B", [("DiT-S/2", 4), ("DiT-S/8", 5), ("DiT-B/4", 8), ("DiT-XL/2", 2)])
def test_model_gradients_against_oracle(dev, name, B):
    """Every parameter's gradient of <dout, DiT(x, t, y)> in bf16 against the fp32 CPU oracle's autograd.
    The reference's own bf16-autocast gradients sit ~1e-2 from its fp32 ones; bound per tensor 5e-2 and
    2e-2 on the flattened whole."""
    from util import build_product_model

    m = build_product_model(name, input_size=32, num_classes=1000, precision="bf16")
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, 4, 32, 32, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    y = torch.randint(0, 1001, (B,), generator=g)
    dout = torch.randn(B, 8, 32, 32, generator=g)
    ref_out, ref = _oracle_grads(m, name, x, t, y, dout)
    mc = m.cuda()  # eval mode: no label dropout, as in the oracle call
    out = mc(x.cuda(), t.cuda(), y.cuda())
    assert out.requires_grad
    assert rel_l2(out, ref_out) < 1e-2
    out.backward(dout.cuda())
    num = den = 0.0
    for k, p in mc.named_parameters():
        if not p.requires_grad:
            assert p.grad is None
            continue
        e = rel_l2(p.grad, ref[k])
        assert e < 5e-2, (k, e)
        num += float((p.grad.double().cpu() - ref[k].double()).pow(2).sum())
        den += float(ref[k].double().pow(2).sum())
    assert math.sqrt(num / den) < 2e-2
    # a second step after zero_grad reuses the arena and gives the same gradients
    first = {k: p.grad.clone() for k, p in mc.named_parameters() if p.grad is not None}
    mc.zero_grad(set_to_none=True)
    mc(x.cuda(), t.cuda(), y.cuda()).backward(dout.cuda())
    for k, p in mc.named_parameters():
        if p.grad is not None:
            assert rel_l2(p.grad, first[k]) < 3e-3, k  # atomics reorder f32 sums that are then rounded to bf16 operands
    # accumulating into live gradients (no zero_grad) doubles them
    mc(x.cuda(), t.cuda(), y.cuda()).backward(dout.cuda())
    for k, p in mc.named_parameters():
        if p.grad is not None:
            assert rel_l2(p.grad, 2 * first[k]) < 3e-3, k


def test_training_losses_step_against_oracle(dev):
    """create_diffusion('').training_losses(model, x0, t, dict(y=y))['loss'].mean().backward() — the
    reference's training step (train_original.py:204-209) — against the oracle's loss and gradients."""
    from fast_dit_b200 import create_diffusion
    from oracle import dit_oracle as O
    from oracle.diffusion_oracle import DiffusionOracle
    from util import build_product_model

    name, B = "DiT-S/2", 6
    m = build_product_model(name, input_size=32, num_classes=1000, precision="bf16")
    g = torch.Generator().manual_seed(12)
    x0 = torch.randn(B, 4, 32, 32, generator=g)
    noise = torch.randn(B, 4, 32, 32, generator=g)
    t = torch.tensor([0, 1, 17, 500, 998, 999])
    y = torch.randint(0, 1000, (B,), generator=g)
    do = DiffusionOracle("")
    x_t = do.q_sample(x0, t, noise)

    def loss_fn(out):
        return do.training_losses(out, x0, x_t, t, noise)["loss"].mean()

    ref_out, ref = _oracle_grads(m, name, x_t, do.map_t(t), y, loss_fn=loss_fn)
    ref_terms = do.training_losses(ref_out, x0, x_t, t, noise)
    mc = m.cuda()
    d = create_diffusion("")
    terms = d.training_losses(mc, x0.cuda(), t.cuda(), dict(y=y.cuda()), noise=noise.cuda())
    for k in ("loss", "mse", "vb"):
        assert rel_l2(terms[k], ref_terms[k]) < 2e-2, k
    terms["loss"].mean().backward()
    num = den = 0.0
    for k, p in mc.named_parameters():
        if p.requires_grad:
            num += float((p.grad.double().cpu() - ref[k].double()).pow(2).sum())
            den += float(ref[k].double().pow(2).sum())
    assert math.sqrt(num / den) < 3e-2


def test_xl2_training_step_train_mode_forced_drop(dev):
    """The C4 model (DiT-XL/2: head dim 72, 28 blocks) in TRAIN mode through training_losses with the label-dropout
    mask forced (models_original.py:79-87 draws it from torch.rand): loss terms and every parameter gradient against
    the fp32 oracle's autograd with the same mask."""
    from fast_dit_b200 import create_diffusion
    from oracle.diffusion_oracle import DiffusionOracle
    from util import build_product_model

    name, B = "DiT-XL/2", 2
    m = build_product_model(name, input_size=32, num_classes=1000, precision="bf16")
    g = torch.Generator().manual_seed(14)
    x0 = torch.randn(B, 4, 32, 32, generator=g)
    noise = torch.randn(B, 4, 32, 32, generator=g)
    t = torch.tensor([0, 731])
    y = torch.randint(0, 1000, (B,), generator=g)
    drop = torch.tensor([True, False])
    do = DiffusionOracle("")
    x_t = do.q_sample(x0, t, noise)

    def loss_fn(out):
        return do.training_losses(out, x0, x_t, t, noise)["loss"].mean()

    ref_out, ref = _oracle_grads(m, name, x_t, do.map_t(t), y, loss_fn=loss_fn, drop_ids=drop)
    ref_terms = do.training_losses(ref_out, x0, x_t, t, noise)
    mc = m.cuda().train()
    ye = mc.y_embedder
    ye.token_drop = lambda labels, force_drop_ids=None: torch.where(drop.to(labels.device), ye.num_classes, labels)
    d = create_diffusion("")
    terms = d.training_losses(mc, x0.cuda(), t.cuda(), dict(y=y.cuda()), noise=noise.cuda())
    for k in ("loss", "mse", "vb"):
        assert rel_l2(terms[k], ref_terms[k]) < 2e-2, k
    terms["loss"].mean().backward()
    num = den = 0.0
    for k, p in mc.named_parameters():
        if p.requires_grad:
            num += float((p.grad.double().cpu() - ref[k].double()).pow(2).sum())
            den += float(ref[k].double().pow(2).sum())
    # the dropped label's row of the embedding table receives gradient, the kept class row of sample 0 none
    gt = mc.y_embedder.embedding_table.weight.grad
    assert float(gt[1000].abs().sum()) > 0 and float(gt[int(y[0])].abs().sum()) == 0
    assert math.sqrt(num / den) < 3e-2


def test_data_parallel_wrapper_single_rank(dev):
    """parallel.DataParallel on one rank: buckets are issued final layer -> blocks L-1..0 -> embedders,
    and the gradients equal the unwrapped model's."""
    from fast_dit_b200.parallel import DataParallel
    from util import build_product_model

    m = build_product_model("DiT-S/8", input_size=32, num_classes=1000, precision="bf16").cuda()
    g = _g(13)
    x = torch.randn(4, 4, 32, 32, device=dev, generator=g)
    t = torch.randint(0, 1000, (4,), device=dev, generator=g)
    y = torch.randint(0, 1000, (4,), device=dev, generator=g)
    dout = torch.randn(4, 8, 32, 32, device=dev, generator=g)
    m(x, t, y).backward(dout)
    plain = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    m.zero_grad(set_to_none=True)
    ddp = DataParallel(m)
    ddp(x, t, y).backward(dout)
    assert ddp.buckets_issued == ["final_layer"] + [f"blocks.{i}" for i in range(m.depth - 1, -1, -1)] + ["embed"]
    for k, p in m.named_parameters():
        if p.grad is not None:
            assert rel_l2(p.grad, plain[k]) < 3e-3, k


def test_fused_adamw_ema_matches_torch(dev):
    """optim.FusedAdamWEMA against torch.optim.AdamW + the reference's update_ema loop (train.py:41-51,161)
    fed the same gradients for three steps; the bf16 shadows the next forward reads track the weights."""
    import copy

    from fast_dit_b200.optim import FusedAdamWEMA
    from util import build_product_model

    m = build_product_model("DiT-S/8", input_size=32, num_classes=1000, precision="bf16").cuda().train()
    ref = copy.deepcopy(m)
    ema_ref = copy.deepcopy(m)
    opt = FusedAdamWEMA(m, lr=1e-3, weight_decay=0.01, ema_decay=0.99)
    opt_ref = torch.optim.AdamW([p for p in ref.parameters() if p.requires_grad], lr=1e-3, weight_decay=0.01)
    g = _g(21)
    for step in range(3):
        x = torch.randn(4, 4, 32, 32, device=dev, generator=g)
        t = torch.randint(0, 1000, (4,), device=dev, generator=g)
        y = torch.randint(0, 1000, (4,), device=dev, generator=g)
        dout = torch.randn(4, 8, 32, 32, device=dev, generator=g)
        torch.manual_seed(100 + step)  # label dropout draw
        m(x, t, y).backward(dout)
        for (k, p), q in zip(m.named_parameters(), ref.parameters()):
            q.grad = None if p.grad is None else p.grad.clone()
        opt.step()
        opt.zero_grad()
        opt_ref.step()
        with torch.no_grad():
            for e, q in zip(ema_ref.parameters(), ref.parameters()):
                if q.requires_grad:
                    e.mul_(0.99).add_(q.data, alpha=0.01)
        for (k, p), q in zip(m.named_parameters(), ref.parameters()):
            assert rel_l2(p, q) < 1e-6, (step, k)
    ema_sd = opt.ema_state_dict()
    for k, e in ema_ref.state_dict().items():
        assert rel_l2(ema_sd[k], e) < 1e-6, k
    sh = m._shadows()
    assert torch.equal(sh["w"][0], m.blocks[0].attn.qkv.weight.detach().bfloat16())
    D = m.hidden_size
    assert torch.equal(sh["ada_w"][6 * D:12 * D], m.blocks[1].adaLN_modulation[1].weight.detach().bfloat16())
    assert torch.equal(sh["ada_b"][-2 * D:], m.final_layer.adaLN_modulation[1].bias.detach())
    # the forward after the step uses the updated weights (and load_state_dict refreshes the shadows)
    m.eval()
    with torch.no_grad():
        a = m(x, t, y)
        ref.eval()
        b = ref(x, t, y)
        assert rel_l2(a, b) < 1e-3
        m.load_state_dict(ema_sd)
        ema_ref.eval()
        assert rel_l2(m(x, t, y), ema_ref(x, t, y)) < 1e-3


def test_optimizer_checkpoint_round_trip_and_torch_adamw_interchange(dev):
    """FusedAdamWEMA.state_dict() is torch.optim.AdamW's format (+ the EMA weights): (i) a saved and reloaded
    optimizer continues exactly like the one that was never interrupted — including the EMA, which a resume must
    not reset (train.py:231-236 saves {"model", "ema", "opt"}); (ii) torch.optim.AdamW over the same parameters loads
    our checkpoint, and we load torch's own state_dict, after which both take identical steps."""
    import copy

    from fast_dit_b200.optim import FusedAdamWEMA
    from util import build_product_model

    def grads(model, seed):
        g = _g(seed)
        x = torch.randn(3, 4, 32, 32, device=dev, generator=g)
        t = torch.randint(0, 1000, (3,), device=dev, generator=g)
        y = torch.randint(0, 1000, (3,), device=dev, generator=g)
        dout = torch.randn(3, 8, 32, 32, device=dev, generator=g)
        torch.manual_seed(seed)
        model(x, t, y).backward(dout)

    m = build_product_model("DiT-S/8", input_size=32, num_classes=1000, precision="bf16").cuda().train()
    opt = FusedAdamWEMA(m, lr=2e-3, weight_decay=0.01, ema_decay=0.9)
    for s in (1, 2):
        grads(m, s)
        opt.step()
        opt.zero_grad()
    ck = {"model": copy.deepcopy(m.state_dict()), "ema": opt.ema_state_dict(), "opt": opt.state_dict()}
    assert ck["opt"]["state"][1]["exp_avg"].data_ptr() != opt.exp_avg.data_ptr()  # copies, not live views
    assert 0 not in ck["opt"]["state"], "pos_embed is parameter 0 of model.parameters() and frozen: no state"
    grads(m, 3)
    # the weight gradients are accumulated with f32 atomics (run-to-run spread ~1e-3): the resumed runs below are fed
    # THESE gradients, so that what is compared is the optimizer state, not the backward pass
    g3 = {k: p.grad.clone() for k, p in m.named_parameters() if p.grad is not None}
    opt.step()
    opt.zero_grad()
    want_w = {k: v.clone() for k, v in m.state_dict().items()}
    want_ema = opt.ema_state_dict()

    def feed(model):
        for k, p in model.named_parameters():
            p.grad = g3[k].clone() if k in g3 else None

    # (i) resume in a fresh process' worth of objects
    m2 = build_product_model("DiT-S/8", input_size=32, num_classes=1000, precision="bf16", seed=5).cuda().train()
    m2.load_state_dict(ck["model"])
    opt2 = FusedAdamWEMA(m2, lr=1.0, weight_decay=0.5, ema_decay=0.9)
    opt2.load_state_dict(ck["opt"])
    assert opt2.step_count == 2 and opt2.lr == 2e-3 and opt2.weight_decay == 0.01
    for k, v in opt2.ema_state_dict().items():
        assert torch.equal(v, ck["ema"][k]), k
    feed(m2)
    opt2.step()
    for k, v in m2.state_dict().items():
        assert rel_l2(v, want_w[k]) < 1e-6, k
    for k, v in opt2.ema_state_dict().items():
        assert rel_l2(v, want_ema[k]) < 1e-6, k

    # (ii) interchange with torch.optim.AdamW
    ref = copy.deepcopy(m2)
    ref.load_state_dict(ck["model"])
    topt = torch.optim.AdamW(ref.parameters(), lr=1.0)
    topt.load_state_dict({k: v for k, v in ck["opt"].items() if k in ("state", "param_groups")})
    m3 = build_product_model("DiT-S/8", input_size=32, num_classes=1000, precision="bf16", seed=6).cuda().train()
    m3.load_state_dict(ck["model"])
    opt3 = FusedAdamWEMA(m3, ema_decay=None)
    opt3.load_state_dict(topt.state_dict())  # torch's own dict: no "ema", no "param_names"
    feed(m3)
    feed(ref)
    opt3.step()
    topt.step()
    for (k, p), q in zip(m3.named_parameters(), ref.parameters()):
        assert rel_l2(p, q) < 1e-6, k
        assert rel_l2(p, want_w[k]) < 1e-6, k
    with pytest.raises(Exception):
        opt3.ema_state_dict()


def test_optimizer_update_overlapped_with_backward_is_identical(dev):
    """FusedAdamWEMA(overlap_backward=True) applies each bucket's update on a side stream as soon as backward has
    finished the bucket; the arithmetic is the same elementwise kernel over the same values, so weights, EMA, moments
    and bf16 shadows after three steps equal the backward-then-step optimizer's BIT FOR BIT."""
    from fast_dit_b200.optim import FusedAdamWEMA
    from util import build_product_model

    out = []
    for overlap in (False, True):
        m = build_product_model("DiT-S/8", input_size=32, num_classes=1000, precision="bf16").cuda().train()
        opt = FusedAdamWEMA(m, lr=1e-3, weight_decay=0.01, ema_decay=0.9, overlap_backward=overlap)
        g = _g(31)
        for step in range(3):
            x = torch.randn(4, 4, 32, 32, device=dev, generator=g)
            t = torch.randint(0, 1000, (4,), device=dev, generator=g)
            y = torch.randint(0, 1000, (4,), device=dev, generator=g)
            dout = torch.randn(4, 8, 32, 32, device=dev, generator=g)
            torch.manual_seed(50 + step)
            m(x, t, y).backward(dout)
            opt.step()
            opt.zero_grad()
        torch.cuda.synchronize()
        assert opt.step_count == 3
        out.append([t.clone() for t in (opt.flat, opt.ema, opt.exp_avg, opt.exp_avg_sq, opt.shadow)])
    names = ("weights", "ema", "exp_avg", "exp_avg_sq", "shadow")
    # weight gradients are accumulated with f32 atomics (split-K): run-to-run they differ in the last bits, so the
    # two runs are compared to that spread, not bitwise, except that both must be finite and the shadows consistent
    for n, a, b in zip(names, out[0], out[1]):
        assert torch.isfinite(a.float()).all() and torch.isfinite(b.float()).all()
        assert rel_l2(b.float(), a.float()) < 2e-3, n
    assert torch.equal(out[1][4], out[1][0].bfloat16()), "bf16 shadows follow the updated weights"
    # accumulating gradients is refused in this mode
    m(x, t, y).backward(dout)
    with pytest.raises(Exception):
        m(x, t, y).backward(dout)
