"""Per-kernel numerics on the GPU: every libditb200 entry point against a plain torch
fp32 statement of the same op (all of these are floating-point kernels).  End-to-end parity
against the oracle lives in test_parity_gpu.py."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


@pytest.fixture(scope="module")
def ops():
    from fast_dit_b200 import ops as o

    return o


def test_lib_loaded_and_initialised(dev):
    from fast_dit_b200 import _lib

    lib = _lib.ensure_init(0)
    assert lib.ditb200_sm_count() >= 100


@pytest.mark.parametrize("D", [384, 768, 1024, 1152, 320])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_ln_modulate(ops, dev, D, out_dtype):
    g = torch.Generator(device=dev).manual_seed(0)
    B, T = 3, 37
    x = torch.randn(B * T, D, device=dev, generator=g) * 2 + 0.3
    mod = torch.randn(B, 6 * D, device=dev, generator=g) * 0.5
    shift, scale = mod[:, :D], mod[:, D:2 * D]
    stats = torch.empty(B * T, 2, device=dev)
    y = ops.ln_modulate(x, shift, scale, T, out_dtype=out_dtype, stats=stats)
    ref = F.layer_norm(x, (D,), eps=1e-6).view(B, T, D) * (1 + scale[:, None]) + shift[:, None]
    ref = ref.view(B * T, D)
    tol = 2e-6 if out_dtype == torch.float32 else 4e-3
    assert rel_l2(y.float(), ref) < tol
    assert rel_l2(stats[:, 0], x.mean(1)) < 1e-5
    assert rel_l2(stats[:, 1], (x.var(1, unbiased=False) + 1e-6).rsqrt()) < 1e-5


@pytest.mark.parametrize("D,T", [(384, 37), (1152, 256), (768, 64), (1024, 16)])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_ln_modulate_resid(ops, dev, D, T, out_dtype):
    """x_out = x + gate * y fused in front of LayerNorm + modulate (models_original.py:120-121, 19-20)."""
    g = torch.Generator(device=dev).manual_seed(21)
    B = 3
    x = torch.randn(B * T, D, device=dev, generator=g) * 2 + 0.5
    y = torch.randn(B * T, D, device=dev, generator=g).bfloat16()
    mod = torch.randn(B, 6 * D, device=dev, generator=g) * 0.5
    gate, shift, scale = mod[:, 2 * D:3 * D], mod[:, :D], mod[:, D:2 * D]
    stats = torch.empty(B * T, 2, device=dev)
    x_out, h = ops.ln_modulate_resid(x, y, gate, shift, scale, T, out_dtype=out_dtype, stats=stats)
    xr = x.double() + gate.double().repeat_interleave(T, 0) * y.double()
    assert rel_l2(x_out, xr) < 1e-6
    ln = F.layer_norm(xr, (D,), eps=1e-6)
    ref = ln * (1 + scale.double().repeat_interleave(T, 0)) + shift.double().repeat_interleave(T, 0)
    assert rel_l2(h.float(), ref) < (1e-5 if out_dtype == torch.float32 else 4e-3)
    assert rel_l2(stats[:, 0], xr.mean(1)) < 1e-5
    only, none = ops.ln_modulate_resid(x, y, gate, shift, scale, T, want_out=False)
    assert none is None and torch.equal(only, x_out)


@pytest.mark.parametrize("p,C,H,D", [(2, 4, 32, 1152), (4, 4, 32, 768), (8, 4, 32, 384), (2, 4, 64, 384)])
def test_patch_embed(ops, dev, p, C, H, D):
    g = torch.Generator(device=dev).manual_seed(1)
    B = 3
    x = torch.randn(B, C, H, H, device=dev, generator=g)
    w = torch.randn(D, C, p, p, device=dev, generator=g) * 0.1
    b = torch.randn(D, device=dev, generator=g) * 0.1
    T = (H // p) ** 2
    pos = torch.randn(T, D, device=dev, generator=g)
    y = ops.patch_embed(x, w, b, pos, p)
    ref = F.conv2d(x, w, b, stride=p).flatten(2).transpose(1, 2) + pos[None]
    assert rel_l2(y.view(B, T, D), ref) < 2e-6
    yb = ops.patch_embed(x, w, b, pos, p, round_bf16=True)
    assert rel_l2(yb.view(B, T, D), ref) < 6e-3


def test_timestep_embedding(ops, dev):
    t = torch.tensor([0, 1, 500, 999, 123], device=dev)
    y = ops.timestep_embedding(t, 256)
    half = 128
    freqs = torch.exp(-math.log(10000) * torch.arange(half, dtype=torch.float32, device=dev) / half)
    args = t[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    assert (y - ref).abs().max() < 2e-4  # |arg| up to 999: a 1-ulp difference in freq moves cos by ~6e-5
    # known-answer values harvested from the reference
    assert abs(float(y[1, 0]) - 0.5403023362) < 1e-6
    assert abs(float(y[1, 128]) - 0.8414709568) < 1e-6
    assert abs(float(y[3, 127]) - 0.9942431450) < 2e-4


@pytest.mark.parametrize("M,N,K", [(4, 1152, 256), (64, 6912, 1152), (7, 130, 52)])
def test_small_linear(ops, dev, M, N, K):
    g = torch.Generator(device=dev).manual_seed(2)
    a = torch.randn(M, K, device=dev, generator=g)
    w = torch.randn(N, K, device=dev, generator=g) * 0.05
    b = torch.randn(N, device=dev, generator=g)
    add = torch.randn(M, N, device=dev, generator=g)
    y = ops.small_linear(a, w, b)
    assert rel_l2(y, F.linear(a, w, b)) < 2e-6
    y = ops.small_linear(a, w, b, silu_in=True)
    assert rel_l2(y, F.linear(F.silu(a), w, b)) < 2e-6
    y = ops.small_linear(a, w, b, silu_out=True, add=add)
    assert rel_l2(y, F.silu(F.linear(a, w, b)) + add) < 2e-6
    y = ops.small_linear(a, w.bfloat16(), b)
    assert rel_l2(y, F.linear(a, w.bfloat16().float(), b)) < 2e-6


def test_label_embed(ops, dev):
    g = torch.Generator(device=dev).manual_seed(3)
    table = torch.randn(1001, 384, device=dev, generator=g)
    y = torch.tensor([0, 5, 1000, 999], device=dev)
    add = torch.randn(4, 384, device=dev, generator=g)
    out = ops.label_embed(y, table, add)
    assert torch.equal(out, add + table[y])
    assert torch.equal(ops.label_embed(y, table), table[y])


def _gemm_ref(a, w, bias, epi, resid=None, gate=None, T=1):
    from fast_dit_b200 import _lib as L

    y = a.double() @ w.double().t() + bias.double()
    if epi == L.EPI_BIAS_GELU:
        y = F.gelu(y, approximate="tanh")
    elif epi == L.EPI_BIAS_SILU:
        y = F.silu(y)
    elif epi == L.EPI_BIAS_GATE_RESID:
        M = a.shape[0]
        gg = gate.double().repeat_interleave(T, dim=0)[:M]
        y = resid.double() + gg * y
    return y


@pytest.mark.parametrize("M,N,K", [(256, 384, 384), (200, 1152, 192), (1024, 1536, 384), (70, 72, 40)])
@pytest.mark.parametrize("epi", [0, 1, 2, 3])
def test_gemm_fp32(ops, dev, M, N, K, epi):
    from fast_dit_b200 import _lib as L

    g = torch.Generator(device=dev).manual_seed(4)
    a = torch.randn(M, K, device=dev, generator=g)
    w = torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)
    bias = torch.randn(N, device=dev, generator=g)
    T = 64
    B = (M + T - 1) // T
    resid = torch.randn(M, N, device=dev, generator=g)
    gate = torch.randn(B, 6 * N, device=dev, generator=g)[:, 2 * N:3 * N]
    kw = {}
    if epi == L.EPI_BIAS_GATE_RESID:
        kw = dict(resid=resid.clone(), gate=gate, rows_per_gate=T)
    y = ops.gemm(a, w, bias, epilogue=epi, **kw)
    ref = _gemm_ref(a, w, bias, epi, resid, gate, T)
    assert rel_l2(y, ref) < 2e-6


TC_SHAPES = [
    (256, 256, 64),      # one k-block
    (256, 384, 384),
    (1024, 1152, 1152),  # DiT-XL proj
    (1000, 3456, 1152),  # ragged M, QKV
    (512, 4608, 1152),   # fc1
    (512, 1152, 4608),   # fc2, long K
    (64, 6912, 1152),    # adaLN: M below one tile
    (300, 200, 72),      # ragged everything (N % 8 == 0, K % 8 == 0)
]


@pytest.mark.parametrize("cta_group,tile_n", [(0, 0), (1, 128), (1, 192), (1, 256), (2, 128), (2, 192), (2, 256)])
@pytest.mark.parametrize("M,N,K", TC_SHAPES)
def test_gemm_tcgen05_bias(ops, dev, M, N, K, cta_group, tile_n):
    g = torch.Generator(device=dev).manual_seed(5)
    a = (torch.randn(M, K, device=dev, generator=g)).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    ref = _gemm_ref(a, w, bias, 0)
    y32 = ops.gemm(a, w, bias, out_dtype=torch.float32, tile_n=tile_n, cta_group=cta_group)
    assert rel_l2(y32, ref) < 1e-5, "f32 output of the bf16 tensor-core GEMM"
    y16 = ops.gemm(a, w, bias, tile_n=tile_n, cta_group=cta_group)
    assert y16.dtype == torch.bfloat16
    assert rel_l2(y16.float(), ref) < 4e-3


@pytest.mark.parametrize("tile_n", [128, 256])
@pytest.mark.parametrize("M,N,K", [(1024, 1152, 1152), (1000, 3456, 1152), (2048, 512, 4608), (300, 200, 72)])
def test_gemm_tcgen05_multicast_cluster(ops, dev, M, N, K, tile_n):
    """cta_group=4: two CTA pairs per cluster share the B tile through TMA multicast (ragged M: the second
    pair's rows may lie past the matrix)."""
    g = torch.Generator(device=dev).manual_seed(15)
    a = (torch.randn(M, K, device=dev, generator=g)).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    ref = _gemm_ref(a, w, bias, 0)
    y32 = ops.gemm(a, w, bias, out_dtype=torch.float32, tile_n=tile_n, cta_group=4)
    assert rel_l2(y32, ref) < 1e-5
    same = ops.gemm(a, w, bias, out_dtype=torch.float32, tile_n=tile_n, cta_group=2)
    assert torch.equal(y32, same), "same tile shape, same accumulation order: bit-identical to the pair kernel"


def test_gemm_tcgen05_narrow_last_column_matches_full_tiles(ops, dev):
    """N = 1152 with 256-wide tiles ends in a 128-wide column that runs as a narrow tcgen05.mma under the
    LPT schedule; the result must be bit-identical to covering N with 128-wide tiles (same k order)."""
    g = torch.Generator(device=dev).manual_seed(16)
    M, N, K = 4096, 1152, 1152
    a = (torch.randn(M, K, device=dev, generator=g)).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    y256 = ops.gemm(a, w, bias, out_dtype=torch.float32, tile_n=256, cta_group=2)
    y128 = ops.gemm(a, w, bias, out_dtype=torch.float32, tile_n=128, cta_group=2)
    assert torch.equal(y256, y128)
    assert rel_l2(y256, _gemm_ref(a, w, bias, 0)) < 1e-5


@pytest.mark.parametrize("M,N,K,cg,bn", [(16384, 3456, 1152, 2, 256), (16384, 1152, 1152, 2, 192), (8200, 1160, 1152, 0, 0),
                                        (16384, 1152, 1152, 1, 256), (40000, 384, 64, 2, 128)])
def test_gemm_dynamic_tile_scheduler_matches_static(ops, dev, M, N, K, cg, bn):
    """ditb200_gemm_args.dynamic_sched = 1: one cluster per tile, running clusters cancel and absorb the pending ones through
    cluster launch control.  Same tiles, same k order: the result must be bit-identical to the static schedule."""
    g = torch.Generator(device=dev).manual_seed(23)
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    kw = dict(tile_n=bn, cta_group=cg) if cg else {}
    y_static = ops.gemm(a, w, bias, out_dtype=torch.bfloat16, **kw)
    prev = ops.set_gemm_dynamic(True)
    try:
        for _ in range(3):  # back-to-back launches: cancelled clusters of one grid must not leak into the next
            y_dyn = ops.gemm(a, w, bias, out_dtype=torch.bfloat16, **kw)
        torch.cuda.synchronize()
    finally:
        ops.set_gemm_dynamic(prev)
    assert torch.equal(y_static, y_dyn)
    assert rel_l2(y_dyn.float()[:2048], _gemm_ref(a[:2048], w, bias, 0)) < 1e-2


@pytest.mark.parametrize("epi", [1, 2, 3])
@pytest.mark.parametrize("M,N,K,T", [(1024, 1152, 1152, 256), (768, 1536, 384, 64), (200, 384, 1536, 16)])
def test_gemm_tcgen05_epilogues(ops, dev, M, N, K, T, epi):
    from fast_dit_b200 import _lib as L

    g = torch.Generator(device=dev).manual_seed(6)
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    B = (M + T - 1) // T
    resid = torch.randn(M, N, device=dev, generator=g)
    gate = torch.randn(B, 6 * N, device=dev, generator=g)[:, 2 * N:3 * N]
    ref = _gemm_ref(a, w, bias, epi, resid, gate, T)
    if epi == L.EPI_BIAS_GATE_RESID:
        x = resid.clone()
        y = ops.gemm(a, w, bias, epilogue=epi, resid=x, gate=gate, rows_per_gate=T)
        assert y.data_ptr() == x.data_ptr(), "gated residual updates the stream in place"
        assert rel_l2(y, ref) < 1e-5
    else:
        y = ops.gemm(a, w, bias, epilogue=epi, out_dtype=torch.float32)
        assert rel_l2(y, ref) < 2e-4  # fast exp in the activation
        y = ops.gemm(a, w, bias, epilogue=epi)
        assert rel_l2(y.float(), ref) < 4e-3


def _attn_ref(qkv, B, T, H, hd):
    q, k, v = qkv.double().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    o = F.scaled_dot_product_attention(q, k, v)
    return o.transpose(1, 2).reshape(B * T, H * hd)


@pytest.mark.parametrize("B,T,H,hd", [(2, 256, 6, 64), (2, 256, 16, 72), (3, 64, 12, 64), (2, 16, 6, 64),
                                      (1, 1024, 4, 72), (2, 100, 3, 72)])
def test_attention_f32(ops, dev, B, T, H, hd):
    g = torch.Generator(device=dev).manual_seed(7)
    qkv = torch.randn(B * T, 3 * H * hd, device=dev, generator=g)
    lse = torch.empty(B, H, T, device=dev)
    o = ops.attention(qkv, B, T, H, hd, lse=lse)
    assert rel_l2(o, _attn_ref(qkv, B, T, H, hd)) < 2e-6
    q, k, _ = qkv.double().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    ref_lse = torch.logsumexp(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    assert rel_l2(lse, ref_lse) < 1e-5


@pytest.mark.parametrize("B,T,H,hd", [(2, 256, 6, 64), (2, 256, 16, 72), (3, 64, 12, 64), (2, 16, 6, 64),
                                      (1, 1024, 4, 72), (2, 100, 3, 72)])
def test_attention_bf16(ops, dev, B, T, H, hd):
    g = torch.Generator(device=dev).manual_seed(8)
    qkv = torch.randn(B * T, 3 * H * hd, device=dev, generator=g).bfloat16()
    lse = torch.empty(B, H, T, device=dev)
    o = ops.attention(qkv, B, T, H, hd, lse=lse)
    ref = _attn_ref(qkv.float(), B, T, H, hd)
    assert rel_l2(o.float(), ref) < 6e-3
    q, k, _ = qkv.double().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    ref_lse = torch.logsumexp(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    assert rel_l2(lse, ref_lse) < 1e-4


@pytest.mark.parametrize("B,T,H,hd", [(1, 128, 1, 64), (3, 128, 5, 72), (20, 256, 16, 72), (40, 128, 12, 64),
                                      (11, 256, 16, 80), (1, 512, 1, 64), (2, 768, 3, 72), (5, 1024, 16, 72),
                                      (40, 512, 6, 64)])
def test_attention_tcgen05_persistent(ops, dev, B, T, H, hd):
    """The tcgen05/TMEM forward kernels (single score tile for T in {128, 256}; KV-blocked online softmax for
    T = 512, 768, 1024): more work items than SMs, so every CTA runs several items through its K/V ring and
    both barrier phases; hd 72/80 exercise the zero-filled second channel chunk; scores scaled up so the row
    maximum (and the rescaling of the running output) matters."""
    g = torch.Generator(device=dev).manual_seed(18)
    qkv = (torch.randn(B * T, 3 * H * hd, device=dev, generator=g) * 2.0).bfloat16()
    lse = torch.empty(B, H, T, device=dev)
    o = ops.attention(qkv, B, T, H, hd, lse=lse)
    assert torch.isfinite(o.float()).all()
    ref = _attn_ref(qkv.float(), B, T, H, hd)
    assert rel_l2(o.float(), ref) < 6e-3
    q, k, _ = qkv.double().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    ref_lse = torch.logsumexp(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1)
    assert rel_l2(lse, ref_lse) < 1e-4
    again = ops.attention(qkv, B, T, H, hd)
    assert torch.equal(o, again), "no atomics, fixed schedule: run-to-run identical"


@pytest.mark.parametrize("T,hd,growth", [(1024, 72, 1.0), (1024, 72, 0.25), (512, 64, 1.0), (768, 72, -1.0)])
def test_attention_kv_blocked_running_maximum(ops, dev, T, hd, growth):
    """The KV-blocked forward (T > 256) moves its reference maximum lazily: only when a row outgrew it by more than
    2^8 (csrc/attention_tc.cu).  Keys whose magnitude rises with the key index make every 128-key block raise the
    row maxima — by far more than 2^8 for growth = 1 (rescale of the running output at every block), by less for
    growth = 0.25 (blocks evaluated against a stale maximum); growth < 0: the first block holds the maximum and
    all later ones sit far below it."""
    B, H = 3, 5
    g = torch.Generator(device=dev).manual_seed(23)
    qkv = torch.randn(B, T, 3, H, hd, device=dev, generator=g)
    blk = torch.arange(T, device=dev) // 128
    ramp = 1.0 + growth * blk.float() if growth > 0 else 4.0 / (1.0 + blk.float())
    q_dir = torch.randn(B, 1, H, hd, device=dev, generator=g)
    qkv[:, :, 0] += 2.0 * q_dir                                        # queries share a direction per head ...
    qkv[:, :, 1] = qkv[:, :, 1] * 0.5 + q_dir * ramp[None, :, None, None]  # ... along which the keys grow block by block
    qkv = qkv.reshape(B * T, 3 * H * hd).bfloat16()
    lse = torch.empty(B, H, T, device=dev)
    o = ops.attention(qkv, B, T, H, hd, lse=lse)
    assert torch.isfinite(o.float()).all() and torch.isfinite(lse).all()
    q, k, _ = qkv.double().view(B, T, 3, H, hd).permute(2, 0, 3, 1, 4).unbind(0)
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    bmax = s.view(B, H, T, T // 128, 128).amax(-1).cummax(-1).values     # running maximum after each block
    jump = (bmax[..., 1:] - bmax[..., :-1]).amax() * 1.4426950408889634
    if growth == 1.0:
        assert jump > 8.0, "the data must push some running maximum past the 2^8 threshold"
    assert rel_l2(o.float(), _attn_ref(qkv.float(), B, T, H, hd)) < 6e-3
    assert rel_l2(lse, torch.logsumexp(s, dim=-1)) < 1e-4


@pytest.mark.parametrize("p,Cout,T,D", [(2, 8, 256, 1152), (4, 8, 64, 768), (8, 8, 16, 384), (2, 4, 256, 384),
                                         (2, 8, 100, 384), (2, 8, 49, 768), (2, 8, 1024, 1024), (4, 2, 64, 1152)])
def test_final_layer(ops, dev, p, Cout, T, D):
    g = torch.Generator(device=dev).manual_seed(9)
    B = 3
    x = torch.randn(B * T, D, device=dev, generator=g)
    mod = torch.randn(B, 2 * D, device=dev, generator=g) * 0.3
    shift, scale = mod[:, :D], mod[:, D:]
    NO = p * p * Cout
    w = torch.randn(NO, D, device=dev, generator=g) / math.sqrt(D)
    b = torch.randn(NO, device=dev, generator=g)
    y = ops.final_layer(x, shift, scale, w, b, T, p, Cout)
    h = F.layer_norm(x, (D,), eps=1e-6).view(B, T, D) * (1 + scale[:, None]) + shift[:, None]
    z = F.linear(h, w, b)
    hp = int(T ** 0.5)
    z = z.reshape(B, hp, hp, p, p, Cout)
    ref = torch.einsum("nhwpqc->nchpwq", z).reshape(B, Cout, hp * p, hp * p)
    assert rel_l2(y, ref) < 3e-6
    yb = ops.final_layer(x, shift, scale, w, b, T, p, Cout, round_bf16=True)
    assert rel_l2(yb, ref) < 8e-3


def test_cfg_combine(ops, dev):
    g = torch.Generator(device=dev).manual_seed(10)
    raw = torch.randn(6, 8, 32, 32, device=dev, generator=g)
    out = ops.cfg_combine(raw, 3, 4.0)
    eps, rest = raw[:, :3], raw[:, 3:]
    c, u = torch.split(eps, 3, dim=0)
    half = u + 4.0 * (c - u)
    ref = torch.cat([torch.cat([half, half], 0), rest], 1)
    assert torch.equal(out, ref), "CFG combine is bit-exact (no FMA contraction)"


def test_cast_bf16(ops, dev):
    g = torch.Generator(device=dev).manual_seed(11)
    for n in (1, 3, 4, 1027, 1 << 20):
        x = torch.randn(n, device=dev, generator=g)
        assert torch.equal(ops.cast_bf16(x), x.bfloat16())


def _guarded(shape, dtype, dev, fill=0.0):
    """An output tensor carved out of a larger allocation whose head and tail (4 KB each) hold a sentinel pattern: a
    kernel that writes outside its output corrupts them.  compute-sanitizer is closed on this GPU pool
    (profiles/r02_sanitizer.md), so out-of-bounds WRITES are hunted this way on ragged shapes."""
    n = 1
    for s in shape:
        n *= s
    pad = 4096 // torch.empty(0, dtype=dtype).element_size()
    buf = torch.full((n + 2 * pad,), 12345.0, device=dev, dtype=dtype)
    out = buf[pad:pad + n].view(shape)
    out.fill_(fill)
    return buf, out, pad


def _guards_intact(buf, pad):
    return bool((buf[:pad] == 12345.0).all()) and bool((buf[-pad:] == 12345.0).all())


@pytest.mark.parametrize("M,N,K", [(300, 200, 72), (1000, 3456, 1152), (257, 1160, 64), (129, 136, 1152), (8200, 1160, 128)])
def test_gemm_epilogues_do_not_write_outside_the_output(ops, dev, M, N, K):
    from fast_dit_b200 import _lib as L

    g = torch.Generator(device=dev).manual_seed(40)
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    T = 64
    gate = torch.randn((M + T - 1) // T, N, device=dev, generator=g)
    for epi, dt in ((L.EPI_BIAS, torch.bfloat16), (L.EPI_BIAS, torch.float32), (L.EPI_BIAS_GELU, torch.bfloat16),
                    (L.EPI_BIAS_GATE_RESID, torch.float32)):
        for kw in ({}, dict(tile_n=256, cta_group=2), dict(tile_n=128, cta_group=1), dict(reverse_m=True)):
            buf, out, pad = _guarded((M, N), dt, dev)
            if epi == L.EPI_BIAS_GATE_RESID:
                ops.gemm(a, w, bias, epilogue=epi, resid=out, gate=gate, rows_per_gate=T, **kw)
            else:
                ops.gemm(a, w, bias, epilogue=epi, out=out, **kw)
            torch.cuda.synchronize()
            assert _guards_intact(buf, pad), (epi, dt, kw)
            assert torch.isfinite(out.float()).all()
    # weight gradient: split-K with f32 atomics into an accumulating output
    dy = torch.randn(M, N, device=dev, generator=g).bfloat16()
    buf, out, pad = _guarded((N, K), torch.float32, dev)
    ops.gemm(dy, a, None, out=out, trans_a=True, trans_w=True, split_k=3)
    torch.cuda.synchronize()
    assert _guards_intact(buf, pad)
    assert rel_l2(out, dy.double().t() @ a.double()) < 1e-4


@pytest.mark.parametrize("B,T,H,hd", [(3, 128, 5, 72), (2, 256, 16, 72), (1, 512, 3, 72), (2, 100, 3, 72), (3, 64, 12, 64)])
def test_attention_and_layernorm_do_not_write_outside_their_outputs(ops, dev, B, T, H, hd):
    g = torch.Generator(device=dev).manual_seed(41)
    D = H * hd
    qkv = torch.randn(B * T, 3 * D, device=dev, generator=g).bfloat16()
    for rev in (False, True):
        buf, out, pad = _guarded((B * T, D), torch.bfloat16, dev)
        ops.attention(qkv, B, T, H, hd, out=out, reverse=rev)
        torch.cuda.synchronize()
        assert _guards_intact(buf, pad), ("attention", rev)
        assert rel_l2(out.float(), _attn_ref(qkv.float(), B, T, H, hd)) < 6e-3
    x = torch.randn(B * T, D, device=dev, generator=g)
    mod = torch.randn(B, 3 * D, device=dev, generator=g)
    for dt in (torch.bfloat16, torch.float32):
        buf, out, pad = _guarded((B * T, D), dt, dev)
        ops.ln_modulate(x, mod[:, :D], mod[:, D:2 * D], T, out_dtype=dt, out=out, reverse=True)
        torch.cuda.synchronize()
        assert _guards_intact(buf, pad), ("ln_modulate", dt)
    if D in (384, 768, 1024, 1152):
        y = torch.randn(B * T, D, device=dev, generator=g).bfloat16()
        buf, xo, pad = _guarded((B * T, D), torch.float32, dev)
        ops.ln_modulate_resid(x, y, mod[:, 2 * D:], mod[:, :D], mod[:, D:2 * D], T, x_out=xo)
        torch.cuda.synchronize()
        assert _guards_intact(buf, pad), "ln_modulate_resid"
