"""Host-side mirror of the reference interface: registry, module attributes, state_dict
layout, respacing, tables.  Runs without a GPU."""
import numpy as np
import pytest
import torch

from util import golden

from fast_dit_b200 import DiT_models, create_diffusion
from fast_dit_b200.diffusion import space_timesteps
from fast_dit_b200.diffusion import gaussian_diffusion as gd


def test_registry_has_the_references_twelve_models():
    assert sorted(DiT_models) == sorted(f"DiT-{s}/{p}" for s in ("S", "B", "L", "XL") for p in (2, 4, 8))
    import fast_dit_b200.models as M
    assert M.DiT_XL_2 is DiT_models["DiT-XL/2"] and M.DiT_S_8 is DiT_models["DiT-S/8"]


def test_module_attributes_and_state_dict_layout():
    """SURVEY.md §8(b) and Appendix B."""
    m = DiT_models["DiT-S/2"](input_size=32, num_classes=1000)
    assert (m.in_channels, m.out_channels, m.patch_size, m.num_heads, m.learn_sigma) == (4, 8, 2, 6, True)
    assert m.x_embedder.num_patches == 256 and m.y_embedder.num_classes == 1000
    sd = m.state_dict()
    D = 384
    expect = {
        "pos_embed": (1, 256, D), "x_embedder.proj.weight": (D, 4, 2, 2), "x_embedder.proj.bias": (D,),
        "t_embedder.mlp.0.weight": (D, 256), "t_embedder.mlp.2.weight": (D, D),
        "y_embedder.embedding_table.weight": (1001, D),
        "blocks.0.attn.qkv.weight": (3 * D, D), "blocks.0.attn.qkv.bias": (3 * D,),
        "blocks.11.attn.proj.weight": (D, D), "blocks.5.mlp.fc1.weight": (4 * D, D),
        "blocks.5.mlp.fc2.weight": (D, 4 * D), "blocks.3.adaLN_modulation.1.weight": (6 * D, D),
        "final_layer.linear.weight": (32, D), "final_layer.adaLN_modulation.1.weight": (2 * D, D),
    }
    for k, shp in expect.items():
        assert tuple(sd[k].shape) == shp, k
    assert len(sd) == 1 + 2 + 4 + 1 + 12 * 10 + 4
    assert not m.pos_embed.requires_grad
    assert sum(p.numel() for p in m.parameters()) == 32963360
    # default init is adaLN-Zero: gates and the output layer start at exactly zero (MO:207-216)
    assert float(sd["blocks.0.adaLN_modulation.1.weight"].abs().max()) == 0.0
    assert float(sd["final_layer.linear.weight"].abs().max()) == 0.0
    m2 = DiT_models["DiT-S/2"](input_size=32, num_classes=1000, learn_sigma=False, class_dropout_prob=0.0)
    assert m2.out_channels == 4 and m2.y_embedder.embedding_table.weight.shape[0] == 1000


def test_state_dict_roundtrip_and_deepcopy():
    import copy
    m = DiT_models["DiT-S/8"](input_size=16, num_classes=10)
    ema = copy.deepcopy(m)  # train.py:153 does this
    ema.load_state_dict(m.state_dict())
    for (k1, p1), (k2, p2) in zip(m.named_parameters(), ema.named_parameters()):
        assert k1 == k2 and torch.equal(p1, p2)


@pytest.mark.parametrize("n,spec,expect_len", [(1000, "250", 250), (1000, "10", 10), (1000, "ddim50", 50),
                                               (300, [10, 15, 20], 45), (1000, "", None)])
def test_space_timesteps(n, spec, expect_len):
    if spec == "":
        d = create_diffusion("")
        assert d.num_timesteps == 1000 and d.timestep_map == list(range(1000))
        return
    s = space_timesteps(n, spec)
    assert len(s) == expect_len
    if spec == [10, 15, 20]:
        assert sorted(s)[:12] == [0, 11, 22, 33, 44, 55, 66, 77, 88, 99, 100, 107]


def test_space_timesteps_errors():
    with pytest.raises(ValueError):
        space_timesteps(1000, "ddim600")  # no integer stride gives exactly 600 steps
    with pytest.raises(ValueError):
        space_timesteps(10, "20")


def test_tables_match_reference_fixtures():
    fx = golden("diffusion_tables.npz")
    for key in sorted({k.rsplit("|", 1)[0] for k in fx.files}):
        sched, spec = key.split("|")
        d = create_diffusion(spec, noise_schedule=sched)
        assert np.array_equal(np.array(d.timestep_map), fx[key + "|timestep_map"])
        for t in ("betas", "alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                  "posterior_variance", "posterior_log_variance_clipped", "posterior_mean_coef1",
                  "posterior_mean_coef2", "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod"):
            assert np.array_equal(getattr(d, t), fx[f"{key}|{t}"]), (key, t)


def test_create_diffusion_type_selection():
    d = create_diffusion("250")
    assert d.model_mean_type == gd.ModelMeanType.EPSILON and d.model_var_type == gd.ModelVarType.LEARNED_RANGE
    assert d.loss_type == gd.LossType.MSE and d.num_timesteps == 250
    assert create_diffusion("", learn_sigma=False).model_var_type == gd.ModelVarType.FIXED_LARGE
    assert create_diffusion("", learn_sigma=False, sigma_small=True).model_var_type == gd.ModelVarType.FIXED_SMALL
    assert create_diffusion("", predict_xstart=True).model_mean_type == gd.ModelMeanType.START_X
    assert create_diffusion("", use_kl=True).loss_type == gd.LossType.RESCALED_KL
    assert create_diffusion("", rescale_learned_sigmas=True).loss_type == gd.LossType.RESCALED_MSE


def test_model_refuses_cpu_inputs():
    from fast_dit_b200._lib import Ditb200Error
    m = DiT_models["DiT-S/8"](input_size=16, num_classes=10).eval()
    with pytest.raises(Ditb200Error):
        m(torch.randn(2, 4, 16, 16), torch.zeros(2, dtype=torch.long), torch.zeros(2, dtype=torch.long))


# ------------------------------------------------------------------ GEMM tile schedule (host replay of the kernel's TileSched)
def _schedule(M, N, tile_m, bn, part, split_k, pairs):
    import ctypes
    from fast_dit_b200 import _lib as L

    lib = L.load()
    cap = 1 << 16
    buf = (ctypes.c_int * (5 * cap))()
    n = lib.ditb200_debug_tile_schedule(M, N, tile_m, bn, part, split_k, pairs, ctypes.cast(buf, ctypes.c_void_p), cap)
    assert 0 < n <= cap
    return np.frombuffer(buf, dtype=np.int32)[: 5 * n].reshape(n, 5).copy()


@pytest.mark.parametrize("M,N,bn,part,split_k", [
    (16384, 3456, 256, 128, 1),   # qkv at C3: 13 full columns + a 128-wide one
    (16384, 1152, 256, 128, 1),   # proj
    (16384, 1152, 192, 0, 1),     # fc2: exact 192-wide cover
    (16384, 4608, 256, 0, 1),     # fc1
    (8200, 1160, 192, 16, 1),     # ragged rows, 16-wide last column
    (1152, 4608, 256, 0, 4),      # weight gradient, split-K
    (300, 384, 256, 128, 1),      # fewer tiles than CTA pairs
])
def test_gemm_static_schedule_covers_every_tile_once(M, N, bn, part, split_k):
    """Every (tile row, tile column, k part) is processed by exactly one CTA pair, the narrow last column has the
    width the host computed, and the longest-first order leaves no pair more than one full tile behind."""
    tile_m, pairs = 256, 74
    rows = _schedule(M, N, tile_m, bn, part, split_k, pairs)
    m_tiles, n_tiles = -(-M // tile_m), -(-N // bn)
    keys = [tuple(r[[1, 2, 4]]) for r in rows]
    assert len(keys) == len(set(keys)) == m_tiles * n_tiles * split_k
    assert {k[0] for k in keys} == set(range(m_tiles)) and {k[1] for k in keys} == set(range(n_tiles))
    last = rows[rows[:, 2] == n_tiles - 1]
    assert (last[:, 3] == (part if part else bn)).all()
    assert (rows[rows[:, 2] < n_tiles - 1][:, 3] == bn).all()
    # balance in columns of work: a narrow tile never counts for less than half a full one (it still loads all of A)
    cost = np.where(rows[:, 3] < bn, np.maximum(rows[:, 3], bn // 2), bn) / split_k
    load = np.bincount(rows[:, 0], weights=cost, minlength=pairs)
    used = load[load > 0]
    assert used.max() - used.min() <= bn + 1e-9
    # within a pair: full tiles first, narrow ones last
    for p in range(pairs):
        w = rows[rows[:, 0] == p][:, 3]
        assert (np.diff(w) <= 0).all()


@pytest.mark.parametrize("name,M,N,K,trans_w,expect", [
    ("qkv", 16384, 3456, 1152, 0, (2, 256, 128)),
    ("proj", 16384, 1152, 1152, 0, (2, 256, 128)),   # short k loop: 256 + narrow column (39.5 vs 41.4 us measured)
    ("fc1", 16384, 4608, 1152, 0, (2, 256, 0)),
    ("fc2", 16384, 1152, 4608, 0, (2, 192, 0)),      # long k loop: exact 192-wide cover (134.1 vs 136.4 us measured)
    ("adaLN", 64, 66816, 1152, 0, (1, 256, 0)),       # M <= 128: single-CTA tiles
    ("odd", 8200, 1160, 1152, 0, (2, 192, 16)),
])
def test_gemm_tile_chooser_pins_the_measured_choices(name, M, N, K, trans_w, expect):
    """The automatic tile choice for the C3 shapes is the one the B200 measurements in DESIGN.md section 4 support."""
    import ctypes
    from fast_dit_b200 import _lib as L

    out = (ctypes.c_int * 3)()
    assert L.load().ditb200_debug_gemm_plan(M, N, K, trans_w, 1, 148, ctypes.cast(out, ctypes.c_void_p)) == 0
    assert tuple(out) == expect, name


@pytest.mark.parametrize("T,hd,backward,expect", [
    (256, 72, 0, 1), (128, 72, 0, 1), (256, 64, 0, 1), (256, 80, 0, 1),       # whole-row tcgen05 forward
    (1024, 72, 0, 2), (512, 64, 0, 2), (768, 72, 0, 2),                        # K/V-blocked tcgen05 forward (C5)
    (64, 72, 0, 0), (100, 64, 0, 0), (300, 72, 0, 0),                          # every other T: mma.sync flash kernel
    (256, 72, 1, 1), (256, 64, 1, 1), (1024, 72, 1, 0), (64, 72, 1, 0),        # backward
    (256, 96, 0, -1), (64, 80, 0, -1), (256, 80, 1, -1),                       # head dims without a kernel
])
def test_attention_dispatch(T, hd, backward, expect):
    """Which kernel family ditb200_attention_fwd / _bwd pick for (T, hd): DESIGN.md section 4's table, pinned."""
    from fast_dit_b200 import _lib as L

    assert L.load().ditb200_debug_attention_path(T, hd, backward) == expect


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the unmodified reference from baseline/_ref where it is installed, else the oracle
    port, on the host cores) runs without a GPU and prints ONE JSON line with the keys the driver reads; same metric /
    unit / workload naming as the GPU arm."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--ref-images", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "img/s" and d["higher_is_better"] is True
    assert "250-step" in d["metric"] and d["config"]["workload"].startswith("DiT-XL/2")
    assert d["value"] > 0 and d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0
    installed = os.path.isfile(os.path.join(root, "baseline", "_ref", "models_original.py"))
    assert d["cpu_baseline"]["kind"] == ("reference" if installed else "port")
    assert d["cpu_baseline"]["cores"] >= 1 and d["gpu_launches"] == 0
    assert d["config"]["workload"].endswith("random-init weights") and "32 kept images/GPU" in d["config"]["workload"]


@pytest.mark.parametrize("M,N,K,expect_table", [(16384, 1152, 4608, True), (16384, 1152, 1152, True), (8192, 3456, 1152, True),
                                                (16384, 4608, 1152, False), (16384, 3456, 1152, True), (1024, 1152, 1152, False),
                                                (16384, 1160, 1152, False)])
def test_gemm_explicit_two_width_schedule(M, N, K, expect_table):
    """Where mixing 256-wide and (256 + N mod 256) / 2-wide tiles balances better than any uniform cover, the GEMM runs an
    explicit per-pair tile list (N = 1152 on 74 pairs: 42 pairs x four 256-wide + 32 pairs x five 192-wide tiles instead
    of six rounds of 192).  Host replay: every output element is covered exactly once, widths are legal MMA widths, and
    the longest pair is at least 2 % shorter than under the formula schedule."""
    import ctypes
    from fast_dit_b200 import _lib as L

    lib = L.load()
    cap = 4096
    buf = (ctypes.c_int * (4 * cap))()
    loads = (ctypes.c_int * 2)()
    n = lib.ditb200_debug_gemm_table(M, N, K, 74, ctypes.cast(buf, ctypes.c_void_p), cap, ctypes.cast(loads, ctypes.c_void_p))
    assert (n > 0) == expect_table
    if not expect_table:
        return
    rows = np.frombuffer(buf, dtype=np.int32)[: 4 * n].reshape(n, 4)
    m_tiles = -(-M // 256)
    cover = np.zeros((m_tiles, N // 16), dtype=np.int32)
    for p, m, c0, w in rows:
        assert 0 <= p < 74 and 0 <= m < m_tiles and w % 16 == 0 and 128 <= w <= 256 and c0 % 16 == 0 and c0 + w <= N
        cover[m, c0 // 16:(c0 + w) // 16] += 1
    assert (cover == 1).all()
    per_pair = np.bincount(rows[:, 0], weights=rows[:, 3] + 104, minlength=74)
    assert per_pair.max() == loads[1] and loads[1] * 100 <= loads[0] * 98
    if (M, N) == (16384, 1152):
        assert n == 56 * 5 + 8 * 6 and sorted(set(per_pair.tolist())) == [4 * 360.0, 5 * 296.0]


def test_adaln_weight_gradient_dispatch(monkeypatch):
    """Training takes the outer-product kernel for the adaLN weight gradient up to 64 images per GPU (C4: 32) and the
    tcgen05 GEMM over the batch above (C2: 256); shapes the kernel does not tile never take it."""
    from fast_dit_b200 import ops

    monkeypatch.delenv("DITB200_ADALN_SIMT_MAX_N", raising=False)
    assert ops.adaln_wgrad_preferred(32, 2 * 1152, 1152) and ops.adaln_wgrad_preferred(64, 2 * 768, 768)
    assert not ops.adaln_wgrad_preferred(65, 2 * 768, 768) and not ops.adaln_wgrad_preferred(256, 2 * 768, 768)
    assert ops.adaln_wgrad_ok(256, 2 * 768, 768) and not ops.adaln_wgrad_ok(257, 2 * 768, 768)
    assert not ops.adaln_wgrad_preferred(8, 2 * 200, 200)  # D not a multiple of 128
    monkeypatch.setenv("DITB200_ADALN_SIMT_MAX_N", "256")
    assert ops.adaln_wgrad_preferred(256, 2 * 768, 768)
