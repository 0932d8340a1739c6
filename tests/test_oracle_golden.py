"""Pin the CPU oracle against fixtures recorded from the unmodified reference
(tools/gen_golden.py).  Runs without a GPU."""
import numpy as np
import pytest
import torch

from util import build_product_model, check_checksums, golden, rel_l2

from oracle import dit_oracle as O
from oracle.diffusion_oracle import DiffusionOracle

TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
          "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
          "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
          "posterior_mean_coef1", "posterior_mean_coef2"]


def _spec(s):
    return [int(v) for v in s.split(",")] if "," in s else s


def test_oracle_tables_bit_exact():
    fx = golden("diffusion_tables.npz")
    keys = sorted({k.rsplit("|", 1)[0] for k in fx.files})
    assert len(keys) == 10
    for key in keys:
        sched, spec = key.split("|")
        d = DiffusionOracle(_spec(spec), noise_schedule=sched)
        assert np.array_equal(np.array(d.timestep_map), fx[key + "|timestep_map"])
        for t in TABLES:
            assert np.array_equal(getattr(d.tab, t), fx[f"{key}|{t}"]), (key, t)


def test_survey_appendix_a_known_answers():
    """Spot values quoted in SURVEY.md Appendix A (harvested from the reference)."""
    d = DiffusionOracle("250")
    assert d.timestep_map[:5] == [0, 4, 8, 12, 16] and d.timestep_map[-3:] == [991, 995, 999]
    assert abs(d.tab.betas[1] - 0.000599065564476) < 1e-15
    assert abs(d.tab.posterior_mean_coef2[-1] - 0.960455307481) < 1e-11
    assert DiffusionOracle("10").timestep_map == [0, 111, 222, 333, 444, 555, 666, 777, 888, 999]
    pe = O.sincos_pos_embed(1152, 16)
    assert pe.shape == (256, 1152)
    assert abs(pe[1, 0] - 0.8414709848) < 1e-9 and pe[1, 576] == 0 and pe[16, 0] == 0 and abs(pe[16, 288] - 1) < 1e-12
    assert abs(pe[17, 0] - 0.8414709848) < 1e-9 and abs(pe[17, 864] - 0.5403023059) < 1e-9
    te = O.timestep_embedding(torch.tensor([0, 1, 500, 999]), 256)
    assert abs(float(te[1, 0]) - 0.5403023362) < 1e-7 and abs(float(te[1, 128]) - 0.8414709568) < 1e-7
    assert abs(float(te[3, 127]) - 0.9942431450) < 1e-6 and abs(float(te[3, 255]) - 0.1071472168) < 1e-6


def test_oracle_tiny_model_forward():
    fx = golden("dit_tiny.npz")
    sd = {k[3:]: torch.from_numpy(fx[k]) for k in fx.files if k.startswith("sd.")}
    kw = {k[3:]: fx[k].item() for k in fx.files if k.startswith("kw.")}
    cfg = O.DiTConfig(**kw)
    x, t, y = (torch.from_numpy(fx[k]) for k in ("x", "t", "y"))
    with torch.no_grad():
        assert rel_l2(O.dit_forward(sd, cfg, x, t, y), fx["out"]) < 1e-6
        assert rel_l2(O.dit_forward_with_cfg(sd, cfg, x, t, torch.from_numpy(fx["ycfg"]), float(fx["cfg_scale"])),
                      fx["out_cfg"]) < 1e-6
        drop = torch.from_numpy(fx["drop"]) == 1
        assert rel_l2(O.dit_forward(sd, cfg, x, t, y, drop_ids=drop), fx["out_drop"]) < 1e-6


@pytest.mark.parametrize("tag,name,lat", [("s2_seed0", "DiT-S/2", 32), ("s8_seed0", "DiT-S/8", 32),
                                          ("b4_seed0", "DiT-B/4", 32)])
def test_oracle_seeded_models(tag, name, lat):
    fx = golden(f"dit_{tag}.npz")
    m = build_product_model(name, input_size=lat, num_classes=1000)
    assert sum(p.numel() for p in m.parameters()) == int(fx["nparams"])
    check_checksums(m, fx)  # construction under the same seed reproduces the reference's weights
    cfg = O.config_for(name, input_size=lat)
    x, t, y = (torch.from_numpy(fx[k]) for k in ("x", "t", "y"))
    with torch.no_grad():
        assert rel_l2(O.dit_forward(m.state_dict(), cfg, x, t, y), fx["out"]) < 1e-6
        assert rel_l2(O.dit_forward_with_cfg(m.state_dict(), cfg, x, t, torch.from_numpy(fx["ycfg"]), 4.0),
                      fx["out_cfg"]) < 1e-6


def test_oracle_diffusion_kats():
    fx = golden("diffusion_kat.npz")
    x, out8, noise, x0 = (torch.from_numpy(fx[k]) for k in ("x", "out8", "noise", "x0"))
    cases = {"lr250": ("250", {}), "lr1000": ("", {}), "fl250": ("250", {"learn_sigma": False}),
             "fs250": ("250", {"learn_sigma": False, "sigma_small": True}),
             "x0_250": ("250", {"predict_xstart": True}),
             "cos100": ("100", {"noise_schedule": "squaredcos_cap_v2"})}
    for tag, (spec, kw) in cases.items():
        d = DiffusionOracle(spec, **kw)
        t = torch.from_numpy(fx[tag + "|t"])
        mo = out8 if kw.get("learn_sigma", True) else out8[:, :4].contiguous()
        for clip in (False, True):
            c = f"{tag}|clip{int(clip)}|"
            pm = d.p_mean_variance(mo, x, t, clip_denoised=clip)
            for k in ("mean", "variance", "log_variance", "pred_xstart"):
                assert np.array_equal(pm[k].numpy(), fx[c + "pmv." + k]), (c, k)
            ps = d.p_sample(mo, x, t, noise, clip_denoised=clip)
            assert np.array_equal(ps["sample"].numpy(), fx[c + "p_sample"]), c
        assert np.array_equal(d.q_sample(x0, t, noise).numpy(), fx[tag + "|q_sample"])
        if tag + "|tl.loss" in fx.files:
            tl = d.training_losses(mo, x0, d.q_sample(x0, t, noise), t, noise)
            for k in ("loss", "mse", "vb"):
                assert np.allclose(tl[k].numpy(), fx[f"{tag}|tl.{k}"], rtol=1e-6, atol=1e-7), (tag, k)


def test_oracle_sample_loop_matches_reference_trajectory():
    """BASELINE.json configs[0]: DiT-S/2, 10-step CFG sampling on CPU."""
    fx = golden("sample_s2_10step.npz")
    m = build_product_model("DiT-S/2", input_size=32, num_classes=1000)
    cfg = O.config_for("DiT-S/2", input_size=32)
    sd = m.state_dict()
    d = DiffusionOracle("10")
    z, y = torch.from_numpy(fx["z"]), torch.from_numpy(fx["y"])
    base = int(fx["noise_seed_base"])
    k = {"i": 0}

    def noise(i, x):
        g = torch.Generator().manual_seed(base + k["i"])
        k["i"] += 1
        return torch.randn(x.shape, generator=g)

    with torch.no_grad():
        final, traj = d.p_sample_loop(lambda x, t, **kw: O.dit_forward_with_cfg(sd, cfg, x, t, kw["y"], 4.0),
                                      z.shape, z, noise, clip_denoised=False, model_kwargs=dict(y=y))
    for i, s in enumerate(traj):
        assert rel_l2(s, fx["traj"][i]) < 1e-5, i


def _fork_case(tag):
    """Inputs of a dit_fork_*.npz fixture: the large dino_feat tensor is replayed from the stored generator seed."""
    fx = golden(f"dit_{tag}.npz")
    kw = {k[3:]: fx[k].item() for k in fx.files if k.startswith("kw.")}
    g = torch.Generator().manual_seed(int(fx["gen_seed"]))
    n, lat = fx["x"].shape[0], kw["input_size"]
    x = torch.randn(n, 4, lat, lat, generator=g)
    dino = torch.randn(n, kw["dino_feat_size"], lat, lat, generator=g)
    assert np.array_equal(x.numpy(), fx["x"]) and abs(float(dino.double().sum()) - float(fx["dino_sum"])) < 1e-6
    return fx, kw, x, dino, torch.from_numpy(fx["t"]), torch.from_numpy(fx["y"])


@pytest.mark.parametrize("tag", ["fork_small", "fork_p4"])
def test_oracle_fork_dino_variant(tag):
    """oracle/dit_dino_oracle.py against the UNMODIFIED fork model (/root/reference/models.py: 9-chunk adaLN, DINO
    cross-attention in the 14th and 16th block, c = t), and the product module's construction: same parameter count,
    same state_dict keys, bit-identical initial weights under the same seed."""
    from fast_dit_b200.models_dino import DiT, DiT_models
    from oracle.dit_dino_oracle import dit_dino_forward

    fx, kw, x, dino, t, y = _fork_case(tag)
    torch.manual_seed(0)
    m = DiT(**kw)
    O.rerandomise_zero_params(m.named_parameters())
    m.eval()
    check_checksums(m, fx)
    assert sum(p.numel() for p in m.parameters()) == int(fx["nparams"])
    cfg = O.DiTConfig(**{k: v for k, v in kw.items() if k != "dino_feat_size"})
    with torch.no_grad():
        out = dit_dino_forward(m.state_dict(), cfg, x, t, dino, y)
    assert rel_l2(out, fx["out"]) < 1e-6
    assert sorted(DiT_models) == sorted(f"DiT-{s}/{p}" for s in ("S", "B", "L", "XL") for p in (2, 4, 8))
