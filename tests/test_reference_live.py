"""Checks against the LIVE reference, only where /root/reference exists (the build container)."""
import os
import sys

import pytest
import torch

from util import ROOT

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(REF), reason="reference tree not present on this box")


@pytest.fixture(scope="module")
def MO():
    sys.path[:0] = [os.path.join(ROOT, "oracle", "timm_standin"), os.path.join(REF, "train_options")]
    import models_original
    return models_original


@pytest.mark.parametrize("name,kw", [("DiT-S/2", dict(input_size=32)), ("DiT-B/8", dict(input_size=16, num_classes=7)),
                                     ("DiT-S/4", dict(input_size=32, learn_sigma=False, class_dropout_prob=0.0))])
def test_init_is_bit_identical_to_reference(MO, name, kw):
    from fast_dit_b200 import DiT_models
    torch.manual_seed(3)
    ref = MO.DiT_models[name](**kw)
    torch.manual_seed(3)
    mine = DiT_models[name](**kw)
    a, b = ref.state_dict(), mine.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert torch.equal(a[k], b[k]), k


def test_oracle_equals_reference_forward(MO):
    from oracle import dit_oracle as O
    torch.manual_seed(0)
    ref = MO.DiT_models["DiT-S/4"](input_size=32).eval()
    O.rerandomise_zero_params(ref.named_parameters())
    g = torch.Generator().manual_seed(5)
    x, t, y = torch.randn(3, 4, 32, 32, generator=g), torch.randint(0, 1000, (3,), generator=g), torch.randint(0, 1000, (3,), generator=g)
    with torch.no_grad():
        assert torch.equal(ref(x, t, y), O.dit_forward(ref.state_dict(), O.config_for("DiT-S/4", input_size=32), x, t, y))


@pytest.fixture(scope="module")
def RD():
    """The reference's own `diffusion` package (pure numpy/torch, importable here)."""
    sys.path[:0] = [REF]
    import diffusion as ref_diffusion
    return ref_diffusion


SPECS = ["", "1000", "500", "250", "100", "50", "10", "1", "ddim25", "ddim50", "ddim100", "ddim250", "10,10,10,10",
         "25,10,5", "4,3,2,1,1", "300,300,300"]


@pytest.mark.parametrize("spec", SPECS)
def test_space_timesteps_equals_reference(RD, spec):
    from diffusion.respace import space_timesteps as ref_space
    from fast_dit_b200.diffusion import space_timesteps

    assert space_timesteps(1000, spec or [1000]) == ref_space(1000, spec or [1000])


@pytest.mark.parametrize("spec", ["ddim600", "ddim999", "2000", "600,600"])
def test_space_timesteps_rejects_what_the_reference_rejects(RD, spec):
    from diffusion.respace import space_timesteps as ref_space
    from fast_dit_b200.diffusion import space_timesteps

    with pytest.raises(ValueError):
        ref_space(1000, spec)
    with pytest.raises(ValueError):
        space_timesteps(1000, spec)


@pytest.mark.parametrize("spec", ["", "250", "ddim50", "25,10,5"])
@pytest.mark.parametrize("kw", [dict(), dict(noise_schedule="squaredcos_cap_v2"), dict(learn_sigma=False),
                                dict(predict_xstart=True), dict(sigma_small=True, learn_sigma=False),
                                dict(rescale_learned_sigmas=True), dict(diffusion_steps=400)])
def test_create_diffusion_equals_reference(RD, spec, kw):
    """Same enum choices, same kept timesteps and bit-identical float64 tables for every constructor switch."""
    import numpy as np
    from fast_dit_b200 import create_diffusion

    ref = RD.create_diffusion(spec, **kw)
    mine = create_diffusion(spec, **kw)
    assert list(mine.timestep_map) == list(ref.timestep_map)
    assert mine.num_timesteps == ref.num_timesteps and mine.original_num_steps == ref.original_num_steps
    for enum in ("model_mean_type", "model_var_type", "loss_type"):
        assert getattr(mine, enum).name == getattr(ref, enum).name
    for name in ("betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
                 "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
                 "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
                 "posterior_mean_coef1", "posterior_mean_coef2"):
        a, b = np.asarray(getattr(ref, name)), np.asarray(getattr(mine, name))
        assert a.dtype == b.dtype and np.array_equal(a, b), name


@pytest.mark.parametrize("spec,kw", [("250", {}), ("", {}), ("100", {"noise_schedule": "squaredcos_cap_v2"}),
                                     ("250", {"learn_sigma": False}), ("50", {"learn_sigma": False, "sigma_small": True}),
                                     ("250", {"predict_xstart": True}), ("ddim50", {})])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_diffusion_oracle_equals_reference_on_random_inputs(RD, spec, kw, seed):
    """The oracle's step arithmetic against the live reference on fresh random inputs (the golden fixtures pin a
    handful of fixed ones): p_mean_variance, p_sample, q_sample bit-exact; training losses to float rounding."""
    import numpy as np
    from oracle.diffusion_oracle import DiffusionOracle

    ref = RD.create_diffusion(spec, **kw)
    d = DiffusionOracle(spec, **kw)
    g = torch.Generator().manual_seed(100 + seed)
    n, C = 5, 4
    learn = kw.get("learn_sigma", True)
    x = torch.randn(n, C, 8, 8, generator=g) * 1.3
    x0 = torch.randn(n, C, 8, 8, generator=g)
    out = torch.randn(n, 2 * C if learn else C, 8, 8, generator=g) * (2.0 if seed == 2 else 0.7)
    noise = torch.randn(n, C, 8, 8, generator=g)
    t = torch.randint(0, ref.num_timesteps, (n,), generator=g)
    t[0] = 0  # the step that drops the noise term / switches to the decoder NLL
    stub = lambda x_, t_, **_: out  # noqa: E731
    for clip in (False, True):
        a = ref.p_mean_variance(stub, x, t, clip_denoised=clip)
        b = d.p_mean_variance(out, x, t, clip_denoised=clip)
        for k in ("mean", "variance", "log_variance", "pred_xstart"):
            assert torch.equal(a[k], b[k]), (k, clip)
        torch.manual_seed(7 + seed)
        a = ref.p_sample(stub, x, t, clip_denoised=clip)
        torch.manual_seed(7 + seed)
        nz = torch.randn_like(x)
        assert torch.equal(a["sample"], d.p_sample(out, x, t, nz, clip_denoised=clip)["sample"])
    assert torch.equal(ref.q_sample(x0, t, noise), d.q_sample(x0, t, noise))
    if not spec.startswith("ddim"):
        a = ref.training_losses(stub, x0, t, noise=noise)
        b = d.training_losses(out, x0, d.q_sample(x0, t, noise), t, noise)
        for k in a:
            assert np.allclose(a[k].numpy(), b[k].numpy(), rtol=1e-6, atol=1e-7), k


@pytest.mark.parametrize("name,kw,lat", [("DiT-S/2", dict(input_size=16), 16), ("DiT-B/8", dict(input_size=32, num_classes=10), 32),
                                         ("DiT-S/4", dict(input_size=32, learn_sigma=False), 32),
                                         ("DiT-S/8", dict(input_size=64, in_channels=3), 64)])
def test_oracle_equals_reference_forward_variants(MO, name, kw, lat):
    """forward, forward_with_cfg and the training-mode label dropout (same torch RNG draw) on several geometries."""
    from oracle import dit_oracle as O
    torch.manual_seed(1)
    ref = MO.DiT_models[name](**kw).eval()
    O.rerandomise_zero_params(ref.named_parameters())
    cfg = O.config_for(name, **kw)
    classes, C = kw.get("num_classes", 1000), kw.get("in_channels", 4)
    g = torch.Generator().manual_seed(11)
    x, t = torch.randn(4, C, lat, lat, generator=g), torch.randint(0, 1000, (4,), generator=g)
    y = torch.randint(0, classes, (4,), generator=g)
    sd = ref.state_dict()
    with torch.no_grad():
        assert torch.equal(ref(x, t, y), O.dit_forward(sd, cfg, x, t, y))
        ycfg = torch.cat([y[:2], torch.full((2,), classes)])
        assert torch.equal(ref.forward_with_cfg(x, t, ycfg, 4.0), O.dit_forward_with_cfg(sd, cfg, x, t, ycfg, 4.0))
        # training mode: LabelEmbedder.token_drop draws torch.rand(N) < p from the global generator (MO:79-87)
        ref.train()
        torch.manual_seed(5)
        out_train = ref(x, t, y)
        torch.manual_seed(5)
        drop = torch.rand(4) < ref.y_embedder.dropout_prob
        assert torch.equal(out_train, O.dit_forward(sd, cfg, x, t, y, drop_ids=drop))
