"""The rest of GaussianDiffusion's public surface on the CUDA path, against fixtures recorded from the UNMODIFIED
reference (tools/gen_golden.py gen_diffusion_api -> tests/golden/diffusion_api.npz): the standalone q / p helpers,
classifier guidance (cond_fn) through p_sample and ddim_sample, ddim_reverse_sample, the KL loss family,
_vb_terms_bpd, _prior_bpd, calc_bpd_loop and the MSE loss for fixed-variance and x0-predicting models.

Tolerances: kernels that only multiply / add table entries are written with separately rounded operations in the
reference's order -> bit-exact; anything through exp / log / tanh -> 2e-6 .. 3e-5 relative (libm vs CUDA)."""
import numpy as np
import pytest
import torch

from util import golden, rel_l2

pytestmark = pytest.mark.gpu

CASES = {"lr250": ("250", {}), "lr1000": ("", {}), "fl250": ("250", {"learn_sigma": False}),
         "fs250": ("250", {"learn_sigma": False, "sigma_small": True}), "x0_250": ("250", {"predict_xstart": True}),
         "cos100": ("100", {"noise_schedule": "squaredcos_cap_v2"})}


def cond_fn_fixture(x, t, **kw):
    """Same exactly-rounded arithmetic as tools/gen_golden.py cond_fn_fixture."""
    return x * 0.5 - 0.25 + (t.float() * 0.001).view(-1, 1, 1, 1)


def _cuda(fx, *keys):
    return [torch.from_numpy(fx[k]).cuda() for k in keys]


def _eq(got, want):
    return torch.equal(got.cpu(), torch.from_numpy(want))


@pytest.mark.parametrize("tag", list(CASES))
def test_standalone_helpers_bit_exact(tag):
    from fast_dit_b200 import create_diffusion

    fx = golden("diffusion_api.npz")
    spec, kw = CASES[tag]
    d = create_diffusion(spec, **kw)
    x, out8, x0 = _cuda(fx, "x", "out8", "x0")
    t = torch.from_numpy(fx[tag + "|t"]).cuda()
    for k, v in zip(("mean", "variance", "log_variance"), d.q_mean_variance(x0, t)):
        assert v.shape == x0.shape and _eq(v, fx[f"{tag}|qmv.{k}"]), k
    for k, v in zip(("mean", "variance", "log_variance"), d.q_posterior_mean_variance(x0, x, t)):
        assert _eq(v, fx[f"{tag}|qpost.{k}"]), k
    assert _eq(d._predict_xstart_from_eps(x, t, out8[:, :4].contiguous()), fx[tag + "|x0_from_eps"])
    assert _eq(d._predict_eps_from_xstart(x, t, x0), fx[tag + "|eps_from_x0"])
    assert rel_l2(d._prior_bpd(x0), fx[tag + "|prior_bpd"]) < 2e-6


@pytest.mark.parametrize("tag", list(CASES))
def test_classifier_guidance_and_ddim_reverse(tag, monkeypatch):
    from fast_dit_b200 import create_diffusion
    from fast_dit_b200.diffusion import gaussian_diffusion as gd

    fx = golden("diffusion_api.npz")
    spec, kw = CASES[tag]
    d = create_diffusion(spec, **kw)
    x, out8, noise = _cuda(fx, "x", "out8", "noise")
    t = torch.from_numpy(fx[tag + "|t"]).cuda()
    mo = out8 if kw.get("learn_sigma", True) else out8[:, :4].contiguous()
    stub = lambda *a, **k: mo  # noqa: E731
    monkeypatch.setattr(gd, "_randn_like", lambda z: noise.clone())
    for clip in (False, True):
        c = f"{tag}|clip{int(clip)}|"
        r = d.p_sample(stub, x, t, clip_denoised=clip, cond_fn=cond_fn_fixture, model_kwargs={})
        assert rel_l2(r["sample"], fx[c + "p_sample_cond"]) < 2e-6
        assert _eq(r["pred_xstart"], fx[c + "p_sample_cond.pred"])
        r = d.ddim_sample(stub, x, t, clip_denoised=clip, cond_fn=cond_fn_fixture, model_kwargs={}, eta=0.3)
        assert rel_l2(r["sample"], fx[c + "ddim_cond"]) < 3e-6
        assert _eq(r["pred_xstart"], fx[c + "ddim_cond.pred"])
        r = d.ddim_reverse_sample(stub, x, t, clip_denoised=clip)
        assert rel_l2(r["sample"], fx[c + "ddim_rev"]) < 2e-6
        assert _eq(r["pred_xstart"], fx[c + "ddim_rev.pred"])
        r = d.ddim_reverse_sample(stub, x, t, clip_denoised=clip, cond_fn=cond_fn_fixture, model_kwargs={})
        assert rel_l2(r["sample"], fx[c + "ddim_rev_cond"]) < 2e-6
    with pytest.raises(AssertionError):
        d.ddim_reverse_sample(stub, x, t, eta=0.5)


@pytest.mark.parametrize("tag", list(CASES))
def test_vb_terms_and_loss_families(tag):
    from fast_dit_b200 import create_diffusion
    from fast_dit_b200.diffusion.gaussian_diffusion import LossType

    fx = golden("diffusion_api.npz")
    spec, kw = CASES[tag]
    x, out8, noise, x0, w = _cuda(fx, "x", "out8", "noise", "x0", "w")
    t = torch.from_numpy(fx[tag + "|t"]).cuda()
    mo = out8 if kw.get("learn_sigma", True) else out8[:, :4].contiguous()
    d = create_diffusion(spec, **kw)
    x_t = d.q_sample(x0, t, noise=noise)
    for clip in (False, True):
        c = f"{tag}|clip{int(clip)}|"
        r = d._vb_terms_bpd(lambda *a, **k: mo, x0, x_t, t, clip_denoised=clip)
        assert np.allclose(r["output"].cpu().numpy(), fx[c + "vb.output"], rtol=3e-5, atol=1e-6), c
        assert rel_l2(r["pred_xstart"], fx[c + "vb.pred"]) < 1e-6
    # MSE family: every mean / variance type, loss terms and the gradient wrt the model output
    mo_g = mo.clone().requires_grad_(True)
    tl = d.training_losses(lambda *a, **k: mo_g, x0, t, noise=noise)
    want_keys = {k.split("mse.")[1] for k in fx.files if k.startswith(tag + "|mse.")} - {"grad"}
    assert set(tl) == want_keys
    for k in tl:
        assert np.allclose(tl[k].detach().cpu().numpy(), fx[f"{tag}|mse.{k}"], rtol=3e-5, atol=1e-6), (tag, k)
    (tl["loss"] * w).sum().backward()
    assert rel_l2(mo_g.grad, fx[tag + "|mse.grad"]) < 3e-5
    # KL family: the gradient reaches the mean channels too
    for lt in ("RESCALED_KL", "KL"):
        dk = create_diffusion(spec, use_kl=True, **kw)
        assert dk.loss_type == LossType.RESCALED_KL
        dk.loss_type = getattr(LossType, lt)
        mo_g = mo.clone().requires_grad_(True)
        tl = dk.training_losses(lambda *a, **k: mo_g, x0, t, noise=noise)
        assert set(tl) == {"loss"}
        assert np.allclose(tl["loss"].detach().cpu().numpy(), fx[f"{tag}|{lt}.loss"], rtol=3e-5, atol=1e-5), (tag, lt)
        (tl["loss"] * w).sum().backward()
        assert rel_l2(mo_g.grad, fx[f"{tag}|{lt}.grad"]) < 5e-5, (tag, lt)


@pytest.mark.parametrize("tag,kw", [("bpd10", {}), ("bpd10_fl", {"learn_sigma": False})])
def test_calc_bpd_loop(tag, kw, monkeypatch):
    from fast_dit_b200 import create_diffusion
    from fast_dit_b200.diffusion import gaussian_diffusion as gd

    fx = golden("diffusion_api.npz")
    out8, x0 = _cuda(fx, "out8", "x0")
    mo = out8 if kw.get("learn_sigma", True) else out8[:, :4].contiguous()
    d = create_diffusion("10", **kw)
    step = {"i": 0}

    def seeded(z):
        g = torch.Generator().manual_seed(500 + step["i"])
        step["i"] += 1
        return torch.randn(z.shape, generator=g).to(z.device)

    monkeypatch.setattr(gd, "_randn_like", seeded)
    for clip in (True, False):
        step["i"] = 0
        r = d.calc_bpd_loop(lambda xx, ts, **k: mo * (1.0 + ts.float().view(-1, 1, 1, 1) * 0.001), x0,
                            clip_denoised=clip, model_kwargs={})
        assert set(r) == {"total_bpd", "prior_bpd", "vb", "xstart_mse", "mse"}
        for k, v in r.items():
            want = fx[f"{tag}|clip{int(clip)}|{k}"]
            assert v.shape == want.shape, k
            assert np.allclose(v.cpu().numpy(), want, rtol=5e-5, atol=1e-5), (tag, clip, k)
