"""Generate tests/golden/*.npz from the UNMODIFIED reference (build container only).

    python tools/gen_golden.py

Imports /root/reference/train_options/models_original.py and /root/reference/diffusion with
oracle/timm_standin on sys.path (timm is not installed here), runs them on CPU in fp32 and
records small input/output fixtures.  The fixtures travel to the GPU box; /root/reference does
not.  Protocol: torch.manual_seed(0) -> construct -> re-randomise every
all-zero parameter with N(0, 0.02^2) from Generator(1234) -> eval().
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle", "timm_standin"), os.path.join(REF, "train_options"), REF]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import diffusion as ref_diffusion  # noqa: E402  (the reference's package)
import models_original as MO  # noqa: E402
from diffusion import gaussian_diffusion as ref_gd  # noqa: E402
from oracle.dit_oracle import rerandomise_zero_params  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
torch.backends.cuda.matmul.allow_tf32 = False
torch.set_num_threads(os.cpu_count())

TABLES = ["betas", "alphas_cumprod", "alphas_cumprod_prev", "alphas_cumprod_next", "sqrt_alphas_cumprod",
          "sqrt_one_minus_alphas_cumprod", "log_one_minus_alphas_cumprod", "sqrt_recip_alphas_cumprod",
          "sqrt_recipm1_alphas_cumprod", "posterior_variance", "posterior_log_variance_clipped",
          "posterior_mean_coef1", "posterior_mean_coef2"]


def build(name=None, seed=0, **kw):
    torch.manual_seed(seed)
    m = MO.DiT_models[name](**kw) if name else MO.DiT(**kw)
    rerandomise_zero_params(m.named_parameters())
    return m.eval()


def checksums(m):
    return {"ck." + k: np.array([float(v.double().sum()), float(v.double().abs().sum())])
            for k, v in m.state_dict().items()}


def inputs(n, lat, seed=0, classes=1000):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(n, 4, lat, lat, generator=g), torch.randint(0, 1000, (n,), generator=g),
            torch.randint(0, classes, (n,), generator=g))


def gen_tables():
    d = {}
    for spec in ["", "250", "10", "ddim50", "25,10,5"]:
        for sched in ["linear", "squaredcos_cap_v2"]:
            df = ref_diffusion.create_diffusion(spec, noise_schedule=sched)
            key = f"{sched}|{spec}"
            d[key + "|timestep_map"] = np.array(df.timestep_map, dtype=np.int64)
            for t in TABLES:
                d[key + "|" + t] = getattr(df, t)
    np.savez_compressed(os.path.join(OUT, "diffusion_tables.npz"), **d)
    print("diffusion_tables", len(d))


def gen_tiny():
    """A whole tiny model with its weights: covers D % 128 != 0, one-k-block GEMMs, T = 16."""
    kw = dict(input_size=8, patch_size=2, in_channels=4, hidden_size=64, depth=2, num_heads=1, num_classes=10)
    m = build(None, **kw)
    x, t, y = inputs(6, 8, seed=3, classes=10)
    with torch.no_grad():
        out = m(x, t, y)
        ycfg = torch.cat([y[:3], torch.full((3,), 10)])
        out_cfg = m.forward_with_cfg(x, t, ycfg, 1.5)
        drop = torch.tensor([1, 0, 0, 1, 0, 1])
        m.train()
        # training-mode forward with a forced drop mask (token_drop's force_drop_ids path, MO:84-85)
        c_t = m.t_embedder(t)
        c = c_t + m.y_embedder(y, True, force_drop_ids=drop)
        h = m.x_embedder(x) + m.pos_embed
        for blk in m.blocks:
            h = blk(h, c)
        out_drop = m.unpatchify(m.final_layer(h, c))
        m.eval()
    d = {"sd." + k: v.numpy() for k, v in m.state_dict().items()}
    d.update(x=x.numpy(), t=t.numpy(), y=y.numpy(), out=out.numpy(), ycfg=ycfg.numpy(), out_cfg=out_cfg.numpy(),
             drop=drop.numpy(), out_drop=out_drop.numpy(), cfg_scale=np.array(1.5))
    d.update({"kw." + k: np.array(v) for k, v in kw.items()})
    np.savez_compressed(os.path.join(OUT, "dit_tiny.npz"), **d)
    print("dit_tiny", sum(v.size for v in d.values()))


def gen_seeded(name, lat, n, tag):
    """Real configurations: weights are reproduced from the seed (checksums pin them)."""
    m = build(name, input_size=lat, num_classes=1000)
    x, t, y = inputs(n, lat, seed=0)
    with torch.no_grad():
        out = m(x, t, y)
        d = dict(x=x.numpy(), t=t.numpy(), y=y.numpy(), out=out.numpy(), name=np.array(name), lat=np.array(lat))
        if n % 2 == 0:
            ycfg = torch.cat([y[: n // 2], torch.full((n // 2,), 1000)])
            d.update(ycfg=ycfg.numpy(), out_cfg=m.forward_with_cfg(x, t, ycfg, 4.0).numpy())
    d.update(checksums(m))
    d["nparams"] = np.array(sum(p.numel() for p in m.parameters()))
    np.savez_compressed(os.path.join(OUT, f"dit_{tag}.npz"), **d)
    print(tag, "params", int(d["nparams"]), "out std", float(out.std()))
    return m


def gen_diffusion_kat():
    d = {}
    g = torch.Generator().manual_seed(0)
    x = torch.randn(4, 4, 16, 16, generator=g)
    out8 = torch.randn(4, 8, 16, 16, generator=g) * 0.5
    noise = torch.randn(4, 4, 16, 16, generator=g)
    x0 = torch.randn(4, 4, 16, 16, generator=g).clamp(-1.2, 1.2)
    d.update(x=x.numpy(), out8=out8.numpy(), noise=noise.numpy(), x0=x0.numpy())
    orig = ref_gd.th.randn_like
    ref_gd.th.randn_like = lambda z: noise.clone()
    try:
        for tag, spec, kw, t in [
            ("lr250", "250", {}, [249, 0, 17, 100]),
            ("lr1000", "", {}, [0, 637, 999, 1]),
            ("fl250", "250", {"learn_sigma": False}, [249, 0, 17, 100]),
            ("fs250", "250", {"learn_sigma": False, "sigma_small": True}, [249, 0, 17, 100]),
            ("x0_250", "250", {"predict_xstart": True}, [249, 0, 17, 100]),
            ("cos100", "100", {"noise_schedule": "squaredcos_cap_v2"}, [99, 0, 50, 1]),
        ]:
            df = ref_diffusion.create_diffusion(spec, **kw)
            tt = torch.tensor(t)
            mo = out8 if kw.get("learn_sigma", True) else out8[:, :4].contiguous()
            stub = lambda *a, **k: mo  # noqa: E731
            for clip in (False, True):
                pm = df.p_mean_variance(stub, x, tt, clip_denoised=clip)
                ps = df.p_sample(stub, x, tt, clip_denoised=clip)
                dd = df.ddim_sample(stub, x, tt, clip_denoised=clip, eta=0.3)
                c = f"{tag}|clip{int(clip)}|"
                for k in ("mean", "variance", "log_variance", "pred_xstart"):
                    d[c + "pmv." + k] = pm[k].numpy()
                d[c + "p_sample"] = ps["sample"].numpy()
                d[c + "ddim"] = dd["sample"].numpy()
            d[tag + "|t"] = np.array(t)
            d[tag + "|q_sample"] = df.q_sample(x0, tt, noise=noise).numpy()
            if kw.get("learn_sigma", True) and not kw.get("predict_xstart"):
                tl = df.training_losses(stub, x0, tt, noise=noise)
                for k in ("loss", "mse", "vb"):
                    d[tag + "|tl." + k] = tl[k].numpy()
                # gradient of sum(loss * w) wrt the model output
                mo_g = mo.clone().requires_grad_(True)
                w = torch.tensor([0.25, 1.0, -0.5, 2.0])
                tl = df.training_losses(lambda *a, **k: mo_g, x0, tt, noise=noise)
                (tl["loss"] * w).sum().backward()
                d[tag + "|tl.grad"] = mo_g.grad.numpy()
                d[tag + "|tl.w"] = w.numpy()
    finally:
        ref_gd.th.randn_like = orig
    np.savez_compressed(os.path.join(OUT, "diffusion_kat.npz"), **d)
    print("diffusion_kat", len(d))


def cond_fn_fixture(x, t, **kw):
    """Stand-in for grad log p(y | x): exactly-rounded IEEE f32 ops only, so CPU and GPU evaluate it identically."""
    return x * 0.5 - 0.25 + (t.float() * 0.001).view(-1, 1, 1, 1)


API_CASES = [
    ("lr250", "250", {}, [249, 0, 17, 100]),
    ("lr1000", "", {}, [0, 637, 999, 1]),
    ("fl250", "250", {"learn_sigma": False}, [249, 0, 17, 100]),
    ("fs250", "250", {"learn_sigma": False, "sigma_small": True}, [249, 0, 17, 100]),
    ("x0_250", "250", {"predict_xstart": True}, [249, 0, 17, 100]),
    ("cos100", "100", {"noise_schedule": "squaredcos_cap_v2"}, [99, 0, 50, 1]),
]


def gen_diffusion_api():
    """The rest of GaussianDiffusion's surface: standalone q/p helpers, classifier guidance (cond_fn) through
    p_sample / ddim_sample, ddim_reverse_sample, the KL loss family, _vb_terms_bpd, _prior_bpd, calc_bpd_loop, and
    the MSE loss for fixed-variance / x0-predicting models."""
    d = {}
    g = torch.Generator().manual_seed(5)
    x = torch.randn(4, 4, 16, 16, generator=g)
    out8 = torch.randn(4, 8, 16, 16, generator=g) * 0.5
    noise = torch.randn(4, 4, 16, 16, generator=g)
    x0 = torch.randn(4, 4, 16, 16, generator=g).clamp(-1.2, 1.2)
    w = torch.tensor([0.25, 1.0, -0.5, 2.0])
    d.update(x=x.numpy(), out8=out8.numpy(), noise=noise.numpy(), x0=x0.numpy(), w=w.numpy())
    orig = ref_gd.th.randn_like
    ref_gd.th.randn_like = lambda z: noise.clone()
    try:
        for tag, spec, kw, t in API_CASES:
            df = ref_diffusion.create_diffusion(spec, **kw)
            tt = torch.tensor(t)
            mo = out8 if kw.get("learn_sigma", True) else out8[:, :4].contiguous()
            stub = lambda *a, **k: mo  # noqa: E731
            d[tag + "|t"] = np.array(t)
            for k, v in zip(("mean", "variance", "log_variance"), df.q_mean_variance(x0, tt)):
                d[f"{tag}|qmv.{k}"] = v.numpy()
            for k, v in zip(("mean", "variance", "log_variance"), df.q_posterior_mean_variance(x0, x, tt)):
                d[f"{tag}|qpost.{k}"] = v.numpy()
            d[tag + "|x0_from_eps"] = df._predict_xstart_from_eps(x, tt, out8[:, :4]).numpy()
            d[tag + "|eps_from_x0"] = df._predict_eps_from_xstart(x, tt, x0).numpy()
            for clip in (False, True):
                c = f"{tag}|clip{int(clip)}|"
                r = df.p_sample(stub, x, tt, clip_denoised=clip, cond_fn=cond_fn_fixture, model_kwargs={})
                d[c + "p_sample_cond"], d[c + "p_sample_cond.pred"] = r["sample"].numpy(), r["pred_xstart"].numpy()
                r = df.ddim_sample(stub, x, tt, clip_denoised=clip, cond_fn=cond_fn_fixture, model_kwargs={}, eta=0.3)
                d[c + "ddim_cond"], d[c + "ddim_cond.pred"] = r["sample"].numpy(), r["pred_xstart"].numpy()
                r = df.ddim_reverse_sample(stub, x, tt, clip_denoised=clip)
                d[c + "ddim_rev"], d[c + "ddim_rev.pred"] = r["sample"].numpy(), r["pred_xstart"].numpy()
                r = df.ddim_reverse_sample(stub, x, tt, clip_denoised=clip, cond_fn=cond_fn_fixture, model_kwargs={})
                d[c + "ddim_rev_cond"] = r["sample"].numpy()
                x_t = df.q_sample(x0, tt, noise=noise)
                r = df._vb_terms_bpd(stub, x0, x_t, tt, clip_denoised=clip)
                d[c + "vb.output"], d[c + "vb.pred"] = r["output"].numpy(), r["pred_xstart"].numpy()
            d[tag + "|prior_bpd"] = df._prior_bpd(x0).numpy()
            # MSE-family loss and gradient for every mean / variance type
            mo_g = mo.clone().requires_grad_(True)
            tl = df.training_losses(lambda *a, **k: mo_g, x0, tt, noise=noise)
            (tl["loss"] * w).sum().backward()
            for k in tl:
                d[f"{tag}|mse.{k}"] = tl[k].detach().numpy()
            d[tag + "|mse.grad"] = mo_g.grad.numpy()
            # KL family (GD:735-746): RESCALED_KL is what create_diffusion(use_kl=True) builds; KL by loss_type
            for lt in ("RESCALED_KL", "KL"):
                dk = ref_diffusion.create_diffusion(spec, use_kl=True, **kw)
                dk.loss_type = getattr(ref_gd.LossType, lt)
                mo_g = mo.clone().requires_grad_(True)
                tl = dk.training_losses(lambda *a, **k: mo_g, x0, tt, noise=noise)
                assert set(tl) == {"loss"}
                (tl["loss"] * w).sum().backward()
                d[f"{tag}|{lt}.loss"] = tl["loss"].detach().numpy()
                d[f"{tag}|{lt}.grad"] = mo_g.grad.numpy()
    finally:
        ref_gd.th.randn_like = orig
    # calc_bpd_loop over a 10-step process; every step's noise from its own seeded generator
    for tag, kw in [("bpd10", {}), ("bpd10_fl", {"learn_sigma": False})]:
        df = ref_diffusion.create_diffusion("10", **kw)
        mo = out8 if kw.get("learn_sigma", True) else out8[:, :4].contiguous()
        step = {"i": 0}

        def seeded(zl):
            gg = torch.Generator().manual_seed(500 + step["i"])
            step["i"] += 1
            return torch.randn(zl.shape, generator=gg)

        ref_gd.th.randn_like = seeded
        try:
            for clip in (True, False):
                step["i"] = 0
                r = df.calc_bpd_loop(lambda xx, ts, **k: mo * (1.0 + ts.float().view(-1, 1, 1, 1) * 0.001), x0,
                                     clip_denoised=clip, model_kwargs={})
                for k, v in r.items():
                    d[f"{tag}|clip{int(clip)}|{k}"] = v.numpy()
        finally:
            ref_gd.th.randn_like = orig
    np.savez_compressed(os.path.join(OUT, "diffusion_api.npz"), **d)
    print("diffusion_api", len(d))


def import_fork_models():
    """The fork's top-level models.py, UNMODIFIED, imported with empty stand-ins for the visualisation libraries it
    pulls in at module import (umap, cv2, matplotlib: SURVEY.md §0.1) and oracle/timm_standin for timm."""
    import importlib.util
    import types

    for name in ("umap", "cv2", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:  # noqa: BLE001
                sys.modules[name] = types.ModuleType(name)
    spec = importlib.util.spec_from_file_location("ref_fork_models", os.path.join(REF, "models.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def gen_fork_dino():
    """the fork's DINO cross-attention DiT (models.py:506-601, 624-754): 9-chunk adaLN,
    CrossAttention (LayerNorm on q and k, fused kv Linear without bias) in the 14th and 16th block, c = t only.
    forward(x, t, dino_feat, y) in eval mode (dropout off) on a 16-block model."""
    import contextlib
    import io

    FK = import_fork_models()
    for tag, kw, n in [("fork_small", dict(input_size=32, patch_size=2, hidden_size=384, depth=16, num_heads=6,
                                            dino_feat_size=64, num_classes=10), 3),
                       ("fork_p4", dict(input_size=32, patch_size=4, hidden_size=768, depth=16, num_heads=12,
                                         dino_feat_size=768, num_classes=10), 2)]:
        torch.manual_seed(0)
        with contextlib.redirect_stdout(io.StringIO()):  # the constructor prints per block
            m = FK.DiT(**kw)
        rerandomise_zero_params(m.named_parameters())
        m.eval()
        g = torch.Generator().manual_seed(21)
        lat = kw["input_size"]
        x = torch.randn(n, 4, lat, lat, generator=g)
        dino = torch.randn(n, kw["dino_feat_size"], lat, lat, generator=g)
        t = torch.randint(0, 1000, (n,), generator=g)
        y = torch.randint(0, 10, (n,), generator=g)
        with torch.no_grad(), contextlib.redirect_stdout(io.StringIO()):
            out = m(x, t, dino, y)
            out_nodino_blocks = None
        # dino_feat is large (n x 768 x 32 x 32): the fixture keeps the generator seed; tests replay the draws
        # (x first, then dino_feat) and check x against the stored copy
        d = dict(x=x.numpy(), gen_seed=np.array(21), t=t.numpy(), y=y.numpy(), out=out.numpy(),
                 dino_sum=np.array(float(dino.double().sum())))
        d.update({"kw." + k: np.array(v) for k, v in kw.items()})
        d.update(checksums(m))
        d["nparams"] = np.array(sum(p.numel() for p in m.parameters()))
        np.savez_compressed(os.path.join(OUT, f"dit_{tag}.npz"), **d)
        print(tag, "params", int(d["nparams"]), "out std", float(out.std()), "keys", len(m.state_dict()))


def gen_sample_loop(model):
    """BASELINE.json configs[0]: DiT-S/2, 10-step CFG-4.0 sampling (n=2 kept images, batch 4 with the
    null-class half), every step's noise drawn from its own seeded CPU generator."""
    df = ref_diffusion.create_diffusion("10")
    n = 2
    g = torch.Generator().manual_seed(7)
    z = torch.randn(n, 4, 32, 32, generator=g)
    y = torch.randint(0, 1000, (n,), generator=g)
    z = torch.cat([z, z], 0)
    yy = torch.cat([y, torch.full((n,), 1000)])
    step = {"i": 0}

    def seeded(zl):
        gg = torch.Generator().manual_seed(1000 + step["i"])
        step["i"] += 1
        return torch.randn(zl.shape, generator=gg)

    orig = ref_gd.th.randn_like
    ref_gd.th.randn_like = seeded
    try:
        traj = []
        with torch.no_grad():
            for out in df.p_sample_loop_progressive(model.forward_with_cfg, z.shape, z, clip_denoised=False,
                                                    model_kwargs=dict(y=yy, cfg_scale=4.0), device="cpu"):
                traj.append(out["sample"].numpy())
    finally:
        ref_gd.th.randn_like = orig
    np.savez_compressed(os.path.join(OUT, "sample_s2_10step.npz"), z=z.numpy(), y=yy.numpy(),
                        traj=np.stack(traj), noise_seed_base=np.array(1000), cfg_scale=np.array(4.0))
    print("sample loop: final std", float(traj[-1].std()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    if len(sys.argv) > 1 and sys.argv[1] == "api":  # only the fixtures added in round 2
        gen_diffusion_api()
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "fork":
        gen_fork_dino()
        sys.exit(0)
    gen_tables()
    gen_diffusion_api()
    gen_fork_dino()
    gen_tiny()
    gen_diffusion_kat()
    s2 = gen_seeded("DiT-S/2", 32, 4, "s2_seed0")
    gen_sample_loop(s2)
    del s2
    gen_seeded("DiT-B/4", 32, 4, "b4_seed0")
    gen_seeded("DiT-S/8", 32, 2, "s8_seed0")
    gen_seeded("DiT-XL/2", 32, 2, "xl2_seed0")
    gen_seeded("DiT-XL/2", 64, 1, "xl2_512_seed0")
