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
