"""Shared helpers for the test-suite (fixtures loading, the weight protocol)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def rel_l2(a, b):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def build_product_model(name=None, seed=0, device=None, precision="bf16", **kw):
    """SURVEY.md §8c protocol on the product module: seed -> construct -> re-randomise the
    all-zero parameters -> eval."""
    from fast_dit_b200.models import DiT, DiT_models
    from oracle.dit_oracle import rerandomise_zero_params

    torch.manual_seed(seed)
    m = DiT_models[name](precision=precision, **kw) if name else DiT(precision=precision, **kw)
    rerandomise_zero_params(m.named_parameters())
    m.eval()
    if device is not None:
        m = m.to(device)
    return m


def check_checksums(model, fx, rtol=0.0):
    sd = model.state_dict()
    keys = [k[3:] for k in fx.files if k.startswith("ck.")]
    assert sorted(keys) == sorted(sd.keys())
    for k in keys:
        v = sd[k].double().cpu()
        got = np.array([float(v.sum()), float(v.abs().sum())])
        assert np.allclose(got, fx["ck." + k], rtol=1e-12, atol=1e-9), k
