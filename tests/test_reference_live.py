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
