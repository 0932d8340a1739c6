#!/usr/bin/env python
"""Per-shape breakdown of the GEMMs (and every other kernel) of one training step.

    python tools/train_profile.py [--workload c4]"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fast_dit_b200 import DiT_models, create_diffusion, ops  # noqa: E402
from fast_dit_b200.optim import FusedAdamWEMA  # noqa: E402
from fast_dit_b200.utils import rerandomise_zero_params  # noqa: E402

W = {"c2": ("DiT-B/4", 32, 256), "c4": ("DiT-XL/2", 32, 32)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="c4")
    a = ap.parse_args()
    name, lat, n = W[a.workload]
    dev = torch.device("cuda")
    torch.manual_seed(0)
    model = DiT_models[name](input_size=lat, num_classes=1000, precision="bf16")
    rerandomise_zero_params(model)
    model = model.to(dev).train()
    opt = FusedAdamWEMA(model, lr=1e-4, weight_decay=0.0, ema_decay=0.9999)
    diffusion = create_diffusion("")
    x = torch.randn(n, 4, lat, lat, device=dev)
    y = torch.randint(0, 1000, (n,), device=dev)

    def step():
        t = torch.randint(0, diffusion.num_timesteps, (n,), device=dev)
        loss = diffusion.training_losses(model, x, t, dict(y=y))["loss"].mean()
        loss.backward()
        opt.step()
        opt.zero_grad()

    for _ in range(3):
        step()
    reps = 3
    with ops.profile() as prof:
        for _ in range(reps):
            step()
    rows = sorted(prof.summary_by_tag().items(), key=lambda kv: -kv[1][1])
    total = sum(v[1] for _, v in rows) / reps
    print(f"{name} {n} img/GPU: {total:.2f} ms of kernels per step")
    for (kname, tag), (cnt, ms, fl) in rows[:45]:
        tf = f"{fl / ms / 1e9:7.0f} TF" if fl else ""
        print(f"  {ms / reps:8.3f} ms  {cnt // reps:4d}x  {ms / cnt * 1e3:8.1f} us  {kname:18s} {tag or '':50s} {tf}")


if __name__ == "__main__":
    main()
