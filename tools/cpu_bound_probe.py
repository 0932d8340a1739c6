#!/usr/bin/env python
"""Is the training step CPU-bound?  Compares the host time to ENQUEUE one step with the device time to run it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_dit_b200 import DiT_models, create_diffusion, ops
from fast_dit_b200.optim import FusedAdamWEMA
from fast_dit_b200.utils import rerandomise_zero_params

name, lat, n = ("DiT-XL/2", 32, 32) if len(sys.argv) < 2 or sys.argv[1] == "c4" else ("DiT-B/4", 32, 256)
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
torch.cuda.synchronize()
enq, tot = [], []
for _ in range(5):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    l0 = ops.LAUNCHES
    step()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    enq.append((t1 - t0) * 1e3); tot.append((t2 - t0) * 1e3)
print(f"{name}: enqueue {sum(enq)/5:.2f} ms/step (host), complete {sum(tot)/5:.2f} ms/step, {ops.LAUNCHES - l0} library launches/step")
