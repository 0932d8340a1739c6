#!/usr/bin/env python
"""Times ditb200_ln_modulate at the C3 shape (M=16384, D=1152): DITB200_LN_THREADS=128|256|512 python tools/ln_probe.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fast_dit_b200 import ops
dev = torch.device("cuda")
M, D, T = int(os.environ.get("M", 16384)), 1152, 256
xs = [torch.randn(M, D, device=dev) for _ in range(4)]
mod = torch.randn(M // T, 6 * D, device=dev)
outs = [torch.empty(M, D, device=dev, dtype=torch.bfloat16) for _ in range(4)]
for i in range(8): ops.ln_modulate(xs[i % 4], mod[:, :D], mod[:, D:2 * D], T, out=outs[i % 4])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(100): ops.ln_modulate(xs[i % 4], mod[:, :D], mod[:, D:2 * D], T, out=outs[i % 4])
e1.record(); torch.cuda.synchronize()
us = e0.elapsed_time(e1) * 10
print(f"ln_modulate M={M} threads={os.environ.get('DITB200_LN_THREADS', '256')}: {us:.1f} us  {M * D * 6 / us / 1e6:.2f} TB/s")
