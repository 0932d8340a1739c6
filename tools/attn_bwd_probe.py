#!/usr/bin/env python
"""Checks ditb200_attention_bwd against fp64 torch autograd and times it.

    python tools/attn_bwd_probe.py --b 32 --t 256 --h 16 --hd 72
    DITB200_ATTN_MMA_SYNC=1 python tools/attn_bwd_probe.py ...     # mma.sync kernels for comparison"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from fast_dit_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=32)
    ap.add_argument("--t", type=int, default=256)
    ap.add_argument("--h", type=int, default=16)
    ap.add_argument("--hd", type=int, default=72)
    ap.add_argument("--iters", type=int, default=30)
    a = ap.parse_args()
    dev = torch.device("cuda")
    B, T, H, hd = a.b, a.t, a.h, a.hd
    D = H * hd
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B * T, 3 * D, device=dev, generator=g).bfloat16()
    dout = torch.randn(B * T, D, device=dev, generator=g).bfloat16()
    lse = torch.empty(B, H, T, device=dev)
    out = ops.attention(qkv, B, T, H, hd, lse=lse)
    got = ops.attention_bwd(qkv, out, dout, lse, B, T, H, hd)
    torch.cuda.synchronize()
    nb = min(B, 4)  # reference on a few images
    q = qkv[: nb * T].double().view(nb, T, 3, H, hd).permute(2, 0, 3, 1, 4).contiguous().requires_grad_(True)
    o = F.scaled_dot_product_attention(q[0], q[1], q[2]).transpose(1, 2).reshape(nb * T, D)
    o.backward(dout[: nb * T].double())
    ref = q.grad.permute(1, 3, 0, 2, 4).reshape(nb * T, 3 * D)
    errs = [float((got[: nb * T, j * D:(j + 1) * D].double() - ref[:, j * D:(j + 1) * D]).norm() / ref[:, j * D:(j + 1) * D].norm())
            for j in range(3)]
    print(f"B={B} T={T} H={H} hd={hd}: rel-L2 dq {errs[0]:.3e} dk {errs[1]:.3e} dv {errs[2]:.3e}  "
          f"non-finite {int((~torch.isfinite(got.float())).sum())}", flush=True)
    for _ in range(3):
        ops.attention_bwd(qkv, out, dout, lse, B, T, H, hd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        ops.attention_bwd(qkv, out, dout, lse, B, T, H, hd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    print(f"  {ms * 1e3:.1f} us/launch (incl. dsum)  {10.0 * B * H * T * T * hd / ms / 1e9:.1f} TF  "
          f"({'mma.sync' if os.environ.get('DITB200_ATTN_MMA_SYNC') else 'tcgen05'})", flush=True)


if __name__ == "__main__":
    main()
