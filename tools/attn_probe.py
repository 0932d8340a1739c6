#!/usr/bin/env python
"""Checks ditb200_attention_fwd against an fp32 torch reference and times it.

    python tools/attn_probe.py --b 64 --t 256 --h 16 --hd 72 [--iters 50]
    DITB200_ATTN_MMA_SYNC=1 python tools/attn_probe.py ...      # the mma.sync kernel for comparison"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fast_dit_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--t", type=int, default=256)
    ap.add_argument("--h", type=int, default=16)
    ap.add_argument("--hd", type=int, default=72)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--scale", type=float, default=1.0)
    a = ap.parse_args()
    dev = torch.device("cuda")
    B, T, H, hd = a.b, a.t, a.h, a.hd
    D = H * hd
    g = torch.Generator(device=dev).manual_seed(0)
    qkvs = [(torch.randn(B * T, 3 * D, device=dev, generator=g) * a.scale).bfloat16() for _ in range(3)]
    qkv = qkvs[0]
    lse = torch.empty(B, H, T, device=dev)
    out = ops.attention(qkv, B, T, H, hd, lse=lse)
    torch.cuda.synchronize()
    q, k, v = (x.reshape(B, T, H, hd).permute(0, 2, 1, 3).float() for x in qkv.float().chunk(3, dim=1))
    s = (q @ k.transpose(-1, -2)) * hd ** -0.5
    ref = (torch.softmax(s, -1) @ v).permute(0, 2, 1, 3).reshape(B * T, D)
    ref_lse = torch.logsumexp(s, -1)
    err = float((out.float() - ref).norm() / ref.norm())
    err_lse = float((lse - ref_lse).abs().max())
    bad = int((~torch.isfinite(out.float())).sum())
    print(f"B={B} T={T} H={H} hd={hd}: rel-L2 {err:.3e}  max|lse err| {err_lse:.3e}  non-finite {bad}", flush=True)
    for i in range(5):
        ops.attention(qkvs[i % 3], B, T, H, hd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.iters):
        ops.attention(qkvs[i % 3], B, T, H, hd)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    fl = 4.0 * B * H * T * T * hd
    print(f"  {ms * 1e3:.1f} us/launch  {fl / ms / 1e9:.1f} TF  ({'mma.sync' if os.environ.get('DITB200_ATTN_MMA_SYNC') else 'tcgen05'})", flush=True)


if __name__ == "__main__":
    main()
