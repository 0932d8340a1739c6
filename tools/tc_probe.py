"""Bring-up probe for the tcgen05 GEMM: one configuration per process so a deadlocked
kernel can be killed by `timeout` without taking the rest of the run with it.

    python tools/tc_probe.py --cg 1 --bn 128 --m 256 --n 256 --k 64 [--bench]
"""
import argparse
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from fast_dit_b200 import ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cg", type=int, default=1)
    ap.add_argument("--bn", type=int, default=128)
    ap.add_argument("--m", type=int, default=256)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--k", type=int, default=64)
    ap.add_argument("--epi", type=int, default=0)
    ap.add_argument("--bench", action="store_true")
    ap.add_argument("--pattern", default="randn")
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(0)
    M, N, K = a.m, a.n, a.k
    if a.pattern == "randn":
        A = torch.randn(M, K, device=dev, generator=g).bfloat16()
        W = (torch.randn(N, K, device=dev, generator=g) / math.sqrt(K)).bfloat16()
    else:  # structured: A = row index marker, W = identity-like, to localise layout bugs
        A = (torch.arange(M, device=dev)[:, None] % 64 + torch.arange(K, device=dev)[None] % 7).bfloat16()
        W = torch.zeros(N, K, device=dev)
        W[torch.arange(N, device=dev), torch.arange(N, device=dev) % K] = 1.0
        W = W.bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    T = 64
    resid = torch.randn(M, N, device=dev, generator=g)
    gate = torch.randn((M + T - 1) // T, N, device=dev, generator=g)
    kw = {}
    if a.epi == 2:
        kw = dict(resid=resid.clone(), gate=gate, rows_per_gate=T)
    torch.cuda.synchronize()
    t0 = time.time()
    y = ops.gemm(A, W, bias, epilogue=a.epi, out_dtype=torch.float32, tile_n=a.bn, cta_group=a.cg, **kw)
    torch.cuda.synchronize()
    ref = A.double() @ W.double().t() + bias.double()
    if a.epi == 1:
        ref = torch.nn.functional.gelu(ref, approximate="tanh")
    elif a.epi == 3:
        ref = torch.nn.functional.silu(ref)
    elif a.epi == 2:
        ref = resid.double() + gate.double().repeat_interleave(T, 0)[:M] * ref
    err = (y.double() - ref)
    rel = float(err.norm() / ref.norm())
    print(f"cfg cg={a.cg} bn={a.bn} M={M} N={N} K={K} epi={a.epi}: rel_l2={rel:.3e} "
          f"max_abs={float(err.abs().max()):.3e} first-call {1e3*(time.time()-t0):.1f} ms", flush=True)
    if rel > 1e-3:
        bad = (err.abs() > 1e-2 * ref.abs().max())
        rb = bad.view(-1, N)[: (M // 32) * 32].view(M // 32, 32, N).any(1).float()
        print("bad fraction per 32-row block:", [round(float(v), 2) for v in rb.mean(1)[:16]])
        cbn = (N // 32) * 32
        cb = bad[:, :cbn].view(M, cbn // 32, 32).any(2).float()
        print("bad fraction per 32-col block:", [round(float(v), 2) for v in cb.mean(0)[:16]])
        print("y[0,:8] ", y[0, :8].tolist())
        print("ref[0,:8]", ref[0, :8].float().tolist())
    if a.bench and rel < 1e-2:
        yb = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
        for _ in range(3):
            ops.gemm(A, W, bias, epilogue=a.epi if a.epi != 2 else 0, out=yb, tile_n=a.bn, cta_group=a.cg)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        iters = 20
        ev0.record()
        for _ in range(iters):
            ops.gemm(A, W, bias, epilogue=a.epi if a.epi != 2 else 0, out=yb, tile_n=a.bn, cta_group=a.cg)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / iters
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        # cuBLAS for context
        for _ in range(3):
            torch.nn.functional.linear(A, W)
        ev0.record()
        for _ in range(iters):
            torch.nn.functional.linear(A, W)
        ev1.record()
        torch.cuda.synchronize()
        ms2 = ev0.elapsed_time(ev1) / iters
        tf2 = 2.0 * M * N * K / (ms2 * 1e-3) / 1e12
        print(f"  bench: {ms*1e3:.1f} us  {tf:.1f} TFLOP/s   (cuBLAS {ms2*1e3:.1f} us {tf2:.1f} TFLOP/s)", flush=True)


if __name__ == "__main__":
    main()
