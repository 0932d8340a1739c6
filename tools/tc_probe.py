#!/usr/bin/env python
"""Times ditb200_gemm (tcgen05 engine) for one shape over a list of tile configs, next to cuBLAS.

    python tools/tc_probe.py --m 16384 --n 1152 --k 1152 --cfgs 2x256,2x144,2x128 [--epi 2] [--check]

Rotates over enough operand sets to exceed the 126 MB L2, CUDA events on the launching stream."""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fast_dit_b200 import _lib as L, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--m", type=int, default=16384)
    ap.add_argument("--n", type=int, default=1152)
    ap.add_argument("--k", type=int, default=1152)
    ap.add_argument("--cfgs", default="0x0")
    ap.add_argument("--epi", type=int, default=0, help="0 bias, 1 gelu, 2 gate+resid, 4 mul dgelu (bf16 aux_in)")
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--sets", type=int, default=4)
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--trans-w", action="store_true")
    ap.add_argument("--trans-a", action="store_true")
    ap.add_argument("--split-k", type=int, default=0)
    ap.add_argument("--no-cublas", action="store_true")
    ap.add_argument("--inplace", action="store_true", help="epi 2: update the residual stream in place (as the model does)")
    a = ap.parse_args()
    dev = torch.device("cuda")
    M, N, K = a.m, a.n, a.k
    g = torch.Generator(device=dev).manual_seed(0)
    T = 256
    sets = []
    for _ in range(a.sets):
        A = (torch.randn((K, M) if a.trans_a else (M, K), device=dev, generator=g) * 0.5).bfloat16()
        W = (torch.randn((K, N) if a.trans_w else (N, K), device=dev, generator=g) * 0.05).bfloat16()
        bias = torch.randn(N, device=dev, generator=g)
        kw = dict(bias=bias)
        if a.epi == 2:
            res = torch.randn(M, N, device=dev, generator=g)
            kw.update(epilogue=L.EPI_BIAS_GATE_RESID, resid=res,
                      gate=torch.randn(max(1, M // T), N, device=dev, generator=g) * 0.1, rows_per_gate=T,
                      out=res if a.inplace else torch.empty(M, N, device=dev))
        elif a.epi == 1:
            kw.update(epilogue=L.EPI_BIAS_GELU, out=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
        elif a.epi == 4:
            kw.update(epilogue=L.EPI_MUL_DGELU, aux_in=torch.randn(M, N, device=dev, generator=g).bfloat16(),
                      out=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
        elif a.split_k > 1:
            kw.update(out=torch.empty(M, N, device=dev), out_dtype=torch.float32, split_k=a.split_k)
        else:
            kw.update(out=torch.empty(M, N, device=dev, dtype=torch.bfloat16))
        sets.append((A, W, kw))
    flops = 2.0 * M * N * K

    def run(cg, bn, i):
        A, W, kw = sets[i % len(sets)]
        return ops.gemm(A, W, cta_group=cg, tile_n=bn, trans_a=a.trans_a, trans_w=a.trans_w, **kw)

    def time_it(fn):
        for i in range(5):
            fn(i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.iters):
            fn(i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.iters

    tag = f"M={M} N={N} K={K} epi={a.epi}" + (" tA" if a.trans_a else "") + (" tW" if a.trans_w else "")
    for cfg in a.cfgs.split(","):
        cg, bn = (int(v) for v in cfg.split("x"))
        try:
            if a.check:
                A, W, kw = sets[0]
                kw2 = dict(kw)
                if a.epi == 2:
                    res0 = kw["resid"].clone()
                    kw2["out"] = kw["resid"] if a.inplace else torch.empty(M, N, device=dev)
                got = ops.gemm(A, W, cta_group=cg, tile_n=bn, trans_a=a.trans_a, trans_w=a.trans_w, **kw2).float()
                Af = A.float().t() if a.trans_a else A.float()
                Wf = W.float() if a.trans_w else W.float().t()
                ref = Af @ Wf + kw["bias"]
                if a.epi == 1:
                    ref = torch.nn.functional.gelu(ref, approximate="tanh")
                if a.epi == 4:
                    u = kw["aux_in"].float().requires_grad_(True)
                    torch.nn.functional.gelu(u, approximate="tanh").sum().backward()
                    ref = ref * u.grad
                if a.epi == 2:
                    ref = res0 + kw["gate"].repeat_interleave(T, 0)[:M] * ref
                err = float((got - ref).norm() / ref.norm())
                print(f"  check cg={cg} bn={bn}: rel-L2 {err:.2e}")
            ms = time_it(lambda i: run(cg, bn, i))
            print(f"{tag} cg={cg} bn={bn:3d}: {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TF", flush=True)
        except Exception as e:  # noqa: BLE001
            print(f"{tag} cg={cg} bn={bn}: FAILED {e}", flush=True)
    if not a.no_cublas and not a.trans_a and not a.trans_w:
        outs = [torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in sets]
        ms = time_it(lambda i: torch.matmul(sets[i % len(sets)][0], sets[i % len(sets)][1].t(), out=outs[i % len(sets)]))
        print(f"{tag} cuBLAS      : {ms * 1e3:8.1f} us  {flops / ms / 1e9:7.1f} TF", flush=True)


if __name__ == "__main__":
    main()
