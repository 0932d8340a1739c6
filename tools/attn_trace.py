#!/usr/bin/env python
"""Timeline of attn_fwd_tc_kernel's hand-offs (needs a library built with -DDITB200_ATTN_TRACE:
`DITB200_NVCC_EXTRA=-DDITB200_ATTN_TRACE tools/ab.sh build trace`, then on the GPU box
`cp ab/trace.so fast_dit_b200/lib/libditb200.so; python tools/attn_trace.py`).

CTA 0 and CTA 100 record clock64() for their first 8 work items; printed in SM clocks relative to the CTA's first
event.  Columns per softmax group g (query tile g): S = s_full seen, M = max pass done, P = p_full arrived (exp pass
done; X = its turn on the MUFU pipe granted, token builds), O = o_full seen, F = s_free arrived (O in registers), E = output stored.  MMA warp per tile: p = p_full seen,
o = P.V issued, w = next-S waits done, s = next S issued.  Producer: k = kv_empty seen, l = loads issued."""
import argparse
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from fast_dit_b200 import _lib, ops  # noqa: E402

SLOTS = {"S0": 0, "M0": 1, "P0": 2, "O0": 3, "F0": 4, "E0": 5, "S1": 6, "M1": 7, "P1": 8, "O1": 9, "F1": 10, "E1": 11,
         "p0": 12, "o0": 13, "w0": 14, "s0": 15, "p1": 16, "o1": 17, "w1": 18, "s1": 19, "k": 20, "l": 21,
         "X0": 23, "X1": 24}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--b", type=int, default=64)
    ap.add_argument("--t", type=int, default=256)
    ap.add_argument("--h", type=int, default=16)
    ap.add_argument("--hd", type=int, default=72)
    a = ap.parse_args()
    dev = torch.device("cuda")
    B, T, H, hd = a.b, a.t, a.h, a.hd
    g = torch.Generator(device=dev).manual_seed(0)
    qkv = torch.randn(B * T, 3 * H * hd, device=dev, generator=g).bfloat16()
    for _ in range(3):
        ops.attention(qkv, B, T, H, hd)
    torch.cuda.synchronize()
    lib = _lib.load()
    fn = lib.ditb200_attn_trace_read  # AttributeError: not a trace build
    fn.restype, fn.argtypes = C.c_int, [C.c_void_p, C.c_int]
    n = 2 * 8 * 32
    buf = (C.c_ulonglong * n)()
    rc = fn(buf, n)
    assert rc == 0, rc
    for cta in range(2):
        rows = [[buf[(cta * 8 + it) * 32 + s] for s in range(32)] for it in range(8)]
        t0 = min(v for r in rows for v in r if v)
        print(f"== CTA {'0' if cta == 0 else '100'} (clocks since its first event)")
        print("it " + " ".join(f"{k:>6}" for k in SLOTS))
        for it, r in enumerate(rows):
            print(f"{it:2d} " + " ".join(f"{(r[s] - t0) if r[s] else -1:6d}" for s in SLOTS.values()))
        # durations per group and item
        print("per group: s_full wait | max pass | exp pass | o_full wait | O read | output | period")
        for gi in range(2):
            o = 6 * gi
            for it in range(1, 8):
                r, pr = rows[it], rows[it - 1]
                if not r[o + 5] or not pr[o + 5]:
                    continue
                print(f"  g{gi} it{it}: {r[o] - pr[o + 5]:6d} {r[o + 1] - r[o]:6d} {r[o + 2] - r[o + 1]:6d} "
                      f"{r[o + 3] - r[o + 2]:6d} {r[o + 4] - r[o + 3]:6d} {r[o + 5] - r[o + 4]:6d} | {r[o + 2] - pr[o + 2]:6d}")
        print("MMA warp per tile: p_full seen -> P.V issued | -> waits done | -> S issued;  o_full latency = O(group) - p(mma)")
        for it in range(0, 7):
            r = rows[it]
            for t in range(2):
                m = 12 + 4 * t
                if not r[m + 3]:
                    continue
                print(f"  it{it} t{t}: issue_o {r[m + 1] - r[m]:6d}  waits {r[m + 2] - r[m + 1]:6d}  issue_s {r[m + 3] - r[m + 2]:6d}"
                      f"   p_full arrive->seen {r[m] - r[6 * t + 2]:6d}  PV issued->o_full seen {r[6 * t + 3] - r[m + 1]:6d}"
                      f"  s_free->S issued {r[m + 3] - r[6 * t + 4]:6d}  S issued->s_full seen(next) "
                      f"{(rows[it + 1][6 * t] - r[m + 3]) if rows[it + 1][6 * t] else -1:6d}")


if __name__ == "__main__":
    main()
