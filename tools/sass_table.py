#!/usr/bin/env python
"""Per-kernel SASS mnemonic table of libditb200.so: which kernels really contain tcgen05 / TMEM / TMA instructions.

    python tools/sass_table.py > profiles/rNN_sass_mnemonics.md

UTCHMMA = tcgen05.mma (kind::f16), LDTM / STTM = tcgen05.ld / tcgen05.st (tensor memory), UTMALDG / UTMASTG = TMA
tensor load / store (cp.async.bulk.tensor), UTMAPF = tensormap prefetch, SYNCS = mbarrier ops, HMMA = mma.sync
(legacy tensor-core path), MUFU = special-function unit, ACQBULK / PREEXIT = griddepcontrol.wait /
.launch_dependents (programmatic dependent launch), UCGABAR = cluster barrier."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "fast_dit_b200", "lib", "libditb200.so")
COLS = ["UTCHMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "SYNCS", "HMMA", "MUFU", "ACQBULK", "PREEXIT", "UCGABAR"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["cu++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    counts, order, cur, i = {}, [], None, 0
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = names[i] if i < len(names) else m.group(1)
            i += 1
            cur = cur.replace("void ", "").replace("ditb200::", "").replace("(int)", "").replace("(bool)", "")
            cur = cur[:cur.index(">(") + 1] if ">(" in cur else cur.split("(")[0]
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        if cur is None:
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for c in COLS:
                if op.startswith(c):
                    counts[cur][c] += 1
            counts[cur]["_all"] += 1
    print("# SASS mnemonics per kernel — `cuobjdump -sass fast_dit_b200/lib/libditb200.so` (sm_100a)\n")
    print(__doc__.split("\n\n", 2)[2].strip() + "\n")
    print("| kernel | instrs | " + " | ".join(COLS) + " |")
    print("|---|---|" + "---|" * len(COLS))
    tot = collections.Counter()
    for k in sorted(order, key=lambda k: -counts[k]["UTCHMMA"] * 10000 - counts[k]["_all"]):
        c = counts[k]
        tot.update(c)
        print(f"| `{k}` | {c['_all']} | " + " | ".join(str(c[x]) if c[x] else "" for x in COLS) + " |")
    print(f"| **total ({len(order)} kernels)** | {tot['_all']} | " + " | ".join(str(tot[x]) for x in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
