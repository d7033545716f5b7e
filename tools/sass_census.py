#!/usr/bin/env python
"""Opcode census of libplc.so per kernel (cuobjdump -sass): proves which kernels are tcgen05 / TMA / TMEM code.

    python tools/sass_census.py [pl-convlstm-gan_b200/libplc.so] > profiles/r02_sass_census.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "SYNCS", "HMMA", "MUFU.TANH", "MUFU.EX2",
        "REDG", "ATOMG", "LDG", "STG", "LDS", "STS"]


def main():
    lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "pl-convlstm-gan_b200", "libplc.so")
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            base = op.split(".")[0]
            cur[base] += 1
            if op.startswith("UTCHMMA") and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
            if op.startswith("MUFU.TANH"):
                cur["MUFU.TANH"] += 1
            if op.startswith("MUFU.EX2"):
                cur["MUFU.EX2"] += 1
    print(f"# SASS opcode census of `{os.path.relpath(lib, ROOT)}` (cuobjdump -sass, sm_100a)\n")
    print("UTCHMMA = tcgen05.mma (`.2CTA` = cta_group::2), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld,")
    print("UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync (must be 0 everywhere).\n")
    print("| kernel | instr | " + " | ".join(KEYS) + " |")
    print("|---|---:|" + "---:|" * len(KEYS))
    tot = collections.Counter()
    for name, c in sorted(per.items(), key=lambda kv: -kv[1]["_total"]):
        d = demangle(name)
        d = d.replace("(int)", "").replace("void ", "")
        d = d[:d.index(">(") + 1] if ">(" in d else re.sub(r"\(.*", "", d)
        print(f"| `{d[-90:]}` | {c['_total']} | " + " | ".join(str(c[k]) for k in KEYS) + " |")
        tot.update(c)
    print(f"| **all {len(per)} kernels** | {tot['_total']} | " + " | ".join(str(tot[k]) for k in KEYS) + " |")


if __name__ == "__main__":
    main()
