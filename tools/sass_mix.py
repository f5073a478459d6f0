#!/usr/bin/env python
"""Instruction mix of the shipped library: `cuobjdump -sass libsfb200.so` -> per-kernel counts of the mnemonics that tell a
Blackwell-native kernel from a recompiled one (UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA load,
UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, HMMA = legacy mma.sync).   usage: python tools/sass_mix.py > profiles/<name>.md"""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "sample_factory_b200", "libsfb200.so")
COLS = ["UTCHMMA", "UTMALDG", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "F2FP", "LDS", "LD.E", "LDG", "STG", "ATOM|RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    for m, d in zip(names, out):
        demangle[m] = re.sub(r"\(.*$", "", d).replace("void ", "")
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, cur = OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle.get(m.group(1), m.group(1))
            counts[cur] = {c: 0 for c in COLS}
            continue
        if cur is None or "/*" not in line:
            continue
        mm = re.search(r"\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not mm:
            continue
        op = mm.group(1)
        for c in COLS:
            if any(op == alt or op.startswith(alt + ".") or (alt == "LD.E" and op.startswith("LD.E")) for alt in c.split("|")):
                counts[cur][c] += 1
    print(f"# SASS instruction mix of libsfb200.so (`tools/sass_mix.py`; cubins: {', '.join(archs)})\n")
    print("tcgen05.mma = `UTCHMMA`, tcgen05.ld / st = `LDTM` / `STTM`, TMA = `UTMALDG`, tcgen05.commit / mbarrier = `UTCBAR` / `SYNCS`; "
          "`HMMA` (legacy mma.sync) does not occur anywhere in the library.  `F2FP` = the fp16 operand split (cvt.rn.f16x2.f32); `LD.E` = "
          "generic loads (shared-memory accesses show up as `LDS`: the address-space fix of round 2).\n")
    print("| kernel | " + " | ".join(COLS) + " |")
    print("|---|" + "---:|" * len(COLS))
    tot = {c: 0 for c in COLS}
    for k, v in counts.items():
        for c in COLS:
            tot[c] += v[c]
        if v["UTCHMMA"] or v["UTMALDG"] or v["LDTM"]:
            print(f"| `{k}` | " + " | ".join(str(v[c]) for c in COLS) + " |")
    print("| **library total** | " + " | ".join(str(tot[c]) for c in COLS) + " |")


if __name__ == "__main__":
    sys.exit(main())
