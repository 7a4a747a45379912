#!/usr/bin/env python
"""SASS evidence per kernel of librubix_b200.so (cuobjdump -sass, run in the build container): instruction
count and the mnemonics that show how each kernel moves data -- 128-bit global loads / stores, TMA bulk copies
(UBLKCP) and mbarrier waits (SYNCS), packed FFMA2, shared-memory traffic, atomics / reductions, barriers.

    python tools/sass_summary.py > profiles/r01_sass_v11.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rubix_b200", "librubix_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
demangle = lambda n: subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
kern, name = collections.OrderedDict(), None
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        name = m.group(1)
        kern[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and name:
        kern[name][m.group(1)] += 1
KEYS = [("LDG.128", r"^LDG\.E\.128"), ("LDG.64", r"^LDG\.E\.64"), ("LDG (<= 32 bit)", r"^LDG(?!\.E\.(128|64))"),
        ("STG.128", r"^STG\.E\.128"), ("STG (<= 64 bit)", r"^STG(?!\.E\.128)"), ("UBLKCP (TMA bulk)", r"^UBLKCP"), ("SYNCS (mbarrier)", r"^SYNCS"), ("LDS", r"^LDS"),
        ("STS", r"^STS"), ("FFMA2", r"^FFMA2"), ("FFMA", r"^FFMA$|^FFMA\."), ("MUFU", r"^MUFU"),
        ("RED (no-return atomics)", r"^RED"), ("ATOM / ATOMG / ATOMS", r"^ATOM"), ("BAR", r"^BAR"), ("SHFL", r"^SHFL")]
print(f"# {os.path.basename(so)}: SASS mnemonic counts per kernel (static instruction counts, sm_100a)\n")
for n, c in kern.items():
    d = demangle(n)
    if "rbx::" not in d:
        continue
    short = re.sub(r"\(.*", "", d).replace("void ", "")
    tot = sum(c.values())
    parts = []
    for label, pat in KEYS:
        v = sum(k for op, k in c.items() if re.match(pat, op))
        if v:
            parts.append(f"{label} {v}")
    print(f"{short}: {tot} instructions; " + ", ".join(parts))
