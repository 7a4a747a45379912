#!/usr/bin/env python
"""Aggregate an ncu capture per CUDA source line (run in the build container, no GPU needed).

    python tools/ncu_lines.py gpurun_out/prof.ncu-rep fused_cube_kernelILi0 [top]

ncu's ``--page source --csv`` gives per-SASS-instruction samples / executed counts; ``nvdisasm -g``
of the matching cubin (extracted from rubix_b200/librubix_b200.so, which must be the binary that was
profiled) gives the source line of every SASS offset.  Joining the two on the instruction offset
yields "which source lines cost what".
"""

import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def line_map(symbol_substr):
    so = os.path.join(ROOT, "rubix_b200", "librubix_b200.so")
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", so], cwd=tmp, capture_output=True)
    for f in os.listdir(tmp):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, f)], capture_output=True, text=True).stdout
        if symbol_substr not in txt:
            continue
        out, cur, active = {}, None, False
        for ln in txt.splitlines():
            if ln.startswith("//---") and ".text." in ln:
                active = symbol_substr in ln
            if not active:
                continue
            m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
            m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
            if m:
                out[int(m.group(1), 16)] = (cur, m.group(2).strip())
        if out:
            return out
    raise SystemExit(f"symbol {symbol_substr} not found in {so}")


def main():
    rep, sym = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    lm = line_map(sym)
    kern = re.sub(r"^\d+", "", re.sub(r"ILi\d+.*", "", sym).replace("_ZN3rbx", ""))
    res = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-id", f"::regex:{kern}:1"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(res)))
    h = next(i for i, r in enumerate(rows) if "Address" in r and "# Samples" in r)
    hdr = rows[h]
    ia, isamp, iex = hdr.index("Address"), hdr.index("# Samples"), hdr.index("Instructions Executed")
    base = int(rows[h + 1][ia], 16)
    per_line, sass = {}, {}
    ts = ti = 0
    for r in rows[h + 1:]:
        try:
            off, s, n = int(r[ia], 16) - base, int(r[isamp]), int(r[iex])
        except (ValueError, IndexError):
            continue
        loc, text = lm.get(off, (None, "?"))
        a = per_line.setdefault(loc, [0, 0])
        a[0] += s
        a[1] += n
        ts += s
        ti += n
        sass.setdefault(loc, []).append((n, s, text))
    src_cache = {}

    def src(loc):
        if not loc:
            return ""
        path = os.path.join(ROOT, "rubix_b200", "csrc", loc[0])
        if path not in src_cache:
            src_cache[path] = open(path).read().splitlines() if os.path.exists(path) else []
        L = src_cache[path]
        return L[loc[1] - 1].strip()[:100] if 0 < loc[1] <= len(L) else ""

    print(f"total stall samples {ts}, warp instructions executed {ti}")
    print("-- by stall samples --")
    for loc, (s, n) in sorted(per_line.items(), key=lambda kv: -kv[1][0])[:top]:
        print(f"{100 * s / ts:5.1f}% smp {100 * n / ti:5.1f}% inst  {loc}  {src(loc)}")
    print("-- by instructions executed --")
    for loc, (s, n) in sorted(per_line.items(), key=lambda kv: -kv[1][1])[:top]:
        print(f"{100 * n / ti:5.1f}% inst {100 * s / ts:5.1f}% smp  {loc}  {src(loc)}")


if __name__ == "__main__":
    main()
