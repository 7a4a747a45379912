#!/usr/bin/env python
"""Summarise an ncu --set full capture (run in the build container): key metrics, stall reasons,
opcode mix and the heaviest shared-memory instructions of the first profiled kernel.

    python tools/ncu_summary.py gpurun_out/prof_fused.ncu-rep [n_particles]
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
npart = float(sys.argv[2]) if len(sys.argv) > 2 else None
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
want = ["gpu__time_duration.sum", "sm__cycles_elapsed.avg", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "smsp__thread_inst_executed_per_inst_executed.ratio"]
print("kernel:", m.get("Kernel Name", ("", ""))[1])
for k in want:
    if k in m:
        print(f"{k:75s} {m[k][0]:16s} {m[k][1]}")
print("-- stall reasons (warps per issue-active cycle) --")
st = [(h, float(v)) for h, v in zip(hdr, vals) if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
for h, v in sorted(st, key=lambda x: -x[1])[:8]:
    print(f"  {h.split('issue_stalled_')[1].replace('_per_issue_active.ratio', ''):25s} {v:.3f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2 = rows[1]
ix = {h: i for i, h in enumerate(h2)}
data = rows[2:]


def f(r, k):
    try:
        return float(r[ix[k]])
    except Exception:
        return 0.0


ops = collections.Counter()
for r in data:
    s = r[ix["Source"]].strip()
    mm = re.match(r"(@!?U?P\d+\s+)?([A-Z0-9_.]+)", s)
    ops[(mm.group(2) if mm else s).split(".")[0]] += f(r, "Instructions Executed")
tot = sum(ops.values())
print(f"-- opcode mix: {tot / 1e6:.1f} M warp instructions" + (f", {tot / npart:.0f} per particle" if npart else "") + " --")
for k, v in ops.most_common(18):
    print(f"  {k:10s} {v / 1e6:8.1f} M {v / tot * 100:5.1f} %")
sh = sorted((r for r in data if f(r, "L1 Wavefronts Shared") > 0), key=lambda r: -f(r, "L1 Wavefronts Shared"))
ws, wi = sum(f(r, "L1 Wavefronts Shared") for r in sh), sum(f(r, "L1 Wavefronts Shared Ideal") for r in sh)
print(f"-- shared-memory wavefronts: {ws / 1e6:.1f} M, ideal {wi / 1e6:.1f} M --")
for r in sh[:12]:
    print(f"  {r[ix['Source']].strip()[:40]:40s} exec {f(r, 'Instructions Executed') / 1e6:6.2f} M  wavefronts {f(r, 'L1 Wavefronts Shared') / 1e6:7.2f} M  ideal {f(r, 'L1 Wavefronts Shared Ideal') / 1e6:6.2f} M")

# ---- optional: merge the per-particle counts into profiles/fused_ncu.json (bench.py's roofline reads them) ---------
#   python tools/ncu_summary.py REP N_PARTICLES KEY [TXT_NAME]     e.g. KEY = linear_10000000  (method_particles[_spaxels])
if len(sys.argv) > 3 and npart:
    import json
    import os
    # FP32 flops per warp instruction and lane (arithmetic only: compares, selects, min/max and conversions count 0)
    FLOPS = {"FFMA": 2, "FFMA2": 4, "FADD": 1, "FMUL": 1, "FADD2": 2, "FMUL2": 2, "MUFU": 1}
    lanes = float(m.get("smsp__thread_inst_executed_per_inst_executed.ratio", ("", "32"))[1])
    _bytes = lambda k: float(m[k][1]) * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[m[k][0]]
    flops = sum(v * FLOPS.get(k, 0) for k, v in ops.items()) * lanes
    rec = {
        "warp_inst_per_particle": tot / npart,
        "flop_per_particle": flops / npart,
        "fp32_warp_inst_per_particle": sum(v for k, v in ops.items() if k in FLOPS) / npart,
        "dram_bytes_per_launch": int(_bytes("dram__bytes_read.sum") + _bytes("dram__bytes_write.sum")),
        "kernel_ms_under_ncu": float(m["gpu__time_duration.sum"][1]) * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[m["gpu__time_duration.sum"][0]],
        "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][1]),
        "lsu_wavefront_pipe_pct": float(m["l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"][1]),
        "smem_wavefronts_per_particle": ws / npart, "smem_wavefronts_ideal_per_particle": wi / npart,
        "registers": int(m["launch__registers_per_thread"][1]), "block_size": int(m["launch__block_size"][1]),
        "kernel": m.get("Kernel Name", ("", ""))[1],
        "flop_rule": "per lane: FFMA 2, FFMA2 4, FADD / FMUL / MUFU 1, FADD2 / FMUL2 2; everything else 0",
        "capture": (sys.argv[4] if len(sys.argv) > 4 else os.path.basename(rep)) + " (ncu --set full --clock-control none)",
    }
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "fused_ncu.json")
    try:
        table = json.load(open(path))
    except (OSError, ValueError):
        table = {}
    table[sys.argv[3]] = rec
    json.dump(table, open(path, "w"), indent=1)
    print(f"-- profiles/fused_ncu.json[{sys.argv[3]}]: {rec['warp_inst_per_particle']:.1f} warp inst, "
          f"{rec['flop_per_particle']:.0f} flop per particle")
