#!/bin/bash
TAG=${1:-cubic}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for v in 1 2; do
  RBX_FUSED_VARIANT=$v timeout -s KILL 200 python bench.py --particles 1000000 --method cubic --no-cpu --no-e2e > $OUT/bench_cubic_v${v}_1e6.json 2>> $OUT/bench.err
done
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_cubic*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f parity %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("ok")))
    except Exception as e: print(f, "ERR", e)
PY
