#!/bin/bash
# cubic cube kernel: one warp per array (variant 1) against the pair variants
TAG=${1:-cubicv}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="--steps 10 --warmup 3 --no-cpu --no-e2e --no-stage --method cubic"
for v in 1 2 4 3; do
  RBX_FUSED_VARIANT=$v timeout -s KILL 200 python bench.py $B --particles 1000000 > $OUT/bench_1e6_cubic_v$v.json 2>> $OUT/bench.err
  RBX_FUSED_VARIANT=$v timeout -s KILL 200 python bench.py $B > $OUT/bench_1e7_cubic_v$v.json 2>> $OUT/bench.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f parity %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity", {}).get("ok")))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $OUT/bench.err
