#!/bin/bash
TAG=${1:-v2prof}
OUT=gpurun_out/$TAG
mkdir -p $OUT
export RBX_FUSED_VARIANT=2
for w in 6 5 4; do
  RBX_FUSED_WARPS=$w timeout -s KILL 200 python bench.py --particles 1000000 --no-cpu --no-e2e --no-parity > $OUT/bench_v2_w${w}_1e6.json 2>> $OUT/bench.err
  RBX_FUSED_WARPS=$w timeout -s KILL 200 python bench.py --particles 10000000 --no-cpu --no-e2e --no-parity > $OUT/bench_v2_w${w}_1e7.json 2>> $OUT/bench.err
done
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:fused_cube_warp -s 3 -c 1 -o $OUT/prof_v2_linear_1000000 -f python bench.py --particles 1000000 --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/ncu.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_v*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"]))
    except Exception as e: print(f, "ERR", e)
PY
