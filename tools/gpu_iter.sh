#!/bin/bash
# One short box visit for a cube-kernel iteration: the parity tests that exercise the cube kernels, then kernel-only
# bench lines (10^7 / 1.25e6 / 10^6 linear, 10^6 cubic).  PROF=linear|cubic adds one ncu --set full capture at 10^6.
#   gpurun --timeout 900 -- 'bash tools/gpu_iter.sh TAG'
TAG=${1:-iter}
OUT=gpurun_out/$TAG
mkdir -p $OUT
KSEL=${KSEL:-"fused or cube or scale or wide or knife or pipeline_host or slab"}
timeout -s KILL 600 python -m pytest tests -m gpu -q -x -k "$KSEL" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
grep -E "passed|failed|Error" $OUT/pytest.log | tail -3
B="--steps 10 --warmup 3 --no-cpu --no-e2e --no-stage"
timeout -s KILL 200 python bench.py $B > $OUT/bench_1e7.json 2> $OUT/bench.err
timeout -s KILL 200 python bench.py $B --particles 1250000 > $OUT/bench_1250000.json 2>> $OUT/bench.err
timeout -s KILL 200 python bench.py $B --particles 1000000 > $OUT/bench_1e6.json 2>> $OUT/bench.err
timeout -s KILL 200 python bench.py $B --particles 1000000 --method cubic > $OUT/bench_1e6_cubic.json 2>> $OUT/bench.err
timeout -s KILL 200 python bench.py $B --method cubic > $OUT/bench_1e7_cubic.json 2>> $OUT/bench.err
if [ -n "$PROF" ]; then
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:fused_cube_warp -s 3 -c 1 \
    -o $OUT/prof_fused_${PROF}_1000000 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage \
    --particles 1000000 --method $PROF > $OUT/ncu_$PROF.log 2>&1
fi
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
