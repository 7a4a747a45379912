#!/bin/bash
# Cube-kernel variants side by side (RBX_FUSED_VARIANT seeds the option table at load time): unset = 7 arrays x 2 warps,
# 2 = 6 arrays x 2 warps, 1 = one warp per array.   Usage: bash tools/gpu_variants.sh TAG
TAG=${1:-variants}
OUT=gpurun_out/$TAG
mkdir -p $OUT
# a deadlock in the turn protocol must not hang the box: short hard timeouts first
timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_cube_synthetic or alternate_code_paths" > $OUT/pytest_first.log 2>&1; echo "first pytest rc=$?" | tee -a $OUT/pytest_first.log
tail -3 $OUT/pytest_first.log
if ! grep -q " passed" $OUT/pytest_first.log || grep -q "failed" $OUT/pytest_first.log; then echo "STOP: first tests not green"; exit 1; fi
for v in 0 2 1; do
  for n in 1000000 10000000; do
    RBX_FUSED_VARIANT=$([ $v = 0 ] && echo -1 || echo $v) timeout -s KILL 200 python bench.py --particles $n --no-cpu --no-e2e > $OUT/bench_v${v}_$n.json 2>> $OUT/bench.err
  done
done
timeout -s KILL 900 python -m pytest tests -m gpu -q -k "fused or cube or knife or doppler or group or pipeline_host" --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -3
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:fused_cube_warp -s 3 -c 1 -o $OUT/prof_fused_linear_1000000 -f python bench.py --particles 1000000 --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/ncu.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_v*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f parity %s ok %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("max_abs_err_over_max"), d.get("parity",{}).get("ok")))
    except Exception as e: print(f, "ERR", e)
PY
