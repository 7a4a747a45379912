#!/bin/bash
# quick cube-kernel check: a few parity tests, kernel times at 10^6 / 1.25e6 / 10^7, optional ncu capture
TAG=${1:-kern}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "fused_cube_synthetic or alternate_code_paths or knife or 1e6" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
grep -E "passed|failed" $OUT/pytest.log | tail -2
for n in 1000000 1250000 10000000; do
  timeout -s KILL 200 python bench.py --particles $n --no-cpu --no-e2e --no-parity > $OUT/bench_$n.json 2>> $OUT/bench.err
done
if [ -n "$2" ]; then
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:fused_cube_warp -s 3 -c 1 -o $OUT/prof_linear_10000000 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/ncu.log 2>&1
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.load(open(f)); print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f" % (d["ms_per_step"], d["roofline"]["kernel_ms"]))
    except Exception as e: print(f, "ERR", e)
PY
