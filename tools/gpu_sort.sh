#!/bin/bash
TAG=${1:-sort}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 120 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fused_cube_synthetic or alternate_code_paths or large_fov" > $OUT/pytest_first.log 2>&1; echo "first pytest rc=$?" | tee -a $OUT/pytest_first.log
tail -3 $OUT/pytest_first.log
if ! grep -q " passed" $OUT/pytest_first.log || grep -q "failed" $OUT/pytest_first.log; then echo "STOP: first tests not green"; tail -40 $OUT/pytest_first.log; exit 1; fi
timeout -s KILL 900 python -m pytest tests -m gpu -q -k "${KEXPR:-fused or cube or pipeline_host or dust}" --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -3
for n in 1000000 10000000; do
  timeout -s KILL 200 python bench.py --particles $n --no-cpu --no-e2e > $OUT/bench_own_$n.json 2>> $OUT/bench.err
  RBX_SORT_IMPL=1 timeout -s KILL 200 python bench.py --particles $n --no-cpu --no-e2e > $OUT/bench_cub_$n.json 2>> $OUT/bench.err
done
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_1e6.csv python bench.py --particles 1000000 --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/under_ncu.log 2>&1
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_1e7.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/under_ncu7.log 2>&1
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.load(open(f))
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f parity %s launches %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("ok"), d["gpu_launches"]))
    except Exception as e: print(f, "ERR", e)
PY
