#!/bin/bash
# Short GPU-box visit while iterating on the cube kernel: the fused-cube parity tests, bench lines at 10^6 / 10^7 and
# one ncu --set full capture of the cube kernel at 10^6.   Usage: bash tools/gpu_quick.sh TAG [pytest -k expression]
TAG=${1:-quick}
KEXPR=${2:-"fused or cube or knife or doppler or group or pipeline_host"}
OUT=gpurun_out/$TAG
mkdir -p $OUT
tools/fma_peak > $OUT/fp32_peak.json 2>/dev/null
timeout 1200 python -m pytest tests -m gpu -q -k "$KEXPR" --durations=8 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python bench.py --particles 1000000 --no-cpu --no-e2e > $OUT/bench_1e6.json 2> $OUT/bench.err
timeout 300 python bench.py --particles 1000000 --no-cpu --no-e2e --method cubic > $OUT/bench_1e6_cubic.json 2>> $OUT/bench.err
timeout 300 python bench.py --no-cpu --no-e2e > $OUT/bench_1e7.json 2>> $OUT/bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_cube_warp -s 3 -c 1 -o $OUT/prof_fused_linear_1000000 -f python bench.py --particles 1000000 --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/ncu.log 2>&1
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -3
python - <<PY
import json
for f in ("bench_1e6","bench_1e6_cubic","bench_1e7"):
    try:
        d=json.load(open("$OUT/%s.json"%f))
        print(f, "ms/step %.4f kernel_ms %.4f parity %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("max_abs_err_over_max")))
    except Exception as e: print(f, "ERR", e)
PY
