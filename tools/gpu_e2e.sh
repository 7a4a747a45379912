#!/bin/bash
TAG=${1:-e2e}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests -m gpu -q -k "1e6 or 1e7 or pipeline_host or radix" --durations=5 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
grep -E "passed|failed" $OUT/pytest_gpu.log | tail -2
timeout -s KILL 300 python bench.py --no-cpu > $OUT/bench_1e7.json 2> $OUT/bench.err
timeout -s KILL 300 python bench.py --particles 1000000 --no-cpu > $OUT/bench_1e6.json 2>> $OUT/bench.err
timeout -s KILL 300 python bench.py --particles 1250000 --no-cpu --no-e2e > $OUT/bench_1250k.json 2>> $OUT/bench.err
for r in 70 90; do RBX_HOST_RATIO=$r timeout -s KILL 300 python bench.py --no-cpu --no-parity > $OUT/bench_1e7_ratio$r.json 2>> $OUT/bench.err; done
RBX_HOST_CHUNKS=1 timeout -s KILL 300 python bench.py --particles 1000000 --no-cpu --no-parity > $OUT/bench_1e6_chunks1.json 2>> $OUT/bench.err
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.load(open(f)); e=d.get("e2e",{})
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f e2e %s packed %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], e.get("ms_per_step"), e.get("packed",{}).get("ms_per_step")))
    except Exception as e: print(f, "ERR", e)
PY
