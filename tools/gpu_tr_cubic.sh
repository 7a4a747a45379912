#!/bin/bash
TAG=${1:-trc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -E "passed|failed|rror" $OUT/pytest.log | tail -5
B="--steps 10 --warmup 3 --no-cpu --no-e2e --no-stage --method cubic"
for tr in 1 0; do
for n in 10000000 1000000; do
  RBX_FUSED_TR=$tr timeout -s KILL 200 python bench.py $B --particles $n > $OUT/bench_cubic_tr${tr}_$n.json 2>> $OUT/bench.err
done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f parity %s %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity", {}).get("ok"), d.get("parity", {}).get("max_abs_err_over_max")))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $OUT/bench.err
