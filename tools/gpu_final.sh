#!/bin/bash
# Final artefacts of a build on one GPU (trimmed gpu_check.sh: output stays below gpurun's 64 MiB): bench lines, launch
# lists, one ncu --set full capture of the selected cube kernel per shard size.   bash tools/gpu_final.sh TAG
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --particles 1000000 --no-cpu > $OUT/bench_1e6.json 2>> $OUT/bench.err
timeout 600 python bench.py --particles 1000000 --method cubic --no-cpu > $OUT/bench_1e6_cubic.json 2>> $OUT/bench.err
timeout 600 python bench.py --method cubic --no-cpu --no-stage > $OUT/bench_1e7_cubic.json 2>> $OUT/bench.err
timeout 600 python bench.py --particles 10000000 --spaxels 150 --no-cpu --steps 5 --no-stage > $OUT/bench_1e7_s150.json 2>> $OUT/bench.err
timeout 600 python bench.py --particles 1000000 --galaxies 8 --no-cpu --steps 5 --no-stage > $OUT/bench_survey8.json 2>> $OUT/bench.err
for N in 10000000 1250000 1000000; do
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$N.csv python bench.py --particles $N --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage > $OUT/under_ncu_$N.log 2>&1
done
# the selected warp kernel is the first of the (up to two) warp-kernel launches of a step: even launch indices
for cfg in "linear 10000000 4" "linear 5000000 4" "linear 2500000 4" "linear 1250000 4" "linear 1000000 4" "cubic 1000000 4" "cubic 10000000 4"; do
  set -- $cfg
  timeout -s KILL 400 ncu --set full --clock-control none -k regex:fused_cube_warp -s $3 -c 1 -o $OUT/prof_fused_$1_$2 -f python bench.py --particles $2 --method $1 --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage > $OUT/ncu_fused_$1_$2.log 2>&1
done
du -sh $OUT
python - <<PY
import json
for f in ["bench","bench_1e6","bench_1e6_cubic","bench_1e7_cubic","bench_1e7_s150","bench_survey8"]:
    try:
        d=json.load(open("$OUT/"+f+".json")); print(f, round(d["ms_per_step"],4), d.get("roofline",{}).get("kernel_ms"), d.get("e2e",{}).get("ms_per_step"), d.get("parity",{}).get("ok"), d.get("roofline_psf_lsf",{}).get("frac"))
    except Exception as e: print(f, "ERR", e)
PY
