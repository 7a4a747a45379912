#!/bin/bash
# One GPU-box visit: FP32 peak, parity tests, smoke, bench lines, ncu launch list and full captures.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag] [quick]
TAG=${1:-run}
QUICK=${2:-}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt
nproc > $OUT/nproc.txt
tools/fma_peak > $OUT/fp32_peak.json 2> $OUT/fp32_peak.err; cat $OUT/fp32_peak.json
timeout 1500 python -m pytest tests -m gpu -x -q --durations=15 > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --particles 1000000 > $OUT/bench_1e6.json 2>> $OUT/bench.err
timeout 600 python bench.py --particles 1000000 --method cubic --no-cpu > $OUT/bench_1e6_cubic.json 2>> $OUT/bench.err
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench.json; cat $OUT/bench_1e6.json
if [ -n "$QUICK" ]; then exit 0; fi
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
timeout 600 python bench.py --particles 10000000 --spaxels 150 --no-cpu --steps 5 > $OUT/bench_1e7_s150.json 2>> $OUT/bench.err
timeout 600 python bench.py --particles 1000000 --galaxies 8 --no-cpu --steps 5 > $OUT/bench_survey8.json 2>> $OUT/bench.err
timeout 600 python tools/bench_stages.py > $OUT/stages.json 2> $OUT/stages.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_1e6.csv python bench.py --particles 1000000 --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/bench_under_ncu_1e6.log 2>&1
for cfg in "linear 10000000" "linear 5000000" "linear 2500000" "linear 1250000" "linear 1000000" "cubic 1000000"; do
  set -- $cfg
  timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:fused_cube_warp -s 4 -c 2 -o $OUT/prof_fused_$1_$2 -f python bench.py --particles $2 --method $1 --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e > $OUT/ncu_fused_$1_$2.log 2>&1
done
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:march -s 2 -c 1 -o $OUT/prof_march_s150 -f python tools/prof_conv.py > $OUT/ncu_march.log 2>&1
