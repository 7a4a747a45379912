#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench lines, stage timings, ncu launch list and full captures.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $OUT/smoke.log
timeout 600 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
timeout 600 python bench.py --method cubic --no-cpu > $OUT/bench_cubic.json 2>> $OUT/bench.err
timeout 600 python bench.py --particles 10000000 --no-cpu --steps 5 > $OUT/bench_1e7.json 2>> $OUT/bench.err
timeout 600 python bench.py --particles 10000000 --spaxels 150 --no-cpu --steps 5 > $OUT/bench_1e7_s150.json 2>> $OUT/bench.err
timeout 600 python bench.py --galaxies 8 --no-cpu --steps 5 > $OUT/bench_survey8.json 2>> $OUT/bench.err
timeout 600 python tools/bench_stages.py > $OUT/stages.json 2> $OUT/stages.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/bench_under_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_cube -s 3 -c 1 -o $OUT/prof_fused -f python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_fused.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_cube -s 3 -c 1 -o $OUT/prof_fused_cubic -f python bench.py --steps 2 --warmup 3 --no-cpu --method cubic > $OUT/ncu_fused_cubic.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:march -s 2 -c 1 -o $OUT/prof_march_s150 -f python tools/prof_conv.py > $OUT/ncu_march.log 2>&1
tail -3 $OUT/pytest_gpu.log; tail -2 $OUT/smoke.log; cat $OUT/bench.json
timeout 300 python tools/bench_dusty.py > $OUT/dusty.json 2> $OUT/dusty.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/launches_dusty.csv python tools/bench_dusty.py --particles 200000 --gas 200000 --staged 20000 --reps 1 > $OUT/dusty_under_ncu.log 2>&1
cat $OUT/dusty.json
