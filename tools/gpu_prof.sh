#!/bin/bash
# parity tests + bench + ncu launch list + full capture of the fused kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --method cubic > gpurun_out/bench_quick_cubic.json 2>> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick_cubic.json
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --particles 10000000 > gpurun_out/bench_quick_1e7.json 2>> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick_1e7.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fused_cube -s 3 -c 1 -o gpurun_out/prof_fused -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
