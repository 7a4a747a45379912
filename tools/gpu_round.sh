#!/bin/bash
# One gpurun call: GPU tests, smoke, bench, ncu launch list, ncu full capture of the top kernel.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_linear.json 2> gpurun_out/bench_linear.err; echo "bench rc=$?"; cat gpurun_out/bench_linear.json
python bench.py --steps 10 --warmup 3 --method cubic --no-cpu > gpurun_out/bench_cubic.json 2> gpurun_out/bench_cubic.err; cat gpurun_out/bench_cubic.json
python bench.py --steps 5 --warmup 3 --particles 10000000 --no-cpu > gpurun_out/bench_1e7.json 2> gpurun_out/bench_1e7.err; cat gpurun_out/bench_1e7.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err; cat gpurun_out/bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:fused_cube -s 3 -c 1 -o gpurun_out/prof_fused -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_full.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:psf_lsf -s 3 -c 1 -o gpurun_out/prof_psflsf -f python bench.py --steps 1 --warmup 3 --no-cpu >> gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
