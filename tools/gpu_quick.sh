#!/bin/bash
# quick GPU check: parity tests (fail fast) + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cat gpurun_out/bench_quick.json; tail -5 gpurun_out/bench_quick.err
timeout 600 python tools/bench_stages.py > gpurun_out/stages.json 2> gpurun_out/stages.err; cat gpurun_out/stages.json; tail -3 gpurun_out/stages.err
