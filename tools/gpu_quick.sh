#!/bin/bash
# quick GPU check: parity tests (fail fast) + one bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.log
for bits in 16 20; do
RBX_SORT_BITS=$bits timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; tail -5 gpurun_out/bench_quick.err
python -c "
import json;d=json.load(open('gpurun_out/bench_quick.json'));print('bits $bits step ms',d['ms_per_step'],'kernel ms',d['roofline']['kernel_ms'],'e2e ms',d['e2e']['ms_per_step'], d['clocks'])"
done
