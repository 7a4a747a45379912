#!/bin/bash
TAG=${1:-prep}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:prep_kernel -s 3 -c 1 -o $OUT/prof_prep_1e7 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage > $OUT/ncu.log 2>&1
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:radix_pass -s 7 -c 1 -o $OUT/prof_radix_1e7 -f python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage > $OUT/ncu2.log 2>&1
ls -la $OUT
