#!/bin/bash
TAG=${1:-hs}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 300 python -m pytest tests -m gpu -q -x -k "build_cube_host or transposed or pipeline_host or 1e6" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/pytest.log
timeout -s KILL 200 python tools/e2e_shard_probe.py 2>&1 | tee $OUT/probe.txt | tail -20
