#!/bin/bash
TAG=${1:-pc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 28 python -m pytest tests/test_gpu_factories.py -m gpu -q -x -k "rubix_pipeline" > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^\[" $OUT/pytest.log | tail -5
