#!/bin/bash
TAG=${1:-sc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 110 python -m pytest tests/test_gpu_sort.py tests/test_gpu_reference_vectors.py -m gpu -q -s --runxfail > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^\[" $OUT/pytest.log | tail -25
timeout -s KILL 60 python tools/sort_probe.py $OUT/sort_probe.json > $OUT/probe.log 2>&1; echo "probe rc=$?"; tail -4 $OUT/probe.log
