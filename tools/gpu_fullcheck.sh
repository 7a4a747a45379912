#!/bin/bash
# the whole GPU suite, then smoke() if the remaining limit allows
TAG=${1:-fc}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 125 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; grep -v "^\[" $OUT/pytest.log | tail -12
timeout -s KILL 25 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/smoke.log
