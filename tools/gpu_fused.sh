#!/bin/bash
# fused-kernel iteration: parity tests + bench lines (optionally PROF=1 for an ncu capture)
TAG=${1:-fused}
mkdir -p gpurun_out/$TAG
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/$TAG/pytest.log
for impl in warp group; do
RBX_FUSED_IMPL=$impl timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/$TAG/bench_$impl.json 2> gpurun_out/$TAG/bench_$impl.err; tail -3 gpurun_out/$TAG/bench_$impl.err
python -c "
import json;d=json.load(open('gpurun_out/$TAG/bench_$impl.json'));print('$impl step ms',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['kernel_ms'],4),'e2e ms',round(d['e2e']['ms_per_step'],3), 'value', round(d['value']/1e6,1))"
done
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --method cubic > gpurun_out/$TAG/bench_cubic.json 2>> gpurun_out/$TAG/bench.err
python -c "
import json;d=json.load(open('gpurun_out/$TAG/bench_cubic.json'));print('cubic step ms',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['kernel_ms'],4),'e2e ms',round(d['e2e']['ms_per_step'],3), 'value', round(d['value']/1e6,1))"
if [ -n "$PROF" ]; then
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fused_cube -s 3 -c 1 -o gpurun_out/$TAG/prof_fused -f python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/$TAG/ncu_full.log 2>&1; tail -2 gpurun_out/$TAG/ncu_full.log
fi
