#!/bin/bash
# PSF/LSF parity tests + conv stage timings (+ optional ncu capture of the marching kernel: PROF=1)
TAG=${1:-conv}
mkdir -p gpurun_out/$TAG
timeout 300 python -m pytest tests -m gpu -x -q -k "psf or lsf" 2>&1 | tail -3
timeout 300 python tools/bench_stages.py --conv-only > gpurun_out/$TAG/conv.json 2> gpurun_out/$TAG/conv.err; tail -3 gpurun_out/$TAG/conv.err
python -c "
import json; d=json.load(open('gpurun_out/$TAG/conv.json'))
for k,v in d.items(): print(k, round(v['ms_median'],4), round(v['achieved_gbs']))"
if [ -n "$PROF" ]; then
timeout 300 ncu --set full --clock-control none --import-source on -k regex:march -s 2 -c 1 -o gpurun_out/$TAG/prof_march -f python tools/prof_conv.py > gpurun_out/$TAG/ncu.log 2>&1; tail -2 gpurun_out/$TAG/ncu.log
fi
