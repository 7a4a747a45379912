#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "psf or lsf" 2>&1 | tail -5
timeout 600 python tools/bench_stages.py --reps 10 > gpurun_out/stages.json 2> gpurun_out/stages.err; tail -3 gpurun_out/stages.err
python -c "
import json;d=json.load(open('gpurun_out/stages.json'))
for k,v in d.items(): print(k, {a:(round(b,4) if isinstance(b,float) else b) for a,b in v.items()})"
