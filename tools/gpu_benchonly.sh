#!/bin/bash
TAG=${1:-bo}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 50 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('$OUT/bench.json'));r=d['roofline'];print(d['ms_per_step'],d['value'],r['frac'],r['issue_frac'],r['hbm_frac'],d['e2e']['ms_per_step'],d['parity']['ok'],d['cpu_baseline']['value'])"
