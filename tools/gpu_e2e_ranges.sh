#!/bin/bash
# rbx_pipeline_host range schedules at 10^7 particles (e2e only): equal ranges against a short last range
TAG=${1:-e2er}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="--steps 3 --warmup 3 --no-cpu --no-parity --no-stage"
for cfg in "5 100" "4 100" "6 100" "5 85" "6 85" "6 75" "7 80" "5 70"; do
  set -- $cfg
  RBX_HOST_CHUNKS=$1 RBX_HOST_RATIO=$2 timeout -s KILL 120 python bench.py $B > $OUT/e2e_$1_$2.json 2>> $OUT/bench.err
  python -c "
import json;d=json.load(open('$OUT/e2e_$1_$2.json'));e=d['e2e'];print('chunks $1 ratio $2: e2e %.3f ms  packed %.3f ms'%(e['ms_per_step'],e['packed']['ms_per_step']))"
done
