#!/bin/bash
TAG=${1:-psubtr}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="--steps 10 --warmup 3 --no-cpu --no-e2e --no-stage --no-parity"
for cfg in "1250000 512" "1250000 644" "1250000 768" "1000000 512" "1000000 640" "10000000 2048" "10000000 1536" "10000000 3072" "10000000 4096" "2500000 512" "2500000 768" "2500000 1024" "5000000 1024" "5000000 1536" "5000000 2048"; do
  set -- $cfg
  RBX_PSUB=$2 timeout -s KILL 100 python bench.py $B --particles $1 > $OUT/b_$1_$2.json 2>> $OUT/bench.err
  python -c "
import json;d=json.load(open('$OUT/b_$1_$2.json'));print('N=$1 psub $2: step %.4f kernel %.4f'%(d['ms_per_step'],d['roofline']['kernel_ms']))"
done
