#!/bin/bash
# fine sweep of the work-item size at one shard size (does the queue model of DESIGN section 5 predict the kernel time?)
TAG=${1:-psubfine}
N=${N:-1250000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="--steps 10 --warmup 3 --no-cpu --no-e2e --no-stage --no-parity --particles $N"
for ts in 3 2; do
for p in ${PSUBS:-384 448 512 576 644 704 768 896}; do
  RBX_PSUB=$p RBX_TAIL_SHIFT=$ts timeout -s KILL 100 python bench.py $B > $OUT/b_${p}_$ts.json 2>> $OUT/bench.err
  python -c "
import json;d=json.load(open('$OUT/b_${p}_$ts.json'));print('N=$N psub $p ts $ts: step %.4f kernel %.4f'%(d['ms_per_step'],d['roofline']['kernel_ms']))"
done
done
