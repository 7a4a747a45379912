#!/bin/bash
# work-item size sweep for the 12-warp cube kernel (option psub), per particle count
TAG=${1:-psub}
OUT=gpurun_out/$TAG
mkdir -p $OUT
for n in 1000000 1250000 2500000 5000000 10000000; do
  for ps in 128 256 512 1024 2048; do
    RBX_PSUB=$ps timeout -s KILL 120 python bench.py --particles $n --steps 6 --warmup 3 --no-cpu --no-e2e --no-parity 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('n=$n psub=$ps ms/step %.4f kernel_ms %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))" | tee -a $OUT/sweep.txt
  done
done
# one rank's work of the 8-GPU large-FOV case on one GPU: launch list
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file $OUT/launches_s150_shard.csv python tools/emulate_slab_rank.py > $OUT/emulate.log 2>&1
tail -3 $OUT/emulate.log
