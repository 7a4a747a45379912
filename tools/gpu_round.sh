#!/bin/bash
# tests + kernel-only bench lines + a launch list of the 1.25e6-particle shard and of the 10^7 step
TAG=${1:-round}
bash tools/gpu_iter.sh $TAG
OUT=gpurun_out/$TAG
for N in 1250000 10000000; do
B="--steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage --particles $N"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$N.csv python bench.py $B > $OUT/under_ncu_$N.log 2>&1
python - <<PY
import csv
lines=[l for l in open("$OUT/launches_$N.csv") if not l.startswith('==')]
rows=list(csv.DictReader(lines))
names=[(x['Kernel Name'][:60], float(x['Metric Value'])) for x in rows if x.get('Metric Name')=='gpu__time_duration.sum']
idx=max(i for i,(n,_) in enumerate(names) if 'prep_kernel' in n)
print("N=$N")
for n,v in names[idx-1:idx+10]: print(f'  {v/1000:9.1f} us  {n}')
PY
done
