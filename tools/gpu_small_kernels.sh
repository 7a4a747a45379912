#!/bin/bash
# launch list + full captures of the small kernels of one 1.25e6-particle shard step (the 8-GPU strong-scaling shard)
TAG=${1:-smallk}
N=${N:-1250000}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="--steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage --particles $N"
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_$N.csv python bench.py $B > $OUT/under_ncu.log 2>&1
for k in ${KERNELS:-prep_kernel segment_kernel reduce_partials}; do
  timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:$k -s ${SKIP:-3} -c 1 -o $OUT/prof_${k}_$N -f python bench.py $B > $OUT/ncu_$k.log 2>&1
done
python - <<PY
import csv
lines=[l for l in open("$OUT/launches_$N.csv") if not l.startswith('==')]
rows=list(csv.DictReader(lines))
names=[(x['Kernel Name'][:60], float(x['Metric Value'])) for x in rows if x.get('Metric Name')=='gpu__time_duration.sum']
idx=max(i for i,(n,_) in enumerate(names) if 'prep_kernel' in n)
for n,v in names[idx-3:idx+12]: print(f'  {v/1000:9.1f} us  {n}')
PY
