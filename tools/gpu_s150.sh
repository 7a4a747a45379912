#!/bin/bash
# config 4 on one GPU + one rank's device work of the 8-GPU large-FOV step (launch list), after the tests
TAG=${1:-s150}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 600 python -m pytest tests -m gpu -q -x > $OUT/pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $OUT/pytest.log
B="--steps 5 --warmup 3 --no-cpu --no-e2e --no-stage"
timeout -s KILL 200 python bench.py $B --spaxels 150 > $OUT/bench_1e7_s150.json 2> $OUT/bench.err
timeout -s KILL 200 python bench.py $B --particles 1250000 > $OUT/bench_1250000.json 2>> $OUT/bench.err
timeout -s KILL 200 python bench.py $B > $OUT/bench_1e7.json 2>> $OUT/bench.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_s150_shard.csv python tools/emulate_slab_rank.py > $OUT/emulate.log 2>&1
python - <<PY
import json, glob, csv
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f parity %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity", {}).get("ok")))
    except Exception as e:
        print(f, "ERR", e)
lines=[l for l in open("$OUT/launches_s150_shard.csv") if not l.startswith('==')]
rows=list(csv.DictReader(lines))
names=[(x['Kernel Name'][:60], float(x['Metric Value'])) for x in rows if x.get('Metric Name')=='gpu__time_duration.sum']
idx=max(i for i,(n,_) in enumerate(names) if 'prep_kernel' in n)
for n,v in names[idx:idx+13]: print(f'  {v/1000:9.1f} us  {n}')
PY
