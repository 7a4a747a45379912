#!/bin/bash
TAG=${1:-small}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout -s KILL 300 python -m pytest tests -m gpu -q -k "rotated or fused_cube_synthetic or 1e7" > $OUT/pytest.log 2>&1; echo "pytest rc=$?" | tee -a $OUT/pytest.log
grep -E "passed|failed|Error" $OUT/pytest.log | tail -3
timeout -s KILL 200 python bench.py --no-cpu > $OUT/bench_1e7.json 2> $OUT/bench.err
timeout -s KILL 200 python bench.py --particles 1000000 --no-cpu > $OUT/bench_1e6.json 2>> $OUT/bench.err
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_1e7.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage > $OUT/under_ncu7.log 2>&1
python - <<PY
import json,glob,csv
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d=json.load(open(f)); e=d.get("e2e",{})
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f e2e %s psf_lsf %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], e.get("ms_per_step"), d.get("roofline_psf_lsf",{}).get("frac")))
    except Exception as e: print(f, "ERR", e)
lines=[l for l in open("$OUT/launches_1e7.csv") if not l.startswith('==')]
rows=list(csv.DictReader(lines))
names=[(x['Kernel Name'][:50], float(x['Metric Value'])) for x in rows if x.get('Metric Name')=='gpu__time_duration.sum']
idx=max(i for i,(n,_) in enumerate(names) if 'prep_kernel' in n)
for n,v in names[idx:idx+5]: print(f'  {v/1000:9.1f} us  {n}')
PY
