#!/bin/bash
# average ncu time of kernels matching $1 in a short bench run (cold-cache, serialised): bash tools/gpu_kernel_time.sh prep
mkdir -p gpurun_out/kt
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:$1 -c 40 --csv --log-file gpurun_out/kt/$1.csv python bench.py --steps 3 --warmup 3 --no-cpu ${@:2} > /dev/null 2>&1
python - <<PY
import csv
rows=list(csv.reader(open('gpurun_out/kt/$1.csv')))
hi=[i for i,r in enumerate(rows) if 'Kernel Name' in r][0]
h=rows[hi]; ix={k:i for i,k in enumerate(h)}
v=[]
for r in rows[hi+1:]:
    if len(r)<len(h) or r[ix['Metric Name']]!='gpu__time_duration.sum': continue
    x=float(r[ix['Metric Value']].replace(',','')); u=r[ix['Metric Unit']]
    v.append(x/1000 if u=='ns' else (x*1000 if u=='ms' else x))
v.sort()
print('$1', 'n', len(v), 'median us', v[len(v)//2] if v else None, 'min', v[0] if v else None)
PY
