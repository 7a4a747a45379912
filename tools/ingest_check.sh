#!/bin/bash
# Copy what a tools/gpu_check.sh visit produced into profiles/ (run in the build container):
#   bash tools/ingest_check.sh gpurun_out/TAG vNN
D=$1; V=$2
for cfg in "linear 10000000" "linear 5000000" "linear 2500000" "linear 1250000" "linear 1000000" "cubic 1000000"; do
  set -- $cfg
  rep=$D/prof_fused_$1_$2.ncu-rep
  [ -f $rep ] || continue
  txt=profiles/r02_fused_${V}_$1_$2_summary.txt
  python tools/ncu_summary.py $rep $2 $1_$2 $txt > $txt
done
[ -f $D/prof_march_s150.ncu-rep ] && python tools/ncu_summary.py $D/prof_march_s150.ncu-rep > profiles/r02_march_s150_${V}_summary.txt
for f in bench bench_1e6 bench_1e6_cubic bench_1e7_s150 bench_survey8 bench_reference stages; do
  [ -s $D/$f.json ] && cp $D/$f.json profiles/r02_${f}_${V}.json
done
[ -s $D/launches.csv ] && cp $D/launches.csv profiles/r02_launches_${V}_1e7.csv
[ -s $D/launches_1e6.csv ] && cp $D/launches_1e6.csv profiles/r02_launches_${V}_1e6.csv
[ -s $D/fp32_peak.json ] && cp $D/fp32_peak.json profiles/fp32_peak.json
[ -s $D/pytest_gpu.log ] && tail -25 $D/pytest_gpu.log > profiles/r02_pytest_gpu_${V}.txt
python tools/sass_summary.py > profiles/r02_sass_${V}.txt 2>/dev/null
ls profiles | grep _${V} | wc -l
