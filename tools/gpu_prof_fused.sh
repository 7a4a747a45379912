#!/bin/bash
# one ncu --set full capture of the selected warp cube kernel:  N=10000000 METHOD=linear bash tools/gpu_prof_fused.sh TAG
TAG=${1:-pf}; N=${N:-10000000}; METHOD=${METHOD:-linear}
OUT=gpurun_out/$TAG
mkdir -p $OUT
# both warp kernels are launched (one returns at once): capture two launches, keep the long one
timeout -s KILL 400 ncu --set full --clock-control none --import-source on -k regex:fused_cube_warp -s 6 -c 2 -o $OUT/prof_fused_${METHOD}_$N -f python bench.py --particles $N --method $METHOD --steps 2 --warmup 3 --no-cpu --no-parity --no-e2e --no-stage > $OUT/ncu.log 2>&1
tail -3 $OUT/ncu.log
