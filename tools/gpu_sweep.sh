#!/bin/bash
# work-item size sweep for the cube kernel (RBX_PSUB / RBX_SMALL_SHIFT / RBX_TAIL_SHIFT)
for n in 1000000 10000000; do
for ps in 128 256 512 1024 2048; do for ss in 1 2 3; do for ts in 2 3; do
if [ $n = 10000000 ] && [ $ps -lt 512 ]; then continue; fi
if [ $n = 1000000 ] && [ $ps -gt 512 ]; then continue; fi
RBX_PSUB=$ps RBX_SMALL_SHIFT=$ss RBX_TAIL_SHIFT=$ts python bench.py --steps 6 --warmup 3 --no-cpu --particles $n 2>/dev/null | python -c "
import json,sys;d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('n $n psub $ps small>>$ss tail>>$ts step ms',round(d['ms_per_step'],4),'kernel ms',round(d['roofline']['kernel_ms'],4))"
done; done; done; done
