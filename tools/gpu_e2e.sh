mkdir -p gpurun_out/s6r
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for c in 1 2 3 4; do RBX_HOST_CHUNKS=$c python bench.py --steps 10 --warmup 3 --no-cpu | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('1e6 chunks $c step ms',round(d['ms_per_step'],4),'e2e ms',round(d['e2e']['ms_per_step'],3))"; done
for c in 1 2 4 6 8; do RBX_HOST_CHUNKS=$c python bench.py --steps 5 --warmup 3 --no-cpu --particles 10000000 | python -c "
import json,sys;d=json.loads(sys.stdin.read());print('1e7 chunks $c step ms',round(d['ms_per_step'],4),'e2e ms',round(d['e2e']['ms_per_step'],3))"; done
