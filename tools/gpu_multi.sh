#!/bin/bash
# multi-GPU check (gpurun --gpus N): default bench and the large-FOV slab-sharded configuration
N=${1:-2}
TAG=${2:-multi}
mkdir -p gpurun_out/$TAG
timeout 300 python -m pytest tests -m gpu -x -q -k "slab or psf or lsf" 2>&1 | tail -2
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N "${@:2}"; }
timeout 300 bash -c "$(declare -f run); N=$N; run 29521 --steps 10 --warmup 3 --no-cpu" > gpurun_out/$TAG/bench_n$N.json 2> gpurun_out/$TAG/bench_n$N.err
timeout 400 bash -c "$(declare -f run); N=$N; run 29522 --steps 5 --warmup 3 --no-cpu --spaxels 150 --particles 1250000" > gpurun_out/$TAG/bench_n${N}_s150.json 2> gpurun_out/$TAG/bench_n${N}_s150.err
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --spaxels 150 --particles 1250000 > gpurun_out/$TAG/bench_n1_s150.json 2> gpurun_out/$TAG/bench_n1_s150.err
for f in bench_n$N bench_n${N}_s150 bench_n1_s150; do tail -2 gpurun_out/$TAG/$f.err | cut -c1-300; python -c "
import json;d=json.loads(open('gpurun_out/$TAG/$f.json').read().strip().splitlines()[-1]);print('$f', 'n_gpus',d['n_gpus'],'step ms',round(d['ms_per_step'],4),'e2e ms',round(d['e2e']['ms_per_step'],3), 'value M/s', round(d['value']/1e6,1), d['config']['parallelism'][:60])"; done
