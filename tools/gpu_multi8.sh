#!/bin/bash
# 8-GPU visit: strong-scaling lines of config 3 (N = 8, 4), config 4 (150 x 150, reduce-scatter) and config 5 (64 galaxies).
TAG=${1:-multi8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt
run() {  # name, nproc, args...
  name=$1; N=$2; shift; shift
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"
}
run bench_n8 8 --steps 10 --warmup 3
run bench_n8_s150 8 --steps 5 --warmup 3 --spaxels 150
run bench_n8_survey 8 --steps 5 --warmup 3 --particles 1000000 --galaxies 8
run bench_n4 4 --steps 10 --warmup 3
run bench_n4_s150 4 --steps 5 --warmup 3 --spaxels 150
NCCL_DEBUG=INFO timeout -s KILL 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 8 --steps 3 --warmup 3 --no-e2e --no-parity 2>&1 | grep -E "NVLS|Channel|Connected|nranks|Using network" | head -30 > $OUT/nccl_info.txt
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g ms/step %.4f kernel_ms %.4f parity %s e2e_ms %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("ok"), d.get("e2e",{}).get("ms_per_step")))
    except Exception as e: print(f, "ERR", e)
PY
