#!/bin/bash
TAG=${1:-n2}; N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
run() { name=$1; shift
  timeout -s KILL 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"; tail -2 $OUT/$name.err; }
run bench_n$N --steps 5 --warmup 3 --no-cpu --no-stage
run bench_n${N}_s150 --steps 3 --warmup 3 --spaxels 150 --no-cpu --no-stage --particles 4000000
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g ms/step %.4f kernel_ms %.4f parity %s e2e_ms %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("ok"), d.get("e2e",{}).get("ms_per_step")))
    except Exception as e: print(f, "ERR", e)
PY
