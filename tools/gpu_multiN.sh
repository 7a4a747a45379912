#!/bin/bash
# N-GPU visit (gpurun --gpus N): strong-scaling line of config 3, and with FULL=1 config 4 (150 x 150, reduce-scatter),
# config 5 (survey batch) and the cubic method.   bash tools/gpu_multiN.sh TAG N
TAG=${1:-multi}; N=${2:-8}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt
run() {  # name, args...
  name=$1; shift
  timeout -s KILL 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"
}
run bench_n$N --steps 10 --warmup 3 --no-cpu
if [ -n "$FULL" ]; then
run bench_n${N}_s150 --steps 5 --warmup 3 --spaxels 150 --no-cpu
run bench_n${N}_survey --steps 5 --warmup 3 --particles 1000000 --galaxies 8 --no-cpu
run bench_n${N}_cubic --steps 10 --warmup 3 --method cubic --no-cpu
fi
if [ -n "$MGPUTEST" ]; then
timeout -s KILL 600 python -m pytest tests -m gpu -q -k "two_ranks" > $OUT/mgpu_test.log 2>&1; tail -3 $OUT/mgpu_test.log
fi
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g ms/step %.4f kernel_ms %.4f parity %s e2e_ms %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("ok"), d.get("e2e",{}).get("ms_per_step")))
    except Exception as e: print(f, "ERR", e)
PY
