#!/bin/bash
# Multi-GPU visit (gpurun --gpus N): the NCCL worker test, strong-scaling bench lines for config 3 (MUSE) and config 4
# (150 x 150, reduce-scatter), optionally the survey batch.   Usage: bash tools/gpu_multi.sh TAG N [survey]
TAG=${1:-multi}
N=${2:-2}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name --format=csv > $OUT/smi.txt
timeout -s KILL 600 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "two_ranks" -s > $OUT/pytest_mgpu.log 2>&1; echo "mgpu pytest rc=$?" | tee -a $OUT/pytest_mgpu.log
tail -5 $OUT/pytest_mgpu.log
run() {  # name, args...
  name=$1; shift
  timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N "$@" > $OUT/$name.json 2> $OUT/$name.err; echo "$name rc=$?"
}
run bench_n${N} --steps 10 --warmup 3
run bench_n${N}_s150 --steps 5 --warmup 3 --spaxels 150 --no-cpu
if [ -n "$3" ]; then run bench_n${N}_survey --steps 5 --warmup 3 --particles 1000000 --galaxies 8 --no-cpu; fi
python - <<PY
import json,glob
for f in sorted(glob.glob("$OUT/bench_n*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], "value %.4g ms/step %.4f kernel_ms %.4f parity %s e2e_ms %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity",{}).get("ok"), d.get("e2e",{}).get("ms_per_step")))
    except Exception as e: print(f, "ERR", e)
PY
