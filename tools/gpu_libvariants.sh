#!/bin/bash
# A/B of compile-time variants: every rubix_b200/variants/lib_*.so takes the library's place in turn (on the box only)
#   ARGS="--method cubic" VARS="2 4" bash tools/gpu_libvariants.sh TAG
TAG=${1:-libv}
OUT=gpurun_out/$TAG
mkdir -p $OUT
B="--steps 10 --warmup 3 --no-cpu --no-e2e --no-stage $ARGS"
for lib in rubix_b200/variants/lib_*.so; do
  name=$(basename $lib .so)
  cp $lib rubix_b200/librubix_b200.so
  for v in ${VARS:-2}; do
    for n in ${SIZES:-1000000 10000000}; do
      RBX_FUSED_VARIANT=$v timeout -s KILL 200 python bench.py $B --particles $n > $OUT/bench_${name}_v${v}_$n.json 2>> $OUT/bench.err
    done
  done
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$OUT/bench_*.json")):
    try:
        d = json.load(open(f))
        print(f.split("/")[-1], "ms/step %.4f kernel_ms %.4f parity %s" % (d["ms_per_step"], d["roofline"]["kernel_ms"], d.get("parity", {}).get("ok")))
    except Exception as e:
        print(f, "ERR", e)
PY
tail -3 $OUT/bench.err
