#!/bin/bash
# A/B bench of kernel variants on the GPU box: tools/ab.sh <variant names...>   (built by tools/variant.sh)
mkdir -p gpurun_out
for V in "$@"; do
  for WL in ${WLS:-cfg3 cfg3b}; do
    DSMGP_LIB_PATH=$PWD/build_variants/libdsmgp_$V.so timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/ab_${V}_$WL.json 2> gpurun_out/ab_${V}_$WL.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/ab_${V}_$WL.json").read().strip().splitlines()[-1])
    ph=d["phases_ms_per_step"]
    print("$V $WL total %.2f gram %.2f potrf %.2f inv %.2f | potrf TF %.2f inv TF %.2f" % (d["ms_per_step"], ph["gram_ms"], ph["potrf_ms"], ph["inverse_ms"], d["roofline"]["potrf_tflops"], d["roofline"]["inverse_tflops"]))
except Exception as e:
    print("$V $WL failed", e); print(open("gpurun_out/ab_${V}_$WL.err").read()[-1500:])
PY
  done
done
