#!/bin/bash
# quick GPU iteration: parity subset + bench on cfg3 / cfg3b.  usage: tools/quick.sh <tag> [pytest-args]
TAG=${1:-q}
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
for WL in cfg3 cfg3b; do
  timeout 300 python bench.py --workload $WL --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_${TAG}_$WL.json 2> gpurun_out/bench_${TAG}_$WL.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/bench_${TAG}_$WL.json").read().strip().splitlines()[-1])
    print("$WL", round(d["ms_per_step"],2), d["phases_ms_per_step"], "potrf TF", round(d["roofline"]["potrf_tflops"],2), "inv TF", round(d["roofline"]["inverse_tflops"],2))
except Exception as e:
    print("$WL failed", e); print(open("gpurun_out/bench_${TAG}_$WL.err").read()[-1500:])
PY
done
