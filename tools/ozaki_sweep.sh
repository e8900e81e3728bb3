#!/bin/bash
# ms per evaluation of cfg3 against the smallest expert (block rows) that takes the INT8 split path
for nb in 6 8 10 12 16; do
  DSMGP_OZAKI_MIN_NB=$nb python bench.py --workload cfg3 --steps 4 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; i=d['int8_split']; print('min_nb $nb: total %.3f  potrf %.3f inverse %.3f | gemm %.2f slice %.2f fp64 tile %.2f' % (d['ms_per_step'], p['potrf_ms'], p['inverse_ms'], i['gemm_ms'], i['slice_ms'], i['fp64_tile_ms']))"
done
