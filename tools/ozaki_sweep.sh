#!/bin/bash
# inverse_ms of cfg3 for split depth / minimum size / slice count of the INT8 path
export DSMGP_OZAKI=1
for cfg in "1 8 8" "2 8 8" "2 6 8" "2 4 8" "3 4 8" "1 8 7" "2 6 7"; do
  set -- $cfg
  DSMGP_OZAKI_DEPTH=$1 DSMGP_OZAKI_MIN_NB=$2 DSMGP_OZAKI_SLICES=$3 python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('depth $1 min_nb $2 S $3: inverse %.3f potrf %.3f total %.3f' % (p['inverse_ms'], p['potrf_ms'], d['ms_per_step']))"
done
