#!/bin/bash
# ms per evaluation of the bench workloads with the INT8 split path off / on (and its minimum expert size)
for wl in cfg2 cfg3b cfg4; do
for cfg in "0 8" "1 8" "1 12" "1 16"; do
  set -- $cfg
  DSMGP_OZAKI=$1 DSMGP_OZAKI_MIN_NB=$2 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); p=d['phases_ms_per_step']; print('$wl ozaki $1 min_nb $2: potrf %.3f inverse %.3f grad %.3f total %.3f' % (p['potrf_ms'], p['inverse_ms'], p['grad_ms'], d['ms_per_step']))"
done
done
