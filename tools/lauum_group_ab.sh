#!/bin/bash
# A/B of the lauum3 task-list grouping (L2 reuse by scheduling) on the mathematical-gradient cfg3 and on cfg4 (IsoSE needs LAUUM).
for G in 1 2 4; do
  DSMGP_LAUUM_GROUP=$G python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict --mathematical 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg3 math G=$G grad_ms', round(d['phases_ms_per_step']['grad_ms'],3), 'total', round(d['ms_per_step'],3))"
  DSMGP_LAUUM_GROUP=$G python bench.py --workload cfg4 --steps 3 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('cfg4 G=$G grad_ms', round(d['phases_ms_per_step']['grad_ms'],3), 'total', round(d['ms_per_step'],3))"
done
CMD="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict"
$CMD > /dev/null 2>&1 && ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:trtri3_kernel -s 3 -c 1 --csv --log-file gpurun_out/trtri3_default_r02.csv $CMD > /dev/null 2>&1
grep -E "dram__bytes|gpu__time|hit_rate" gpurun_out/trtri3_default_r02.csv | awk -F, '{print $(NF-2), $(NF)}'
