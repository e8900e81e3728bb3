#!/bin/bash
# A/B of the trtri3 task-list grouping (L2 reuse by scheduling): kernel time per (G, stagger) and ncu DRAM bytes of two settings.
CMD="python bench.py --workload cfg3 --steps 5 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict"
for cfg in "1 1" "2 4" "4 4" "4 8" "4 16" "8 8"; do set -- $cfg; DSMGP_TRTRI_GROUP=$1 DSMGP_TRTRI_STAGGER=$2 $CMD 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('G=$1 stagger=$2', 'inverse_ms', round(d['phases_ms_per_step']['inverse_ms'],3), 'total', round(d['ms_per_step'],3))"; done
CMD1="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict"
for cfg in "4 8" "2 4"; do set -- $cfg; DSMGP_TRTRI_GROUP=$1 DSMGP_TRTRI_STAGGER=$2 $CMD1 > /dev/null 2>&1 && DSMGP_TRTRI_GROUP=$1 DSMGP_TRTRI_STAGGER=$2 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:trtri3_kernel -s 3 -c 1 --csv --log-file gpurun_out/trtri3_g$1_s$2.csv $CMD1 > /dev/null 2>&1; grep -E "dram__bytes_read|gpu__time" gpurun_out/trtri3_g$1_s$2.csv | awk -F, '{print $(NF-2), $(NF)}'; done
