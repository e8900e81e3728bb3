#!/bin/bash
for oz in 0 1; do
DSMGP_OZAKI=$oz python bench.py --workload cfg3 --steps 20 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('OZAKI=$oz', {k: round(v,3) for k,v in d['phases_ms_per_step'].items()}, round(d['ms_per_step'],3), d['clocks'], 'e2e', d['e2e']['value'])"
done
nvidia-smi --query-gpu=power.draw,power.limit,clocks.sm,clocks.max.sm,temperature.gpu --format=csv
