#!/bin/bash
# final single-GPU bench lines of round 2 (INT8 split path on by default; *_fp64 = DSMGP_OZAKI=0)
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_r02h.json 2> gpurun_out/bench_r02h.err; echo "default rc=$?"
DSMGP_OZAKI=0 python bench.py --no-sub-records > gpurun_out/bench_r02h_fp64.json 2>/dev/null; echo "fp64 rc=$?"
python bench.py --impl reference > gpurun_out/bench_r02h_reference.json 2>/dev/null; echo "reference rc=$?"
for wl in cfg2 cfg4 cfg3b; do python bench.py --workload $wl --no-cpu-baseline --no-sub-records > gpurun_out/bench_r02h_$wl.json 2>/dev/null; echo "$wl rc=$?"; done
for f in gpurun_out/bench_r02h.json gpurun_out/bench_r02h_fp64.json gpurun_out/bench_r02h_cfg2.json gpurun_out/bench_r02h_cfg4.json gpurun_out/bench_r02h_cfg3b.json; do
tail -1 $f | python -c "
import json,sys; d=json.loads(sys.stdin.read()); r=d['roofline']; print('$f', round(d['ms_per_step'],3), 'ms', round(d['value'],3), 'e2e', round(d['e2e']['value'],3), r['kernel'], round(r['achieved'],1), round(r['frac'],3), d['clocks'], (d.get('int8_split') or {}).get('share_of_flops_on_int8'), d.get('predict',{}).get('wall_ms'), (d.get('scale_cfg5') or {}).get('seconds_per_evaluation'), (d.get('mathematical') or {}).get('ms_per_step'))"
done
