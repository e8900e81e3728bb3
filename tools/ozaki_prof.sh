#!/bin/bash
# S = 7 vs 8 A/B and one ncu --set full capture of the INT8 block-product kernel inside the cfg3 evaluation.
export DSMGP_OZAKI=1
mkdir -p gpurun_out /tmp/prof
DSMGP_OZAKI_SLICES=7 timeout 400 python tools/ozaki_ab.py cfg3 2>&1 | tail -6
CMD="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict"
$CMD > /dev/null 2>&1 || exit 1
for K in gemm_kernel slice_kernel; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$K -s 7 -c 1 -f -o /tmp/prof/$K $CMD > gpurun_out/ncu_run_oz_$K.log 2>&1
  echo "full $K rc=$?"
  { python tools/ncu_summary.py /tmp/prof/$K.ncu-rep 14; echo; python tools/ncu_lines.py /tmp/prof/$K.ncu-rep 14; } > gpurun_out/ncu_full_oz_$K.txt 2>&1
  ncu -i /tmp/prof/$K.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); H=rows[0]; V=rows[-1]
for k,v in zip(H,V):
    if any(t in k for t in ('lts__t_bytes.sum','lts__throughput','l1tex__m_xbar2l1tex_read_bytes.sum','sm__throughput','dram__throughput','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','lts__t_sector_hit_rate','sm__inst_executed_pipe_tensor','tensor')): print(k,'=',v)
" > gpurun_out/ncu_l2_oz_$K.txt 2>&1
  head -40 gpurun_out/ncu_l2_oz_$K.txt
done
