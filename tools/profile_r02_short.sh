#!/bin/bash
# launch list and the slicing-kernel capture of the final library (the rest of profile_r02.sh is unchanged by the last commits)
CMD="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records"
mkdir -p gpurun_out /tmp/prof
$CMD > gpurun_out/plain_r02.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"potrf2|trtri3|eval2|gemm_kernel|slice_kernel|parts_kernel|setflags|gram_fit|rows_kernel|alpha_reduce|predict3|route_kernel|mix_kernel|solve3|lauum3|gather_kernel" -c 600 --csv --log-file gpurun_out/launches_r02.csv $CMD > gpurun_out/ncu_list_r02.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:slice_kernel -s 27 -c 1 -f -o /tmp/prof/oz_slice_kernel $CMD > gpurun_out/ncu_run_oz_slice_kernel_r02.log 2>&1
echo "full oz_slice rc=$?"
{ echo "# ncu --set full --clock-control none --import-source on -k regex:oz_slice_kernel (one launch of: $CMD)"; python tools/ncu_summary.py /tmp/prof/oz_slice_kernel.ncu-rep 12; echo; echo "# stall samples by CUDA source line"; python tools/ncu_lines.py /tmp/prof/oz_slice_kernel.ncu-rep 14; } > gpurun_out/ncu_full_oz_slice_kernel_r02.txt 2>&1
head -14 gpurun_out/ncu_full_oz_slice_kernel_r02.txt
