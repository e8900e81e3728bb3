#!/bin/bash
# Round-2 profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one `ncu --set full` capture per
# top kernel, summarised ON THE GPU BOX (the .ncu-rep files exceed what gpurun_out carries back).  usage: tools/profile_r02.sh <tag>
set -u
TAG=${1:-r02}
CMD="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records"
CMDM="python bench.py --workload cfg3 --steps 1 --warmup 3 --no-cpu-baseline --no-sub-records --no-predict --mathematical"
mkdir -p gpurun_out /tmp/prof
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | head -c 400; echo
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"potrf2|trtri3|eval2|gemm_kernel|slice_kernel|parts_kernel|setflags|gram_fit|rows_kernel|alpha_reduce|predict3|route_kernel|mix_kernel|solve3|lauum3|gather_kernel" -c 600 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
summ() {   # kernel, rep
  { echo "# ncu --set full --clock-control none --import-source on -k regex:$1 (one launch of: $3)"; python tools/ncu_summary.py $2 12; echo; echo "# stall samples by CUDA source line"; python tools/ncu_lines.py $2 14; } > gpurun_out/ncu_full_$1_$TAG.txt 2>&1
}
# INT8 split path (default): the block-product kernel (4 launches per evaluation: -s 13 = 2nd launch of the 4th evaluation = A22 -= L21 L21^T),
# the slicing kernel (pass 1 of the L21 slicing) and the fused FP64 tile launch (factorisation + inverse tiles of the first diagonal ranges)
ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 13 -c 1 -f -o /tmp/prof/oz_gemm_kernel $CMD > gpurun_out/ncu_run_oz_gemm_kernel_$TAG.log 2>&1
echo "full oz_gemm rc=$?"; summ oz_gemm_kernel /tmp/prof/oz_gemm_kernel.ncu-rep "$CMD"
ncu --set full --clock-control none --import-source on -k regex:slice_kernel -s 27 -c 1 -f -o /tmp/prof/oz_slice_kernel $CMD > gpurun_out/ncu_run_oz_slice_kernel_$TAG.log 2>&1
echo "full oz_slice rc=$?"; summ oz_slice_kernel /tmp/prof/oz_slice_kernel.ncu-rep "$CMD"
ncu --set full --clock-control none --import-source on -k regex:eval2_kernel -s 6 -c 1 -f -o /tmp/prof/eval2_kernel $CMD > gpurun_out/ncu_run_eval2_kernel_$TAG.log 2>&1
echo "full eval2 rc=$?"; summ eval2_kernel /tmp/prof/eval2_kernel.ncu-rep "$CMD"
ncu --set full --clock-control none --import-source on -k regex:gram_fit_kernel -s 3 -c 1 -f -o /tmp/prof/gram_fit_kernel $CMD > gpurun_out/ncu_run_gram_fit_kernel_$TAG.log 2>&1
echo "full gram_fit rc=$?"; summ gram_fit_kernel /tmp/prof/gram_fit_kernel.ncu-rep "$CMD"
# FP64 tile pipelines alone (DSMGP_OZAKI=0): the kernels of the first half of the round
for K in potrf2_kernel trtri3_kernel; do
  DSMGP_OZAKI=0 ncu --set full --clock-control none --import-source on -k regex:$K -s 3 -c 1 -f -o /tmp/prof/$K $CMD > gpurun_out/ncu_run_${K}_$TAG.log 2>&1
  echo "full $K rc=$?"; summ $K /tmp/prof/$K.ncu-rep "DSMGP_OZAKI=0 $CMD"
done
ncu --set full --clock-control none --import-source on -k regex:predict3_kernel -s 1 -c 1 -f -o /tmp/prof/predict3_kernel $CMD > gpurun_out/ncu_run_predict3_$TAG.log 2>&1
echo "full predict3 rc=$?"; summ predict3_kernel /tmp/prof/predict3_kernel.ncu-rep "$CMD (40,000 test points, device path)"
$CMDM > gpurun_out/plain_math_$TAG.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:lauum3_kernel -s 3 -c 1 -f -o /tmp/prof/lauum3_kernel $CMDM > gpurun_out/ncu_run_lauum3_$TAG.log 2>&1
echo "full lauum3 rc=$?"; summ lauum3_kernel /tmp/prof/lauum3_kernel.ncu-rep "$CMDM"
# DRAM traffic of the inverse with the G x G task grouping (L2 reuse by scheduling)
for G in 1 4; do
  DSMGP_OZAKI=0 DSMGP_TRTRI_GROUP=$G $CMD > /dev/null 2>&1 && DSMGP_OZAKI=0 DSMGP_TRTRI_GROUP=$G ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sector_hit_rate.pct --clock-control none -k regex:trtri3_kernel -s 3 -c 1 --csv --log-file gpurun_out/trtri3_group${G}_$TAG.csv $CMD > /dev/null 2>&1
  echo "trtri3 group $G rc=$?"
done
ls -la gpurun_out/ | tail -25
