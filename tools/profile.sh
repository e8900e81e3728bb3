#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then one full capture per top kernel.
# usage: tools/profile.sh <tag> [workload]
set -u
TAG=${1:-r01}
WL=${2:-cfg3}
CMD="python bench.py --workload $WL --steps 1 --warmup 3 --no-cpu-baseline"
mkdir -p gpurun_out
$CMD > gpurun_out/plain_$TAG.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -1 gpurun_out/plain_$TAG.log | head -c 600; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_list_$TAG.log 2>&1
echo "launch list rc=$?"
for K in ${KERNELS:-trtri_kernel potrf_panel_kernel}; do
  SKIP=${SKIP:-3}
  [ "$K" = "potrf_panel_kernel" ] && SKIP=${SKIP_PANEL:-127}
  ncu --set full --clock-control none --import-source on -k regex:$K -s $SKIP -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > gpurun_out/ncu_full_${K}_$TAG.log 2>&1
  echo "full $K rc=$?"
done
ls -la gpurun_out/
