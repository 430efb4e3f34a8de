#!/bin/bash
# Profiling recipe (B200_PROFILING.md): plain run first, then the launch list, then --set full
# captures of the top kernels.  Run under gpurun on ONE GPU; outputs land in gpurun_out/.
# usage: bash profiles/run_ncu.sh <tag> [batch]
TAG=${1:-r02}
BATCH=${2:-2048}
CMD="python bench.py --steps 1 --warmup 3 --batch $BATCH --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
$CMD > gpurun_out/ncu_plain_$TAG.json 2> gpurun_out/ncu_plain_$TAG.err || { echo "plain run failed"; tail -5 gpurun_out/ncu_plain_$TAG.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > /dev/null 2>&1
for K in lstm4_bwd_kernel lstm4_fwd_kernel gemm_f16_2sm_astat_kernel frontend_train_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:$K -s 6 -c 1 -f -o gpurun_out/prof_${K}_$TAG $CMD > /dev/null 2>&1
done
# the 13 CTA-pair GEMM launches of the timed step's backward pass (3 warm-up steps x 13 skipped): dW / dX traffic
ncu --set full --clock-control none -k regex:gemm_f16_2sm_kernel -s 39 -c 13 -f -o gpurun_out/prof_gemm_f16_2sm_kernel_$TAG $CMD > /dev/null 2>&1
ls -la gpurun_out/ | grep $TAG
