#!/bin/bash
# Profiling recipe (B200_PROFILING.md). Run on the GPU box:  gpurun -- 'bash profiles/run_ncu.sh r01f f16x3'
#   gpurun_out/<tag>_launches.csv : every launch of one timed bench step with its device time
#                                   (serialised, cold cache: compare SHARES with bench.py's event timing)
#   gpurun_out/<tag>_step.ncu-rep : --set full on one level-0 fused flow step (flow_step_f16_kernel<12,..>)
#   gpurun_out/<tag>_gate.ncu-rep : --set full on the level-0 ConvLSTM gate conv (conv3x3_f16_kernel)
TAG=${1:-r01}
PREC=${2:-f16x3}
S=${3:-1024}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --samples $S --no-cpu-baseline --no-train --precision $PREC"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-310} -c ${NCU_COUNT:-420} --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
if [ -z "$NCU_LIST_ONLY" ]; then
  ncu --set full --clock-control none --import-source on -k regex:flow_step_f16_kernel -s 35 -c 1 \
      -o gpurun_out/${TAG}_step -f $CMD > gpurun_out/${TAG}_step.log 2>&1
  ncu --set full --clock-control none --import-source on -k regex:conv3x3_f16_kernel -s 11 -c 1 \
      -o gpurun_out/${TAG}_gate -f $CMD > gpurun_out/${TAG}_gate.log 2>&1
fi
ls -la gpurun_out
