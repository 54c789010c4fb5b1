#!/bin/bash
# Profiling recipe (B200_PROFILING.md). Run on the GPU box:  gpurun -- 'bash profiles/run_ncu.sh r01b f16x3'
# Produces gpurun_out/<tag>_launches.csv (per-launch device time, serialised/cold-cache: compare SHARES)
# and, with a third argument (kernel regex), gpurun_out/<tag>_<name>.ncu-rep (--set full on that kernel).
TAG=${1:-r01}
PREC=${2:-f16x3}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --samples 256 --no-cpu-baseline --precision $PREC"
ncu --metrics gpu__time_duration.sum --clock-control none -s ${NCU_SKIP:-310} -c ${NCU_COUNT:-420} --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
if [ -n "$3" ]; then
  ncu --set full --clock-control none --import-source on -k regex:$3 -s ${4:-60} -c ${5:-3} \
      -o gpurun_out/${TAG}_${6:-top} -f $CMD > gpurun_out/${TAG}_${6:-top}.log 2>&1
fi
ls -la gpurun_out
