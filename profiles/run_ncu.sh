#!/bin/bash
# Profiling recipe (B200_PROFILING.md). Run on the GPU box:  gpurun -- 'bash profiles/run_ncu.sh r01'
# Produces gpurun_out/<tag>_launches.csv (per-launch device time, serialised/cold-cache: compare SHARES)
# and gpurun_out/<tag>_conv.ncu-rep / <tag>_pointwise.ncu-rep (--set full on the top kernels).
TAG=${1:-r01}
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 3 --samples 256 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 600 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:conv3x3 -s 260 -c 6 \
    -o gpurun_out/${TAG}_conv -f $CMD > gpurun_out/${TAG}_conv.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flow_pointwise -s 60 -c 3 \
    -o gpurun_out/${TAG}_pointwise -f $CMD > gpurun_out/${TAG}_pointwise.log 2>&1
ls -la gpurun_out
