#!/bin/bash
# One point of the training scaling curves (strong: global batch 64; weak: 64 per GPU) on N GPUs of this box.
#   tools/run_train_scale.sh N   -> gpurun_out/r02_train_scale_N.jsonl (two JSON lines)
N=${1:-1}
OUT=gpurun_out/r02_train_scale_${N}.jsonl
: > $OUT
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
$L bench.py --gpus $N --workload train --steps 10 --warmup 3 --no-cpu-baseline --global-batch 64 2>>gpurun_out/r02_train_scale_${N}.err >> $OUT
$L bench.py --gpus $N --workload train --steps 10 --warmup 3 --no-cpu-baseline --global-batch $((64*N)) 2>>gpurun_out/r02_train_scale_${N}.err >> $OUT
python - <<PY
import json
for l in open("$OUT"):
    d=json.loads(l); print($N, d["config"]["global_batch"], round(d["value"],2), "steps/s", round(d["ms_per_step"],1), "ms", d["allreduce"]["ms_per_step"], d["cuda_graph"])
PY
