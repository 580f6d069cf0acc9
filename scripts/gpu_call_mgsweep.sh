#!/bin/bash
# usage: gpu_call_mgsweep.sh NGPUS  -- tuning sweep of the split step's boundary shapes
N=${1:-2}
mkdir -p gpurun_out
OUT=gpurun_out/mg${N}_sweep.log
: > $OUT
run() {
  echo "== $*" >> $OUT
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus $N --steps 300 --warmup 5 --quick 2>> gpurun_out/mg${N}_sweep.err | grep quick >> $OUT
}
run PYCS_X=default
run PYCS_SPLIT_EDGE_ROWS=16
run PYCS_SPLIT_EDGE_ROWS=32 PYCS_SPLIT_BAND=8
run PYCS_SPLIT_EDGE_ROWS=48
run PYCS_SPLIT_BAND=24 PYCS_SPLIT_EDGE_ROWS=24
run PYCS_GRAPH=1
PYCS_STEP_PROFILE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 300 --warmup 5 --quick > gpurun_out/mg${N}_profile.log 2>&1
grep "step profile" gpurun_out/mg${N}_profile.log >> $OUT
cut -c1-200 $OUT
