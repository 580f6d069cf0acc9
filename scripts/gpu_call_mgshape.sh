#!/bin/bash
# usage: gpu_call_mgshape.sh NGPUS R E : one-kernel step on several GPUs, shapes of the CTA table (band / edge chunk / interior chunk)
N=${1:-2}; R=${2:-86}; E=${3:-78}
mkdir -p gpurun_out
q() { echo "== $*"; env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus $N --steps 300 --warmup 10 --quick 2>gpurun_out/mgs${N}_q.err | grep quick | cut -c1-100; }
( q PYCS_X=0
  q PYCS_SPLIT_BAND=$R PYCS_SPLIT_EDGE_ROWS=$R PYCS_SPLIT_ROWS=$R
  q PYCS_SPLIT_BAND=$R PYCS_SPLIT_EDGE_ROWS=$E PYCS_SPLIT_ROWS=$R
  q PYCS_SPLIT_BAND=$E PYCS_SPLIT_EDGE_ROWS=$E PYCS_SPLIT_ROWS=$R
  q PYCS_SPLIT_BAND=$R PYCS_SPLIT_EDGE_ROWS=$R PYCS_SPLIT_ROWS=$R PYCS_GRAPH=0 PYCS_PDL=2
  q PYCS_ONEKERNEL=0 ) | tee gpurun_out/mgs${N}_shapes.log
