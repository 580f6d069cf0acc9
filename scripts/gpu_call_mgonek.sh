#!/bin/bash
# usage: gpu_call_mgonek.sh NGPUS [full]: one-kernel step on several GPUs -- parity worker, quick timings of the shapes, driver-shaped bench
N=${1:-2}
mkdir -p gpurun_out
W=gpurun_out/mgo${N}_worker.log
if [ "$2" = "full" ]; then A=""; else A="1536"; fi
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tests/mgpu_worker.py $A > $W 2>&1
echo "worker rc=$?" >> $W
grep -v "^rank [1-9]" $W | tail -12
q() { env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus $N --steps 300 --warmup 10 --quick 2>gpurun_out/mgo${N}_q.err | grep quick | cut -c1-100; }
( echo "onekernel (graph):"; q PYCS_X=0; echo "onekernel launched:"; q PYCS_GRAPH=0; echo "split (two streams, launched):"; q PYCS_ONEKERNEL=0 ) | tee gpurun_out/mgo${N}_quick.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mgo${N}_bench_20.json 2> gpurun_out/mgo${N}_bench_20.err
cut -c1-330 gpurun_out/mgo${N}_bench_20.json
