#!/bin/bash
# usage: gpu_call_mgfinal.sh NGPUS : parity worker (N=1536 only) + driver-shaped bench + quick variants
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tests/mgpu_worker.py 1536 > gpurun_out/mgf${N}_worker.log 2>&1
echo "worker rc=$?" >> gpurun_out/mgf${N}_worker.log
tail -2 gpurun_out/mgf${N}_worker.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mgf${N}_bench_20.json 2> gpurun_out/mgf${N}_bench_20.err
cut -c1-330 gpurun_out/mgf${N}_bench_20.json
q() { env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus $N --steps 300 --warmup 5 --quick 2>gpurun_out/mgf${N}_q.err | grep quick | cut -c1-100; }
( echo "default:"; q PYCS_X=0; echo "xkernel:"; q PYCS_MG_XKERNEL=1; echo "graph:"; q PYCS_GRAPH=1 ) | tee gpurun_out/mgf${N}_quick.log
