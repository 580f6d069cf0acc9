#!/bin/bash
# usage: gpu_call_mg8.sh NGPUS : parity worker (short) + driver-shaped bench + quick long run + profile
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tests/mgpu_worker.py 1536 > gpurun_out/mg${N}_worker.log 2>&1
echo "worker rc=$?" >> gpurun_out/mg${N}_worker.log
tail -3 gpurun_out/mg${N}_worker.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/mg${N}_bench_20.json 2> gpurun_out/mg${N}_bench_20.err
cut -c1-330 gpurun_out/mg${N}_bench_20.json
PYCS_STEP_PROFILE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29651 bench.py --gpus $N --steps 300 --warmup 5 --quick > gpurun_out/mg${N}_profile.log 2>&1
grep "pycs" gpurun_out/mg${N}_profile.log | head -12
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29652 bench.py --gpus $N --steps 300 --warmup 5 --quick 2>/dev/null | grep quick | cut -c1-120
