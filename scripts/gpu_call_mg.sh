#!/bin/bash
# usage: gpu_call_mg.sh NGPUS
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 tests/mgpu_worker.py > gpurun_out/mg${N}_worker.log 2>&1
echo "worker rc=$?" >> gpurun_out/mg${N}_worker.log
tail -4 gpurun_out/mg${N}_worker.log
for S in 20 400; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29642 bench.py --gpus $N --steps $S --warmup 5 > gpurun_out/mg${N}_bench_$S.json 2> gpurun_out/mg${N}_bench_$S.err
done
PYCS_GRAPH=0 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus $N --steps 400 --warmup 5 > gpurun_out/mg${N}_bench_400_nograph.json 2> gpurun_out/mg${N}_bench_400_nograph.err
for f in gpurun_out/mg${N}_bench_*.json; do echo $f; cut -c1-330 $f; done
tail -3 gpurun_out/mg${N}_bench_20.err
