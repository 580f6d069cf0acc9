#!/bin/bash
# round 2, first GPU call: parity suite, sweep of the instantiated tuning points, driver-shaped bench lines, launch list
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/tests.log
timeout 300 python scripts/sweep_fused.py 1536 cs > gpurun_out/sweep_cs.log 2>&1
timeout 200 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_20.json 2> gpurun_out/bench_20.err
PYCS_SPLIT=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench_20_split.json 2> gpurun_out/bench_20_split.err
timeout 200 python bench.py --steps 2000 --warmup 20 --no-cpu > gpurun_out/bench_2000.json 2> gpurun_out/bench_2000.err
PYCS_SPLIT=1 timeout 200 python bench.py --steps 2000 --warmup 20 --no-cpu > gpurun_out/bench_2000_split.json 2> gpurun_out/bench_2000_split.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ncu_bench.log 2>&1
tail -3 gpurun_out/tests.log
cat gpurun_out/sweep_cs.log | tail -12
cat gpurun_out/bench_20.json | cut -c1-600
