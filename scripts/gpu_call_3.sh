#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/tests3.log 2>&1
echo "tests rc=$?" >> gpurun_out/tests3.log
tail -6 gpurun_out/tests3.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench3_20.json 2> gpurun_out/bench3_20.err
cut -c1-400 gpurun_out/bench3_20.json; tail -3 gpurun_out/bench3_20.err
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused2b -s 6 -c 1 -f -o gpurun_out/r2_v2b python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu3.log 2>&1
ls -la gpurun_out/*.ncu-rep
