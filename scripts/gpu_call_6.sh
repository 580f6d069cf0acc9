#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests6.log 2>&1
echo "tests rc=$?" >> gpurun_out/tests6.log
tail -8 gpurun_out/tests6.log
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench6_20.json 2> gpurun_out/bench6_20.err
cut -c1-300 gpurun_out/bench6_20.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/ncu6.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fused2b -s 6 -c 1 -f -o gpurun_out/r2_v2b_final python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu6b.log 2>&1
python __graft_entry__.py smoke > gpurun_out/smoke6.log 2>&1; tail -3 gpurun_out/smoke6.log
