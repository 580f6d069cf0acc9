#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "basis or fused or config3 or lean or separable or lazily or smoke" > gpurun_out/tests4.log 2>&1
echo "tests rc=$?" >> gpurun_out/tests4.log
tail -12 gpurun_out/tests4.log
timeout 300 python scripts/bench_configs.py 400 > gpurun_out/configs4.jsonl 2> gpurun_out/configs4.err
cut -c1-330 gpurun_out/configs4.jsonl
