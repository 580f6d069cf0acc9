#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/tests2.log 2>&1
echo "tests rc=$?" >> gpurun_out/tests2.log
tail -5 gpurun_out/tests2.log
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench2_20.json 2> gpurun_out/bench2_20.err
PYCS_GRAPH=0 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench2_20_nograph.json 2> gpurun_out/bench2_20_nograph.err
PYCS_SPLIT=1 timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/bench2_20_split.json 2> gpurun_out/bench2_20_split.err
timeout 300 python scripts/bench_configs.py 400 > gpurun_out/configs2.jsonl 2> gpurun_out/configs2.err
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_case.py > gpurun_out/memcheck.log 2>&1
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_case.py > gpurun_out/racecheck.log 2>&1
tail -3 gpurun_out/memcheck.log gpurun_out/racecheck.log
cut -c1-300 gpurun_out/bench2_20.json
