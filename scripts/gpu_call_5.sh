#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do PYCS_ISSUE=$v timeout 120 python scripts/time_kernel.py >> gpurun_out/issue_variants.log 2>&1; done
cat gpurun_out/issue_variants.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/tests5.log 2>&1
echo "tests rc=$?" >> gpurun_out/tests5.log
tail -8 gpurun_out/tests5.log
timeout 300 python scripts/bench_configs.py 400 > gpurun_out/configs5.jsonl 2> gpurun_out/configs5.err
cut -c1-200 gpurun_out/configs5.jsonl
