#!/bin/bash
# hoisted CFL numbers (32) and the barrier probes (64, 128: wrong results, timing only)
mkdir -p gpurun_out
L=gpurun_out/var2_sweep.log
: > $L
for v in 0 32 64 128 96 0 32; do PYCS_VARIANT=$v timeout 120 python scripts/time_kernel.py >> $L 2>&1; done
cat $L
PYCS_VARIANT=32 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_matches_operator_path and default" 2>&1 | tail -2
