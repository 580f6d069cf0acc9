#!/bin/bash
# march variants of the step kernel (PYCS_VARIANT, csrc/fused2b.cu): CUDA-event timing at N=1536 + parity of the combined ones
mkdir -p gpurun_out
L=gpurun_out/var1_sweep.log
: > $L
for v in 0 1 2 3 5 8 16 24 11 27 0 1 3; do PYCS_VARIANT=$v timeout 120 python scripts/time_kernel.py >> $L 2>&1; done
cat $L
for v in 27 24 5; do
  echo "== parity PYCS_VARIANT=$v" >> gpurun_out/var1_parity.log
  PYCS_VARIANT=$v timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_matches_operator_path and default" >> gpurun_out/var1_parity.log 2>&1
done
grep -E "passed|failed|error|==" gpurun_out/var1_parity.log
