#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/onek3_sweep.log
: > $L
q() { echo "== $*" >> $L; env "$@" timeout 120 python bench.py --quick --steps 400 --warmup 20 2>>gpurun_out/onek3.err | cut -c1-120 >> $L; }
q PYCS_ONEKERNEL=0
q PYCS_ONEKERNEL=1
q PYCS_ONEKERNEL=0
q PYCS_ONEKERNEL=1
cat $L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "(fused_step_shapes and not 1536) or fused_matches or separable or basis" > gpurun_out/onek3_tests.log 2>&1
tail -3 gpurun_out/onek3_tests.log
