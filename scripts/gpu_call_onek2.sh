#!/bin/bash
# block-wide closing sum + early first copies; serial vs one-kernel step (graph / launched / launched + PDL) on one GPU
mkdir -p gpurun_out
L=gpurun_out/onek2_sweep.log
: > $L
timeout 120 python scripts/time_kernel.py >> $L 2>&1
q() { echo "== $*" >> $L; env "$@" timeout 120 python bench.py --quick --steps 400 --warmup 20 2>>gpurun_out/onek2.err | cut -c1-120 >> $L; }
q PYCS_ONEKERNEL=0
q PYCS_ONEKERNEL=1
q PYCS_ONEKERNEL=1 PYCS_GRAPH=0
q PYCS_ONEKERNEL=1 PYCS_GRAPH=0 PYCS_PDL=2
q PYCS_ONEKERNEL=0
q PYCS_ONEKERNEL=1 PYCS_GRAPH=0 PYCS_PDL=2
cat $L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "(fused_step_shapes and not 1536) or fused_matches or separable or basis" > gpurun_out/onek2_tests.log 2>&1
tail -3 gpurun_out/onek2_tests.log
PYCS_GRAPH=0 PYCS_PDL=2 timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_matches or separable or basis or config2" > gpurun_out/onek2_tests_pdl2.log 2>&1
tail -3 gpurun_out/onek2_tests_pdl2.log
