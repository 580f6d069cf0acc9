#!/bin/bash
# one-kernel step (GH = 2: ghost fill inside the step kernel) on one GPU: parity, then timing against the serial step
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_step_shapes and not 1536" > gpurun_out/onek_tests.log 2>&1
tail -4 gpurun_out/onek_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_matches or separable or basis or config2 or (fused_step_shapes and 1536 and onekernel)" > gpurun_out/onek_tests2.log 2>&1
tail -4 gpurun_out/onek_tests2.log
L=gpurun_out/onek_sweep.log
: > $L
for o in 0 1 0 1; do
  echo "== PYCS_ONEKERNEL=$o" >> $L
  PYCS_ONEKERNEL=$o timeout 120 python bench.py --quick --steps 400 --warmup 20 2>>gpurun_out/onek.err | cut -c1-120 >> $L
done
for o in 0 1; do
  echo "== configs PYCS_ONEKERNEL=$o" >> $L
  PYCS_ONEKERNEL=$o timeout 200 python scripts/bench_configs.py 400 2>>gpurun_out/onek.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'][:40], d['ms_per_step'], d['launches_per_step'])" >> $L
done
cat $L
timeout 300 compute-sanitizer --tool memcheck python scripts/sanitize_case.py > gpurun_out/onek_memcheck.log 2>&1; tail -3 gpurun_out/onek_memcheck.log
