#!/bin/bash
# programmatic dependent launches in the serial step: timing (graph / launched, PDL on / off) and parity
mkdir -p gpurun_out
L=gpurun_out/pdl_sweep.log
: > $L
for g in 1 0; do for p in 0 1; do
  echo "== PYCS_GRAPH=$g PYCS_PDL=$p" >> $L
  PYCS_GRAPH=$g PYCS_PDL=$p timeout 120 python bench.py --quick --steps 400 --warmup 20 2>>gpurun_out/pdl.err | cut -c1-120 >> $L
done; done
cat $L
for p in 0 1; do
  echo "== configs PYCS_PDL=$p" >> $L
  PYCS_PDL=$p timeout 200 python scripts/bench_configs.py 400 2>>gpurun_out/pdl.err | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print(d['config'][:40], d['ms_per_step'])" >> $L
done
tail -16 $L
timeout 600 python -m pytest tests/test_gpu_parity.py -q -x -k "fused or config1 or config2 or config3_n768_deformational or separable or basis" > gpurun_out/pdl_tests.log 2>&1
tail -3 gpurun_out/pdl_tests.log
PYCS_GRAPH=0 timeout 300 python -m pytest tests/test_gpu_parity.py -q -x -k "fused_matches or separable or basis" > gpurun_out/pdl_tests_nograph.log 2>&1
tail -3 gpurun_out/pdl_tests_nograph.log
