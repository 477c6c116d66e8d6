#!/bin/bash
# Round 2, third GPU call: POA variants (parity + timing on the config-4 shape and inside the config-3 step), pooled workspaces
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02c}
timeout 900 python -m pytest tests/test_gpu_poa.py tests/test_gpu_zz_poa_variants.py tests/test_gpu_cluster_call.py tests/test_gpu_call.py tests/test_gpu_ksw.py -q -x -s 2>&1 | tail -8 | tee gpurun_out/gpu_tests_$TAG.txt
for v in 0 7 71 135 199 263 455 487; do
  echo "== variant $v"
  SVB_POA_VARIANT=$v SVB_POA_TIMING=1 timeout 300 python tools/bench_call.py --clusters 12000 --pairs 0 --cpu-seconds 0.5 2>&1 | \
    grep -E "k_poa phases|\"kernel\"" | cut -c1-400
done | tee gpurun_out/poa_variants_$TAG.txt
for v in 0 455 487; do
  echo "== bench, variant $v"
  SVB_POA_VARIANT=$v SVB_POA_TIMING=1 timeout 900 python bench.py --no-config2 --no-cpu-baseline --no-call-stage 2>gpurun_out/bench_v${v}_$TAG.err | tee gpurun_out/bench_v${v}_$TAG.txt | \
    python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('value %.0f reads/s  step %.1f ms  stages %s' % (d['value'], d['ms_per_step'], json.dumps(d['stages_ms'])))
print('e2e %.0f reads/s  %s' % (d['e2e']['value'], json.dumps(d['e2e']['stages_ms'])))
print(json.dumps(d['parity_vs_planted']), json.dumps(d['kernels']['k_poa']))"
  tail -2 gpurun_out/bench_v${v}_$TAG.err
done
