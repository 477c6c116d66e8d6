#!/bin/bash
# Round 2, fourth GPU call: the CLI on svb_cluster_batch + svb_call_batch (call / pipeline / clipped / cluster tests), ILP POA rows
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02d}
timeout 1200 python -m pytest tests/test_gpu_poa.py tests/test_gpu_zz_poa_variants.py tests/test_gpu_cluster_call.py tests/test_gpu_call.py tests/test_gpu_cli.py tests/test_gpu_zy_pipeline.py tests/test_gpu_zz_clipped.py tests/test_cluster_cpu.py tests/test_clipper_cpu.py tests/test_gpu_search.py -q -x -s 2>&1 | tail -12 | tee gpurun_out/gpu_tests_$TAG.txt
for v in 455 967 1479 3015 3527; do
  echo "== variant $v"
  SVB_POA_VARIANT=$v SVB_POA_TIMING=1 timeout 300 python tools/bench_call.py --clusters 12000 --pairs 0 --cpu-seconds 0.5 2>&1 | \
    grep -E "k_poa phases|\"kernel\"" | tail -2 | cut -c1-400
done | tee gpurun_out/poa_variants_$TAG.txt
for v in 967 1479 3015 3527; do
  echo "== bench, variant $v"
  SVB_POA_VARIANT=$v timeout 900 python bench.py --no-config2 --no-cpu-baseline --no-call-stage 2>gpurun_out/bench_v${v}_$TAG.err > gpurun_out/bench_v${v}_$TAG.txt
  python tools/bench_brief.py gpurun_out/bench_v${v}_$TAG.txt
done
