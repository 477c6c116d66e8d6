#!/bin/bash
# Round 2, ninth GPU call: blocked 4-columns-per-lane POA rows (BLK4), batched N-run closed form, search without the tail launch
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02i}
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_zz_poa_variants.py tests/test_gpu_poa.py tests/test_gpu_zz_stream_pack2.py -q -s 2>&1 | tail -8 | tee gpurun_out/gpu_tests_$TAG.txt
for v in 8647 10695; do
  echo "== variant $v"
  SVB_POA_VARIANT=$v SVB_POA_TIMING=1 timeout 300 python tools/bench_call.py --clusters 12000 --pairs 0 --cpu-seconds 0.5 2>&1 | \
    grep -E "k_poa phases|\"kernel\"" | tail -2 | cut -c1-400
done | tee gpurun_out/poa_variants_$TAG.txt
for v in 8647 10695; do
  echo "== bench, variant $v"
  SVB_POA_VARIANT=$v timeout 900 python bench.py --no-config2 --no-cpu-baseline --no-call-stage 2>gpurun_out/bench_v${v}_$TAG.err > gpurun_out/bench_v${v}_$TAG.txt
  python tools/bench_brief.py gpurun_out/bench_v${v}_$TAG.txt
done
echo "== bench, SVB_SEARCH_TAIL=0"
SVB_SEARCH_TAIL=0 timeout 900 python bench.py --no-config2 --no-cpu-baseline --no-call-stage 2>gpurun_out/bench_notail_$TAG.err > gpurun_out/bench_notail_$TAG.txt
python tools/bench_brief.py gpurun_out/bench_notail_$TAG.txt
