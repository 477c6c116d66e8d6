#!/bin/bash
# Round 2, first GPU call: the GPU suite, the skipped streamed 2-bit test, inflate timing, POA / ksw2 variant sweep.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02a}
nproc; lscpu | grep -E "Model name|Socket|NUMA|^CPU\(s\)" ; free -g | head -2
timeout 900 python -m pytest tests -q -m gpu -s 2>&1 | tail -25 | tee gpurun_out/gpu_tests_$TAG.txt
SVB_TEST_STREAM_PACK2=1 SVB_SEARCH_STATS=1 timeout 200 python -m pytest tests/test_gpu_zz_stream_pack2.py -q -s 2>&1 | tail -15 | tee gpurun_out/stream_pack2_$TAG.txt
for q in "" "--quals"; do timeout 300 python tools/bench_inflate.py --mb 256 $q 2>&1 | tail -1; done | tee gpurun_out/inflate_$TAG.txt
N=12000
for v in 0 1 2 4 7 8 16 31 32 63; do
  echo "== variant $v"
  SVB_POA_VARIANT=$v SVB_POA_TIMING=1 timeout 300 python tools/bench_call.py --clusters $N --pairs 0 --cpu-seconds 0.5 2>&1 | \
    grep -E "k_poa phases|\"kernel\"|clusters_per_s|GCUPS" | cut -c1-400
done | tee gpurun_out/poa_variants_$TAG.txt
for g in 16 8; do for v in 0 63; do for k in 1 6; do
  echo "== variant $v, $g lanes per cluster, $k launch buckets"
  SVB_POA_BUCKETS=$k SVB_POA_GROUP=$g SVB_POA_VARIANT=$v SVB_POA_TIMING=1 timeout 300 python tools/bench_call.py --clusters $N --pairs 0 --cpu-seconds 0.5 2>&1 | \
    grep -E "k_poa phases|\"kernel\"|clusters_per_s|GCUPS" | cut -c1-400
done; done; done | tee -a gpurun_out/poa_variants_$TAG.txt
for v in 0 1 2 3; do
  echo "== ksw variant $v"
  SVB_KSW_VARIANT=$v timeout 300 python tools/bench_call.py --clusters 0 --pairs 20000 --max-len 3000 --cpu-seconds 0.5 2>&1 | grep k_ksw | cut -c1-300
  SVB_KSW_VARIANT=$v timeout 300 python tools/bench_call.py --clusters 0 --pairs 20000 --cpu-seconds 0.5 2>&1 | grep k_ksw | cut -c1-300
done | tee gpurun_out/ksw_variants_$TAG.txt
