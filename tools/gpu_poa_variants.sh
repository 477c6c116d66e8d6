#!/bin/bash
# POA kernel variants (SVB_POA_VARIANT, svdss_b200/csrc/poa_kernel.cuh) side by side on the config-4 shape:
# parity first, then clusters/s + GCUPS + phase shares per variant.  TAG names the outputs.
mkdir -p gpurun_out
TAG=${TAG:-r02}
N=${N:-12000}
timeout 900 python -m pytest tests/test_gpu_poa.py tests/test_gpu_zz_poa_variants.py -x -q -m gpu -s 2>&1 | tail -6
for v in 0 1 2 3 4 7 8 15 16 31 32 39 63; do  # POA
  echo "== variant $v"
  SVB_POA_VARIANT=$v SVB_POA_TIMING=1 timeout 600 python tools/bench_call.py --clusters $N --pairs 0 --cpu-seconds 0.5 2>&1 | \
    grep -E "k_poa phases|\"kernel\"|clusters_per_s|GCUPS" | cut -c1-400
done | tee gpurun_out/poa_variants_$TAG.txt
for g in 16 8; do for v in 0 7 31 63; do for k in 1 6; do
  echo "== variant $v, $g lanes per cluster, $k launch buckets"
  SVB_POA_BUCKETS=$k SVB_POA_GROUP=$g SVB_POA_VARIANT=$v SVB_POA_TIMING=1 timeout 600 python tools/bench_call.py --clusters $N --pairs 0 --cpu-seconds 0.5 2>&1 | \
    grep -E "k_poa phases|\"kernel\"|clusters_per_s|GCUPS" | cut -c1-400
done; done; done | tee -a gpurun_out/poa_variants_$TAG.txt
# ksw2 backtrack variant on a pipeline-like shape (pairs up to 3 kb) and on the config-5 shape
for v in 0 1 2 3; do
  echo "== ksw variant $v"
  SVB_KSW_VARIANT=$v timeout 600 python tools/bench_call.py --clusters 0 --pairs 20000 --max-len 3000 --cpu-seconds 0.5 2>&1 | grep k_ksw | cut -c1-300
  SVB_KSW_VARIANT=$v timeout 600 python tools/bench_call.py --clusters 0 --pairs 20000 --cpu-seconds 0.5 2>&1 | grep k_ksw | cut -c1-300
done | tee -a gpurun_out/poa_variants_$TAG.txt
