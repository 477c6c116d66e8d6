#!/bin/bash
# Round 2, eleventh GPU call: the index builder on a repeat-rich 1 Gb reference, the search tests incl. the new repeat-rich case
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02k}
timeout 900 python -m pytest tests/test_gpu_search.py tests/test_gpu_cli.py tests/test_gpu_zy_pipeline.py -q 2>&1 | tail -4 | tee gpurun_out/gpu_tests_$TAG.txt
timeout 900 python tools/bench_index_repeats.py --mb 200 2>&1 | tail -1 | tee gpurun_out/index_repeats_200_$TAG.txt
timeout 1500 python tools/bench_index_repeats.py --mb 1000 2>&1 | tail -1 | tee gpurun_out/index_repeats_1000_$TAG.txt
