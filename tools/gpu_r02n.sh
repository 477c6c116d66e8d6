#!/bin/bash
# Round 2, fourteenth GPU call: quick check of the pinned record arrays / flat host sweep, cluster tests
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02n}
timeout 600 python -m pytest tests/test_gpu_cluster_call.py tests/test_cluster_cpu.py tests/test_gpu_call.py -q 2>&1 | tail -3
timeout 900 python bench.py --no-config2 --no-cpu-baseline --no-call-stage 2>gpurun_out/bench_$TAG.err > gpurun_out/bench_$TAG.txt
python tools/bench_brief.py gpurun_out/bench_$TAG.txt; tail -2 gpurun_out/bench_$TAG.err
