#!/bin/bash
# Round 2, twenty-fourth GPU call: the library built with --default-stream per-thread -- whole GPU suite, smoke, the line
set -x
mkdir -p gpurun_out
TAG=r02x
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests_$TAG.txt
python __graft_entry__.py smoke 2>&1 | tail -1
timeout 900 python bench.py --steps 10 --warmup 3 --no-config2 --no-cpu-baseline --no-call-stage 2> gpurun_out/bench_$TAG.err > gpurun_out/bench_$TAG.txt
python tools/bench_brief.py gpurun_out/bench_$TAG.txt
