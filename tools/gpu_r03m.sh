#!/bin/bash
# Round 2, last full validation: whole GPU suite, smoke, the line as the driver runs it (both arms)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r03m}
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests_$TAG.txt
python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 1500 python bench.py 2>gpurun_out/bench_full_$TAG.err > gpurun_out/bench_full_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_full_$TAG.txt 2>/dev/null | cut -c1-400
( time timeout 900 python bench.py --impl reference 2>gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.txt ) 2>&1 | tail -4
cut -c1-200 gpurun_out/bench_ref_$TAG.txt
