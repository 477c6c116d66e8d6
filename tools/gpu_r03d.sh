#!/bin/bash
# Round 2, thirtieth GPU call: whole GPU suite, smoke, and the round's line exactly as the driver runs it (both arms), on the final code
set -x
mkdir -p gpurun_out
TAG=r03d
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests_$TAG.txt
python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 1500 python bench.py 2>gpurun_out/bench_full_$TAG.err > gpurun_out/bench_full_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_full_$TAG.txt
( time timeout 900 python bench.py --impl reference 2>gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.txt ) 2>&1 | tail -4
cut -c1-300 gpurun_out/bench_ref_$TAG.txt
