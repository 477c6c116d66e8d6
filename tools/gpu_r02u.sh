#!/bin/bash
# Round 2, twenty-first GPU call: the reader with its three stages side by side (read | inflate | parse), then the whole GPU
# suite and the round's line (both arms) on the current code.
set -x
mkdir -p gpurun_out
TAG=r02u
timeout 1500 python tools/bench_bamread.py --records 30000 --repeat 48 --gpu-inflate 2>&1 | tail -1 | tee gpurun_out/bamread_$TAG.txt
timeout 1800 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests_$TAG.txt
python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 1500 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_full_$TAG.err > gpurun_out/bench_full_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_full_$TAG.txt
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.txt ) 2>&1 | tail -4
cut -c1-300 gpurun_out/bench_ref_$TAG.txt
