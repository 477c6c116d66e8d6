#!/bin/bash
# Round 2, fifth GPU call: the whole GPU suite (N-run closed form, text pad, CLI on the new C ABI), the reference arm at full size
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02e}
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | tail -12 | tee gpurun_out/gpu_tests_$TAG.txt
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.txt ) 2>&1 | tail -4
cut -c1-1800 gpurun_out/bench_ref_$TAG.txt; tail -3 gpurun_out/bench_ref_$TAG.err
( time timeout 1500 python bench.py --steps 5 --warmup 3 2>gpurun_out/bench_full_$TAG.err > gpurun_out/bench_full_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_full_$TAG.txt; tail -5 gpurun_out/bench_full_$TAG.err
