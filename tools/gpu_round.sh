#!/bin/bash
# full GPU pass: tests, smoke, bench (+reference arm), ncu evidence
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python __graft_entry__.py smoke
timeout 1200 python bench.py 2>&1 | tee gpurun_out/bench_full.txt | tail -3
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tee gpurun_out/bench_reference.txt | tail -2
TAG=${TAG:-r01} READS=1000000 bash tools/gpu_profile.sh
