#!/bin/bash
set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke
python tools/rank_sweep.py 2>&1 | tee gpurun_out/rank_sweep.txt | tail -8
timeout 600 python bench.py --ref-bp 200000000 --reads 50000 --steps 2 --warmup 1 --cpu-seconds 3 2>&1 | tee gpurun_out/bench_small.txt | tail -20
timeout 1500 python bench.py --steps 2 --warmup 3 2>&1 | tee gpurun_out/bench_full.txt | tail -20
