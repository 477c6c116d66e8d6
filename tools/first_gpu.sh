#!/bin/bash
# first GPU run: environment probe, tests, rank microbench for both block sizes
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
nproc; free -g | head -2; lscpu | grep -E "Model name|^CPU\(s\)|Flags" | cut -c1-300
python -m pytest tests -x -q -m gpu 2>&1 | tail -25
python tools/rank_sweep.py 2>&1 | tee gpurun_out/rank_sweep.txt | tail -30
