#!/bin/bash
# Round 2, twenty-ninth GPU call: the reader in view mode (no copy of unsearched reads) with the three-stage pipeline, CLI tests
set -x
mkdir -p gpurun_out
timeout 1500 python tools/bench_bamread.py --records 30000 --repeat 48 --gpu-inflate 2>&1 | tail -1 | tee gpurun_out/bamread_r03c.txt
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_zy_pipeline.py -x -q 2>&1 | tail -3 | tee gpurun_out/cli_r03c.txt
