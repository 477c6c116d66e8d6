#!/bin/bash
# Round 2, thirty-second GPU call: the BAM loader of `search` on the device (svb_bamstream_*): CLI parity tests, then the loader alone
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -25 | tee gpurun_out/cli_r03g.txt
timeout 1500 python tools/bench_bamread.py --records 30000 --repeat 24 --gpu-inflate 2>&1 | tail -3 | tee gpurun_out/bamread_r03g.txt
