#!/bin/bash
# Round 2, seventeenth GPU call: members-per-warp sweep of the device inflate, the reader with --gpu-inflate next to the host one,
# and the CLI parity test of the wiring.
set -x
mkdir -p gpurun_out
for mb in 256 1024; do timeout 900 python tools/bench_inflate.py --mb $mb --mpw 1,2,4,8,16,32 2>&1 | tail -1; done | tee gpurun_out/inflate_r02q.txt
timeout 900 python tools/bench_bamread.py --records 30000 --repeat 8 --gpu-inflate 2>&1 | tail -1 | tee gpurun_out/bamread_r02q.txt
timeout 900 python -m pytest tests/test_gpu_cli.py tests/test_gpu_zz_inflate.py -x -q 2>&1 | tail -5 | tee gpurun_out/cli_r02q.txt
