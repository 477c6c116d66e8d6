#!/bin/bash
# Round 2, twentieth GPU call: the BAM reader with --gpu-inflate on a 6.6 GB file (pinning time apart), and the CLI end to end
# (smooth | index | search | call) with and without --gpu-inflate on the same files.
set -x
mkdir -p gpurun_out
timeout 1500 python tools/bench_bamread.py --records 30000 --repeat 48 --gpu-inflate 2>&1 | tail -1 | tee gpurun_out/bamread_r02t.txt
timeout 1500 python tools/bench_e2e.py --raw --ref-bp 50000000 --gpu-inflate --threads 16 2> gpurun_out/e2e_r02t.err | tail -1 | tee gpurun_out/e2e_r02t.txt
tail -5 gpurun_out/e2e_r02t.err
