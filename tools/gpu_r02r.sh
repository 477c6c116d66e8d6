#!/bin/bash
# Round 2, eighteenth GPU call: the warp-per-member inflate kernel -- parity (both kernels), GB/s next to the thread kernel,
# and the BAM reader with --gpu-inflate next to the host one on a 3.3 GB file.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_inflate.py tests/test_gpu_cli.py -x -q -s 2>&1 | tail -15 | tee gpurun_out/cli_r02r.txt
for mb in 256 1024; do
  timeout 900 python tools/bench_inflate.py --mb $mb 2>&1 | tail -1
  SVB_INFLATE_KERNEL=thread timeout 900 python tools/bench_inflate.py --mb $mb 2>&1 | tail -1
done | tee gpurun_out/inflate_r02r.txt
timeout 900 python tools/bench_inflate.py --mb 512 --quals 2>&1 | tail -1 | tee -a gpurun_out/inflate_r02r.txt
timeout 1200 python tools/bench_bamread.py --records 30000 --repeat 24 --gpu-inflate 2>&1 | tail -1 | tee gpurun_out/bamread_r02r.txt
