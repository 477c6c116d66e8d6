#!/bin/bash
# Round 2, nineteenth GPU call: warp inflate kernel with queued matches (batched copies), the reader with fixed pinned buffers.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_zz_inflate.py tests/test_gpu_cli.py -x -q 2>&1 | tail -4 | tee gpurun_out/cli_r02s.txt
for mb in 256 1024; do
  timeout 900 python tools/bench_inflate.py --mb $mb 2>&1 | tail -1
  SVB_INFLATE_OCC12=1 timeout 900 python tools/bench_inflate.py --mb $mb 2>&1 | tail -1
done | tee gpurun_out/inflate_r02s.txt
timeout 900 python tools/bench_inflate.py --mb 512 --quals 2>&1 | tail -1 | tee -a gpurun_out/inflate_r02s.txt
timeout 1200 python tools/bench_bamread.py --records 30000 --repeat 24 --gpu-inflate 2>&1 | tail -1 | tee gpurun_out/bamread_r02s.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_bgzf_inflate_warp -c 2 -o gpurun_out/r02s_inflate python tools/bench_inflate.py --mb 256 > gpurun_out/ncu_r02s.log 2>&1
ls -la gpurun_out/*.ncu-rep
