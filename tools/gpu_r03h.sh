#!/bin/bash
# Round 2, thirty-third GPU call: device BAM loader with the segmented record walk: CLI parity, lap times, loader alone
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_cli.py -x -q 2>&1 | tail -25 | tee gpurun_out/cli_r03h.txt
timeout 1500 python tools/bench_bamread.py --records 30000 --repeat 24 --gpu-inflate --laps 2>&1 | tail -4 | tee gpurun_out/bamread_r03h.txt
f=$(ls -t /tmp/tmp*/s.bam 2>/dev/null | head -1); echo "file: $f"
