#!/bin/bash
# Round 2: ncu of the device BAM loader's kernels (one 64 MiB window of the BAM-shaped test file)
set -x
mkdir -p gpurun_out
python - <<'E'
import subprocess, sys, os
sys.path.insert(0, os.getcwd())
# a BAM of ~1 window: reuse the generator of tools/bench_bamread.py by importing its pieces would be longer than running it once with a kept file
E
timeout 900 python tools/bench_bamread.py --records 16000 --repeat 1 --gpu-inflate --keep gpurun_out/one_window.bam > gpurun_out/bamread_r03q.txt 2>&1
ls -la gpurun_out/one_window.bam
SVB_BGZF_GPU_MIN_BYTES=0 SVB_BAMREAD_DEVICE=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_bam|k_bgzf" -c 12 -o gpurun_out/r03q_bamstream -f svdss_b200/SVDSS _bamread gpurun_out/one_window.bam --gpu-inflate > gpurun_out/ncu_r03q.log 2>&1
tail -3 gpurun_out/ncu_r03q.log
rm -f gpurun_out/one_window.bam
ls -la gpurun_out/*.ncu-rep | tail -2
