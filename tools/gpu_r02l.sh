#!/bin/bash
# Round 2, twelfth GPU call: the search kernel's configurations side by side (mop / cpa / tma / lane groups) on configs[1]
set -x
mkdir -p gpurun_out
timeout 1200 python tools/cfg_sweep.py 1000000 2>gpurun_out/cfg_sweep_r02l.err | tee gpurun_out/cfg_sweep_r02l.txt
tail -2 gpurun_out/cfg_sweep_r02l.err
