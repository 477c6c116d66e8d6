#!/bin/bash
set -x
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_ksw_extd2 -s 1 -c 1 -o gpurun_out/prof_ksw -f \
  python tools/bench_call.py --clusters 0 --pairs 4000 --cpu-seconds 0.5 > gpurun_out/prof_ksw.log 2>&1
tail -2 gpurun_out/prof_ksw.log
ncu --set full --clock-control none --import-source on -k regex:k_poa -s 1 -c 1 -o gpurun_out/prof_poa -f \
  python tools/bench_call.py --clusters 2400 --pairs 0 --cpu-seconds 0.5 > gpurun_out/prof_poa.log 2>&1
tail -2 gpurun_out/prof_poa.log
