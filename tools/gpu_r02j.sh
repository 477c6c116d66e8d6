#!/bin/bash
# Round 2, tenth GPU call: whole GPU suite, the round's line (both arms), ncu of the POA kernel in both regimes
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02j}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests_$TAG.txt
python __graft_entry__.py smoke 2>&1 | tail -1
( time timeout 1500 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_full_$TAG.err > gpurun_out/bench_full_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_full_$TAG.txt
( time timeout 900 python bench.py --impl reference --steps 3 --warmup 1 2>gpurun_out/bench_ref_$TAG.err > gpurun_out/bench_ref_$TAG.txt ) 2>&1 | tail -4
cut -c1-300 gpurun_out/bench_ref_$TAG.txt
export SVB_NO_STREAM=1
SVB_POA_VARIANT=455 ncu --set full --clock-control none --import-source on -k regex:k_poa -s 1 -c 1 -o gpurun_out/prof_poa4_$TAG -f \
  timeout 900 python tools/bench_call.py --clusters 12000 --pairs 0 --cpu-seconds 0.2 > gpurun_out/prof_poa4_$TAG.log 2>&1
tail -1 gpurun_out/prof_poa4_$TAG.log | cut -c1-300
SVB_PROFILE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_poa" -c 1 \
  -o gpurun_out/prof_poa3_$TAG -f timeout 900 python bench.py --steps 1 --warmup 0 --no-config2 --no-cpu-baseline --no-call-stage > gpurun_out/prof_poa3_$TAG.log 2>&1
tail -1 gpurun_out/prof_poa3_$TAG.log | cut -c1-200
ls -la gpurun_out | tail -6
