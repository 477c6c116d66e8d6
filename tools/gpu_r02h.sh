#!/bin/bash
# Round 2, eighth GPU call: whole GPU suite, final-size line, then the ncu evidence of the same step
# (launch list with device times + one --set full capture of the POA, search and ksw2 kernels)
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02h}
timeout 1500 python -m pytest tests -q -m gpu 2>&1 | tail -6 | tee gpurun_out/gpu_tests_$TAG.txt
( time timeout 1500 python bench.py --steps 10 --warmup 3 2>gpurun_out/bench_full_$TAG.err > gpurun_out/bench_full_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_full_$TAG.txt
# the streamed e2e kernel spins on flags written by copies queued behind it: under ncu (kernels serialised) that never
# completes, so the profiled runs use the plain upload path
export SVB_NO_STREAM=1
SVB_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_$TAG.csv timeout 900 python bench.py --steps 2 --warmup 1 --no-config2 --no-cpu-baseline --no-call-stage > gpurun_out/launches_bench_$TAG.log 2>&1
tail -2 gpurun_out/launches_bench_$TAG.log | cut -c1-300
SVB_PROFILE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"k_poa|k_sfs_search_mop|k_ksw_extd2|k_cl_" -c 7 \
  -o gpurun_out/prof_$TAG -f timeout 1200 python bench.py --steps 1 --warmup 0 --no-config2 --no-cpu-baseline --no-call-stage > gpurun_out/prof_bench_$TAG.log 2>&1
tail -2 gpurun_out/prof_bench_$TAG.log | cut -c1-300
ls -la gpurun_out | tail -8
