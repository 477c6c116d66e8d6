#!/bin/bash
# ncu evidence: launch list of the timed region + one full capture of the search kernel
set -x
mkdir -p gpurun_out
READS=${READS:-100000}
# the streamed e2e kernel spins on chunk-ready flags written by copies queued behind it: under ncu
# (kernels serialised) that never completes, so the profiled runs use the plain upload path
export SVB_NO_STREAM=1
TAG=${TAG:-sfs}
if [ -z "$SKIP_LAUNCHES" ]; then
SVB_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_$TAG.csv timeout 600 python bench.py --reads $READS --steps 2 --warmup 1 --no-cpu-baseline --no-rank-walk > gpurun_out/launches_bench_$TAG.log 2>&1
tail -3 gpurun_out/launches_bench_$TAG.log
fi
SVB_PROFILE=1 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_sfs_search -c 1 \
  -o gpurun_out/prof_$TAG -f timeout 900 python bench.py --reads $READS --steps 1 --warmup 0 --no-cpu-baseline --no-rank-walk > gpurun_out/prof_bench_$TAG.log 2>&1
tail -3 gpurun_out/prof_bench_$TAG.log
ls -la gpurun_out
