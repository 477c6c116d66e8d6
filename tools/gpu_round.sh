#!/bin/bash
# full GPU pass: tests, smoke, bench (+reference arm), ncu evidence.  TAG names the outputs.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r01}
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
python __graft_entry__.py smoke
SVB_SEARCH_STATS=1 timeout 1200 python bench.py 2>gpurun_out/bench_full_$TAG.err | tee gpurun_out/bench_full_$TAG.txt | cut -c1-400
grep "k_sfs_search_mop\]" gpurun_out/bench_full_$TAG.err | head -3
timeout 1200 python bench.py --impl reference --steps 2 --warmup 1 2>&1 | tail -1 | tee gpurun_out/bench_reference_$TAG.txt | cut -c1-300
# launch list of the timed region (plain upload path: see tools/gpu_profile.sh) + one full-set capture
export SVB_NO_STREAM=1
SVB_PROFILE=1 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
  --log-file gpurun_out/launches_$TAG.csv timeout 600 python bench.py --reads 250000 --steps 2 --warmup 1 --no-cpu-baseline --no-rank-walk > gpurun_out/launches_bench_$TAG.log 2>&1
tail -2 gpurun_out/launches_bench_$TAG.log | cut -c1-200
SVB_PROFILE=1 timeout 900 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:k_sfs_search -c 1 \
  -o gpurun_out/prof_$TAG -f python bench.py --reads 1000000 --steps 1 --warmup 0 --no-cpu-baseline --no-rank-walk > gpurun_out/prof_bench_$TAG.log 2>&1
tail -2 gpurun_out/prof_bench_$TAG.log | cut -c1-200
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ls -la gpurun_out | tail -8
