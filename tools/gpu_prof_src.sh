#!/bin/bash
# steady-state (READS large enough that every thread walks several reads) source-level capture with
# the minimum of replay passes
set -x
mkdir -p gpurun_out
READS=${READS:-1000000}
TAG=${TAG:-src}
export SVB_NO_STREAM=1
SVB_PROFILE=1 timeout ${NCU_TIMEOUT:-900} ncu --profile-from-start off --clock-control none --import-source on -k regex:k_sfs_search ${NCU_SKIP:+--launch-skip $NCU_SKIP} -c 1 \
  --section SourceCounters --section WarpStateStats --section SchedulerStats \
  --metrics dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,gpu__time_duration.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active \
  -o gpurun_out/prof_$TAG -f python bench.py --reads $READS --steps 1 --warmup 0 --no-cpu-baseline --no-rank-walk $EXTRA > gpurun_out/prof_bench_$TAG.log 2>&1
tail -3 gpurun_out/prof_bench_$TAG.log | cut -c1-300
ncu -i gpurun_out/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --csv > gpurun_out/prof_${TAG}_source.csv 2>/dev/null
ncu -i gpurun_out/prof_$TAG.ncu-rep --page source --print-source cuda --csv > gpurun_out/prof_${TAG}_cuda.csv 2>/dev/null
ls -la gpurun_out | tail -4
