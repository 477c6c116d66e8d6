#!/bin/bash
# Round 2, twenty-third GPU call: pool_available with the driver's figure cached -- lap times again, then the line
set -x
mkdir -p gpurun_out
SVB_STAGE_STATS=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-config2 --no-cpu-baseline --no-call-stage 2> gpurun_out/stage_r02w.err > gpurun_out/stage_r02w.txt
grep "svb-stage" gpurun_out/stage_r02w.err | grep "pool_available\|call: " | tail -40
timeout 900 python bench.py --steps 10 --warmup 3 --no-config2 --no-cpu-baseline --no-call-stage 2> gpurun_out/bench_r02w.err > gpurun_out/bench_r02w.txt
python tools/bench_brief.py gpurun_out/bench_r02w.txt
