#!/bin/bash
# Round 2, twenty-second GPU call: host-side lap times of the call stage (SVB_STAGE_STATS=1) on the bench's step
set -x
mkdir -p gpurun_out
SVB_STAGE_STATS=1 timeout 900 python bench.py --steps 2 --warmup 3 --no-config2 --no-cpu-baseline --no-call-stage 2> gpurun_out/stage_r02v.err > gpurun_out/stage_r02v.txt
grep "svb-stage" gpurun_out/stage_r02v.err | tail -120
python tools/bench_brief.py gpurun_out/stage_r02v.txt
