#!/bin/bash
# Round 2, twenty-fifth GPU call: the pipelined step (search + cluster of batch k+1 while batch k is called)
set -x
mkdir -p gpurun_out
TAG=r02y
timeout 900 python bench.py --steps 10 --warmup 3 --no-config2 --no-cpu-baseline --no-call-stage 2> gpurun_out/bench_$TAG.err > gpurun_out/bench_$TAG.txt
tail -3 gpurun_out/bench_$TAG.err
python tools/bench_brief.py gpurun_out/bench_$TAG.txt
python - <<'E'
import json
j=json.loads([l for l in open("gpurun_out/bench_r02y.txt") if l.startswith("{")][-1])
print(json.dumps(j["pipeline"]["sequential"]))
print(j["ms_per_step"], j["ms_per_step_cuda_events"], j["e2e"]["ms_per_step"], j["gpu_launches"])
E
