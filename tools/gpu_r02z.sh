#!/bin/bash
# Round 2, twenty-sixth GPU call: POA slot cap for chain-bound batches (SVB_POA_SLOT_HEADROOM) under the pipelined step
set -x
mkdir -p gpurun_out
SVB_STAGE_STATS=1 timeout 600 python bench.py --steps 1 --warmup 1 --no-config2 --no-cpu-baseline --no-call-stage --no-pipeline 2>&1 | grep "row chains" | tail -1
for h in 0 0.4 0.6 0.8 1.0; do
  SVB_POA_SLOT_HEADROOM=$h timeout 900 python bench.py --steps 10 --warmup 3 --no-config2 --no-cpu-baseline --no-call-stage 2> gpurun_out/bench_r02z_$h.err > gpurun_out/bench_r02z_$h.txt
  echo "headroom $h"; python tools/bench_brief.py gpurun_out/bench_r02z_$h.txt | head -2
  python - <<E
import json
j=json.loads([l for l in open("gpurun_out/bench_r02z_$h.txt") if l.startswith("{")][-1])
q=j["pipeline"]["sequential"]; print("  sequential", round(q["ms_per_step"],1), round(q["e2e_ms_per_step"],1), {k:round(v,1) for k,v in q["stages_ms"].items()}, " pipelined", round(j["ms_per_step"],1), round(j["e2e"]["ms_per_step"],1))
E
done
