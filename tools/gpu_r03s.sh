#!/bin/bash
# Round 2 (4 GPUs): the line under torchrun at N = 4 on the final code
set -x
mkdir -p gpurun_out
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29573 bench.py --gpus 4 --steps 5 --warmup 3 2>gpurun_out/bench_n4_r03s.err > gpurun_out/bench_n4_r03s.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_n4_r03s.txt 2>/dev/null | head -2 | cut -c1-300
python - <<E
import json
j=json.loads([l for l in open("gpurun_out/bench_n4_r03s.txt") if l.startswith("{")][-1])
p=j["pipeline"]; print(p["value_pipelined"], p["e2e_pipelined"], round(p["sequential"]["value"]), round(p["sequential"]["e2e_value"]))
print(j["n_gpus"], round(j["value"]), round(j["e2e"]["value"]), j["e2e"]["sv_records_on_rank0"])
E
