#!/bin/bash
# Round 2, thirty-first GPU call (8 GPUs): the pipelined step under torchrun at N = 8
set -x
mkdir -p gpurun_out
TAG=${TAG:-r03e}
nvidia-smi -L | wc -l; nproc
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29563 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/bench_n8_$TAG.err > gpurun_out/bench_n8_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_n8_$TAG.txt
python - <<E
import json
j=json.loads([l for l in open("gpurun_out/bench_n8_$TAG.txt") if l.startswith("{")][-1])
print(json.dumps(j["pipeline"]["sequential"]))
print(j["n_gpus"], j["value"], j["e2e"]["value"], j["e2e"]["sv_records_on_rank0"], j["clocks"])
E
