#!/bin/bash
# Round 2, twenty-eighth GPU call (2 GPUs): the pipelined step under torchrun (NCCL gather on the calling thread while the worker searches)
set -x
mkdir -p gpurun_out
TAG=r03b
nvidia-smi -L | head -8; nproc
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29553 bench.py --gpus 2 --steps 10 --warmup 3 2>gpurun_out/bench_n2_$TAG.err > gpurun_out/bench_n2_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_n2_$TAG.txt; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n2_$TAG.err | tail -6
python - <<'E'
import json
j=json.loads([l for l in open("gpurun_out/bench_n2_r03b.txt") if l.startswith("{")][-1])
print(json.dumps(j["pipeline"]["sequential"]))
print(j["n_gpus"], j["value"], j["e2e"]["value"], j["e2e"]["sv_records_on_rank0"])
E
