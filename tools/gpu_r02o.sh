#!/bin/bash
# Round 2, thirteenth GPU call (8 GPUs): both arms of bench.py under torchrun
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02o}
nvidia-smi -L | head -8; nproc
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29643 bench.py --gpus 8 --steps 5 --warmup 3 2>gpurun_out/bench_n8_$TAG.err > gpurun_out/bench_n8_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_n8_$TAG.txt; grep -v "^\*\|OMP_NUM" gpurun_out/bench_n8_$TAG.err | tail -8
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29644 bench.py --impl reference --gpus 8 --steps 3 --warmup 1 2>gpurun_out/bench_n8_ref_$TAG.err > gpurun_out/bench_n8_ref_$TAG.txt ) 2>&1 | tail -4
cut -c1-400 gpurun_out/bench_n8_ref_$TAG.txt
