#!/bin/bash
# Round 2, seventh GPU call (2 GPUs): both arms of bench.py under torchrun, as the driver launches them
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02g}
nvidia-smi -L; nproc
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 2>gpurun_out/bench_n2_$TAG.err > gpurun_out/bench_n2_$TAG.txt ) 2>&1 | tail -4
python tools/bench_brief.py gpurun_out/bench_n2_$TAG.txt; tail -6 gpurun_out/bench_n2_$TAG.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 2>gpurun_out/bench_n2_ref_$TAG.err > gpurun_out/bench_n2_ref_$TAG.txt ) 2>&1 | tail -4
cut -c1-600 gpurun_out/bench_n2_ref_$TAG.txt; tail -4 gpurun_out/bench_n2_ref_$TAG.err
