#!/bin/bash
# Round 2, second GPU call: the new GPU tests (Clusterer + pcall through the C ABI), then bench.py on a scaled-down
# workload (both arms) to shake it out, then the full-size line.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02b}
timeout 600 python -m pytest tests/test_gpu_cluster_call.py -q -x 2>&1 | tail -15 | tee gpurun_out/gpu_cluster_call_$TAG.txt
timeout 600 python bench.py --ref-bp 60000000 --reads 40000 --svs 400 --config2-reads 30000 --cpu-records 8000 --steps 2 --warmup 1 2>gpurun_out/bench_small_$TAG.err | tee gpurun_out/bench_small_$TAG.txt | cut -c1-3000
tail -5 gpurun_out/bench_small_$TAG.err
timeout 600 python bench.py --impl reference --ref-bp 60000000 --reads 40000 --svs 400 --cpu-records 8000 --steps 2 --warmup 1 2>gpurun_out/bench_small_ref_$TAG.err | tee gpurun_out/bench_small_ref_$TAG.txt | cut -c1-1500
tail -5 gpurun_out/bench_small_ref_$TAG.err
SVB_POA_TIMING=1 timeout 1500 python bench.py 2>gpurun_out/bench_full_$TAG.err | tee gpurun_out/bench_full_$TAG.txt | cut -c1-6000
tail -30 gpurun_out/bench_full_$TAG.err
