#!/bin/bash
# Round 2, sixteenth GPU call: device inflate (thread per member) on bigger windows -- does it scale with the members per launch?
set -x
mkdir -p gpurun_out
for mb in 512 1024 2048; do timeout 900 python tools/bench_inflate.py --mb $mb 2>&1 | tail -1; done | tee gpurun_out/inflate_r02p.txt
timeout 600 python tools/bench_bamread.py --records 40000 2>&1 | tail -1 | tee gpurun_out/bamread_r02p.txt
