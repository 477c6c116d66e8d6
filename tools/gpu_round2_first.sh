#!/bin/bash
# First GPU call of round 2, in the order that loses least if the box time runs out:
# the GPU suite (the zy / zz files have never met a GPU), smoke, the bench line, the 2-bit transport of the e2e arm
# next to the 4-bit one, then the POA / ksw2 variant sweep.  TAG names the outputs.
set -x
mkdir -p gpurun_out
TAG=${TAG:-r02a}
timeout 1500 python -m pytest tests -q -m gpu -s 2>&1 | tail -25 | tee gpurun_out/gpu_tests_$TAG.txt
python __graft_entry__.py smoke 2>&1 | tail -2
# the streamed 2-bit transport is skipped by default (its first run in round 1 did not finish in 38 s): look at it here, bounded
SVB_TEST_STREAM_PACK2=1 SVB_SEARCH_STATS=1 timeout 240 python -m pytest tests/test_gpu_zz_stream_pack2.py -q -s 2>&1 | tail -15 | tee gpurun_out/stream_pack2_$TAG.txt
SVB_SEARCH_STATS=1 timeout 1200 python bench.py 2>gpurun_out/bench_full_$TAG.err | tee gpurun_out/bench_full_$TAG.txt | cut -c1-600
# e2e through svb_sfs_batch_bam4 with every chunk re-packed to 2 bits on the host (the line's e2e object is the one to read)
SVB_STREAM_PACK2=1 timeout 900 python bench.py --no-cpu-baseline --no-rank-walk --no-call-stage 2>gpurun_out/bench_pack2_$TAG.err | \
  tee gpurun_out/bench_pack2_$TAG.txt | python -c "
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1]); e = d['e2e']
print('pack2 e2e %.0f reads/s, %.1f ms/step, h2d %.2f GB/step' % (e['value'], e['ms_per_step'], e['h2d_bytes_per_step'] / 1e9))"
# device inflate of BGZF members (building block): GB/s with and without random qualities
for q in "" "--quals"; do timeout 600 python tools/bench_inflate.py --mb 256 $q 2>&1 | tail -1; done | tee gpurun_out/inflate_$TAG.txt
TAG=$TAG N=${N:-12000} bash tools/gpu_poa_variants.sh
ls -la gpurun_out | tail -12
