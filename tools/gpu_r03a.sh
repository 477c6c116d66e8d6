#!/bin/bash
# Round 2, twenty-seventh GPU call: CTA-rows POA at 3 CTAs per SM (variant 12743, 167 registers) next to 2 (6599)
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_poa_variants.py -x -q -s 2>&1 | tail -4
for v in 12743 6599; do
  SVB_POA_VARIANT=$v timeout 900 python bench.py --steps 6 --warmup 3 --no-config2 --no-cpu-baseline --no-call-stage --no-pipeline 2> gpurun_out/bench_r03a_$v.err > gpurun_out/bench_r03a_$v.txt
  echo "variant $v"; python tools/bench_brief.py gpurun_out/bench_r03a_$v.txt 2>/dev/null | head -1
done
