#!/bin/bash
# parity tests of the search path + bench line (no CPU baseline): the inner loop of kernel tuning
timeout 900 python -m pytest tests/test_gpu_search.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --no-cpu-baseline --steps ${STEPS:-2} 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('value %.0f reads/s  ms/step %.1f  kernel_ms %.1f  frac %.3f  Gext/s %.2f  e2e %.0f (%.1f ms)  clocks %s' % (d['value'], d['ms_per_step'], r['kernel_ms'], r['frac'], r['extensions_per_s']/1e9, d['e2e']['value'], d['e2e']['ms_per_step'], d['clocks']))"
