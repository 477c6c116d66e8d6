"""GPU: the k_ksw_extd2 variants of SVB_KSW_VARIANT (1: backtrack in windows of 32 steps with the traceback
bytes prefetched along the predicted diagonal; 2: checkpointed traceback, bands recomputed during the backtrack;
ksw_kernel.cuh) give the same scores and CIGARs as the default kernel and the oracle.  Off by default until
they have been measured; child process like the POA variants."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

CHILD = r"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.getcwd(), "tests")); sys.path.insert(0, os.getcwd())
import oracle
from ksw_cases import make_pairs, planted_pairs
from svdss_b200 import capi
rng = np.random.default_rng(61)
pairs = make_pairs(rng, 300) + planted_pairs(rng, 60)
qc, qo = oracle.concat([p[0] for p in pairs]); tc, to = oracle.concat([p[1] for p in pairs])
os.environ["SVB_KSW_VARIANT"] = "0"
a = capi.ksw_extd2_batch(qc, qo, tc, to)
times = ["0: %.2f" % a.kernel_ms]
for variant in (1, 2, 3):
    os.environ["SVB_KSW_VARIANT"] = str(variant)
    b = capi.ksw_extd2_batch(qc, qo, tc, to)
    for k, (q, t) in enumerate(pairs):
        assert int(a.score[k]) == int(b.score[k]) and a.cigar_of(k) == b.cigar_of(k), (variant, k)
        if k % 5 == 0:
            sc, cg = oracle.ksw_extd2(q, t)
            assert sc == int(b.score[k]) and cg == b.cigar_of(k), (variant, k)
    times.append("%d: %.2f" % (variant, b.kernel_ms))
os.environ["SVB_KSW_VARIANT"] = "3"
os.environ["SVB_KSW_TB_BYTES"] = str(8 << 20)                           # several waves with the checkpointed layout
b = capi.ksw_extd2_batch(qc, qo, tc, to)
assert b.waves > 1 and all(int(a.score[k]) == int(b.score[k]) and a.cigar_of(k) == b.cigar_of(k) for k in range(len(pairs)))
print("KSW_VARIANT_OK kernel ms by variant  " + "  ".join(times))
"""


def test_variants_equal_default_kernel():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=root, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0 and "KSW_VARIANT_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    print(r.stdout.strip())
