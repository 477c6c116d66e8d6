"""GPU: the k_poa variants of SVB_POA_VARIANT (previous row's scores in shared memory, in1 traceback, warp-wide
remain[] / re-rank, packed per-rank row records, speculative traceback, windowed graph update; poa_kernel.cuh) give
the same consensus, status and cell count as variant 0 -- the kernel measured in round 1 -- and as the banded oracle.
They run in a child process so that a fault in one cannot take the CUDA context of the other tests with it."""
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
from poa_cases import make_cluster
from svdss_b200 import capi
rng = np.random.default_rng(31)
clusters = [make_cluster(rng, n_reads=int(rng.integers(2, 20)), tlen=int(rng.integers(40, 400)), rate=0.01)[1] for _ in range(40)]
clusters += [make_cluster(rng)[1] for _ in range(10)]                      # config-4 shapes: 20-60 reads x 200-2000 bp
clusters += [[], [clusters[0][0]]]
clusters.append([rng.integers(0, 4, size=int(rng.integers(150, 260))).astype(np.uint8) for _ in range(12)])   # overflow -> rerun with worst-case wcap
os.environ["SVB_POA_VARIANT"] = "0"
a = capi.poa_batch(clusters)
times = ["0: %.2f" % a.kernel_ms]
for variant, group in ((455, 32), (3527, 32), (4551, 32), (6599, 32), (12743, 32)):
    os.environ["SVB_POA_VARIANT"] = str(variant)
    b = capi.poa_batch(clusters)
    assert a.cells == b.cells, (variant, a.cells, b.cells)
    for c, reads in enumerate(clusters):
        assert np.array_equal(a.consensus(c), b.consensus(c)), (variant, c)
        if c % 4 == 0 and reads:
            assert np.array_equal(b.consensus(c), oracle.poa_consensus(reads, band=True)), (variant, group, c)
    times.append("%d/g%d: %.2f" % (variant, group, b.kernel_ms))
print("POA_VARIANTS_OK kernel ms by variant  " + "  ".join(times))
"""


def test_variants_equal_default_kernel():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=root, capture_output=True, text=True, timeout=150)
    assert r.returncode == 0 and "POA_VARIANTS_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    print(r.stdout.strip())
