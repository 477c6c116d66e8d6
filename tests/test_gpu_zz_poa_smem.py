"""GPU: k_poa<true> (SVB_POA_SMEM=1: the previous row's scores kept in shared memory) gives the same
consensus, status and cell count as the default kernel and as the banded oracle.  The variant is
off by default until it has been measured; it runs in a child process so that a fault in it cannot
take the CUDA context of the other tests with it."""
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
os.environ["SVB_POA_SMEM"] = "0"
a = capi.poa_batch(clusters)
os.environ["SVB_POA_SMEM"] = "1"
b = capi.poa_batch(clusters)
assert a.cells == b.cells, (a.cells, b.cells)
for c, reads in enumerate(clusters):
    assert np.array_equal(a.consensus(c), b.consensus(c)), c
    if c % 4 == 0 and reads:
        assert np.array_equal(b.consensus(c), oracle.poa_consensus(reads, band=True)), c
print("POA_SMEM_OK kernel ms default %.2f smem %.2f" % (a.kernel_ms, b.kernel_ms))
"""


def test_smem_variant_equals_default_kernel():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=root, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0 and "POA_SMEM_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    print(r.stdout.strip())
