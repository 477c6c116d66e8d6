"""GPU: the 2-bit transport of the streamed search (SVB_STREAM_PACK2=1: every chunk re-packed on the host by
svb_pack2_chunk, decoded by the unpacking CTAs of k_sfs_search_mop<.., .., true>, N positions patched before the
chunk's flag goes up) gives the same SFS tables and extension counts as the 4-bit transport, with about half
the bytes across PCIe.  Green on a B200 since round 2 (profiles/r02a_variants_sweep.txt); the transport itself stays
opt-in: on the bench host (16 cores per GPU) re-packing 7.6 GB per million reads on the host costs more than the PCIe
time it saves (2.0 M against 5.6 M reads/s end to end, profiles/r02b_bench_full.txt).  Child process with a time limit."""
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
from svdss_b200 import capi, synth
contigs = synth.make_reference(500_000, seed=51, contigs=3)
reads = synth.make_reads(contigs, 400, seed=52, mean_len=7000, sd_len=2500, min_len=150, max_len=18000)
reads += synth.make_reads(contigs, 40, seed=53, mean_len=3001, sd_len=400, min_len=301, max_len=5001, raw_hifi=True)
reads.insert(5, np.zeros(0, np.uint8))
reads.insert(77, reads[10][:1].copy())
reads[20] = reads[20].copy(); reads[20][100:104] = 5
reads[300] = np.full(2600, 5, np.uint8)                               # a read of N only (restarts inside it have a closed form since round 2)
cat, offs = oracle.concat(contigs)
idx = capi.Index.build(cat, offs, block_bytes=128)
seq4, s4o, lq = capi.pack_bam4(reads)
seq4 = seq4.copy()
seq4[int(s4o[20]) + 50] = (5 << 4) | 5                                # an IUPAC code decodes to N as well
os.environ["SVB_STREAM_MIN_BYTES"] = "1"
for chunk in ("65536", "100032", "1048576"):
    os.environ["SVB_STREAM_CHUNK_BYTES"] = chunk
    for assemble in (False, True):
        os.environ["SVB_STREAM_PACK2"] = "0"
        a = idx.sfs_batch_bam4(seq4, s4o, lq, assemble=assemble)
        os.environ["SVB_STREAM_PACK2"] = "1"
        b = idx.sfs_batch_bam4(seq4, s4o, lq, assemble=assemble)
        assert a.n_sfs == b.n_sfs and a.n_ext == b.n_ext, (chunk, assemble, a.n_sfs, b.n_sfs, a.n_ext, b.n_ext)
        assert all(a.per_read(i) == b.per_read(i) for i in range(len(reads))), (chunk, assemble)
        assert b.h2d_bytes < 0.75 * a.h2d_bytes, (a.h2d_bytes, b.h2d_bytes)
os.environ["SVB_PACK2_EXC_CAP"] = "16"                                # chunks with more N than that fall back to the 4-bit form
os.environ["SVB_STREAM_CHUNK_BYTES"] = "65536"
c = idx.sfs_batch_bam4(seq4, s4o, lq, assemble=True)
assert c.n_sfs == b.n_sfs and c.n_ext == b.n_ext and all(c.per_read(i) == b.per_read(i) for i in range(len(reads)))
print("STREAM_PACK2_OK h2d bytes 4-bit %d, 2-bit %d" % (a.h2d_bytes, b.h2d_bytes))
"""


def test_two_bit_transport_equals_four_bit_transport():
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", CHILD], cwd=root, capture_output=True, text=True, timeout=200)
    assert r.returncode == 0 and "STREAM_PACK2_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-4000:]
    print(r.stdout.strip())
