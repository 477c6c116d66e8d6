"""GPU: a ropebwt3-format `.fmd` written from the GPU-built BWT (`SVDSS index --fmd`) holds the oracle's
BWT, and `search` on it (decode -> invert -> re-index on the GPU), as well as on its `index --from-fmd`
conversion, prints exactly what `search` prints on the library's own index file.
(Named to run after the kernel parity tests: this one exercises host glue, not a kernel.)"""
import os
import subprocess

import numpy as np
import pytest

import rld_model
from common import oracle_index
from svdss_b200 import build, synth

pytestmark = pytest.mark.gpu
L = "$ACGTN"


def dec(a):
    return "".join(L[int(x)] for x in a)


def test_fmd_round_trip_through_the_shell(tmp_path):
    build.build_lib()
    exe = build.build_host()
    contigs = synth.make_reference(60_000, seed=81, contigs=3, n_repeats=3, n_nruns=1, nrun_len=60)
    fa = str(tmp_path / "ref.fa")
    with open(fa, "w") as f:
        for i, c in enumerate(contigs):
            f.write(">chr%d\n%s\n" % (i + 1, dec(c)))
    idx, fmd, idx2 = str(tmp_path / "ref.svb"), str(tmp_path / "ref.fmd"), str(tmp_path / "ref2.svb")
    r = subprocess.run([exe, "index", "-t4", "-d", "-o", idx, "--fmd", fmd, fa], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    T, SA, bwt = oracle_index(contigs)
    assert rld_model.decode(rld_model.parse(fmd)) == bwt.tobytes()
    r = subprocess.run([exe, "index", "--from-fmd", fmd, "-o", idx2], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(idx2, "rb").read() == open(idx, "rb").read()
    reads = synth.make_reads(contigs, 60, seed=82, mean_len=3000, sd_len=800, min_len=200, max_len=6000)
    fq = str(tmp_path / "reads.fq")
    with open(fq, "w") as f:
        for i, rd in enumerate(reads):
            f.write("@r%03d\n%s\n+\n%s\n" % (i, dec(rd), "I" * len(rd)))
    outs = []
    for index in (idx, fmd, idx2):
        r = subprocess.run([exe, "search", "--index", index, "--fastx", fq, "--threads", "2"], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        outs.append(r.stdout)
    assert outs[0] and outs[1] == outs[0] and outs[2] == outs[0]
    # the same through the C ABI: svb_index_load tells the two formats apart by their magic
    from svdss_b200 import capi
    a, b = capi.Index.load(idx), capi.Index.load(fmd)
    assert a.n == b.n == len(bwt)
    cat, offs = synth.concat(reads)
    ra, rb = a.sfs_batch(cat, offs, assemble=True), b.sfs_batch(cat, offs, assemble=True)
    assert ra.n_sfs == rb.n_sfs > 0 and all(ra.per_read(i) == rb.per_read(i) for i in range(len(reads)))
