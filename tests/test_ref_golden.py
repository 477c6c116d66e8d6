"""Vectors written by the REFERENCE binary (oracle/build_ref.sh + oracle/ref_vectors.sh, for whoever has the sources
the reference's CMake fetches -- this container does not, so these tests skip here and the oracle stays "parity
unpinned").  When tests/golden/ref_world.* exist they pin, on the fixture world of
tests/sv_world.py: the SFS set per read (oracle, and the GPU through the CLI), the ropebwt3 .fmd layout
(host/rld.hpp), the cluster file and the VCF body.

What a disagreement would mean: .sfs -- the SFS statement itself (it is index-independent, so this would be a bug);
.fmd -- the RLD layout restated from memory; .clusters -- only std::sort's tie order may legitimately differ (we sort
stably); .vcf -- POA tie-breaks (abPOA's adaptive band, heaviest-bundle ties) and ksw2's CIGAR ties (H >= E >= F >= E2
>= F2 with strict > to switch, left-aligned gaps) are restated from memory: a difference there is confined to records
whose alignment has equal-score alternatives."""
import os
import subprocess

import pytest

import oracle
from common import oracle_index, fm_results
from sv_world import make_world

HERE = os.path.dirname(os.path.abspath(__file__))
G = os.path.join(HERE, "golden")
need = pytest.mark.skipif(not os.path.exists(os.path.join(G, "ref_world.sfs")),
                          reason="no reference-made vectors (oracle/build_ref.sh needs sources this container does not have)")


def parse_sfs(path):
    out, name = {}, None
    for line in open(path):
        f = line.split()
        if len(f) < 4:
            continue
        if f[0] != "*":
            name = f[0]
        out.setdefault(name, set()).add((int(f[1]), int(f[2])))
    return out


@need
def test_oracle_sfs_equals_the_reference(tmp_path):
    w = make_world(str(tmp_path))
    ref = parse_sfs(os.path.join(G, "ref_world.sfs"))
    T, SA, bwt = oracle_index(w["contigs"])
    import numpy as np
    L = {c: i for i, c in enumerate("$ACGTN")}
    recs = [r for r in w["records"] if not (r["flag"] & 0x904) and len(r["seq"]) >= 100]
    res, _ = fm_results(oracle.FMIndex(bwt), [np.array([L[c] for c in r["seq"]], np.uint8) for r in recs])
    for r, e in zip(recs, res):
        assert set(oracle.assemble(e)) == ref.get(r["qname"], set()), r["qname"]


@need
@pytest.mark.gpu
def test_gpu_cli_equals_the_reference(tmp_path):
    from svdss_b200 import build
    exe = build.build_host()
    w = make_world(str(tmp_path))
    # the reference's own index file first: pins host/rld.hpp
    r = subprocess.run([exe, "search", "--index", os.path.join(G, "ref_world.fmd"), "--bam", w["bam"], "--noputative", "--threads", "4"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sfs = tmp_path / "gpu.sfs"
    sfs.write_text(r.stdout)
    assert parse_sfs(str(sfs)) == parse_sfs(os.path.join(G, "ref_world.sfs"))
    cl = tmp_path / "gpu.clusters"
    r = subprocess.run([exe, "call", "--threads", "4", "--reference", w["fa"], "--bam", w["bam"], "--sfs", os.path.join(G, "ref_world.sfs"),
                        "--clusters", str(cl)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert sorted(cl.read_text().splitlines()) == sorted(open(os.path.join(G, "ref_world.clusters")).read().splitlines())
    body = [l for l in r.stdout.splitlines() if not l.startswith("#")]
    assert body == open(os.path.join(G, "ref_world.vcf")).read().splitlines()
