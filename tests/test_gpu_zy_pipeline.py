"""GPU: the whole `run_svdss` sequence on our shell -- SVDSS smooth | index | search | call -- on a
raw-HiFi-shaped diploid sample (substitutions and 1-bp indels on top of planted INS/DEL): the planted
SVs must come back with exact type and length, nothing else may be called, and the SFS file must be
what the oracle finds on the smoothed reads."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from bam_writer import write_bam, read_bam
from common import oracle_index, fm_results
from svdss_b200 import build, synth

pytestmark = pytest.mark.gpu
L = "$ACGTN"


def dec(a):
    return "".join(L[int(x)] for x in a)


def test_smooth_index_search_call(tmp_path):
    build.build_lib()
    exe = build.build_host()
    contigs = synth.make_reference(400_000, seed=201, contigs=2, n_repeats=3, n_nruns=1, nrun_len=40)
    names = ["chrA", "chrB"]
    cat = synth.make_sv_catalogue(contigs, 14, seed=202, min_len=50, max_len=900, margin=4000, spacing=6000)
    alns = synth.make_sample_alignments(contigs, cat, coverage=14, seed=203, mean_len=6000, sd_len=1200, min_len=2000,
                                        max_len=12000, tag_hp=True, clip_rate=0.05, sub_rate=0.001, indel_rate=0.0005)
    fa = str(tmp_path / "ref.fa")
    with open(fa, "w") as f:
        for n, c in zip(names, contigs):
            f.write(">%s\n%s\n" % (n, dec(c)))
    bam = str(tmp_path / "raw.bam")
    write_bam(bam, [(n, len(c)) for n, c in zip(names, contigs)],
              [dict(qname=a["qname"], flag=0, tid=a["tid"], pos=a["pos"], mapq=60, seq=dec(a["seq"]), cigar=a["cigar"],
                    tags={"HP": ("C", a["hp"])}) for a in alns])
    smoothed = str(tmp_path / "smoothed.bam")
    with open(smoothed, "wb") as f:
        r = subprocess.run([exe, "smooth", "--reference", fa, "--bam", bam, "--threads", "4"], stdout=f, stderr=subprocess.PIPE, text=True)
    assert r.returncode == 0, r.stderr
    idx = str(tmp_path / "ref.svb")
    r = subprocess.run([exe, "index", "-d", "-o", idx, fa], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([exe, "search", "--index", idx, "--bam", smoothed, "--threads", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    sfs_text = r.stdout
    # the SFS file against the oracle on the smoothed reads that pass the XF filter
    _, _, recs = read_bam(smoothed)
    keep = [x for x in recs if x["tags"]["XF"][1] == 0]
    assert 0 < len(keep) < len(recs)
    T, SA, bwt = oracle_index(contigs)
    exp, _ = fm_results(oracle.FMIndex(bwt), [oracle.encode_nt6(x["seq"]) for x in keep])
    want = set()
    for x, e in zip(keep, exp):
        for qs, ln in oracle.assemble(e):
            want.add((x["qname"], qs, ln))
    got, name = set(), None
    for line in sfs_text.splitlines():
        f = line.split("\t")
        name = f[0] if f[0] != "*" else name
        got.add((name, int(f[1]), int(f[2])))
    assert got == want and len(got) > 20
    sfs = str(tmp_path / "sample.sfs")
    open(sfs, "w").write(sfs_text)
    r = subprocess.run([exe, "call", "--reference", fa, "--bam", smoothed, "--sfs", sfs, "--threads", "4"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    calls = [l.split("\t") for l in r.stdout.splitlines() if not l.startswith("#")]
    used = set()
    covered = 0
    for sv in cat:
        support = 0
        for a in alns:      # reads whose alignment carries this event
            ref = a["pos"]
            for ln, op in a["cigar"]:
                if op in "ID" and ln == sv["len"] and ref == sv["pos"] + 1 and a["tid"] == sv["contig"]:
                    support += 1
                if op in "MD":
                    ref += ln
        if support < 3:
            continue
        covered += 1
        for k, f in enumerate(calls):
            if k in used or f[0] != names[sv["contig"]] or abs(int(f[1]) - (sv["pos"] + 1)) > 30:
                continue
            if ("SVTYPE=%s;" % sv["type"]) in f[7] and ("SVLEN=%d;" % (sv["len"] if sv["type"] == "INS" else -sv["len"])) in f[7]:
                used.add(k)
                break
    assert covered >= 8 and len(used) == covered        # every well-supported planted SV is called, exactly
    assert len(calls) <= covered + 3                     # and (almost) nothing else
