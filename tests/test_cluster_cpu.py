"""The Clusterer of `SVDSS call` through the CLI (`SVDSS call --cluster-only --clusters FILE`: BAM scan on the host,
svb_cluster_batch on the GPU) against the literal Python transcription in tests/cluster_model.py -- those tests
carry the gpu marker; the rest of the file (usage errors, the oracle pipeline as a checker, fuzz_ratio, and that the
stage fails loudly without a device) runs on the CPU."""
import os
import subprocess

import pytest

import cluster_model
from sv_world import make_world
from svdss_b200 import build


@pytest.fixture(scope="module")
def exe():
    build.build_lib()
    return build.build_host()


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    return make_world(str(tmp_path_factory.mktemp("clu")))


def run_clusterer(exe, w, threads, extra=()):
    out = os.path.join(w["d"], "clusters_%d.txt" % threads)
    r = subprocess.run([exe, "call", "--reference", w["fa"], "--bam", w["bam"], "--sfs", w["sfs"], "--threads", str(threads),
                        "--cluster-only", "--clusters", out] + list(extra), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert r.stdout == ""
    return open(out).read(), r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 3, 4])
def test_clusters_match_the_transcription(exe, world, threads):
    got, log = run_clusterer(exe, world, threads)
    exp = cluster_model.run(world["records"], world["names"], world["ref_seqs"], world["sfs_by_read"], threads=threads)
    assert len(exp) >= 8                                        # the planted SVs produce clusters
    assert got == cluster_model.clusters_text(exp)
    assert "Placing SFSs on reference genome" in log


@pytest.mark.gpu
def test_every_planted_sv_has_a_cluster(exe, world):
    got, _ = run_clusterer(exe, world, 4)
    spans = []
    for line in got.splitlines():
        reg, n = line.split("\t")[:2]
        chrom, se = reg.rsplit(":", 1)
        s, e = se.split("-")
        spans.append((chrom, int(s), int(e), int(n)))
    n_checked = 0
    for sv in world["catalogue"]:
        chrom = world["names"][sv["contig"]]
        lo, hi = sv["pos"] + 1, sv["pos"] + 1 + (sv["len"] if sv["type"] == "DEL" else 0)
        support = 0
        for r in world["records"]:        # reads whose alignment carries this event
            ref = r["pos"]
            for ln, op in r["cigar"]:
                if op in "ID" and ref == lo and ln == sv["len"] and r["tid"] == sv["contig"] and r["mapq"] >= 20 and r["flag"] == 0:
                    support += 1
                if op in "MD":
                    ref += ln
        if support < 2:
            continue
        n_checked += 1
        assert any(c == chrom and s <= lo + 1 and hi <= e and n >= 2 for c, s, e, n in spans), sv
    assert n_checked >= 8


@pytest.mark.gpu
def test_min_mapq_and_weight_flags(exe, world):
    # mapq 0 keeps the low-mapq reads; weight 100 leaves no cluster with sub-reads
    got0, _ = run_clusterer(exe, world, 2, ["--min-mapq", "0"])
    exp0 = cluster_model.run(world["records"], world["names"], world["ref_seqs"], world["sfs_by_read"], threads=2, min_mapq=0)
    assert got0 == cluster_model.clusters_text(exp0)
    got1, _ = run_clusterer(exe, world, 2, ["--min-cluster-weight", "100"])
    assert got1 == ""


def test_call_usage_needs_inputs(exe, world):
    r = subprocess.run([exe, "call", "--reference", world["fa"], "--bam", world["bam"]], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr
    r = subprocess.run([exe, "call", "--reference", world["fa"], "--bam", os.path.join(world["d"], "none.bam"), "--sfs", world["sfs"],
                        "--cluster-only"], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot read BAM" in r.stderr


def test_cluster_stage_fails_loudly_without_a_device(exe, world):
    from svdss_b200 import capi
    if capi.lib().svb_device_count() >= 1:
        pytest.skip("a GPU is present")
    r = subprocess.run([exe, "call", "--reference", world["fa"], "--bam", world["bam"], "--sfs", world["sfs"], "--cluster-only",
                        "--clusters", os.path.join(world["d"], "x.txt")], capture_output=True, text=True)
    assert r.returncode == 1 and "svb_cluster_batch" in r.stderr and r.stdout == ""        # no CPU fallback in the product


def test_oracle_pipeline_recovers_planted_svs(world):
    """The checker itself: Clusterer transcription -> oracle POA -> oracle ksw2 -> CIGAR walk must call
    the planted SVs (type and length exact) -- this is what the GPU `call` is compared with."""
    import call_model
    clusters = cluster_model.run(world["records"], world["names"], world["ref_seqs"], world["sfs_by_read"], threads=4)
    lines, st = call_model.call_vcf_lines(world["ref_seqs"], clusters, threads=4, return_stats=True)
    found = 0
    for sv in world["catalogue"]:
        chrom = world["names"][sv["contig"]]
        for l in lines:
            f = l.split("\t")
            if f[0] == chrom and abs(int(f[1]) - (sv["pos"] + 1)) <= 20 and ("SVTYPE=%s;" % sv["type"]) in f[7] \
                    and ("SVLEN=%d;" % (sv["len"] if sv["type"] == "INS" else -sv["len"])) in f[7]:
                found += 1
                break
    assert found >= 9, (found, len(lines))
    assert st["after"] <= st["before_chain"]


def test_fuzz_ratio_matches_lcs_definition(exe):
    """filter_sv_chains' rapidfuzz::fuzz::ratio (caller.cpp:455-458) = 100 * 2*LCS / (|a|+|b|): the shell's
    bit-parallel LCS (multi-word carry) against a plain DP, incl. the rapidfuzz doc example (96.55)."""
    import numpy as np
    import call_model
    rng = np.random.default_rng(5)
    cases = [("this is a test", "this is a test!"), ("A", "A"), ("ACGT", "TGCA")]
    for n, m in [(63, 64), (64, 65), (130, 257), (700, 650), (1, 300)]:
        a = "".join(rng.choice(list("ACGT"), size=n))
        b = list(a[:m]) + list(rng.choice(list("ACGT"), size=max(0, m - n)))
        for k in rng.integers(0, len(b), size=len(b) // 10):
            b[k] = "ACGT"[int(rng.integers(4))]
        cases.append((a, "".join(b)))
    for a, b in cases:
        r = subprocess.run([exe, "_ratio", a, b], capture_output=True, text=True)
        assert r.returncode == 0
        assert abs(float(r.stdout) - call_model.fuzz_ratio(a, b)) < 1e-5, (a, b)
    assert abs(call_model.fuzz_ratio(*cases[0]) - 96.5517241) < 1e-6


@pytest.mark.gpu
def test_bam_reader_records_across_inflate_windows(exe, world, tmp_path):
    """The BAM reader parses records in place inside an inflated window and copies only the ones that
    straddle two windows; a one-block window forces that path on (nearly) every record. Same clusters,
    and `smooth` writes the same records."""
    base, _ = run_clusterer(exe, world, 4)
    env = dict(os.environ, SVB_BGZF_WINDOW="1")
    out = os.path.join(world["d"], "clusters_win.txt")
    r = subprocess.run([exe, "call", "--reference", world["fa"], "--bam", world["bam"], "--sfs", world["sfs"], "--threads", "4",
                        "--cluster-only", "--clusters", out], capture_output=True, text=True, env=env)
    assert r.returncode == 0, r.stderr
    assert open(out).read() == base
    outs = []
    for e in (os.environ, env):
        p = str(tmp_path / ("s%d.bam" % len(outs)))
        with open(p, "wb") as f:
            r = subprocess.run([exe, "smooth", "--reference", world["fa"], "--bam", world["bam"]], stdout=f, stderr=subprocess.PIPE, env=dict(e))
        assert r.returncode == 0
        from bam_writer import read_bam
        outs.append(read_bam(p))
    assert outs[0] == outs[1] and len(outs[0][2]) > 100
