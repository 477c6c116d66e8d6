"""CPU: the Clusterer kernels' source (svdss_b200/csrc/cluster_core.cuh + cluster_host.hpp) compiled for the host
(tests/emul/cluster_emul.cpp: the two kernels become loops over the same per-item functions) against the literal
Python transcription of clusterer.cpp in tests/cluster_model.py -- clusters, sub-read windows, coverage, RVEC,
clips and output order for several --threads."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import cluster_model
from cluster_common import aln_batch, compare, ref_of
from sv_world import make_world
from test_clipper_cpu import clip_world  # noqa: F401  (fixture: SFSs inside soft / hard clips, a chromosome without sequence)
from svdss_b200 import capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emul", "cluster_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libcluster_emul.so")
    deps = [src] + [os.path.join(ROOT, "svdss_b200", "csrc", f) for f in ("cluster_core.cuh", "cluster_host.hpp")] + \
        [os.path.join(ROOT, "include", "svdss_b200.h")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    return make_world(str(tmp_path_factory.mktemp("clue")))


@pytest.fixture(scope="module")
def world_untagged(tmp_path_factory):
    return make_world(str(tmp_path_factory.mktemp("clue2")), seed=91, tag_hp=False, n_svs=16, coverage=10)


@pytest.mark.parametrize("threads", [1, 3, 4])
def test_emulated_kernels_match_the_transcription(emul, world, threads):
    recs, alns = aln_batch(world)
    res = capi.cluster_batch(alns, ref_of(world), threads=threads, emul=emul)
    exp = cluster_model.run(world["records"], world["names"], world["ref_seqs"], world["sfs_by_read"], threads=threads)
    assert len(exp) >= 8
    compare(res, exp, recs, world["names"])


def test_untagged_world_and_flags(emul, world_untagged):
    w = world_untagged
    recs, alns = aln_batch(w, min_mapq=0)
    res = capi.cluster_batch(alns, ref_of(w), threads=2, emul=emul)
    exp = cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=2, min_mapq=0)
    compare(res, exp, recs, w["names"])
    recs, alns = aln_batch(w)
    res = capi.cluster_batch(alns, ref_of(w), threads=2, min_cluster_weight=100, emul=emul)
    assert not res.placed.any() and res.small_clusters == res.n


def test_clips_match_the_transcription(emul, world):
    recs, alns = aln_batch(world)
    res = capi.cluster_batch(alns, ref_of(world), threads=3, clipped=True, emul=emul)
    clips = []
    exp = cluster_model.run(world["records"], world["names"], world["ref_seqs"], world["sfs_by_read"], threads=3, clips_out=clips)
    compare(res, exp, recs, world["names"])
    # the reference's order: accepted reads dealt to thread slots, per-slot vectors inserted at the front (clusterer.cpp:24)
    acc = [i for i in range(alns.n) if alns.sfs_offs[i + 1] > alns.sfs_offs[i]]
    slots = [[] for _ in range(3)]
    for n, i in enumerate(acc):
        lp, ll, rp, rl = (int(x) for x in res.clip[i])
        if ll > 0:
            slots[n % 3].append((recs[i]["qname"], world["names"][recs[i]["tid"]], lp, ll, True))
        if rl > 0:
            slots[n % 3].append((recs[i]["qname"], world["names"][recs[i]["tid"]], rp, rl, False))
    got = []
    for t in range(3):
        got[0:0] = slots[t]
    assert len(clips) >= 3 and got == clips


def test_descending_sfs_order_is_the_reference_quirk(emul, world):
    # --noassemble .sfs files list a read's SFSs in descending qs: the scan of clusterer.cpp:183 never looks left of
    # the previous hit, so all but the first are placed differently -- the restatement must follow, not "fix" it
    rev = {k: list(reversed(v)) for k, v in world["sfs_by_read"].items()}
    recs, alns = aln_batch(world, sfs_by_read=rev)
    res = capi.cluster_batch(alns, ref_of(world), threads=2, emul=emul)
    exp = cluster_model.run(world["records"], world["names"], world["ref_seqs"], rev, threads=2)
    compare(res, exp, recs, world["names"])


def test_raw_reads_with_indel_columns_in_the_flanks(emul, tmp_path):
    # raw-HiFi-shaped alignments: hundreds of CIGAR ops per read, 1-bp I/D columns inside the 100-column flanks
    w = make_world(str(tmp_path), seed=33, sub_rate=0.002, indel_rate=0.004, n_svs=10, coverage=6)
    recs, alns = aln_batch(w)
    assert np.diff(alns.cigar_offs).max() > 30
    res = capi.cluster_batch(alns, ref_of(w), threads=4, emul=emul)
    exp = cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=4)
    assert len(exp) >= 5
    compare(res, exp, recs, w["names"])


@pytest.mark.parametrize("threads,clipped", [(1, True), (3, True), (4, False)])
def test_clip_world_counters_and_clips(emul, clip_world, threads, clipped):
    w = clip_world
    recs = [r for r in w["records"] if cluster_model.primary(r) and r.get("mapq", 60) >= 20]
    sfs = {i: [(q, l) for q, l, _ in w["sfs_by_read"][r["qname"]]] for i, r in enumerate(recs) if r["qname"] in w["sfs_by_read"]}
    alns = capi.AlnBatch.from_records(recs, sfs)
    seqs = [w["ref_seqs"].get(n, "") for n in w["names"]]
    ref = capi.RefSeqs.from_strings(w["names"], seqs)
    ref.len[[i for i, n in enumerate(w["names"]) if n not in w["ref_seqs"]]] = -1        # chrC: in the BAM header, not in the FASTA
    res = capi.cluster_batch(alns, ref, threads=threads, clipped=clipped, emul=emul)
    clips = [] if clipped else None
    exp = cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=threads, clips_out=clips)
    compare(res, exp, recs, w["names"])
    if clipped:
        assert (res.unplaced, res.s_unplaced, res.e_unplaced, res.unknown) == (1, 1, 1, 0)      # whole / hardL / hardR
        acc = [i for i in range(alns.n) if alns.sfs_offs[i + 1] > alns.sfs_offs[i]]
        slots = [[] for _ in range(threads)]
        for n, i in enumerate(acc):
            lp, ll, rp, rl = (int(x) for x in res.clip[i])
            if ll > 0:
                slots[n % threads].append((recs[i]["qname"], w["names"][recs[i]["tid"]], lp, ll, True))
            if rl > 0:
                slots[n % threads].append((recs[i]["qname"], w["names"][recs[i]["tid"]], rp, rl, False))
        got = []
        for t in range(threads):
            got[0:0] = slots[t]
        assert got == clips and len(clips) == 4 + 3 + 6 + 5 + 2 + 1
    else:
        assert (res.unplaced, res.s_unplaced, res.e_unplaced) == (1, 11, 18)
