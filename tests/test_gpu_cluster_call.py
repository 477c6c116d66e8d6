"""GPU: svb_cluster_batch (Clusterer::run, clusterer.cpp:8-52) and svb_call_batch (Caller::pcall, caller.cpp:311-406)
through the C ABI against the literal Python transcriptions (tests/cluster_model.py, tests/call_model.py over the
oracle's POA and ksw2): clusters, sub-reads, coverage, RVEC and order; then jobs, consensus, score, CIGAR and SV
records -- with the reads as host nt6 bytes, as BAM 4-bit host bytes and resident on the device, and the reference
as host ASCII and as the index's device-resident text."""
import numpy as np
import pytest

import call_model
import cluster_model
import oracle
from cluster_common import aln_batch, compare, ref_of
from sv_world import make_world
from svdss_b200 import capi

pytestmark = pytest.mark.gpu

NT6 = {c: i for i, c in enumerate("$ACGTN")}
NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    return make_world(str(tmp_path_factory.mktemp("gclu")))


@pytest.mark.parametrize("threads", [1, 4])
def test_cluster_batch_matches_the_transcription(world, threads):
    recs, alns = aln_batch(world)
    res = capi.cluster_batch(alns, ref_of(world), threads=threads, clipped=True)
    exp = cluster_model.run(world["records"], world["names"], world["ref_seqs"], world["sfs_by_read"], threads=threads)
    assert len(exp) >= 8 and res.launches == 3 and res.kernel_ms > 0
    compare(res, exp, recs, world["names"])


def test_cluster_batch_on_raw_reads(tmp_path):
    w = make_world(str(tmp_path), seed=33, sub_rate=0.002, indel_rate=0.004, n_svs=10, coverage=6)
    recs, alns = aln_batch(w)
    res = capi.cluster_batch(alns, ref_of(w), threads=3)
    compare(res, cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=3), recs, w["names"])


def test_cluster_batch_rejects_unsorted_and_handles_empty(world):
    recs, alns = aln_batch(world)
    bad = capi.AlnBatch(alns.tid[::-1], alns.pos[::-1], alns.hp, alns.cigar_offs, alns.cigar, alns.sfs_offs, alns.sfs_qs, alns.sfs_len)
    with pytest.raises(capi.SvbError):
        capi.cluster_batch(bad, ref_of(world))
    none = capi.AlnBatch(alns.tid, alns.pos, alns.hp, alns.cigar_offs, alns.cigar, np.zeros(alns.n + 1, np.int64), [], [])
    assert capi.cluster_batch(none, ref_of(world)).n == 0


def expected_jobs(world, clusters, useht):
    enc = lambda s: oracle.CHAR26[np.frombuffer(s.encode(), np.uint8)]
    jobs = []
    for ci, cl in enumerate(clusters):
        if len(cl["subreads"]) < 2:
            continue
        cov = (sum(cl["cov"]),) + tuple(cl["cov"])
        for sub, cv in call_model.split_cluster(cl["subreads"], cov, useht, np.float32(0.97)):
            cons = oracle.poa_consensus([enc(x[1]) for x in sub], band=True)
            window = world["ref_seqs"][cl["chrom"]][cl["s"]:cl["e"] + 1]
            score, cig = oracle.ksw_extd2(cons, enc(window))
            svs, rpos, cpos = [], cl["s"], 0
            for l, op in cig:
                if op == "M":
                    rpos += l; cpos += l
                elif op == "I":
                    if l >= 25:
                        svs.append((0, rpos, l, cpos))
                    cpos += l
                else:
                    if l >= 25:
                        svs.append((1, rpos, l, cpos))
                    rpos += l
            jobs.append(dict(cluster=ci, names=[x[0] for x in sub], cov=cv, cons=cons, score=score, cigar=cig, svs=svs))
    return jobs


def check_calls(calls, res, recs, jobs):
    placed = [c for c in range(res.n) if res.placed[c]]
    assert calls.n_jobs == len(jobs) and len(jobs) >= 8
    k = 0
    for j, e in enumerate(jobs):
        assert placed.index(int(calls.job_cluster[j])) == e["cluster"]
        subs = calls.job_sub[int(calls.job_sub_offs[j]):int(calls.job_sub_offs[j + 1])]
        assert [recs[int(res.sub_aln[s])]["qname"] for s in subs] == e["names"]
        assert tuple(int(x) for x in calls.job_cov[j]) == tuple(e["cov"])
        assert np.array_equal(calls.consensus(j), e["cons"])
        assert int(calls.score[j]) == e["score"] and calls.cigar_of(j) == e["cigar"]
        assert int(calls.job_nv[j]) == len(e["svs"])
        for sv in e["svs"]:
            assert (int(calls.sv_job[k]), int(calls.sv_type[k]), int(calls.sv_pos[k]), int(calls.sv_len[k]), int(calls.sv_cpos[k])) == (j,) + sv
            k += 1
    assert k == calls.n_svs and k >= 8


@pytest.mark.parametrize("tag_hp,useht", [(True, True), (True, False), (False, True)])
def test_call_batch_matches_the_model(tmp_path, tag_hp, useht):
    import torch
    w = make_world(str(tmp_path), tag_hp=tag_hp, seed=75 + int(useht))
    recs, alns = aln_batch(w)
    ref = ref_of(w)
    res = capi.cluster_batch(alns, ref, threads=4)
    clusters = cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=4)
    compare(res, clusters, recs, w["names"])
    jobs = expected_jobs(w, clusters, useht)
    # reads as host nt6 bytes
    offs = np.zeros(len(recs) + 1, np.int64)
    offs[1:] = np.cumsum([len(r["seq"]) for r in recs])
    nt6 = np.array([NT6[c] for r in recs for c in r["seq"]], np.uint8)
    check_calls(capi.call_batch(res, capi.ReadSeqs(nt6, offs[:-1], capi.SVB_SEQ_NT6), ref, useht=useht), res, recs, jobs)
    # as BAM stores them: 4 bits per base, every read on a byte boundary
    b4, o4 = [], [0]
    for r in recs:
        c = [NT16[x] for x in r["seq"]] + [0]
        b4 += [(c[i] << 4) | c[i + 1] for i in range(0, len(r["seq"]), 2)]
        o4.append(len(b4))
    check_calls(capi.call_batch(res, capi.ReadSeqs(np.array(b4, np.uint8), o4[:-1], capi.SVB_SEQ_BAM4), ref, useht=useht), res, recs, jobs)
    # resident on the device, next to the index's own copy of the reference
    cat, coffs = oracle.concat(w["contigs"])
    idx = capi.Index.build(cat, coffs, device=0)
    dref = idx.ref()
    assert np.array_equal(dref.len, np.diff(coffs))
    res_d = capi.cluster_batch(alns, dref, threads=4)
    compare(res_d, clusters, recs, w["names"])
    t = torch.from_numpy(nt6).cuda()
    some = offs[:-1].copy()
    used = set(int(a) for a in res.sub_aln)
    some[[i for i in range(len(recs)) if i not in used]] = -1          # reads that were never searched have no sequence
    calls = capi.call_batch(res_d, capi.ReadSeqs(t.data_ptr(), some, capi.SVB_SEQ_NT6, capi.SVB_MEM_DEVICE), dref, useht=useht)
    check_calls(calls, res, recs, jobs)
    assert calls.launches >= 4 and calls.poa_kernel_ms > 0 and calls.ksw_kernel_ms > 0
