"""CPU: oracle/call_oracle.c (Clusterer::run and Caller::pcall restated in C over aligned-pair vectors, OpenMP; the
checker and CPU baseline of the `call` side) against the literal Python transcriptions tests/cluster_model.py and
tests/call_model.py, and a scaled slice of the config-3 generator through oracle search -> cluster -> call."""
import numpy as np
import pytest

import call_model
import cluster_model
import oracle
from cluster_common import aln_batch, compare, ref_of
from common import oracle_index, fm_results
from sv_world import make_world
from svdss_b200 import capi, synth
from test_gpu_cluster_call import NT6, check_calls, expected_jobs


@pytest.mark.parametrize("threads,kw", [(1, {}), (4, {}), (3, dict(seed=33, sub_rate=0.002, indel_rate=0.004, n_svs=10, coverage=6))])
def test_orc_cluster_matches_the_transcription(tmp_path, threads, kw):
    w = make_world(str(tmp_path), **kw)
    recs, alns = aln_batch(w)
    res = oracle.cluster(alns, ref_of(w), threads=threads, clipped=True)
    exp = cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=threads)
    assert len(exp) >= 5
    compare(res, exp, recs, w["names"])


@pytest.mark.parametrize("tag_hp,useht", [(True, True), (True, False), (False, True)])
def test_orc_call_matches_the_model(tmp_path, tag_hp, useht):
    w = make_world(str(tmp_path), tag_hp=tag_hp, seed=75 + int(useht))
    recs, alns = aln_batch(w)
    ref = ref_of(w)
    res = oracle.cluster(alns, ref, threads=4)
    clusters = cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=4)
    jobs = expected_jobs(w, clusters, useht)
    offs = np.zeros(len(recs) + 1, np.int64)
    offs[1:] = np.cumsum([len(r["seq"]) for r in recs])
    nt6 = np.array([NT6[c] for r in recs for c in r["seq"]], np.uint8)
    check_calls(oracle.call(res, capi.ReadSeqs(nt6, offs[:-1], capi.SVB_SEQ_NT6), ref, useht=useht), res, recs, jobs)


def region_world(ref_bp=600_000, n_reads=900, n_svs=24, seed=11):
    """a scaled config-3 slice: the bench's generators on a small reference"""
    contigs = synth.make_reference(ref_bp, seed=seed, contigs=3, n_repeats=6, n_nruns=2, nrun_len=100)
    cat_ref, coffs = synth.concat(contigs)
    cat = synth.make_sv_catalogue_arrays(coffs, n_svs=n_svs, seed=seed + 1, min_len=50, max_len=1500)
    # the generator keeps SVs 30 kb off contig ends; at this size that leaves few: relax by re-deriving for a small reference
    S = synth.make_sample_region(coffs, cat, n_reads, 0, coverage=20.0, seed=seed + 2, mean_len=8000, sd_len=1500, min_len=2000, max_len=14000)
    reads = synth.materialize_segments_numpy(cat_ref, S["segs"])
    return contigs, cat_ref, coffs, cat, S, reads


def test_config3_generator_through_the_oracle_pipeline():
    contigs, cat_ref, coffs, cat, S, reads = region_world()
    assert len(cat["gpos"]) >= 3 and (S["xf"] == 0).sum() >= 20
    ro = S["segs"]["read_offs"]
    # every searched read spells its CIGAR: M blocks equal the reference, I blocks equal the catalogue allele
    for k, a in enumerate(S["searched"][:200]):
        seq = reads[ro[k]:ro[k + 1]]
        assert len(seq) == S["l_qseq"][a]
        q, r = 0, int(coffs[S["tid"][a]] + S["pos"][a])
        for c in S["cigar"][S["cigar_offs"][a]:S["cigar_offs"][a + 1]]:
            ln, op = int(c >> 4), int(c & 15)
            if op == 0:
                assert np.array_equal(seq[q:q + ln], cat_ref[r:r + ln]); q += ln; r += ln
            elif op in (1, 4):
                q += ln
            else:
                r += ln
    # search with the oracle, then cluster + call on the CPU
    T, SA, bwt = oracle_index(contigs)
    fm = oracle.FMIndex(bwt)
    res, _ = fm_results(fm, [reads[ro[k]:ro[k + 1]] for k in range(len(S["searched"]))])
    n = len(S["tid"])
    sfs_offs = np.zeros(n + 1, np.int64)
    qs, ln = [], []
    per = {int(a): oracle.assemble(res[k]) for k, a in enumerate(S["searched"])}
    for a in range(n):
        for q, l in per.get(a, []):
            qs.append(q); ln.append(l)
        sfs_offs[a + 1] = len(qs)
    alns = capi.AlnBatch(S["tid"], S["pos"], S["hp"], S["cigar_offs"], S["cigar"], sfs_offs, qs, ln)
    ref = capi.RefSeqs(cat_ref, coffs[:-1], np.diff(coffs), capi.SVB_SEQ_NT6)
    cl = oracle.cluster(alns, ref, threads=4)
    seq_offs = np.full(n, -1, np.int64)
    seq_offs[S["searched"]] = ro[:-1]
    calls = oracle.call(cl, capi.ReadSeqs(reads, seq_offs, capi.SVB_SEQ_NT6), ref)
    # every planted SV that at least two reads carry comes back with exact type and length, at its anchor
    carried = np.bincount(S["sv_of_read"][S["sv_of_read"] >= 0], minlength=len(cat["gpos"]))
    want = {(int(cat["contig"][k]), bool(cat["is_del"][k]), int(cat["len"][k]), int(cat["pos"][k])) for k in np.nonzero(carried >= 2)[0]}
    got = {(int(cl.tid[calls.job_cluster[j]]), bool(t), int(l), int(p)) for j, t, l, p in zip(calls.sv_job, calls.sv_type, calls.sv_len, calls.sv_pos)}
    assert len(want) >= 3
    for tid, is_del, l, pos in want:
        assert any(g[0] == tid and g[1] == is_del and g[2] == l and abs(g[3] - (pos + 1)) <= l + 2 for g in got), (tid, is_del, l, pos, sorted(got))
    assert len(got) <= len(want) + 2


def test_batch_helpers_equal_their_scalar_forms():
    rng = np.random.default_rng(5)
    seqs, so, co = synth.gen_clusters_fast(12, seed=9, lo=60, hi=300)
    assert len(co) == 13 and (np.diff(co) >= 20).all() and seqs.max() <= 3
    cons = oracle.poa_batch(seqs, so, co, threads=2)
    for c in range(12):
        one = [seqs[so[i]:so[i + 1]] for i in range(co[c], co[c + 1])]
        assert np.array_equal(cons[c], oracle.poa_consensus(one, band=True))
        # half the reads differ from the template by one planted event: lengths within 3 %
        ls = np.diff(so)[co[c]:co[c + 1]]
        assert ls.max() - ls.min() <= 0.07 * ls.max() + 2
    qc, qo, tc, to = synth.gen_pairs_fast(40, seed=10, lo=100, hi=900)
    sc, cells = oracle.ksw_batch(qc, qo, tc, to, threads=2)
    assert cells == int((np.diff(qo) * np.diff(to)).sum())
    for p in range(40):
        s1, cig = oracle.ksw_extd2(qc[qo[p]:qo[p + 1]], tc[to[p]:to[p + 1]])
        assert s1 == sc[p]
        big = [l for l, op in cig if op != "M" and l >= 40]
        assert len(big) >= 1                                   # the planted INS / DEL shows in the alignment
    # assemble_batch on emit-order records
    offs, qs, ln = [0], [], []
    for _ in range(50):
        k = int(rng.integers(0, 6))
        q = np.sort(rng.choice(500, size=k, replace=False))[::-1]
        qs += q.tolist(); ln += rng.integers(1, 80, k).tolist(); offs.append(len(qs))
    aq, al, cnt = oracle.assemble_batch(offs, qs, ln)
    m = 0
    for r in range(50):
        exp = oracle.assemble(list(zip(qs[offs[r]:offs[r + 1]], ln[offs[r]:offs[r + 1]])))
        assert list(zip(aq[m:m + cnt[r]].tolist(), al[m:m + cnt[r]].tolist())) == exp
        m += int(cnt[r])
