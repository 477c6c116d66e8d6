"""CPU-only: the oracle's three statements of the SFS path agree with each other and with the
golden fixtures (see oracle/sfs_oracle.c header for why parity is otherwise unpinned)."""
import numpy as np
import pytest

import oracle
import ref_model
from common import load_golden, random_case, oracle_index, fm_results
from svdss_b200 import synth


def test_golden_fixtures_all_statements():
    n = 0
    for contigs, reads, raw, asm in load_golden():
        T, SA, bwt = oracle_index(contigs)
        fm = oracle.FMIndex(bwt)
        got_fm, _ = fm_results(fm, reads)
        for r, exp_raw, exp_asm, g in zip(reads, raw, asm, got_fm):
            assert oracle.sfs_spec(T, SA, r) == exp_raw
            assert g == exp_raw
            assert oracle.assemble(exp_raw) == exp_asm
            n += len(exp_raw)
    assert n > 300


@pytest.mark.parametrize("seed", range(4))
def test_literal_transcription_vs_definition_vs_c(seed):
    rng = np.random.default_rng(100 + seed)
    contigs, reads = random_case(rng, with_n=bool(seed % 2), n_reads=25, max_contig=160)
    fmd = ref_model.NaiveFMD(contigs)
    T, SA, bwt = oracle_index(contigs)
    assert SA.tolist() == fmd.sa
    assert bwt.tolist() == fmd.bwt
    fm = oracle.FMIndex(bwt)
    assert fm.acc.tolist() == fmd.acc
    got_fm, _ = fm_results(fm, reads)
    for r, g in zip(reads, got_fm):
        if len(r) == 0:
            assert g == []
            continue
        a = ref_model.ping_pong_search(fmd, [int(x) for x in r] + [0])
        assert a == ref_model.sfs_definition(contigs, r) == oracle.sfs_spec(T, SA, r) == g
        assert ref_model.assemble(a) == oracle.assemble(a)
        # starts and ends strictly decrease
        for x, y in zip(a, a[1:]):
            assert y[0] < x[0] and y[0] + y[1] < x[0] + x[1]


def test_rank2a_matches_naive_occ():
    rng = np.random.default_rng(7)
    contigs, _ = random_case(rng, with_n=True, n_reads=0)
    fmd = ref_model.NaiveFMD(contigs)
    T, SA, bwt = oracle_index(contigs)
    fm = oracle.FMIndex(bwt)
    n = len(T)
    for _ in range(200):
        k = int(rng.integers(0, n + 1))
        l = int(rng.integers(k, n + 1))
        ok, ol = fm.rank2a(k, l)
        assert ok.tolist() == fmd.occ[k].tolist() and ol.tolist() == fmd.occ[l].tolist()


def test_config1_spec_vs_port():
    """SURVEY 8(d) config 1 shape at reduced read count: 1 Mb reference with planted repeats and
    N runs, smoothed-shaped and raw-HiFi-shaped 15 kb reads; SA-narrowing spec == FM port."""
    contigs = synth.make_reference(1_000_000, seed=1)
    T, SA, bwt = oracle_index(contigs)
    fm = oracle.FMIndex(bwt)
    reads = synth.make_reads(contigs, 60, seed=2) + synth.make_reads(contigs, 20, seed=3, raw_hifi=True)
    got, ext = fm_results(fm, reads)
    assert ext > sum(len(r) for r in reads) * 0.9
    tot = 0
    for r, g in zip(reads, got):
        assert oracle.sfs_spec(T, SA, r) == g
        tot += len(g)
    assert tot > 100


def test_assemble_edge_cases():
    assert oracle.assemble([]) == []
    assert oracle.assemble([(5, 3)]) == [(5, 3)]
    # touching intervals are NOT merged (assembler.cpp:41 uses <=)
    assert oracle.assemble([(8, 2), (5, 3)]) == [(5, 3), (8, 2)]
    assert oracle.assemble([(7, 4), (5, 3)]) == [(5, 6)]
    assert ref_model.assemble([(7, 4), (5, 3), (20, 2), (10, 1)]) == oracle.assemble([(7, 4), (5, 3), (20, 2), (10, 1)])


def test_property_three_statements_agree():
    """hypothesis: on arbitrary small references and reads (N included, both strands, empty and
    one-base reads) the literal ping_pong.cpp transcription over a bidirectional FMD, the
    occurrence-only definition and the C oracle give the same SFS list, with strictly
    decreasing starts and ends, and every SFS is absent from the reference while both of its
    one-base-shorter ends occur."""
    from hypothesis import given, settings, strategies as st

    base = st.integers(min_value=1, max_value=5)
    contig = st.lists(base, min_size=1, max_size=60)

    @settings(max_examples=60, deadline=None)
    @given(st.lists(contig, min_size=1, max_size=3), st.lists(st.lists(base, min_size=0, max_size=50), min_size=1, max_size=4),
           st.data())
    def check(contigs, reads, data):
        contigs = [np.array(c, np.uint8) for c in contigs]
        # half of the reads are cut out of the reference (either strand) so that long matches occur
        for i in range(len(reads)):
            if data.draw(st.booleans()):
                c = contigs[data.draw(st.integers(0, len(contigs) - 1))]
                a = data.draw(st.integers(0, len(c) - 1)); b = data.draw(st.integers(a + 1, len(c)))
                piece = c[a:b] if data.draw(st.booleans()) else synth.revcomp6(c[a:b])
                reads[i] = reads[i][:len(reads[i]) // 2] + [int(x) for x in piece] + reads[i][len(reads[i]) // 2:]
        fmd = ref_model.NaiveFMD(contigs)
        T, SA, bwt = oracle_index(contigs)
        fm = oracle.FMIndex(bwt)
        arrs = [np.array(r, np.uint8) for r in reads]
        got_fm, _ = fm_results(fm, arrs)
        strands = [c.tobytes() for c in contigs] + [synth.revcomp6(c).tobytes() for c in contigs]
        occurs = lambda w: any(w.tobytes() in s for s in strands)
        for r, g in zip(arrs, got_fm):
            if len(r) == 0:
                assert g == []
                continue
            a = ref_model.ping_pong_search(fmd, [int(x) for x in r] + [0])
            assert a == ref_model.sfs_definition(contigs, r) == oracle.sfs_spec(T, SA, r) == g
            for x, y in zip(a, a[1:]):
                assert y[0] < x[0] and y[0] + y[1] < x[0] + x[1]
            for qs, ln in a:
                w = r[qs:qs + ln]
                assert not occurs(w) and (ln == 1 or (occurs(w[1:]) and occurs(w[:-1])))

    check()
