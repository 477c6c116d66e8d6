"""GPU parity: svb_poa_batch vs the POA restatement. The kernel follows the banded oracle's rules, so
it must equal orc_poa(band=1) exactly; against the exact (un-banded) oracle the tolerance of
SURVEY 8(c) applies (normalised edit distance <= 1 %)."""
import numpy as np
import pytest

import oracle
from poa_cases import make_cluster
from svdss_b200 import capi

pytestmark = pytest.mark.gpu


def test_small_clusters_bit_exact_vs_banded_oracle():
    rng = np.random.default_rng(21)
    clusters, tpls = [], []
    for it in range(60):
        tpl, reads = make_cluster(rng, n_reads=int(rng.integers(2, 25)), tlen=int(rng.integers(40, 500)), rate=0.01)
        clusters.append(reads); tpls.append(tpl)
    clusters.append([tpls[0]])                       # single read
    clusters.append([])                              # empty cluster -> ""
    clusters.append([np.zeros(0, np.uint8), tpls[1], tpls[1]])
    res = capi.poa_batch(clusters)
    assert res.n_clusters == len(clusters)
    for c, reads in enumerate(clusters):
        exp = oracle.poa_consensus(reads, band=True) if reads else np.zeros(0, np.uint8)
        got = res.consensus(c)
        assert np.array_equal(got, exp), (c, len(got), len(exp), oracle.edit_distance(got, exp))
    assert res.consensus_string(len(clusters) - 3) == "".join("ACGTN"[b] for b in tpls[0])
    assert res.cells > 0


def test_config4_shape_tolerance_and_alleles():
    """20-60 reads x 200-2000 bp, 0.1 % noise, half the reads with a planted indel: consensus within
    1 % of the exact POA and equal to the banded oracle"""
    rng = np.random.default_rng(22)
    clusters = [make_cluster(rng)[1] for _ in range(24)]
    res = capi.poa_batch(clusters)
    for c, reads in enumerate(clusters):
        got = res.consensus(c)
        assert np.array_equal(got, oracle.poa_consensus(reads, band=True)), c
        exact = oracle.poa_consensus(reads, band=False)
        assert oracle.edit_distance(got, exact) <= 0.01 * len(exact) + 1
    # SV allele: majority of reads carries a 120 bp insertion -> consensus carries it (what ksw2
    # then reports as an INS, caller.cpp:371-383)
    t = rng.integers(0, 4, size=900).astype(np.uint8)
    alt = np.concatenate([t[:400], rng.integers(0, 4, size=120).astype(np.uint8), t[400:]])
    res2 = capi.poa_batch([[alt, t, alt, alt, t, alt, alt]])
    assert np.array_equal(res2.consensus(0), alt)


def test_workspace_overflow_rerun(monkeypatch):
    """divergent reads blow the heuristic node capacity; the library reruns those clusters with
    worst-case capacities and still matches the oracle"""
    rng = np.random.default_rng(23)
    reads = [rng.integers(0, 4, size=int(rng.integers(150, 260))).astype(np.uint8) for _ in range(12)]  # unrelated
    _, ok = make_cluster(rng, n_reads=8, tlen=300)
    res = capi.poa_batch([reads, ok])
    assert np.array_equal(res.consensus(0), oracle.poa_consensus(reads, band=True))
    assert np.array_equal(res.consensus(1), oracle.poa_consensus(ok, band=True))
