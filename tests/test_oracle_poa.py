"""CPU-only: the POA restatement (oracle/poa_oracle.c). Exact (un-banded) consensus recovers the
template on noisy clusters; the adaptive band stays within the 1 % tolerance of the exact one."""
import numpy as np

import oracle
from poa_cases import make_cluster, noisy_copy


def test_identical_and_single_reads():
    rng = np.random.default_rng(0)
    t = rng.integers(0, 4, size=257).astype(np.uint8)
    assert np.array_equal(oracle.poa_consensus([t]), t)
    assert np.array_equal(oracle.poa_consensus([t, t, t], band=True), t)
    assert len(oracle.poa_consensus([np.zeros(0, np.uint8)])) == 0


def test_majority_wins_and_planted_indel():
    rng = np.random.default_rng(1)
    t = rng.integers(0, 4, size=400).astype(np.uint8)
    alt = np.concatenate([t[:150], rng.integers(0, 4, size=60).astype(np.uint8), t[150:]])   # 60 bp insertion allele
    reads = [alt, alt, alt, t, alt, t]
    for band in (False, True):
        c = oracle.poa_consensus(reads, band=band)
        assert np.array_equal(c, alt)
    reads = [t, t, alt, t]
    assert np.array_equal(oracle.poa_consensus(reads), t)
    sub = t.copy(); sub[77] = (sub[77] + 1) % 4
    assert np.array_equal(oracle.poa_consensus([sub, t, t]), t)


def test_noisy_clusters_exact_vs_banded_tolerance():
    rng = np.random.default_rng(2)
    for it in range(25):
        tpl, reads = make_cluster(rng, n_reads=int(rng.integers(3, 30)), tlen=int(rng.integers(60, 700)), rate=0.01)
        cu = oracle.poa_consensus(reads, band=False)
        cb, st = oracle.poa_consensus(reads, band=True, return_stats=True)
        assert oracle.edit_distance(cu, tpl) <= max(2, 0.02 * len(tpl))
        assert oracle.edit_distance(cb, cu) <= 0.01 * len(tpl) + 1          # SURVEY 8(c) tolerance
        assert st[1] >= len(tpl) // 2


def test_edit_distance():
    a = np.array([0, 1, 2, 3], np.uint8)
    assert oracle.edit_distance(a, a) == 0
    assert oracle.edit_distance(a, a[:2]) == 2
    assert oracle.edit_distance(a, np.array([0, 2, 2, 3, 3], np.uint8)) == 2
