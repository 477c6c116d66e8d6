"""GPU parity: svb_ksw_extd2_batch vs the ksw2 restatement -- scores AND CIGARs bit-exact."""
import numpy as np
import pytest

import oracle
from ksw_cases import make_pairs, planted_pairs
from svdss_b200 import capi

pytestmark = pytest.mark.gpu


def _run(pairs, **kw):
    qc, qo = oracle.concat([p[0] for p in pairs])
    tc, to = oracle.concat([p[1] for p in pairs])
    return capi.ksw_extd2_batch(qc, qo, tc, to, **kw)


def _check(pairs, res, **kw):
    assert res.n_pairs == len(pairs)
    for i, (q, t) in enumerate(pairs):
        sc, cig = oracle.ksw_extd2(q, t, **kw)
        assert int(res.score[i]) == sc, (i, len(q), len(t))
        assert res.cigar_of(i) == cig, (i, len(q), len(t))


def test_small_random_pairs_exact():
    rng = np.random.default_rng(11)
    pairs = make_pairs(rng, 600, max_len=150)
    pairs += [(np.zeros(0, np.uint8), np.array([1, 2], np.uint8)), (np.array([1], np.uint8), np.zeros(0, np.uint8)),
              (np.array([2], np.uint8), np.array([2], np.uint8)), (np.array([2], np.uint8), np.array([3], np.uint8))]
    res = _run(pairs)
    _check(pairs, res)
    assert res.cigar_string(len(pairs) - 2) == "1M"


def test_band_boundaries_and_multi_band_pairs():
    """targets around the 128-row band edge and several bands deep; queries around the 32-column
    prefetch edge"""
    rng = np.random.default_rng(12)
    pairs = []
    for tl in (1, 3, 4, 5, 127, 128, 129, 255, 256, 257, 384, 700):
        for ql in (1, 31, 32, 33, 64, 65, 300):
            t = rng.integers(0, 4, size=tl).astype(np.uint8)
            if ql <= tl:
                s = int(rng.integers(0, tl - ql + 1))
                q = t[s:s + ql].copy()
            else:
                q = np.concatenate([t, rng.integers(0, 4, size=ql - tl).astype(np.uint8)])
            if ql > 4:
                q[int(rng.integers(0, ql))] = 4
            pairs.append((np.ascontiguousarray(q, np.uint8), t))
    _check(pairs, _run(pairs))


def test_planted_sv_pairs_config5_shape():
    rng = np.random.default_rng(13)
    pairs = planted_pairs(rng, 300, lo=100, hi=3000)
    res = _run(pairs)
    _check(pairs, res)
    # the planted event comes out as ONE long gap (two-piece cost crossover at k = 25)
    big = sum(1 for i in range(len(pairs)) if any(l >= 25 and op in "ID" for l, op in res.cigar_of(i)))
    assert big > 0.9 * len(pairs)


def test_other_parameters_and_waves(monkeypatch):
    rng = np.random.default_rng(14)
    pairs = make_pairs(rng, 200, max_len=300, min_len=20)
    kw = dict(a=2, b=-4, sc_n=-1, q=4, e=2, q2=24, e2=1)   # abPOA-like convex penalties
    res = _run(pairs, match=2, mismatch=-4, sc_n=-1, gapo=4, gape=2, gapo2=24, gape2=1)
    _check(pairs, res, **kw)
    # tiny traceback budget -> many waves, same answers
    monkeypatch.setenv("SVB_KSW_TB_BYTES", str(1 << 20))
    res2 = _run(pairs)
    assert res2.waves > 1
    _check(pairs, res2)


def test_large_pair_properties():
    """a 6 kb x 6.5 kb pair: score equals the independent optimum, CIGAR re-scores to it"""
    rng = np.random.default_rng(15)
    t = rng.integers(0, 4, size=6500).astype(np.uint8)
    q = np.concatenate([t[:2000], t[2500:]])
    q[rng.integers(0, len(q), size=12)] = 0
    res = _run([(q, t)])
    sc = int(res.score[0])
    assert sc == oracle.affine2_score(q, t)
    assert oracle.cigar_score(q, t, res.cigar_of(0)) == sc
    assert any(l == 500 and op == "D" for l, op in res.cigar_of(0))
