"""CPU-only: the ksw2 extd2 restatement (oracle/ksw_oracle.c) is pinned by an independent optimum
(min-cost Gotoh with two affine pieces), by re-scoring its own CIGAR, and by ksw2's documented
left-alignment of gaps."""
import numpy as np

import oracle
from ksw_cases import make_pairs


def test_score_is_the_global_optimum_and_cigar_rescored():
    rng = np.random.default_rng(3)
    for q, t in make_pairs(rng, 250, max_len=140):
        sc, cig = oracle.ksw_extd2(q, t)
        assert sc == oracle.affine2_score(q, t)
        assert oracle.cigar_score(q, t, cig) == sc
        assert sum(l for l, op in cig if op in "MI") == len(q)
        assert sum(l for l, op in cig if op in "MD") == len(t)
        for (l1, o1), (l2, o2) in zip(cig, cig[1:]):
            assert o1 != o2 and l1 > 0 and l2 > 0


def test_empty_inputs_follow_ksw_reset_extz():
    a = np.array([0, 1, 2], np.uint8)
    z = np.zeros(0, np.uint8)
    assert oracle.ksw_extd2(z, a) == (oracle.KSW_NEG_INF, [])
    assert oracle.ksw_extd2(a, z) == (oracle.KSW_NEG_INF, [])


def test_gap_left_alignment_and_long_gap_piece():
    enc = lambda s: oracle.CHAR26[np.frombuffer(s.encode(), np.uint8)]
    t = enc("ACGTACGT" + "AG" * 40 + "TTGACCA" * 3)
    q = np.concatenate([t[:18], t[78:]])          # 60 bp deleted inside the AG repeat
    sc, cig = oracle.ksw_extd2(q, t)
    assert cig == [(8, "M"), (60, "D"), (41, "M")]   # leftmost placement (flag lacks KSW_EZ_RIGHT)
    assert sc == 49 - (41 + 60)                      # long piece: min(16+2k, 41+k) at k=60
    ins = np.random.default_rng(0).integers(0, 4, 70).astype(np.uint8)
    q2 = np.concatenate([t[:30], ins, t[30:]])
    sc2, cig2 = oracle.ksw_extd2(q2, t)
    assert [op for _, op in cig2] == ["M", "I", "M"] and cig2[1][0] == 70
    # N scores -e2 = -1 against anything (mat[24] == 0, no KSW_EZ_GENERIC_SC)
    n = enc("ACGTNACGT")
    assert oracle.ksw_extd2(n, enc("ACGTAACGT"))[0] == 8 - 1
    assert oracle.ksw_extd2(n, n)[0] == 8 - 1
