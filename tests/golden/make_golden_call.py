"""Generates tests/golden/ksw_small.json and tests/golden/poa_small.json -- known-answer vectors for the `call`
half of the path.

The reference ships no golden vectors and ksw2 / abPOA are not vendored in its tree, so nothing here is the
output of the reference itself (parity stays "unpinned").  What the files do pin:
  * ksw2: every expected score is the optimum of an INDEPENDENT statement of the problem (min-cost Gotoh DP with
    two affine pieces, oracle.affine2_score) and the expected CIGAR re-scores to exactly that optimum; the
    CIGAR's tie-breaks are those of the ksw2 restatement (oracle/ksw_oracle.c: ksw_backtrack, gaps leftmost).
  * POA: the expected consensus is the banded restatement's (oracle/poa_oracle.c, abPOA's adaptive band and
    heaviest bundling with every tie-break fixed in its header); at generation time it must sit within the
    SURVEY 8(c) tolerance of the exact (un-banded) POA and, for the planted-allele clusters, be the allele.
The committed files also freeze the oracle: tests/test_golden_call_cpu.py fails if a later edit changes what it
computes.  Run:  python tests/golden/make_golden_call.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle  # noqa: E402
from ksw_cases import make_pairs, planted_pairs  # noqa: E402
from poa_cases import make_cluster  # noqa: E402

L = "ACGTN"


def dec(a):
    return "".join(L[int(x)] for x in a)


def ksw_cases():
    rng = np.random.default_rng(20240611)
    pairs = make_pairs(rng, 70, max_len=120)
    pairs += planted_pairs(rng, 6, lo=150, hi=500)
    t = rng.integers(0, 4, size=129).astype(np.uint8)                      # around the kernel's 128-row band edge
    pairs += [(t[:127].copy(), t), (np.concatenate([t, t[:40]]), t), (t[5:6].copy(), t[5:6].copy())]
    out = []
    for q, t in pairs:
        sc, cig = oracle.ksw_extd2(q, t)
        assert sc == oracle.affine2_score(q, t), "restatement is not optimal"
        assert oracle.cigar_score(q, t, cig) == sc
        out.append({"q": dec(q), "t": dec(t), "score": int(sc), "cigar": "".join("%d%s" % (l, op) for l, op in cig)})
    return out


def poa_cases():
    rng = np.random.default_rng(20240612)
    out = []
    for it in range(16):
        tpl, reads = make_cluster(rng, n_reads=int(rng.integers(2, 14)), tlen=int(rng.integers(40, 260)), rate=0.01)
        cb = oracle.poa_consensus(reads, band=True)
        cu = oracle.poa_consensus(reads, band=False)
        assert oracle.edit_distance(cb, cu) <= 0.01 * len(tpl) + 1
        out.append({"reads": [dec(r) for r in reads], "consensus": dec(cb), "template": dec(tpl),
                    "edit_distance_to_exact_poa": int(oracle.edit_distance(cb, cu))})
    # planted alleles: the majority carries a 60 bp insertion / a 45 bp deletion
    t = rng.integers(0, 4, size=300).astype(np.uint8)
    ins = np.concatenate([t[:120], rng.integers(0, 4, size=60).astype(np.uint8), t[120:]])
    dele = np.concatenate([t[:200], t[245:]])
    for allele, reads in ((ins, [ins, t, ins, ins, t, ins]), (dele, [dele, dele, t, dele]), (t, [t, ins, t, t, dele])):
        cb = oracle.poa_consensus(reads, band=True)
        assert np.array_equal(cb, allele) and np.array_equal(oracle.poa_consensus(reads, band=False), allele)
        out.append({"reads": [dec(r) for r in reads], "consensus": dec(cb), "template": dec(allele), "edit_distance_to_exact_poa": 0})
    out.append({"reads": [dec(t)], "consensus": dec(t), "template": dec(t), "edit_distance_to_exact_poa": 0})   # one read: itself
    return out


def main():
    k = ksw_cases()
    with open(os.path.join(HERE, "ksw_small.json"), "w") as f:
        json.dump({"parameters": "caller.cpp:333-337,348-349: match 1, mismatch -9, N 0 in the table but scored -e2, gap min(16+2k, 41+k), global, full band",
                   "alphabet": L, "cases": k}, f, indent=0)
    p = poa_cases()
    with open(os.path.join(HERE, "poa_small.json"), "w") as f:
        json.dump({"parameters": "caller.cpp:259-271 over abpoa_init_para defaults: global, match 2, mismatch 4, gaps 4/2 and 24/1, band 10 + 0.01 * qlen, heaviest bundling, one consensus",
                   "alphabet": L, "cases": p}, f, indent=0)
    print("ksw pairs: %d, POA clusters: %d" % (len(k), len(p)))


if __name__ == "__main__":
    main()
