"""Generates tests/golden/sfs_small.json.

The reference ships no golden vectors and cannot be built offline (ropebwt3 is not vendored), so
these fixtures come from tests/ref_model.py -- a pure-Python literal transcription of
ping_pong.cpp:4-49 over a naive bidirectional FMD index with rb3_fmd_set_intv / rb3_fmd_extend
semantics -- and every expected value is double-checked against the index-free definition
(brute-force substring tests).  Run:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import ref_model  # noqa: E402

L = "$ACGTN"


def dec(a):
    return "".join(L[int(x)] for x in a)


def comp(a):
    a = np.asarray(a, np.uint8)
    return np.where((a >= 1) & (a <= 4), 5 - a, a).astype(np.uint8)


def main():
    rng = np.random.default_rng(20240607)
    cases = []
    for ci in range(14):
        alpha = [1, 2, 3, 4] if ci % 3 else [1, 2, 3, 4, 5]
        ncont = int(rng.integers(1, 4))
        contigs = [rng.choice(alpha, size=int(rng.integers(30, 260))).astype(np.uint8) for _ in range(ncont)]
        if ci % 4 == 1:  # planted repeat + N run
            c = contigs[0]
            k = len(c) // 4
            c[-k:] = c[:k]
            c[len(c) // 2: len(c) // 2 + 7] = 5
        fmd = ref_model.NaiveFMD(contigs)
        reads = []
        for ri in range(12):
            c = contigs[int(rng.integers(ncont))]
            a = int(rng.integers(0, len(c) - 5))
            b = int(rng.integers(a + 5, len(c) + 1))
            r = c[a:b].copy()
            if rng.random() < 0.5:
                r = comp(r[::-1])
            for _ in range(int(rng.integers(0, 4))):
                if len(r) < 3:
                    break
                p = int(rng.integers(0, len(r)))
                k = int(rng.integers(3))
                if k == 0:
                    r[p] = rng.choice(alpha)
                elif k == 1:
                    r = np.insert(r, p, rng.choice(alpha, size=int(rng.integers(1, 9))))
                else:
                    r = np.delete(r, slice(p, p + int(rng.integers(1, 5))))
            if ri == 10:
                r = np.array([5] * 4 + list(r[:6]) + [5], np.uint8)   # N-rich read
            if ri == 11:
                r = r[:1]                                            # length-1 read
            reads.append(np.ascontiguousarray(r, np.uint8))
        exp = []
        for r in reads:
            got = ref_model.ping_pong_search(fmd, [int(x) for x in r] + [0])
            assert got == ref_model.sfs_definition(contigs, r)
            exp.append({"raw": got, "assembled": ref_model.assemble(got)})
        cases.append({"contigs": [dec(c) for c in contigs], "reads": [dec(r) for r in reads], "sfs": exp})
    with open(os.path.join(HERE, "sfs_small.json"), "w") as f:
        json.dump({"generator": "tests/golden/make_golden.py", "overlap": -1, "cases": cases}, f, indent=0)
    print("cases", len(cases), "sfs", sum(len(e["raw"]) for c in cases for e in c["sfs"]))


if __name__ == "__main__":
    main()
