"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

import oracle
from svdss_b200 import synth

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    with open(os.path.join(HERE, "golden", "sfs_small.json")) as f:
        g = json.load(f)
    cases = []
    for c in g["cases"]:
        contigs = [oracle.encode_nt6(s) for s in c["contigs"]]
        reads = [oracle.encode_nt6(s) for s in c["reads"]]
        raw = [[tuple(x) for x in e["raw"]] for e in c["sfs"]]
        asm = [[tuple(x) for x in e["assembled"]] for e in c["sfs"]]
        cases.append((contigs, reads, raw, asm))
    return cases


def random_case(rng, with_n=True, n_reads=40, max_contig=300):
    alpha = [1, 2, 3, 4, 5] if with_n else [1, 2, 3, 4]
    contigs = [rng.choice(alpha, size=int(rng.integers(20, max_contig))).astype(np.uint8)
               for _ in range(int(rng.integers(1, 4)))]
    reads = []
    for _ in range(n_reads):
        c = contigs[int(rng.integers(len(contigs)))]
        a = int(rng.integers(0, len(c)))
        b = int(rng.integers(a, len(c))) + 1
        r = c[a:b].copy()
        if rng.random() < 0.5:
            r = synth.revcomp6(r)
        for _m in range(int(rng.integers(0, 4))):
            if len(r) == 0:
                break
            p = int(rng.integers(0, len(r)))
            k = int(rng.integers(3))
            if k == 0:
                r[p] = rng.choice(alpha)
            elif k == 1:
                r = np.insert(r, p, rng.choice(alpha, size=int(rng.integers(1, 6))))
            else:
                r = np.delete(r, slice(p, p + int(rng.integers(1, 4))))
        reads.append(np.ascontiguousarray(r, np.uint8))
    return contigs, reads


def oracle_index(contigs):
    T = oracle.build_text(contigs)
    SA = oracle.suffix_array(T)
    bwt = oracle.bwt_from_sa(T, SA)
    return T, SA, bwt


def fm_results(fm, reads):
    cat, offs = oracle.concat(reads)
    counts, ooff, qs, ln, ext = fm.search_batch(cat, offs)
    return [list(zip(qs[ooff[i]:ooff[i + 1]].tolist(), ln[ooff[i]:ooff[i + 1]].tolist()))
            for i in range(len(reads))], ext
