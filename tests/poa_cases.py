"""Cluster generators for the POA tests (SURVEY 8d config 4 shape at test sizes)."""
import numpy as np


def noisy_copy(rng, tpl, rate=0.001, indel=None):
    r = tpl.copy()
    if indel is not None:
        kind, p, L = indel
        if kind == "I":
            r = np.concatenate([r[:p], rng.integers(0, 4, size=L).astype(np.uint8), r[p:]])
        else:
            r = np.concatenate([r[:p], r[p + L:]])
    n = rng.binomial(len(r), rate)
    for p in sorted(rng.integers(0, max(1, len(r)), size=n).tolist(), reverse=True):
        k = int(rng.integers(3))
        if k == 0:
            r[p] = (r[p] + rng.integers(1, 4)) % 4
        elif k == 1:
            r = np.insert(r, p, rng.integers(0, 4))
        elif len(r) > 2:
            r = np.delete(r, p)
    return np.ascontiguousarray(r, np.uint8)


def make_cluster(rng, n_reads=None, tlen=None, rate=0.001, frac_indel=0.5, max_indel_frac=0.03):
    """reads = template + `rate` sub/indel noise; a fraction carries one planted INS or DEL whose
    length stays within the 0.97 length-ratio window of split_cluster_by_len (caller.cpp:78-97)"""
    n_reads = n_reads or int(rng.integers(20, 61))
    tlen = tlen or int(np.exp(rng.uniform(np.log(200), np.log(2000))))
    tpl = rng.integers(0, 4, size=tlen).astype(np.uint8)
    reads = []
    for _ in range(n_reads):
        indel = None
        if rng.random() < frac_indel:
            L = int(rng.integers(1, max(2, int(tlen * max_indel_frac))))
            indel = ("I" if rng.random() < 0.5 else "D", int(rng.integers(1, max(2, tlen - L - 1))), L)
        reads.append(noisy_copy(rng, tpl, rate, indel))
    return tpl, reads
