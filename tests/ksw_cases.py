"""Pair generators shared by the ksw tests (SURVEY 8d config 5 shape at test sizes)."""
import numpy as np


def mutate(rng, t, max_indel=60, alphabet=4):
    q = t.copy()
    for _ in range(int(rng.integers(0, 4))):
        k = int(rng.integers(3))
        p = int(rng.integers(0, max(1, len(q))))
        if k == 0 and len(q):
            q[p] = rng.integers(0, 5)
        elif k == 1:
            q = np.insert(q, p, rng.integers(0, alphabet, size=int(rng.integers(1, max_indel))).astype(np.uint8))
        elif len(q) > 3:
            q = np.delete(q, slice(p, p + int(rng.integers(1, max_indel))))
    return np.ascontiguousarray(q, np.uint8)


def make_pairs(rng, n, max_len=140, min_len=1):
    out = []
    for it in range(n):
        tl = int(rng.integers(min_len, max_len))
        t = rng.integers(0, 4 if it % 3 else 5, size=tl).astype(np.uint8)
        q = mutate(rng, t, max_indel=max(2, max_len // 2))
        if it % 17 == 0:
            q = rng.integers(0, 4, size=int(rng.integers(1, max_len))).astype(np.uint8)   # unrelated
        if it % 23 == 0:
            t = np.tile(rng.integers(0, 4, size=3).astype(np.uint8), tl // 3 + 1)[:tl]       # tandem repeat
            q = np.delete(t, slice(tl // 3, tl // 3 + min(12, tl // 2))) if tl > 8 else t.copy()
        if len(q) == 0:
            q = t[:1].copy()
        out.append((np.ascontiguousarray(q, np.uint8), np.ascontiguousarray(t, np.uint8)))
    return out


def planted_pairs(rng, n, lo=100, hi=3000):
    """config-5 shape: target length ~logU[lo,hi]; query = target with one planted INS/DEL U[50,1000]
    (clamped to the target) + 0.2 % noise."""
    out = []
    for _ in range(n):
        tl = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        t = rng.integers(0, 4, size=tl).astype(np.uint8)
        q = t.copy()
        L = int(min(rng.integers(50, 1001), max(1, tl // 2)))
        p = int(rng.integers(1, max(2, tl - L - 1)))
        if rng.random() < 0.5:
            q = np.concatenate([q[:p], rng.integers(0, 4, size=L).astype(np.uint8), q[p:]])
        else:
            q = np.concatenate([q[:p], q[p + L:]])
        nn = rng.binomial(len(q), 0.002)
        pos = rng.integers(0, len(q), size=nn)
        q[pos] = (q[pos] + rng.integers(1, 4, size=nn)) % 4
        out.append((np.ascontiguousarray(q, np.uint8), t))
    return out
