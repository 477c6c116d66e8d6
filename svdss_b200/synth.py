"""Synthetic HiFi-shaped workloads of SURVEY.md section 8(d) (fixed seeds, numpy, host side).

The reference ships no data (tests/README.md points at an external tarball), so every parity case and
every bench line runs on these generators.  nt6 codes: $=0 A=1 C=2 G=3 T=4 N=5 (ping_pong.hpp:46-52).
The device-side generator used at full scale lives in csrc/synth.cu and follows the same recipe.
"""
import numpy as np


def comp6(a):
    a = np.asarray(a, np.uint8)
    return np.where((a >= 1) & (a <= 4), 5 - a, a).astype(np.uint8)


def revcomp6(a):
    return comp6(a[::-1])


def make_reference(length=1_000_000, seed=1, n_repeats=20, n_nruns=5, nrun_len=200,
                   contigs=1):
    """i.i.d. uniform ACGT + planted exact repeats (1-5 kb) + N runs. Returns list of nt6 arrays."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(1, 5, size=length, dtype=np.uint8)
    for _ in range(n_repeats):
        ln = int(rng.integers(1000, 5001)) if length >= 50_000 else int(rng.integers(20, max(21, length // 20)))
        if ln * 2 >= length:
            continue
        src = int(rng.integers(0, length - ln))
        dst = int(rng.integers(0, length - ln))
        seg = ref[src:src + ln].copy()
        if rng.random() < 0.5:
            seg = revcomp6(seg)
        ref[dst:dst + ln] = seg
    for _ in range(n_nruns):
        ln = min(nrun_len, max(1, length // 50))
        p = int(rng.integers(0, length - ln))
        ref[p:p + ln] = 5
    # split into contigs with GRCh38-like decreasing proportions
    if contigs <= 1:
        return [ref]
    w = np.linspace(2.0, 0.5, contigs)
    cuts = np.floor(np.cumsum(w / w.sum()) * length).astype(np.int64)
    cuts[-1] = length
    out, b = [], 0
    for c in cuts:
        out.append(ref[b:c].copy())
        b = int(c)
    return [c for c in out if len(c)]


def make_reads(contigs, n_reads=1000, seed=2, mean_len=15000, sd_len=2000, min_len=5000,
               max_len=25000, event_rate=0.5, raw_hifi=False):
    """Smoothed-shaped reads: exact reference copies (either strand) with Poisson(event_rate) planted
    events per read (INS of i.i.d. bases / DEL, U[30,500] bp; or a 100-2000 bp random soft-clip
    tail).  raw_hifi adds 0.1% substitutions + 0.05% 1-bp indels.  Returns list of nt6 arrays."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for c in contigs], np.int64)
    reads = []
    for _ in range(n_reads):
        L = int(np.clip(rng.normal(mean_len, sd_len), min_len, max_len))
        ci = int(rng.choice(len(contigs), p=lens / lens.sum()))
        c = contigs[ci]
        L = min(L, len(c))
        st = int(rng.integers(0, len(c) - L + 1))
        r = c[st:st + L].copy()
        if rng.random() < 0.5:
            r = revcomp6(r)
        for _e in range(int(rng.poisson(event_rate))):
            kind = int(rng.integers(0, 3))
            if kind == 0 and len(r) > 700:      # INS
                p = int(rng.integers(100, len(r) - 100))
                ins = rng.integers(1, 5, size=int(rng.integers(30, 501)), dtype=np.uint8)
                r = np.concatenate([r[:p], ins, r[p:]])
            elif kind == 1 and len(r) > 1400:   # DEL
                d = int(rng.integers(30, 501))
                p = int(rng.integers(100, len(r) - d - 100))
                r = np.concatenate([r[:p], r[p + d:]])
            else:                               # soft-clip tail (random sequence)
                t = rng.integers(1, 5, size=int(rng.integers(100, 2001)), dtype=np.uint8)
                r = np.concatenate([r, t]) if rng.random() < 0.5 else np.concatenate([t, r])
        if raw_hifi:
            n_sub = rng.binomial(len(r), 0.001)
            pos = rng.integers(0, len(r), size=n_sub)
            r[pos] = ((r[pos] - 1 + rng.integers(1, 4, size=n_sub)) % 4 + 1).astype(np.uint8)
            n_id = rng.binomial(len(r), 0.0005)
            for p in sorted(rng.integers(1, len(r) - 1, size=n_id).tolist(), reverse=True):
                if rng.random() < 0.5:
                    r = np.delete(r, p)
                else:
                    r = np.insert(r, p, rng.integers(1, 5, dtype=np.uint8))
        reads.append(np.ascontiguousarray(r, np.uint8))
    return reads


def concat(seqs):
    offs = np.zeros(len(seqs) + 1, np.int64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs])
    cat = (np.ascontiguousarray(np.concatenate(seqs), np.uint8) if len(seqs) and offs[-1]
           else np.zeros(0, np.uint8))
    return cat, offs
