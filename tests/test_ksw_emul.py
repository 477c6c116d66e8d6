"""CPU: the ksw2 kernel source itself (svdss_b200/csrc/ksw_kernel.cuh) compiled for the host with the
lock-step warp emulator of tests/emul/ -- the default kernel and the SVB_KSW_VARIANT bits (1: backtrack in
windows of 32 steps with the traceback bytes prefetched, 2: checkpointed traceback, bands recomputed during the
backtrack) -- must give the oracle's score and CIGAR."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emul", "ksw_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libksw_emul.so")
    deps = [src, os.path.join(HERE, "emul", "warp_emul.hpp"), os.path.join(ROOT, "svdss_b200", "csrc", "ksw_kernel.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = C.CDLL(out)
    lib.emul_ksw.restype = C.c_int
    return lib


def run(lib, pairs, variant):
    qs = [np.ascontiguousarray(q, np.uint8) for q, _ in pairs]
    ts = [np.ascontiguousarray(t, np.uint8) for _, t in pairs]
    qo = np.zeros(len(pairs) + 1, np.int64); qo[1:] = np.cumsum([len(x) for x in qs])
    to = np.zeros(len(pairs) + 1, np.int64); to[1:] = np.cumsum([len(x) for x in ts])
    qc = np.concatenate(qs + [np.zeros(1, np.uint8)]); tc = np.concatenate(ts + [np.zeros(1, np.uint8)])
    co = np.zeros(len(pairs) + 1, np.int64); co[1:] = np.cumsum([len(q) + len(t) + 2 for q, t in pairs])
    cig = np.zeros(int(co[-1]) + 1, np.uint32)
    ncig = np.zeros(len(pairs), np.int32)
    score = np.zeros(len(pairs), np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    k = oracle.KSW_PARAMS
    rc = lib.emul_ksw(p(qc), p(qo), p(tc), p(to), len(pairs), variant, k["a"], k["b"], k["sc_n"], k["q"], k["e"], k["q2"], k["e2"],
                      p(score), p(cig), p(co), p(ncig))
    assert rc == 0
    return [(int(score[i]), [(int(c >> 4), "MID"[int(c & 0xf)]) for c in cig[int(co[i]):int(co[i]) + int(ncig[i])]]) for i in range(len(pairs))]


def make_pairs(seed, n, lo, hi):
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        tl = int(rng.integers(lo, hi))
        t = rng.integers(0, 4, size=tl).astype(np.uint8)
        q = t.copy()
        kind = int(rng.integers(0, 4))
        L = int(rng.integers(3, max(4, tl // 3)))
        p = int(rng.integers(1, max(2, tl - L - 1)))
        if kind == 0:
            q = np.concatenate([t[:p], rng.integers(0, 4, size=L).astype(np.uint8), t[p:]])     # insertion (long ones take the second gap piece)
        elif kind == 1:
            q = np.concatenate([t[:p], t[p + L:]])                                               # deletion
        elif kind == 2:
            q = rng.integers(0, 4, size=int(rng.integers(1, hi))).astype(np.uint8)               # unrelated
        m = rng.random(len(q)) < 0.03
        q[m] = rng.integers(0, 5, size=int(m.sum()))                                             # substitutions, some N (code 4)
        out.append((q, t))
    return out


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_emulated_kernel_equals_oracle(emul, variant):
    pairs = make_pairs(51, 24, 1, 90) + make_pairs(52, 6, 120, 300)      # one band, and several bands (KBAND = 128 target rows)
    if variant & 2:                                                       # checkpointed traceback needs >= 4 bands (target > 384)
        pairs = pairs[:8] + make_pairs(53, 5, 390, 700)
        rng = np.random.default_rng(54)
        t = rng.integers(0, 4, size=600).astype(np.uint8)
        pairs.append((np.concatenate([t[:100], t[420:]]), t))             # a deletion longer than two bands: the path runs down one column
        pairs.append((np.concatenate([t[:300], rng.integers(0, 4, size=200).astype(np.uint8), t[300:]]), t))   # long insertion inside a band
        pairs.append((t[:40].copy(), t))                                  # short query against a long target
    pairs.append((np.zeros(0, np.uint8), pairs[0][1]))                   # empty query: KSW_NEG_INF, no CIGAR
    got = run(emul, pairs, variant)
    for i, (q, t) in enumerate(pairs):
        if len(q) == 0 or len(t) == 0:
            assert got[i] == (-0x40000000, []), i
            continue
        sc, cg = oracle.ksw_extd2(q, t)
        assert got[i] == (sc, cg), (i, len(q), len(t))
