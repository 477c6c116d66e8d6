"""CPU: the POA kernel source itself (svdss_b200/csrc/poa_kernel.cuh) compiled for the host with the
lock-step warp emulator of tests/emul/ -- the default kernel k_poa<0> and the variants selected by
SVB_POA_VARIANT (bit 1 previous row in shared memory, 2 first predecessor from in1 in the traceback,
4 remain[] / re-rank by the whole warp, 8 windowed graph update, 16 windowed traceback, 32 leaner DP rows) -- must give the banded oracle's
consensus, and the same cell count as each other.  This is how a kernel variant written without GPU
time gets checked before it is ever launched."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from poa_cases import make_cluster

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emul", "poa_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libpoa_emul.so")
    deps = [src, os.path.join(HERE, "emul", "warp_emul.hpp"), os.path.join(ROOT, "svdss_b200", "csrc", "poa_kernel.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = C.CDLL(out)
    lib.emul_poa.restype = C.c_int
    return lib


def run(lib, clusters, smem, worst_case=False, group=32, swcap=0, ncap_limit=None):
    seqs = [np.ascontiguousarray(s, np.uint8) for cl in clusters for s in cl]
    so = np.zeros(len(seqs) + 1, np.int64)
    so[1:] = np.cumsum([len(s) for s in seqs])
    cat = np.concatenate(seqs + [np.zeros(1, np.uint8)])
    co = np.zeros(len(clusters) + 1, np.int64)
    co[1:] = np.cumsum([len(cl) for cl in clusters])
    # capacities as svb_poa_batch computes them (pass 0 heuristic / pass 1 worst case)
    ncap = wcap = 4
    lmax = 1
    for cl in clusters:
        ls = [len(s) for s in cl if len(s)]
        if not ls:
            continue
        lmax = max(lmax, max(ls))
        w = 10 + int(0.01 * max(ls))
        if worst_case:
            nc, wc = sum(ls) + 2, max(ls) + 1
        else:
            nc = min(sum(ls) + 2, 2 * max(ls) + 32 * len(ls) + 64)
            wc = min(max(ls) + 1, 2 * w + 1 + 4 * (max(ls) - min(ls)) + 96)
        ncap, wcap = max(ncap, nc), max(wcap, wc)
    if ncap_limit:
        ncap = min(ncap, ncap_limit)
    wcap = (wcap + 31) & ~31
    ecap = 3 * ncap + 64
    cap = np.zeros(len(clusters) + 1, np.int64)
    cap[1:] = np.cumsum([(2 * max([len(s) for s in cl] + [0]) + 64 + 15) & ~15 for cl in clusters])
    cons = np.zeros(int(cap[-1]) + 1, np.uint8)
    clen = np.zeros(len(clusters), np.int32)
    status = np.zeros(len(clusters), np.int32)
    cells = C.c_ulonglong(0)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    rc = lib.emul_poa(p(cat), p(so), p(co), len(clusters), int(smem), int(group), ncap, ecap, wcap, lmax, int(swcap), p(cons), p(cap), p(clen), p(status), C.byref(cells))
    assert rc == 0, "emulated warp deadlocked or shared buffer too small (%d)" % rc
    return [cons[int(cap[i]):int(cap[i]) + int(clen[i])].copy() for i in range(len(clusters))], status, cells.value


def clusters_small(seed, n):
    rng = np.random.default_rng(seed)
    out = [make_cluster(rng, n_reads=int(rng.integers(2, 9)), tlen=int(rng.integers(30, 160)), rate=0.02)[1] for _ in range(n)]
    out.append([out[0][0]])                              # single read
    out.append([])                                       # empty cluster
    out.append([np.zeros(0, np.uint8), out[1][0], out[1][0]])
    return out


@pytest.mark.parametrize("smem", [0, 1, 2, 4, 32, 39, 68, 71, 135, 199, 263, 455, 487, 512, 1024, 967, 1479])
def test_emulated_kernel_equals_banded_oracle(emul, smem):
    clusters = clusters_small(41, 14 if smem in (0, 39, 455, 487, 967, 1479) else 5)
    got, status, cells = run(emul, clusters, smem)
    assert not status.any() and cells > 0
    for c, reads in enumerate(clusters):
        exp = oracle.poa_consensus(reads, band=True) if reads else np.zeros(0, np.uint8)
        assert np.array_equal(got[c], exp), (c, len(got[c]), len(exp))


def test_variants_agree_on_wider_rows_and_planted_alleles(emul):
    """reads long enough for rows of more than one 32-column step, planted indels (multi-predecessor rows,
    bands that move left and right), an SV allele carried by the majority"""
    rng = np.random.default_rng(42)
    clusters = [make_cluster(rng, n_reads=int(rng.integers(4, 8)), tlen=int(rng.integers(250, 420)), rate=0.01)[1] for _ in range(3)]
    t = rng.integers(0, 4, size=300).astype(np.uint8)
    alt = np.concatenate([t[:140], rng.integers(0, 4, size=40).astype(np.uint8), t[140:]])
    clusters.append([alt, t, alt, alt, t, alt])
    a, sa, ca = run(emul, clusters, 0)
    for variant in (7, 39, 71, 135, 263, 455, 487, 512, 1024, 967, 1479):
        b, sb, cb = run(emul, clusters, variant)
        assert ca == cb and np.array_equal(sa, sb)
        for c, reads in enumerate(clusters):
            assert np.array_equal(a[c], b[c]), (variant, c)
            assert np.array_equal(b[c], oracle.poa_consensus(reads, band=True)), (variant, c)
        assert np.array_equal(b[-1], alt)


def test_overflow_status_and_worst_case_rerun(emul):
    """a node capacity too small for one cluster: the kernel flags that cluster (status bit 1) and leaves the
    clusters next to it alone; with worst-case capacities (svb_poa_batch's second pass) it matches the oracle"""
    rng = np.random.default_rng(43)
    reads = [rng.integers(0, 4, size=int(rng.integers(60, 100))).astype(np.uint8) for _ in range(10)]
    _, ok = make_cluster(rng, n_reads=4, tlen=80)
    for smem, group in ((0, 32), (39, 32), (39, 8), (455, 32), (487, 8)):
        got1, status, _ = run(emul, [reads, ok, ok], smem, group=group, ncap_limit=150)   # too few nodes for the first cluster only
        assert status[0] != 0 and status[1] == 0 and status[2] == 0
        assert np.array_equal(got1[1], oracle.poa_consensus(ok, band=True)) and np.array_equal(got1[2], got1[1])
        got, status2, _ = run(emul, [reads], smem, worst_case=True, group=group)
        assert status2[0] == 0
        assert np.array_equal(got[0], oracle.poa_consensus(reads, band=True))


@pytest.mark.parametrize("group,variant", [(16, 0), (8, 0), (16, 39), (8, 39), (8, 7), (16, 455), (8, 487)])
def test_sub_warp_groups(emul, group, variant):
    """G lanes per cluster: 32/G clusters run side by side in one warp, each group with its own control flow
    (clusters of different sizes, so the groups diverge and finish at different times)"""
    rng = np.random.default_rng(44)
    clusters = clusters_small(45, 7)
    clusters += [make_cluster(rng, n_reads=5, tlen=int(t), rate=0.01)[1] for t in (260, 90, 330)]
    got, status, cells = run(emul, clusters, variant, group=group)
    ref, status0, cells0 = run(emul, clusters, 0)
    assert cells == cells0 and np.array_equal(status, status0)
    for c, reads in enumerate(clusters):
        assert np.array_equal(got[c], ref[c]), c
        exp = oracle.poa_consensus(reads, band=True) if reads else np.zeros(0, np.uint8)
        assert np.array_equal(got[c], exp), c


def test_rows_wider_than_the_shared_copy_fall_back_to_the_workspace(emul):
    """swcap smaller than most bands: rows alternate between the shared copy and the workspace"""
    rng = np.random.default_rng(46)
    clusters = [make_cluster(rng, n_reads=5, tlen=int(t), rate=0.01)[1] for t in (60, 200, 340)]
    ref, _, cells0 = run(emul, clusters, 0)
    for swcap, variant in ((24, 39), (33, 39), (24, 487), (33, 455), (24, 967), (33, 1479)):
        got, status, cells = run(emul, clusters, variant, swcap=swcap)
        assert cells == cells0 and not status.any()
        for c in range(len(clusters)):
            assert np.array_equal(got[c], ref[c]), (swcap, c)


def test_config4_shaped_cluster(emul):
    """SURVEY 8(d) config-4 shape (20-60 reads x 200-2000 bp, 0.1 % noise, half the reads with a planted indel):
    the default kernel and the most different build (all variant bits, 8 lanes per cluster)"""
    rng = np.random.default_rng(47)
    clusters = [make_cluster(rng, n_reads=21, tlen=700)[1], make_cluster(rng, n_reads=24, tlen=260)[1]]
    exp = [oracle.poa_consensus(c, band=True) for c in clusters]
    for variant, group in ((0, 32), (39, 8), (39, 32), (455, 16), (455, 32), (487, 32), (967, 32), (1479, 32)):
        got, status, _ = run(emul, clusters, variant, group=group)
        assert not status.any()
        assert all(np.array_equal(g, e) for g, e in zip(got, exp)), (variant, group)
