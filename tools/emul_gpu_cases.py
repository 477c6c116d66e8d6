"""The cluster / pair sets of the GPU parity tests (tests/test_gpu_poa.py, tests/test_gpu_ksw.py) run through the
warp emulator instead of a GPU: every kernel variant against the oracle, a few minutes of CPU.  What to run after
touching poa_kernel.cuh / ksw_kernel.cuh when no GPU is at hand.
  python tools/emul_gpu_cases.py [poa] [ksw]"""
import ctypes as C
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np  # noqa: E402
import oracle  # noqa: E402
from ksw_cases import make_pairs, planted_pairs  # noqa: E402
from poa_cases import make_cluster  # noqa: E402
import test_ksw_emul as TK  # noqa: E402
import test_poa_emul as TP  # noqa: E402


def build(name):
    src = os.path.join(ROOT, "tests", "emul", name + "_emul.cpp")
    out = os.path.join(ROOT, "tests", "emul", "_build", "lib%s_emul.so" % name)
    os.makedirs(os.path.dirname(out), exist_ok=True)
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = C.CDLL(out)
    getattr(lib, "emul_" + name).restype = C.c_int
    return lib


def poa():
    lib = build("poa")

    def check(name, clusters, variants):
        exp = [oracle.poa_consensus(c, band=True) if c else np.zeros(0, np.uint8) for c in clusters]
        for variant, group in variants:
            t = time.time()
            got, status, _ = TP.run(lib, clusters, variant, group=group)
            redo = [c for c in range(len(clusters)) if status[c]]
            if redo:   # second pass of svb_poa_batch: worst-case capacities
                got2, status2, _ = TP.run(lib, [clusters[c] for c in redo], variant, group=group, worst_case=True)
                assert not status2.any()
                for k, c in enumerate(redo):
                    got[c] = got2[k]
            bad = [c for c in range(len(clusters)) if not np.array_equal(got[c], exp[c])]
            print("poa %-8s variant %2d group %2d  mismatches %s  reruns %d  %.0f s" % (name, variant, group, bad, len(redo), time.time() - t), flush=True)
            assert not bad
    rng = np.random.default_rng(21)      # test_small_clusters_bit_exact_vs_banded_oracle
    clusters, tpls = [], []
    for _ in range(60):
        tpl, reads = make_cluster(rng, n_reads=int(rng.integers(2, 25)), tlen=int(rng.integers(40, 500)), rate=0.01)
        clusters.append(reads); tpls.append(tpl)
    clusters += [[tpls[0]], [], [np.zeros(0, np.uint8), tpls[1], tpls[1]]]
    check("small", clusters, [(0, 32), (63, 32), (63, 8)])
    rng = np.random.default_rng(22)      # test_config4_shape_tolerance_and_alleles
    check("config4", [make_cluster(rng)[1] for _ in range(24)], [(0, 32), (63, 16)])
    rng = np.random.default_rng(23)      # test_workspace_overflow_rerun
    reads = [rng.integers(0, 4, size=int(rng.integers(150, 260))).astype(np.uint8) for _ in range(12)]
    check("overflow", [reads, make_cluster(rng, n_reads=8, tlen=300)[1]], [(0, 32), (63, 8)])


def ksw():
    lib = build("ksw")

    def check(name, pairs, variants):
        exp = [oracle.ksw_extd2(q, t) if len(q) and len(t) else (-0x40000000, []) for q, t in pairs]
        for v in variants:
            t0 = time.time()
            got = TK.run(lib, pairs, v)
            bad = [i for i in range(len(pairs)) if got[i] != exp[i]]
            print("ksw %-10s variant %d  mismatches %s  %.0f s" % (name, v, bad, time.time() - t0), flush=True)
            assert not bad
    check("small600", make_pairs(np.random.default_rng(31), 600, max_len=150), [0, 3])
    check("planted40", planted_pairs(np.random.default_rng(33), 40, lo=100, hi=3000), [0, 1, 2, 3])
    check("mid200", make_pairs(np.random.default_rng(34), 200, max_len=300, min_len=20), [0, 3])


if __name__ == "__main__":
    what = sys.argv[1:] or ["poa", "ksw"]
    if "ksw" in what:
        ksw()
    if "poa" in what:
        poa()
