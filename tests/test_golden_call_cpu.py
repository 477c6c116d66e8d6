"""CPU: the committed known-answer vectors of the `call` half (tests/golden/ksw_small.json, poa_small.json;
tests/golden/make_golden_call.py made them and says what they do and do not pin).  The oracle must still
compute them, the ksw2 scores must still be the optimum of the independent Gotoh statement, and the CUDA kernel
sources themselves -- ksw_kernel.cuh / poa_kernel.cuh compiled for the host against the warp emulator -- must
give the same answers before they ever meet a GPU (tests/test_gpu_zx_golden_call.py is the GPU twin)."""
import json
import os

import numpy as np

import oracle
from test_ksw_emul import emul as ksw_emul, run as ksw_run        # noqa: F401  (fixture + harness)
from test_poa_emul import emul as poa_emul, run as poa_run        # noqa: F401

HERE = os.path.dirname(os.path.abspath(__file__))


def enc(s):
    return np.array(["ACGTN".index(c) for c in s], np.uint8)


def load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        d = json.load(f)
    assert d["alphabet"] == "ACGTN" and len(d["cases"]) >= 20
    return d["cases"]


def cigar_string(cig):
    return "".join("%d%s" % (l, op) for l, op in cig)


def test_oracle_reproduces_the_ksw_vectors():
    for k, c in enumerate(load("ksw_small.json")):
        q, t = enc(c["q"]), enc(c["t"])
        sc, cig = oracle.ksw_extd2(q, t)
        assert (sc, cigar_string(cig)) == (c["score"], c["cigar"]), k
        assert oracle.affine2_score(q, t) == c["score"], k                 # independent optimum
        assert oracle.cigar_score(q, t, cig) == c["score"], k


def test_oracle_reproduces_the_poa_vectors():
    for k, c in enumerate(load("poa_small.json")):
        reads = [enc(r) for r in c["reads"]]
        cb = oracle.poa_consensus(reads, band=True)
        assert "".join("ACGTN"[int(b)] for b in cb) == c["consensus"], k
        exact = oracle.poa_consensus(reads, band=False)
        assert oracle.edit_distance(cb, exact) == c["edit_distance_to_exact_poa"] <= 0.01 * len(c["template"]) + 1, k


def test_emulated_ksw_kernel_gives_the_vectors(ksw_emul):                  # noqa: F811
    cases = load("ksw_small.json")
    pairs = [(enc(c["q"]), enc(c["t"])) for c in cases]
    for variant in (0, 1):
        got = ksw_run(ksw_emul, pairs, variant)
        for k, c in enumerate(cases):
            assert (got[k][0], cigar_string(got[k][1])) == (c["score"], c["cigar"]), (variant, k)


def test_emulated_poa_kernel_gives_the_vectors(poa_emul):                  # noqa: F811
    cases = load("poa_small.json")
    clusters = [[enc(r) for r in c["reads"]] for c in cases]
    got, status, cells = poa_run(poa_emul, clusters, 0)
    assert not status.any() and cells > 0
    for k, c in enumerate(cases):
        assert "".join("ACGTN"[int(b)] for b in got[k]) == c["consensus"], k
