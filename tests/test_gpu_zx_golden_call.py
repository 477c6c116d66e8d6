"""GPU: svb_ksw_extd2_batch and svb_poa_batch against the committed known-answer vectors of the `call` half
(tests/golden/ksw_small.json, poa_small.json; CPU twin: tests/test_golden_call_cpu.py) -- score and CIGAR
string as Caller::pcall builds it (caller.cpp:352-355), consensus string as Caller::run_poa returns it
(caller.cpp:295-297).  Nothing under oracle/ is used here: the fixtures are the checker."""
import json
import os

import numpy as np
import pytest

from svdss_b200 import capi, synth

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def enc(s):
    return np.array(["ACGTN".index(c) for c in s], np.uint8)


def load(name):
    with open(os.path.join(HERE, "golden", name)) as f:
        return json.load(f)["cases"]


def test_ksw_extd2_batch_gives_the_vectors():
    cases = load("ksw_small.json")
    qc, qo = synth.concat([enc(c["q"]) for c in cases])
    tc, to = synth.concat([enc(c["t"]) for c in cases])
    res = capi.ksw_extd2_batch(qc, qo, tc, to)
    assert res.n_pairs == len(cases)
    for k, c in enumerate(cases):
        assert (int(res.score[k]), res.cigar_string(k)) == (c["score"], c["cigar"]), k


def test_poa_batch_gives_the_vectors():
    cases = load("poa_small.json")
    res = capi.poa_batch([[enc(r) for r in c["reads"]] for c in cases])
    assert res.n_clusters == len(cases)
    for k, c in enumerate(cases):
        assert res.consensus_string(k) == c["consensus"], k
