"""CPU: ropebwt3 `.fmd` support of the shell (svdss_b200/host/rld.hpp) through the `_fmd`
hook: the C++ writer and reader against each other and against the independent Python statement in
tests/rld_model.py, the frame table through a rank that enters the stream the way rld_locate_blk does,
and the BWT inversion against the oracle's suffix-array BWT of {S, rc(S)}.
Parity with ropebwt3 itself is unpinned (its source is not in the reference tree, see rld.hpp)."""
import os
import subprocess

import numpy as np
import pytest

import oracle
import rld_model
from common import oracle_index
from svdss_b200 import build, synth


@pytest.fixture(scope="module")
def exe():
    build.build_lib()
    return build.build_host()


def runs_to_bwt(rng, n_runs, mean_run, long_runs=()):
    parts, last = [], -1
    for k in range(n_runs):
        c = int(rng.integers(0, 6))
        if c == last:
            c = (c + 1) % 6
        last = c
        l = int(rng.geometric(1.0 / mean_run))
        parts.append(np.full(l, c, np.uint8))
    for pos, l, c in long_runs:
        parts.insert(pos, np.full(l, c, np.uint8))
    return np.concatenate(parts)


def cxx(exe, *args):
    r = subprocess.run([exe, "_fmd"] + list(args), capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


CASES = [("short_runs", 3000, 1.3, ()), ("genome_like", 20000, 1.5, ()), ("long_runs", 400, 60.0, ((10, 20000, 2), (200, 70000, 4), (390, 16384, 1))),
         ("single_symbol", 1, 5000.0, ()), ("tiny", 3, 1.0, ())]


@pytest.mark.parametrize("name,n_runs,mean_run,long_runs", CASES)
def test_writer_reader_and_python_model_agree(exe, tmp_path, name, n_runs, mean_run, long_runs):
    rng = np.random.default_rng(len(name) * 7 + n_runs)
    bwt = runs_to_bwt(rng, n_runs, mean_run, long_runs)
    raw, fmd, back = str(tmp_path / "b.bin"), str(tmp_path / "b.fmd"), str(tmp_path / "back.bin")
    bwt.tofile(raw)
    cxx(exe, "encode", raw, fmd)
    cxx(exe, "decode", fmd, back)
    assert np.array_equal(np.fromfile(back, np.uint8), bwt)                 # C++ writer -> C++ reader
    f = rld_model.parse(fmd)
    assert f["asize"] == 6 and f["sbits"] == 3 and sum(f["mcnt"]) == len(bwt)
    assert rld_model.decode(f) == bwt.tobytes()                             # C++ writer -> Python reader (block counters checked too)
    if long_runs:
        ssize = 8
        types = {f["words"][o] >> 62 for o in range(0, len(f["words"]) >> 3 << 3, ssize)}
        assert types == {0, 1}                                              # u16 and u32 counter blocks both occur
    py = str(tmp_path / "py.fmd")
    rld_model.encode(py, bwt.tolist())
    cxx(exe, "decode", py, back)
    assert np.array_equal(np.fromfile(back, np.uint8), bwt)                 # Python writer -> C++ reader
    assert rld_model.parse(py)["words"] == f["words"]                       # and the two writers emit the same stream


def test_frames_give_rank_entry_points(exe, tmp_path):
    rng = np.random.default_rng(9)
    bwt = runs_to_bwt(rng, 30000, 1.4, ((100, 30000, 3), (20000, 9000, 1)))
    raw, fmd = str(tmp_path / "b.bin"), str(tmp_path / "b.fmd")
    bwt.tofile(raw)
    cxx(exe, "encode", raw, fmd)
    f = rld_model.parse(fmd)
    assert f["n_frames"] > 20
    cum = np.zeros((6, len(bwt) + 1), np.int64)
    for c in range(6):
        cum[c, 1:] = np.cumsum(bwt == c)
    ks = list(rng.integers(0, len(bwt), 300)) + [0, 1, len(bwt) - 1]
    for k in ks:
        assert rld_model.rank_all(f, int(k)) == [int(cum[c, k]) for c in range(6)], k
    # frame offsets are block starts inside the stream and never the closing block
    last = len(f["words"]) >> 3 << 3
    for k in range(f["n_frames"]):
        assert f["frame"][k * 7] % 8 == 0 and f["frame"][k * 7] < max(last, 1)


def test_reader_rejects_damaged_files(exe, tmp_path):
    rng = np.random.default_rng(3)
    bwt = runs_to_bwt(rng, 2000, 1.5)
    raw, fmd, back = str(tmp_path / "b.bin"), str(tmp_path / "b.fmd"), str(tmp_path / "back.bin")
    bwt.tofile(raw)
    cxx(exe, "encode", raw, fmd)
    data = bytearray(open(fmd, "rb").read())
    for name, mutate in [("truncated", lambda d: d[:len(d) // 2]), ("magic", lambda d: b"RLD\2" + d[4:]),
                         ("counts", lambda d: d[:24] + bytes([d[24] ^ 1]) + d[25:]),
                         ("payload", lambda d: d[:24 + 48 + 16 * 8 + 3] + bytes([d[24 + 48 + 16 * 8 + 3] ^ 0x55]) + d[24 + 48 + 16 * 8 + 4:])]:
        bad = str(tmp_path / (name + ".fmd"))
        open(bad, "wb").write(bytes(mutate(bytes(data))))
        r = subprocess.run([exe, "_fmd", "decode", bad, back], capture_output=True, text=True)
        assert r.returncode == 1 and "[critical]" in r.stderr, name


@pytest.mark.parametrize("seed,n_contigs", [(1, 1), (2, 3), (3, 6)])
def test_fmd_of_a_collection_inverts_to_its_forward_strands(exe, tmp_path, seed, n_contigs):
    contigs = synth.make_reference(30_000, seed=seed, contigs=n_contigs, n_repeats=3, n_nruns=1, nrun_len=40)
    T, SA, bwt = oracle_index(contigs)                                       # text S0 $ rc(S0) $ S1 $ ..., sentinels by position
    raw, fmd, fa = str(tmp_path / "b.bin"), str(tmp_path / "b.fmd"), str(tmp_path / "c.fa")
    bwt.tofile(raw)
    cxx(exe, "encode", raw, fmd)
    cxx(exe, "contigs", fmd, fa)
    lines = open(fa).read().split("\n")
    got = [l for l in lines if l and not l.startswith(">")]
    assert got == ["".join("$ACGTN"[int(x)] for x in c) for c in contigs]


def test_collection_without_reverse_strands_is_refused(exe, tmp_path):
    rng = np.random.default_rng(4)
    seqs = [rng.integers(1, 5, 500).astype(np.uint8) for _ in range(3)]
    T = np.concatenate([np.concatenate([s, [0]]) for s in seqs]).astype(np.uint8)   # forward strands only (ropebwt3 build -R)
    SA = oracle.suffix_array(T)
    bwt = oracle.bwt_from_sa(T, SA)
    raw, fmd, fa = str(tmp_path / "b.bin"), str(tmp_path / "b.fmd"), str(tmp_path / "c.fa")
    bwt.tofile(raw)
    cxx(exe, "encode", raw, fmd)
    r = subprocess.run([exe, "_fmd", "contigs", fmd, fa], capture_output=True, text=True)
    assert r.returncode == 1 and "reverse complement" in r.stderr


def test_search_on_an_fmd_index_reaches_the_gpu_build(exe, tmp_path):
    """`search --index x.fmd` decodes and inverts on the host, then needs the GPU like any index."""
    contigs = synth.make_reference(5_000, seed=5, contigs=2, n_repeats=1, n_nruns=0)
    T, SA, bwt = oracle_index(contigs)
    raw, fmd = str(tmp_path / "b.bin"), str(tmp_path / "b.fmd")
    bwt.tofile(raw)
    cxx(exe, "encode", raw, fmd)
    fq = tmp_path / "r.fq"
    fq.write_text("@r1\n" + "ACGT" * 40 + "\n+\n" + "I" * 160 + "\n")
    r = subprocess.run([exe, "search", "--index", fmd, "--fastx", str(fq)], capture_output=True, text=True)
    assert "ropebwt3 FMD index: 4 sequences (2 after dropping reverse strands), 5000 bp" in r.stderr
    if r.returncode != 0:                                                     # no GPU here: the build must fail loudly
        assert "svb_index_build" in r.stderr and r.stdout == ""


def test_stream_longer_than_one_chunk(exe, tmp_path):
    """More than 2^23 words: the last word of a chunk carries no pairs (rld_get_stail) and the block
    before it is one word shorter.  C++ writer -> C++ reader only (the Python model is too slow here)."""
    rng = np.random.default_rng(11)
    sym = rng.integers(0, 6, 150_000_000, dtype=np.uint8)        # runs of ~1.2 symbols, ~4.6 bits each
    raw, fmd, back = str(tmp_path / "b.bin"), str(tmp_path / "b.fmd"), str(tmp_path / "back.bin")
    sym.tofile(raw)
    cxx(exe, "encode", raw, fmd)
    assert os.path.getsize(fmd) > (1 << 23) * 8 + 4096
    with open(fmd, "rb") as f:
        f.seek(32 + 48 + ((1 << 23) - 1) * 8)                        # header: magic, u32, reserved u64, n_bytes, n_frames, mcnt[6]
        assert f.read(8) == b"\0" * 8                               # the unused word at the end of chunk 0
    cxx(exe, "decode", fmd, back)
    assert np.array_equal(np.fromfile(back, np.uint8), sym)
