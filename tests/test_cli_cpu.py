"""CPU-only: the C++14 SVDSS shell builds and keeps the reference's argument validation and exit
codes (main.cpp:27-31,56-66,78-81)."""
import subprocess

import pytest

from svdss_b200 import build


@pytest.fixture(scope="module")
def exe():
    build.build_lib()
    return build.build_host()


def test_usage_and_exit_codes(exe):
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 1 and "Usage" in r.stderr and r.stdout == ""
    r = subprocess.run([exe, "search", "--index", "x.idx"], capture_output=True, text=True)   # no --bam/--fastx
    assert r.returncode == 1 and "--fastx" in r.stderr
    r = subprocess.run([exe, "search", "--fastx", "r.fq"], capture_output=True, text=True)    # no --index
    assert r.returncode == 1
    r = subprocess.run([exe, "frobnicate"], capture_output=True, text=True)
    assert r.returncode == 1
    r = subprocess.run([exe, "search", "--help"], capture_output=True, text=True)
    assert r.returncode == 0 and "--noputative" in r.stderr
    r = subprocess.run([exe, "--version"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("SVDSS, ")
    r = subprocess.run([exe, "search", "--bogus"], capture_output=True, text=True)
    assert r.returncode == 1


def test_search_fails_loudly_without_index_or_gpu(exe, tmp_path):
    fq = tmp_path / "r.fq"
    fq.write_text("@r1\nACGT\n+\nIIII\n")
    r = subprocess.run([exe, "search", "--index", str(tmp_path / "none.idx"), "--fastx", str(fq)],
                       capture_output=True, text=True)
    assert r.returncode == 1 and "svb_index_load" in r.stderr and r.stdout == ""


def test_cxxopts_spellings_are_accepted(exe, tmp_path):
    """config.cpp:26-57: every registered option parses, also as `--option=value`; --binary and
    --append are registered but read nowhere."""
    fq = tmp_path / "r.fq"
    fq.write_text("@r1\nACGT\n+\nIIII\n")
    r = subprocess.run([exe, "search", "--index=" + str(tmp_path / "none.idx"), "--fastx=" + str(fq), "--threads=2", "--bsize=10",
                        "--binary", "--append", "x", "--omax=5"], capture_output=True, text=True)
    assert r.returncode == 1 and "svb_index_load" in r.stderr and "unknown option" not in r.stderr
    r = subprocess.run([exe, "call", "--reference=x.fa", "--clipped"], capture_output=True, text=True)   # still needs --bam/--sfs
    assert r.returncode == 1 and "Usage" in r.stderr


def test_index_takes_ropebwt3_build_options(exe, tmp_path):
    """`SVDSS index` is ropebwt3's `build` (main.cpp:34-37; run_svdss:142 `-t N -d FA -o OUT`, README.md:113
    `-t16 -d FA > OUT`): getopt clusters and tuning flags parse, strand/format changes are refused."""
    fa = tmp_path / "r.fa"
    fa.write_text(">a\nACGTACGTAC\n>b\nGGGTTTAACC\n")
    for args in (["-t16", "-d", str(fa)], ["-t", "4", "-d", str(fa), "-o", str(tmp_path / "x.idx")], ["-dt2", "-m", "1G", "-l512", "-2s", str(fa)],
                 ["-L", str(fa), str(fa)]):
        r = subprocess.run([exe, "index"] + args, capture_output=True, text=True)
        assert "unknown option" not in r.stderr and "needs a value" not in r.stderr, r.stderr
        assert "indexing" in r.stderr                      # reached the GPU build (which fails loudly without a GPU)
        if r.returncode != 0:
            assert "svb_index_build" in r.stderr
    r = subprocess.run([exe, "index", "-d", str(fa), str(fa)], capture_output=True, text=True)
    assert "indexing 4 sequences, 40 bp" in r.stderr      # several input files, like ropebwt3 build
    r = subprocess.run([exe, "index", "-L", str(tmp_path / "r.fa")], capture_output=True, text=True)
    assert "indexing 4 sequences" in r.stderr             # -L: every line is a sequence (headers included, as in ropebwt3)
    for bad in (["-R", str(fa)], ["-F", str(fa)], ["-b", str(fa)], ["-q", str(fa)], ["-o"], []):
        r = subprocess.run([exe, "index"] + bad, capture_output=True, text=True)
        assert r.returncode == 1 and "indexing" not in r.stderr, bad
    r = subprocess.run([exe, "index", "-i", str(fa)], capture_output=True, text=True)
    assert r.returncode == 1 and "not a ropebwt3 FMD" in r.stderr


def test_bam_reader_windows_and_truncated_files(exe, tmp_path):
    """host/io.hpp BgzfSource: the file is read a window at a time and the members are found in memory -- the records
    come out the same whatever the window (smaller than a member, cutting members in two, one for all), and a
    file that ends inside a member is an error, not a shorter file."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import numpy as np
    from bam_writer import write_bam
    rng = np.random.default_rng(5)
    recs = [dict(qname="r%d" % i, flag=0, tid=0, pos=100 * i, mapq=60, seq="".join("ACGT"[int(x)] for x in rng.integers(0, 4, 2500)),
                 cigar=[(2500, "M")], tags={"XF": ("C", i % 2)}) for i in range(150)]
    path = str(tmp_path / "t.bam")
    write_bam(path, [("chr1", 1_000_000)], recs)
    seen = set()
    for window in ("7", "3000", "40000", str(64 << 20)):
        r = subprocess.run([exe, "_bamread", path], capture_output=True, text=True, env=dict(os.environ, SVB_BGZF_WINDOW=window))
        assert r.returncode == 0, r.stderr
        j = json.loads(r.stdout.strip().splitlines()[-1])
        assert j["records"] == 150 and j["kept"] == 75 and j["bases"] == 150 * 2500
        seen.add(j["seq_sum"])
    # the payload limit of a window (the device path's pinned buffers have a fixed size): the rest of the read is carried over
    for window, payload in (("50000", "65536"), (str(64 << 20), "70000"), (str(64 << 20), "200000")):
        r = subprocess.run([exe, "_bamread", path], capture_output=True, text=True, env=dict(os.environ, SVB_BGZF_WINDOW=window, SVB_BGZF_PAYLOAD=payload))
        assert r.returncode == 0, r.stderr
        j = json.loads(r.stdout.strip().splitlines()[-1])
        assert j["records"] == 150 and j["kept"] == 75
        seen.add(j["seq_sum"])
    assert len(seen) == 1
    raw = open(path, "rb").read()
    assert len(raw) > 50000
    cut = str(tmp_path / "cut.bam")
    open(cut, "wb").write(raw[:len(raw) * 2 // 3])
    for window in ("3000", str(64 << 20)):
        r = subprocess.run([exe, "_bamread", cut], capture_output=True, text=True, env=dict(os.environ, SVB_BGZF_WINDOW=window))
        assert r.returncode != 0
