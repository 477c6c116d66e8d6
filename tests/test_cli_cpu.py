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
