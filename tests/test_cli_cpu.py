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
