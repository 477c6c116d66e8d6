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
