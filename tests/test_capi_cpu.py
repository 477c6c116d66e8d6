"""CPU-only checks of the drop-in boundary: the C-ABI library builds, loads, exports every symbol
include/svdss_b200.h declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from svdss_b200 import build, capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    build.build_lib()
    return capi.lib()


def test_header_symbols_exported(L):
    hdr = open(os.path.join(ROOT, "include", "svdss_b200.h")).read()
    declared = set(re.findall(r"SVB_API\s+[\w\s\*]+?\b(svb_\w+)\s*\(", hdr))
    assert declared, "no declarations parsed"
    assert declared == set(capi.EXPORTS)
    for name in sorted(declared):
        assert hasattr(L, name), name


def test_no_cpu_fallback_without_gpu(L):
    if L.svb_device_count() > 0:
        pytest.skip("a GPU is present")
    bwt = np.array([1, 2, 0, 3], np.uint8)
    with pytest.raises(capi.SvbError) as ei:
        capi.Index.from_bwt(bwt)
    assert ei.value.code == -5 and "no CPU fallback" in str(ei.value)
    with pytest.raises(capi.SvbError):
        capi.suffix_array(np.array([1, 2, 0], np.uint8))
    with pytest.raises(capi.SvbError) as ei:                                 # the device inflate has no host twin behind it either
        capi.bgzf_inflate_device([b"\x03\x00"], [0])
    assert ei.value.code == -5


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "svdss_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", ".hpp")):
                src = open(os.path.join(dp, f), errors="replace").read()
                assert not re.search(r"^\s*(import|from)\s+oracle\b", src, re.M), f
                # comments may cite the oracle's rule list; code may not include, link or load it
                assert "liboracle" not in src, f
                assert not re.search(r"#\s*include[^\n]*oracle", src), f
                assert not re.search(r"(dlopen|CDLL|LoadLibrary)[^\n]*oracle", src), f


def test_pack2_host_round_trip(L):
    """svb_pack2_host (host-only utility): 4 bits -> 2 bits per base, exceptions flagged, every length mod 4,
    reads long enough for the SIMD path, odd input offsets, and a thread count of 1 and of all cores"""
    rng = np.random.default_rng(8)
    lens = [0, 1, 2, 3, 4, 5, 6, 7, 8, 63, 64, 65, 127, 128, 129, 130, 131, 257, 1000, 4099, 15001] + [int(x) for x in rng.integers(1, 3000, 40)]
    reads = [rng.integers(1, 5, size=l).astype(np.uint8) for l in lens]        # nt6 codes A C G T
    for k in (5, 17, 30, 44):                                                  # reads with an N somewhere (first, last, middle base)
        if len(reads[k]):
            reads[k][[0, len(reads[k]) - 1, len(reads[k]) // 2][k % 3]] = 5
    seq4, offs, lq = capi.pack_bam4(reads)
    for threads in (1, 0):
        out, ooffs, exc = capi.pack2_host(seq4, offs, lq, threads=threads)
        for i, r in enumerate(reads):
            has_n = bool((r == 5).any())
            assert bool(exc[i]) == has_n, (i, len(r))
            if has_n:
                continue
            p = out[int(ooffs[i]):int(ooffs[i + 1])]
            assert len(p) == (len(r) + 3) // 4
            codes = np.stack([(p >> 6) & 3, (p >> 4) & 3, (p >> 2) & 3, p & 3], axis=1).reshape(-1)[:len(r)]
            assert np.array_equal(codes + 1, r), (i, len(r))
            if len(r) % 4:                                                      # pad positions are zero
                assert not (int(p[-1]) & ((1 << (2 * (4 - len(r) % 4))) - 1))
