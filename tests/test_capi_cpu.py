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
