"""CPU: the device side of the 2-bit read transport (svdss_b200/csrc/unpack2.cuh) compiled for the host with
the warp emulator, fed by the host side (svb_pack2_host): pack -> unpack must give back the nt6 bytes of the
reads, for the whole batch at once and range by range the way the streamed pipeline decodes chunks."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from svdss_b200 import build, capi

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul():
    build.build_lib()
    src = os.path.join(HERE, "emul", "unpack2_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libunpack2_emul.so")
    deps = [src, os.path.join(HERE, "emul", "warp_emul.hpp"), os.path.join(ROOT, "svdss_b200", "csrc", "unpack2.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = C.CDLL(out)
    lib.emul_unpack2.restype = C.c_int
    return lib


@pytest.mark.parametrize("chunk", [0, 4096])
def test_pack2_then_unpack2_is_the_identity(emul, chunk):
    rng = np.random.default_rng(12)
    lens = [1, 2, 3, 4, 5, 15, 16, 17, 31, 33, 64, 100, 257, 1023, 1024, 1025, 5000] + [int(x) for x in rng.integers(1, 2500, 25)]
    reads = [rng.integers(1, 5, size=l).astype(np.uint8) for l in lens]
    seq4, s4o, lq = capi.pack_bam4(reads)
    pk, pko, exc = capi.pack2_host(seq4, s4o, lq)
    assert not exc.any()
    offs = np.zeros(len(reads) + 1, np.int64)
    offs[1:] = np.cumsum(lens)
    total = int(offs[-1])
    pk_pad = np.concatenate([pk, np.zeros(16, np.uint8)])                       # the decoder reads whole 8-byte words
    out = np.full(total + 64, 0xEE, np.uint8)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    # 8-byte aligned base like a cudaMalloc'ed buffer
    assert pk_pad.ctypes.data % 8 == 0 and out.ctypes.data % 16 == 0
    rc = emul.emul_unpack2(p(pk_pad), p(pko), p(offs), C.c_int64(len(reads)), C.c_int64(total), C.c_int64(chunk), p(out))
    assert rc == 0
    assert np.array_equal(out[:total], np.concatenate(reads))
    assert (out[total:] == 0xEE).all()                                          # nothing written past the batch
