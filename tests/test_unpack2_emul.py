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


@pytest.mark.parametrize("chunk_bytes", [777, 4096])
def test_streamed_chunks_deliver_exactly_what_each_range_needs(emul, chunk_bytes):
    """The streamed pipeline in miniature: per chunk of base positions the host packs the covering bytes
    (svb_pack2_chunk), they are copied into the device-side packed array (here: an array that starts as garbage),
    the range is decoded (unpack2_read through the emulator) and the positions without a 2-bit form are patched to N.
    Reads span chunks, start and end anywhere, and some carry N runs."""
    rng = np.random.default_rng(14)
    lens = [int(x) for x in rng.integers(1, 2200, 30)] + [1, 2, 3, 5, 4096, 5000]
    reads = [rng.integers(1, 5, size=l).astype(np.uint8) for l in lens]
    reads[3][:] = 5                                                             # a read of N only
    reads[7][len(reads[7]) // 2:len(reads[7]) // 2 + 40] = 5                    # an N run
    reads[-1][[0, 4999]] = 5                                                    # first and last base
    seq4, s4o, lq = capi.pack_bam4(reads)
    offs = np.zeros(len(reads) + 1, np.int64); offs[1:] = np.cumsum(lens)
    pko = np.zeros(len(reads) + 1, np.int64); pko[1:] = np.cumsum((np.array(lens, np.int64) + 3) // 4)
    total = int(offs[-1])
    dev_pk = np.full(int(pko[-1]) + 16, 0xFF, np.uint8)                         # "device" copy of the packed bytes: nothing delivered yet
    out = np.full(total + 64, 0xEE, np.uint8)
    stage = np.zeros(chunk_bytes // 4 + 8192, np.uint8)
    exc = np.zeros(1 << 16, np.int64)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    L = capi.lib()
    r_lo = 0
    n_chunks = (total + chunk_bytes - 1) // chunk_bytes
    n_exc_total = 0
    for c in range(n_chunks):
        o, nb = c * chunk_bytes, min(chunk_bytes, total - c * chunk_bytes)
        while r_lo + 1 < len(reads) and offs[r_lo + 1] <= o:                    # reads reaching into the chunk, as sfs_batch_streamed finds them
            r_lo += 1
        r_hi = r_lo
        while r_hi + 1 < len(reads) and offs[r_hi + 1] < o + nb:
            r_hi += 1
        pa2, pe2, n_exc = C.c_int64(0), C.c_int64(0), C.c_int64(0)
        rc = L.svb_pack2_chunk(p(seq4), p(s4o), p(offs), p(pko), C.c_int64(r_lo), C.c_int64(r_hi), C.c_int64(o), C.c_int64(nb), p(stage),
                               C.c_int64(len(stage)), C.byref(pa2), C.byref(pe2), p(exc), C.c_int64(len(exc)), C.byref(n_exc), 0)
        assert rc == 0
        dev_pk[pa2.value:pe2.value] = stage[:pe2.value - pa2.value]             # the H2D copy of this chunk
        rc = emul.emul_unpack2_range(p(dev_pk), p(pko), p(offs), C.c_int64(r_lo), C.c_int64(r_hi), C.c_int64(o), C.c_int64(o + nb), p(out))
        assert rc == 0
        out[exc[:n_exc.value]] = 5                                              # the patch the last unpacking CTA applies
        n_exc_total += n_exc.value
        assert np.array_equal(out[o:o + nb], np.concatenate(reads)[o:o + nb]), c   # the chunk is complete before its flag would go up
    assert n_exc_total == int(sum((r == 5).sum() for r in reads))
    assert (out[total:] == 0xEE).all()
