"""CPU: the device inflate (svdss_b200/csrc/inflate_kernel.cuh: one thread per BGZF member, RFC 1951 restated)
compiled for the host with the warp emulator and compared with zlib: stored, fixed-Huffman and dynamic blocks,
multi-block streams, back-references at the maximum distance, the empty EOF member, a real BAM written by the
test writer, and corrupt / truncated members, which must be reported and never write outside their range."""
import ctypes as C
import os
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul_lib():
    src = os.path.join(HERE, "emul", "inflate_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libinflate_emul.so")
    deps = [src, os.path.join(HERE, "emul", "warp_emul.hpp"), os.path.join(ROOT, "svdss_b200", "csrc", "inflate_kernel.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    lib = C.CDLL(out)
    lib.emul_bgzf_inflate.restype = C.c_int
    lib.emul_bgzf_inflate_mpw.restype = C.c_int
    lib.emul_bgzf_inflate_warp.restype = C.c_int
    return lib


class Emul:
    def __init__(self, lib, kernel):
        self.lib, self.kernel = lib, kernel


@pytest.fixture(params=["thread", "warp"])
def emul(emul_lib, request):
    """both kernels: one thread per member (k_bgzf_inflate) and one warp per member (k_bgzf_inflate_warp)"""
    return Emul(emul_lib, request.param)


def raw_deflate(data, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, flush_every=0):
    c = zlib.compressobj(level, zlib.DEFLATED, -15, 8, strategy)
    if not flush_every:
        return c.compress(data) + c.flush()
    out = b""
    for a in range(0, len(data), flush_every):                     # several blocks in one stream
        out += c.compress(data[a:a + flush_every]) + c.flush(zlib.Z_FULL_FLUSH)
    return out + c.flush()


def payloads():
    rng = np.random.default_rng(7)
    dna = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=60000))
    text = (b"@SQ\tSN:chr1\tLN:248956422\n" * 400)[:0xff00]
    noise = bytes(rng.integers(0, 256, size=5000, dtype=np.uint8))
    far = noise[:300] + bytes(32768 - 300) + noise[:300] + b"tail"      # a match at distance 32768
    cases = [
        b"", b"A", b"ACGT" * 3, dna, text, noise, far, bytes(65280), bytes(range(256)) * 100,
        dna[:1000] + noise[:1000] + dna[:1000],
    ]
    out = []
    for d in cases:
        out.append((d, raw_deflate(d, 6)))
        out.append((d, raw_deflate(d, 1)))
        out.append((d, raw_deflate(d, 9, flush_every=4001)))
    out.append((dna, raw_deflate(dna, 0)))                              # stored blocks (65535-byte pieces)
    out.append((b"hello hello hello", raw_deflate(b"hello hello hello", 9, zlib.Z_FIXED)))   # fixed code with matches
    out.append((dna[:5000], raw_deflate(dna[:5000], 6, zlib.Z_FIXED)))
    out.append((dna[:30000] + text[:20000], raw_deflate(dna[:30000] + text[:20000], 6, flush_every=7000)))   # dynamic + empty stored blocks
    out.append((dna[:9000], raw_deflate(dna[:9000], 6, zlib.Z_HUFFMAN_ONLY)))
    out.append((bytes(40000), raw_deflate(bytes(40000), 6, zlib.Z_RLE)))
    return out


def run(lib, comps, out_lens, pad=64, mpw=32):
    io = np.zeros(len(comps) + 1, np.int64); io[1:] = np.cumsum([len(c) for c in comps])
    oo = np.zeros(len(comps) + 1, np.int64); oo[1:] = np.cumsum(out_lens)
    comp = np.frombuffer(b"".join(comps) + b"\xff" * 16, np.uint8).copy()   # what follows the last member is not zero
    out = np.full(int(oo[-1]) + pad, 0xEE, np.uint8)
    status = np.full(len(comps), -7, np.int32)
    p = lambda a: a.ctypes.data_as(C.c_void_p)
    if lib.kernel == "warp":
        assert lib.lib.emul_bgzf_inflate_warp(p(comp), p(io), p(oo), len(comps), p(out), p(status)) == 0
    else:
        assert lib.lib.emul_bgzf_inflate_mpw(p(comp), p(io), p(oo), len(comps), p(out), p(status), mpw) == 0
    return out, oo, status


def test_every_block_type_equals_zlib(emul):
    cases = payloads()
    assert len(cases) > 32                                               # more than one warp of members
    for d, c in cases:
        assert zlib.decompress(c, -15) == d
    out, oo, status = run(emul, [c for _, c in cases], [len(d) for d, _ in cases])
    assert (status == 0).all(), status
    for k, (d, _) in enumerate(cases):
        assert out[int(oo[k]):int(oo[k + 1])].tobytes() == d, k
    assert (out[int(oo[-1]):] == 0xEE).all()


@pytest.mark.parametrize("mpw", [1, 5])
def test_fewer_members_per_warp(emul, mpw):
    """small windows spread over more warps (svb_bgzf_inflate_device picks 1..32 members per warp)"""
    cases = payloads()[:23]
    out, oo, status = run(emul, [c for _, c in cases], [len(d) for d, _ in cases], mpw=mpw)
    assert (status == 0).all(), status
    for k, (d, _) in enumerate(cases):
        assert out[int(oo[k]):int(oo[k + 1])].tobytes() == d, k
    assert (out[int(oo[-1]):] == 0xEE).all()


def test_members_with_their_gzip_trailer_attached(emul):
    """BgzfSource hands the members over as they lie in the file, CRC32 + ISIZE still behind every deflate stream"""
    cases = payloads()[:12]
    comps = [c + zlib.crc32(d).to_bytes(4, "little") + len(d).to_bytes(4, "little") for d, c in cases]
    out, oo, status = run(emul, comps, [len(d) for d, _ in cases])
    assert (status == 0).all(), status
    for k, (d, _) in enumerate(cases):
        assert out[int(oo[k]):int(oo[k + 1])].tobytes() == d, k


def test_members_of_a_bam_file(emul, tmp_path):
    from bam_writer import write_bam
    rng = np.random.default_rng(8)
    recs = [dict(qname="r%d" % i, flag=0, tid=0, pos=100 * i, mapq=60, seq="".join("ACGT"[int(x)] for x in rng.integers(0, 4, 3000)),
                 cigar=[(3000, "M")], tags={"XF": ("C", 0)}) for i in range(60)]
    path = str(tmp_path / "t.bam")
    write_bam(path, [("chr1", 1_000_000)], recs)
    from svdss_b200 import capi
    comps, lens = capi.bgzf_members(open(path, "rb").read())             # SAM spec 4.1: gzip members with a BC extra subfield
    assert len(comps) >= 3 and lens[-1] == 0                            # the EOF member
    out, oo, status = run(emul, comps, lens)
    assert (status == 0).all()
    want = b"".join(zlib.decompress(c, -15) for c in comps)
    assert out[:int(oo[-1])].tobytes() == want and want[:4] == b"BAM\x01"


def test_corrupt_members_are_reported_and_stay_in_their_range(emul):
    rng = np.random.default_rng(9)
    d = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=20000))
    good = raw_deflate(d, 6)
    comps, lens = [good, good[:len(good) // 2], good, good, b"\x07" + good, good], [len(d), len(d), len(d) - 5, len(d) + 5, len(d), len(d)]
    for trial in range(40):                                              # random damage: any status, never a write outside
        b = bytearray(good)
        for _ in range(3):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        comps.append(bytes(b)); lens.append(len(d))
    out, oo, status = run(emul, comps, lens)
    assert status[0] == 0 and status[5] == 0
    assert status[1] != 0 and status[2] != 0 and status[3] != 0 and status[4] != 0    # truncated, too long, too short, reserved block type
    assert out[int(oo[0]):int(oo[1])].tobytes() == d and out[int(oo[5]):int(oo[6])].tobytes() == d
    assert (out[int(oo[-1]):] == 0xEE).all()
    for k in range(6, len(comps)):
        try:
            ok = zlib.decompress(comps[k], -15) == d
        except zlib.error:
            ok = False
        if status[k] == 0:
            assert out[int(oo[k]):int(oo[k + 1])].tobytes() == zlib.decompress(comps[k], -15)
        if ok:
            assert status[k] == 0
