"""GPU: svb_bgzf_inflate_device (inflate_kernel.cuh: one warp per BGZF member, and the first kernel with one
thread per member behind SVB_INFLATE_KERNEL=thread) against zlib -- every block
type, a BAM written by the test writer, and corrupt members, which must be reported (SVB_EIO, per-member status)
while the healthy members of the same call still come out right.  CPU twin through the warp emulator:
tests/test_inflate_emul.py.  New at the end of round 1 and never run on a GPU: last in the suite."""
import zlib

import numpy as np
import pytest

from svdss_b200 import capi
from test_inflate_emul import payloads, raw_deflate

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["warp", "thread"])
def kernel(request, monkeypatch):
    """the library reads SVB_INFLATE_KERNEL at every call"""
    monkeypatch.setenv("SVB_INFLATE_KERNEL", request.param)
    return request.param


def test_random_damage_never_leaves_its_range():
    """three flipped bits per member: any status, but a member reported good equals zlib and no member touches its neighbours"""
    rng = np.random.default_rng(19)
    d = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=30000)) + bytes(20000)
    good = raw_deflate(d, 6)
    comps = [good]
    for _ in range(200):
        b = bytearray(good)
        for _ in range(3):
            b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
        comps.append(bytes(b))
    comps.append(good)
    r = capi.bgzf_inflate_device(comps, [len(d)] * len(comps), check_status=False)
    assert r.status[0] == 0 and r.status[-1] == 0
    assert r.out[:len(d)].tobytes() == d and r.out[int(r.out_offs[-2]):].tobytes() == d
    n_good = 0
    for k in range(1, len(comps) - 1):
        if r.status[k] == 0:
            assert r.out[int(r.out_offs[k]):int(r.out_offs[k + 1])].tobytes() == zlib.decompress(comps[k], -15)
            n_good += 1
    print("damaged members that still inflate: %d of 200" % n_good)


def test_every_block_type_equals_zlib():
    cases = payloads()
    r = capi.bgzf_inflate_device([c for _, c in cases], [len(d) for d, _ in cases])
    assert (r.status == 0).all()
    for k, (d, _) in enumerate(cases):
        assert r.out[int(r.out_offs[k]):int(r.out_offs[k + 1])].tobytes() == d, k
    assert capi.bgzf_inflate_device([], []).rc == 0


def test_a_bam_file_window(tmp_path):
    from bam_writer import write_bam
    rng = np.random.default_rng(8)
    recs = [dict(qname="r%d" % i, flag=0, tid=0, pos=100 * i, mapq=60, seq="".join("ACGT"[int(x)] for x in rng.integers(0, 4, 9000)),
                 cigar=[(9000, "M")], tags={"XF": ("C", 0)}) for i in range(400)]
    path = str(tmp_path / "t.bam")
    write_bam(path, [("chr1", 1_000_000)], recs)
    comps, sizes = capi.bgzf_members(open(path, "rb").read())
    assert len(comps) > 64 and sizes[-1] == 0
    r = capi.bgzf_inflate_device(comps, sizes)
    want = b"".join(zlib.decompress(c, -15) for c in comps)
    assert r.out.tobytes() == want and want[:4] == b"BAM\x01"
    print("inflated %d members, %.1f MB in %.3f ms on the device" % (len(comps), len(want) / 1e6, r.kernel_ms))


def test_corrupt_members_are_reported():
    rng = np.random.default_rng(9)
    d = bytes(rng.choice(np.frombuffer(b"ACGT", np.uint8), size=20000))
    good = raw_deflate(d, 6)
    comps, sizes = [good, good[:len(good) // 2], good, b"\x07" + good, good], [len(d), len(d), len(d) - 5, len(d), len(d)]
    r = capi.bgzf_inflate_device(comps, sizes, check_status=False)
    assert r.rc == -74                                                    # SVB_EIO
    assert r.status[0] == 0 and r.status[4] == 0 and r.status[1] != 0 and r.status[2] != 0 and r.status[3] != 0
    assert r.out[:len(d)].tobytes() == d and r.out[int(r.out_offs[4]):].tobytes() == d
    with pytest.raises(capi.SvbError):
        capi.bgzf_inflate_device(comps, sizes)
