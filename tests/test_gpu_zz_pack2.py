"""GPU: the 2-bit read transport end to end on its two halves -- svb_pack2_host on the CPU, k_unpack2 on the
device (svb_unpack2_device) -- gives back the reads, at every length mod 16 and for a batch of 15 kb reads."""
import numpy as np
import pytest

from svdss_b200 import capi

pytestmark = pytest.mark.gpu


def test_pack2_host_then_unpack2_device():
    rng = np.random.default_rng(13)
    lens = list(range(1, 70)) + [255, 256, 257, 1023, 1024, 1025] + [int(x) for x in rng.integers(5000, 25000, 60)]
    reads = [rng.integers(1, 5, size=l).astype(np.uint8) for l in lens]
    seq4, s4o, lq = capi.pack_bam4(reads)
    pk, pko, exc = capi.pack2_host(seq4, s4o, lq)
    assert not exc.any()
    offs = np.zeros(len(reads) + 1, np.int64)
    offs[1:] = np.cumsum(lens)
    got = capi.unpack2_device(pk, pko, offs)
    assert np.array_equal(got, np.concatenate(reads))
