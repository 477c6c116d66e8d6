"""GPU parity tests proper: the CUDA path (through the C ABI) against the oracle, bit-exact."""
import os

import numpy as np
import pytest

import oracle
from common import load_golden, random_case, oracle_index, fm_results
from svdss_b200 import capi, synth

pytestmark = pytest.mark.gpu
BLOCKS = [64, 128]


def _gpu_sfs(idx, reads, assemble):
    cat, offs = oracle.concat(reads)
    res = idx.sfs_batch(cat, offs, assemble=assemble)
    assert res.n_reads == len(reads)
    return [res.per_read(i) for i in range(len(reads))], res


@pytest.mark.parametrize("seed", range(3))
def test_suffix_array_small(seed):
    rng = np.random.default_rng(seed)
    contigs, _ = random_case(rng, with_n=bool(seed % 2), n_reads=0, max_contig=2000)
    if seed == 2:  # long exact repeats and an N run longer than the 21-symbol first-pass key
        c = contigs[0]
        contigs.append(np.concatenate([c, c[: len(c) // 2], np.full(90, 5, np.uint8), c[::-1]]))
    T = oracle.build_text(contigs)
    assert capi.suffix_array(T).tolist() == oracle.suffix_array(T).tolist()


@pytest.mark.parametrize("bb", BLOCKS)
def test_index_build_bwt_and_acc(bb):
    contigs = synth.make_reference(300_000, seed=11, contigs=3)
    T, SA, bwt = oracle_index(contigs)
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs, block_bytes=bb)
    assert idx.n == len(T) and idx.block_bytes == bb
    assert idx.acc == oracle.FMIndex(bwt).acc.tolist()
    assert np.array_equal(idx.bwt(), bwt)


@pytest.mark.parametrize("bb", BLOCKS)
def test_rank2a_parity(bb):
    rng = np.random.default_rng(5)
    contigs = synth.make_reference(200_000, seed=12, contigs=2)
    T, SA, bwt = oracle_index(contigs)
    fm = oracle.FMIndex(bwt)
    idx = capi.Index.from_bwt(bwt, block_bytes=bb)
    n = len(bwt)
    k = rng.integers(0, n + 1, size=4000)
    l = np.minimum(n, k + rng.integers(0, 5000, size=4000) * (rng.random(4000) < 0.7))
    edge = np.array([0, 1, 127, 128, 129, 255, 256, 257, n - 1, n], np.int64)
    k = np.concatenate([k, edge, np.zeros(1, np.int64)])
    l = np.concatenate([l, edge, np.array([n], np.int64)])
    ok, ol = idx.rank2a(k, l)
    for i in range(len(k)):
        a, b = fm.rank2a(int(k[i]), int(l[i]))
        assert ok[i].tolist() == a.tolist() and ol[i].tolist() == b.tolist(), i


@pytest.mark.parametrize("bb", BLOCKS)
def test_golden_fixtures(bb):
    for contigs, reads, raw, asm in load_golden():
        cat, offs = oracle.concat(contigs)
        idx = capi.Index.build(cat, offs, block_bytes=bb)
        got_raw, _ = _gpu_sfs(idx, reads, assemble=False)
        got_asm, _ = _gpu_sfs(idx, reads, assemble=True)
        assert got_raw == raw
        assert got_asm == asm


@pytest.mark.parametrize("bb", BLOCKS)
@pytest.mark.parametrize("seed", range(3))
def test_random_small_parity(bb, seed):
    rng = np.random.default_rng(50 + seed)
    contigs, reads = random_case(rng, with_n=bool(seed % 2), n_reads=300, max_contig=400)
    reads += [np.zeros(0, np.uint8), np.array([5], np.uint8), np.array([1], np.uint8),
              np.full(40, 5, np.uint8), contigs[0].copy(), synth.revcomp6(contigs[-1])]
    T, SA, bwt = oracle_index(contigs)
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs, block_bytes=bb)
    exp = [oracle.sfs_spec(T, SA, r) if len(r) else [] for r in reads]
    got_raw, res = _gpu_sfs(idx, reads, assemble=False)
    assert got_raw == exp
    got_asm, _ = _gpu_sfs(idx, reads, assemble=True)
    assert got_asm == [oracle.assemble(e) for e in exp]
    # whole contigs (either strand) occur in the index: no SFS
    assert got_raw[-1] == [] and got_raw[-2] == []
    fm_exp, ext = fm_results(oracle.FMIndex(bwt), reads)
    assert res.n_ext == ext  # same number of rb3_fmd_extend calls as the literal CPU port


def test_empty_batch_and_bad_args():
    contigs = [np.array([1, 2, 3, 4, 1, 1, 2], np.uint8)]
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs)
    res = idx.sfs_batch(np.zeros(0, np.uint8), np.zeros(1, np.int64))
    assert res.n_reads == 0 and res.n_sfs == 0
    with pytest.raises(capi.SvbError):
        idx.sfs_batch(np.array([1, 2], np.uint8), np.array([0, 2], np.int64), overlap=3)
    with pytest.raises(capi.SvbError):
        capi.Index.build(cat, offs, block_bytes=96)
    # overlap == 0 is the "relaxed" branch (ping_pong.cpp:44-45): begin -= 1 after every SFS
    r = np.array([1, 2, 3, 3, 3, 3, 4, 1, 1, 2], np.uint8)
    res = idx.sfs_batch(r, np.array([0, len(r)], np.int64), overlap=0, assemble=False)
    assert res.n_sfs >= 1


@pytest.mark.parametrize("bb", BLOCKS)
def test_config1_full_parity(bb):
    """1 Mb reference (planted repeats, N runs), 1000 smoothed-shaped 15 kb
    reads + 200 raw-HiFi-shaped reads. SFS sets bit-identical to the oracle."""
    contigs = synth.make_reference(1_000_000, seed=1)
    reads = synth.make_reads(contigs, 1000, seed=2) + synth.make_reads(contigs, 200, seed=3, raw_hifi=True)
    T, SA, bwt = oracle_index(contigs)
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs, block_bytes=bb)
    assert np.array_equal(idx.bwt(), bwt)
    exp, ext = fm_results(oracle.FMIndex(bwt), reads)
    for i in range(0, len(reads), 37):  # SA-narrowing statement on a sample, FM port on everything
        assert oracle.sfs_spec(T, SA, reads[i]) == exp[i]
    got_raw, res = _gpu_sfs(idx, reads, assemble=False)
    assert got_raw == exp
    assert res.n_ext == ext
    rank_ext = res.n_ext - res.n_text_ext       # extensions answered from index blocks or the K-mer table
    assert 0 < res.n_blocks_touched      # (the tail kernel's sprints fetch blocks for links that are not used)
    if bb == 128:   # the default kernel; the 64-byte-block lane-group kernels are rank walk only
        assert res.n_text_ext > res.n_ext // 2      # smoothed reads: most of the walk is a located match
    got_asm, _ = _gpu_sfs(idx, reads, assemble=True)
    assert got_asm == [oracle.assemble(e) for e in exp]
    # resident path returns the same thing
    dr = capi.DeviceReads(*oracle.concat(reads))
    res2 = idx.sfs_resident(dr, assemble=False)
    assert [res2.per_read(i) for i in range(len(reads))] == exp


def test_index_save_load_roundtrip(tmp_path):
    contigs = synth.make_reference(100_000, seed=4, contigs=2)
    reads = synth.make_reads(contigs, 50, seed=5, mean_len=3000, sd_len=500, min_len=500, max_len=6000)
    cat, offs = oracle.concat(contigs)
    for bb in BLOCKS:
        idx = capi.Index.build(cat, offs, block_bytes=bb)
        a, _ = _gpu_sfs(idx, reads, assemble=True)
        p = os.path.join(tmp_path, "idx%d.svb" % bb)
        idx.save(p)
        idx2 = capi.Index.load(p)
        assert idx2.n == idx.n and idx2.acc == idx.acc and idx2.block_bytes == bb
        b, rb = _gpu_sfs(idx2, reads, assemble=True)
        assert a == b
        assert bb == 64 or rb.n_text_ext > 0   # text, SA samples and contig starts came back from the file
    with pytest.raises(capi.SvbError):
        capi.Index.load(os.path.join(tmp_path, "missing.svb"))


def test_medium_scale_properties():
    """Size-independent properties on a 16 Mb multi-contig reference (32 M-symbol BWT): the BWT is
    a permutation of the text; every emitted SFS is minimal-absent (checked by the CPU port on the
    GPU-built BWT and by direct substring search on a sample); assemble output is sorted,
    non-overlapping and idempotent."""
    contigs = synth.make_reference(16_000_000, seed=21, contigs=5, n_repeats=40, n_nruns=6, nrun_len=3000)
    reads = synth.make_reads(contigs, 400, seed=22)
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs, block_bytes=128)
    T = oracle.build_text(contigs)
    bwt = idx.bwt()
    assert np.array_equal(np.bincount(bwt, minlength=6), np.bincount(T, minlength=6))
    fm = oracle.FMIndex(bwt)
    exp, ext = fm_results(fm, reads)
    got, res = _gpu_sfs(idx, reads, assemble=False)
    assert got == exp and res.n_ext == ext
    strands = [c.tobytes() for c in contigs] + [synth.revcomp6(c).tobytes() for c in contigs]
    occurs = lambda w: any(w.tobytes() in s for s in strands)
    checked = 0
    for r, g in zip(reads[:40], got[:40]):
        for qs, ln in g[:3]:
            w = r[qs:qs + ln]
            assert not occurs(w)
            assert ln == 1 or (occurs(w[1:]) and occurs(w[:-1]))
            checked += 1
    assert checked > 20
    asm, _ = _gpu_sfs(idx, reads, assemble=True)
    for a in asm:
        assert a == sorted(a)
        for x, y in zip(a, a[1:]):
            assert x[0] + x[1] <= y[0]
        assert oracle.assemble(a) == a


@pytest.mark.parametrize("cfg,bb", [("mop", 128), ("cpa", 128), ("tma", 128), ("8x1", 128), ("4x2", 128), ("2x4", 128), ("1x8", 128),
                                    ("4x1", 64), ("2x2", 64), ("1x4", 64)])
def test_all_kernel_configs_agree(cfg, bb, monkeypatch):
    """every lane-group / staging configuration of the search kernel returns the oracle's SFS sets"""
    monkeypatch.setenv("SVB_SEARCH_CFG", cfg)
    contigs = synth.make_reference(400_000, seed=31, contigs=3)
    reads = synth.make_reads(contigs, 300, seed=32, mean_len=6000, sd_len=1500, min_len=300, max_len=12000)
    reads += synth.make_reads(contigs, 60, seed=33, mean_len=3000, sd_len=500, min_len=300, max_len=5000, raw_hifi=True)
    T, SA, bwt = oracle_index(contigs)
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs, block_bytes=bb)
    exp, ext = fm_results(oracle.FMIndex(bwt), reads)
    got, res = _gpu_sfs(idx, reads, assemble=False)
    assert got == exp and res.n_ext == ext
    got_asm, _ = _gpu_sfs(idx, reads, assemble=True)
    assert got_asm == [oracle.assemble(e) for e in exp]


def test_streamed_host_batch_matches_resident(monkeypatch):
    """svb_sfs_batch streams the read bytes in chunks behind the running kernel; force that path
    (tiny chunks) and compare with the oracle and with the resident path"""
    contigs = synth.make_reference(600_000, seed=41, contigs=2)
    reads = synth.make_reads(contigs, 500, seed=42, mean_len=8000, sd_len=2500, min_len=200, max_len=20000)
    reads.insert(3, np.zeros(0, np.uint8))
    T, SA, bwt = oracle_index(contigs)
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs, block_bytes=128)
    exp, ext = fm_results(oracle.FMIndex(bwt), reads)
    monkeypatch.setenv("SVB_STREAM_MIN_BYTES", "1")
    monkeypatch.setenv("SVB_STREAM_CHUNK_BYTES", "65536")
    for assemble in (False, True):
        got, res = _gpu_sfs(idx, reads, assemble=assemble)
        assert got == (exp if not assemble else [oracle.assemble(e) for e in exp])
        assert res.n_ext == ext
    monkeypatch.setenv("SVB_NO_STREAM", "1")
    got2, _ = _gpu_sfs(idx, reads, assemble=False)
    assert got2 == exp


def test_located_match_mode_equals_rank_walk(monkeypatch):
    """The located-match mode (unique interval on a sampled SA row -> byte compare against the text,
    forward phase mirrored onto the other strand) must give the SFS sets AND the extension count of
    the pure rank walk: many short contigs (mirror / sentinel / text-start edges), N runs, reads
    ending inside matches, both assemble modes; an index built from a bare BWT has no text."""
    rng = np.random.default_rng(77)
    contigs = [rng.integers(1, 5, size=int(rng.integers(40, 4000)), dtype=np.uint8) for _ in range(60)]
    contigs[3][100:130] = 5
    contigs.append(contigs[5][200:900].copy())                 # an exact repeat: intervals of size 2
    contigs.append(synth.revcomp6(contigs[7][50:1500]))
    reads = []
    for _ in range(600):
        c = contigs[int(rng.integers(len(contigs)))]
        a = int(rng.integers(0, max(1, len(c) - 30)))
        b = int(rng.integers(a + 1, len(c) + 1))
        r = c[a:b].copy()
        if rng.random() < 0.5:
            r = synth.revcomp6(r)
        k = int(rng.integers(0, 4))
        if k == 1 and len(r) > 10:
            p = int(rng.integers(0, len(r))); r[p] = (r[p] % 4) + 1
        elif k == 2:
            p = int(rng.integers(0, len(r) + 1)); r = np.concatenate([r[:p], rng.integers(1, 5, size=int(rng.integers(1, 60)), dtype=np.uint8), r[p:]])
        elif k == 3 and len(r) > 40:
            p = int(rng.integers(0, len(r) - 20)); r = np.concatenate([r[:p], r[p + int(rng.integers(1, 20)):]])
        reads.append(np.ascontiguousarray(r, np.uint8))
    T, SA, bwt = oracle_index(contigs)
    cat, offs = oracle.concat(contigs)
    exp, ext = fm_results(oracle.FMIndex(bwt), reads)
    for spec_i in range(0, len(reads), 41):
        assert oracle.sfs_spec(T, SA, reads[spec_i]) == exp[spec_i]
    idx = capi.Index.build(cat, offs, block_bytes=128)
    got, res = _gpu_sfs(idx, reads, assemble=False)
    assert got == exp and res.n_ext == ext and res.n_text_ext > 0
    got_asm, _ = _gpu_sfs(idx, reads, assemble=True)
    assert got_asm == [oracle.assemble(e) for e in exp]
    monkeypatch.setenv("SVB_SEARCH_TEXT", "0")
    got0, res0 = _gpu_sfs(idx, reads, assemble=False)
    assert got0 == exp and res0.n_ext == ext and res0.n_text_ext == 0
    # K-mer jump table off as well: the plain rank walk touches the most blocks
    monkeypatch.setenv("SVB_SEARCH_JUMP", "0")
    got00, res00 = _gpu_sfs(idx, reads, assemble=False)
    assert got00 == exp and res00.n_ext == ext and res00.n_blocks_touched > res00.n_ext // 2
    monkeypatch.delenv("SVB_SEARCH_TEXT")
    got01, res01 = _gpu_sfs(idx, reads, assemble=False)      # located matches without the jump table
    assert got01 == exp and res01.n_ext == ext and res01.n_text_ext > 0
    monkeypatch.delenv("SVB_SEARCH_JUMP")
    idx_b = capi.Index.from_bwt(bwt, block_bytes=128)
    got_b, res_b = _gpu_sfs(idx_b, reads, assemble=False)
    assert got_b == exp and res_b.n_text_ext == 0


def test_bam4_packed_input_matches_byte_input(monkeypatch):
    """svb_sfs_batch_bam4: reads handed over as BAM stores them (4-bit nt16, byte-aligned per read) are
    decoded on the GPU (ping_pong.cpp:90-94) and must give the oracle's SFS sets: odd and even lengths,
    empty reads, N and IUPAC codes (all -> code 5), plain upload and the streamed path with tiny chunks
    (chunk boundaries inside reads and inside bytes)."""
    contigs = synth.make_reference(500_000, seed=51, contigs=3)
    reads = synth.make_reads(contigs, 400, seed=52, mean_len=7000, sd_len=2500, min_len=150, max_len=18000)
    reads += synth.make_reads(contigs, 40, seed=53, mean_len=3001, sd_len=400, min_len=301, max_len=5001, raw_hifi=True)
    reads.insert(5, np.zeros(0, np.uint8))
    reads.insert(77, reads[10][:1].copy())
    reads[20] = reads[20].copy(); reads[20][100:104] = 5
    T, SA, bwt = oracle_index(contigs)
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs, block_bytes=128)
    exp, ext = fm_results(oracle.FMIndex(bwt), reads)
    seq4, s4o, lq = capi.pack_bam4(reads)
    seq4 = seq4.copy()
    # an IUPAC code (R = 5) where read 20 has its N run decodes to N as well
    b = int(s4o[20]) + 50
    seq4[b] = (5 << 4) | 5
    monkeypatch.setenv("SVB_NO_STREAM", "1")
    res = idx.sfs_batch_bam4(seq4, s4o, lq, assemble=False)
    assert [res.per_read(i) for i in range(len(reads))] == exp and res.n_ext == ext
    assert res.h2d_bytes < sum(len(r) for r in reads) * 0.6 + 16 * len(reads) + 64
    monkeypatch.delenv("SVB_NO_STREAM")
    monkeypatch.setenv("SVB_STREAM_MIN_BYTES", "1")
    for chunk in ("65536", "100032"):
        monkeypatch.setenv("SVB_STREAM_CHUNK_BYTES", chunk)
        for assemble in (False, True):
            res = idx.sfs_batch_bam4(seq4, s4o, lq, assemble=assemble)
            got = [res.per_read(i) for i in range(len(reads))]
            assert got == (exp if not assemble else [oracle.assemble(e) for e in exp])
            assert res.n_ext == ext
    empty = idx.sfs_batch_bam4(np.zeros(0, np.uint8), np.zeros(1, np.int64), np.zeros(0, np.int32))
    assert empty.n_reads == 0 and empty.n_sfs == 0
    with pytest.raises(capi.SvbError):
        idx.sfs_batch_bam4(seq4, s4o, lq + 4)          # lengths that do not fit their packed bytes


def test_runs_of_n_longer_than_any_in_the_text_have_a_closed_form():
    """ping_pong.cpp:12-47 on a run of N longer than the longest run of N in the text: every base of the run is its own
    restart of 2 L extensions.  The kernel answers those restarts without the index; SFSs and extension counts must be
    the walk's, for runs at the read's start, end, middle, for runs of exactly L and L + 1, and for a read that is N only
    (8 kb: 3.2 M serial extensions on one lane in round 1) -- inside a time that shows no lane walked them."""
    import time
    contigs = synth.make_reference(300_000, seed=8, contigs=2, n_nruns=4, nrun_len=200)
    L = 200
    assert max(len(s) for c in contigs for s in "".join("N" if x == 5 else "a" for x in c).split("a")) == L
    rng = np.random.default_rng(9)
    c0 = contigs[0]
    base = lambda a, n: c0[a:a + n].copy()
    N = lambda n: np.full(n, 5, np.uint8)
    reads = [np.concatenate([base(1000, 3000), N(L + 1), base(5000, 2000)]),          # L + 1 in the middle: one closed-form restart
             np.concatenate([base(1000, 3000), N(L), base(5000, 2000)]),              # exactly L: the walk as it is
             np.concatenate([N(700), base(9000, 2500)]),                              # run at the read's start
             np.concatenate([base(12000, 2500), N(650)]),                             # run at the read's end
             np.concatenate([base(20000, 900), N(450), base(30000, 40), N(300), base(40000, 900)]),
             N(8000), N(L + 1), N(L), N(3),
             np.concatenate([base(50000, 1200), rng.integers(1, 5, 300).astype(np.uint8), N(1000), rng.integers(1, 5, 200).astype(np.uint8)])]
    T, SA, bwt = oracle_index(contigs)
    exp, ext = fm_results(oracle.FMIndex(bwt), reads)
    assert len(exp[5]) > 7000 and ext > 3_000_000
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs)
    _gpu_sfs(idx, reads[:2], assemble=False)                                          # warm-up
    t = time.perf_counter()
    got, res = _gpu_sfs(idx, reads, assemble=False)
    dt = time.perf_counter() - t
    assert got == exp and res.n_ext == ext
    # what is left to walk: the <= L restarts next to the end of a run (the run's neighbour decides, no closed form): ~L^2 extensions
    assert res.kernel_ms < 100.0, res.kernel_ms
    got1, res1 = _gpu_sfs(idx, [reads[5]], assemble=False)                            # the read that is N only, alone
    assert got1 == [exp[5]] and res1.kernel_ms < 10.0, res1.kernel_ms                 # "an 8 kb all-N read finishes in under 10 ms"
    got_asm, _ = _gpu_sfs(idx, reads, assemble=True)
    assert got_asm == [oracle.assemble(e) for e in exp]
    os.environ["SVB_SEARCH_NO_NRUN"] = "1"                                            # the plain walk agrees (and shows what the closed form saves)
    try:
        got2, res2 = _gpu_sfs(idx, reads, assemble=False)
    finally:
        del os.environ["SVB_SEARCH_NO_NRUN"]
    assert got2 == exp and res2.n_ext == ext
    print("N-run closed form: kernel %.2f ms (call %.1f ms; the all-N read alone %.2f ms) vs %.1f ms walking" % (res.kernel_ms, dt * 1e3, res1.kernel_ms, res2.kernel_ms))


def test_index_builder_on_a_repeat_rich_reference():
    """Segmental duplications (exact and 1 %-diverged copies, either strand) and tandem arrays: most suffixes stay
    unresolved after the first pass of the suffix sort and go through prefix doubling.  SA and BWT must equal the
    oracle's, and the search on that index must agree with the oracle port."""
    rng = np.random.default_rng(31)
    n = 400_000
    ref = rng.integers(1, 5, n).astype(np.uint8)
    for _ in range(12):                                          # segmental duplications, half reverse-complemented
        L = int(rng.integers(5_000, 30_000))
        src, dst = int(rng.integers(0, n - L)), int(rng.integers(0, n - L))
        seg = ref[src:src + L].copy()
        if rng.random() < 0.5:
            seg = synth.revcomp6(seg)
        if rng.random() < 0.7:
            m = rng.random(L) < 0.01
            seg[m] = (seg[m] % 4) + 1
        ref[dst:dst + L] = seg
    for _ in range(10):                                          # tandem arrays
        unit = rng.integers(1, 5, int(rng.integers(1, 40))).astype(np.uint8)
        L = int(rng.integers(500, 6_000))
        dst = int(rng.integers(0, n - L))
        ref[dst:dst + L] = np.tile(unit, L // len(unit) + 1)[:L]
    contigs = [ref[:150_000].copy(), ref[150_000:].copy()]
    T, SA, bwt = oracle_index(contigs)
    cat, offs = oracle.concat(contigs)
    idx = capi.Index.build(cat, offs)
    assert np.array_equal(idx.bwt(), bwt)
    assert np.array_equal(capi.suffix_array(T), SA)
    reads = synth.make_reads(contigs, 150, seed=32, mean_len=4000, sd_len=1500, min_len=300, max_len=9000)
    exp, ext = fm_results(oracle.FMIndex(bwt), reads)
    got, res = _gpu_sfs(idx, reads, assemble=False)
    assert got == exp and res.n_ext == ext
