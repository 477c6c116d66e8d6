"""CPU: the device BAM loader's per-item code (svdss_b200/csrc/bam_core.cuh: record walk by guessed segments + exact
linking, record parse with the aux walk, the load filters of ping_pong.cpp:66-75 and the XF rule of :196-203, nt16 -> nt6)
compiled for the host and held against a plain Python BAM parser -- on windows cut anywhere, with every aux type, and with
payload bytes that look like records (the segment guesses must then be wrong without the walk's answer changing)."""
import ctypes as C
import os
import struct
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


@pytest.fixture(scope="module")
def emul():
    src = os.path.join(HERE, "emul", "bam_emul.cpp")
    out = os.path.join(HERE, "emul", "_build", "libbam_emul.so")
    deps = [src, os.path.join(ROOT, "svdss_b200", "csrc", "bam_core.cuh")]
    if not os.path.exists(out) or any(os.path.getmtime(d) > os.path.getmtime(out) for d in deps):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", out, src])
    return C.CDLL(out)


def record(qname, flag, tid, pos, seq4, l_qseq, qual, cigar=(), aux=b"", mapq=60):
    body = struct.pack("<iiBBHHHiiii", tid, pos, len(qname) + 1, mapq, 4680, len(cigar), flag, l_qseq, -1, -1, 0)
    body += qname + b"\0" + b"".join(struct.pack("<I", c) for c in cigar) + seq4 + qual + aux
    return struct.pack("<i", len(body)) + body


def fake_chain(k=3):
    """k small byte strings that look like records, one after the other"""
    return b"".join(record(b"zz%d" % i, 0, 0, 5 + i, b"\x12\x48", 4, b"\x10" * 4) for i in range(k))


def make_payload(rng, n, adversarial):
    recs, truth = [], []
    for i in range(n):
        l = int(rng.integers(0, 400)) if i % 11 == 0 else int(rng.integers(100, 3000))
        seq4 = rng.integers(0, 256, (l + 1) // 2, dtype=np.uint8).tobytes()
        qual = bytes(rng.integers(0, 94, l, dtype=np.uint8))
        if adversarial and l >= 200 and i % 3 == 0:
            f = fake_chain(int(rng.integers(1, 5)))
            at = int(rng.integers(0, l - len(f))) if l > len(f) else 0
            qual = (qual[:at] + f + qual[at + len(f):])[:l]
        flag = [0, 16, 4, 256, 2048, 1024][int(rng.integers(0, 6))] if i % 4 == 0 else (16 if i % 2 else 0)
        tid = int(rng.integers(-1, 3)) if i % 13 == 0 else int(rng.integers(0, 3))
        aux, xf, hp = b"", None, None
        kinds = rng.permutation(8)[:int(rng.integers(0, 6))]
        for kd in kinds:
            if kd == 0:
                xf = int(rng.integers(0, 3)); aux += b"XF" + [b"C", b"c", b"S", b"i"][i % 4] + {0: struct.pack("<B", xf), 1: struct.pack("<b", xf), 2: struct.pack("<H", xf), 3: struct.pack("<i", xf)}[i % 4]
            elif kd == 1:
                hp = int(rng.integers(0, 3)); aux += b"HPc" + struct.pack("<b", hp)
            elif kd == 2:
                aux += b"RGZ" + b"grp%d" % i + b"\0"
            elif kd == 3:
                aux += b"NMI" + struct.pack("<I", i)
            elif kd == 4:
                aux += b"fqf" + struct.pack("<f", 1.5)
            elif kd == 5:
                aux += b"tyA" + b"x"
            elif kd == 6:
                k = int(rng.integers(0, 5)); aux += b"baBs" + struct.pack("<i", k) + b"\1\0" * k
            elif kd == 7:
                aux += b"hxH" + b"1AE3" + b"\0" + b"ddd" + struct.pack("<d", 2.5)
        qn = b"read/%d/ccs" % i
        cigar = [(l << 4) | 0] if l else []
        recs.append(record(qn, flag, tid, 100 + i, seq4, l, qual, cigar, aux))
        truth.append(dict(qname=qn, flag=flag, tid=tid, l_qseq=l, xf=xf or 0, has_xf=xf is not None, hp=hp or 0, seq4=seq4))
    return recs, truth


def py_walk(win, start, total):
    p, offs = start, []
    while p + 4 <= total:
        bs = struct.unpack_from("<i", win, p)[0]
        assert bs >= 32
        if p + 4 + bs > total:
            break
        offs.append(p + 4)
        p += 4 + bs
    return offs, p


def run_walk(lib, win, start, total, n_seg, n_ref=3):
    a = np.frombuffer(win, np.uint8)
    cap = max(16, (total - start) // 36 + 1)
    off = np.zeros(cap, np.int64)
    res = np.zeros(4, np.int64)
    lib.emul_bam_walk(a.ctypes.data_as(C.c_void_p), C.c_int64(start), C.c_int64(total), C.c_int(n_seg), C.c_int(n_ref), off.ctypes.data_as(C.c_void_p),
                      C.c_int64(cap), res.ctypes.data_as(C.c_void_p))
    return off[:int(res[0])].tolist(), int(res[1]), int(res[2]), int(res[3])


@pytest.mark.parametrize("adversarial", [False, True])
def test_walk_equals_the_serial_chain(emul, adversarial):
    rng = np.random.default_rng(11 + adversarial)
    recs, _ = make_payload(rng, 400, adversarial)
    header = b"BAM\1" + bytes(rng.integers(0, 256, 77, dtype=np.uint8))      # the caller skips it
    win = header + b"".join(recs) + b"\0" * 64
    full = len(win) - 64
    joined_any = 0
    for total in (full, full - 1, full - 40, full // 2, len(header) + len(recs[0]) - 1, len(header) + 3, len(header)):
        want, end = py_walk(win, len(header), total)
        for n_seg in (0, 1, 2, 7, 64, 512):
            got, gend, err, joined = run_walk(emul, win, len(header), total, n_seg)
            assert err == 0 and got == want and gend == end, (total, n_seg)
            joined_any += joined
    assert joined_any > 100                                                 # the guesses were used, not only the fallback
    if adversarial:
        # the planted byte strings do look like records to the guess: some segments start on one and never meet the chain
        a = np.frombuffer(win, np.uint8)
        got, _, _, joined = run_walk(emul, win, len(header), full, 512)
        assert joined < 512


def test_bad_block_size_is_an_error(emul):
    rng = np.random.default_rng(5)
    recs, _ = make_payload(rng, 50, False)
    body = b"".join(recs[:20]) + struct.pack("<i", 7) + b"".join(recs[20:])
    for n_seg in (0, 1, 16):
        got, _, err, _ = run_walk(emul, body + b"\0" * 64, 0, len(body), n_seg)
        assert err == 1 and len(got) == 20


@pytest.mark.parametrize("putative", [1, 0])
def test_parse_filters_and_tags(emul, putative):
    rng = np.random.default_rng(21)
    recs, truth = make_payload(rng, 300, True)
    win = b"".join(recs) + b"\0" * 64
    offs, _ = py_walk(win, 0, len(win) - 64)
    assert len(offs) == 300
    n = len(offs)
    a = np.frombuffer(win, np.uint8)
    ro = np.array(offs, np.int64)
    cols = [np.zeros(n, np.int32) for _ in range(7)]
    so = np.zeros(n, np.int64)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    err = emul.emul_bam_parse(p(a), p(ro), C.c_int64(n), C.c_int(putative), *[p(c) for c in cols], p(so))
    assert err == 0
    tid, lq, xf, hp, flag, state, nlen = cols
    seen = set()
    for i, t in enumerate(truth):
        assert (tid[i], lq[i], flag[i], hp[i]) == (t["tid"], t["l_qseq"], t["flag"], t["hp"]), i
        assert xf[i] == t["xf"]
        if t["flag"] & (0x4 | 0x800 | 0x100):
            want = 0
        elif t["l_qseq"] < 100:
            want = 3
        else:
            want = 1 if (putative and t["has_xf"] and t["xf"] != 0) else 2
        assert state[i] == want, i
        seen.add(want)
        assert nlen[i] == (len(t["qname"]) if want in (1, 2) else 0)
        assert win[so[i]:so[i] + (t["l_qseq"] + 1) // 2] == t["seq4"]
        if want == 2 and i % 17 == 0:
            out = np.zeros(t["l_qseq"], np.uint8)
            s4 = np.frombuffer(t["seq4"], np.uint8)
            emul.emul_bam_decode(p(s4), C.c_int32(t["l_qseq"]), p(out))
            codes = np.stack([s4 >> 4, s4 & 15], 1).reshape(-1)[:t["l_qseq"]]
            lut = np.full(16, 5, np.uint8); lut[1], lut[2], lut[4], lut[8] = 1, 2, 3, 4
            assert np.array_equal(out, lut[codes])
    assert seen == ({0, 1, 2, 3} if putative else {0, 2, 3})


def test_malformed_records_are_reported(emul):
    good = record(b"r1", 0, 0, 1, b"\x12" * 60, 120, b"\x20" * 120, [(120 << 4)], b"XFC\2")
    cases = [record(b"r2", 0, 0, 1, b"\x12" * 60, 120, b"\x20" * 100),              # qualities cut short
             record(b"r3", 0, 0, 1, b"\x12" * 60, 120, b"\x20" * 120, aux=b"XFq\1"),  # unknown aux type
             record(b"r4", 0, 0, 1, b"\x12" * 60, 120, b"\x20" * 120, aux=b"baBs" + struct.pack("<i", 1000))]   # B array past the record
    for bad in cases:
        win = good + bad + b"\0" * 64
        offs, _ = py_walk(win, 0, len(win) - 64)
        a = np.frombuffer(win, np.uint8)
        ro = np.array(offs, np.int64)
        cols = [np.zeros(2, np.int32) for _ in range(7)]
        so = np.zeros(2, np.int64)
        p = lambda x: x.ctypes.data_as(C.c_void_p)
        assert emul.emul_bam_parse(p(a), p(ro), C.c_int64(2), C.c_int(1), *[p(c) for c in cols], p(so)) == 1
        assert cols[5][0] == 1 and cols[5][1] == 0            # the good record parsed, the bad one dropped


def test_alignment_mode_cigar_endpos_and_cg_tag(emul):
    """the Clusterer's view of a record (clusterer.cpp:58-153): no length filter, bam_endpos, and the CG:B,I convention
    of records with more than 65535 CIGAR ops (SAM spec 4.2.2)"""
    M, I, D, N, S, EQ, X = 0, 1, 2, 3, 4, 7, 8
    op = lambda n, o: (n << 4) | o
    seq = lambda l: (b"\x12" * ((l + 1) // 2), b"\x20" * l)
    recs, want = [], []
    s4, q = seq(150)
    recs.append(record(b"a", 0, 1, 1000, s4, 150, q, [op(10, S), op(100, M), op(5, I), op(20, D), op(35, EQ)], b"HPc\1", mapq=37))
    want.append((1, 1000, 1000 + 100 + 20 + 35, 5, 37))
    s4, q = seq(50)                                                         # short reads stay (the search loader drops them)
    recs.append(record(b"bb", 16, 0, 7, s4, 50, q, [op(50, M)]))
    want.append((1, 7, 57, 1, 60))
    s4, q = seq(120)
    recs.append(record(b"c", 4, 0, 7, s4, 120, q, [op(120, M)]))            # unmapped: dropped
    want.append((0, 7, 127, 1, 60))
    s4, q = seq(120)
    recs.append(record(b"d", 0, 0, 500, s4, 120, q, []))                     # no CIGAR: bam_endpos = pos + 1
    want.append((1, 500, 501, 0, 60))
    real = [op(60, M), op(3, D), op(30, X), op(2, N), op(30, M)]
    s4, q = seq(120)
    cg = b"CGBI" + struct.pack("<i", len(real)) + b"".join(struct.pack("<I", c) for c in real)
    recs.append(record(b"long", 0, 2, 9000, s4, 120, q, [op(120, S), op(125, N)], b"NMI" + struct.pack("<I", 3) + cg + b"XFC\0"))
    want.append((1, 9000, 9000 + 60 + 3 + 30 + 2 + 30, len(real), 60))
    s4, q = seq(120)
    recs.append(record(b"notcg", 0, 2, 9000, s4, 120, q, [op(100, S), op(125, N)], cg))   # the field is not the placeholder: it stands
    want.append((1, 9000, 9125, 2, 60))
    win = b"".join(recs) + b"\0" * 64
    offs, _ = py_walk(win, 0, len(win) - 64)
    n = len(offs)
    a = np.frombuffer(win, np.uint8)
    ro = np.array(offs, np.int64)
    cols = [np.zeros(n, np.int32) for _ in range(5)]
    co = np.zeros(n, np.int64)
    mq = np.zeros(n, np.int32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    assert emul.emul_bam_parse_aln(p(a), p(ro), C.c_int64(n), p(cols[0]), p(cols[1]), p(cols[2]), p(cols[3]), p(cols[4]), p(co), p(mq)) == 0
    state, nlen, pos, endpos, ncig = cols
    for i, (st, ps, en, nc, mapq) in enumerate(want):
        assert (state[i], pos[i], endpos[i], ncig[i], mq[i]) == (st, ps, en, nc, mapq), i
    got_real = [struct.unpack_from("<I", win, int(co[4]) + 4 * k)[0] for k in range(len(real))]
    assert got_real == real
    assert nlen[1] == 2 and nlen[2] == 0
