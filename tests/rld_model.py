"""Independent Python statement of ropebwt3's `.fmd` (rld0 "RLD\\3") format for tests of
svdss_b200/host/rld.hpp: an encoder, a sequential decoder and a frame-based rank that follows
rld_locate_blk.  Restated from memory like rld.hpp (ropebwt3 is not in the reference tree), so the
two agree by construction on what was remembered; what these tests catch are coding slips, the
chunk/block bookkeeping and the frame table."""
import struct

LBITS = 23
M64 = (1 << 64) - 1


def ilog2(v):
    return v.bit_length() - 1


def offset0(asize, ty):
    a1 = asize + 1
    return [(a1 * 16 + 63) // 64, (a1 * 32 + 63) // 64, a1][ty]


def stail_of(o, ssize, lbits=LBITS):
    return o + ssize - (2 if (o + ssize) % (1 << lbits) == 0 else 1)


def header(words, o, asize):
    """(type, [total, per-code counts...]) of the block at word o."""
    ty = words[o] >> 62
    vals = []
    for j in range(asize + 1):
        if ty == 0:
            v = words[o + (j >> 2)] >> (16 * (j & 3)) & 0xffff
            if j == 3:
                v &= 0x3fff
        elif ty == 1:
            v = words[o + (j >> 1)] >> (32 * (j & 1)) & 0xffffffff
            if j == 1:
                v &= 0x3fffffff
        else:
            v = words[o + j] & (0x3fffffffffffffff if j == 0 else M64)
        vals.append(v)
    return ty, vals


def parse(path):
    data = open(path, "rb").read()
    assert data[:4] == b"RLD\3"
    # rld_dump: magic, u32 asize << 16 | sbits, 8 reserved bytes (zero), n_bytes (a multiple of 8), n_frames, then mcnt[1..asize]
    a, reserved, n_bytes, n_frames = struct.unpack_from("<IQQQ", data, 4)
    assert reserved == 0 and n_bytes % 8 == 0
    n_words = n_bytes // 8
    asize, sbits = a >> 16, a & 0xffff
    o = 32
    mcnt = list(struct.unpack_from("<%dQ" % asize, data, o)); o += 8 * asize
    words = list(struct.unpack_from("<%dQ" % n_words, data, o)); o += 8 * n_words
    frame = list(struct.unpack_from("<%dQ" % (n_frames * (asize + 1)), data, o)); o += 8 * n_frames * (asize + 1)
    assert o == len(data)
    return dict(asize=asize, sbits=sbits, mcnt=mcnt, words=words, frame=frame, n_frames=n_frames)


def block_runs(f, o):
    """(length, symbol) pairs of the block at word o, read bit by bit (Elias delta + abits symbol)."""
    asize, ssize, words = f["asize"], 1 << f["sbits"], f["words"]
    abits = ilog2(asize) + 1
    ty = words[o] >> 62
    beg, end = (o + offset0(asize, ty)) * 64, (stail_of(o, ssize) + 1) * 64

    def bit(i):                              # the closing block is cut after its counters: zeros beyond
        return words[i >> 6] >> (63 - (i & 63)) & 1 if (i >> 6) < len(words) else 0

    def bits(i, n):
        v = 0
        for k in range(n):
            v = v << 1 | bit(i + k)
        return v
    runs, i = [], beg
    while i < end:
        z = 0
        while i + z < end and z < 6 and bit(i + z) == 0:
            z += 1
        if z == 6 or i + z >= end:
            break
        nb = bits(i + z, z + 1) - 1          # gamma: z zeros, then N + 1 in z + 1 bits
        i += 2 * z + 1
        ln = (1 << nb) | bits(i, nb)
        i += nb
        runs.append((ln, bits(i, abits)))
        i += abits
        assert i <= end
    return runs


def decode(f):
    ssize, words = 1 << f["sbits"], f["words"]
    last = len(words) >> f["sbits"] << f["sbits"]
    out = bytearray()
    for o in range(0, last, ssize):
        runs = block_runs(f, o)
        _, h = header(words, o + ssize, f["asize"])
        assert h[0] == sum(l for l, _ in runs)
        for c in range(f["asize"]):
            assert h[1 + c] == sum(l for l, s in runs if s == c)
        for l, c in runs:
            out += bytes([c]) * l
    assert block_runs(f, last) == [] and len(words) == last + offset0(f["asize"], words[last] >> 62)
    for c in range(f["asize"]):
        assert f["mcnt"][c] == out.count(bytes([c]))
    return bytes(out)


def rank_all(f, k):
    """Occurrences of every code in bwt[0, k) through the frame table, the way rld_locate_blk enters
    the stream: frame -> walk whole blocks by their counters -> decode inside the block."""
    asize, ssize, words = f["asize"], 1 << f["sbits"], f["words"]
    n = sum(f["mcnt"])
    n_blks = len(words) // ssize + 1
    ibits = ilog2(n // n_blks) + 4 if n // n_blks > 0 else 3
    assert f["n_frames"] == ((n + (1 << ibits) - 1) >> ibits) + 1
    z = f["frame"][(k >> ibits) * (asize + 1):(k >> ibits) * (asize + 1) + asize + 1]
    o, cnt = z[0], list(z[1:])
    s = sum(cnt)
    assert s <= k
    while True:
        _, h = header(words, o + ssize, asize)
        if s + h[0] > k:
            break
        for c in range(asize):
            cnt[c] += h[1 + c]
        s += h[0]
        o += ssize
    for l, c in block_runs(f, o):
        t = min(l, k - s)
        cnt[c] += t
        s += t
        if s == k:
            break
    assert s == k
    return cnt


def encode(path, bwt, asize=6, sbits=3):
    """Writer: same stream rules as rld.hpp (a pair never fills a block to its last bit)."""
    ssize = 1 << sbits
    abits = ilog2(asize) + 1
    words = [0] * ssize
    cnt, mcnt = [0] * (asize + 1), [0] * (asize + 1)
    st = dict(shead=0, bitpos=offset0(asize, 0) * 64)

    def next_block():
        st["shead"] += ssize
        sh = st["shead"]
        words.extend([0] * ssize)
        d = [cnt[i] - mcnt[i] for i in range(asize + 1)]
        ty = 0 if d[0] < 0x4000 else 1 if d[0] < 0x40000000 else 2
        for j, v in enumerate(d):
            if ty == 0:
                words[sh + (j >> 2)] |= v << (16 * (j & 3))
            elif ty == 1:
                words[sh + (j >> 1)] |= v << (32 * (j & 1))
            else:
                words[sh + j] = v
        words[sh] |= ty << 62
        st["bitpos"] = (sh + offset0(asize, ty)) * 64
        mcnt[:] = cnt

    i, n = 0, len(bwt)
    while i < n:
        j = i + 1
        while j < n and bwt[j] == bwt[i]:
            j += 1
        l, c = j - i, bwt[i]
        nb = ilog2(l)
        z = ilog2(nb + 1)
        code = ((nb + 1) << nb | (l ^ (1 << nb))) << abits | c
        w = 2 * z + 1 + nb + abits
        end = (stail_of(st["shead"], ssize) + 1) * 64
        if st["bitpos"] + w >= end:          # would touch or pass the last bit of the block
            next_block()
        for k in range(w):
            if code >> (w - 1 - k) & 1:
                p = st["bitpos"] + k
                words[p >> 6] |= 1 << (63 - (p & 63))
        st["bitpos"] += w
        cnt[0] += l
        cnt[c + 1] += l
        i = j
    next_block()
    n_words = st["bitpos"] // 64
    del words[n_words:]
    with open(path, "wb") as f:
        f.write(b"RLD\3" + struct.pack("<IQQQ", asize << 16 | sbits, 0, 8 * n_words, 0) + struct.pack("<%dQ" % asize, *cnt[1:]) +
                struct.pack("<%dQ" % n_words, *words))
