"""Minimal BAM writer for tests (no samtools/htslib offline): BGZF blocks (deflate + BC extra field +
EOF block), header with @SQ, records with the fields the search path reads."""
import struct
import zlib

NT16 = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}


def _bgzf_block(data):
    comp = zlib.compressobj(6, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    bsize = len(body) + 25
    head = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return head + body + struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data))


def write_bam(path, refs, records):
    """refs: [(name, length)]; records: dicts with qname, flag, tid, pos, mapq, seq (ACGTN string),
    cigar [(len, op_char)], tags {"XF": int, "HP": int} (written as typed ints)."""
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    out = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs)))
    for name, ln in refs:
        out += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    for r in records:
        seq = r["seq"]
        qn = r["qname"].encode() + b"\0"
        cig = r.get("cigar", [(len(seq), "M")] if seq else [])
        cigb = b"".join(struct.pack("<I", (l << 4) | "MIDNSHP=X".index(op)) for l, op in cig)
        sb = bytearray((len(seq) + 1) // 2)
        for i, ch in enumerate(seq):
            sb[i >> 1] |= NT16.get(ch, 15) << (0 if i & 1 else 4)
        aux = b""
        for k, v in r.get("tags", {}).items():
            if isinstance(v, tuple):     # (type_char, value)
                ty, val = v
                aux += k.encode() + ty.encode() + struct.pack({"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[ty], val)
            elif isinstance(v, str):
                aux += k.encode() + b"Z" + v.encode() + b"\0"
            else:
                aux += k.encode() + b"i" + struct.pack("<i", v)
        core = struct.pack("<iiBBHHHiiii", r.get("tid", 0), r.get("pos", 0), len(qn), r.get("mapq", 60), 4680,
                           len(cig), r.get("flag", 0), len(seq), -1, -1, 0)
        qual = r.get("qual")
        body = core + qn + cigb + bytes(sb) + (bytes(qual) if qual is not None else b"\xff" * len(seq)) + aux
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as f:
        data = bytes(out)
        for o in range(0, len(data), 60000):
            f.write(_bgzf_block(data[o:o + 60000]))
        f.write(_bgzf_block(b""))


def read_bam(path):
    """Parse a BAM written by anything (BGZF = concatenated gzip members). Returns (header text, refs,
    records): records are dicts qname, flag, tid, pos, mapq, cigar [(len, op)], seq, qual (bytes), tags
    {tag: (type, value)} in file order (integer and Z tags decoded, others kept raw)."""
    import gzip
    data = gzip.decompress(open(path, "rb").read())
    assert data[:4] == b"BAM\1"
    l_text = struct.unpack_from("<i", data, 4)[0]
    text = data[8:8 + l_text].decode()
    o = 8 + l_text
    n_ref = struct.unpack_from("<i", data, o)[0]; o += 4
    refs = []
    for _ in range(n_ref):
        ln = struct.unpack_from("<i", data, o)[0]; o += 4
        name = data[o:o + ln - 1].decode(); o += ln
        refs.append((name, struct.unpack_from("<i", data, o)[0])); o += 4
    recs = []
    while o < len(data):
        bs = struct.unpack_from("<i", data, o)[0]; o += 4
        b = data[o:o + bs]; o += bs
        tid, pos, l_qn, mapq, _bin, n_cig, flag, l_seq = struct.unpack_from("<iiBBHHHi", b, 0)
        p = 32
        qname = b[p:p + l_qn - 1].decode(); p += l_qn
        cig = [(c >> 4, "MIDNSHP=X"[c & 15]) for c in struct.unpack_from("<%dI" % n_cig, b, p)]; p += 4 * n_cig
        sb = b[p:p + (l_seq + 1) // 2]; p += (l_seq + 1) // 2
        seq = "".join("=ACMGRSVTWYHKDBN"[(sb[i >> 1] >> (0 if i & 1 else 4)) & 15] for i in range(l_seq))
        qual = b[p:p + l_seq]; p += l_seq
        tags = {}
        while p < len(b):
            tag, ty = b[p:p + 2].decode(), chr(b[p + 2]); p += 3
            if ty in "cCsSiI":
                fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[ty]
                tags[tag] = (ty, struct.unpack_from(fmt, b, p)[0]); p += struct.calcsize(fmt)
            elif ty == "Z":
                e = b.index(b"\0", p); tags[tag] = (ty, b[p:e].decode()); p = e + 1
            elif ty == "A":
                tags[tag] = (ty, chr(b[p])); p += 1
            elif ty == "f":
                tags[tag] = (ty, struct.unpack_from("<f", b, p)[0]); p += 4
            else:
                raise ValueError("tag type " + ty)
        recs.append(dict(qname=qname, flag=flag, tid=tid, pos=pos, mapq=mapq, cigar=cig, seq=seq, qual=bytes(qual), tags=tags))
    return text, refs, recs
