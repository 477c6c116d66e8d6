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
        body = core + qn + cigb + bytes(sb) + b"\xff" * len(seq) + aux
        out += struct.pack("<i", len(body)) + body
    with open(path, "wb") as f:
        data = bytes(out)
        for o in range(0, len(data), 60000):
            f.write(_bgzf_block(data[o:o + 60000]))
        f.write(_bgzf_block(b""))
