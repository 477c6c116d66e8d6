"""CPU-only: host-side mirror of the reference interface (.sfs text format, parser, assembler)."""
import io
import os

from svdss_b200 import host
import oracle
import ref_model


def test_sfs_struct_and_assembler_match_reference_semantics():
    s = host.SFS("r1", 10, 5, 2)
    assert (s.qs, s.qe, s.l, s.htag, s.chrom) == (10, 15, 5, 2, "")   # sfs.hpp:43-50
    pairs = [(30, 4), (7, 4), (5, 3), (20, 2), (10, 1), (21, 9)]
    got = host.Assembler.assemble([host.SFS("r", q, l, 0) for q, l in pairs])
    assert [(x.qs, x.l) for x in got] == ref_model.assemble(pairs) == oracle.assemble(pairs)


def test_sfs_text_roundtrip(tmp_path):
    buf = io.StringIO()
    pp = host.PingPong(index=None, out=buf)
    batch = {"readA": [host.SFS("readA", 3, 17, 0), host.SFS("readA", 40, 20, 0)],
             "readB": [], "readC": [host.SFS("readC", 0, 25, 1)]}
    pp.output_batch(batch)
    txt = buf.getvalue()
    # ping_pong.cpp:227-228: name only on a read's first line, trailing TAB before the newline
    assert txt == "readA\t3\t17\t0\t\n*\t40\t20\t0\t\nreadC\t0\t25\t1\t\n"
    p = os.path.join(tmp_path, "x.sfs")
    open(p, "w").write(txt)
    parsed = host.parse_sfsfile(p)      # sfs.cpp:5-30
    assert list(parsed) == ["readA", "readC"]
    assert [(s.qs, s.l, s.htag) for s in parsed["readA"]] == [(3, 17, 0), (40, 20, 0)]
    assert parsed["readC"][0].qname == "readC" and parsed["readC"][0].htag == 1


def test_fastx_reader(tmp_path):
    p = os.path.join(tmp_path, "r.fq")
    open(p, "w").write("@r1 desc\nACGT\nNN\n+\nIIIIII\n@r2\nTTTT\n+\n@@@@\n")
    assert list(host.read_fastx(p)) == [("r1", b"ACGTNN"), ("r2", b"TTTT")]
    p2 = os.path.join(tmp_path, "r.fa")
    open(p2, "w").write(">c1\nACG\nTTA\n>c2\n\nGG\n")
    assert list(host.read_fastx(p2)) == [("c1", b"ACGTTA"), ("c2", b"GG")]
    assert host.NT6[ord("a")] == 1 and host.NT6[ord("N")] == 5 and host.NT6[ord("=")] == 5
