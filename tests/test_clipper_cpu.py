"""CPU: `SVDSS call --clipped` (reference clipper.cpp, clusterer.cpp:211-226,339-345, caller.cpp:37-55)
against the literal Python transcriptions in tests/cluster_model.py and tests/clipper_model.py.
The clip extraction runs through `call --cluster-only --clipped --clips FILE` (svb_cluster_batch: those two tests
carry the gpu marker; tests/test_cluster_emul.py holds the same kernel source against the same world on the CPU);
Clipper::call runs through the `_clipper` hook of the shell and needs no GPU."""
import os
import subprocess

import numpy as np
import pytest

import clipper_model
import cluster_model
from bam_writer import write_bam
from svdss_b200 import build


@pytest.fixture(scope="module")
def exe():
    build.build_lib()
    return build.build_host()


def _seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, n))


@pytest.fixture(scope="module")
def clip_world(tmp_path_factory):
    """Two contigs; reads whose SFSs sit inside soft clips (left, right, both), inside hard-clipped
    reads, spanning the whole read, and ordinary placed ones."""
    d = str(tmp_path_factory.mktemp("clip"))
    rng = np.random.default_rng(5)
    names = ["chrA", "chrB", "chrC"]
    seqs = {n: _seq(rng, 40_000) for n in names}
    fa = os.path.join(d, "ref.fa")
    with open(fa, "w") as f:
        for n in names[:2]:                                   # chrC is in the BAM header but not in the FASTA
            f.write(">%s\n%s\n" % (n, seqs[n]))
    records, sfs = [], {}

    def add(qname, tid, pos, cigar, sfs_list, mapq=60, flag=0):
        qlen = sum(l for l, op in cigar if op in "MIS=X")
        records.append(dict(qname=qname, flag=flag, tid=tid, pos=pos, mapq=mapq, cigar=cigar, seq=_seq(rng, qlen), hp=None))
        if sfs_list:
            sfs[qname] = [(qs, ln, 0) for qs, ln in sfs_list]

    # insertion site on chrA: right clips ending at 20000, left clips starting at 19990 (Clipper::call only
    # pairs a left clip with the first right clip at a HIGHER position, clipper.cpp:107-122,164-186)
    for i in range(4):
        add("insR%d" % i, 0, 17000 + 10 * i, [(3000 - 10 * i, "M"), (400 + i, "S")], [(3100, 150)])
    for i in range(3):
        add("insL%d" % i, 0, 19990, [(350 + 5 * i, "S"), (2500, "M")], [(40, 120)])
    # deletion site on chrB: six right clips at 10000, five left clips at 15000
    for i in range(6):
        add("delR%d" % i, 1, 8000, [(2000, "M"), (300, "S")], [(2050, 100), (2200, 50)])
    for i in range(5):
        add("delL%d" % i, 1, 15000, [(280, "S"), (10, "M"), (5, "I"), (1990, "M")], [(20, 200)])
    add("both", 0, 30000, [(200, "S"), (1000, "M"), (200, "S")], [(10, 100), (1250, 100)])    # a left and a right clip
    add("lone", 0, 5000, [(150, "S"), (1500, "M")], [(10, 100)])                              # w = 1: filtered
    add("hardL", 0, 6000, [(100, "H"), (1500, "M")], [(0, 50)])                               # first op is not S: s_unplaced
    add("hardR", 0, 7000, [(1500, "M"), (100, "H")], [(1480, 20)])                            # last op is not S: e_unplaced
    add("whole", 0, 9000, [(1200, "M")], [(0, 1200)])                                         # covers the read: unplaced
    add("placed", 0, 12000, [(800, "M"), (60, "I"), (800, "M")], [(780, 100)])                # an ordinary SFS
    add("placed2", 0, 12100, [(700, "M"), (60, "I"), (800, "M")], [(680, 100)])
    add("lowq", 1, 15000, [(280, "S"), (2000, "M")], [(20, 200)], mapq=3)                     # dropped by --min-mapq
    add("onC", 2, 1000, [(100, "S"), (1000, "M")], [(10, 50)])                                # chromosome without sequence
    add("nosfs", 1, 8000, [(2000, "M"), (300, "S")], [])                                      # no SFS: not a clip
    records.sort(key=lambda r: (r["tid"], r["pos"]))
    bam = os.path.join(d, "clip.bam")
    write_bam(bam, [(n, 40_000) for n in names],
              [dict(qname=r["qname"], flag=r["flag"], tid=r["tid"], pos=r["pos"], mapq=r["mapq"], seq=r["seq"], cigar=r["cigar"], tags={})
               for r in records])
    sfsf = os.path.join(d, "clip.sfs")
    with open(sfsf, "w") as f:
        for r in records:
            for k, (qs, ln, h) in enumerate(sfs.get(r["qname"], [])):
                f.write("%s\t%d\t%d\t%d\t\n" % (r["qname"] if k == 0 else "*", qs, ln, h))
    return dict(d=d, fa=fa, bam=bam, sfs=sfsf, names=names, ref_seqs={n: seqs[n] for n in names[:2]}, records=records, sfs_by_read=sfs)


def _clip_lines(clips):
    return "".join("%s\t%s\t%d\t%d\t%s\n" % (n, c, p, l, "L" if st else "R") for n, c, p, l, st in clips)


@pytest.mark.gpu
@pytest.mark.parametrize("threads", [1, 3, 4])
def test_clips_match_the_transcription(exe, clip_world, threads):
    w = clip_world
    out = os.path.join(w["d"], "clips_%d.tsv" % threads)
    r = subprocess.run([exe, "call", "--reference", w["fa"], "--bam", w["bam"], "--sfs", w["sfs"], "--threads", str(threads),
                        "--cluster-only", "--clipped", "--clips", out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exp = []
    cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=threads, clips_out=exp)
    assert len(exp) == 4 + 3 + 6 + 5 + 2 + 1
    assert open(out).read() == _clip_lines(exp)
    assert "1/1/1 unplaced SFSs. 0 erroneus SFSs. %d clipped SFSs." % len(exp) in r.stderr   # whole / hardL / hardR


@pytest.mark.gpu
def test_without_clipped_the_same_sfss_count_as_unplaced(exe, clip_world):
    w = clip_world
    out = os.path.join(w["d"], "clips_off.tsv")
    r = subprocess.run([exe, "call", "--reference", w["fa"], "--bam", w["bam"], "--sfs", w["sfs"], "--cluster-only", "--clips", out],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert open(out).read() == ""
    # left: 3 insL + 5 delL + both + lone + hardL; right: 4 insR + 2*6 delR + both + hardR
    assert "1/11/18 unplaced SFSs. 0 erroneus SFSs. 0 clipped SFSs." in r.stderr


def run_hook(exe, d, tag, fa, clips, regions, threads):
    cf, rf = os.path.join(d, "c_%s.tsv" % tag), os.path.join(d, "r_%s.tsv" % tag)
    with open(cf, "w") as f:
        f.write(_clip_lines([c[:5] for c in clips]))
    with open(rf, "w") as f:
        f.write("".join("%d %d\n" % r for r in regions))
    r = subprocess.run([exe, "_clipper", "--reference", fa, "--clips-in", cf, "--regions-in", rf, "--threads", str(threads), "--verbose"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    stages = {"combined": {"R": [], "L": []}, "clustered": {"R": [], "L": []}}
    for line in r.stderr.splitlines():
        t = line.split("\t")
        if t[0] in stages:
            stages[t[0]][t[1]].append(("", t[2], int(t[3]), int(t[4]), t[1] == "L", int(t[5])))
    return r.stdout.splitlines(), stages


def test_clipper_on_the_world_clips(exe, clip_world):
    w = clip_world
    clips = []
    cluster_model.run(w["records"], w["names"], w["ref_seqs"], w["sfs_by_read"], threads=4, clips_out=clips)
    clips = [c + (0,) for c in clips]
    got, st = run_hook(exe, w["d"], "world", w["fa"], clips, [], 4)
    p, rcl, lcl = clipper_model.call(clips, w["names"][:2], w["ref_seqs"], 4, [], st["combined"]["R"], st["combined"]["L"])
    assert st["clustered"]["R"] == rcl and st["clustered"]["L"] == lcl
    assert got == clipper_model.vcf_lines(p)
    # the planted events: INS at chrA:20000 (4 right clips beat 3 left clips -> s = 20000, l = longest clip), DEL chrB:10000-15000
    assert len(got) == 2
    ins = [l for l in got if "SVTYPE=INS" in l][0].split("\t")
    assert ins[0] == "chrA" and ins[1] == "20000" and "SVLEN=403;" in ins[7] and "WEIGHT=4;" in ins[7] and ins[4] == "<INS>"
    dele = [l for l in got if "SVTYPE=DEL" in l][0].split("\t")
    assert dele[0] == "chrB" and dele[1] == "10000" and "SVLEN=-5001;" in dele[7] and "WEIGHT=6;" in dele[7] and "IMPRECISE" in dele[7]
    # a called SV within 1000 bp of the insertion breakpoints removes that call, the deletion stays
    got2, _ = run_hook(exe, w["d"], "world_reg", w["fa"], clips, [(20500 - 1000, 20600 + 1000)], 4)
    assert [l.split("\t")[0] for l in got2] == ["chrB"]


@pytest.mark.parametrize("seed,threads", [(1, 1), (2, 2), (3, 4), (4, 3), (5, 4), (6, 8)])
def test_clipper_random_clip_sets(exe, tmp_path, seed, threads):
    rng = np.random.default_rng(100 + seed)
    names = ["c%d" % i for i in range(6)]
    seqs = {n: _seq(rng, 120_000) for n in names}
    fa = str(tmp_path / "r.fa")
    with open(fa, "w") as f:
        for n in names:
            f.write(">%s\n%s\n" % (n, seqs[n]))
    # breakpoints: a few hot positions per chromosome (so that reads pile up), chains within 1000 bp,
    # positions below 1000 (unsigned wrap in cluster()), the same coordinate on different chromosomes
    hot = {n: sorted(set(int(x) for x in np.concatenate([rng.integers(0, 110_000, 10), rng.integers(0, 1500, 2),
                                                          [30_000, 30_400, 30_900, 31_500, 36_000]]))) for n in names}
    clips = []
    for i in range(700):
        n = names[int(rng.integers(0, 6))]
        p = hot[n][int(rng.integers(0, len(hot[n])))]
        if rng.random() < 0.2:
            p += int(rng.integers(0, 3000))
        name = "r%d" % (i if rng.random() < 0.9 else int(rng.integers(0, 700)))      # some duplicated read names
        clips.append((name, n, p, int(rng.integers(1, 3000)), bool(rng.random() < 0.5), 0))
    regions = [(int(a) - 1000, int(a) + int(b) + 1000) for a, b in zip(rng.integers(0, 110_000, 12), rng.integers(1, 5000, 12))]
    got, st = run_hook(exe, str(tmp_path), "rnd", fa, clips, regions, threads)
    p, rcl, lcl = clipper_model.call(clips, names, seqs, threads, regions, st["combined"]["R"], st["combined"]["L"])
    assert st["clustered"]["R"] == rcl and st["clustered"]["L"] == lcl
    assert len(rcl) > 10 and len(lcl) > 10
    assert got == clipper_model.vcf_lines(p)
    assert any("SVTYPE=INS" in l for l in got) and any("SVTYPE=DEL" in l for l in got)


def test_partner_left_of_every_clip_is_the_first_clip(exe, tmp_path):
    """clipper.cpp:118-121 recurses with end = 0u - 1 here (out-of-bounds read); we return clip 0."""
    fa = str(tmp_path / "r.fa")
    with open(fa, "w") as f:
        f.write(">c0\n%s\n" % ("ACGT" * 5000))
    clips = [("a", "c0", 5000, 100, True, 0), ("b", "c0", 5000, 120, True, 0),          # left clips at 5000 (w = 2)
             ("c", "c0", 5300, 80, False, 0), ("d", "c0", 5300, 90, False, 0),          # right clips at 5300 and 9000
             ("e", "c0", 9000, 80, False, 0), ("f", "c0", 9000, 90, False, 0)]
    got, _ = run_hook(exe, str(tmp_path), "edge", fa, clips, [], 2)
    assert len(got) == 1 and got[0].split("\t")[:2] == ["c0", "5300"] and "SVTYPE=INS" in got[0] and "SVLEN=120;" in got[0]
