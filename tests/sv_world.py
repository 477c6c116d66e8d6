"""Shared fixture for the `call` tests: a small diploid sample with planted SVs, its coordinate-sorted
BAM (own writer), the .sfs file of the oracle's SFS search, and the record list for the Python models."""
import os

import numpy as np

import oracle
from bam_writer import write_bam
from common import oracle_index, fm_results
from svdss_b200 import synth

L = "$ACGTN"


def dec(a):
    return "".join(L[int(x)] for x in a)


def make_world(d, ref_bp=90_000, n_svs=12, coverage=8, seed=71, tag_hp=True, mean_len=4000, sub_rate=0.0, indel_rate=0.0):
    contigs = synth.make_reference(ref_bp, seed=seed, contigs=2, n_repeats=4, n_nruns=1, nrun_len=50)
    names = ["chrA", "chrB"]
    cat = synth.make_sv_catalogue(contigs, n_svs, seed=seed + 1, min_len=50, max_len=600, margin=1500, spacing=1500)
    alns = synth.make_sample_alignments(contigs, cat, coverage=coverage, seed=seed + 2, mean_len=mean_len, sd_len=800,
                                        min_len=1000, max_len=8000, tag_hp=tag_hp, sub_rate=sub_rate, indel_rate=indel_rate)
    fa = os.path.join(d, "ref.fa")
    with open(fa, "w") as f:
        for n, c in zip(names, contigs):
            s = dec(c)
            f.write(">%s desc\n" % n)
            for o in range(0, len(s), 60):
                f.write((s[o:o + 60].lower() if (o // 60) % 7 == 3 else s[o:o + 60]) + "\n")   # some soft-masked lines
    # SFSs of every read from the CPU oracle (assembled, ascending qs) -> .sfs text (ping_pong.cpp:227-228)
    T, SA, bwt = oracle_index(contigs)
    fm = oracle.FMIndex(bwt)
    res, _ = fm_results(fm, [a["seq"] for a in alns])
    rng = np.random.default_rng(seed + 3)
    records, sfs_by_read, lines = [], {}, []
    for i, a in enumerate(alns):
        rec = dict(qname=a["qname"], flag=0, tid=a["tid"], pos=a["pos"], mapq=60, cigar=a["cigar"], seq=dec(a["seq"]),
                   hp=a["hp"] if a["hp"] else None)
        if tag_hp and rng.random() < 0.15:
            rec["hp"] = None                                  # untagged read in a tagged sample
        if rng.random() < 0.03:
            rec["mapq"] = 5                                   # dropped by --min-mapq in both passes
        records.append(rec)
        asm = oracle.assemble(res[i])
        if asm:
            htag = rec["hp"] or 0
            sfs_by_read[a["qname"]] = [(qs, ln, htag) for qs, ln in asm]
            for k, (qs, ln) in enumerate(asm):
                lines.append("%s\t%d\t%d\t%d\t\n" % (a["qname"] if k == 0 else "*", qs, ln, htag))
    # records the filters must drop
    records.insert(3, dict(qname="unmapped_x", flag=4, tid=0, pos=0, mapq=0, cigar=[], seq="ACGT" * 30, hp=None))
    sec = dict(records[5]); sec["flag"] = 0x100; records.insert(6, sec)
    sup = dict(records[9]); sup["flag"] = 0x800; records.insert(10, sup)
    bam = os.path.join(d, "sample.bam")
    write_bam(bam, [(n, len(c)) for n, c in zip(names, contigs)],
              [dict(qname=r["qname"], flag=r["flag"], tid=r["tid"], pos=r["pos"], mapq=r["mapq"], seq=r["seq"], cigar=r["cigar"],
                    tags=({"HP": ("C", r["hp"])} if r["hp"] else {})) for r in records])
    sfs = os.path.join(d, "sample.sfs")
    with open(sfs, "w") as f:
        f.write("".join(lines))
    ref_seqs = {n: dec(c) for n, c in zip(names, contigs)}
    return dict(d=d, fa=fa, bam=bam, sfs=sfs, names=names, contigs=contigs, ref_seqs=ref_seqs, records=records,
                sfs_by_read=sfs_by_read, catalogue=cat)
