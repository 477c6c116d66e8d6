"""Literal Python transcription of the reference Smoother (smoother.cpp:90-345) for tests."""
import math


def mismatch_rate(rec, ref_seq):                       # compute_maxaccuracy, smoother.cpp:310-339
    ref_off, ins, mat, clip = rec["pos"], 0, 0, 0
    nm = nmm = 0
    for ln, op in rec["cigar"]:
        if op in "M=X":
            for j in range(ln):
                if ref_seq[ref_off + j] == rec["seq"][mat + ins + clip + j]:
                    nm += 1
                else:
                    nmm += 1
            ref_off += ln; mat += ln
        elif op == "I":
            ins += ln
        elif op == "D":
            ref_off += ln
        elif op == "S":
            clip += ln
        else:
            break
    return nmm / nm if nm else (math.inf if nmm else math.nan)


def percentile(x, q):                                  # smoother.cpp:254-263
    idx = (len(x) - 1) * q
    lo, hi = math.floor(idx), math.ceil(idx)
    h = idx - lo
    return (1.0 - h) * x[lo] + h * x[hi]


def accept(rec, ref_names, ref_seqs, min_mapq=20):     # smoother.cpp:508-540
    if rec["flag"] & (0x4 | 0x800 | 0x100) or rec["mapq"] < min_mapq or len(rec["seq"]) < 2:
        return False
    return ref_names[rec["tid"]] in ref_seqs


def smooth_read(rec, ref_seq, al_accuracy, min_indel_length=20):   # smoother.cpp:90-239
    """Returns (xf, new record or None): seq, qual, cigar of the rebuilt entry when xf == 0."""
    seq, qual = rec["seq"], rec["qual"]
    new_seq, new_qual, new_cigar = [], bytearray(), []
    ref_off, ins, mat, clip, m_diff = rec["pos"], 0, 0, 0, 0
    nm = nmm = 0
    ignore = True
    for ln, op in rec["cigar"]:
        at = clip + mat + ins
        if op in "M=X":
            new_seq.append(ref_seq[ref_off:ref_off + ln]); new_qual += qual[at:at + ln]
            for j in range(ln):
                if ref_seq[ref_off + j] == seq[at + j]:
                    nm += 1
                else:
                    nmm += 1
            ref_off += ln; mat += ln
            if new_cigar and new_cigar[-1][1] == "M":
                new_cigar[-1] = (new_cigar[-1][0] + ln + m_diff, "M")
            else:
                new_cigar.append((ln + m_diff, "M"))
            m_diff = 0
        elif op == "I":
            if ln > min_indel_length:
                ignore = False
                new_seq.append(seq[at:at + ln]); new_qual += qual[at:at + ln]
                new_cigar.append((ln, "I"))
            ins += ln
        elif op == "D":
            if ln <= min_indel_length:
                new_seq.append(ref_seq[ref_off:ref_off + ln]); new_qual += qual[at:at + ln]
                m_diff += ln
            else:
                ignore = False
                new_cigar.append((ln, "D"))
            ref_off += ln
        elif op == "S":
            ignore = False
            new_seq.append(seq[at:at + ln]); new_qual += qual[at:at + ln]
            clip += ln
            new_cigar.append((ln, "S"))
        else:
            break
    rate = nmm / nm if nm else (math.inf if nmm else math.nan)
    if rate > al_accuracy:
        return 1, None
    if ignore:
        return 2, None
    return 0, dict(seq="".join(new_seq), qual=bytes(new_qual), cigar=new_cigar)


def run(records, ref_names, ref_seqs, min_mapq=20, accp=0.98):
    """Expected output records (same dict keys as bam_writer.read_bam) in output order."""
    acc = []
    for rec in records:
        if len(acc) >= 10000:
            break
        if accept(rec, ref_names, ref_seqs, min_mapq):
            acc.append(mismatch_rate(rec, ref_seqs[ref_names[rec["tid"]]]))
    acc.sort()
    import numpy as np
    al = percentile(acc, float(np.float32(accp)))        # config.hpp:77 keeps accp as a float
    out = []
    for rec in records:
        if not accept(rec, ref_names, ref_seqs, min_mapq):
            continue
        xf, new = smooth_read(rec, ref_seqs[ref_names[rec["tid"]]], al)
        o = dict(rec)
        if new:
            o.update(new)
        o["xf"] = xf
        out.append(o)
    return out, al
