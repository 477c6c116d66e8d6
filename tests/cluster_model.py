"""Literal Python transcription of the reference Clusterer (clusterer.cpp) for tests: the aligned-pair
lists are materialised exactly like bam.cpp:92-134 and every loop follows the reference line by line,
so that the C++ host shell (svdss_b200/host/clusterer.hpp, which walks CIGARs instead) is checked
against an independent statement.  Records are dicts: qname, flag, tid, pos, mapq, cigar [(len, op)],
seq (ACGTN str), hp (int or None)."""


def get_aligned_pairs(pos, cigar):                      # bam.cpp:92-134
    out, ref, rd = [], pos, 0
    for ln, op in cigar:
        if op in "M=X":
            for _ in range(ln):
                out.append((rd, ref)); rd += 1; ref += 1
        elif op in "IS":
            for _ in range(ln):
                out.append((rd, -1)); rd += 1
        elif op in "DN":
            for _ in range(ln):
                out.append((-1, ref)); ref += 1
    return out


def get_unique_kmers(alpairs, k, from_end, cseq):       # clusterer.cpp:350-403
    if len(alpairs) < k:
        return (-1, -1)
    kmers = {}
    i = 0
    while i < len(alpairs) - k + 1:
        skip = False
        for j in range(i, i + k):
            if alpairs[j][0] == -1 or alpairs[j][1] == -1:
                skip = True; i = j + 1
                break
        if skip:
            continue
        km = cseq[alpairs[i][1]:alpairs[i][1] + k]
        kmers[km] = kmers.get(km, 0) + 1
        i += 1
    last = (-1, -1)
    i = 0
    while i < len(alpairs) - k + 1:
        off = len(alpairs) - k - i if from_end else i
        skip = False
        for j in range(off, off + k):
            if alpairs[j][0] == -1 or alpairs[j][1] == -1:
                skip = True; i += j - off
                break
        if skip:
            i += 1
            continue
        last = alpairs[off]
        if kmers.get(cseq[alpairs[off][1]:alpairs[off][1] + k], 0) == 1:
            break
        i += 1
    return last


def extend_alignment(rec, sfs_list, chrom, cseq, flank=100, ksize=7, clips=None):   # clusterer.cpp:156-345
    """clips: None (no --clipped) or a list that receives (qname, chrom, p, l, starting) tuples."""
    alpairs = get_aligned_pairs(rec["pos"], rec["cigar"])
    last_pos = 0
    local = []
    lclip = rclip = (0, 0)
    for qs_, l_, htag in sfs_list:
        s, e = qs_, qs_ + l_ - 1
        aln_start = aln_end = refs = refe = -1
        for i in range(last_pos, len(alpairs)):
            q, r = alpairs[i]
            if q == -1 or r == -1:
                continue
            elif q < s:
                last_pos = i; refs = r; aln_start = i
            elif q > e:
                refe = r; aln_end = i
                break
        if refs == -1 and refe == -1:                    # :206-210
            continue
        elif refs == -1:                                 # :211-218
            ln, op = rec["cigar"][0]
            if op == "S" and clips is not None:
                lclip = (rec["pos"], ln)
            continue
        elif refe == -1:                                 # :219-226
            ln, op = rec["cigar"][-1]
            if op == "S" and clips is not None:
                rclip = (endpos(rec), ln)
            continue
        local_al = []
        last_r = refs - 1
        for i in range(aln_start, aln_end + 1):
            q, r = alpairs[i]
            if r == -1:
                if refs <= last_r <= refe:
                    local_al.append((q, r))
            else:
                last_r = r
                if refs <= r <= refe:
                    local_al.append((q, r))
            if q != -1 and r != -1 and r >= refe:
                break
        pre = alpairs[max(0, aln_start - flank):aln_start]
        post = alpairs[aln_end + 1:aln_end + 1 + flank]
        prek = get_unique_kmers(pre, ksize, True, cseq)
        postk = get_unique_kmers(post, ksize, False, cseq)
        if prek[0] == -1 or prek[1] == -1:
            prek = local_al[0]
        if postk[0] == -1 or postk[1] == -1:
            postk = local_al[-1]
        if -1 in prek or -1 in postk:
            continue
        if prek[1] > postk[1] + ksize:
            continue
        local.append(dict(chrom=chrom, qname=rec["qname"], rs=prek[1], re=postk[1] + ksize, qs=prek[0], qe=postk[0] + ksize, htag=htag))
    merged = []
    for x in local:
        for m in merged:
            if (x["rs"] <= m["rs"] <= x["re"]) or (m["rs"] <= x["rs"] <= m["re"]):
                m["rs"] = min(m["rs"], x["rs"]); m["re"] = max(m["re"], x["re"])
                m["qs"] = min(m["qs"], x["qs"]); m["qe"] = max(m["qe"], x["qe"])
                break
        else:
            merged.append(dict(x))
    if clips is not None:                                # :339-345
        if lclip[1] > 0:
            clips.append((rec["qname"], chrom, lclip[0], lclip[1], True))
        if rclip[1] > 0:
            clips.append((rec["qname"], chrom, rclip[0], rclip[1], False))
    return merged


def primary(rec):
    return not (rec.get("flag", 0) & (0x4 | 0x800 | 0x100))


def endpos(rec):
    span = sum(l for l, op in rec["cigar"] if op in "MDN=X")
    return rec["pos"] + (span if span else 1)


def run(records, ref_names, ref_seqs, sfs_by_read, threads=4, min_mapq=20, min_cluster_weight=2, clips_out=None):
    """Returns the clusters in the reference's order for `threads`: list of dicts chrom, s, e, cov
    (cov0,cov1,cov2), reads [(0/1, hp)], subreads [(name, seq, hp)].  With `clips_out` (a list) the
    run is `--clipped`: it receives the Clusterer's clips (qname, chrom, p, l, starting) in the
    reference's order (per-slot vectors inserted at the front, clusterer.cpp:24)."""
    # pass 1: accepted reads dealt round-robin to thread slots (clusterer.cpp:109-133)
    p_ext = [[] for _ in range(threads)]
    p_clips = [[] for _ in range(threads)]
    n = 0
    for rec in records:
        if not primary(rec) or rec.get("mapq", 60) < min_mapq or rec["qname"] not in sfs_by_read:
            continue
        chrom = ref_names[rec["tid"]]
        if chrom in ref_seqs:
            p_ext[n % threads].extend(extend_alignment(rec, sfs_by_read[rec["qname"]], chrom, ref_seqs[chrom],
                                                       clips=p_clips[n % threads] if clips_out is not None else None))
        n += 1
    if clips_out is not None:
        for t in range(threads):
            clips_out[0:0] = p_clips[t]
    ext = [x for t in p_ext for x in t]
    if not ext:
        return []
    ext.sort(key=lambda x: (x["chrom"].encode(), x["rs"]))          # stable, like the shell
    dist = int(max(x["re"] - x["rs"] for x in ext) * 1.1)
    intervals, prev_i, prev_e, prev_chrom = [], 0, ext[0]["re"], ext[0]["chrom"]
    for i in range(1, len(ext)):
        x = ext[i]
        if x["chrom"] != prev_chrom:
            prev_chrom = x["chrom"]; intervals.append((prev_i, i - 1)); prev_i = i; prev_e = x["re"]
        elif x["rs"] - prev_e > dist:
            intervals.append((prev_i, i - 1)); prev_e = x["re"]; prev_i = i
    intervals.append((prev_i, len(ext) - 1))
    maps = [dict() for _ in range(threads)]
    for ii, (a, b) in enumerate(intervals):
        mine = maps[ii % threads]
        j = a
        low, high, last_j = ext[j]["rs"], ext[j]["re"], j
        j += 1
        while j <= b:
            x = ext[j]
            if x["rs"] <= high:
                low = min(low, x["rs"]); high = max(high, x["re"])
            else:
                mine.setdefault((low, high), []).extend(ext[last_j:j])
                low, high, last_j = x["rs"], x["re"], j
            j += 1
        mine.setdefault((low, high), []).extend(ext[last_j:b + 1])
    raw = [m[k] for m in maps for k in sorted(m)]
    # fill_clusters (clusterer.cpp:478-610)
    out = []
    for sfss in raw:
        chrom = sfss[0]["chrom"]
        names = set(x["qname"] for x in sfss)
        min_s = min(x["rs"] for x in sfss); max_e = max(x["re"] for x in sfss)
        if len(names) < min_cluster_weight:
            continue
        cov = [0, 0, 0]
        locus, sub = [], []
        beg, end = max(0, min_s - 1), max_e
        for rec in records:
            if ref_names[rec["tid"]] != chrom or not (rec["pos"] < end and endpos(rec) > beg):
                continue
            if not primary(rec) or rec.get("mapq", 60) < min_mapq:
                continue
            hp = rec.get("hp") or 0
            cov[hp] += 1
            locus.append([0, 3 if hp == 0 else hp])
            if rec["qname"] not in names:
                continue
            locus[-1][0] = 1
            alpairs = get_aligned_pairs(rec["pos"], rec["cigar"])
            qs = qe = -1
            for q, r in reversed(alpairs):
                if q == -1 or r == -1:
                    continue
                if r <= min_s:
                    qs = q
                    break
            for q, r in alpairs:
                if q == -1 or r == -1:
                    continue
                if r >= max_e:
                    qe = q
                    break
            if qs == -1 or qe == -1:
                continue
            sub.append((rec["qname"], rec["seq"][qs:qe + 1], hp))
        c = dict(chrom=chrom, s=min_s, e=max_e, subreads=sub, cov=None, reads=[])
        if len(sub) >= min_cluster_weight:
            c["cov"] = tuple(cov); c["reads"] = [tuple(x) for x in locus]
        out.append(c)
    return out


def clusters_text(clusters):                             # clusterer.cpp:613-626
    return "".join("%s:%d-%d\t%d%s\n" % (c["chrom"], c["s"] + 1, c["e"] + 1, len(c["subreads"]),
                                         "".join("\t%s:%s" % (n, s) for n, s, _ in c["subreads"])) for c in clusters)
