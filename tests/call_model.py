"""Python restatement of the `call` compute core for tests: split_cluster (untagged branch,
caller.cpp:100-150 + :78-97), run_poa -> ksw2 -> CIGAR walk -> SV (caller.cpp:311-406), SV::operator<<
(sv.cpp:53-80).  POA and ksw2 come from the oracle."""
import numpy as np

import oracle


def get_len(sub):          # clusterer.hpp:103-111, unsigned integer mean
    return sum(len(s[1]) for s in sub) // len(sub)


def split_cluster_by_len(subreads, min_ratio=np.float32(0.97)):
    subreads = list(subreads)
    groups = []
    for sr in subreads:
        for g in groups:
            cl, sl = np.float32(get_len(g)), np.float32(len(sr[1]))
            if min(cl, sl) / max(cl, sl) >= min_ratio:
                g.append(sr)
                break
        else:
            groups.append([sr])
    return groups


def _two_largest(groups):
    i1 = i2 = -1
    v1 = v2 = 0
    for i, g in enumerate(groups):
        if len(g) > v1:
            v2, i2, v1, i1 = v1, i1, len(g), i
        elif len(g) > v2:
            v2, i2 = len(g), i
    return [i for i in (i1, i2) if i != -1]


def _largest(groups):
    v, k = 0, -1
    for i, g in enumerate(groups):
        if len(g) > v:
            v, k = len(g), i
    return k


def split_cluster(subreads, cov=(0, 0, 0, 0), useht=True, min_ratio=np.float32(0.97)):
    """caller.cpp:100-255. subreads: (name, seq[, htag]); cov = (cov, cov0, cov1, cov2).
    Returns [(subreads, (cov, cov0, cov1, cov2))], at most two."""
    sr3 = [(x[0], x[1], x[2] if len(x) > 2 else 0) for x in subreads]
    c0 = [x for x in sr3 if not (useht and x[2] in (1, 2))]
    c1 = [x for x in sr3 if useht and x[2] == 1]
    c2 = [x for x in sr3 if useht and x[2] == 2]
    cv, cv0, cv1, cv2 = cov
    if not c1 and not c2:
        groups = split_cluster_by_len(c0, min_ratio)
        return [(groups[i], (cv, cv0, -1, -1)) for i in _two_largest(groups)]
    both = (1 if c1 else 0) + (2 if c2 else 0)
    g1, g2 = split_cluster_by_len(c1, min_ratio), split_cluster_by_len(c2, min_ratio)
    add1, add2 = [0] * len(g1), [0] * len(g2)
    fresh, fresh_cov0 = [], cv0

    def best(groups, sl):          # best_ratio is an int in the reference: it truncates to 0 or 1
        b, br = -1, -1
        for i, g in enumerate(groups):
            cl = np.float32(get_len(g))
            r = min(cl, sl) / max(cl, sl)
            if r >= min_ratio and r > np.float32(br):
                b, br = i, int(r)
        return b, br

    for sr in c0:
        sl = np.float32(len(sr[1]))
        b1, r1 = best(g1, sl)
        b2, r2 = best(g2, sl)
        if both == 1:
            if b1 == -1:
                fresh.append(sr)
            else:
                g1[b1].append(sr); add1[b1] += 1; fresh_cov0 -= 1
        elif both == 2:
            if b2 == -1:
                fresh.append(sr)
            else:
                g2[b2].append(sr); add2[b2] += 1; fresh_cov0 -= 1
        else:
            if b1 != -1 and r1 > r2:
                g1[b1].append(sr); add1[b1] += 1; fresh_cov0 -= 1
            elif b2 != -1 and r2 > r1:
                g2[b2].append(sr); add2[b2] += 1; fresh_cov0 -= 1
    out = []
    k = _largest(g1)
    if k != -1:
        out.append((g1[k], (cv, -1, cv1 + add1[k], -1)))
    k = _largest(g2)
    if k != -1:
        out.append((g2[k], (cv, -1, -1, cv2 + add2[k])))
    if both != 3:
        gn = split_cluster_by_len(fresh, min_ratio)
        k = _largest(gn)
        if k != -1:
            out.append((gn[k], (cv, fresh_cov0, -1, -1)))
    return out


def fuzz_ratio(a, b):
    """rapidfuzz::fuzz::ratio = 100 * 2*LCS / (|a|+|b|) (normalised Indel similarity)."""
    if not a and not b:
        return 100.0
    if not a or not b:
        return 0.0
    A = np.frombuffer(a.encode(), np.uint8)
    prev = np.zeros(len(b) + 1, np.int64)
    Bv = np.frombuffer(b.encode(), np.uint8)
    for ch in A:
        # LCS row: cur[j] = max(prev[j], cur[j-1], prev[j-1] + (ch == b[j-1])) -> running max of a candidate row
        cand = np.maximum(prev[1:], prev[:-1] + (Bv == ch))
        cur = np.concatenate([[0], np.maximum.accumulate(cand)])
        prev = cur
    return 100.0 * 2.0 * float(prev[-1]) / (len(a) + len(b))


def call_vcf_lines(ref, clusters, min_cluster_weight=2, min_sv_length=25, useht=True, threads=4, min_ratio=0.97,
                   return_stats=False):
    """ref: dict chrom -> upper-case str; clusters: either tuples (chrom, s, e, [(name, seq)]) -- the
    --clusters-in form: htag 0, cov = cov0 = n, empty RVEC -- or the dicts of cluster_model.run (0-based
    inclusive s,e).  Returns the VCF record lines in the order the shell prints them."""
    enc = lambda s: oracle.CHAR26[np.frombuffer(s.encode(), np.uint8)]
    p_recs = [[] for _ in range(threads)]
    for ci, cl in enumerate(clusters):
        if isinstance(cl, tuple):
            chrom, s, e, subreads = cl
            n = len(subreads)
            cov, rvec = (n, n, 0, 0), ""
        else:
            chrom, s, e, subreads = cl["chrom"], cl["s"], cl["e"], cl["subreads"]
            if len(subreads) < min_cluster_weight:
                continue
            cov = (sum(cl["cov"]),) + tuple(cl["cov"])
            rvec = "-".join("%d:%d" % x for x in cl["reads"])
        if len(subreads) < min_cluster_weight:
            continue
        for sub, (cv, cv0, cv1, cv2) in split_cluster(subreads, cov, useht, np.float32(min_ratio)):
            cons_codes = oracle.poa_consensus([enc(x[1]) for x in sub], band=True)
            cons = "".join("ACGTN"[c] for c in cons_codes)
            window = ref[chrom][s:e + 1]
            score, cig = oracle.ksw_extd2(enc(cons), enc(window))
            cigar_str = "".join("%d%s" % (l, op) for l, op in cig)
            rpos, cpos, nv, svs = s, 0, 0, []
            reads = ",".join(x[0] for x in sub)
            for l, op in cig:
                if op == "M":
                    rpos += l; cpos += l
                elif op == "I":
                    if l >= min_sv_length:
                        a = ref[chrom][rpos - 1]
                        svs.append(("INS", rpos, a, a + cons[cpos:cpos + l], l)); nv += 1
                    cpos += l
                else:
                    if l >= min_sv_length:
                        svs.append(("DEL", rpos, ref[chrom][rpos - 1:rpos + l], ref[chrom][rpos - 1], l)); nv += 1
                    rpos += l
            for typ, pos, refall, altall, l in svs:
                end = pos + len(refall) - 1
                idx = "%s_%s:%d-%d_%d" % (typ, chrom, pos, end, l)
                info = ("VARTYPE=SV;SVTYPE=%s;SVLEN=%d;END=%d;WEIGHT=%d;COV=%d;COV0=%d;COV1=%d;COV2=%d;AS=%d;NV=%d;"
                        "CIGAR=%s;RVEC=%s;READS=%s" % (typ, -l if typ == "DEL" else l, end, len(sub), cv, cv0, cv1, cv2, score, nv,
                                                     cigar_str, rvec, reads))
                line = "%s\t%d\t%s\t%s\t%s\t.\tPASS\t%s\tGT:GQ\t0/1:100" % (chrom, pos, idx, refall, altall, info)
                p_recs[ci % threads].append(dict(chrom=chrom, s=pos, e=end, type=typ, l=l, w=len(sub), refall=refall, altall=altall, line=line))
    recs = []
    for t in range(threads):                 # svs.insert(svs.begin(), ...), caller.cpp:17-22
        recs = p_recs[t] + recs
    key = lambda r: (r["chrom"].encode(), r["s"])
    recs.sort(key=key)
    # clean_dups (caller.cpp:409-427)
    keep, last = [], None
    for r in recs:
        k = (r["chrom"], r["s"], r["refall"], r["altall"])
        if k != last:
            keep.append(r)
        last = k
    n_before = len(keep)
    recs = keep
    # filter_sv_chains (caller.cpp:430-475)
    if len(recs) >= 2:
        keep, prev, reset = [], recs[0], False
        for sv in recs[1:]:
            if reset:
                reset = False; prev = sv
                continue
            if sv["chrom"] == prev["chrom"] and sv["s"] - prev["e"] < 2 * sv["l"] and prev["type"] == sv["type"]:
                w_r = min(sv["w"], prev["w"]) / max(sv["w"], prev["w"])
                l_r = min(sv["l"], prev["l"]) / max(sv["l"], prev["l"])
                if sv["s"] - prev["s"] < 100 and w_r >= 0.9 and l_r >= float(np.float32(min_ratio)):
                    sim = fuzz_ratio(sv["refall"], prev["refall"]) if sv["type"] == "DEL" else fuzz_ratio(sv["altall"], prev["altall"])
                    if sim > 70:
                        keep.append(sv if sv["w"] > prev["w"] else prev)
                        reset = True
                        continue
            keep.append(prev)
            prev = sv
        keep.append(prev)
        recs = keep
        recs.sort(key=key)
    if return_stats:
        return [r["line"] for r in recs], dict(before_chain=n_before, after=len(recs))
    return [r["line"] for r in recs]
