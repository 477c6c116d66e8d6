"""Python restatement of the `call` compute core for tests: split_cluster (untagged branch,
caller.cpp:100-150 + :78-97), run_poa -> ksw2 -> CIGAR walk -> SV (caller.cpp:311-406), SV::operator<<
(sv.cpp:53-80).  POA and ksw2 come from the oracle."""
import numpy as np

import oracle


def get_len(sub):          # clusterer.hpp:103-111, unsigned integer mean
    return sum(len(s[1]) for s in sub) // len(sub)


def split_cluster_by_len(subreads, min_ratio=np.float32(0.97)):
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


def split_cluster(subreads):
    groups = split_cluster_by_len(subreads)
    i1 = i2 = -1
    v1 = v2 = 0
    for i, g in enumerate(groups):
        if len(g) > v1:
            v2, i2, v1, i1 = v1, i1, len(g), i
        elif len(g) > v2:
            v2, i2 = len(g), i
    return [groups[i] for i in (i1, i2) if i != -1]


def call_vcf_lines(ref, clusters, min_cluster_weight=2, min_sv_length=25):
    """ref: dict chrom -> upper-case str; clusters: list of (chrom, s, e, [(name, seq)]) with 0-based
    inclusive s,e. Returns the VCF record lines sorted by (chrom, s) like the shell prints them."""
    enc = lambda s: oracle.CHAR26[np.frombuffer(s.encode(), np.uint8)]
    recs = []
    for chrom, s, e, subreads in clusters:
        if len(subreads) < min_cluster_weight:
            continue
        n = len(subreads)
        for sub in split_cluster(subreads):
            cons_codes = oracle.poa_consensus([enc(sq) for _, sq in sub], band=True)
            cons = "".join("ACGTN"[c] for c in cons_codes)
            window = ref[chrom][s:e + 1]
            score, cig = oracle.ksw_extd2(enc(cons), enc(window))
            cigar_str = "".join("%d%s" % (l, op) for l, op in cig)
            rpos, cpos, nv, svs = s, 0, 0, []
            reads = ",".join(nm for nm, _ in sub)
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
                        "CIGAR=%s;RVEC=;READS=%s" % (typ, -l if typ == "DEL" else l, end, len(sub), n, n, -1, -1, score, nv,
                                                     cigar_str, reads))
                recs.append((chrom, pos, "%s\t%d\t%s\t%s\t%s\t.\tPASS\t%s\tGT:GQ\t0/1:100" % (chrom, pos, idx, refall, altall, info)))
    recs.sort(key=lambda r: (r[0], r[1]))
    return [r[2] for r in recs]
