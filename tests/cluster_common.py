"""Shared helpers of the Clusterer tests: the records Clusterer::run keeps of a world (tests/sv_world.py) as an
svb_alns_t, and the Python model's clusters in the shape svb_clusters_t has."""
import numpy as np

import cluster_model
from svdss_b200 import capi


def kept_records(world, min_mapq=20):
    """primary records with mapq >= min_mapq, BAM order (clusterer.cpp:116-122)"""
    return [r for r in world["records"] if cluster_model.primary(r) and r.get("mapq", 60) >= min_mapq]


def aln_batch(world, min_mapq=20, sfs_by_read=None):
    recs = kept_records(world, min_mapq)
    sfs_by_read = world["sfs_by_read"] if sfs_by_read is None else sfs_by_read
    sfs = {i: [(q, l) for q, l, _ in sfs_by_read[r["qname"]]] for i, r in enumerate(recs) if r["qname"] in sfs_by_read}
    return recs, capi.AlnBatch.from_records(recs, sfs)


def ref_of(world):
    return capi.RefSeqs.from_strings(world["names"], [world["ref_seqs"][n] for n in world["names"]])


def as_model_shape(res, recs, names):
    """svb_clusters_t -> the dicts cluster_model.run returns (only placed clusters appear there)"""
    out = []
    for c in range(res.n):
        if not res.placed[c]:
            continue
        a, b = int(res.sub_offs[c]), int(res.sub_offs[c + 1])
        sub = []
        for k in range(a, b):
            r = recs[int(res.sub_aln[k])]
            sub.append((r["qname"], r["seq"][int(res.sub_qs[k]):int(res.sub_qe[k]) + 1], int(res.sub_hp[k])))
        d = dict(chrom=names[int(res.tid[c])], s=int(res.s[c]), e=int(res.e[c]), subreads=sub, cov=None, reads=[])
        if len(sub) >= 2 or res.rvec_offs[c + 1] > res.rvec_offs[c]:
            pass
        out.append(d)
    return out


def compare(res, exp, recs, names, min_cluster_weight=2):
    got = as_model_shape(res, recs, names)
    assert len(got) == len(exp)
    k = 0
    for c in range(res.n):
        if not res.placed[c]:
            continue
        g, e = got[k], exp[k]
        k += 1
        assert (g["chrom"], g["s"], g["e"]) == (e["chrom"], e["s"], e["e"])
        assert g["subreads"] == [tuple(x) for x in e["subreads"]]
        rv = res.rvec[int(res.rvec_offs[c]):int(res.rvec_offs[c + 1])]
        if len(e["subreads"]) >= min_cluster_weight:
            assert (int(res.cov0[c]), int(res.cov1[c]), int(res.cov2[c])) == tuple(e["cov"])
            assert [(int(x) & 1, int(x) >> 1) for x in rv] == [tuple(x) for x in e["reads"]]
        else:
            assert e["cov"] is None and len(rv) == 0 and (int(res.cov0[c]), int(res.cov1[c]), int(res.cov2[c])) == (0, 0, 0)
