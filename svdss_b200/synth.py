"""Synthetic HiFi-shaped workloads of SURVEY.md section 8(d) (fixed seeds, numpy, host side).

The reference ships no data (tests/README.md points at an external tarball), so every parity case and
every bench line runs on these generators.  nt6 codes: $=0 A=1 C=2 G=3 T=4 N=5 (ping_pong.hpp:46-52).
The device-side generator used at full scale lives in csrc/synth.cu and follows the same recipe.
"""
import numpy as np


def comp6(a):
    a = np.asarray(a, np.uint8)
    return np.where((a >= 1) & (a <= 4), 5 - a, a).astype(np.uint8)


def revcomp6(a):
    return comp6(a[::-1])


def make_reference(length=1_000_000, seed=1, n_repeats=20, n_nruns=5, nrun_len=200,
                   contigs=1):
    """i.i.d. uniform ACGT + planted exact repeats (1-5 kb) + N runs. Returns list of nt6 arrays."""
    rng = np.random.default_rng(seed)
    ref = rng.integers(1, 5, size=length, dtype=np.uint8)
    for _ in range(n_repeats):
        ln = int(rng.integers(1000, 5001)) if length >= 50_000 else int(rng.integers(20, max(21, length // 20)))
        if ln * 2 >= length:
            continue
        src = int(rng.integers(0, length - ln))
        dst = int(rng.integers(0, length - ln))
        seg = ref[src:src + ln].copy()
        if rng.random() < 0.5:
            seg = revcomp6(seg)
        ref[dst:dst + ln] = seg
    for _ in range(n_nruns):
        ln = min(nrun_len, max(1, length // 50))
        p = int(rng.integers(0, length - ln))
        ref[p:p + ln] = 5
    # split into contigs with GRCh38-like decreasing proportions
    if contigs <= 1:
        return [ref]
    w = np.linspace(2.0, 0.5, contigs)
    cuts = np.floor(np.cumsum(w / w.sum()) * length).astype(np.int64)
    cuts[-1] = length
    out, b = [], 0
    for c in cuts:
        out.append(ref[b:c].copy())
        b = int(c)
    return [c for c in out if len(c)]


def make_reads(contigs, n_reads=1000, seed=2, mean_len=15000, sd_len=2000, min_len=5000,
               max_len=25000, event_rate=0.5, raw_hifi=False):
    """Smoothed-shaped reads: exact reference copies (either strand) with Poisson(event_rate) planted
    events per read (INS of i.i.d. bases / DEL, U[30,500] bp; or a 100-2000 bp random soft-clip
    tail).  raw_hifi adds 0.1% substitutions + 0.05% 1-bp indels.  Returns list of nt6 arrays."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for c in contigs], np.int64)
    reads = []
    for _ in range(n_reads):
        L = int(np.clip(rng.normal(mean_len, sd_len), min_len, max_len))
        ci = int(rng.choice(len(contigs), p=lens / lens.sum()))
        c = contigs[ci]
        L = min(L, len(c))
        st = int(rng.integers(0, len(c) - L + 1))
        r = c[st:st + L].copy()
        if rng.random() < 0.5:
            r = revcomp6(r)
        for _e in range(int(rng.poisson(event_rate))):
            kind = int(rng.integers(0, 3))
            if kind == 0 and len(r) > 700:      # INS
                p = int(rng.integers(100, len(r) - 100))
                ins = rng.integers(1, 5, size=int(rng.integers(30, 501)), dtype=np.uint8)
                r = np.concatenate([r[:p], ins, r[p:]])
            elif kind == 1 and len(r) > 1400:   # DEL
                d = int(rng.integers(30, 501))
                p = int(rng.integers(100, len(r) - d - 100))
                r = np.concatenate([r[:p], r[p + d:]])
            else:                               # soft-clip tail (random sequence)
                t = rng.integers(1, 5, size=int(rng.integers(100, 2001)), dtype=np.uint8)
                r = np.concatenate([r, t]) if rng.random() < 0.5 else np.concatenate([t, r])
        if raw_hifi:
            n_sub = rng.binomial(len(r), 0.001)
            pos = rng.integers(0, len(r), size=n_sub)
            r[pos] = ((r[pos] - 1 + rng.integers(1, 4, size=n_sub)) % 4 + 1).astype(np.uint8)
            n_id = rng.binomial(len(r), 0.0005)
            for p in sorted(rng.integers(1, len(r) - 1, size=n_id).tolist(), reverse=True):
                if rng.random() < 0.5:
                    r = np.delete(r, p)
                else:
                    r = np.insert(r, p, rng.integers(1, 5, dtype=np.uint8))
        reads.append(np.ascontiguousarray(r, np.uint8))
    return reads


def concat(seqs):
    offs = np.zeros(len(seqs) + 1, np.int64)
    if len(seqs):
        offs[1:] = np.cumsum([len(s) for s in seqs])
    cat = (np.ascontiguousarray(np.concatenate(seqs), np.uint8) if len(seqs) and offs[-1]
           else np.zeros(0, np.uint8))
    return cat, offs


# ---------------------------------------------------------------------------------------------
# Scale generator (config 2/3): reads are described as *segments* so that the bases can be
# materialised on the device from a device-resident reference without a 15 GB host round trip.
# A segment is (kind, start, len, rev): kind 0 = reference slice [start, start+len) of the
# concatenated contigs (reverse-complemented when rev), kind 1 = `len` hashed random bases.
# ---------------------------------------------------------------------------------------------
_M1 = np.int64(-7046029254386353131)   # 0x9E3779B97F4A7C15 as int64
_M2 = np.int64(-4658895280553007687)   # 0xBF58476D1CE4E5B9 as int64


def make_read_segments(contig_offs, n_reads, seed=4, mean_len=15000, sd_len=2000, min_len=5000,
                       max_len=25000, event_rate=0.5, max_events=4):
    """Vectorised smoothed-shaped read recipe. contig_offs: int64[m+1] offsets of the contigs in the
    concatenated reference. Returns dict of numpy arrays describing all segments + read offsets."""
    rng = np.random.default_rng(seed)
    contig_offs = np.asarray(contig_offs, np.int64)
    clen = np.diff(contig_offs)
    L = np.clip(rng.normal(mean_len, sd_len, n_reads), min_len, max_len).astype(np.int64)
    ci = rng.choice(len(clen), size=n_reads, p=clen / clen.sum())
    L = np.minimum(L, clen[ci])
    st = contig_offs[ci] + (rng.random(n_reads) * (clen[ci] - L + 1)).astype(np.int64)
    rev = rng.random(n_reads) < 0.5
    nev = np.minimum(rng.poisson(event_rate, n_reads), max_events)
    E = max_events
    kind = rng.integers(0, 3, size=(n_reads, E))            # 0 INS, 1 DEL, 2 clip
    evlen = rng.integers(30, 501, size=(n_reads, E))
    cliplen = rng.integers(100, 2001, size=(n_reads, E))
    pos = np.sort((rng.random((n_reads, E)) * np.maximum(L - 200, 1)[:, None]).astype(np.int64) + 100, axis=1)
    active = np.arange(E)[None, :] < nev[:, None]
    # piece list per read: [clip_front?] then for each event e: template[cur:pos_e], (INS random | DEL skip)
    seg_kind, seg_start, seg_len, seg_rev, seg_read = [], [], [], [], []

    def add(mask, k, s, ln, rv):
        idx = np.nonzero(mask & (ln > 0))[0]
        seg_kind.append(np.full(len(idx), k, np.int8)); seg_start.append(s[idx]); seg_len.append(ln[idx])
        seg_rev.append(rv[idx]); seg_read.append(idx)

    order_key = []  # (read, slot) ordering: slot increases along the read
    slot = 0
    is_clip = active & (kind == 2)
    front = is_clip & (rng.random((n_reads, E)) < 0.5)
    back = is_clip & ~front
    zeros = np.zeros(n_reads, np.int64)
    for e in range(E):   # front clips first
        add(front[:, e], 1, zeros, cliplen[:, e].astype(np.int64), np.zeros(n_reads, bool)); order_key.append(slot); slot += 1
    cur = np.zeros(n_reads, np.int64)
    for e in range(E):
        ev = active[:, e] & (kind[:, e] != 2)
        p = np.clip(pos[:, e], cur, L)
        # template piece [cur, p)
        a, b = cur, np.where(ev, p, cur)
        ln = b - a
        start = np.where(rev, st + L - b, st + a)
        add(ev, 0, start, ln, rev); order_key.append(slot); slot += 1
        ins = ev & (kind[:, e] == 0)
        add(ins, 1, zeros, evlen[:, e].astype(np.int64), np.zeros(n_reads, bool)); order_key.append(slot); slot += 1
        dele = ev & (kind[:, e] == 1)
        cur = np.where(ev, p, cur)
        cur = np.where(dele, np.minimum(cur + evlen[:, e], L), cur)
    ln = L - cur
    start = np.where(rev, st, st + cur)
    add(np.ones(n_reads, bool), 0, start, ln, rev); order_key.append(slot); slot += 1
    for e in range(E):
        add(back[:, e], 1, zeros, cliplen[:, e].astype(np.int64), np.zeros(n_reads, bool)); order_key.append(slot); slot += 1
    slots = np.concatenate([np.full(len(r), k, np.int64) for r, k in zip(seg_read, order_key)])
    seg_read = np.concatenate(seg_read); seg_kind = np.concatenate(seg_kind)
    seg_start = np.concatenate(seg_start); seg_len = np.concatenate(seg_len); seg_rev = np.concatenate(seg_rev)
    o = np.lexsort((slots, seg_read))
    seg_read, seg_kind, seg_start, seg_len, seg_rev = seg_read[o], seg_kind[o], seg_start[o], seg_len[o], seg_rev[o]
    seg_out = np.zeros(len(seg_len) + 1, np.int64)
    seg_out[1:] = np.cumsum(seg_len)
    read_len = np.bincount(seg_read, weights=seg_len, minlength=n_reads).astype(np.int64)
    read_offs = np.zeros(n_reads + 1, np.int64)
    read_offs[1:] = np.cumsum(read_len)
    return {"kind": seg_kind, "start": seg_start, "len": seg_len, "rev": seg_rev, "read": seg_read,
            "out": seg_out, "read_offs": read_offs, "seed": seed, "n_events": nev}


def _hash_bases_np(gidx, seed):
    h = (gidx.astype(np.int64) + np.int64(seed)) * _M1
    h = h ^ ((h >> 29) & np.int64(0x7FFFFFFFF))
    h = h * _M2
    return (((h >> 33) & 3) + 1).astype(np.uint8)


def materialize_segments_numpy(ref_cat, segs):
    """Host materialiser (tests)."""
    total = int(segs["out"][-1])
    out = np.empty(total, np.uint8)
    seg_id = np.repeat(np.arange(len(segs["len"])), segs["len"])
    gidx = np.arange(total, dtype=np.int64)
    w = gidx - segs["out"][seg_id]
    rv = segs["rev"][seg_id]
    src = np.where(rv, segs["start"][seg_id] + segs["len"][seg_id] - 1 - w, segs["start"][seg_id] + w)
    is_ref = segs["kind"][seg_id] == 0
    b = ref_cat[np.where(is_ref, src, 0)]
    b = np.where(rv, comp6(b), b)
    hidx = gidx
    if "key" in segs:      # kind 2: bases hashed from key + offset in the segment (the same allele in every read)
        hidx = np.where(segs["kind"][seg_id] == 2, segs["key"][seg_id] + w, gidx)
    out[:] = np.where(is_ref, b, _hash_bases_np(hidx, segs["seed"]))
    return out


def materialize_segments_torch(ref_cat_t, segs, out_t=None, chunk_segs=40_000):
    """Device materialiser: ref_cat_t uint8 CUDA tensor of the concatenated contigs. Returns a uint8
    CUDA tensor with 64 spare bytes at the end (the search kernel's read-window padding)."""
    import torch
    dev = ref_cat_t.device
    total = int(segs["out"][-1])
    out = out_t if out_t is not None else torch.zeros(total + 64, dtype=torch.uint8, device=dev)
    S = len(segs["len"])
    M1 = torch.tensor(int(_M1), dtype=torch.int64, device=dev)
    M2 = torch.tensor(int(_M2), dtype=torch.int64, device=dev)
    for s0 in range(0, S, chunk_segs):
        s1 = min(S, s0 + chunk_segs)
        ln = torch.from_numpy(segs["len"][s0:s1]).to(dev)
        o0 = int(segs["out"][s0]); o1 = int(segs["out"][s1])
        if o1 == o0:
            continue
        seg_id = torch.repeat_interleave(torch.arange(s1 - s0, device=dev), ln)
        gidx = torch.arange(o0, o1, dtype=torch.int64, device=dev)
        so = torch.from_numpy(segs["out"][s0:s1]).to(dev)
        st = torch.from_numpy(segs["start"][s0:s1]).to(dev)
        rv = torch.from_numpy(segs["rev"][s0:s1]).to(dev)
        kd = torch.from_numpy(segs["kind"][s0:s1]).to(dev)
        w = gidx - so[seg_id]
        rvv = rv[seg_id]
        src = torch.where(rvv, st[seg_id] + ln[seg_id] - 1 - w, st[seg_id] + w)
        is_ref = kd[seg_id] == 0
        b = ref_cat_t[torch.where(is_ref, src, torch.zeros_like(src))]
        cb = torch.where((b >= 1) & (b <= 4), 5 - b, b)
        b = torch.where(rvv, cb, b)
        hidx = gidx
        if "key" in segs:
            ky = torch.from_numpy(segs["key"][s0:s1]).to(dev)
            hidx = torch.where(kd[seg_id] == 2, ky[seg_id] + w, gidx)
        h = (hidx + int(segs["seed"])) * M1
        h = h ^ ((h >> 29) & 0x7FFFFFFFF)
        h = h * M2
        rb = (((h >> 33) & 3) + 1).to(torch.uint8)
        out[o0:o1] = torch.where(is_ref, b, rb)
    return out


# ---------------------------------------------------------------------------------------------
# SV-carrying sample (config 3 shape): a diploid sample = reference + planted INS/DEL catalogue,
# reads drawn from the two haplotypes with their true alignment (pos + CIGAR) to the reference, as
# a smoothed BAM would hold them (forward-strand sequence, M blocks identical to the reference).
# ---------------------------------------------------------------------------------------------
def make_sv_catalogue(contigs, n_svs, seed=5, min_len=50, max_len=5000, margin=2000, spacing=3000):
    """Planted het/hom INS/DEL: list of dicts (contig, pos, type, len, seq, gt) sorted by (contig, pos).
    `pos` is the 0-based reference base AFTER which the event happens (VCF-style anchor = pos)."""
    rng = np.random.default_rng(seed)
    lens = np.array([len(c) for c in contigs], np.int64)
    out, taken = [], [[] for _ in contigs]
    tries = 0
    while len(out) < n_svs and tries < 100 * n_svs:
        tries += 1
        ci = int(rng.choice(len(contigs), p=lens / lens.sum()))
        ln = int(np.exp(rng.uniform(np.log(min_len), np.log(max_len))))
        if lens[ci] < 2 * margin + ln + 10:
            continue
        pos = int(rng.integers(margin, lens[ci] - margin - ln))
        if any(abs(pos - p) < spacing + ln + l2 for p, l2 in taken[ci]):
            continue
        typ = "INS" if rng.random() < 0.5 else "DEL"
        if typ == "DEL" and (contigs[ci][pos:pos + ln + 2] == 5).any():
            continue
        seq = rng.integers(1, 5, size=ln, dtype=np.uint8) if typ == "INS" else None
        gt = (1, 1) if rng.random() < 0.4 else ((1, 0) if rng.random() < 0.5 else (0, 1))
        taken[ci].append((pos, ln))
        out.append(dict(contig=ci, pos=pos, type=typ, len=ln, seq=seq, gt=gt))
    out.sort(key=lambda s: (s["contig"], s["pos"]))
    return out


def _add_noise(rng, pieces, cigar, sub_rate, indel_rate):
    """sequencing noise on the M blocks of one alignment: substitutions in place, 1-bp insertions and
    deletions that split the block (first / last 20 bases of a block stay clean)"""
    out_p, out_c = [], []
    pi = 0
    for ln, op in cigar:
        if op == "D":
            out_c.append((ln, op))
            continue
        blk = pieces[pi].copy(); pi += 1
        if op != "M" or ln < 60:
            out_p.append(blk); out_c.append((ln, op))
            continue
        subs = np.nonzero(rng.random(ln) < sub_rate)[0]
        blk[subs] = (blk[subs] % 4) + 1
        cuts = sorted(set((np.nonzero(rng.random(ln - 40) < indel_rate)[0] + 20).tolist()))
        last = 0
        for cpos in cuts:
            if cpos - last < 2:
                continue
            out_p.append(blk[last:cpos]); out_c.append((cpos - last, "M"))
            if rng.random() < 0.5:
                out_p.append(rng.integers(1, 5, size=1, dtype=np.uint8)); out_c.append((1, "I")); last = cpos
            else:
                out_c.append((1, "D")); last = cpos + 1
        out_p.append(blk[last:]); out_c.append((ln - last, "M"))
    return out_p, out_c


def make_sample_alignments(contigs, catalogue, coverage=10, seed=6, mean_len=15000, sd_len=2000, min_len=1000,
                           max_len=25000, tag_hp=True, names=None, clip_rate=0.1, sub_rate=0.0, indel_rate=0.0):
    """Reads of a diploid sample with true alignments. Returns a list of dicts: qname, tid, pos, cigar
    [(len, op)], seq (nt6 array, forward strand), hp (1/2), has_event (bool -> XF:i:0 else XF:i:2).
    Sorted by (tid, pos) like a coordinate-sorted BAM.  sub_rate / indel_rate > 0 make the reads
    raw-HiFi-shaped (substitutions inside M blocks, 1-bp I/D splitting them): the input of `smooth`."""
    rng = np.random.default_rng(seed)
    recs = []
    rid = 0
    for ci, c in enumerate(contigs):
        n_reads = max(1, int(len(c) * coverage / mean_len))
        for hap in (0, 1):
            evs = [s for s in catalogue if s["contig"] == ci and s["gt"][hap]]
            for _ in range((n_reads + (1 - hap)) // 2):
                L = int(np.clip(rng.normal(mean_len, sd_len), min_len, max_len))
                L = min(L, len(c))
                a = int(rng.integers(0, len(c) - L + 1))
                b = a + L
                pieces, cigar, ref = [], [], a
                has_event = False
                for s in evs:
                    p = s["pos"]
                    # the event must sit well inside the read: both flanks >= 150 reference bases
                    lo = p + 1
                    hi = p + 1 + (s["len"] if s["type"] == "DEL" else 0)
                    if lo - a < 150 or b - hi < 150 or lo < ref:
                        continue
                    pieces.append(c[ref:lo]); cigar.append((lo - ref, "M"))
                    if s["type"] == "INS":
                        pieces.append(s["seq"]); cigar.append((s["len"], "I"))
                    else:
                        cigar.append((s["len"], "D"))
                    ref = hi
                    has_event = True
                pieces.append(c[ref:b]); cigar.append((b - ref, "M"))
                if rng.random() < clip_rate:      # soft-clipped random tail (either end)
                    t = rng.integers(1, 5, size=int(rng.integers(100, 2001)), dtype=np.uint8)
                    if rng.random() < 0.5:
                        pieces.append(t); cigar.append((len(t), "S"))
                    else:
                        pieces.insert(0, t); cigar.insert(0, (len(t), "S"))
                    has_event = True
                if sub_rate > 0 or indel_rate > 0:
                    pieces, cigar = _add_noise(rng, pieces, cigar, sub_rate, indel_rate)
                seq = np.ascontiguousarray(np.concatenate(pieces), np.uint8)
                recs.append(dict(qname=(names[rid] if names else "read_%06d" % rid), tid=ci, pos=a, cigar=cigar, seq=seq,
                                 hp=hap + 1 if tag_hp else 0, has_event=has_event))
                rid += 1
    recs.sort(key=lambda r: (r["tid"], r["pos"]))
    return recs


# ---------------------------------------------------------------------------------------------
# `call` stage workloads: config 4 = clusters of 20-60 reads x 200-2000 bp with 0.1 % noise,
# half the reads carrying one planted INS/DEL; config 5 = consensus x window pairs, target length
# log-uniform, one planted INS/DEL of 50-1000 bp + 0.2 % noise.  Sequences are codes 0..3.
def gen_clusters(rng, n, lo=200, hi=2000):
    out = []
    for _ in range(n):
        tlen = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        nr = int(rng.integers(20, 61))
        tpl = rng.integers(0, 4, size=tlen).astype(np.uint8)
        reads = []
        for _r in range(nr):
            r = tpl.copy()
            m = rng.random(tlen) < 0.001
            r[m] = (r[m] + rng.integers(1, 4, size=int(m.sum()))) % 4
            if rng.random() < 0.5:
                L = int(rng.integers(1, max(2, int(tlen * 0.03))))
                p = int(rng.integers(1, max(2, tlen - L - 1)))
                r = np.concatenate([r[:p], rng.integers(0, 4, size=L).astype(np.uint8), r[p:]]) if rng.random() < 0.5 \
                    else np.concatenate([r[:p], r[p + L:]])
            reads.append(r)
        out.append(reads)
    return out


def gen_pairs(rng, n, lo=100, hi=10000):
    out = []
    for _ in range(n):
        tl = int(np.exp(rng.uniform(np.log(lo), np.log(hi))))
        t = rng.integers(0, 4, size=tl).astype(np.uint8)
        L = int(min(rng.integers(50, 1001), max(1, tl // 2)))
        p = int(rng.integers(1, max(2, tl - L - 1)))
        q = np.concatenate([t[:p], rng.integers(0, 4, size=L).astype(np.uint8), t[p:]]) if rng.random() < 0.5 \
            else np.concatenate([t[:p], t[p + L:]])
        m = rng.random(len(q)) < 0.002
        q[m] = (q[m] + rng.integers(1, 4, size=int(m.sum()))) % 4
        out.append((q, t))
    return out


# ---------------------------------------------------------------------------------------------
# Config 3 at scale: a 30x coordinate-sorted smoothed BAM over the config-2 reference with a
# planted SV catalogue, described as arrays (no per-read Python): what `search` + `call` see of it.
# ---------------------------------------------------------------------------------------------
def make_sv_catalogue_arrays(contig_offs, n_svs=20000, seed=5, min_len=50, max_len=5000):
    """Planted het/hom INS/DEL over the whole reference, one per equal-sized cell of the concatenated contigs
    (inside the cell's central half, so two SVs never share a read), cells touching a contig end dropped.
    Returns dict of arrays sorted by position: gpos (global 0-based base AFTER which the event happens),
    contig, pos (contig-relative), is_del, len, hapmask (bit 0 / 1 = carried by haplotype 1 / 2), id."""
    rng = np.random.default_rng(seed)
    contig_offs = np.asarray(contig_offs, np.int64)
    total = int(contig_offs[-1])
    cell = total // n_svs
    g = (np.arange(n_svs, dtype=np.int64) * cell + cell // 4 + (rng.random(n_svs) * (cell // 2)).astype(np.int64))
    ln = np.exp(rng.uniform(np.log(min_len), np.log(max_len), n_svs)).astype(np.int64)
    is_del = rng.random(n_svs) < 0.5
    u = rng.random(n_svs)
    hapmask = np.where(u < 0.4, 3, np.where(u < 0.7, 1, 2)).astype(np.int8)
    ci = np.searchsorted(contig_offs, g, side="right") - 1
    keep = (g - contig_offs[ci] > 30000) & (contig_offs[ci + 1] - g - ln > 30000)
    idx = np.nonzero(keep)[0]
    return dict(gpos=g[idx], contig=ci[idx].astype(np.int32), pos=(g - contig_offs[ci])[idx], is_del=is_del[idx], len=ln[idx],
                hapmask=hapmask[idx], id=idx.astype(np.int64))


def make_sample_region(contig_offs, cat, n_reads, g_start, coverage=30.0, seed=6, mean_len=15000, sd_len=2000, min_len=5000,
                       max_len=25000, clip_rate=0.05, tag_hp=True):
    """`n_reads` records of a coordinate-sorted `coverage`x smoothed BAM starting at global position g_start (a slice of
    config 3: rank r of N takes the r-th slice).  Reads are forward-strand copies of one haplotype of the sample (the
    reference with the catalogue's SVs of that haplotype applied); a read carries an SV when the event lies >= 150
    reference bases inside it; `clip_rate` of the reads get a 100-2000 bp random soft-clipped tail.  XF = 0 for reads
    with an event or a clip (what the Smoother tags as worth searching), 2 otherwise (smoother.cpp:213-231).
    Returns dict: tid, pos, hp, xf, l_qseq, cigar_offs, cigar (BAM encoding), and `segs` -- the segment recipe of
    materialize_segments_* for the XF == 0 reads only (`searched` = their record indices): kind 0 reference slice,
    kind 1 bases hashed from the output position (clips), kind 2 bases hashed from key + offset (the inserted
    allele: identical in every read that carries it)."""
    rng = np.random.default_rng(seed)
    contig_offs = np.asarray(contig_offs, np.int64)
    total = int(contig_offs[-1])
    span = int(n_reads * mean_len / coverage)
    L = np.clip(rng.normal(mean_len, sd_len, n_reads), min_len, max_len).astype(np.int64)
    a = np.sort((g_start + (rng.random(n_reads) * span).astype(np.int64)) % max(1, total - max_len - 1))
    ci = np.searchsorted(contig_offs, a, side="right") - 1
    L = np.minimum(L, contig_offs[ci + 1] - a)          # never across a contig end
    short = L < min_len
    a[short] = contig_offs[ci[short] + 1] - min_len
    L[short] = min_len
    o = np.lexsort((a, ci))
    a, L, ci = a[o], L[o], ci[o]
    b = a + L
    hap = (rng.random(n_reads) < 0.5).astype(np.int8)   # 0 / 1
    # the (at most one) SV inside each read: first catalogue entry with gpos + 1 - a >= 150
    k = np.searchsorted(cat["gpos"], a + 149, side="left")
    k = np.minimum(k, len(cat["gpos"]) - 1)
    lo = cat["gpos"][k] + 1
    hi = lo + np.where(cat["is_del"][k], cat["len"][k], 0)
    carries = (lo - a >= 150) & (b - hi >= 150) & (((cat["hapmask"][k] >> hap) & 1) == 1)
    is_ins = carries & ~cat["is_del"][k]
    is_del = carries & cat["is_del"][k]
    evlen = np.where(carries, cat["len"][k], 0)
    clip = rng.random(n_reads) < clip_rate
    clip_front = clip & (rng.random(n_reads) < 0.5)
    clip_back = clip & ~clip_front
    cliplen = np.where(clip, rng.integers(100, 2001, n_reads), 0).astype(np.int64)
    xf = np.where(carries | clip, 0, 2).astype(np.int32)
    # CIGAR: [S] M [I|D M] [S]
    m1 = np.where(carries, lo - a, L)
    m2 = np.where(carries, b - hi, 0)
    n_ops = 1 + clip.astype(np.int64) + 2 * carries.astype(np.int64)
    cigar_offs = np.zeros(n_reads + 1, np.int64)
    cigar_offs[1:] = np.cumsum(n_ops)
    cigar = np.zeros(int(cigar_offs[-1]), np.uint32)
    p = cigar_offs[:-1].copy()
    f = np.nonzero(clip_front)[0]; cigar[p[f]] = (cliplen[f] << 4) | 4; p[f] += 1
    cigar[p] = (m1 << 4) | 0; p += 1
    f = np.nonzero(carries)[0]
    cigar[p[f]] = (evlen[f] << 4) | np.where(is_del[f], 2, 1); p[f] += 1
    cigar[p[f]] = (m2[f] << 4) | 0; p[f] += 1
    f = np.nonzero(clip_back)[0]; cigar[p[f]] = (cliplen[f] << 4) | 4
    l_qseq = (m1 + m2 + np.where(is_ins, evlen, 0) + cliplen).astype(np.int64)
    # segments of the searched (XF == 0) reads, in read order
    s_idx = np.nonzero(xf == 0)[0]
    ns = len(s_idx)
    slot_kind = np.stack([np.where(clip_front[s_idx], 1, -1), np.zeros(ns, np.int64), np.where(is_ins[s_idx], 2, -1),
                          np.where(carries[s_idx], 0, -1), np.where(clip_back[s_idx], 1, -1)], axis=1)
    slot_len = np.stack([cliplen[s_idx], m1[s_idx], evlen[s_idx], m2[s_idx], cliplen[s_idx]], axis=1)
    slot_start = np.stack([np.zeros(ns, np.int64), a[s_idx], np.zeros(ns, np.int64), hi[s_idx], np.zeros(ns, np.int64)], axis=1)
    slot_key = np.zeros((ns, 5), np.int64)
    slot_key[:, 2] = (cat["id"][k[s_idx]] + 1) * 8192
    use = (slot_kind >= 0) & (slot_len > 0)
    seg_read = np.repeat(np.arange(ns), 5).reshape(ns, 5)[use]
    seg_kind = slot_kind[use].astype(np.int8)
    seg_len = slot_len[use]
    seg_out = np.zeros(len(seg_len) + 1, np.int64)
    seg_out[1:] = np.cumsum(seg_len)
    read_offs = np.zeros(ns + 1, np.int64)
    read_offs[1:] = np.cumsum(l_qseq[s_idx])
    segs = {"kind": seg_kind, "start": slot_start[use], "len": seg_len, "rev": np.zeros(len(seg_len), bool), "read": seg_read,
            "key": slot_key[use], "out": seg_out, "read_offs": read_offs, "seed": seed}
    return dict(tid=ci.astype(np.int32), pos=(a - contig_offs[ci]).astype(np.int32), hp=(hap + 1 if tag_hp else 0 * hap).astype(np.int32),
                xf=xf, l_qseq=l_qseq.astype(np.int32), cigar_offs=cigar_offs, cigar=cigar, searched=s_idx, segs=segs,
                sv_of_read=np.where(carries, k, -1), span=span, g_start=int(g_start))


def _hash_codes(idx, seed):
    """codes 0..3 hashed from 64-bit indices"""
    return (_hash_bases_np(np.asarray(idx, np.int64), seed) - 1).astype(np.uint8)


def gen_clusters_fast(n, seed=6, lo=200, hi=2000):
    """Config 4 at scale, vectorised: the recipe of gen_clusters (template of log-uniform length, 20-60 reads per cluster,
    0.1 % substitutions, half the reads with one INS or DEL of up to 3 % of the template) as arrays.
    Returns (codes uint8 concatenated, seq_offs int64[n_reads + 1], cluster_offs int64[n + 1])."""
    rng = np.random.default_rng(seed)
    tlen = np.exp(rng.uniform(np.log(lo), np.log(hi), n)).astype(np.int64)
    nr = rng.integers(20, 61, n)
    toff = np.zeros(n + 1, np.int64)
    toff[1:] = np.cumsum(tlen)
    pool = rng.integers(0, 4, int(toff[-1]), dtype=np.uint8)
    co = np.zeros(n + 1, np.int64)
    co[1:] = np.cumsum(nr)
    R = int(co[-1])
    cl = np.repeat(np.arange(n), nr)
    tl = tlen[cl]
    ev = rng.random(R) < 0.5
    is_ins = ev & (rng.random(R) < 0.5)
    is_del = ev & ~is_ins
    L = np.where(ev, (rng.random(R) * np.maximum(1, (tl * 0.03).astype(np.int64) - 1)).astype(np.int64) + 1, 0)
    p = (rng.random(R) * np.maximum(1, tl - L - 2)).astype(np.int64) + 1
    L = np.minimum(L, np.maximum(tl - p - 1, 0))
    rl = tl + np.where(is_ins, L, 0) - np.where(is_del, L, 0)
    so = np.zeros(R + 1, np.int64)
    so[1:] = np.cumsum(rl)
    tot = int(so[-1])
    rid = np.repeat(np.arange(R), rl)
    w = np.arange(tot, dtype=np.int64) - so[rid]                    # offset in the read
    pr, Lr = p[rid], L[rid]
    in_ins = is_ins[rid] & (w >= pr) & (w < pr + Lr)
    src = np.where(w < pr, w, np.where(is_ins[rid], w - Lr, np.where(is_del[rid], w + Lr, w)))
    out = pool[toff[cl[rid]] + np.where(in_ins, 0, src)]
    gi = np.arange(tot, dtype=np.int64)
    out = np.where(in_ins, _hash_codes(gi, seed + 1), out)
    sub = rng.random(tot) < 0.001
    out = np.where(sub, (out + 1 + _hash_codes(gi, seed + 2) % 3) % 4, out).astype(np.uint8)
    return np.ascontiguousarray(out), so, co


def gen_pairs_fast(n, seed=7, lo=100, hi=10000):
    """Config 5 at scale, vectorised: target of log-uniform length, query = target with one planted INS or DEL of 50-1000 bp
    (at most half the target) + 0.2 % substitutions.  Returns (q codes, q_offs, t codes, t_offs)."""
    rng = np.random.default_rng(seed)
    tl = np.exp(rng.uniform(np.log(lo), np.log(hi), n)).astype(np.int64)
    to = np.zeros(n + 1, np.int64)
    to[1:] = np.cumsum(tl)
    t = rng.integers(0, 4, int(to[-1]), dtype=np.uint8)
    L = np.minimum(rng.integers(50, 1001, n), np.maximum(1, tl // 2))
    p = (rng.random(n) * np.maximum(1, tl - L - 2)).astype(np.int64) + 1
    is_ins = rng.random(n) < 0.5
    ql = tl + np.where(is_ins, L, -L)
    qo = np.zeros(n + 1, np.int64)
    qo[1:] = np.cumsum(ql)
    tot = int(qo[-1])
    rid = np.repeat(np.arange(n), ql)
    w = np.arange(tot, dtype=np.int64) - qo[rid]
    pr, Lr, ins = p[rid], L[rid], is_ins[rid]
    in_ins = ins & (w >= pr) & (w < pr + Lr)
    src = np.where(w < pr, w, np.where(ins, w - Lr, w + Lr))
    q = t[to[rid] + np.where(in_ins, 0, src)]
    gi = np.arange(tot, dtype=np.int64)
    q = np.where(in_ins, _hash_codes(gi, seed + 1), q)
    sub = rng.random(tot) < 0.002
    q = np.where(sub, (q + 1 + _hash_codes(gi, seed + 2) % 3) % 4, q).astype(np.uint8)
    return np.ascontiguousarray(q), qo, t, to
