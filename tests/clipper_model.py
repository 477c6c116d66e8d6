"""Literal Python transcription of the reference Clipper (clipper.cpp:3-215) and of the collection
loop in Caller::run (caller.cpp:37-55), for tests of svdss_b200/host/clipper.hpp.

Clips are tuples (name, chrom, p, l, starting, w).  One thing cannot be transcribed: combine()
walks a std::unordered_map<uint, ...>, whose iteration order belongs to the C++ library.  The model
therefore takes the order of the combined breakpoints of a side as an argument (the shell prints it
with --verbose), checks that it is a permutation of what combine() must produce, and transcribes
everything downstream of it.  Arithmetic on positions is reduced mod 2**32 where the reference
computes in `uint`."""

U32 = 1 << 32


def remove_duplicates(clips):                              # clipper.cpp:5-15
    seen, out = set(), []
    for c in clips:
        if c[0] not in seen:
            seen.add(c[0])
            out.append(c)
    return out


def combine_set(clips, chromosomes):                       # clipper.cpp:17-52, as per-slot lists of unordered groups
    """Returns [slot 3, slot 2, slot 1, slot 0] where a slot is a list of per-chromosome groups and a
    group is the set of combined clips of one chromosome (order inside a group = container order)."""
    d = {}
    for c in clips:
        d.setdefault(c[1], {}).setdefault(c[2], []).append(c)
    slots = [[] for _ in range(4)]
    for i, chrom in enumerate(chromosomes):
        grp = set()
        for p, cs in d.get(chrom, {}).items():
            grp.add(("", chrom, p, max(c[3] for c in cs), cs[0][4], len(cs)))
        if grp:
            slots[i % 4].append(grp)
    return [slots[3], slots[2], slots[1], slots[0]]


def check_combined_order(order, clips, chromosomes):
    """`order`: combined clips as the shell produced them.  Must be the groups of combine_set in
    sequence, each group in some order."""
    k = 0
    for slot in combine_set(clips, chromosomes):
        for grp in slot:
            got = order[k:k + len(grp)]
            assert set(got) == grp and len(got) == len(grp), (got, grp)
            k += len(grp)
    assert k == len(order)


def filter_lowcovered(clips, w):                           # :54-63
    return [c for c in clips if c[5] >= w]


def filter_tooclose(clips, regions):                       # :97-106, closed intervals
    return [c for c in clips if not any(lo <= c[2] + 1 and c[2] <= hi for lo, hi in regions)]


def cluster(clips, r):                                     # :67-95
    by_pos = {}
    for c in clips:
        found = False
        for key in sorted(by_pos):
            if (key - r) % U32 <= c[2] and c[2] <= (key + r) % U32:
                found = True
                k = by_pos[key]
                by_pos[key] = (k[0], k[1], k[2], max(k[3], c[3]), k[4], k[5] + c[5])
        if not found:
            by_pos[c[2]] = c
    return [by_pos[k] for k in sorted(by_pos)]


def binary_search(clips, begin, end, query):               # :107-122 (begin, end are uint there)
    if begin > end or begin >= len(clips):
        return -1
    m = (begin + end) // 2
    if clips[m][2] == query[2]:
        return m + 1 if m + 1 < len(clips) else m
    elif clips[m][2] > query[2]:
        if m > 0 and clips[m - 1][2] < query[2]:
            return m
        if m == 0:
            return 0        # reference: recursion with end = UINT_MAX, out-of-bounds read (see clipper.hpp)
        return binary_search(clips, begin, m - 1, query)
    else:
        return binary_search(clips, m + 1, end, query)


def preprocess(clips, chromosomes, regions, order):
    u = remove_duplicates(clips)
    check_combined_order(order, u, chromosomes)
    v = filter_lowcovered(order, 2)
    v = filter_tooclose(v, regions)
    v = cluster(v, 1000)
    return sorted(v, key=lambda c: c[2])


def call(clips, chromosomes, seqs, threads, regions, r_order, l_order):   # :124-215
    """Returns (per-thread lists of (type, chrom, s, refbase, w, l), rclips, lclips)."""
    rclips = preprocess([c for c in clips if not c[4]], chromosomes, regions, r_order)
    lclips = preprocess([c for c in clips if c[4]], chromosomes, regions, l_order)
    p = [[] for _ in range(threads)]
    if not lclips or not rclips:
        return p, rclips, lclips
    for i, lc in enumerate(lclips):
        r = binary_search(rclips, 0, len(rclips) - 1, lc)
        if r == -1:
            continue
        rc = rclips[r]
        if rc[5] == 0:
            continue
        if abs(rc[2] - lc[2]) < 1000:
            s = lc[2] if lc[5] > rc[5] else rc[2]
            p[i % threads].append(("INS", lc[1], s, seqs[lc[1]][s], max(lc[5], rc[5]), max(lc[3], rc[3])))
    for i, rc in enumerate(rclips):
        l = binary_search(lclips, 0, len(lclips) - 1, rc)
        if l == -1:
            continue
        lc = lclips[l]
        if lc[5] == 0:
            continue
        d = (lc[2] - rc[2]) % U32
        if 2000 <= d <= 50000:
            w = max(lc[5], rc[5])
            if w >= 5:
                p[i % threads].append(("DEL", rc[1], rc[2], seqs[rc[1]][rc[2]], w, d + 1))
    return p, rclips, lclips


def vcf_lines(p):                                          # caller.cpp:45-54 + sv.cpp:7-27,53-80
    out = []
    for slot in p:
        out[0:0] = slot
    lines = []
    for ty, chrom, s, refbase, w, l in out:
        e = s + len(refbase) - 1
        lines.append("%s\t%d\t%s_%s:%d-%d_%d\t%s\t<%s>\t.\tPASS\tVARTYPE=SV;SVTYPE=%s;SVLEN=%d;END=%d;WEIGHT=%d;COV=0;COV0=0;COV1=0;"
                     "COV2=0;AS=0;NV=0;CIGAR=.;RVEC=;READS=;IMPRECISE\tGT:GQ\t./.:0"
                     % (chrom, s, ty, chrom, s, e, l, refbase, ty, ty, -l if ty == "DEL" else l, e, w))
    return lines
