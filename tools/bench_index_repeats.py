"""The GPU index builder on a repeat-rich reference: a reference of --mb megabases with --dup-frac of
its bases inside segmental duplications (10-200 kb copies at 1 % divergence, half of them reverse-complemented) and
tandem arrays (units of 2-60 bp repeated over 1-20 kb) -- the worst case for the first pass of the suffix sort
(21-symbol keys) and for the located-match mode.  Prints one JSON line: build time, what SVB_INDEX_STATS reported
(unresolved suffixes after the first pass against the 2^31 limit, doubling rounds), and two size-independent checks
through the search: reads copied from the reference carry no SFS, reads with a planted 40-base random insertion carry
one that covers it; plus the share of extensions the located-match mode answered.
  python tools/bench_index_repeats.py [--mb 1000] [--dup-frac 0.10]"""
import argparse, io, json, os, sys, time, contextlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=1000)
    ap.add_argument("--dup-frac", type=float, default=0.10)
    ap.add_argument("--contigs", type=int, default=8)
    a = ap.parse_args()
    import torch
    from svdss_b200 import capi
    dev = torch.device("cuda", 0)
    n = a.mb * 1_000_000
    g = torch.Generator(device=dev); g.manual_seed(11)
    ref = torch.randint(1, 5, (n,), dtype=torch.uint8, device=dev, generator=g)
    rng = np.random.default_rng(12)
    dup_bases = 0
    n_dups = 0
    while dup_bases < a.dup_frac * n * 0.8:
        L = int(rng.integers(10_000, 200_001))
        src, dst = int(rng.integers(0, n - L)), int(rng.integers(0, n - L))
        seg = ref[src:src + L].clone()
        if rng.random() < 0.5:
            seg = (5 - seg).flip(0)
        m = torch.rand(L, device=dev, generator=g) < 0.01
        seg = torch.where(m, (seg % 4) + 1, seg)
        ref[dst:dst + L] = seg
        dup_bases += L; n_dups += 1
    tand_bases = 0
    n_tand = 0
    while tand_bases < a.dup_frac * n * 0.2:
        unit = torch.randint(1, 5, (int(rng.integers(2, 61)),), dtype=torch.uint8, device=dev, generator=g)
        L = int(rng.integers(1_000, 20_001))
        dst = int(rng.integers(0, n - L))
        ref[dst:dst + L] = unit.repeat(L // len(unit) + 1)[:L]
        tand_bases += L; n_tand += 1
    offs = np.linspace(0, n, a.contigs + 1).astype(np.int64)
    offs_t = torch.from_numpy(offs).to(dev)
    torch.cuda.synchronize()
    os.environ["SVB_INDEX_STATS"] = "1"
    rd, wr = os.pipe()
    saved = os.dup(2); os.dup2(wr, 2)
    t = time.perf_counter()
    try:
        idx = capi.Index.build_device(ref.data_ptr(), offs_t.data_ptr(), a.contigs, device=0)
    finally:
        os.dup2(saved, 2); os.close(wr)
    dt = time.perf_counter() - t
    log = os.read(rd, 1 << 20).decode(errors="replace").strip().splitlines()
    # size-independent checks through the search
    host = None
    reads, want = [], []
    for k in range(4000):
        L = 5000
        c = int(rng.integers(0, a.contigs))
        s = int(rng.integers(offs[c], offs[c + 1] - L))
        r = ref[s:s + L].cpu().numpy().copy()
        if k % 2:
            r = np.where((r >= 1) & (r <= 4), 5 - r, r)[::-1].copy()
        if k % 4 >= 2:                                   # plant a 40-base random insertion
            p = int(rng.integers(500, L - 500))
            r = np.concatenate([r[:p], rng.integers(1, 5, 40).astype(np.uint8), r[p:]])
            want.append((p, p + 40))
        else:
            want.append(None)
        reads.append(np.ascontiguousarray(r, np.uint8))
    cat = np.concatenate(reads); ro = np.zeros(len(reads) + 1, np.int64); ro[1:] = np.cumsum([len(r) for r in reads])
    res = idx.sfs_batch(cat, ro, assemble=True)
    clean_ok = sum(1 for k, w in enumerate(want) if w is None and res.offs[k + 1] == res.offs[k])
    clean_n = sum(1 for w in want if w is None)
    ins_ok = 0
    for k, w in enumerate(want):
        if w is None:
            continue
        sf = res.per_read(k)
        ins_ok += any(q <= w[0] + 39 and q + l >= w[0] + 1 for q, l in sf) and all(q + l > w[0] - 40 and q < w[1] + 40 for q, l in sf)
    print(json.dumps({"reference_mb": a.mb, "contigs": a.contigs, "segmental_duplications": n_dups, "duplicated_bases": dup_bases,
                      "tandem_arrays": n_tand, "tandem_bases": tand_bases, "bwt_symbols": idx.n, "build_s": round(dt, 2),
                      "builder_log": log, "exact_copies_without_sfs": "%d / %d" % (clean_ok, clean_n),
                      "planted_insertions_found_and_nothing_else": "%d / %d" % (ins_ok, len(want) - clean_n),
                      "located_match_share_of_extensions": round(res.n_text_ext / max(1, res.n_ext), 4)}), flush=True)


if __name__ == "__main__":
    main()
