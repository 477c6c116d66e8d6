"""End-to-end `SVDSS index | search | call` through the C++ shell on a scaled config-3 sample
(SURVEY 8d): diploid sample with a planted INS/DEL catalogue, coordinate-sorted smoothed-shaped BAM
(XF:i:0 on reads carrying an event or clip, XF:i:2 otherwise; HP tags), reads ~15 kb at --coverage.
Prints one JSON line: wall-clock of each stage, reads/s over all BAM records and over the records
actually searched, and SV recall/precision against the planted catalogue.
  python tools/bench_e2e.py [--ref-bp 20000000] [--coverage 30] [--svs 200]"""
import argparse, json, os, struct, subprocess, sys, tempfile, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from svdss_b200 import build, synth

NT16 = np.zeros(6, np.uint8); NT16[1:6] = [1, 2, 4, 8, 15]      # nt6 code -> BAM 4-bit code (A C G T N)


def bgzf_block(data, level=1):
    comp = zlib.compressobj(level, zlib.DEFLATED, -15)
    body = comp.compress(data) + comp.flush()
    return (struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, len(body) + 25) + body +
            struct.pack("<II", zlib.crc32(data) & 0xffffffff, len(data)))


def write_bam_fast(path, refs, recs):
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join("@SQ\tSN:%s\tLN:%d\n" % r for r in refs)
    head = bytearray(b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(refs)))
    for name, ln in refs:
        head += struct.pack("<i", len(name) + 1) + name.encode() + b"\0" + struct.pack("<i", ln)
    with open(path, "wb") as f:
        buf = bytearray(head)

        def flush(final=False):
            nonlocal buf
            while len(buf) >= 60000 or (final and buf):
                f.write(bgzf_block(bytes(buf[:60000]))); del buf[:60000]
        for r in recs:
            seq = r["seq"]
            c = NT16[seq]
            if len(c) & 1:
                c = np.concatenate([c, [0]]).astype(np.uint8)
            packed = ((c[0::2] << 4) | c[1::2]).astype(np.uint8).tobytes()
            qn = r["qname"].encode() + b"\0"
            cigb = b"".join(struct.pack("<I", (l << 4) | "MIDNSHP=X".index(op)) for l, op in r["cigar"])
            aux = b"XFC" + struct.pack("<B", 0 if r["has_event"] else 2)
            if r["hp"]:
                aux += b"HPC" + struct.pack("<B", r["hp"])
            core = struct.pack("<iiBBHHHiiii", r["tid"], r["pos"], len(qn), 60, 4680, len(r["cigar"]), 0, len(seq), -1, -1, 0)
            body = core + qn + cigb + packed + b"\xff" * len(seq) + aux
            buf += struct.pack("<i", len(body)) + body
            flush()
        flush(final=True)
        f.write(bgzf_block(b""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--ref-bp", type=int, default=20_000_000)
    ap.add_argument("--contigs", type=int, default=4)
    ap.add_argument("--coverage", type=float, default=30.0)
    ap.add_argument("--svs", type=int, default=200)
    ap.add_argument("--threads", type=int, default=4)
    ap.add_argument("--keep", default="")
    ap.add_argument("--gpu-inflate", action="store_true", help="run smooth / search / call a second time with --gpu-inflate: same bytes out, wall clock next to the host inflate")
    ap.add_argument("--raw", action="store_true", help="raw-HiFi-shaped input: run `SVDSS smooth` first (0.1%% substitutions, 0.05%% 1-bp indels)")
    a = ap.parse_args()
    build.build_lib(); exe = build.build_host()
    d = a.keep or tempfile.mkdtemp(prefix="svb_e2e_")
    os.makedirs(d, exist_ok=True)
    t0 = time.perf_counter()
    contigs = synth.make_reference(a.ref_bp, seed=3, contigs=a.contigs, n_repeats=20, n_nruns=5)
    names = ["chr%d" % (i + 1) for i in range(len(contigs))]
    cat = synth.make_sv_catalogue(contigs, a.svs, seed=5, min_len=50, max_len=5000, margin=20000, spacing=20000)
    recs = synth.make_sample_alignments(contigs, cat, coverage=a.coverage, seed=6, mean_len=15000, sd_len=2000, min_len=5000,
                                        max_len=25000, tag_hp=True, clip_rate=0.05, sub_rate=0.001 if a.raw else 0.0,
                                        indel_rate=0.0005 if a.raw else 0.0)
    L = np.frombuffer(b"$ACGTN", np.uint8)
    fa = os.path.join(d, "ref.fa")
    with open(fa, "wb") as f:
        for n, c in zip(names, contigs):
            f.write(b">" + n.encode() + b"\n" + L[c].tobytes() + b"\n")
    bam = os.path.join(d, "sample.bam")
    write_bam_fast(bam, [(n, len(c)) for n, c in zip(names, contigs)], recs)
    gen_s = time.perf_counter() - t0
    n_reads = len(recs); n_bases = int(sum(len(r["seq"]) for r in recs)); n_searched = sum(1 for r in recs if r["has_event"])

    def stage(args, out=None):
        t = time.perf_counter()
        r = subprocess.run(args, stdout=open(out, "wb") if out else subprocess.PIPE, stderr=subprocess.PIPE)
        dt = time.perf_counter() - t
        if r.returncode != 0:
            sys.stderr.write(r.stderr.decode()); raise SystemExit("stage failed: " + " ".join(args))
        return dt, r.stderr.decode()
    idx, sfs, vcf = os.path.join(d, "ref.svb"), os.path.join(d, "sample.sfs"), os.path.join(d, "calls.vcf")
    t_smooth = None
    if a.raw:
        smoothed = os.path.join(d, "smoothed.bam")
        t_smooth, log_smooth = stage([exe, "smooth", "--reference", fa, "--bam", bam, "--threads", str(a.threads)], smoothed)
        sys.stderr.write(log_smooth)
        bam = smoothed
    t_index, _ = stage([exe, "index", "-t", str(a.threads), "-d", "-o", idx, fa])
    t_search, _ = stage([exe, "search", "--index", idx, "--bam", bam, "--threads", str(a.threads)], sfs)
    t_call, log_call = stage([exe, "call", "--reference", fa, "--bam", bam, "--sfs", sfs, "--threads", str(a.threads)], vcf)
    dev = None
    if a.gpu_inflate:
        g = ["--gpu-inflate"]
        dev = {}
        if a.raw:
            dev["smooth_s"], _ = stage([exe, "smooth", "--reference", fa, "--bam", os.path.join(d, "sample.bam"), "--threads", str(a.threads)] + g, smoothed + ".dev")
            dev["smoothed_bam_identical"] = open(smoothed, "rb").read() == open(smoothed + ".dev", "rb").read()
            os.remove(smoothed + ".dev")
        dev["search_s"], _ = stage([exe, "search", "--index", idx, "--bam", bam, "--threads", str(a.threads)] + g, sfs + ".dev")
        dev["call_s"], _ = stage([exe, "call", "--reference", fa, "--bam", bam, "--sfs", sfs, "--threads", str(a.threads)] + g, vcf + ".dev")
        dev["sfs_identical"] = open(sfs, "rb").read() == open(sfs + ".dev", "rb").read()
        dev["vcf_identical"] = open(vcf, "rb").read() == open(vcf + ".dev", "rb").read()
        dev = {k: (round(v, 2) if isinstance(v, float) else v) for k, v in dev.items()}
    calls = [l.split("\t") for l in open(vcf) if not l.startswith("#")]
    hit = 0
    used = set()
    for sv in cat:
        chrom = names[sv["contig"]]
        for k, f in enumerate(calls):
            if k in used or f[0] != chrom or abs(int(f[1]) - (sv["pos"] + 1)) > 50 or ("SVTYPE=%s;" % sv["type"]) not in f[7]:
                continue
            svlen = abs(int(f[7].split("SVLEN=")[1].split(";")[0]))
            if abs(svlen - sv["len"]) <= max(2, sv["len"] // 50):
                hit += 1; used.add(k)
                break
    out = {"workload": "config 3 scaled: %.1f Mb reference x%d contigs, %.0fx coverage, %d planted SVs" % (a.ref_bp / 1e6, len(contigs), a.coverage, len(cat)),
           "bam_records": n_reads, "bases": n_bases, "searched_records": n_searched, "sfs_lines": sum(1 for _ in open(sfs)),
           "smooth_s": None if t_smooth is None else round(t_smooth, 2), "index_s": round(t_index, 2), "search_s": round(t_search, 2), "call_s": round(t_call, 2),
           "reads_per_s_search_call_all_records": round(n_reads / (t_search + t_call), 1),
           "reads_per_s_search_call_searched": round(n_searched / (t_search + t_call), 1),
           "calls": len(calls), "planted": len(cat), "recall": round(hit / max(1, len(cat)), 3),
           "precision": round(len(used) / max(1, len(calls)), 3), "generate_s": round(gen_s, 1),
           "note": "wall clock of the CLI stages (process start, BGZF inflate, index load, GPU work, text output)"}
    if dev is not None:
        out["with_gpu_inflate"] = dev
    print(json.dumps(out), flush=True)
    sys.stderr.write(log_call)


if __name__ == "__main__":
    main()
