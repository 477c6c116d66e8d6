"""GPU: the C++14 shell end to end -- `SVDSS index` then `SVDSS search` on FASTX and on BAM -- must
print byte-for-byte what the reference's output_batch would (ping_pong.cpp:213-236) for the SFS sets
of the oracle: logical batches of --bsize accepted reads, round-robin thread slots, qname order."""
import os
import subprocess

import numpy as np
import pytest

import oracle
from bam_writer import write_bam
from common import oracle_index, fm_results
from svdss_b200 import build, synth

pytestmark = pytest.mark.gpu
L = "$ACGTN"


def dec(a):
    return "".join(L[int(x)] for x in a)


def expected_sfs_text(names, sfs_lists, htags, searched, threads, bsize, assemble):
    bsize = max(threads, (bsize // threads) * threads)
    out = []
    for b0 in range(0, len(names), bsize):
        b1 = min(len(names), b0 + bsize)
        for t in range(threads):
            slot = {}
            for i in range(b0 + t, b1, threads):
                if searched[i]:
                    slot.setdefault(names[i], []).append(i)
            for name in sorted(slot, key=lambda s: s.encode()):
                first = True
                for i in slot[name]:
                    recs = oracle.assemble(sfs_lists[i]) if assemble else sfs_lists[i]
                    for qs, ln in recs:
                        out.append("%s\t%d\t%d\t%d\t\n" % (name if first else "*", qs, ln, htags[i]))
                        first = False
    return "".join(out)


@pytest.fixture(scope="module")
def world(tmp_path_factory):
    d = tmp_path_factory.mktemp("cli")
    build.build_lib()
    exe = build.build_host()
    contigs = synth.make_reference(300_000, seed=61, contigs=3)
    fa = os.path.join(d, "ref.fa")
    with open(fa, "w") as f:
        for i, c in enumerate(contigs):
            s = dec(c)
            f.write(">chr%d some description\n" % (i + 1))
            for o in range(0, len(s), 70):
                f.write(s[o:o + 70] + "\n")
    idx = os.path.join(d, "ref.svb")
    r = subprocess.run([exe, "index", "-t", "4", "-d", "-o", idx, fa], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    T, SA, bwt = oracle_index(contigs)
    reads = synth.make_reads(contigs, 230, seed=62, mean_len=4000, sd_len=1500, min_len=120, max_len=9000)
    return dict(d=d, exe=exe, idx=idx, contigs=contigs, fm=oracle.FMIndex(bwt), reads=reads)


def test_index_to_stdout_equals_file(world):
    r = subprocess.run([world["exe"], "index", os.path.join(world["d"], "ref.fa")], capture_output=True)
    assert r.returncode == 0
    assert r.stdout == open(world["idx"], "rb").read()


@pytest.mark.parametrize("threads,bsize,assemble", [(4, 10000, True), (3, 50, True), (2, 64, False)])
def test_search_fastx(world, threads, bsize, assemble):
    reads = world["reads"]
    names = ["read_%03d" % ((i * 37) % len(reads)) for i in range(len(reads))]   # not in sorted order
    fq = os.path.join(world["d"], "reads.fq")
    with open(fq, "w") as f:
        for n, r in zip(names, reads):
            f.write("@%s extra\n%s\n+\n%s\n" % (n, dec(r), "I" * len(r)))
    exp, _ = fm_results(world["fm"], reads)
    args = [world["exe"], "search", "--index", world["idx"], "--fastx", fq, "--threads", str(threads), "--bsize", str(bsize)]
    if not assemble:
        args.append("--noassemble")
    r = subprocess.run(args, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    want = expected_sfs_text(names, exp, [0] * len(reads), [True] * len(reads), threads, bsize, assemble)
    assert r.stdout == want
    assert want.count("\n") > 50


def test_search_bam_filters_and_tags(world):
    """BAM mode: unmapped/secondary/supplementary and l_qseq < 100 records are dropped on load
    (ping_pong.cpp:66-75), XF != 0 reads are skipped unless --noputative (:202), HP is carried into
    the 4th column, sequences go through the 4-bit -> nt6 decode (:90-94)."""
    reads = world["reads"][:120]
    recs, names, htags, searched, keep = [], [], [], [], []
    rng = np.random.default_rng(7)
    for i, r in enumerate(reads):
        flag = [0, 16, 4, 256, 2048][i % 5] if i % 7 == 0 else (16 if i % 2 else 0)
        xf = [0, 2, 1][i % 3] if i % 4 == 0 else 0
        hp = int(rng.integers(0, 3))
        seq = dec(r)
        if i == 11:
            seq = seq[:99]          # too short -> filtered
        if i == 13:
            seq = seq[:50] + "RYKM" + seq[54:]     # IUPAC codes decode to N (code 5)
        tags = {}
        if i % 4 == 0:
            tags["XF"] = ("C", xf)
        if hp:
            tags["HP"] = ("c", hp)
        tags["NM"] = 3
        tags["RG"] = "grp1"
        recs.append(dict(qname="q%04d" % (997 * i % 1000), flag=flag, tid=i % 3, pos=100 + i, seq=seq, tags=tags))
        dropped = bool(flag & (4 | 256 | 2048)) or len(seq) < 100
        if not dropped:
            arr = oracle.encode_nt6(seq)
            keep.append(arr); names.append(recs[-1]["qname"]); htags.append(hp); searched.append(xf == 0)
    bam = os.path.join(world["d"], "reads.bam")
    write_bam(bam, [("chr1", 1000), ("chr2", 1000), ("chr3", 1000)], recs)
    exp, _ = fm_results(world["fm"], keep)
    for putative in (True, False):
        args = [world["exe"], "search", "--index", world["idx"], "--bam", bam, "--threads", "4", "--bsize", "40"]
        if not putative:
            args.append("--noputative")
        r = subprocess.run(args, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        want = expected_sfs_text(names, exp, htags, searched if putative else [True] * len(names), 4, 40, True)
        assert r.stdout == want
        # the same file with its BGZF windows inflated on the device (k_bgzf_inflate), one window and many small ones
        for window, walk, submit in ((None, None, None), ("3000", None, None), (None, "parallel", None), ("20000", "parallel", None), ("9000", None, "30000")):
            env = dict(os.environ, SVB_BGZF_GPU_MIN_BYTES="0", SVB_BGZF_STATS="1")   # small files stay on the host threads by default
            if window:
                env["SVB_BGZF_WINDOW"] = window
            if walk:
                env["SVB_BAM_WALK"] = walk            # the segmented record walk, which windows under 1 MB would not use
            if submit:
                env["SVB_SEARCH_SUBMIT_BASES"] = submit   # many searches per file: results wait for the reads that complete their logical batch
            g = subprocess.run(args + ["--gpu-inflate"], capture_output=True, text=True, env=env)
            assert g.returncode == 0, g.stderr
            assert g.stdout == want
            assert "BAM records decoded on GPU" in g.stderr      # the device loader (svb_bamstream_*), not the host parser
    assert "\t1\t\n" in want or "\t2\t\n" in want    # some HP tag made it to the output
    # a damaged member is an error with the device inflate too, not a shorter output
    raw = bytearray(open(bam, "rb").read())
    for k in range(len(raw) // 2, len(raw) // 2 + 64):
        raw[k] ^= 0x5A
    bad = os.path.join(world["d"], "damaged.bam")
    open(bad, "wb").write(bytes(raw))
    for extra in ([], ["--gpu-inflate"]):
        r = subprocess.run([world["exe"], "search", "--index", world["idx"], "--bam", bad] + extra, capture_output=True, text=True,
                           env=dict(os.environ, SVB_BGZF_GPU_MIN_BYTES="0"))
        assert r.returncode != 0 or r.stdout != want, extra


def test_call_core_from_clusters_file(world):
    """`SVDSS call --clusters-in`: POA + ksw2 + CIGAR->SV on the GPU vs the Python restatement over
    the oracle; planted INS/DEL alleles must come out as SV records with exact coordinates."""
    import call_model
    from poa_cases import noisy_copy
    rng = np.random.default_rng(71)
    contigs = world["contigs"]
    ref = {"chr%d" % (i + 1): dec(c).replace("N", "A") for i, c in enumerate(contigs)}
    fa = os.path.join(world["d"], "ref_call.fa")
    with open(fa, "w") as f:
        for k, v in ref.items():
            f.write(">%s\n%s\n" % (k, v.lower() if k == "chr2" else v))     # load_chromosomes upper-cases
    clusters, lines = [], []
    code = lambda s: oracle.CHAR26[np.frombuffer(s.encode(), np.uint8)]
    for ci in range(14):
        chrom = "chr%d" % (1 + ci % 3)
        L = int(rng.integers(300, 900))
        s = int(rng.integers(1000, len(ref[chrom]) - L - 1000))
        e = s + L - 1
        window = code(ref[chrom][s:e + 1])
        kind = ci % 3
        p, k = int(rng.integers(60, L - 200)), int(rng.integers(30, 120))
        if kind == 0:
            allele = np.concatenate([window[:p], rng.integers(0, 4, size=k).astype(np.uint8), window[p:]])   # INS
        elif kind == 1:
            allele = np.concatenate([window[:p], window[p + k:]])                                            # DEL
        else:
            allele = window                                                                                   # no SV
        n = int(rng.integers(1, 9)) if ci != 5 else 1      # ci == 5: below --min-cluster-weight
        sub = [("r%d_%d" % (ci, j), "".join("ACGT"[c] for c in noisy_copy(rng, allele, 0.002))) for j in range(n)]
        if ci == 7:   # second length group: reference-length reads next to the allele reads
            sub += [("w%d_%d" % (ci, j), "".join("ACGT"[c] for c in window)) for j in range(3)]
        clusters.append((chrom, s, e, sub))
        lines.append("%s:%d-%d\t%d\t%s" % (chrom, s + 1, e + 1, len(sub), "\t".join("%s:%s" % x for x in sub)))
    cfile = os.path.join(world["d"], "clusters.txt")
    open(cfile, "w").write("\n".join(lines) + "\n")
    sam = os.path.join(world["d"], "poa.sam")
    r = subprocess.run([world["exe"], "call", "--reference", fa, "--clusters-in", cfile, "--poa", sam,
                        "--min-cluster-weight", "2", "--min-sv-length", "25"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    out = r.stdout.splitlines()
    header = [l for l in out if l.startswith("#")]
    body = [l for l in out if not l.startswith("#")]
    assert header[0] == "##fileformat=VCFv4.2" and header[-1].startswith("#CHROM\tPOS\tID\tREF\tALT")
    assert "##contig=<ID=chr1,length=%d>" % len(ref["chr1"]) in header
    want = call_model.call_vcf_lines(ref, clusters)
    assert sorted(body) == sorted(want)
    assert [l.split("\t")[:2] for l in body] == sorted([l.split("\t")[:2] for l in body], key=lambda x: (x[0], int(x[1])))
    assert sum("SVTYPE=INS" in l for l in body) >= 3 and sum("SVTYPE=DEL" in l for l in body) >= 3
    samtxt = open(sam).read().splitlines()
    assert samtxt[0] == "@HD\tVN:1.4" and sum(1 for l in samtxt if not l.startswith("@")) >= 12
