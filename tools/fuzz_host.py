"""Mutation fuzzing of everything the host shell parses without a GPU: BAM (through `call --cluster-only --clipped`
and `smooth`), .sfs, FASTA, cluster files, ropebwt3 .fmd.  A finding = an exit code other than 0 / 1 or a sanitizer
report.  Build the shell with sanitizers first for the stricter run:
  g++ -std=c++14 -O1 -g -fsanitize=address,undefined -fopenmp -pthread -o /tmp/SVDSS_asan svdss_b200/host/svdss_main.cpp \\
      -Lsvdss_b200 -lsvdss_b200 -lz -Wl,-rpath,$PWD/svdss_b200
  ASAN_OPTIONS=detect_leaks=0 python tools/fuzz_host.py --exe /tmp/SVDSS_asan [--iters 200]
Round 1: 0 findings in ~1 500 mutated inputs (plain and sanitized builds)."""
import argparse
import gzip
import os
import random
import struct
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))


def mutate(b, rnd, text=False, header=0):
    b = bytearray(b)
    for _ in range(rnd.randint(1, 6)):
        if len(b) < 2:
            break
        p = rnd.randrange(min(header, len(b))) if header and rnd.random() < 0.3 else rnd.randrange(len(b))
        k = rnd.random()
        if k < 0.45:
            b[p] = rnd.choice(b"\t\n -*:>@+0919ACGTNacgtn\x00\xff") if text else rnd.randrange(256)
        elif k < 0.65:
            if text:
                b[p:p] = rnd.choice([b"\n\n", b"\t\t", b"-1", b"99999999999999999999", b">", b"*\t"])
            else:
                b[p:p + 8] = struct.pack("<q", rnd.choice([-1, 0, 1, 2 ** 31 - 1, -2 ** 31, 2 ** 40, 2 ** 62, 10 ** 6]))
        elif k < 0.85:
            del b[p:p + rnd.randint(1, 300)]
        else:
            b = b[:p]
    return bytes(b)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--exe", default=os.path.join(ROOT, "svdss_b200", "SVDSS"))
    ap.add_argument("--iters", type=int, default=100)
    ap.add_argument("--seed", type=int, default=7)
    a = ap.parse_args()
    from bam_writer import _bgzf_block
    from common import oracle_index
    from sv_world import make_world
    from svdss_b200 import synth
    d = tempfile.mkdtemp(prefix="svb_fuzz_")
    w = make_world(d)
    subprocess.run([a.exe, "call", "--reference", w["fa"], "--bam", w["bam"], "--sfs", w["sfs"], "--cluster-only", "--clusters", d + "/clu.txt"], check=True,
                   capture_output=True)
    T, SA, bwt = oracle_index(synth.make_reference(4000, seed=5, contigs=2, n_repeats=1, n_nruns=1, nrun_len=30))
    bwt.tofile(d + "/b.bin")
    subprocess.run([a.exe, "_fmd", "encode", d + "/b.bin", d + "/b.fmd"], check=True)
    bam = gzip.decompress(open(w["bam"], "rb").read())
    sfs, fa, clu, fmd = (open(p, "rb").read() for p in (w["sfs"], w["fa"], d + "/clu.txt", d + "/b.fmd"))
    rnd = random.Random(a.seed)
    findings = 0

    def run(cmd):
        nonlocal findings
        r = subprocess.run([a.exe] + cmd, capture_output=True, timeout=300)
        if r.returncode not in (0, 1) or b"Sanitizer" in r.stderr or b"runtime error" in r.stderr:
            findings += 1
            print("FINDING", cmd, r.returncode, r.stderr[-400:], flush=True)
    for it in range(a.iters):
        m = mutate(bam, rnd)
        with open(d + "/f.bam", "wb") as f:
            for o in range(0, len(m), 60000):
                f.write(_bgzf_block(m[o:o + 60000]))
            f.write(_bgzf_block(b""))
        for name, src, text, hdr in (("f.sfs", sfs, True, 0), ("f.fa", fa, True, 0), ("f.clu", clu, True, 0), ("f.fmd", fmd, False, 72)):
            open(os.path.join(d, name), "wb").write(mutate(src, rnd, text, hdr))
        run(["call", "--reference", w["fa"], "--bam", d + "/f.bam", "--sfs", w["sfs"], "--cluster-only", "--clipped"])
        run(["smooth", "--reference", w["fa"], "--bam", d + "/f.bam"])
        run(["call", "--reference", w["fa"], "--bam", w["bam"], "--sfs", d + "/f.sfs", "--cluster-only", "--clipped"])
        run(["call", "--reference", d + "/f.fa", "--bam", w["bam"], "--sfs", w["sfs"], "--cluster-only"])
        run(["call", "--reference", w["fa"], "--clusters-in", d + "/f.clu"])
        run(["_fmd", "decode", d + "/f.fmd", d + "/o.bin"])
        run(["_fmd", "contigs", d + "/f.fmd", d + "/o.fa"])
    print("fuzz done: %d iterations x 7 commands, findings: %d" % (a.iters, findings))
    return 1 if findings else 0


if __name__ == "__main__":
    sys.exit(main())
