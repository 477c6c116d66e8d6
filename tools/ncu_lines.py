"""Attribute the SASS execution counts of an `ncu --page source --csv` dump to CUDA source lines, using
the line table of the built library (nvdisasm -g): ncu's own CUDA view loses its metrics in CSV form.
  python tools/ncu_lines.py gpurun_out/prof_X_source.csv k_sfs_search_mop [top_n]
The library must be the build that was profiled."""
import collections, csv, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_with_lines(kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "svdss_b200", "libsvdss_b200.so")], cwd=d, capture_output=True)
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout.splitlines()
        start = next((i for i, l in enumerate(txt) if re.match(r"\s*\.section\s+\.text\.\S*%s" % kernel, l)), None)
        if start is None:
            continue
        seq, cur = [], None
        for l in txt[start + 1:]:
            if re.match(r"\s*\.section", l):
                break
            m = re.search(r'//## File "([^"]+)", line (\d+)', l)
            if m:
                cur = (os.path.basename(m.group(1)), int(m.group(2)))
                continue
            if re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", l):
                seq.append(cur)
        return seq
    raise SystemExit("kernel %s not found in the library" % kernel)


def main(path, kernel, top=40):
    rows = list(csv.reader(open(path)))
    hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    lines = sass_with_lines(kernel)
    if len(lines) != len(data):
        print("# warning: %d SASS lines in the library vs %d in the profile (different build?)" % (len(lines), len(data)))
    agg, smp = collections.Counter(), collections.Counter()
    tot = 0
    for ln, r in zip(lines, data):
        c = int(r[ci["Instructions Executed"]]); tot += c
        agg[ln] += c; smp[ln] += int(r[ci["# Samples"]] or 0)
    src = {}
    print("# %s: warp instructions executed %.4g; share and stall samples per source line" % (kernel, tot))
    for ln, c in agg.most_common(int(top)):
        f, l = ln if ln else ("?", 0)
        if f not in src:
            p = os.path.join(ROOT, "svdss_b200", "csrc", f)
            src[f] = open(p).read().splitlines() if os.path.exists(p) else None
        text = src[f][l - 1].strip()[:100] if src[f] and 0 < l <= len(src[f]) else ""
        print("%5.1f%% %7d  %s:%d  %s" % (100 * c / tot, smp[ln], f, l, text))


if __name__ == "__main__":
    main(*sys.argv[1:4])
