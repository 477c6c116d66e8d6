"""GB/s of svb_bgzf_inflate_device (k_bgzf_inflate, one thread per BGZF member) on a BAM-shaped window: records of
~15 kb reads (4-bit bases, 0xff qualities as the smoothed BAMs of this repo carry, or --quals for random ones),
cut into 0xff00-byte members like htslib does, deflated at --level.  Next to it: zlib on one host core for the same
members (what BgzfSource::fill does per thread).  Prints one JSON line.
  python tools/bench_inflate.py [--mb 256] [--level 6] [--quals]"""
import argparse, json, os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from svdss_b200 import capi


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--mb", type=int, default=256, help="inflated size of the window")
    ap.add_argument("--level", type=int, default=6)
    ap.add_argument("--quals", action="store_true")
    ap.add_argument("--mpw", default="", help="comma list of members-per-warp values to time (SVB_INFLATE_MPW); default: the library's choice only")
    a = ap.parse_args()
    rng = np.random.default_rng(1)
    body = bytearray()
    while len(body) < a.mb << 20:
        l = int(rng.integers(10000, 20000))
        seq = rng.integers(0, 256, size=(l + 1) // 2, dtype=np.uint8)
        seq = (1 << (seq & 3)) | ((1 << ((seq >> 2) & 3)) << 4)                    # two of A C G T (1 2 4 8) per byte
        qual = rng.integers(0, 42, size=l, dtype=np.uint8).tobytes() if a.quals else b"\xff" * l
        body += b"\0" * 36 + b"read/%d/ccs\0" % len(body) + seq.astype(np.uint8).tobytes() + qual + b"XFC\0"
    comps, sizes = [], []
    for o in range(0, len(body), 0xff00):
        d = bytes(body[o:o + 0xff00])
        c = zlib.compressobj(a.level, zlib.DEFLATED, -15)
        comps.append(c.compress(d) + c.flush()); sizes.append(len(d))
    t = time.perf_counter()
    for c in comps:
        zlib.decompress(c, -15)
    host_s = time.perf_counter() - t
    capi.bgzf_inflate_device(comps[:64], sizes[:64])                               # warm-up
    best = None
    for _ in range(3):
        r = capi.bgzf_inflate_device(comps, sizes)
        best = r.kernel_ms if best is None else min(best, r.kernel_ms)
    ok = r.out.tobytes() == bytes(body)
    sweep = {}
    for m in [x for x in a.mpw.split(",") if x]:
        os.environ["SVB_INFLATE_MPW"] = m
        ms, wall = None, None
        for _ in range(2):
            t = time.perf_counter()
            r = capi.bgzf_inflate_device(comps, sizes)
            w = time.perf_counter() - t
            ms = r.kernel_ms if ms is None else min(ms, r.kernel_ms)
            wall = w if wall is None else min(wall, w)
        ok = ok and r.out.tobytes() == bytes(body)
        sweep[m] = {"kernel_ms": round(ms, 2), "call_ms_with_copies_and_python_packing": round(wall * 1e3, 1)}
    os.environ.pop("SVB_INFLATE_MPW", None)
    print(json.dumps({"mpw_sweep": sweep, "kernel": "k_bgzf_inflate" if os.environ.get("SVB_INFLATE_KERNEL") == "thread" else "k_bgzf_inflate_warp", "members": len(comps), "inflated_bytes": len(body), "compressed_bytes": sum(map(len, comps)),
                      "level": a.level, "random_quals": a.quals, "kernel_ms": best, "GB_s_inflated": len(body) / (best * 1e-3) / 1e9,
                      "zlib_one_core_GB_s": len(body) / host_s / 1e9, "identical_to_input": ok}), flush=True)


if __name__ == "__main__":
    main()
