"""Sweep of the search kernel's configurations on configs[1] (1 M unfiltered smoothed-shaped reads by default): the default
micro-op kernel (`mop`: located-match mode + jump table + tail launch) and, with both switched off so that every
extension is an Occ lookup, the block-staging variants of the pure rank walk -- cp.async (`cpa`), TMA bulk copies +
mbarrier (`tma`), and the lane-group kernel (`4x2`) -- so that the TMA-vs-cp.async choice is on record.
usage: python tools/cfg_sweep.py [reads] [ref_bp]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from svdss_b200 import capi, synth

reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_000
ref_bp = int(float(sys.argv[2])) if len(sys.argv) > 2 else bench.REF_BP
class A: pass
a = A(); a.ref_bp = ref_bp; a.contigs = 24; a.block_bytes = 128
idx, ref, offs, setup = bench.build_reference_and_index(a, 0, 0, torch, capi)
segs = synth.make_read_segments(offs, reads, seed=4)
reads_t = synth.materialize_segments_torch(ref, segs)
read_offs = segs["read_offs"]
del ref
offs_t = torch.from_numpy(read_offs).cuda()
dr = capi.DeviceReads(reads_t.data_ptr(), offs_t.data_ptr(), device=0, mem=1, n_reads=reads)
base = None
for name, env in (("mop (default: located match + jump table + tail launch)", {}),
                  ("mop, rank walk only", {"SVB_SEARCH_TEXT": "0", "SVB_SEARCH_JUMP": "0"}),
                  ("cpa, rank walk only (cp.async staging)", {"SVB_SEARCH_CFG": "cpa", "SVB_SEARCH_TEXT": "0", "SVB_SEARCH_JUMP": "0"}),
                  ("tma, rank walk only (cp.async.bulk + mbarrier staging)", {"SVB_SEARCH_CFG": "tma", "SVB_SEARCH_TEXT": "0", "SVB_SEARCH_JUMP": "0"}),
                  ("4x2 lane groups, rank walk only", {"SVB_SEARCH_CFG": "4x2", "SVB_SEARCH_TEXT": "0", "SVB_SEARCH_JUMP": "0"}),
                  ("cpa with located match + jump table", {"SVB_SEARCH_CFG": "cpa"}),
                  ("tma with located match + jump table", {"SVB_SEARCH_CFG": "tma"})):
    os.environ.update(env)
    try:
        idx.sfs_resident(dr)
        rs = [idx.sfs_resident(dr) for _ in range(2)]
    finally:
        for k in env:
            del os.environ[k]
    r = rs[-1]
    ms = float(np.mean([x.kernel_ms for x in rs]))
    sig = (r.n_sfs, r.n_ext, int(r.qs.sum()), int(r.len.sum()))
    base = base or sig
    alg = r.n_blocks_touched * 128 + 2.0 * r.n_text_ext
    print(json.dumps({"config": name, "kernel_ms": round(ms, 2), "reads_per_s": round(reads / ms * 1e3), "G_extensions_per_s": round(r.n_ext / ms / 1e6, 2),
                      "algorithmic_GB_per_s": round(alg / ms / 1e6, 1), "index_blocks": int(r.n_blocks_touched), "text_extensions": int(r.n_text_ext),
                      "same_result": sig == base}), flush=True)
