"""Sweep of lane-group configurations of the search kernel on the config-2 index.
usage: python tools/cfg_sweep.py [reads] [ref_bp]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from svdss_b200 import capi, synth

reads = int(float(sys.argv[1])) if len(sys.argv) > 1 else 300_000
ref_bp = int(float(sys.argv[2])) if len(sys.argv) > 2 else bench.REF_BP
class A: pass
for bb, cfgs in ((128, ["4x2", "tma", "cpa"]),):
    a = A(); a.ref_bp = ref_bp; a.contigs = 24; a.reads = reads; a.block_bytes = bb
    idx, reads_t, read_offs, setup = bench.build_workload(a, 0, 0, torch, capi, synth)
    offs_t = torch.from_numpy(read_offs).cuda()
    dr = capi.DeviceReads(reads_t.data_ptr(), offs_t.data_ptr(), device=0, mem=1, n_reads=reads)
    base = None
    for cfg in cfgs:
        os.environ["SVB_SEARCH_CFG"] = cfg
        idx.sfs_resident(dr)
        r = idx.sfs_resident(dr)
        sig = (r.n_sfs, r.n_ext, int(r.qs.sum()), int(r.len.sum()))
        base = base or sig
        ms, blk = idx.rank_bench(1 << 26, 1, seed=7, iters=3)
        ms2, blk2 = idx.rank_bench(1 << 26, 1 << 20, seed=7, iters=3)
        print(json.dumps({"block_bytes": bb, "cfg": cfg, "kernel_ms": round(r.kernel_ms, 2),
                          "reads_s": round(reads / r.kernel_ms * 1e3), "Gext_s": round(r.n_ext / r.kernel_ms / 1e6, 2),
                          "GB_s": round(r.n_blocks_touched * bb / r.kernel_ms / 1e6, 1), "same_result": sig == base,
                          "rank_d1_Gext_s": round((1 << 26) / ms / 1e6, 2), "rank_d1_GB_s": round(blk * bb / ms / 1e6),
                          "rank_2blk_Gext_s": round((1 << 26) / ms2 / 1e6, 2), "rank_2blk_GB_s": round(blk2 * bb / ms2 / 1e6)}), flush=True)
    dr.close(); idx.close(); del reads_t; torch.cuda.empty_cache()
