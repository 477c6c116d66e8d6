"""Rank microbenchmark sweep (SURVEY 8d "FMD rank GB/s"): random (k, k+delta) extensions on a block
array of the config-2 size.  The block array is built from a synthetic BWT (uniform symbols) --
rank throughput depends on the array size and layout, not on the BWT being a real one."""
import json
import sys
import os
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ctypes as C
import torch
from svdss_b200 import capi

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 6_200_000_000
L = capi.lib()
for bb in (64, 128):
    torch.manual_seed(1)
    bwt = torch.randint(1, 5, (n,), dtype=torch.uint8, device="cuda")
    h = C.c_void_p()
    t = time.time()
    capi.check(L.svb_index_from_bwt(C.c_void_p(bwt.data_ptr()), n, 1, 0, bb, C.byref(h)))
    idx = capi.Index(h)
    del bwt
    torch.cuda.empty_cache()
    for delta in (1, 1 << 10, 1 << 20):
        ms, blk = idx.rank_bench(1 << 26, delta, seed=7, iters=5)
        gbs = blk * bb / ms / 1e6
        print(json.dumps({"block_bytes": bb, "n": n, "delta": delta, "ms": round(ms, 3), "blocks": blk,
                          "Gext_s": round((1 << 26) / ms / 1e6, 3), "GB_s": round(gbs, 1),
                          "build_s": round(time.time() - t, 2)}))
    idx.close()
