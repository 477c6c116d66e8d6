"""world_size-2 gloo test of the N>1 host path: shard reads, search each shard, gather the SFS
tables on rank 0 -- must equal the single-process result. The per-rank search stand-in on CPU is the
oracle's FM port (tests may use the oracle); on the GPU box bench.py runs the same sharding with the
CUDA path and NCCL."""
import os
import sys

import numpy as np
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle
    from common import oracle_index
    from svdss_b200 import parallel, synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    contigs = synth.make_reference(120_000, seed=51, contigs=2)
    reads = synth.make_reads(contigs, 41, seed=52, mean_len=2500, sd_len=900, min_len=100, max_len=6000)
    T, SA, bwt = oracle_index(contigs)
    fm = oracle.FMIndex(bwt)
    cat, offs = oracle.concat(reads)
    cuts = parallel.shard_reads_by_bases(offs, world)
    lo, hi = cuts[rank], cuts[rank + 1]
    sub_offs = np.ascontiguousarray(offs[lo:hi + 1])
    counts, ooff, qs, ln, _ = fm.search_batch(cat, sub_offs)
    got = parallel.gather_sfs(counts, qs, ln, dist, dst=0)
    if rank == 0:
        c_all, o_all, q_all, l_all, _ = fm.search_batch(cat, offs)
        ok = (np.array_equal(got[0], o_all) and np.array_equal(got[1], q_all) and np.array_equal(got[2], l_all))
        q.put(("ok" if ok else "mismatch", int(got[0][-1]), cuts))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_helpers():
    from svdss_b200 import parallel
    assert [parallel.shard_range(10, r, 3) for r in range(3)] == [(0, 4), (4, 7), (7, 10)]
    assert parallel.shard_range(2, 3, 4) == (2, 2)
    offs = np.array([0, 10, 20, 100, 110, 120], np.int64)
    cuts = parallel.shard_reads_by_bases(offs, 2)
    assert cuts[0] == 0 and cuts[-1] == 5 and 0 < cuts[1] < 5


def test_two_rank_gather_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, n_sfs, cuts = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert status == "ok" and n_sfs > 0
