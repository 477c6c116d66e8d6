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


def _call_worker(rank, world, port, q):
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import oracle
    from poa_cases import make_cluster
    from svdss_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(53)
    clusters = [make_cluster(rng, n_reads=int(rng.integers(2, 7)), tlen=int(rng.integers(30, 140)), rate=0.02)[1] for _ in range(13)]
    clusters.insert(4, [])                                             # an empty cluster keeps its (empty) slot
    costs = parallel.cluster_cost([len(c) for c in clusters], [max([len(r) for r in c] + [0]) for c in clusters])
    mine = parallel.shard_clusters_by_cost(costs, world)[rank]
    # per-rank stand-in for svb_poa_batch + svb_ksw_extd2_batch on CPU: the oracle
    cons = [oracle.poa_consensus(clusters[c], band=True) if clusters[c] else np.zeros(0, np.uint8) for c in mine]
    offs = np.zeros(len(mine) + 1, np.int64); offs[1:] = np.cumsum([len(x) for x in cons])
    got = parallel.gather_ragged(mine, offs, np.concatenate(cons + [np.zeros(0, np.uint8)]), len(clusters), dist, dst=0)
    cig = [np.array([(l << 4) | "MID".index(op) for l, op in oracle.ksw_extd2(x, clusters[c][0])[1]], np.uint32) if len(x) else np.zeros(0, np.uint32)
           for c, x in zip(mine, cons)]
    coffs = np.zeros(len(mine) + 1, np.int64); coffs[1:] = np.cumsum([len(x) for x in cig])
    got_c = parallel.gather_ragged(mine, coffs, np.concatenate(cig + [np.zeros(0, np.uint32)]), len(clusters), dist, dst=0)
    if rank == 0:
        ok = got[1].dtype == np.uint8 and got_c[1].dtype == np.uint32
        for c, reads in enumerate(clusters):
            exp = oracle.poa_consensus(reads, band=True) if reads else np.zeros(0, np.uint8)
            ok &= bool(np.array_equal(got[1][got[0][c]:got[0][c + 1]], exp))
            if len(exp):
                ec = [(l << 4) | "MID".index(op) for l, op in oracle.ksw_extd2(exp, reads[0])[1]]
                ok &= got_c[1][got_c[0][c]:got_c[0][c + 1]].tolist() == ec
        q.put(("ok" if ok else "mismatch", [len(m) for m in parallel.shard_clusters_by_cost(costs, world)]))
    dist.barrier()
    dist.destroy_process_group()


def test_cluster_sharding_by_cost():
    from svdss_b200 import parallel
    costs = parallel.cluster_cost([10, 2, 50, 7, 7, 30], [100, 2000, 300, 400, 400, 100])
    sh = parallel.shard_clusters_by_cost(costs, 2)
    assert sorted(np.concatenate(sh).tolist()) == list(range(6))
    assert sh[0].tolist() == [1, 3, 5] and sh[1].tolist() == [0, 2, 4]      # 8e6 | 4.5e6 | 1.12e6 1.12e6 | 3e5 | 1e5, dealt in turn
    tot = [costs[s].sum() for s in sh]
    assert max(tot) / sum(tot) < 0.7
    assert [len(s) for s in parallel.shard_clusters_by_cost(costs, 8)] == [1, 1, 1, 1, 1, 1, 0, 0]


def test_two_rank_call_gather_matches_single_process():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29950 + os.getpid() % 300
    procs = [ctx.Process(target=_call_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    status, sizes = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert status == "ok" and sizes == [7, 7]


def _rows_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from svdss_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    # the SV tables of bench.py's e2e step: [n, 4] int32 per rank, ragged, one rank with nothing to report
    tabs = [np.arange(12, dtype=np.int32).reshape(3, 4) + 100 * r for r in range(world)]
    tabs[1] = np.zeros((0, 4), np.int32)
    got = parallel.gather_rows(tabs[rank], dist, dst=0)
    got1 = parallel.gather_rows(np.array([7 + rank, 9 + rank], np.int32), dist, dst=0)      # 1-D tables become one column
    if rank == 0:
        ok = np.array_equal(got, np.concatenate(tabs, axis=0)) and got1.shape == (2 * world, 1) and got1[:, 0].tolist() == [7 + r + k * 2 for r in range(world) for k in (0, 1)]
        q.put("ok" if ok else "mismatch")
    else:
        assert got is None and got1 is None                                               # payload goes to rank 0 only
    dist.barrier()
    dist.destroy_process_group()


def test_gather_rows_is_ragged_and_lands_on_rank0_only():
    world = 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_rows_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    status = q.get(timeout=240)
    for p in procs:
        p.join(timeout=120)
    assert status == "ok" and all(p.exitcode == 0 for p in procs)
