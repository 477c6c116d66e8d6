"""Read sharding across the GPUs of one box and the final SFS gather.

The path has no exchange step: reads are independent (ping_pong.cpp:191-206), so every rank owns
a contiguous range of the batch and a replica of the index; the only communication is gathering the
per-read SFS tables on rank 0.  One process per GPU, `torch.distributed` for plumbing
(NCCL on the GPU box; the same code runs under gloo in the CPU tests)."""
import numpy as np


def shard_range(n_items, rank, world):
    """contiguous, balanced [lo, hi) of rank"""
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_reads_by_bases(offs, world):
    """cut points that balance BASES (work ~ bases) instead of read counts: list of world+1 indices"""
    offs = np.asarray(offs, np.int64)
    total = int(offs[-1] - offs[0])
    cuts = [0]
    for r in range(1, world):
        cuts.append(int(np.searchsorted(offs - offs[0], total * r // world, side="left")))
    cuts.append(len(offs) - 1)
    return [min(max(c, cuts[i - 1] if i else 0), len(offs) - 1) for i, c in enumerate(cuts)]


def gather_rows(local, dist, dst=0, device="cpu"):
    """Gather int32 tables of shape [n_r, k] (k equal on every rank) on rank `dst`, rank order; None elsewhere.
    One tiny all_gather of the row counts, then ONE padded gather to `dst` -- nobody but `dst` receives payload."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    local = np.ascontiguousarray(local, np.int32)
    if local.ndim == 1:
        local = local.reshape(-1, 1)
    k = local.shape[1]
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.cpu()[0]) for c in counts]
    m = max(max(counts), 1)
    pad = torch.zeros((m, k), dtype=torch.int32, device=device)
    if local.shape[0]:
        pad[:local.shape[0]] = torch.as_tensor(local, device=device)
    out = [torch.zeros_like(pad) for _ in range(world)] if rank == dst else None
    dist.gather(pad, out, dst=dst)
    if rank != dst:
        return None
    return np.concatenate([t.cpu().numpy()[:c] for t, c in zip(out, counts)], axis=0)


def gather_sfs(local_counts, local_qs, local_len, dist, dst=0, device="cpu"):
    """Gather per-read SFS tables on rank `dst` in global read order (the path's only exchange).
    local_counts: int[n_local_reads]; local_qs/local_len: int32[sum(counts)].  int32 on the wire, payload to `dst` only."""
    counts = gather_rows(np.asarray(local_counts, np.int32), dist, dst, device)
    recs = gather_rows(np.stack([np.asarray(local_qs, np.int32), np.asarray(local_len, np.int32)], axis=1) if len(local_qs)
                       else np.zeros((0, 2), np.int32), dist, dst, device)
    if counts is None:
        return None
    counts = counts[:, 0].astype(np.int64)
    offs = np.zeros(len(counts) + 1, np.int64)
    offs[1:] = np.cumsum(counts)
    return offs, recs[:, 0].copy(), recs[:, 1].copy()


# ---- `call`: clusters are independent (caller.cpp:312-313); shard them by cost, gather the ragged results

def cluster_cost(n_reads, max_len):
    """DP work of one cluster: every read is aligned to a graph about as long as the reads."""
    return np.asarray(n_reads, np.float64) * np.asarray(max_len, np.float64) ** 2


def shard_clusters_by_cost(costs, world):
    """Cost-sorted round robin (largest first, ties by index): list of `world` int64 index arrays, each in
    ascending cluster order.  Every rank gets a similar mix of big and small clusters, no communication."""
    costs = np.asarray(costs, np.float64)
    order = np.lexsort((np.arange(len(costs)), -costs))
    return [np.sort(order[r::world]).astype(np.int64) for r in range(world)]


def gather_ragged(local_ids, local_offs, local_payload, n_total, dist, dst=0, device="cpu"):
    """Gather variable-length records (consensus bases, CIGAR words) computed for the clusters `local_ids`
    of this rank on rank `dst`, in global cluster order.  local_offs: int64[len(local_ids) + 1] into
    local_payload (any integer dtype; carried as int64).  Returns (offs int64[n_total + 1], payload) on dst,
    None elsewhere.  Same two-collective shape as gather_sfs: sizes, then one padded all_gather."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    local_ids = np.asarray(local_ids, np.int64)
    local_offs = np.asarray(local_offs, np.int64)
    payload = np.asarray(local_payload)
    lens = np.diff(local_offs)
    sizes = torch.tensor([len(local_ids), len(payload)], dtype=torch.int64, device=device)
    all_sizes = [torch.zeros(2, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(all_sizes, sizes)
    all_sizes = [tuple(int(x) for x in s.cpu()) for s in all_sizes]
    max_c = max(s[0] for s in all_sizes)
    max_p = max(s[1] for s in all_sizes)
    pad = torch.zeros(2 * max_c + max_p, dtype=torch.int64, device=device)
    pad[:len(local_ids)] = torch.as_tensor(local_ids, device=device)
    pad[max_c:max_c + len(lens)] = torch.as_tensor(lens.astype(np.int64), device=device)
    pad[2 * max_c:2 * max_c + len(payload)] = torch.as_tensor(payload.astype(np.int64), device=device)
    out = [torch.zeros_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    if rank != dst:
        return None
    all_len = np.zeros(n_total, np.int64)
    parts = []
    for (nc, npay), t in zip(all_sizes, out):
        t = t.cpu().numpy()
        ids, ln = t[:nc], t[max_c:max_c + nc]
        all_len[ids] = ln
        parts.append((ids, ln, t[2 * max_c:2 * max_c + npay]))
    offs = np.zeros(n_total + 1, np.int64)
    offs[1:] = np.cumsum(all_len)
    res = np.zeros(int(offs[-1]), payload.dtype)
    for ids, ln, pay in parts:
        src = np.zeros(len(ids) + 1, np.int64)
        src[1:] = np.cumsum(ln)
        for k, c in enumerate(ids):
            res[offs[c]:offs[c + 1]] = pay[src[k]:src[k + 1]].astype(payload.dtype)
    return offs, res
