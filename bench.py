#!/usr/bin/env python
"""bench.py -- SFS-extracted reads/sec of the FMD ping-pong search (BASELINE.json configs[1]).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload: 3.1 Gb i.i.d. reference in 24 contigs (seed 3), both strands indexed
(6.2 G-symbol BWT), 1 M smoothed-shaped 15 kb reads per GPU (seed 4 + rank).  A *step* is one pass
of the search over the whole batch.  `value` times the pass with the batch resident in HBM,
`e2e` times svb_sfs_batch() with HOST (pinned) buffers: H2D of the reads and D2H of the SFS table
are inside the timed region.  Multi-GPU: reads shard across ranks (weak scaling, index replicated),
no data-path collective; timing is max over ranks; each rank's pinned staging buffers are allocated on
the NUMA node of its GPU.  At N=1 the line also carries `cpu_baseline` (the CPU port on a bounded
sample), `roofline_rank_walk` / `fmd_rank_microbench` (the "FMD rank GB/s" half of the metric) and
`call_stage` (POA and ksw2 kernels on bounded samples of configs 4 and 5, after everything else).

--impl reference times the CPU port of the same path (oracle/, OpenMP, all host cores) on a bounded
sample of the same workload; the reference binary itself cannot be built offline.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

REF_BP = 3_100_000_000
N_CONTIGS = 24
READS_PER_GPU = 1_000_000
HBM_FALLBACK_GBS = 6650.0


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def contig_offsets(ref_bp, n_contigs):
    w = np.linspace(2.0, 0.5, n_contigs)
    cuts = np.floor(np.cumsum(w / w.sum()) * ref_bp).astype(np.int64)
    cuts[-1] = ref_bp
    return np.concatenate([[0], cuts]).astype(np.int64)


def ncu_traffic(n_reads, block_bytes, key="k_sfs_search_tma"):
    """dram bytes per launch of the search kernel from the committed ncu --set full capture of the
    same workload (config 2 reads; the capture may hold fewer reads per launch than this run: the
    per-read traffic is scaled, and the entry says so); None when the run does not match it."""
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            t = json.load(f)[key]
        if block_bytes == 128 and not os.environ.get("SVB_SEARCH_CFG"):
            return t["dram_bytes"] * (n_reads / t["reads_per_launch"])
    except Exception:
        pass
    return None


def hbm_peak():
    """HBM GB/s from the driver-written MEASURED_PEAKS.json (key `hbm_gbs`, or any numeric entry whose
    key path mentions hbm; the burst figure if both a burst and a sustained one are given: the search
    kernels are timed alone with CUDA events), else the fallback of B200_PROFILING.md."""
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            d = json.load(f)
        if isinstance(d.get("hbm_gbs"), (int, float)):
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        cands = []

        def walk(o, path):
            if isinstance(o, dict):
                for k, v in o.items():
                    walk(v, path + [str(k).lower()])
            elif isinstance(o, (int, float)) and not isinstance(o, bool):
                key = ".".join(path)
                if "hbm" in key and not any(x in key for x in ("tf", "flop", "pct", "frac")):
                    cands.append((key, float(o)))
        walk(d, [])
        if cands:
            cands.sort(key=lambda kv: (0 if "burst" in kv[0] else 1 if "sustain" not in kv[0] else 2))
            key, v = cands[0]
            if v < 100:          # TB/s
                v *= 1000.0
            return v, "measured (MEASURED_PEAKS.json %s)" % key
    except Exception:
        pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks/throttle sampler running during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "50"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                pass
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def gpu_numa_cpus(local, torch):
    """(NUMA node of GPU `local`, the CPUs of that node this process may use), or (None, None)."""
    try:
        p = torch.cuda.get_device_properties(local)
        if all(hasattr(p, k) for k in ("pci_domain_id", "pci_bus_id", "pci_device_id")):
            bdf = "%04x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        else:
            out = subprocess.check_output(["nvidia-smi", "-i", str(local), "--query-gpu=pci.bus_id", "--format=csv,noheader"], text=True)
            bdf = out.strip().lower()[-12:]
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node < 0:
            return None, None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        return (node, cpus) if cpus else (None, None)
    except Exception:
        return None, None


class near_gpu:
    """Host allocations made inside this block are first touched on the NUMA node the GPU hangs off
    (the rank's pinned staging buffers: with several ranks per host, copies that cross the socket
    interconnect share its bandwidth).  The thread's affinity is restored on exit."""

    def __init__(self, local, torch):
        self.node, self.cpus = gpu_numa_cpus(local, torch)
        self.saved = None

    def __enter__(self):
        if self.cpus:
            try:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
            except Exception:
                self.saved = None
        return self

    def __exit__(self, *exc):
        if self.saved is not None:
            try:
                os.sched_setaffinity(0, self.saved)
            except Exception:
                pass
        return False


def build_workload(args, rank, local, torch, capi, synth, need_device_reads=True):
    """Reference on the device, index built on the device, reads materialised on the device."""
    dev = torch.device("cuda", local)
    offs = contig_offsets(args.ref_bp, args.contigs)
    t0 = time.time()
    gen = torch.Generator(device=dev)
    gen.manual_seed(3)
    ref = torch.empty(args.ref_bp, dtype=torch.uint8, device=dev)
    CH = 1 << 30
    for a in range(0, args.ref_bp, CH):  # chunked: randint materialises int64 internally
        b = min(args.ref_bp, a + CH)
        ref[a:b] = torch.randint(1, 5, (b - a,), dtype=torch.uint8, device=dev, generator=gen)
    offs_t = torch.from_numpy(offs).to(dev)
    torch.cuda.synchronize(dev)
    t1 = time.time()
    idx = capi.Index.build_device(ref.data_ptr(), offs_t.data_ptr(), args.contigs, device=local,
                                  block_bytes=args.block_bytes)
    t2 = time.time()
    log("[rank %d] reference %.1fs, index build %.1fs (n=%d, %d-byte blocks, %.2f GB)" %
        (rank, t1 - t0, t2 - t1, idx.n, idx.block_bytes, idx.device_bytes / 1e9))
    segs = synth.make_read_segments(offs, args.reads, seed=4 + rank)
    reads_t = synth.materialize_segments_torch(ref, segs)
    torch.cuda.synchronize(dev)
    del ref
    torch.cuda.empty_cache()
    log("[rank %d] reads: %d, %.2f Gbases, generated in %.1fs" %
        (rank, args.reads, segs["read_offs"][-1] / 1e9, time.time() - t2))
    return idx, reads_t, segs["read_offs"], {"ref_s": t1 - t0, "index_build_s": t2 - t1}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from svdss_b200 import build, capi, synth
    rank, world, local = dist_env()
    build.build_lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsvdss_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    idx, reads_t, read_offs, setup = build_workload(args, rank, local, torch, capi, synth)
    n_reads = len(read_offs) - 1
    offs_t = torch.from_numpy(read_offs).to(dev)
    dreads = capi.DeviceReads(reads_t.data_ptr(), offs_t.data_ptr(), device=local, mem=capi.SVB_MEM_DEVICE,
                              n_reads=n_reads)
    # host copies for the end-to-end arms (pinned): the reads as BAM stores them (4 bits per base, what
    # the reference's loader receives from bam_get_seq) and, for comparison, one nt6 byte per base
    total = int(read_offs[-1])
    host_np = None
    numa = near_gpu(local, torch)
    if world == 1:   # the byte-per-base comparison arm runs on one GPU only (15 GB of pinned memory per rank)
        with numa:
            host = torch.empty(total, dtype=torch.uint8, pin_memory=True)
        host.copy_(reads_t[:total])
        host_np = host.numpy()
    l_qseq = np.diff(read_offs).astype(np.int32)
    seq4_offs = np.zeros(n_reads + 1, np.int64)
    seq4_offs[1:] = np.cumsum((l_qseq.astype(np.int64) + 1) // 2)
    s4o_t = torch.from_numpy(seq4_offs).to(dev)
    packed_t = torch.empty(int(seq4_offs[-1]) + 16, dtype=torch.uint8, device=dev)
    capi.pack4_device(reads_t.data_ptr(), offs_t.data_ptr(), s4o_t.data_ptr(), n_reads, packed_t.data_ptr(), device=local)
    with numa:
        host4 = torch.empty(int(seq4_offs[-1]), dtype=torch.uint8, pin_memory=True)
    host4.copy_(packed_t[:int(seq4_offs[-1])])
    torch.cuda.synchronize(dev)
    del packed_t, s4o_t
    torch.cuda.empty_cache()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        sampler = ClockSampler(local)
        sampler.start()
        if os.environ.get("SVB_PROFILE"):
            torch.cuda.profiler.start()   # ncu --profile-from-start off: capture the timed region only
        ev0 = torch.cuda.Event(enable_timing=True)
        ev1 = torch.cuda.Event(enable_timing=True)
        res = []
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(steps):
            res.append(fn())
        ev1.record()
        barrier()
        wall = (time.perf_counter() - t0) * 1e3
        if os.environ.get("SVB_PROFILE"):
            torch.cuda.profiler.stop()
        clocks = sampler.stop()
        ms = ev0.elapsed_time(ev1)
        t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return res, float(t[0]), float(t[1]), clocks

    assemble = not args.noassemble
    # ---- value: batch resident in HBM
    res, ms_dev, ms_wall, clocks = timed(lambda: idx.sfs_resident(dreads, assemble=assemble), args.steps, args.warmup)
    ms_step = ms_dev / args.steps
    kernel_ms = float(np.mean([r.kernel_ms for r in res]))
    blocks = float(np.mean([r.n_blocks_touched for r in res]))
    n_ext = float(np.mean([r.n_ext for r in res]))
    text_ext = float(np.mean([r.n_text_ext for r in res]))
    launches = int(sum(r.launches for r in res))
    n_sfs = res[-1].n_sfs
    # ---- e2e: host buffers through svb_sfs_batch
    from svdss_b200 import parallel

    def e2e_step(packed=True):
        if packed:
            r = idx.sfs_batch_bam4(host4.data_ptr(), seq4_offs, l_qseq, assemble=assemble)
        else:
            r = idx.sfs_batch(host_np, read_offs, assemble=assemble)
        if world > 1:  # the path's only collective: final gather of the SFS tables on rank 0 (NCCL)
            r.gathered = parallel.gather_sfs(np.diff(r.offs), r.qs, r.len, dist, dst=0, device=dev)
        return r

    res_e, ms_dev_e, ms_wall_e, clocks_e = timed(e2e_step, args.steps, args.warmup)
    ms_step_e = ms_wall_e / args.steps
    assert res_e[-1].n_sfs == n_sfs, "resident and host paths disagree"
    res_b = None
    if host_np is not None:
        res_b, ms_dev_b, ms_wall_b, _ = timed(lambda: e2e_step(False), max(1, args.steps - 1), 1)
        assert res_b[-1].n_sfs == n_sfs, "resident and host (byte) paths disagree"
    peak, peak_src = hbm_peak()
    # algorithmic bytes of one launch: 128 B per distinct index block fetched + 2 B (read byte + text
    # byte) per extension answered in located-match mode
    alg_bytes = blocks * idx.block_bytes + 2.0 * text_ext
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    # the pure FMD rank walk (every extension = an Occ lookup, SURVEY 8d "FMD rank GB/s"): one extra
    # untimed-region pass with the located-match mode switched off, same batch, same kernel
    rank_walk = None
    if text_ext > 0 and not args.no_rank_walk:
        os.environ["SVB_SEARCH_TEXT"] = "0"
        os.environ["SVB_SEARCH_JUMP"] = "0"
        os.environ["SVB_SEARCH_CFG"] = "cpa"     # k_sfs_search_tma<9,1>: the kernel tuned for the pure rank walk
        try:
            idx.sfs_resident(dreads, assemble=assemble)
            rr = [idx.sfs_resident(dreads, assemble=assemble) for _ in range(2)]
        finally:
            del os.environ["SVB_SEARCH_TEXT"]
            del os.environ["SVB_SEARCH_JUMP"]
            del os.environ["SVB_SEARCH_CFG"]
        assert rr[-1].n_sfs == n_sfs and rr[-1].n_ext == res[-1].n_ext, "rank walk and located-match mode disagree"
        rk_ms = float(np.mean([r.kernel_ms for r in rr]))
        rk_bytes = float(np.mean([r.n_blocks_touched for r in rr])) * idx.block_bytes
        rank_walk = {"bound": "hbm", "achieved": rk_bytes / (rk_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": rk_bytes / (rk_ms * 1e-3) / 1e9 / peak, "kernel_ms": rk_ms, "algorithmic_bytes": rk_bytes,
                     "reads_per_s": n_reads / (rk_ms * 1e-3), "extensions_per_s": rr[-1].n_ext / (rk_ms * 1e-3),
                     "traffic": ncu_traffic(n_reads, idx.block_bytes, "k_sfs_search_tma rank walk") if args.ref_bp == REF_BP else None,
                     "kernel": "k_sfs_search_tma<9,1> (SVB_SEARCH_CFG=cpa SVB_SEARCH_TEXT=0 SVB_SEARCH_JUMP=0): every extension is an Occ lookup in a 128 B block"}
    # "FMD rank GB/s" (SURVEY 8d): 2^26 device-generated random (k, k+delta) extensions on this index
    fmd_rank = None
    if not args.no_rank_walk and idx.block_bytes == 128:
        fmd_rank = []
        for delta in (1, 1 << 10, 1 << 20):
            ms, blk = idx.rank_bench(1 << 26, delta, seed=7, iters=3)
            gbs = blk * idx.block_bytes / ms / 1e6
            fmd_rank.append({"delta": delta, "GB_s": gbs, "frac_of_peak": gbs / peak, "blocks_per_extension": blk / float(1 << 26),
                             "G_extensions_per_s": (1 << 26) / ms / 1e6})
    out = {
        "metric": "SFS-extracted reads/sec (FMD ping-pong search)",
        "value": world * n_reads / (ms_step * 1e-3),
        "unit": "reads/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int64", "data": "synthetic",
        "config": {"workload": "configs[1]: FMD ping-pong SFS extraction, %.2f Gb reference (%d contigs, both strands "
                               "indexed), %d smoothed-shaped ~15 kb reads per GPU" % (args.ref_bp / 1e9, args.contigs, n_reads),
                   "reads_per_gpu": n_reads, "bases_per_gpu": total, "index_symbols": idx.n,
                   "index_block_bytes": idx.block_bytes, "index_bytes": idx.device_bytes, "assemble": assemble,
                   "overlap": -1, "parallelism": "read-shard x%d, index replicated, no collective" % world,
                   "l2": "inputs (%.1f GB reads, %.1f GB index) larger than L2" % (total / 1e9, idx.device_bytes / 1e9),
                   "sfs_per_step_rank0": int(n_sfs), "extensions_per_step_rank0": int(n_ext),
                   "index_build_s": round(setup["index_build_s"], 2)},
        "e2e": {"value": world * n_reads / (ms_step_e * 1e-3), "unit": "reads/s",
                "h2d_bytes_per_step": int(res_e[-1].h2d_bytes), "d2h_bytes_per_step": int(res_e[-1].d2h_bytes),
                "ms_per_step": ms_step_e, "device_ms_per_step": ms_dev_e / args.steps,
                "api": "svb_sfs_batch_bam4: pinned host buffer of 4-bit BAM-native reads (bam_get_seq layout), decoded on the GPU",
                "host_buffer_numa_node": numa.node},
        "e2e_nt6_bytes": None if res_b is None else {
            "value": world * n_reads / (ms_wall_b / max(1, args.steps - 1) * 1e-3), "unit": "reads/s",
            "h2d_bytes_per_step": int(res_b[-1].h2d_bytes), "d2h_bytes_per_step": int(res_b[-1].d2h_bytes),
            "api": "svb_sfs_batch: one nt6 byte per base, the reference's in-memory form after its host decode"},
        "gpu_launches": launches + int(sum(r.launches for r in res_e)),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "traffic": ncu_traffic(n_reads, idx.block_bytes, "k_sfs_search_mop") if args.ref_bp == REF_BP else None,
                     "peak_source": peak_src,
                     "kernel": ("k_sfs_search_mop main + tail launch (thread-per-read micro-op pipeline: cp.async-staged 128 B index blocks, warp-cooperative located-match compare, K-mer jump table; parked walks finished warp-per-walk with sprints)"
                                if idx.block_bytes == 128 and not os.environ.get("SVB_SEARCH_CFG")
                                else "k_sfs_search cfg=%s" % os.environ.get("SVB_SEARCH_CFG", "4x1")),
                     "kernel_ms": kernel_ms,
                     "algorithmic_bytes": alg_bytes, "index_blocks": blocks, "text_extensions": text_ext,
                     "extensions_per_s": n_ext / (kernel_ms * 1e-3)},
        "roofline_rank_walk": rank_walk,
        "fmd_rank_microbench": fmd_rank,
        "clocks": clocks,
        "clocks_e2e": clocks_e,
    }
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_baseline(idx, host_np, read_offs, res[-1], assemble, args)
    if rank == 0 and world == 1 and not args.no_call_stage:
        # after everything the line is judged on, in a child process with a time limit: neither an exception
        # nor a crash or a hang of these kernels may cost the line
        out["call_stage"] = call_stage_child(local)
        try:
            out["host_pack2"] = host_pack2_rate(capi, host4.numpy(), seq4_offs, l_qseq)
        except Exception as e:
            out["host_pack2"] = {"error": "%s: %s" % (type(e).__name__, e)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def host_pack2_rate(capi, host4_np, seq4_offs, l_qseq, n=60000):
    """How fast this host re-packs 4-bit reads to 2 bits (svb_pack2_host, all threads): GB/s of 4-bit input on
    the first n reads of the batch.  The number the 2-bit transport stands or falls with:
    above the PCIe rate of the e2e arm it gains, at twice that rate it halves the copy."""
    n = int(min(n, len(l_qseq)))
    lq = np.ascontiguousarray(l_qseq[:n], np.int32)
    so = np.ascontiguousarray(seq4_offs[:n + 1], np.int64)
    oo = np.zeros(n + 1, np.int64)
    oo[1:] = np.cumsum((lq.astype(np.int64) + 3) // 4)
    out = np.ones(max(1, int(oo[-1])), np.uint8)          # touched before the timed calls
    exc = np.ones(max(1, n), np.uint8)
    L = capi.lib()
    best = None
    for _ in range(3):
        t = time.perf_counter()
        rc = L.svb_pack2_host(capi._ptr(host4_np), capi._ptr(so), capi._ptr(lq), n, capi._ptr(out), capi._ptr(oo), capi._ptr(exc), 0)
        dt = time.perf_counter() - t
        if rc != 0:
            raise RuntimeError("svb_pack2_host failed: %d" % rc)
        best = dt if best is None else min(best, dt)
    nbytes = int(so[-1] - so[0])
    return {"GB_s_of_4bit_input": nbytes / best / 1e9, "sample_bytes": nbytes, "threads": os.cpu_count(), "reads_with_other_codes": int(exc[:n].sum())}


def call_stage_child(local, limit_s=240):
    """call_stage_sample() in a child process on the same GPU (`bench.py --call-stage-only`)."""
    try:     # device 0 of the child = the only GPU this (world == 1) run uses
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--call-stage-only"], capture_output=True, text=True,
                           timeout=limit_s, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"error": "call-stage sample did not finish within %d s" % limit_s}
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}
    for line in reversed(r.stdout.splitlines()):
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                break
    return {"error": "child exit %d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:])}


def call_stage_sample(capi, n_clusters=1500, n_pairs=20000):
    """The `call` half of the path on bounded samples of SURVEY 8(d) configs 4 and 5 (they are parity-test
    cases, not the bench line: reported next to it so that one run shows both halves).  POA: clusters of
    20-60 reads x 200-2000 bp; ksw2: consensus x window pairs up to 3 kb (the pipeline's range)."""
    from svdss_b200 import synth
    rng = np.random.default_rng(6)
    cl = synth.gen_clusters(rng, n_clusters)
    capi.poa_batch(cl[:64])
    r = capi.poa_batch(cl)
    out = {"poa": {"clusters": len(cl), "reads": int(sum(len(c) for c in cl)), "kernel_ms": float(r.kernel_ms), "device_ms": float(r.device_ms),
                   "clusters_per_s": len(cl) / (r.device_ms * 1e-3), "GCUPS": r.cells / (r.kernel_ms * 1e-3) / 1e9, "launches": int(r.launches),
                   "variant": int(os.environ.get("SVB_POA_VARIANT", "0")), "lanes_per_cluster": int(os.environ.get("SVB_POA_GROUP", "32")),
                   "api": "svb_poa_batch (host buffers; H2D + kernel + D2H in device_ms)"}}
    pr = synth.gen_pairs(rng, n_pairs, hi=3000)
    qc, qo = synth.concat([p[0] for p in pr])
    tc, to = synth.concat([p[1] for p in pr])
    capi.ksw_extd2_batch(qc[:qo[64]], qo[:65], tc[:to[64]], to[:65])
    k = capi.ksw_extd2_batch(qc, qo, tc, to)
    out["ksw2"] = {"pairs": len(pr), "kernel_ms": float(k.kernel_ms), "device_ms": float(k.device_ms), "pairs_per_s": len(pr) / (k.device_ms * 1e-3),
                   "GCUPS": k.cells / (k.kernel_ms * 1e-3) / 1e9, "waves": int(k.waves), "variant": int(os.environ.get("SVB_KSW_VARIANT", "0")),
                   "api": "svb_ksw_extd2_batch (host buffers)"}
    return out


def cpu_port_time(fm, host_np, read_offs, target_s, threads):
    """time the CPU port on a prefix of the reads sized for ~target_s seconds"""
    n_reads = len(read_offs) - 1
    probe = min(n_reads, 2000)
    offs = np.ascontiguousarray(read_offs[:probe + 1])
    t = time.perf_counter()
    fm.search_batch(host_np, offs, threads=threads, want_output=False)
    dt = max(time.perf_counter() - t, 1e-3)
    m = int(min(n_reads, max(probe, probe * target_s / dt)))
    offs = np.ascontiguousarray(read_offs[:m + 1])
    t = time.perf_counter()
    counts, _, _, _, ext = fm.search_batch(host_np, offs, threads=threads, want_output=False)
    dt = time.perf_counter() - t
    return m, dt, int(ext), counts


def cpu_baseline(idx, host_np, read_offs, gpu_res, assemble, args):
    import oracle
    t = time.time()
    bwt = idx.bwt()
    fm = oracle.FMIndex(bwt)
    del bwt
    log("cpu_baseline: BWT download + CPU block build %.1fs" % (time.time() - t))
    threads = oracle.max_threads()
    m, dt, ext, counts = cpu_port_time(fm, host_np, read_offs, args.cpu_seconds, threads)
    # parity on the sample: raw SFS sets of the CPU port vs the GPU result
    parity = None
    if not assemble:
        parity = bool(np.array_equal(np.diff(gpu_res.offs[:m + 1]), counts))
    else:
        k = min(m, 3000)
        offs = np.ascontiguousarray(read_offs[:k + 1])
        c, ooff, qs, ln, _ = fm.search_batch(host_np, offs, threads=threads)
        ok = True
        for r in range(k):
            exp = oracle.assemble(list(zip(qs[ooff[r]:ooff[r + 1]].tolist(), ln[ooff[r]:ooff[r + 1]].tolist())))
            if exp != gpu_res.per_read(r):
                ok = False
                break
        parity = ok
    return {"value": m / dt, "unit": "reads/s", "cores": threads, "kind": "port",
            "sample": "first %d reads of the same batch, %.1f s, %d extensions (%.1f M ext/s)" % (m, dt, ext, ext / dt / 1e6),
            "parity_vs_gpu_on_sample": parity}


def run_reference(args):
    """CPU port of the reference path on all host cores, same config, bounded sample per step."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    import torch
    import oracle
    from svdss_b200 import build, capi, synth
    build.build_lib()
    torch.cuda.set_device(local)
    # setup (untimed): the 6.2 G-symbol BWT is built on the GPU, then handed to the CPU port
    idx, reads_t, read_offs, setup = build_workload(args, rank, local, torch, capi, synth)
    total = int(read_offs[-1])
    # host sample: the first `sample` reads
    sample = min(len(read_offs) - 1, args.ref_sample)
    nbytes = int(read_offs[sample])
    host_np = reads_t[:nbytes].cpu().numpy()
    offs = np.ascontiguousarray(read_offs[:sample + 1])
    bwt = idx.bwt()
    idx.close()
    del reads_t
    fm = oracle.FMIndex(bwt)
    del bwt
    threads = oracle.max_threads()
    for _ in range(args.warmup):
        fm.search_batch(host_np, np.ascontiguousarray(offs[:min(sample, 500) + 1]), threads=threads, want_output=False)
    t0 = time.perf_counter()
    ext = 0
    for _ in range(args.steps):
        counts, ooff, qs, ln, e = fm.search_batch(host_np, offs, threads=threads)
        ext += e
    dt = time.perf_counter() - t0
    ms_step = dt / args.steps * 1e3
    v = sample / (ms_step * 1e-3)
    out = {"impl": "reference", "metric": "SFS-extracted reads/sec (FMD ping-pong search)", "value": v,
           "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int64", "data": "synthetic",
           "config": {"workload": "configs[1]: FMD ping-pong SFS extraction, %.2f Gb reference, ~15 kb smoothed-shaped reads; "
                                  "CPU port (oracle/, OpenMP) on a bounded sample" % (args.ref_bp / 1e9),
                      "index_symbols": int(fm.n), "index_built_on": "gpu (setup, untimed)",
                      "note": "the reference binary cannot be built offline (ropebwt3/abPOA/ksw2/htslib not vendored)"},
           "cpu_baseline": {"value": v, "unit": "reads/s", "cores": threads, "kind": "port",
                            "sample": "first %d reads per step (%d bases), %.1f M ext/s" % (sample, nbytes, ext / dt / 1e6)},
           "e2e": {"value": v, "unit": "reads/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
           "gpu_launches": 0}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ref-bp", type=int, default=REF_BP)
    ap.add_argument("--contigs", type=int, default=N_CONTIGS)
    ap.add_argument("--reads", type=int, default=READS_PER_GPU)
    ap.add_argument("--block-bytes", type=int, default=0)
    ap.add_argument("--noassemble", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-rank-walk", action="store_true", help="skip the extra pure-rank-walk pass")
    ap.add_argument("--no-call-stage", action="store_true", help="skip the POA / ksw2 samples of the call stage")
    ap.add_argument("--cpu-seconds", type=float, default=15.0)
    ap.add_argument("--ref-sample", type=int, default=20000)
    ap.add_argument("--call-stage-only", action="store_true", help="print the call_stage object alone (the child of the main run)")
    args = ap.parse_args()
    if args.call_stage_only:
        from svdss_b200 import capi
        try:
            obj = call_stage_sample(capi)
        except Exception as e:
            obj = {"error": "%s: %s" % (type(e).__name__, e)}
        print(json.dumps(obj), flush=True)
        return
    _, world, _ = dist_env()
    if world != args.gpus and args.impl == "ours" and world == 1 and args.gpus > 1:
        # convenience: relaunch under torchrun
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29511"] + sys.argv
        raise SystemExit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
