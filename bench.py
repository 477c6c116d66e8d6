#!/usr/bin/env python
"""bench.py -- SFS-extracted reads/sec of SVDSS's hot path, search + call (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (one slice per GPU): 3.1 Gb i.i.d. reference in 24 contigs (seed 3), both strands indexed
(6.2 G-symbol BWT); a 30x coordinate-sorted smoothed-shaped BAM with a planted catalogue of 20 000 het/hom INS/DEL of
50-5000 bp (seed 5); rank r of N owns records [r*R, (r+1)*R) of it (R = --reads, 775 000 = 1/8 of the 6.2 M records,
so that 8 GPUs hold the whole 30x sample).  A *step* is one pass of the path over the rank's slice:
  search   PingPong::process_batch on the records the putative filter keeps (XF == 0, ping_pong.cpp:196-203) -> SFS table
  cluster  Clusterer::run: SFSs placed on the reference, clustered, sub-reads cut (svb_cluster_batch)
  call     Caller::pcall: split_cluster, POA consensus, ksw2 realignment, CIGAR walk -> SV records (svb_call_batch)
`value` = all N slices' records / step time with the searched reads and the reference resident in HBM; `e2e` = the same
step from HOST buffers (4-bit BAM-native reads in pinned memory in, SV records out, copies inside the timed region).
Multi-GPU: slices are independent (weak scaling, index replicated), the only exchange is the gather of the SV records
on rank 0 (NCCL, timed inside e2e); timing is max over ranks.  At N=1 the line also carries `cpu_baseline` (the CPU
port of the same step on a bounded sample), `search_config2` (configs[1]: 1 M unfiltered reads through the search
alone, with the FMD-rank roofline objects) and `call_stage` (configs[3] and [4]: POA- and ksw2-heavy shapes).

--impl reference times the CPU port of the same step (oracle/, OpenMP, all host cores) on a bounded sample of rank
0's slice; the reference binary itself cannot be built offline.
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
READS_PER_GPU = 775_000          # records of the 30x BAM per GPU: 6.2 M / 8
CONFIG2_READS = 1_000_000
N_SVS = 20_000
COVERAGE = 30.0
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


def host_threads():
    """threads the CPU arms may use: the affinity mask, not OMP_NUM_THREADS (torchrun exports OMP_NUM_THREADS=1)"""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def build_reference_and_index(args, rank, local, torch, capi):
    """Reference on the device, index built on the device."""
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
    idx = capi.Index.build_device(ref.data_ptr(), offs_t.data_ptr(), args.contigs, device=local, block_bytes=args.block_bytes)
    t2 = time.time()
    log("[rank %d] reference %.1fs, index build %.1fs (n=%d, %d-byte blocks, %.2f GB)" %
        (rank, t1 - t0, t2 - t1, idx.n, idx.block_bytes, idx.device_bytes / 1e9))
    return idx, ref, offs, {"ref_s": t1 - t0, "index_build_s": t2 - t1}


class Slice:
    """One rank's slice of the config-3 BAM: alignment arrays (host), the searched reads (device nt6 + pinned 4-bit host)."""

    def __init__(self, args, rank, local, ref, offs, torch, capi, synth, n_records=None, want_host=True):
        dev = torch.device("cuda", local)
        self.cat = synth.make_sv_catalogue_arrays(offs, n_svs=args.svs, seed=5)
        n = n_records or args.reads
        span = int(args.reads * 15000 / COVERAGE)
        self.S = S = synth.make_sample_region(offs, self.cat, n, rank * span, coverage=COVERAGE, seed=6 + rank)
        self.n = n
        self.searched = S["searched"]
        self.ns = len(self.searched)
        self.read_offs = S["segs"]["read_offs"]
        self.reads_t = synth.materialize_segments_torch(ref, S["segs"])
        self.offs_t = torch.from_numpy(self.read_offs).to(dev)
        self.bases = int(self.read_offs[-1])
        self.dreads = capi.DeviceReads(self.reads_t.data_ptr(), self.offs_t.data_ptr(), device=local, mem=capi.SVB_MEM_DEVICE, n_reads=self.ns)
        # per-alignment offsets of the searched reads' sequences (-1: never searched, hence never a sub-read)
        self.seq_offs_dev = np.full(n, -1, np.int64)
        self.seq_offs_dev[self.searched] = self.read_offs[:-1]
        self.l_qseq = np.ascontiguousarray(S["l_qseq"][self.searched], np.int32)
        self.seq4_offs = np.zeros(self.ns + 1, np.int64)
        self.seq4_offs[1:] = np.cumsum((self.l_qseq.astype(np.int64) + 1) // 2)
        self.seq_offs_host4 = np.full(n, -1, np.int64)
        self.seq_offs_host4[self.searched] = self.seq4_offs[:-1]
        self.host4 = None
        self.numa = near_gpu(local, torch)
        if want_host:
            s4o_t = torch.from_numpy(self.seq4_offs).to(dev)
            packed_t = torch.empty(int(self.seq4_offs[-1]) + 16, dtype=torch.uint8, device=dev)
            capi.pack4_device(self.reads_t.data_ptr(), self.offs_t.data_ptr(), s4o_t.data_ptr(), self.ns, packed_t.data_ptr(), device=local)
            with self.numa:
                self.host4 = torch.empty(int(self.seq4_offs[-1]) + 16, dtype=torch.uint8, pin_memory=True)
            self.host4.copy_(packed_t)
            torch.cuda.synchronize(dev)
            self.host4_np = self.host4.numpy()

    def pin(self, torch):
        """the record arrays the Clusterer gets every step, in pinned host memory (what a loader that fills them once per
        slice would use): the uploads inside svb_cluster_batch then run at PCIe speed instead of through bounce buffers"""
        S = self.S
        self._pinned = {}
        for k, dt in (("tid", np.int32), ("pos", np.int32), ("hp", np.int32), ("cigar_offs", np.int64), ("cigar", np.uint32)):
            a = np.ascontiguousarray(S[k], dt)
            t = torch.empty(max(1, a.nbytes), dtype=torch.uint8, pin_memory=True)
            v = t.numpy()[:a.nbytes].view(dt)
            v[:] = a
            self._pinned[k] = (t, v)
        t = torch.empty((self.n + 1) * 8, dtype=torch.uint8, pin_memory=True)
        self._so = (t, t.numpy().view(np.int64))
        self._cnt = np.zeros(self.n, np.int64)

    def alns(self, capi, r):
        """svb_alns_t of the slice with the SFS table of search result r (reads of r = self.searched, ascending)"""
        if getattr(self, "_pinned", None) is None:
            cnt = np.zeros(self.n, np.int64)
            cnt[self.searched] = np.diff(r.offs)
            so = np.zeros(self.n + 1, np.int64)
            np.cumsum(cnt, out=so[1:])
            S = self.S
            return capi.AlnBatch(S["tid"], S["pos"], S["hp"], S["cigar_offs"], S["cigar"], so, r.qs, r.len)
        self._cnt[self.searched] = np.diff(r.offs)
        so = self._so[1]
        so[0] = 0
        np.cumsum(self._cnt, out=so[1:])
        P = self._pinned
        return capi.AlnBatch(P["tid"][1], P["pos"][1], P["hp"][1], P["cigar_offs"][1], P["cigar"][1], so, r.qs, r.len)

    def truth(self):
        """planted SVs at least two searched reads of the slice carry: set of (tid, is_del, len, pos)"""
        sv = self.S["sv_of_read"]
        carried = np.bincount(sv[sv >= 0], minlength=len(self.cat["gpos"]))
        c = self.cat
        return {(int(c["contig"][k]), bool(c["is_del"][k]), int(c["len"][k]), int(c["pos"][k])) for k in np.nonzero(carried >= 2)[0]}


def sv_table(cl, calls):
    """int32 [n_svs, 4]: tid, type (0 INS, 1 DEL), pos, len -- what rank 0 gathers"""
    if calls.n_svs == 0:
        return np.zeros((0, 4), np.int32)
    return np.stack([cl.tid[calls.job_cluster[calls.sv_job]], calls.sv_type.astype(np.int32), calls.sv_pos, calls.sv_len], axis=1).astype(np.int32)


def score_calls(table, truth):
    """planted SVs recovered with exact type and length within len + 2 bases of the anchor; records matching no planted SV"""
    got = [tuple(int(x) for x in row) for row in table]
    hit = sum(any(g[0] == t and bool(g[1]) == d and g[3] == l and abs(g[2] - (p + 1)) <= l + 2 for g in got) for t, d, l, p in truth)
    extra = sum(not any(g[0] == t and bool(g[1]) == d and g[3] == l and abs(g[2] - (p + 1)) <= l + 2 for t, d, l, p in truth) for g in got)
    return {"planted_with_two_carriers": len(truth), "recovered_exact_type_and_length": hit, "records": len(got), "records_matching_no_planted_sv": extra}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from svdss_b200 import build, capi, synth, parallel
    rank, world, local = dist_env()
    build.build_lib()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: libsvdss_b200 has no CPU path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    idx, ref, offs, setup = build_reference_and_index(args, rank, local, torch, capi)
    t0 = time.time()
    sl = Slice(args, rank, local, ref, offs, torch, capi, synth)
    sl.pin(torch)
    dref = idx.ref()
    log("[rank %d] slice: %d records (%d searched = %.1f %%, %.2f Gbases), %d planted SVs genome-wide, generated in %.1fs" %
        (rank, sl.n, sl.ns, 100.0 * sl.ns / sl.n, sl.bases / 1e9, len(sl.cat["gpos"]), time.time() - t0))
    assemble = True
    threads = args.threads

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def timed(fn, steps, warmup, many=False):
        """K steps between two barriers; many=True: fn(n) runs n steps itself (the two-stage pipeline) and returns their results"""
        if many:
            fn(warmup)
        else:
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
        if many:
            res = fn(steps)
        else:
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

    class Step:
        pass

    def cluster_side(r):
        st = Step()
        st.search = r
        t = time.perf_counter()
        st.cl = capi.cluster_batch(sl.alns(capi, r), dref, threads=threads, device=local)
        st.t_cluster = (time.perf_counter() - t) * 1e3
        return st

    def call_side(st, reads, gather):
        t = time.perf_counter()
        st.calls = capi.call_batch(st.cl, reads, dref, device=local)
        st.t_call = (time.perf_counter() - t) * 1e3
        st.table = sv_table(st.cl, st.calls)
        st.gathered, st.t_gather = None, 0.0
        if gather and world > 1:   # the path's only collective: SV records to rank 0 (NCCL gather, int32)
            t = time.perf_counter()
            st.gathered = parallel.gather_rows(st.table, dist, dst=0, device=dev)
            st.t_gather = (time.perf_counter() - t) * 1e3
        return st

    reads_dev = capi.ReadSeqs(sl.reads_t.data_ptr(), sl.seq_offs_dev, capi.SVB_SEQ_NT6, capi.SVB_MEM_DEVICE)
    reads_host = capi.ReadSeqs(sl.host4_np, sl.seq_offs_host4, capi.SVB_SEQ_BAM4, capi.SVB_MEM_HOST)

    def front_resident():
        t = time.perf_counter()
        r = idx.sfs_resident(sl.dreads, assemble=assemble)
        ts = (time.perf_counter() - t) * 1e3
        st = cluster_side(r)
        st.t_search = ts
        return st

    def front_e2e():
        t = time.perf_counter()
        r = idx.sfs_batch_bam4(sl.host4.data_ptr(), sl.seq4_offs, sl.l_qseq, assemble=assemble)
        ts = (time.perf_counter() - t) * 1e3
        st = cluster_side(r)
        st.t_search = ts
        return st

    def step_resident():
        return call_side(front_resident(), reads_dev, gather=False)

    def step_e2e():
        return call_side(front_e2e(), reads_host, gather=True)

    # Two-stage pipeline over consecutive batches: a worker thread searches and clusters batch k + 1 while this thread
    # calls batch k.  The library's stream 0 is the calling thread's own stream (--default-stream per-thread), so the two
    # overlap on the device: k_poa on a config-3 batch is bound by the row chain of its biggest clusters and leaves most
    # SMs idle for most of its time.  n fronts and n backs complete inside the timed region.
    from concurrent.futures import ThreadPoolExecutor
    worker = ThreadPoolExecutor(1)

    def pipelined(front, reads, gather):
        def run(n):
            res = []
            if n <= 0:
                return res
            fut = worker.submit(front)
            for k in range(n):
                st = fut.result()
                if k + 1 < n:
                    fut = worker.submit(front)
                res.append(call_side(st, reads, gather))
            return res
        return run

    # ---- value: searched reads + reference resident in HBM
    def pick_mode(step, pipe):
        """Sequential or pipelined?  Decided BEFORE the timed passes by a short probe of both (max over ranks, so every rank
        decides alike): at 8 ranks on one 32-core host the end-to-end leg loses to the contention for host memory what the
        overlap hides (25.3 vs 28.8 M records/s), on one or two ranks it gains 7-8 %."""
        if args.no_pipeline:
            return False, None
        _, _, w_seq, _ = timed(step, 3, 1)
        _, _, w_pipe, _ = timed(pipe, 3, 1, many=True)
        return w_pipe < w_seq, {"sequential_ms_per_step": w_seq / 3, "pipelined_ms_per_step": w_pipe / 3}

    pipe_res = pipelined(front_resident, reads_dev, False)
    use_pipe, probe = pick_mode(step_resident, pipe_res)
    res_seq, ms_dev_seq, ms_wall_seq, clocks_seq = timed(step_resident, args.steps, args.warmup)
    if not use_pipe:
        res, ms_dev, ms_wall, clocks = res_seq, ms_dev_seq, ms_wall_seq, clocks_seq
    else:
        res, ms_dev, ms_wall, clocks = timed(pipe_res, args.steps, args.warmup, many=True)
    ms_step = ms_wall / args.steps

    def mean(f, rs=res):
        return float(np.mean([f(x) for x in rs]))
    last = res[-1]
    # ---- e2e: host buffers through the C ABI
    pipe_e = pipelined(front_e2e, reads_host, True)
    use_pipe_e, probe_e = pick_mode(step_e2e, pipe_e)
    res_e_seq, ms_dev_e_seq, ms_wall_e_seq, clocks_e_seq = timed(step_e2e, args.steps, args.warmup)
    if not use_pipe_e:
        res_e, ms_dev_e, ms_wall_e, clocks_e = res_e_seq, ms_dev_e_seq, ms_wall_e_seq, clocks_e_seq
    else:
        res_e, ms_dev_e, ms_wall_e, clocks_e = timed(pipe_e, args.steps, args.warmup, many=True)
    ms_step_e = ms_wall_e / args.steps
    for x in res + res_e + res_seq + res_e_seq:      # every step of every pass: the same SV table
        assert np.array_equal(x.table, last.table), "two passes of the step disagree"
    peak, peak_src = hbm_peak()
    r0 = last.search
    search_kernel_ms = mean(lambda x: x.search.kernel_ms)
    blocks, text_ext, n_ext = mean(lambda x: x.search.n_blocks_touched), mean(lambda x: x.search.n_text_ext), mean(lambda x: x.search.n_ext)
    kern = {
        "k_sfs_search_mop (main + tail)": {"ms": search_kernel_ms, "bound": "hbm", "algorithmic_bytes": blocks * idx.block_bytes + 2.0 * text_ext,
                                          "what": "128 B per index block fetched + 2 B per located-match extension"},
        "k_cl_extend + k_cl_fill": {"ms": mean(lambda x: x.cl.kernel_ms), "bound": "hbm", "algorithmic_bytes": None, "what": "per-read CIGAR walks; not a bandwidth kernel"},
        "k_poa": {"ms": mean(lambda x: x.calls.poa_kernel_ms), "bound": "hbm", "algorithmic_bytes": 4.0 * mean(lambda x: x.calls.poa_cells),
                  "what": "one 4-byte traceback word per DP cell; integer-SIMT / latency bound, not a bandwidth kernel", "GCUPS": mean(lambda x: x.calls.poa_cells / max(x.calls.poa_kernel_ms, 1e-6) / 1e6)},
        "k_ksw_extd2": {"ms": mean(lambda x: x.calls.ksw_kernel_ms), "bound": "hbm", "algorithmic_bytes": 1.0 * mean(lambda x: x.calls.ksw_cells),
                        "what": "one traceback byte per DP cell; integer-SIMT bound", "GCUPS": mean(lambda x: x.calls.ksw_cells / max(x.calls.ksw_kernel_ms, 1e-6) / 1e6)},
    }
    for k, v in kern.items():
        v["share_of_step"] = v["ms"] / ms_step
        v["achieved_GB_s"] = None if v["algorithmic_bytes"] is None else v["algorithmic_bytes"] / (v["ms"] * 1e-3) / 1e9
    dom = max(kern, key=lambda k: kern[k]["ms"])
    dk = kern[dom]
    achieved = dk["achieved_GB_s"] or 0.0
    launches = int(sum(x.search.launches + x.cl.launches + x.calls.launches for x in res + res_e))   # the two timed regions the line's value / e2e come from
    stages = {"search_ms": mean(lambda x: x.t_search), "cluster_ms": mean(lambda x: x.t_cluster), "call_ms": mean(lambda x: x.t_call),
              "cluster_host_sweep_ms": mean(lambda x: x.cl.host_ms), "call_host_ms": mean(lambda x: x.calls.host_ms),
              "poa_kernel_ms": kern["k_poa"]["ms"], "ksw_kernel_ms": kern["k_ksw_extd2"]["ms"], "search_kernel_ms": search_kernel_ms,
              "call_gather_ms": mean(lambda x: x.calls.gather_ms), "call_poa_stage_ms": mean(lambda x: x.calls.poa_ms), "call_ksw_stage_ms": mean(lambda x: x.calls.ksw_ms),
              "call_device_ms": mean(lambda x: x.calls.device_ms), "poa_reruns": mean(lambda x: x.calls.poa_reruns)}
    stages_e = {"search_ms": mean(lambda x: x.t_search, res_e), "cluster_ms": mean(lambda x: x.t_cluster, res_e), "call_ms": mean(lambda x: x.t_call, res_e),
                "gather_ms": mean(lambda x: x.t_gather, res_e)}
    h2d = int(res_e[-1].search.h2d_bytes + res_e[-1].cl.h2d_bytes + res_e[-1].calls.h2d_bytes)
    d2h = int(res_e[-1].search.d2h_bytes + res_e[-1].cl.d2h_bytes + res_e[-1].calls.d2h_bytes)
    out = {
        "metric": "SFS-extracted reads/sec (search+call)",
        "value": world * sl.n / (ms_step * 1e-3),
        "unit": "reads/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_step,
        "ms_per_step_cuda_events": ms_dev / args.steps,   # events on the launch stream around the same K steps (every call is synchronous at return: the two agree)
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "int32", "data": "synthetic",
        "config": workload_config(args, world),
        "searched_reads_per_s": world * sl.ns / (ms_step * 1e-3),
        "per_step_rank0": {"records": sl.n, "searched_reads": sl.ns, "searched_bases": sl.bases, "sfs": int(r0.n_sfs), "extensions": int(r0.n_ext),
                           "extended_sfs": int(last.cl.n_extended), "clusters": int(last.cl.n), "clusters_placed": int(last.cl.placed.sum()),
                           "sub_reads": int(last.cl.sub_offs[-1]), "poa_jobs": int(last.calls.n_jobs), "poa_cells": int(last.calls.poa_cells),
                           "ksw_cells": int(last.calls.ksw_cells), "sv_records": int(last.calls.n_svs),
                           "unplaced_sfs": [int(last.cl.unplaced), int(last.cl.s_unplaced), int(last.cl.e_unplaced)]},
        "parity_vs_planted": score_calls(last.table, sl.truth()),
        "stages_ms": stages,
        "pipeline": None if args.no_pipeline else {
            "what": "search + cluster of batch k+1 on a worker thread while this thread calls batch k (stream 0 = per-thread default stream); K of each inside the timed region; stage times are those seen while overlapped",
            "value_pipelined": bool(use_pipe), "e2e_pipelined": bool(use_pipe_e),
            "chosen_by": "a probe of 3 steps of either mode before the timed passes (max over ranks)", "probe_value": probe, "probe_e2e": probe_e,
            "sequential": {"value": world * sl.n / (ms_wall_seq / args.steps * 1e-3), "ms_per_step": ms_wall_seq / args.steps,
                           "stages_ms": {"search_ms": mean(lambda x: x.t_search, res_seq), "cluster_ms": mean(lambda x: x.t_cluster, res_seq), "call_ms": mean(lambda x: x.t_call, res_seq),
                                         "poa_kernel_ms": mean(lambda x: x.calls.poa_kernel_ms, res_seq), "search_kernel_ms": mean(lambda x: x.search.kernel_ms, res_seq)},
                           "e2e_value": world * sl.n / (ms_wall_e_seq / args.steps * 1e-3), "e2e_ms_per_step": ms_wall_e_seq / args.steps}},
        "e2e": {"value": world * sl.n / (ms_step_e * 1e-3), "unit": "reads/s",
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": ms_step_e, "stages_ms": stages_e,
                "searched_reads_per_s": world * sl.ns / (ms_step_e * 1e-3),
                "api": "svb_sfs_batch_bam4 (pinned 4-bit BAM-native reads) -> svb_cluster_batch -> svb_call_batch (sub-reads cut from the same host buffer)"
                       + (" -> NCCL gather of the SV records on rank 0" if world > 1 else ""),
                "host_buffer_numa_node": sl.numa.node, "sv_records_on_rank0": None if res_e[-1].gathered is None else int(len(res_e[-1].gathered))},
        "gpu_launches": launches,
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": ncu_traffic(sl.ns, idx.block_bytes, "k_sfs_search_mop config3") if dom.startswith("k_sfs") else ncu_traffic_plain(dom),
                     "peak_source": peak_src, "kernel": dom, "kernel_ms": dk["ms"], "algorithmic_bytes": dk["algorithmic_bytes"],
                     "algorithmic_bytes_definition": dk["what"], "share_of_step": dk["share_of_step"],
                     "note": "dominant kernel of the step by device time; every kernel of the step is listed under `kernels`"},
        "kernels": kern,
        "clocks": clocks,
        "clocks_e2e": clocks_e,
    }
    if rank == 0 and world == 1:
        del sl.reads_t
        if not args.no_config2:
            try:
                out["search_config2"] = search_config2(args, idx, ref, offs, local, torch, capi, synth, timed, peak, peak_src)
            except Exception as e:
                out["search_config2"] = {"error": "%s: %s" % (type(e).__name__, e)}
        if not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = cpu_baseline(args, idx, ref, offs, local, torch, capi, synth)
            except Exception as e:
                out["cpu_baseline"] = {"error": "%s: %s" % (type(e).__name__, e)}
        del ref
        torch.cuda.empty_cache()
        if not args.no_call_stage:
            # in a child process with a time limit: neither an exception nor a crash or a hang of these kernels may cost the line
            out["call_stage"] = call_stage_child(local, full=args.full_call_stage)
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def workload_config(args, world):
    """the `config` object: identical in both arms (what is measured), arm-specific facts live elsewhere on the line"""
    return {"workload": "configs[2] slice per GPU: search+call on %d records of a %.0fx smoothed-shaped BAM (15 kb reads, XF putative filter) over a "
                        "%.2f Gb reference (%d contigs, both strands indexed), %d planted INS/DEL of 50-5000 bp"
                        % (args.reads, COVERAGE, args.ref_bp / 1e9, args.contigs, args.svs),
            "reads_per_gpu": args.reads, "coverage": COVERAGE, "ref_bp": args.ref_bp, "planted_svs": args.svs,
            "putative_filter": True, "assemble": True, "overlap": -1, "min_cluster_weight": 2, "min_sv_length": 25,
            "parallelism": "BAM slice per GPU x%d, index replicated, gather of SV records on rank 0" % world,
            "l2": "searched reads (~1.3 GB) + index (21 GB) + reference text larger than L2"}


def ncu_traffic_plain(key):
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            return json.load(f)[key]["dram_bytes"]
    except Exception:
        return None


def search_config2(args, idx, ref, offs, local, torch, capi, synth, timed, peak, peak_src):
    """BASELINE.json configs[1]: 1 M unfiltered smoothed-shaped reads through the search alone (batch resident in HBM),
    with the roofline objects of the FMD rank half of the metric."""
    dev = torch.device("cuda", local)
    t0 = time.time()
    segs = synth.make_read_segments(offs, args.config2_reads, seed=4)
    reads_t = synth.materialize_segments_torch(ref, segs)
    read_offs = segs["read_offs"]
    offs_t = torch.from_numpy(read_offs).to(dev)
    n_reads = len(read_offs) - 1
    dreads = capi.DeviceReads(reads_t.data_ptr(), offs_t.data_ptr(), device=local, mem=capi.SVB_MEM_DEVICE, n_reads=n_reads)
    torch.cuda.synchronize(dev)
    log("search_config2: %d reads, %.2f Gbases, generated in %.1fs" % (n_reads, read_offs[-1] / 1e9, time.time() - t0))
    steps = max(3, min(args.steps, 10))
    res, ms_dev, ms_wall, clocks = timed(lambda: idx.sfs_resident(dreads, assemble=True), steps, 3)
    ms_step = ms_dev / steps
    kernel_ms = float(np.mean([r.kernel_ms for r in res]))
    blocks = float(np.mean([r.n_blocks_touched for r in res]))
    n_ext = float(np.mean([r.n_ext for r in res]))
    text_ext = float(np.mean([r.n_text_ext for r in res]))
    alg = blocks * idx.block_bytes + 2.0 * text_ext
    out = {"reads": n_reads, "value": n_reads / (ms_step * 1e-3), "unit": "reads/s", "ms_per_step": ms_step, "steps": steps, "sfs": int(res[-1].n_sfs),
           "roofline": {"bound": "hbm", "achieved": alg / (kernel_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": alg / (kernel_ms * 1e-3) / 1e9 / peak,
                        "traffic": ncu_traffic(n_reads, idx.block_bytes, "k_sfs_search_mop") if args.ref_bp == REF_BP else None, "peak_source": peak_src,
                        "kernel": "k_sfs_search_mop main + tail launch", "kernel_ms": kernel_ms, "algorithmic_bytes": alg, "index_blocks": blocks,
                        "text_extensions": text_ext, "extensions_per_s": n_ext / (kernel_ms * 1e-3)},
           "clocks": clocks}
    # e2e of the search alone from pinned 4-bit host reads (and with the 2-bit transport)
    try:
        l_qseq = np.diff(read_offs).astype(np.int32)
        seq4_offs = np.zeros(n_reads + 1, np.int64)
        seq4_offs[1:] = np.cumsum((l_qseq.astype(np.int64) + 1) // 2)
        s4o_t = torch.from_numpy(seq4_offs).to(dev)
        packed_t = torch.empty(int(seq4_offs[-1]) + 16, dtype=torch.uint8, device=dev)
        capi.pack4_device(reads_t.data_ptr(), offs_t.data_ptr(), s4o_t.data_ptr(), n_reads, packed_t.data_ptr(), device=local)
        host4 = torch.empty(int(seq4_offs[-1]), dtype=torch.uint8, pin_memory=True)
        host4.copy_(packed_t[:int(seq4_offs[-1])])
        torch.cuda.synchronize(dev)
        del packed_t, s4o_t
        for name, env in (("e2e", None), ("e2e_pack2", "1")):
            if env:
                os.environ["SVB_STREAM_PACK2"] = env
            try:
                rs, _, wall, _ = timed(lambda: idx.sfs_batch_bam4(host4.data_ptr(), seq4_offs, l_qseq, assemble=True), 3, 2)
            finally:
                os.environ.pop("SVB_STREAM_PACK2", None)
            assert rs[-1].n_sfs == res[-1].n_sfs, "resident and host paths disagree"
            out[name] = {"value": n_reads / (wall / 3 * 1e-3), "unit": "reads/s", "ms_per_step": wall / 3, "h2d_bytes_per_step": int(rs[-1].h2d_bytes),
                         "d2h_bytes_per_step": int(rs[-1].d2h_bytes),
                         "api": "svb_sfs_batch_bam4, pinned 4-bit reads" + (", re-packed to 2 bits per base on the host chunk by chunk (SVB_STREAM_PACK2=1)" if env else "")}
        del host4
    except Exception as e:
        out["e2e_error"] = "%s: %s" % (type(e).__name__, e)
    # the pure FMD rank walk (every extension = an Occ lookup, SURVEY 8d "FMD rank GB/s"): located-match mode and jump table off
    if not args.no_rank_walk:
        os.environ["SVB_SEARCH_TEXT"] = "0"
        os.environ["SVB_SEARCH_JUMP"] = "0"
        os.environ["SVB_SEARCH_CFG"] = "cpa"     # k_sfs_search_tma<9,1>: the kernel tuned for the pure rank walk
        try:
            idx.sfs_resident(dreads, assemble=True)
            rr = [idx.sfs_resident(dreads, assemble=True) for _ in range(2)]
        finally:
            for k in ("SVB_SEARCH_TEXT", "SVB_SEARCH_JUMP", "SVB_SEARCH_CFG"):
                del os.environ[k]
        assert rr[-1].n_sfs == res[-1].n_sfs and rr[-1].n_ext == res[-1].n_ext, "rank walk and located-match mode disagree"
        rk_ms = float(np.mean([r.kernel_ms for r in rr]))
        rk_bytes = float(np.mean([r.n_blocks_touched for r in rr])) * idx.block_bytes
        out["roofline_rank_walk"] = {
            "bound": "hbm", "achieved": rk_bytes / (rk_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s", "frac": rk_bytes / (rk_ms * 1e-3) / 1e9 / peak,
            "kernel_ms": rk_ms, "algorithmic_bytes": rk_bytes, "reads_per_s": n_reads / (rk_ms * 1e-3), "extensions_per_s": rr[-1].n_ext / (rk_ms * 1e-3),
            "traffic": ncu_traffic(n_reads, idx.block_bytes, "k_sfs_search_tma rank walk") if args.ref_bp == REF_BP else None,
            "kernel": "k_sfs_search_tma<9,1> (SVB_SEARCH_CFG=cpa SVB_SEARCH_TEXT=0 SVB_SEARCH_JUMP=0): every extension is an Occ lookup in a 128 B block"}
        if idx.block_bytes == 128:
            fr = []
            for delta in (1, 1 << 10, 1 << 20):
                ms, blk = idx.rank_bench(1 << 26, delta, seed=7, iters=3)
                gbs = blk * idx.block_bytes / ms / 1e6
                fr.append({"delta": delta, "GB_s": gbs, "frac_of_peak": gbs / peak, "blocks_per_extension": blk / float(1 << 26),
                           "G_extensions_per_s": (1 << 26) / ms / 1e6})
            out["fmd_rank_microbench"] = fr
    return out


def host_pack2_rate(capi, host4_np, seq4_offs, l_qseq, n=60000):
    """How fast this host re-packs 4-bit reads to 2 bits (svb_pack2_host, all threads): GB/s of 4-bit input on
    the first n reads of the batch."""
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


def call_stage_child(local, limit_s=420, full=False):
    """call_stage_sample() in a child process on the same GPU (`bench.py --call-stage-only`)."""
    try:     # device 0 of the child = the only GPU this (world == 1) run uses
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--call-stage-only"] + (["--full-call-stage"] if full else []),
                           capture_output=True, text=True, timeout=limit_s, cwd=ROOT)
    except subprocess.TimeoutExpired:
        return {"error": "call-stage sample did not finish within %s s" % limit_s}
    except Exception as e:
        return {"error": "%s: %s" % (type(e).__name__, e)}
    for line in reversed(r.stdout.splitlines()):
        if line.startswith("{"):
            try:
                return json.loads(line)
            except ValueError:
                break
    return {"error": "child exit %d: %s" % (r.returncode, (r.stderr or r.stdout)[-400:])}


def call_stage_sample(capi, full=False):
    """The `call` kernels on the POA- and ksw2-heavy shapes of SURVEY 8(d): config 4 at full size (50 000 clusters of 20-60
    reads x 200-2000 bp) and config 5 (consensus x window pairs of 100 bp - 10 kb; 200 000 of its 1 M pairs by default, all
    with --full-call-stage), next to the scalar CPU oracle on all host cores on a bounded sample of each."""
    import oracle
    from svdss_b200 import synth
    thr = host_threads()
    seqs, so, co = synth.gen_clusters_fast(50_000, seed=6)
    capi.poa_batch_arrays(seqs[:int(so[co[64]])], so[:co[64] + 1], co[:65])
    r = capi.poa_batch_arrays(seqs, so, co)
    out = {"poa": {"config": "configs[3]: 50 000 clusters of 20-60 reads x 200-2000 bp", "clusters": len(co) - 1, "reads": len(so) - 1, "cells": int(r.cells),
                   "kernel_ms": float(r.kernel_ms), "device_ms": float(r.device_ms), "clusters_per_s": (len(co) - 1) / (r.device_ms * 1e-3),
                   "GCUPS": r.cells / (r.kernel_ms * 1e-3) / 1e9, "launches": int(r.launches), "reruns": int(r.reruns),
                   "api": "svb_poa_batch (host buffers; H2D + kernel + D2H in device_ms)"}}
    m = 96 * thr
    t = time.perf_counter()
    cons = oracle.poa_batch(seqs, so, co[:m + 1], threads=thr)
    dt = time.perf_counter() - t
    ok = all(np.array_equal(cons[k], r.consensus(k)) for k in range(m))
    out["poa"]["cpu_oracle"] = {"clusters_per_s": m / dt, "cores": thr, "sample": "first %d clusters, %.1f s" % (m, dt), "identical_consensus": bool(ok)}
    n_pairs = 1_000_000 if full else 200_000
    qc, qo, tc, to = synth.gen_pairs_fast(n_pairs, seed=7)
    capi.ksw_extd2_batch(qc[:qo[64]], qo[:65], tc[:to[64]], to[:65])
    k = capi.ksw_extd2_batch(qc, qo, tc, to)
    out["ksw2"] = {"config": "configs[4]: consensus x window pairs of 100 bp - 10 kb, %d of 1 000 000" % n_pairs, "pairs": n_pairs, "cells": int(k.cells),
                   "kernel_ms": float(k.kernel_ms), "device_ms": float(k.device_ms), "pairs_per_s": n_pairs / (k.device_ms * 1e-3),
                   "GCUPS": k.cells / (k.kernel_ms * 1e-3) / 1e9, "waves": int(k.waves), "api": "svb_ksw_extd2_batch (host buffers)"}
    m = 24 * thr
    t = time.perf_counter()
    sc, cells = oracle.ksw_batch(qc, qo[:m + 1], tc, to[:m + 1], threads=thr)
    dt = time.perf_counter() - t
    out["ksw2"]["cpu_oracle"] = {"GCUPS": cells / dt / 1e9, "pairs_per_s": m / dt, "cores": thr, "sample": "first %d pairs, %.1f s" % (m, dt),
                                 "identical_scores": bool(np.array_equal(sc, k.score[:m]))}
    return out


class CpuPort:
    """The CPU port of one step (oracle/: FM search port, Clusterer and pcall restatements, OpenMP) on the first `n_records`
    records of rank 0's slice.  Used by the cpu_baseline leg and by --impl reference: same sample, same code."""

    def __init__(self, args, idx, ref, offs, local, torch, capi, synth):
        import oracle
        self.oracle, self.capi = oracle, capi
        self.threads = host_threads()
        t = time.time()
        bwt = idx.bwt()
        self.fm = oracle.FMIndex(bwt)
        del bwt
        log("cpu port: BWT download + CPU block build %.1fs" % (time.time() - t))
        # the reference as host nt6 bytes (3.1 GB) and the sample's reads
        self.n = min(args.reads, args.cpu_records)
        sl = Slice(args, 0, local, ref, offs, torch, capi, synth, n_records=None, want_host=False)
        # the first n records of the slice (the slice generator draws the whole slice; a prefix keeps the same records)
        keep = sl.searched < self.n
        self.searched = sl.searched[keep]
        ns = int(keep.sum())
        self.read_offs = np.ascontiguousarray(sl.read_offs[:ns + 1])
        self.reads = sl.reads_t[:int(self.read_offs[-1])].cpu().numpy()
        self.ref_host = ref.cpu().numpy()
        del sl.reads_t
        torch.cuda.empty_cache()
        S = sl.S
        n = self.n
        self.S = {k: np.ascontiguousarray(S[k][:n]) for k in ("tid", "pos", "hp", "xf", "l_qseq")}
        self.S["cigar_offs"] = np.ascontiguousarray(S["cigar_offs"][:n + 1])
        self.S["cigar"] = np.ascontiguousarray(S["cigar"][:int(S["cigar_offs"][n])])
        self.ref = capi.RefSeqs(self.ref_host, offs[:-1], np.diff(offs), capi.SVB_SEQ_NT6)
        self.seq_offs = np.full(n, -1, np.int64)
        self.seq_offs[self.searched] = self.read_offs[:-1]
        self.rd = capi.ReadSeqs(self.reads, self.seq_offs, capi.SVB_SEQ_NT6)
        self.ns = ns
        sv = S["sv_of_read"][:n]
        carried = np.bincount(sv[sv >= 0], minlength=len(sl.cat["gpos"]))
        c = sl.cat
        self.truth = {(int(c["contig"][k]), bool(c["is_del"][k]), int(c["len"][k]), int(c["pos"][k])) for k in np.nonzero(carried >= 2)[0]}

    def step(self):
        o, capi = self.oracle, self.capi
        t0 = time.perf_counter()
        counts, ooff, qs, ln, ext = self.fm.search_batch(self.reads, self.read_offs, threads=self.threads)
        # Assembler::assemble per read (assembler.cpp:34-56)
        aq, al, acnt = o.assemble_batch(ooff, qs, ln)
        t1 = time.perf_counter()
        cnt = np.zeros(self.n, np.int64)
        cnt[self.searched] = acnt
        so = np.zeros(self.n + 1, np.int64)
        np.cumsum(cnt, out=so[1:])
        S = self.S
        alns = capi.AlnBatch(S["tid"], S["pos"], S["hp"], S["cigar_offs"], S["cigar"], so, aq, al)
        cl = o.cluster(alns, self.ref, threads=4, omp_threads=self.threads)
        t2 = time.perf_counter()
        calls = o.call(cl, self.rd, self.ref, omp_threads=self.threads)
        t3 = time.perf_counter()
        return {"search_s": t1 - t0, "cluster_s": t2 - t1, "call_s": t3 - t2, "total_s": t3 - t0, "extensions": int(ext), "cl": cl, "calls": calls}


def cpu_baseline(args, idx, ref, offs, local, torch, capi, synth):
    port = CpuPort(args, idx, ref, offs, local, torch, capi, synth)
    port.fm.search_batch(port.reads, np.ascontiguousarray(port.read_offs[:min(port.ns, 200) + 1]), threads=port.threads, want_output=False)
    r = port.step()
    tab = sv_table(r["cl"], r["calls"])
    return {"value": port.n / r["total_s"], "unit": "reads/s", "cores": port.threads, "kind": "port",
            "sample": "first %d records of rank 0's slice (%d searched), one step: search %.1f s + cluster %.2f s + call %.1f s"
                      % (port.n, port.ns, r["search_s"], r["cluster_s"], r["call_s"]),
            "searched_reads_per_s": port.ns / r["total_s"], "search_M_ext_per_s": r["extensions"] / r["search_s"] / 1e6,
            "parity_vs_planted": score_calls(tab, port.truth), "sv_records": int(len(tab))}


def run_reference(args):
    """CPU port of the same step on all host cores, a bounded sample of rank 0's slice per step."""
    rank, world, local = dist_env()
    if rank != 0:
        return
    import torch
    from svdss_b200 import build, capi, synth
    build.build_lib()
    torch.cuda.set_device(local)
    # setup (untimed): the 6.2 G-symbol BWT is built on the GPU, then handed to the CPU port -- no CPU builder here does it in minutes
    idx, ref, offs, setup = build_reference_and_index(args, rank, local, torch, capi)
    port = CpuPort(args, idx, ref, offs, local, torch, capi, synth)
    del ref
    idx.close()
    torch.cuda.empty_cache()
    for _ in range(min(args.warmup, 1)):
        port.fm.search_batch(port.reads, np.ascontiguousarray(port.read_offs[:min(port.ns, 200) + 1]), threads=port.threads, want_output=False)
    t0 = time.perf_counter()
    rs = [port.step() for _ in range(args.steps)]
    dt = time.perf_counter() - t0
    ms_step = dt / args.steps * 1e3
    v = port.n / (ms_step * 1e-3)
    tab = sv_table(rs[-1]["cl"], rs[-1]["calls"])
    out = {"impl": "reference", "metric": "SFS-extracted reads/sec (search+call)", "value": v,
           "unit": "reads/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "int32", "data": "synthetic",
           "config": workload_config(args, world),
           "searched_reads_per_s": port.ns / (ms_step * 1e-3),
           "cpu_baseline": {"value": v, "unit": "reads/s", "cores": port.threads, "kind": "port",
                            "sample": "first %d records of rank 0's slice per step (%d searched); search %.1f s + cluster %.2f s + call %.1f s per step"
                                      % (port.n, port.ns, np.mean([r["search_s"] for r in rs]), np.mean([r["cluster_s"] for r in rs]), np.mean([r["call_s"] for r in rs]))},
           "parity_vs_planted": score_calls(tab, port.truth),
           "note": "CPU port of the path (oracle/: FM-index search port, Clusterer and pcall restatements over the oracle's POA and ksw2, OpenMP); the reference "
                   "binary cannot be built offline (ropebwt3 / abPOA / ksw2 / htslib are not vendored); index built on the GPU as untimed setup",
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
    ap.add_argument("--reads", type=int, default=READS_PER_GPU, help="BAM records per GPU and step")
    ap.add_argument("--svs", type=int, default=N_SVS, help="planted SVs genome-wide")
    ap.add_argument("--threads", type=int, default=4, help="--threads of `call` (fixes the order of the clusters only)")
    ap.add_argument("--config2-reads", type=int, default=CONFIG2_READS)
    ap.add_argument("--block-bytes", type=int, default=0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-config2", action="store_true", help="skip the configs[1] search-only measurement")
    ap.add_argument("--no-pipeline", action="store_true", help="steps strictly one after the other (no overlap of batch k+1's search with batch k's call)")
    ap.add_argument("--no-rank-walk", action="store_true", help="skip the extra pure-rank-walk pass")
    ap.add_argument("--no-call-stage", action="store_true", help="skip the POA / ksw2 shapes of configs[3] / [4]")
    ap.add_argument("--full-call-stage", action="store_true", help="configs[4] at its full 1 M pairs (about a minute)")
    ap.add_argument("--cpu-records", type=int, default=77_500, help="records per step of the CPU port (cpu_baseline and --impl reference)")
    ap.add_argument("--call-stage-only", action="store_true", help="print the call_stage object alone (the child of the main run)")
    args = ap.parse_args()
    if args.call_stage_only:
        from svdss_b200 import capi
        try:
            if capi.lib().svb_device_count() < 1:
                raise RuntimeError("no CUDA device: libsvdss_b200 has no CPU fallback")
            obj = call_stage_sample(capi, full=args.full_call_stage)
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
