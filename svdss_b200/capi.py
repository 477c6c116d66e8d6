"""ctypes binding of libsvdss_b200.so -- the same C ABI (include/svdss_b200.h) a C++ host links.

There is no CPU fallback: if the library is missing it is an ImportError-class failure, and every
compute entry returns SVB_ECUDA without a CUDA device, surfaced here as SvbError.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsvdss_b200.so")

SVB_MEM_HOST, SVB_MEM_DEVICE = 0, 1


class SvbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libsvdss_b200 error %d: %s" % (code, msg))
        self.code = code


class IndexInfo(C.Structure):
    _fields_ = [("n", C.c_int64), ("acc", C.c_int64 * 7), ("n_blocks", C.c_int64),
                ("block_bytes", C.c_int32), ("block_syms", C.c_int32), ("n_contigs", C.c_int64),
                ("device_bytes", C.c_int64), ("device", C.c_int32)]


class SfsOut(C.Structure):
    _fields_ = [("n_reads", C.c_int64), ("n_sfs", C.c_int64), ("offs", C.POINTER(C.c_int64)),
                ("qs", C.POINTER(C.c_int32)), ("len", C.POINTER(C.c_int32)),
                ("n_ext", C.c_int64), ("n_blocks_touched", C.c_int64), ("kernel_ms", C.c_float),
                ("device_ms", C.c_float), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("launches", C.c_int32), ("block_bytes", C.c_int32), ("n_text_ext", C.c_int64)]


class KswOut(C.Structure):
    _fields_ = [("n_pairs", C.c_int64), ("score", C.POINTER(C.c_int32)), ("cigar_offs", C.POINTER(C.c_int64)),
                ("cigar", C.POINTER(C.c_uint32)), ("n_cigar", C.c_int64), ("cells", C.c_int64),
                ("kernel_ms", C.c_float), ("device_ms", C.c_float), ("h2d_bytes", C.c_int64),
                ("d2h_bytes", C.c_int64), ("launches", C.c_int32), ("waves", C.c_int32)]


class PoaOut(C.Structure):
    _fields_ = [("n_clusters", C.c_int64), ("cons_offs", C.POINTER(C.c_int64)), ("cons", C.POINTER(C.c_uint8)),
                ("status", C.POINTER(C.c_int32)), ("cells", C.c_int64), ("kernel_ms", C.c_float),
                ("device_ms", C.c_float), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64),
                ("launches", C.c_int32), ("reruns", C.c_int32)]


class Alns(C.Structure):
    _fields_ = [("n_aln", C.c_int64), ("tid", C.c_void_p), ("pos", C.c_void_p), ("hp", C.c_void_p), ("cigar_offs", C.c_void_p),
                ("cigar", C.c_void_p), ("sfs_offs", C.c_void_p), ("sfs_qs", C.c_void_p), ("sfs_len", C.c_void_p)]


class Ref(C.Structure):
    _fields_ = [("n_contigs", C.c_int64), ("seq", C.c_void_p), ("start", C.c_void_p), ("len", C.c_void_p), ("name_rank", C.c_void_p),
                ("fmt", C.c_int), ("mem", C.c_int)]


class ClustersOut(C.Structure):
    _fields_ = [("n_clusters", C.c_int64), ("tid", C.POINTER(C.c_int32)), ("s", C.POINTER(C.c_int32)), ("e", C.POINTER(C.c_int32)),
                ("cov0", C.POINTER(C.c_int32)), ("cov1", C.POINTER(C.c_int32)), ("cov2", C.POINTER(C.c_int32)), ("placed", C.POINTER(C.c_uint8)),
                ("sub_offs", C.POINTER(C.c_int64)), ("sub_aln", C.POINTER(C.c_int32)), ("sub_qs", C.POINTER(C.c_int32)),
                ("sub_qe", C.POINTER(C.c_int32)), ("sub_hp", C.POINTER(C.c_int32)), ("rvec_offs", C.POINTER(C.c_int64)),
                ("rvec", C.POINTER(C.c_uint8)), ("clip", C.POINTER(C.c_int32)),
                ("unplaced", C.c_int64), ("s_unplaced", C.c_int64), ("e_unplaced", C.c_int64), ("unknown", C.c_int64),
                ("unextended", C.c_int64), ("small_clusters", C.c_int64), ("small_clusters_2", C.c_int64), ("n_extended", C.c_int64),
                ("max_ext_len", C.c_int32), ("dist", C.c_int32), ("kernel_ms", C.c_float), ("device_ms", C.c_float), ("host_ms", C.c_float),
                ("launches", C.c_int32), ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64)]


class Seqs(C.Structure):
    _fields_ = [("n", C.c_int64), ("seq", C.c_void_p), ("offs", C.c_void_p), ("fmt", C.c_int), ("mem", C.c_int)]


class CallsOut(C.Structure):
    _fields_ = [("n_jobs", C.c_int64), ("job_cluster", C.POINTER(C.c_int32)), ("job_cov", C.POINTER(C.c_int32)),
                ("job_sub_offs", C.POINTER(C.c_int64)), ("job_sub", C.POINTER(C.c_int32)), ("cons_offs", C.POINTER(C.c_int64)),
                ("cons", C.POINTER(C.c_uint8)), ("score", C.POINTER(C.c_int32)), ("cigar_offs", C.POINTER(C.c_int64)),
                ("cigar", C.POINTER(C.c_uint32)), ("n_svs", C.c_int64), ("sv_job", C.POINTER(C.c_int32)), ("sv_type", C.POINTER(C.c_uint8)),
                ("sv_pos", C.POINTER(C.c_int32)), ("sv_len", C.POINTER(C.c_int32)), ("sv_cpos", C.POINTER(C.c_int32)),
                ("job_nv", C.POINTER(C.c_int32)), ("skipped_outside", C.c_int64), ("poa_cells", C.c_int64), ("ksw_cells", C.c_int64),
                ("poa_kernel_ms", C.c_float), ("ksw_kernel_ms", C.c_float), ("gather_ms", C.c_float), ("device_ms", C.c_float),
                ("host_ms", C.c_float), ("launches", C.c_int32), ("poa_reruns", C.c_int32), ("ksw_waves", C.c_int32),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("poa_ms", C.c_float), ("ksw_ms", C.c_float)]


SVB_SEQ_ASCII, SVB_SEQ_NT6, SVB_SEQ_BAM4 = 0, 1, 2

_lib = None

# every symbol include/svdss_b200.h declares (tests check the .so exports all of them)
EXPORTS = [
    "svb_last_error", "svb_device_count", "svb_version",
    "svb_index_build", "svb_index_from_bwt", "svb_index_load", "svb_index_save", "svb_index_free",
    "svb_index_info", "svb_index_get_bwt", "svb_suffix_array",
    "svb_rank2a", "svb_rank_bench",
    "svb_sfs_batch", "svb_sfs_batch_bam4", "svb_pack4_device", "svb_pack2_host", "svb_pack2_chunk", "svb_unpack2_device", "svb_bgzf_inflate_device", "svb_reads_upload", "svb_reads_free", "svb_sfs_resident", "svb_sfs_out_free",
    "svb_ksw_extd2_batch", "svb_ksw_out_free",
    "svb_poa_batch", "svb_poa_out_free",
    "svb_cluster_batch", "svb_clusters_free", "svb_call_batch", "svb_calls_free", "svb_index_ref", "svb_host_alloc_pinned", "svb_host_free_pinned", "svb_bamstream_open", "svb_bamstream_window", "svb_bamstream_pending_bytes", "svb_bamstream_fetch", "svb_bamstream_search", "svb_bamstream_close",
]


def lib():
    """Load libsvdss_b200.so (built in-tree by svdss_b200/build.py). Fails loudly if absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("libsvdss_b200.so not built: run `python -m svdss_b200.build` "
                      "(the CUDA library is the only implementation; there is no CPU path)")
    L = C.CDLL(LIB_PATH)
    vp, i64, i32 = C.c_void_p, C.c_int64, C.c_int
    L.svb_last_error.restype = C.c_char_p
    L.svb_version.restype = C.c_char_p
    L.svb_device_count.restype = i32
    L.svb_index_build.argtypes = [vp, vp, i64, i32, i32, i32, C.POINTER(vp)]
    L.svb_index_from_bwt.argtypes = [vp, i64, i32, i32, i32, C.POINTER(vp)]
    L.svb_index_load.argtypes = [C.c_char_p, i32, C.POINTER(vp)]
    L.svb_index_save.argtypes = [vp, C.c_char_p]
    L.svb_index_free.argtypes = [vp]
    L.svb_index_free.restype = None
    L.svb_index_info.argtypes = [vp, C.POINTER(IndexInfo)]
    L.svb_index_get_bwt.argtypes = [vp, vp]
    L.svb_suffix_array.argtypes = [vp, i64, i32, i32, vp]
    L.svb_rank2a.argtypes = [vp, vp, vp, i64, vp, vp]
    L.svb_rank_bench.argtypes = [vp, i64, i64, C.c_uint64, i32, C.POINTER(C.c_float), C.POINTER(i64)]
    L.svb_sfs_batch.argtypes = [vp, vp, vp, i64, i32, i32, C.POINTER(SfsOut)]
    L.svb_sfs_batch_bam4.argtypes = [vp, vp, vp, vp, i64, i32, i32, C.POINTER(SfsOut)]
    L.svb_pack4_device.argtypes = [vp, vp, vp, i64, i32, vp]
    L.svb_reads_upload.argtypes = [vp, vp, i64, i32, i32, C.POINTER(vp)]
    L.svb_reads_free.argtypes = [vp]
    L.svb_reads_free.restype = None
    L.svb_sfs_resident.argtypes = [vp, vp, i32, i32, C.POINTER(SfsOut)]
    L.svb_sfs_out_free.argtypes = [C.POINTER(SfsOut)]
    L.svb_sfs_out_free.restype = None
    L.svb_ksw_extd2_batch.argtypes = [vp, vp, vp, vp, i64, i32, i32, i32, i32, i32, i32, i32, i32, C.POINTER(KswOut)]
    L.svb_ksw_out_free.argtypes = [C.POINTER(KswOut)]
    L.svb_ksw_out_free.restype = None
    L.svb_poa_batch.argtypes = [vp, vp, vp, i64, i32, C.POINTER(PoaOut)]
    L.svb_poa_out_free.argtypes = [C.POINTER(PoaOut)]
    L.svb_poa_out_free.restype = None
    L.svb_cluster_batch.argtypes = [C.POINTER(Alns), C.POINTER(Ref), i32, i32, i32, i32, i32, i32, C.POINTER(ClustersOut)]
    L.svb_clusters_free.argtypes = [C.POINTER(ClustersOut)]
    L.svb_clusters_free.restype = None
    L.svb_call_batch.argtypes = [C.POINTER(ClustersOut), C.POINTER(Seqs), C.POINTER(Ref), i32, i32, C.c_float, i32, i32, C.POINTER(CallsOut)]
    L.svb_calls_free.argtypes = [C.POINTER(CallsOut)]
    L.svb_calls_free.restype = None
    L.svb_index_ref.argtypes = [vp, C.POINTER(Ref), vp, vp]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise SvbError(rc, lib().svb_last_error().decode(errors="replace"))


def _ptr(a):
    """numpy array -> void*, int -> raw (device) pointer"""
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return a.ctypes.data_as(C.c_void_p)


class SfsResult:
    """Per-batch search result copied out of a svb_sfs_out_t."""

    def __init__(self, out):
        n, m = out.n_reads, out.n_sfs
        self.offs = np.ctypeslib.as_array(out.offs, shape=(n + 1,)).copy() if n >= 0 and out.offs else np.zeros(1, np.int64)
        self.qs = np.ctypeslib.as_array(out.qs, shape=(m,)).copy() if m else np.zeros(0, np.int32)
        self.len = np.ctypeslib.as_array(out.len, shape=(m,)).copy() if m else np.zeros(0, np.int32)
        self.n_reads, self.n_sfs = n, m
        self.n_ext = out.n_ext
        self.n_blocks_touched = out.n_blocks_touched
        self.kernel_ms = out.kernel_ms
        self.device_ms = out.device_ms
        self.h2d_bytes = out.h2d_bytes
        self.d2h_bytes = out.d2h_bytes
        self.launches = out.launches
        self.block_bytes = out.block_bytes
        self.n_text_ext = out.n_text_ext

    def per_read(self, r):
        a, b = int(self.offs[r]), int(self.offs[r + 1])
        return list(zip(self.qs[a:b].tolist(), self.len[a:b].tolist()))


class Index:
    """Device-resident FMD index (stands in for rb3_fmi_t, ping_pong.cpp:243-245)."""

    def __init__(self, handle):
        self._h = handle
        info = IndexInfo()
        check(lib().svb_index_info(self._h, C.byref(info)))
        self.n = info.n
        self.acc = list(info.acc)
        self.n_blocks = info.n_blocks
        self.block_bytes = info.block_bytes
        self.block_syms = info.block_syms
        self.device_bytes = info.device_bytes
        self.device = info.device

    @classmethod
    def build(cls, contigs_cat, offs, device=0, block_bytes=0, mem=SVB_MEM_HOST):
        """contigs_cat/offs: numpy arrays (host) or raw device pointers with mem=SVB_MEM_DEVICE."""
        h = C.c_void_p()
        n = (len(offs) - 1) if not isinstance(offs, (int, np.integer)) else None
        if n is None:
            raise ValueError("pass n_contigs through build_device()")
        check(lib().svb_index_build(_ptr(contigs_cat), _ptr(offs), n, mem, device, block_bytes, C.byref(h)))
        return cls(h)

    @classmethod
    def build_device(cls, seq_ptr, offs_ptr, n_contigs, device=0, block_bytes=0):
        h = C.c_void_p()
        check(lib().svb_index_build(C.c_void_p(seq_ptr), C.c_void_p(offs_ptr), n_contigs, SVB_MEM_DEVICE,
                                    device, block_bytes, C.byref(h)))
        return cls(h)

    @classmethod
    def from_bwt(cls, bwt, device=0, block_bytes=0):
        bwt = np.ascontiguousarray(bwt, np.uint8)
        h = C.c_void_p()
        check(lib().svb_index_from_bwt(_ptr(bwt), len(bwt), SVB_MEM_HOST, device, block_bytes, C.byref(h)))
        return cls(h)

    @classmethod
    def load(cls, path, device=0):
        h = C.c_void_p()
        check(lib().svb_index_load(os.fsencode(path), device, C.byref(h)))
        return cls(h)

    def save(self, path):
        check(lib().svb_index_save(self._h, os.fsencode(path)))

    def ref(self):
        """the indexed contigs as a device-resident RefSeqs (nt6 codes, forward strands)"""
        info = IndexInfo()
        check(lib().svb_index_info(self._h, C.byref(info)))
        st = np.zeros(info.n_contigs, np.int64)
        ln = np.zeros(info.n_contigs, np.int64)
        r = Ref()
        check(lib().svb_index_ref(self._h, C.byref(r), _ptr(st), _ptr(ln)))
        return RefSeqs(int(r.seq), st, ln, SVB_SEQ_NT6, SVB_MEM_DEVICE)

    def bwt(self):
        out = np.empty(self.n, np.uint8)
        check(lib().svb_index_get_bwt(self._h, _ptr(out)))
        return out

    def rank2a(self, k, l):
        k = np.ascontiguousarray(k, np.int64)
        l = np.ascontiguousarray(l, np.int64)
        ok = np.empty((len(k), 6), np.int64)
        ol = np.empty((len(k), 6), np.int64)
        check(lib().svb_rank2a(self._h, _ptr(k), _ptr(l), len(k), _ptr(ok), _ptr(ol)))
        return ok, ol

    def rank_bench(self, n_queries, delta, seed=1, iters=5):
        ms = C.c_float()
        blk = C.c_int64()
        check(lib().svb_rank_bench(self._h, n_queries, delta, seed, iters, C.byref(ms), C.byref(blk)))
        return ms.value, blk.value

    def sfs_batch(self, reads_cat, offs, overlap=-1, assemble=True):
        """svb_sfs_batch with HOST buffers (numpy)."""
        reads_cat = np.ascontiguousarray(reads_cat, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        out = SfsOut()
        check(lib().svb_sfs_batch(self._h, _ptr(reads_cat), _ptr(offs), len(offs) - 1, overlap,
                                  1 if assemble else 0, C.byref(out)))
        try:
            return SfsResult(out)
        finally:
            lib().svb_sfs_out_free(C.byref(out))

    def sfs_batch_bam4(self, seq4, seq4_offs, l_qseq, overlap=-1, assemble=True):
        """svb_sfs_batch_bam4: reads as BAM stores them (4-bit nt16, one read per byte-aligned run).
        seq4 may be a numpy array or an integer host address (pinned buffer)."""
        seq4_offs = np.ascontiguousarray(seq4_offs, np.int64)
        l_qseq = np.ascontiguousarray(l_qseq, np.int32)
        if isinstance(seq4, np.ndarray):
            seq4 = np.ascontiguousarray(seq4, np.uint8)
            p = _ptr(seq4)
        else:
            p = C.c_void_p(int(seq4))
        out = SfsOut()
        check(lib().svb_sfs_batch_bam4(self._h, p, _ptr(seq4_offs), _ptr(l_qseq), len(l_qseq), overlap,
                                       1 if assemble else 0, C.byref(out)))
        try:
            return SfsResult(out)
        finally:
            lib().svb_sfs_out_free(C.byref(out))

    def sfs_resident(self, reads, overlap=-1, assemble=True):
        out = SfsOut()
        check(lib().svb_sfs_resident(self._h, reads._h, overlap, 1 if assemble else 0, C.byref(out)))
        try:
            return SfsResult(out)
        finally:
            lib().svb_sfs_out_free(C.byref(out))

    def close(self):
        if getattr(self, "_h", None):
            lib().svb_index_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DeviceReads:
    """A batch of reads resident in HBM (svb_reads_t)."""

    def __init__(self, reads_cat, offs, device=0, mem=SVB_MEM_HOST, n_reads=None):
        h = C.c_void_p()
        if mem == SVB_MEM_HOST:
            reads_cat = np.ascontiguousarray(reads_cat, np.uint8)
            offs = np.ascontiguousarray(offs, np.int64)
            n_reads = len(offs) - 1
        self._keep = (reads_cat, offs)
        check(lib().svb_reads_upload(_ptr(reads_cat), _ptr(offs), n_reads, mem, device, C.byref(h)))
        self._h = h
        self.n_reads = n_reads

    def close(self):
        if getattr(self, "_h", None):
            lib().svb_reads_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def suffix_array(text, device=0):
    text = np.ascontiguousarray(text, np.uint8)
    sa = np.empty(len(text), np.int64)
    check(lib().svb_suffix_array(_ptr(text), len(text), SVB_MEM_HOST, device, _ptr(sa)))
    return sa


# caller.cpp:333-349: match +1, mismatch -9, N -> -gape2, gaps min(16+2k, 41+k)
KSW_DEFAULTS = dict(match=1, mismatch=-9, sc_n=-1, gapo=16, gape=2, gapo2=41, gape2=1)


class KswResult:
    def __init__(self, out):
        n = out.n_pairs
        self.n_pairs = n
        self.score = np.ctypeslib.as_array(out.score, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        self.cigar_offs = np.ctypeslib.as_array(out.cigar_offs, shape=(n + 1,)).copy()
        m = out.n_cigar
        self.cigar = np.ctypeslib.as_array(out.cigar, shape=(m,)).copy() if m else np.zeros(0, np.uint32)
        self.cells = out.cells
        self.kernel_ms = out.kernel_ms
        self.device_ms = out.device_ms
        self.h2d_bytes = out.h2d_bytes
        self.d2h_bytes = out.d2h_bytes
        self.launches = out.launches
        self.waves = out.waves

    def cigar_of(self, p):
        a, b = int(self.cigar_offs[p]), int(self.cigar_offs[p + 1])
        return [(int(c >> 4), "MID"[int(c & 0xf)]) for c in self.cigar[a:b]]

    def cigar_string(self, p):
        """the string Caller::pcall builds at caller.cpp:352-355"""
        return "".join("%d%s" % (l, op) for l, op in self.cigar_of(p))


def ksw_extd2_batch(q_cat, q_offs, t_cat, t_offs, device=0, **kw):
    """svb_ksw_extd2_batch with HOST buffers: sequences are _char26_table codes 0..4."""
    p = dict(KSW_DEFAULTS)
    p.update(kw)
    q_cat = np.ascontiguousarray(q_cat, np.uint8)
    t_cat = np.ascontiguousarray(t_cat, np.uint8)
    q_offs = np.ascontiguousarray(q_offs, np.int64)
    t_offs = np.ascontiguousarray(t_offs, np.int64)
    if len(q_offs) != len(t_offs):
        raise ValueError("query and target offset arrays differ in length")
    out = KswOut()
    check(lib().svb_ksw_extd2_batch(_ptr(q_cat), _ptr(q_offs), _ptr(t_cat), _ptr(t_offs), len(q_offs) - 1,
                                    p["match"], p["mismatch"], p["sc_n"], p["gapo"], p["gape"], p["gapo2"],
                                    p["gape2"], device, C.byref(out)))
    try:
        return KswResult(out)
    finally:
        lib().svb_ksw_out_free(C.byref(out))


class PoaResult:
    def __init__(self, out):
        n = out.n_clusters
        self.n_clusters = n
        self.cons_offs = np.ctypeslib.as_array(out.cons_offs, shape=(n + 1,)).copy()
        m = int(self.cons_offs[-1])
        self.cons = np.ctypeslib.as_array(out.cons, shape=(m,)).copy() if m else np.zeros(0, np.uint8)
        self.status = np.ctypeslib.as_array(out.status, shape=(n,)).copy() if n else np.zeros(0, np.int32)
        self.cells = out.cells
        self.kernel_ms = out.kernel_ms
        self.device_ms = out.device_ms
        self.h2d_bytes = out.h2d_bytes
        self.d2h_bytes = out.d2h_bytes
        self.launches = out.launches
        self.reruns = out.reruns

    def consensus(self, c):
        return self.cons[int(self.cons_offs[c]):int(self.cons_offs[c + 1])]

    def consensus_string(self, c):
        """the string Caller::run_poa returns (caller.cpp:295-297)"""
        return "".join("ACGTN"[int(b)] for b in self.consensus(c))


def poa_batch(clusters, device=0):
    """clusters: list of lists of uint8 code arrays (0..4). svb_poa_batch with HOST buffers."""
    seqs = [np.ascontiguousarray(s, np.uint8) for cl in clusters for s in cl]
    seq_offs = np.zeros(len(seqs) + 1, np.int64)
    if seqs:
        seq_offs[1:] = np.cumsum([len(s) for s in seqs])
    cat = np.ascontiguousarray(np.concatenate(seqs)) if seqs and seq_offs[-1] else np.zeros(1, np.uint8)
    cl_offs = np.zeros(len(clusters) + 1, np.int64)
    if clusters:
        cl_offs[1:] = np.cumsum([len(cl) for cl in clusters])
    out = PoaOut()
    check(lib().svb_poa_batch(_ptr(cat), _ptr(seq_offs), _ptr(cl_offs), len(clusters), device, C.byref(out)))
    try:
        return PoaResult(out)
    finally:
        lib().svb_poa_out_free(C.byref(out))


def poa_batch_arrays(cat, seq_offs, cl_offs, device=0):
    """svb_poa_batch on ready arrays: codes 0..4 concatenated, int64 offsets of the sequences and of the clusters"""
    cat = np.ascontiguousarray(cat, np.uint8)
    if len(cat) == 0:
        cat = np.zeros(1, np.uint8)
    seq_offs = np.ascontiguousarray(seq_offs, np.int64)
    cl_offs = np.ascontiguousarray(cl_offs, np.int64)
    out = PoaOut()
    check(lib().svb_poa_batch(_ptr(cat), _ptr(seq_offs), _ptr(cl_offs), len(cl_offs) - 1, device, C.byref(out)))
    try:
        return PoaResult(out)
    finally:
        lib().svb_poa_out_free(C.byref(out))


NT16_OF_NT6 = np.array([15, 1, 2, 4, 8, 15], np.uint8)   # nt6 code -> htslib nt16 code (N for $ / N)


def pack_bam4(reads):
    """Host helper (tests): list of nt6 arrays -> (seq4 bytes, byte offsets, l_qseq) in BAM's layout:
    two bases per byte, first base in the high nibble, every read starting on a byte boundary."""
    l_qseq = np.array([len(r) for r in reads], np.int32)
    offs = np.zeros(len(reads) + 1, np.int64)
    offs[1:] = np.cumsum((l_qseq.astype(np.int64) + 1) // 2)
    out = np.zeros(int(offs[-1]), np.uint8)
    for r, o in zip(reads, offs[:-1]):
        c = NT16_OF_NT6[np.minimum(np.asarray(r, np.uint8), 5)]
        if len(c) & 1:
            c = np.concatenate([c, np.zeros(1, np.uint8)])
        out[o:o + len(c) // 2] = (c[0::2] << 4) | c[1::2]
    return out, offs, l_qseq


def pack2_host(seq4, seq4_offs, l_qseq, threads=0):
    """svb_pack2_host: BAM-native 4-bit reads -> 2 bits per base on the host.  Returns (packed bytes, byte offsets
    [n+1], exception flags [n])."""
    seq4 = np.ascontiguousarray(seq4, np.uint8)
    seq4_offs = np.ascontiguousarray(seq4_offs, np.int64)
    l_qseq = np.ascontiguousarray(l_qseq, np.int32)
    n = len(l_qseq)
    out_offs = np.zeros(n + 1, np.int64)
    out_offs[1:] = np.cumsum((l_qseq.astype(np.int64) + 3) // 4)
    out = np.zeros(max(1, int(out_offs[-1])), np.uint8)
    exc = np.zeros(max(1, n), np.uint8)
    check(lib().svb_pack2_host(_ptr(seq4 if len(seq4) else np.zeros(1, np.uint8)), _ptr(seq4_offs), _ptr(l_qseq if n else np.zeros(1, np.int32)), n,
                               _ptr(out), _ptr(out_offs), _ptr(exc), int(threads)))
    return out[:int(out_offs[-1])], out_offs, exc[:n]


def bgzf_members(raw):
    """Walk the gzip members of a BGZF file image (bytes): the raw-deflate payloads and their ISIZE fields
    (SAM spec 4.1: extra subfield BC = member size - 1).  What a caller of svb_bgzf_inflate_device does first."""
    import struct
    comps, sizes, o = [], [], 0
    while o < len(raw):
        if raw[o:o + 4] != b"\x1f\x8b\x08\x04":
            raise ValueError("not a BGZF member at byte %d" % o)
        xlen = struct.unpack_from("<H", raw, o + 10)[0]
        bsize, x = None, o + 12
        while x + 4 <= o + 12 + xlen:
            sl = struct.unpack_from("<H", raw, x + 2)[0]
            if raw[x:x + 2] == b"BC" and sl == 2:
                bsize = struct.unpack_from("<H", raw, x + 4)[0]
            x += 4 + sl
        if bsize is None:
            raise ValueError("gzip member without a BC subfield at byte %d" % o)
        total = bsize + 1
        comps.append(raw[o + 12 + xlen:o + total - 8])
        sizes.append(struct.unpack_from("<I", raw, o + total - 4)[0])
        o += total
    return comps, sizes


class InflateResult:
    pass


def bgzf_inflate_device(comps, sizes, device=0, check_status=True):
    """svb_bgzf_inflate_device: inflate the members (list of payload bytes, list of ISIZE) on the GPU."""
    n = len(comps)
    io = np.zeros(n + 1, np.int64)
    io[1:] = np.cumsum([len(c) for c in comps])
    oo = np.zeros(n + 1, np.int64)
    oo[1:] = np.cumsum(sizes)
    comp = np.frombuffer(b"".join(comps) + b"\0", np.uint8)
    out = np.zeros(max(1, int(oo[-1])), np.uint8)
    status = np.zeros(max(1, n), np.int32)
    ms = C.c_float(0)
    rc = lib().svb_bgzf_inflate_device(_ptr(comp), _ptr(io), _ptr(oo), C.c_int64(n), C.c_int(device), _ptr(out), _ptr(status), C.byref(ms))
    if check_status:
        check(rc)
    r = InflateResult()
    r.rc, r.out, r.out_offs, r.status, r.kernel_ms = rc, out[:int(oo[-1])], oo, status[:n], float(ms.value)
    return r


def unpack2_device(packed, packed_offs, offs, device=0):
    """svb_unpack2_device: decode a batch packed by pack2_host on the GPU -> nt6 bytes (numpy)."""
    packed = np.ascontiguousarray(packed, np.uint8)
    packed_offs = np.ascontiguousarray(packed_offs, np.int64)
    offs = np.ascontiguousarray(offs, np.int64)
    out = np.zeros(max(1, int(offs[-1])), np.uint8)
    check(lib().svb_unpack2_device(_ptr(packed if len(packed) else np.zeros(1, np.uint8)), _ptr(packed_offs), _ptr(offs), len(offs) - 1, device, _ptr(out)))
    return out[:int(offs[-1])]


def pack4_device(d_seq_ptr, d_offs_ptr, d_seq4_offs_ptr, n_reads, d_out_ptr, device=0):
    check(lib().svb_pack4_device(C.c_void_p(d_seq_ptr), C.c_void_p(d_offs_ptr), C.c_void_p(d_seq4_offs_ptr), n_reads,
                                 device, C.c_void_p(d_out_ptr)))


class AlnBatch:
    """svb_alns_t over numpy arrays (kept alive here): the records Clusterer::run keeps of a BAM, with the SFSs of
    each record's read.  cigars: list of [(len, op char)] or (cigar_offs, cigar u32) arrays."""
    OPS = "MIDNSHP=X"

    def __init__(self, tid, pos, hp, cigar_offs, cigar, sfs_offs, sfs_qs, sfs_len):
        self.tid = np.ascontiguousarray(tid, np.int32); self.pos = np.ascontiguousarray(pos, np.int32)
        self.hp = np.ascontiguousarray(hp, np.int32)
        self.cigar_offs = np.ascontiguousarray(cigar_offs, np.int64); self.cigar = np.ascontiguousarray(cigar, np.uint32)
        self.sfs_offs = np.ascontiguousarray(sfs_offs, np.int64)
        self.sfs_qs = np.ascontiguousarray(sfs_qs, np.int32); self.sfs_len = np.ascontiguousarray(sfs_len, np.int32)
        self.n = len(self.tid)
        self.c = Alns(self.n, *[a.ctypes.data for a in (self.tid, self.pos, self.hp, self.cigar_offs, self.cigar, self.sfs_offs,
                                                          self.sfs_qs, self.sfs_len)])

    @classmethod
    def from_records(cls, records, sfs_by_index):
        """records: dicts with tid, pos, hp (or None), cigar [(len, op)]; sfs_by_index[i] = [(qs, len), ...]"""
        co, cg, so, qs, ln = [0], [], [0], [], []
        for i, r in enumerate(records):
            cg += [(l << 4) | cls.OPS.index(op) for l, op in r["cigar"]]
            co.append(len(cg))
            for q, l in sfs_by_index.get(i, []):
                qs.append(q); ln.append(l)
            so.append(len(qs))
        return cls([r["tid"] for r in records], [r["pos"] for r in records], [r.get("hp") or 0 for r in records], co, cg, so, qs, ln)


class RefSeqs:
    """svb_ref_t over a host numpy byte array (ASCII or nt6) or a device pointer."""

    def __init__(self, seq, start, length, fmt=SVB_SEQ_ASCII, mem=SVB_MEM_HOST, name_rank=None):
        self.seq = seq if isinstance(seq, (int, np.integer)) else np.ascontiguousarray(seq, np.uint8)
        self.start = np.ascontiguousarray(start, np.int64); self.len = np.ascontiguousarray(length, np.int64)
        self.name_rank = None if name_rank is None else np.ascontiguousarray(name_rank, np.int32)
        self.c = Ref(len(self.start), int(self.seq) if isinstance(self.seq, (int, np.integer)) else self.seq.ctypes.data,
                     self.start.ctypes.data, self.len.ctypes.data, None if self.name_rank is None else self.name_rank.ctypes.data, fmt, mem)

    @classmethod
    def from_strings(cls, names, seqs):
        """ASCII chromosomes in tid order; name_rank = order of the names as byte strings"""
        cat = np.frombuffer("".join(seqs).encode(), np.uint8)
        ln = np.array([len(x) for x in seqs], np.int64)
        st = np.concatenate([[0], np.cumsum(ln)[:-1]]).astype(np.int64)
        order = sorted(range(len(names)), key=lambda i: names[i].encode())
        rank = np.empty(len(names), np.int32)
        rank[order] = np.arange(len(names), dtype=np.int32)
        return cls(cat, st, ln, SVB_SEQ_ASCII, SVB_MEM_HOST, rank)


class Clusters:
    """svb_clusters_t copied out."""

    def __init__(self, o):
        n = o.n_clusters

        def arr(p, m, dt):
            return np.ctypeslib.as_array(p, shape=(m,)).copy() if m and p else np.zeros(0, dt)
        self.n = n
        self.tid, self.s, self.e = arr(o.tid, n, np.int32), arr(o.s, n, np.int32), arr(o.e, n, np.int32)
        self.cov0, self.cov1, self.cov2 = arr(o.cov0, n, np.int32), arr(o.cov1, n, np.int32), arr(o.cov2, n, np.int32)
        self.placed = arr(o.placed, n, np.uint8)
        self.sub_offs = np.ctypeslib.as_array(o.sub_offs, shape=(n + 1,)).copy() if o.sub_offs else np.zeros(1, np.int64)
        m = int(self.sub_offs[-1])
        self.sub_aln, self.sub_qs, self.sub_qe, self.sub_hp = (arr(p, m, np.int32) for p in (o.sub_aln, o.sub_qs, o.sub_qe, o.sub_hp))
        self.rvec_offs = np.ctypeslib.as_array(o.rvec_offs, shape=(n + 1,)).copy() if o.rvec_offs else np.zeros(1, np.int64)
        self.rvec = arr(o.rvec, int(self.rvec_offs[-1]), np.uint8)
        self.clip = None
        for k in ("unplaced", "s_unplaced", "e_unplaced", "unknown", "unextended", "small_clusters", "small_clusters_2", "n_extended",
                  "max_ext_len", "dist", "kernel_ms", "device_ms", "host_ms", "launches", "h2d_bytes", "d2h_bytes"):
            setattr(self, k, getattr(o, k))


def cluster_batch(alns, ref, threads=4, min_cluster_weight=2, flank=100, ksize=7, clipped=False, device=0, emul=None):
    """svb_cluster_batch; `emul` = a CDLL of tests/emul/cluster_emul.cpp runs the same per-item code on the CPU (tests only)."""
    o = ClustersOut()
    if emul is not None:
        emul.emul_cluster_batch.argtypes = [C.POINTER(Alns), C.POINTER(Ref), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(ClustersOut)]
        emul.emul_clusters_free.argtypes = [C.POINTER(ClustersOut)]
        rc = emul.emul_cluster_batch(C.byref(alns.c), C.byref(ref.c), threads, min_cluster_weight, flank, ksize, int(clipped), C.byref(o))
        if rc != 0:
            raise SvbError(rc, "emul_cluster_batch")
        free = emul.emul_clusters_free
    else:
        check(lib().svb_cluster_batch(C.byref(alns.c), C.byref(ref.c), threads, min_cluster_weight, flank, ksize, int(clipped), device, C.byref(o)))
        free = lib().svb_clusters_free
    try:
        res = Clusters(o)
        if clipped and o.clip:
            res.clip = np.ctypeslib.as_array(o.clip, shape=(alns.n, 4)).copy()
    finally:
        free(C.byref(o))
    return res


class ReadSeqs:
    """svb_seqs_t: sequence i of the batch = seq[offs[i] ...) (offs < 0: not available).  seq: numpy bytes or a device pointer."""

    def __init__(self, seq, offs, fmt=SVB_SEQ_NT6, mem=SVB_MEM_HOST):
        self.seq = seq if isinstance(seq, (int, np.integer)) else np.ascontiguousarray(seq, np.uint8)
        self.offs = np.ascontiguousarray(offs, np.int64)
        self.c = Seqs(len(self.offs), int(self.seq) if isinstance(self.seq, (int, np.integer)) else self.seq.ctypes.data,
                      self.offs.ctypes.data, fmt, mem)


class Calls:
    """svb_calls_t copied out."""

    def __init__(self, o):
        nj, ns = o.n_jobs, o.n_svs

        def arr(p, m, dt):
            return np.ctypeslib.as_array(p, shape=(m,)).copy() if m and p else np.zeros(0, dt)
        self.n_jobs, self.n_svs = nj, ns
        self.job_cluster = arr(o.job_cluster, nj, np.int32)
        self.job_cov = arr(o.job_cov, nj * 4, np.int32).reshape(-1, 4)
        self.job_sub_offs = np.ctypeslib.as_array(o.job_sub_offs, shape=(nj + 1,)).copy()
        self.job_sub = arr(o.job_sub, int(self.job_sub_offs[-1]), np.int32)
        self.cons_offs = np.ctypeslib.as_array(o.cons_offs, shape=(nj + 1,)).copy()
        self.cons = arr(o.cons, int(self.cons_offs[-1]), np.uint8)
        self.score = arr(o.score, nj, np.int32)
        self.cigar_offs = np.ctypeslib.as_array(o.cigar_offs, shape=(nj + 1,)).copy()
        self.cigar = arr(o.cigar, int(self.cigar_offs[-1]), np.uint32)
        self.sv_job, self.sv_pos, self.sv_len, self.sv_cpos = (arr(p, ns, np.int32) for p in (o.sv_job, o.sv_pos, o.sv_len, o.sv_cpos))
        self.sv_type = arr(o.sv_type, ns, np.uint8)
        self.job_nv = arr(o.job_nv, nj, np.int32)
        for k in ("skipped_outside", "poa_cells", "ksw_cells", "poa_kernel_ms", "ksw_kernel_ms", "gather_ms", "device_ms", "host_ms",
                  "launches", "poa_reruns", "ksw_waves", "h2d_bytes", "d2h_bytes", "poa_ms", "ksw_ms"):
            setattr(self, k, getattr(o, k))

    def consensus(self, j):
        return self.cons[int(self.cons_offs[j]):int(self.cons_offs[j + 1])]

    def cigar_of(self, j):
        return [(int(x) >> 4, "MID"[int(x) & 0xf]) for x in self.cigar[int(self.cigar_offs[j]):int(self.cigar_offs[j + 1])]]


def _clusters_struct(cl):
    """a ClustersOut over the numpy arrays of a Clusters object (kept alive by the returned tuple)"""
    o = ClustersOut()
    keep = []
    o.n_clusters = cl.n
    for name, dt in (("tid", np.int32), ("s", np.int32), ("e", np.int32), ("cov0", np.int32), ("cov1", np.int32), ("cov2", np.int32),
                     ("placed", np.uint8), ("sub_offs", np.int64), ("sub_aln", np.int32), ("sub_qs", np.int32), ("sub_qe", np.int32),
                     ("sub_hp", np.int32), ("rvec_offs", np.int64), ("rvec", np.uint8)):
        a = np.ascontiguousarray(getattr(cl, name), dt)
        if a.size == 0:
            a = np.zeros(1, dt)
        keep.append(a)
        setattr(o, name, a.ctypes.data_as(C.POINTER(np.ctypeslib.as_ctypes_type(dt))))
    return o, keep


def call_batch(clusters, reads, ref, min_cluster_weight=2, min_sv_length=25, min_ratio=0.97, useht=True, device=0):
    """svb_call_batch on a Clusters object (from cluster_batch, or any object with the same array attributes)"""
    co, keep = _clusters_struct(clusters)
    o = CallsOut()
    check(lib().svb_call_batch(C.byref(co), C.byref(reads.c), C.byref(ref.c), min_cluster_weight, min_sv_length, min_ratio, int(useht),
                               device, C.byref(o)))
    try:
        return Calls(o)
    finally:
        lib().svb_calls_free(C.byref(o))
