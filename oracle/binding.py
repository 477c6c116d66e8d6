"""ctypes binding of oracle/_build/liboracle.so (built by oracle/Makefile)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")
_lib = None

_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")


def build(force=False):
    """Compile the C oracle (gcc). Idempotent."""
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".c")]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in srcs)):
        return _LIB_PATH
    subprocess.check_call(["make", "-s", "-C", _HERE])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        L.orc_text_len.restype = C.c_int64
        L.orc_text_len.argtypes = [_i64p, C.c_int64]
        L.orc_build_text.restype = None
        L.orc_build_text.argtypes = [_u8p, _i64p, C.c_int64, _u8p]
        L.orc_suffix_array.restype = None
        L.orc_suffix_array.argtypes = [_u8p, C.c_int64, _i64p]
        L.orc_bwt_from_sa.restype = None
        L.orc_bwt_from_sa.argtypes = [_u8p, C.c_int64, _i64p, _u8p]
        L.orc_sfs_spec.restype = C.c_int64
        L.orc_sfs_spec.argtypes = [_u8p, C.c_int64, _i64p, _u8p, C.c_int64, _i32p, _i32p, C.c_int64]
        L.orc_assemble.restype = C.c_int64
        L.orc_assemble.argtypes = [_i32p, _i32p, C.c_int64, _i32p, _i32p]
        L.orc_fm_build.restype = C.c_void_p
        L.orc_fm_build.argtypes = [_u8p, C.c_int64]
        L.orc_fm_free.restype = None
        L.orc_fm_free.argtypes = [C.c_void_p]
        L.orc_fm_acc.restype = None
        L.orc_fm_acc.argtypes = [C.c_void_p, _i64p]
        L.orc_fm_rank2a.restype = None
        L.orc_fm_rank2a.argtypes = [C.c_void_p, C.c_int64, C.c_int64, _i64p, _i64p]
        L.orc_fm_search_batch.restype = C.c_int64
        L.orc_fm_search_batch.argtypes = [C.c_void_p, _u8p, _i64p, C.c_int64, C.c_int,
                                          _i64p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_max_threads.restype = C.c_int
        L.orc_fm_search_batch1.restype = C.c_int64
        L.orc_fm_search_batch1.argtypes = [C.c_void_p, _u8p, _i64p, C.c_int64, C.c_int, _i64p,
                                           C.POINTER(C.c_int64), C.POINTER(C.POINTER(C.c_int32)),
                                           C.POINTER(C.POINTER(C.c_int32))]
        L.orc_free.restype = None
        L.orc_free.argtypes = [C.c_void_p]
        _lib = L
    return _lib


NT6 = np.full(256, 5, np.uint8)  # ping_pong.hpp:46-52 (index 0 -> 0, everything else non-ACGT -> 5)
NT6[0] = 0
for _ch, _v in (("A", 1), ("C", 2), ("G", 3), ("T", 4)):
    NT6[ord(_ch)] = _v
    NT6[ord(_ch.lower())] = _v


def encode_nt6(s):
    if isinstance(s, str):
        s = s.encode()
    return NT6[np.frombuffer(s, np.uint8)]


def concat(seqs):
    """list of uint8 arrays -> (concat, offs int64[n+1])"""
    offs = np.zeros(len(seqs) + 1, np.int64)
    if seqs:
        offs[1:] = np.cumsum([len(s) for s in seqs])
        cat = np.ascontiguousarray(np.concatenate(seqs).astype(np.uint8)) if offs[-1] else np.zeros(0, np.uint8)
    else:
        cat = np.zeros(0, np.uint8)
    return cat, offs


def build_text(contigs):
    """contigs: list of nt6 uint8 arrays -> T = S$rc(S)$..."""
    cat, offs = concat(contigs)
    n = lib().orc_text_len(offs, len(contigs))
    T = np.empty(n, np.uint8)
    lib().orc_build_text(cat if len(cat) else np.zeros(1, np.uint8), offs, len(contigs), T)
    return T


def suffix_array(T):
    SA = np.empty(len(T), np.int64)
    lib().orc_suffix_array(np.ascontiguousarray(T), len(T), SA)
    return SA


def bwt_from_sa(T, SA):
    bwt = np.empty(len(T), np.uint8)
    lib().orc_bwt_from_sa(np.ascontiguousarray(T), len(T), SA, bwt)
    return bwt


def sfs_spec(T, SA, P, cap=None):
    """(qs,len) pairs in the reference's emit order (descending qs) for one nt6 read P."""
    P = np.ascontiguousarray(P, np.uint8)
    cap = cap or max(16, len(P))
    qs = np.empty(cap, np.int32)
    ln = np.empty(cap, np.int32)
    c = lib().orc_sfs_spec(T, len(T), SA, P if len(P) else np.zeros(1, np.uint8), len(P), qs, ln, cap)
    assert c <= cap
    return list(zip(qs[:c].tolist(), ln[:c].tolist()))


def assemble(pairs):
    if not pairs:
        return []
    qs = np.array([p[0] for p in pairs], np.int32)
    ln = np.array([p[1] for p in pairs], np.int32)
    oq = np.empty_like(qs)
    ol = np.empty_like(ln)
    c = lib().orc_assemble(qs, ln, len(pairs), oq, ol)
    return list(zip(oq[:c].tolist(), ol[:c].tolist()))


class FMIndex:
    """CPU port of the FM index + ping-pong search (oracle statement #2; also the CPU baseline)."""

    def __init__(self, bwt):
        bwt = np.ascontiguousarray(bwt, np.uint8)
        self.n = len(bwt)
        self._h = lib().orc_fm_build(bwt, self.n)
        if not self._h:
            raise MemoryError("orc_fm_build")
        self.acc = np.empty(7, np.int64)
        lib().orc_fm_acc(self._h, self.acc)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_fm_free(self._h)
            self._h = None

    def rank2a(self, k, l):
        ok = np.empty(6, np.int64)
        ol = np.empty(6, np.int64)
        lib().orc_fm_rank2a(self._h, int(k), int(l), ok, ol)
        return ok, ol

    def search_batch(self, reads, offs, threads=0, want_output=True):
        """returns (counts int64[n], out_off int64[n+1], qs int32[], len int32[], n_extensions)"""
        reads = np.ascontiguousarray(reads, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        n = len(offs) - 1
        counts = np.zeros(n, np.int64)
        if len(reads) == 0:
            reads = np.zeros(1, np.uint8)
        ext = lib().orc_fm_search_batch(self._h, reads, offs, n, threads, counts, None, None, None)
        if not want_output:
            return counts, None, None, None, ext
        return self.search_batch1(reads, offs, threads)

    def search_batch1(self, reads, offs, threads=0):
        """single pass (counts + SFS table), what the timed reference arm runs"""
        reads = np.ascontiguousarray(reads, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        n = len(offs) - 1
        counts = np.zeros(n, np.int64)
        if len(reads) == 0:
            reads = np.zeros(1, np.uint8)
        n_out = C.c_int64(0)
        pq = C.POINTER(C.c_int32)()
        pl = C.POINTER(C.c_int32)()
        ext = lib().orc_fm_search_batch1(self._h, reads, offs, n, threads, counts, C.byref(n_out), C.byref(pq), C.byref(pl))
        tot = n_out.value
        qs = np.ctypeslib.as_array(pq, shape=(max(tot, 1),))[:tot].copy()
        ln = np.ctypeslib.as_array(pl, shape=(max(tot, 1),))[:tot].copy()
        lib().orc_free(pq)
        lib().orc_free(pl)
        out_off = np.zeros(n + 1, np.int64)
        out_off[1:] = np.cumsum(counts)
        return counts, out_off, qs, ln, ext

    def search_batch2(self, reads, offs, threads=0):
        """two-pass variant (count, then fill caller-sized arrays)"""
        reads = np.ascontiguousarray(reads, np.uint8)
        offs = np.ascontiguousarray(offs, np.int64)
        n = len(offs) - 1
        counts = np.zeros(n, np.int64)
        if len(reads) == 0:
            reads = np.zeros(1, np.uint8)
        ext = lib().orc_fm_search_batch(self._h, reads, offs, n, threads, counts, None, None, None)
        out_off = np.zeros(n + 1, np.int64)
        out_off[1:] = np.cumsum(counts)
        tot = int(out_off[-1])
        qs = np.empty(max(tot, 1), np.int32)
        ln = np.empty(max(tot, 1), np.int32)
        c2 = np.zeros(n, np.int64)
        lib().orc_fm_search_batch(self._h, reads, offs, n, threads, c2,
                                  out_off.ctypes.data_as(C.c_void_p), qs.ctypes.data_as(C.c_void_p),
                                  ln.ctypes.data_as(C.c_void_p))
        assert (c2 == counts).all()
        return counts, out_off, qs[:tot], ln[:tot], ext


def max_threads():
    return lib().orc_max_threads()


# ------------------------------------------------------------------ ksw2 extd2 (oracle/ksw_oracle.c)
KSW_NEG_INF = -0x40000000
KSW_PARAMS = dict(a=1, b=-9, sc_n=-1, q=16, e=2, q2=41, e2=1)   # caller.cpp:333-349
_u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")

CHAR26 = np.full(256, 4, np.uint8)   # caller.hpp:25-37 (_char26_table)
for _i, _v in ((0, 0), (1, 1), (2, 2), (3, 3)):
    CHAR26[_i] = _v
for _ch, _v in (("A", 0), ("C", 1), ("G", 2), ("T", 3), ("U", 3)):
    CHAR26[ord(_ch)] = _v
    CHAR26[ord(_ch.lower())] = _v


def _ksw_lib():
    L = lib()
    if not hasattr(L, "_ksw_ready"):
        ci = C.c_int
        L.orc_ksw_extd2.restype = ci
        L.orc_ksw_extd2.argtypes = [ci, _u8p, ci, _u8p, ci, ci, ci, ci, ci, ci, ci, _u32p, ci, C.POINTER(ci)]
        L.orc_affine2_score.restype = ci
        L.orc_affine2_score.argtypes = [ci, _u8p, ci, _u8p, ci, ci, ci, ci, ci, ci, ci]
        L.orc_cigar_score.restype = ci
        L.orc_cigar_score.argtypes = [ci, _u8p, ci, _u8p, ci, ci, ci, ci, ci, ci, ci, _u32p, ci]
        L._ksw_ready = True
    return L


def _z(a):
    a = np.ascontiguousarray(a, np.uint8)
    return a if len(a) else np.zeros(1, np.uint8)


def ksw_extd2(query, target, **kw):
    """(score, cigar list of (len, op) with op in 'MID') as ksw_extd2_sse(flag=0,w=-1) would return."""
    p = dict(KSW_PARAMS); p.update(kw)
    cap = len(query) + len(target) + 2
    cig = np.zeros(cap, np.uint32)
    n = C.c_int(0)
    sc = _ksw_lib().orc_ksw_extd2(len(query), _z(query), len(target), _z(target), p["a"], p["b"], p["sc_n"],
                                  p["q"], p["e"], p["q2"], p["e2"], cig, cap, C.byref(n))
    return sc, [(int(c >> 4), "MID"[int(c & 0xf)]) for c in cig[:n.value]]


def affine2_score(query, target, **kw):
    p = dict(KSW_PARAMS); p.update(kw)
    return _ksw_lib().orc_affine2_score(len(query), _z(query), len(target), _z(target), p["a"], p["b"], p["sc_n"],
                                        p["q"], p["e"], p["q2"], p["e2"])


def cigar_score(query, target, cigar, **kw):
    p = dict(KSW_PARAMS); p.update(kw)
    cig = np.array([(l << 4) | "MID".index(op) for l, op in cigar], np.uint32)
    if len(cig) == 0:
        cig = np.zeros(1, np.uint32)
    return _ksw_lib().orc_cigar_score(len(query), _z(query), len(target), _z(target), p["a"], p["b"], p["sc_n"],
                                      p["q"], p["e"], p["q2"], p["e2"], cig, len(cigar))


# ------------------------------------------------------------------ POA (oracle/poa_oracle.c)
POA_PARAMS = dict(match=2, mismatch=4, o1=4, e1=2, o2=24, e2=1, wb=10, wf=0.01)   # abpoa_init_para defaults


def _poa_lib():
    L = lib()
    if not hasattr(L, "_poa_ready"):
        ci = C.c_int
        L.orc_poa.restype = ci
        L.orc_poa.argtypes = [_u8p, _i64p, ci, ci, ci, ci, ci, ci, ci, ci, ci, C.c_double, _u8p, ci, C.c_void_p]
        L.orc_edit_distance.restype = ci
        L.orc_edit_distance.argtypes = [_u8p, ci, _u8p, ci]
        L._poa_ready = True
    return L


def poa_consensus(seqs, band=False, return_stats=False, **kw):
    """consensus (codes 0..4) of a cluster's sequences added in input order (Caller::run_poa)."""
    p = dict(POA_PARAMS); p.update(kw)
    cat, offs = concat([np.ascontiguousarray(s, np.uint8) for s in seqs])
    cap = int(2 * max([len(s) for s in seqs] + [1]) + 64)
    out = np.zeros(cap, np.uint8)
    stats = np.zeros(3, np.int64)
    n = _poa_lib().orc_poa(_z(cat), offs, len(seqs), 1 if band else 0, p["match"], p["mismatch"], p["o1"], p["e1"],
                           p["o2"], p["e2"], p["wb"], p["wf"], out, cap, stats.ctypes.data_as(C.c_void_p))
    assert n <= cap
    return (out[:n].copy(), stats) if return_stats else out[:n].copy()


def edit_distance(a, b):
    return _poa_lib().orc_edit_distance(_z(a), len(a), _z(b), len(b))


# ------------------------------------------------------------------ Clusterer + pcall (oracle/call_oracle.c)
def _call_lib():
    L = lib()
    if not hasattr(L, "_call_ready"):
        from svdss_b200 import capi as K      # struct layouts of include/svdss_b200.h only; the library is not loaded
        ci = C.c_int
        L.orc_cluster.restype = ci
        L.orc_cluster.argtypes = [C.POINTER(K.Alns), C.POINTER(K.Ref), ci, ci, ci, ci, ci, ci, C.POINTER(K.ClustersOut)]
        L.orc_clusters_free.restype = None
        L.orc_clusters_free.argtypes = [C.POINTER(K.ClustersOut)]
        L.orc_call.restype = ci
        L.orc_call.argtypes = [C.POINTER(K.ClustersOut), C.POINTER(K.Seqs), C.POINTER(K.Ref), ci, ci, C.c_float, ci, ci, C.POINTER(K.CallsOut)]
        L.orc_calls_free.restype = None
        L.orc_calls_free.argtypes = [C.POINTER(K.CallsOut)]
        L._call_ready = True
    return L


def cluster(alns, ref, threads=4, min_cluster_weight=2, flank=100, ksize=7, clipped=False, omp_threads=0):
    """Clusterer::run on the CPU (literal restatement over aligned-pair vectors); alns / ref: capi.AlnBatch / capi.RefSeqs (host)"""
    from svdss_b200 import capi as K
    L = _call_lib()
    o = K.ClustersOut()
    rc = L.orc_cluster(C.byref(alns.c), C.byref(ref.c), threads, min_cluster_weight, flank, ksize, int(clipped), omp_threads, C.byref(o))
    assert rc == 0
    try:
        res = K.Clusters(o)
        if clipped and o.clip:
            res.clip = np.ctypeslib.as_array(o.clip, shape=(alns.n, 4)).copy()
    finally:
        L.orc_clusters_free(C.byref(o))
    return res


def call(clusters, reads, ref, min_cluster_weight=2, min_sv_length=25, min_ratio=0.97, useht=True, omp_threads=0):
    """Caller::pcall on the CPU over the oracle's POA and ksw2 (OpenMP over jobs); host buffers only"""
    from svdss_b200 import capi as K
    L = _call_lib()
    co, keep = K._clusters_struct(clusters)
    o = K.CallsOut()
    rc = L.orc_call(C.byref(co), C.byref(reads.c), C.byref(ref.c), min_cluster_weight, min_sv_length, min_ratio, int(useht), omp_threads, C.byref(o))
    assert rc == 0
    try:
        return K.Calls(o)
    finally:
        L.orc_calls_free(C.byref(o))


def _batch_lib():
    L = _call_lib()
    if not hasattr(L, "_batch_ready"):
        ci, i64 = C.c_int, C.c_int64
        L.orc_poa_batch.restype = i64
        L.orc_poa_batch.argtypes = [_u8p, _i64p, _i64p, i64, ci, _u8p, _i64p, _i32p]
        L.orc_ksw_batch.restype = i64
        L.orc_ksw_batch.argtypes = [_u8p, _i64p, _u8p, _i64p, i64, ci, _i32p]
        L.orc_assemble_batch.restype = i64
        L.orc_assemble_batch.argtypes = [_i64p, _i32p, _i32p, i64, _i32p, _i32p, _i64p]
        L._batch_ready = True
    return L


def poa_batch(seqs, seq_offs, cluster_offs, threads=0):
    """banded POA consensus of every cluster (orc_poa under OpenMP): list of code arrays"""
    seq_offs = np.ascontiguousarray(seq_offs, np.int64)
    cluster_offs = np.ascontiguousarray(cluster_offs, np.int64)
    n = len(cluster_offs) - 1
    lens = np.diff(seq_offs)
    cap = np.zeros(n + 1, np.int64)
    for c in range(n):
        a, b = int(cluster_offs[c]), int(cluster_offs[c + 1])
        cap[c + 1] = cap[c] + 2 * int(lens[a:b].max() if b > a else 0) + 64
    cons = np.zeros(int(cap[-1]) + 1, np.uint8)
    ln = np.zeros(n + 1, np.int32)
    _batch_lib().orc_poa_batch(_z(seqs), seq_offs, cluster_offs, n, threads, cons, cap, ln)
    return [cons[int(cap[c]):int(cap[c]) + int(ln[c])].copy() for c in range(n)]


def ksw_batch(q, q_offs, t, t_offs, threads=0):
    """scores of orc_ksw_extd2 for every pair under OpenMP; returns (scores int32, cells)"""
    q_offs = np.ascontiguousarray(q_offs, np.int64)
    t_offs = np.ascontiguousarray(t_offs, np.int64)
    n = len(q_offs) - 1
    sc = np.zeros(n + 1, np.int32)
    cells = _batch_lib().orc_ksw_batch(_z(q), q_offs, _z(t), t_offs, n, threads, sc)
    return sc[:n], int(cells)


def assemble_batch(offs, qs, ln):
    """Assembler::assemble for every read of a search_batch result (records in emit order, descending qs).
    Returns (qs, len, count per read) with ascending qs inside a read."""
    offs = np.ascontiguousarray(offs, np.int64)
    n = len(offs) - 1
    qs = np.ascontiguousarray(qs, np.int32)
    ln = np.ascontiguousarray(ln, np.int32)
    oq = np.zeros(len(qs) + 1, np.int32)
    ol = np.zeros(len(qs) + 1, np.int32)
    cnt = np.zeros(n + 1, np.int64)
    m = _batch_lib().orc_assemble_batch(offs, _z32(qs), _z32(ln), n, oq, ol, cnt)
    return oq[:m].copy(), ol[:m].copy(), cnt[:n].copy()


def _z32(a):
    return a if len(a) else np.zeros(1, np.int32)
