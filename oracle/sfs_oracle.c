/*
 * oracle/sfs_oracle.c -- CPU restatement of the SVDSS `search` hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under svdss_b200/ may link, import or call this file.
 * Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference leg.
 *
 * PARITY UNPINNED: the reference ships no golden vectors, KATs or fixtures for this path
 * (reference tests/ = a smoke script + Dockerfiles, data not shipped) and its arithmetic lives in
 * ropebwt3 @0ea3919 which is not vendored and cannot be built offline.  What is pinned instead:
 * the SFS result is index-independent (it only depends on "is P[i..j] a substring of some contig or
 * of its reverse complement"), so three independent statements are cross-checked in tests/:
 *   (1) orc_sfs_spec()   : suffix-array *forward narrowing* (no BWT, no Occ, no LF-mapping),
 *   (2) orc_fm_search()  : literal control flow of ping_pong.cpp:4-49 over an FM index (rank on a
 *                          sampled-Occ block array) -- the CPU port that is also the CPU baseline,
 *   (3) tests/ref_model.py: pure-Python literal transcription over a naive *bidirectional* FMD
 *                          (rb3_fmd_set_intv / rb3_fmd_extend semantics) + brute-force `in` tests.
 *
 * Reference lines followed:
 *   ping_pong.cpp:4-49    PingPong::ping_pong_search (control flow, emit order, restart rule)
 *   ping_pong.cpp:36, ping_pong.hpp:38   complement = 5-c for 1..4 else c
 *   ping_pong.hpp:46-52   seq_nt6_table ($=0 A=1 C=2 G=3 T=4 N/other=5)
 *   assembler.cpp:34-56   Assembler::assemble
 *   config.hpp:82         overlap == -1 always
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static inline uint8_t comp6(uint8_t c) { return (c >= 1 && c <= 4) ? (uint8_t)(5 - c) : c; }

/* ---------------------------------------------------------------------------------------------
 * Text model (ropebwt3 build without -R inserts S$ and rc(S)$ for every record):
 *   T = S_0 $ rc(S_0) $ S_1 $ rc(S_1) $ ...      n = 2 * sum(|S_i| + 1)
 * ------------------------------------------------------------------------------------------- */
ORC_API int64_t orc_text_len(const int64_t *offs, int64_t m) { return 2 * (offs[m] - offs[0] + m); }

ORC_API void orc_build_text(const uint8_t *seqs, const int64_t *offs, int64_t m, uint8_t *T) {
  int64_t p = 0;
  for (int64_t r = 0; r < m; ++r) {
    int64_t b = offs[r], e = offs[r + 1];
    for (int64_t i = b; i < e; ++i) T[p++] = seqs[i];
    T[p++] = 0;
    for (int64_t i = e - 1; i >= b; --i) T[p++] = comp6(seqs[i]);
    T[p++] = 0;
  }
}

/* ---------------------------------------------------------------------------------------------
 * Suffix array by comparison sort.  Sentinels are made distinct by text position: two suffixes
 * equal up to and including a '$' are ordered by position (any fixed order among sentinels gives
 * the same answers to '$'-free pattern queries).
 * ------------------------------------------------------------------------------------------- */
static const uint8_t *g_T;
static int64_t g_n;

static inline int haszero64(uint64_t v) {
  return ((v - 0x0101010101010101ULL) & ~v & 0x8080808080808080ULL) != 0;
}

static int suf_cmp(const void *pa, const void *pb) {
  int64_t a = *(const int64_t *)pa, b = *(const int64_t *)pb;
  if (a == b) return 0;
  const uint8_t *T = g_T;
  int64_t n = g_n, i = a, j = b;
  for (;;) {
    if (i + 8 <= n && j + 8 <= n) {
      uint64_t x, y;
      memcpy(&x, T + i, 8);
      memcpy(&y, T + j, 8);
      if (x == y && !haszero64(x)) { i += 8; j += 8; continue; }
    }
    /* byte loop over (at most) this word */
    for (int k = 0; k < 8; ++k) {
      if (i >= n || j >= n) return a < b ? -1 : 1; /* cannot happen: T ends with '$' */
      uint8_t x = T[i], y = T[j];
      if (x != y) return x < y ? -1 : 1;
      if (x == 0) return a < b ? -1 : 1; /* same sentinel column: order by position */
      ++i; ++j;
    }
  }
}

ORC_API void orc_suffix_array(const uint8_t *T, int64_t n, int64_t *SA) {
  for (int64_t i = 0; i < n; ++i) SA[i] = i;
  g_T = T;
  g_n = n;
  qsort(SA, (size_t)n, sizeof(int64_t), suf_cmp);
}

ORC_API void orc_bwt_from_sa(const uint8_t *T, int64_t n, const int64_t *SA, uint8_t *bwt) {
  for (int64_t i = 0; i < n; ++i) bwt[i] = SA[i] ? T[SA[i] - 1] : T[n - 1];
}

/* ---------------------------------------------------------------------------------------------
 * (1) Spec oracle: SFS by suffix-array forward narrowing.
 *
 * occ(W) <=> occ(rc(W)) because T is closed under reverse complement, so
 *   backward phase (ping_pong.cpp:12-22): grow W = P[b..s] to the left == append comp(P[b]) to
 *     rc(W); we narrow the SA interval of rc(W) by one more character at depth d.
 *   forward phase (ping_pong.cpp:28-37): grow W = P[b..e] to the right == append P[e] to W.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int64_t lo, hi; int64_t d; } sa_iv;

static void sa_narrow(const uint8_t *T, const int64_t *SA, sa_iv *iv, uint8_t c) {
  /* suffixes in [lo,hi) share a '$'-free prefix of length d, so SA[i]+d < n */
  int64_t lo = iv->lo, hi = iv->hi, d = iv->d;
  int64_t a = lo, b = hi;
  while (a < b) { int64_t m = (a + b) >> 1; if (T[SA[m] + d] < c) a = m + 1; else b = m; }
  int64_t first = a;
  b = hi;
  while (a < b) { int64_t m = (a + b) >> 1; if (T[SA[m] + d] <= c) a = m + 1; else b = m; }
  iv->lo = first; iv->hi = a; iv->d = d + 1;
}

/* emits (qs,len) pairs in the reference's order (descending qs); returns count (<= cap written) */
ORC_API int64_t orc_sfs_spec(const uint8_t *T, int64_t n, const int64_t *SA, const uint8_t *P,
                             int64_t l, int32_t *out_qs, int32_t *out_len, int64_t cap) {
  int64_t cnt = 0, s = l - 1;
  while (s >= 0) {
    /* b = largest b <= s with !occ(P[b..s]) */
    sa_iv iv = {0, n, 0};
    int64_t b = s;
    for (;;) {
      sa_narrow(T, SA, &iv, comp6(P[b]));
      if (iv.lo == iv.hi) break;
      if (b == 0) return cnt; /* whole prefix P[0..s] occurs: stop (ping_pong.cpp:24-25) */
      --b;
    }
    /* e = smallest e >= b with !occ(P[b..e]) */
    sa_iv fv = {0, n, 0};
    int64_t e = b;
    for (;;) {
      sa_narrow(T, SA, &fv, P[e]);
      if (fv.lo == fv.hi) break;
      ++e;
      if (e >= l) { fprintf(stderr, "orc_sfs_spec: forward phase ran off the read\n"); abort(); }
    }
    if (cnt < cap) { out_qs[cnt] = (int32_t)b; out_len[cnt] = (int32_t)(e - b + 1); }
    ++cnt;
    if (b == 0) break;
    s = e - 1; /* begin = end + overlap, overlap == -1 (ping_pong.cpp:42-47, config.hpp:82) */
  }
  return cnt;
}

/* ---------------------------------------------------------------------------------------------
 * Assembler::assemble (assembler.cpp:34-56): sort by qs, merge runs of overlapping SFSs.
 * in/out as (qs,len); returns number written to out.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int32_t qs, l; } qsl;
static int qsl_cmp(const void *a, const void *b) {
  int32_t x = ((const qsl *)a)->qs, y = ((const qsl *)b)->qs;
  return x < y ? -1 : x > y;
}
ORC_API int64_t orc_assemble(const int32_t *qs, const int32_t *len, int64_t m, int32_t *oqs,
                             int32_t *olen) {
  if (m == 0) return 0;
  qsl *v = (qsl *)malloc(sizeof(qsl) * (size_t)m);
  for (int64_t i = 0; i < m; ++i) { v[i].qs = qs[i]; v[i].l = len[i]; }
  qsort(v, (size_t)m, sizeof(qsl), qsl_cmp);
  int64_t o = 0, i = 0;
  while (i < m) {
    int64_t j;
    for (j = i + 1; j < m; ++j) {
      if (v[j - 1].qs + v[j - 1].l <= v[j].qs) {
        oqs[o] = v[i].qs; olen[o] = v[j - 1].qs + v[j - 1].l - v[i].qs; ++o;
        i = j;
        break;
      }
    }
    if (j == m) {
      oqs[o] = v[i].qs; olen[o] = v[j - 1].qs + v[j - 1].l - v[i].qs; ++o;
      i = j;
    }
  }
  free(v);
  return o;
}

/* ---------------------------------------------------------------------------------------------
 * (2) CPU port: FM index with a sampled-Occ block array + the literal ping-pong loop.
 *
 * Block = 64 BWT symbols: 6 x u64 absolute Occ counts at block start is too fat for cache, so we
 * keep 4 x u64 counts for A,C,G,T in a 64-byte block together with three 64-bit bit-planes of the
 * nt6 code (code = b0 | b1<<1 | b2<<2).  N and $ ranks (needed only when a read contains N) come
 * from a side array of per-block (cntN) -- kept separate so the hot block stays one cache line.
 *   block layout (64 B): u64 cnt[4]; u64 plane[3]; u64 pad
 * This is a *port* (kind "port" in bench.py): same algorithm as the reference's
 * rb3_fmd_extend -> rld_rank2a, simpler (hence faster) rank structure; see BASELINE.md section 3.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
  int64_t n;
  int64_t acc[7];
  int64_t nblk;
  uint64_t *blk;   /* nblk * 8 u64 */
  int64_t *cntN;   /* nblk  (Occ(N) at block start) */
} orc_fm_t;

ORC_API orc_fm_t *orc_fm_build(const uint8_t *bwt, int64_t n) {
  orc_fm_t *f = (orc_fm_t *)calloc(1, sizeof(orc_fm_t));
  f->n = n;
  f->nblk = n / 64 + 1;
  if (posix_memalign((void **)&f->blk, 64, (size_t)f->nblk * 64)) return NULL;
  f->cntN = (int64_t *)malloc(sizeof(int64_t) * (size_t)f->nblk);
  /* pass 1 (parallel): planes + per-block symbol counts (count of A,C,G,T in B[0..3], N in cntN) */
#pragma omp parallel for schedule(static)
  for (int64_t b = 0; b < f->nblk; ++b) {
    uint64_t *B = f->blk + b * 8;
    uint64_t p0 = 0, p1 = 0, p2 = 0;
    int64_t lc[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 64; ++j) {
      int64_t i = b * 64 + j;
      uint8_t s = i < n ? bwt[i] : 7; /* padding code 7 matches no symbol */
      if (i < n) lc[s]++;
      p0 |= (uint64_t)(s & 1) << j;
      p1 |= (uint64_t)((s >> 1) & 1) << j;
      p2 |= (uint64_t)((s >> 2) & 1) << j;
    }
    B[0] = (uint64_t)lc[1]; B[1] = (uint64_t)lc[2]; B[2] = (uint64_t)lc[3]; B[3] = (uint64_t)lc[4];
    B[4] = p0; B[5] = p1; B[6] = p2; B[7] = (uint64_t)lc[0];
    f->cntN[b] = lc[5];
  }
  /* pass 2 (serial): exclusive prefix sums */
  int64_t c[6] = {0, 0, 0, 0, 0, 0};
  for (int64_t b = 0; b < f->nblk; ++b) {
    uint64_t *B = f->blk + b * 8;
    int64_t t;
    for (int q = 0; q < 4; ++q) { t = (int64_t)B[q]; B[q] = (uint64_t)c[q + 1]; c[q + 1] += t; }
    t = f->cntN[b]; f->cntN[b] = c[5]; c[5] += t;
    c[0] += (int64_t)B[7]; B[7] = 0;
  }
  f->acc[0] = 0;
  for (int s = 0; s < 6; ++s) f->acc[s + 1] = f->acc[s] + c[s];
  return f;
}

ORC_API void orc_fm_free(orc_fm_t *f) {
  if (!f) return;
  free(f->blk); free(f->cntN); free(f);
}
ORC_API void orc_fm_acc(const orc_fm_t *f, int64_t *acc7) { memcpy(acc7, f->acc, sizeof(f->acc)); }

static inline int64_t fm_occ(const orc_fm_t *f, int c, int64_t k) {
  int64_t b = k >> 6; int off = (int)(k & 63);
  const uint64_t *B = f->blk + b * 8;
  uint64_t m = ((c & 1) ? B[4] : ~B[4]) & ((c & 2) ? B[5] : ~B[5]) & ((c & 4) ? B[6] : ~B[6]);
  m &= off ? (~0ULL >> (64 - off)) : 0ULL;
  int64_t base;
  if (c >= 1 && c <= 4) base = (int64_t)B[c - 1];
  else if (c == 5) base = f->cntN[b];
  else base = b * 64 - (int64_t)(B[0] + B[1] + B[2] + B[3]) - f->cntN[b]; /* '$' */
  return base + __builtin_popcountll(m);
}

/* rank of all six symbols at k and l: the shape of rb3_fmi_rank2a (ping_pong.cpp:20,35 call it
 * through rb3_fmd_extend).  Exposed for the rank parity test of the CUDA rank kernel. */
ORC_API void orc_fm_rank2a(const orc_fm_t *f, int64_t k, int64_t l, int64_t *ok6, int64_t *ol6) {
  for (int c = 0; c < 6; ++c) { ok6[c] = fm_occ(f, c, k); ol6[c] = fm_occ(f, c, l); }
}

/* literal ping_pong.cpp:4-49 with the unidirectional reading of rb3_fmd_extend (SURVEY 7 "key
 * insight"): every direction switch restarts from set_intv, so only x[0]/size of a plain backward
 * search is ever observed; the forward phase is a backward search of complemented characters. */
static int64_t fm_ping_pong(const orc_fm_t *f, const uint8_t *P, int64_t l, int32_t *oqs,
                            int32_t *olen, int64_t cap, int64_t *n_ext) {
  int64_t cnt = 0, ext = 0;
  int64_t begin = l - 1;
  while (begin >= 0) {
    uint8_t c = P[begin];
    int64_t k = f->acc[c], s = f->acc[c + 1] - f->acc[c];
    while (s != 0 && begin > 0) {
      --begin;
      c = P[begin];
      int64_t ok = fm_occ(f, c, k), ol = fm_occ(f, c, k + s);
      k = f->acc[c] + ok; s = ol - ok; ++ext;
    }
    if (begin == 0 && s != 0) break;
    int64_t end = begin;
    c = comp6(P[end]);
    k = f->acc[c]; s = f->acc[c + 1] - f->acc[c];
    while (s != 0) {
      ++end;
      c = comp6(P[end]);
      int64_t ok = fm_occ(f, c, k), ol = fm_occ(f, c, k + s);
      k = f->acc[c] + ok; s = ol - ok; ++ext;
    }
    if (cnt < cap) { oqs[cnt] = (int32_t)begin; olen[cnt] = (int32_t)(end - begin + 1); }
    ++cnt;
    if (begin == 0) break;
    begin = end - 1;
  }
  if (n_ext) *n_ext += ext;
  return cnt;
}

/* Batch driver, OpenMP over reads (ping_pong.cpp:329-361 partitions reads round-robin over
 * threads; schedule(dynamic) here is at least as good).  Reads must carry a terminating 0 the way
 * ping_pong.cpp:94 stores them only if callers rely on it; this port never reads P[l].
 * Two-pass ABI: counts[r] always receives the true per-read count; (qs,len) are written at
 * out_off[r].. when out_off != NULL.  Returns total extensions performed. */
ORC_API int64_t orc_fm_search_batch(const orc_fm_t *f, const uint8_t *reads, const int64_t *offs,
                                    int64_t n_reads, int threads, int64_t *counts,
                                    const int64_t *out_off, int32_t *oqs, int32_t *olen) {
  int64_t total_ext = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
#endif
#pragma omp parallel for schedule(dynamic, 16) reduction(+ : total_ext)
  for (int64_t r = 0; r < n_reads; ++r) {
    int64_t l = offs[r + 1] - offs[r];
    int64_t ext = 0;
    if (out_off) {
      int64_t cap = out_off[r + 1] - out_off[r];
      counts[r] = fm_ping_pong(f, reads + offs[r], l, oqs + out_off[r], olen + out_off[r], cap, &ext);
    } else {
      counts[r] = fm_ping_pong(f, reads + offs[r], l, NULL, NULL, 0, &ext);
    }
    total_ext += ext;
  }
  return total_ext;
}


/* Single-pass batch driver used by the timed reference arm: every thread appends (read, qs, len)
 * to its own growing buffer (the reference's process_batch also returns per-thread containers,
 * ping_pong.cpp:176-209), then the buffers are stitched in read order.  Caller frees with
 * orc_free().  Returns total extensions; *n_out / *o_read / *o_qs / *o_len receive the table. */
typedef struct { int32_t *r, *q, *l; int64_t n, cap; } tbuf;
static void tbuf_push(tbuf *b, int32_t r, int32_t q, int32_t l) {
  if (b->n == b->cap) {
    b->cap = b->cap ? b->cap * 2 : 4096;
    b->r = (int32_t *)realloc(b->r, sizeof(int32_t) * (size_t)b->cap);
    b->q = (int32_t *)realloc(b->q, sizeof(int32_t) * (size_t)b->cap);
    b->l = (int32_t *)realloc(b->l, sizeof(int32_t) * (size_t)b->cap);
  }
  b->r[b->n] = r; b->q[b->n] = q; b->l[b->n] = l; b->n++;
}

ORC_API int64_t orc_fm_search_batch1(const orc_fm_t *f, const uint8_t *reads, const int64_t *offs,
                                     int64_t n_reads, int threads, int64_t *counts, int64_t *n_out,
                                     int32_t **o_qs, int32_t **o_len) {
  int64_t total_ext = 0;
#ifdef _OPENMP
  if (threads > 0) omp_set_num_threads(threads);
  int nt = omp_get_max_threads();
#else
  int nt = 1;
#endif
  tbuf *bufs = (tbuf *)calloc((size_t)nt, sizeof(tbuf));
  int32_t *owner = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_reads ? n_reads : 1));
  int64_t *start = (int64_t *)malloc(sizeof(int64_t) * (size_t)(n_reads ? n_reads : 1));
#pragma omp parallel reduction(+ : total_ext)
  {
#ifdef _OPENMP
    int me = omp_get_thread_num();
#else
    int me = 0;
#endif
    tbuf *b = &bufs[me];
    int32_t tq[64], tl[64];
#pragma omp for schedule(dynamic, 16)
    for (int64_t r = 0; r < n_reads; ++r) {
      int64_t l = offs[r + 1] - offs[r], ext = 0;
      owner[r] = me; start[r] = b->n;
      int64_t c = fm_ping_pong(f, reads + offs[r], l, tq, tl, 64, &ext);
      if (c <= 64) {
        for (int64_t i = 0; i < c; ++i) tbuf_push(b, (int32_t)r, tq[i], tl[i]);
      } else { /* rare: redo with a big enough scratch */
        int32_t *bq = (int32_t *)malloc(sizeof(int32_t) * (size_t)c * 2);
        int64_t e2 = 0;
        fm_ping_pong(f, reads + offs[r], l, bq, bq + c, c, &e2);
        for (int64_t i = 0; i < c; ++i) tbuf_push(b, (int32_t)r, bq[i], bq[c + i]);
        free(bq);
      }
      counts[r] = c;
      total_ext += ext;
    }
  }
  int64_t tot = 0;
  for (int64_t r = 0; r < n_reads; ++r) tot += counts[r];
  int32_t *q = (int32_t *)malloc(sizeof(int32_t) * (size_t)(tot ? tot : 1));
  int32_t *ln = (int32_t *)malloc(sizeof(int32_t) * (size_t)(tot ? tot : 1));
  int64_t o = 0;
  for (int64_t r = 0; r < n_reads; ++r) {
    tbuf *b = &bufs[owner[r]];
    memcpy(q + o, b->q + start[r], sizeof(int32_t) * (size_t)counts[r]);
    memcpy(ln + o, b->l + start[r], sizeof(int32_t) * (size_t)counts[r]);
    o += counts[r];
  }
  for (int t = 0; t < nt; ++t) { free(bufs[t].r); free(bufs[t].q); free(bufs[t].l); }
  free(bufs); free(owner); free(start);
  *n_out = tot; *o_qs = q; *o_len = ln;
  return total_ext;
}

ORC_API void orc_free(void *p) { free(p); }

ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
