/* CPU restatement (TEST INFRASTRUCTURE: checker and CPU baseline only, never on the product path) of the `call` side
 * of the hot path, literal where the reference is literal:
 *   orc_cluster  Clusterer::run        clusterer.cpp:8-52  (extend_alignment :156-345 over materialised aligned pairs as
 *                                      bam.cpp:92-134 builds them, get_unique_kmers :350-403, cluster_by_proximity
 *                                      :405-475, fill_clusters :478-610), OpenMP over reads / clusters like the reference
 *   orc_call     Caller::pcall         caller.cpp:311-406  (split_cluster :100-255, split_cluster_by_len :78-97,
 *                                      run_poa -> orc_poa, ksw_extd2_sse -> orc_ksw_extd2, CIGAR walk :359-401),
 *                                      OpenMP schedule(static, 1) over clusters (:312)
 * Inputs and outputs use the plain-C structs of include/svdss_b200.h (layout only; nothing of the library is linked).
 * Parity pin: tests/test_oracle_call.py holds both against the literal Python transcriptions tests/cluster_model.py /
 * tests/call_model.py; the reference itself ships no vectors for this path (parity unpinned). */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/svdss_b200.h"

#define ORC_API __attribute__((visibility("default")))

int orc_poa(const uint8_t *seqs, const int64_t *offs, int n, int band, int match, int mismatch, int o1, int e1, int o2, int e2,
            int wb, double wf, uint8_t *cons, int cap, int64_t *stats);
int orc_ksw_extd2(int ql, const uint8_t *query, int tl, const uint8_t *target, int a, int b, int sc_n, int q, int e, int q2, int e2,
                  uint32_t *cigar, int cap, int *n_cigar);

typedef struct { int q, r; } pair_t;

/* bam.cpp:92-134 */
static pair_t *aligned_pairs(const uint32_t *cig, int n_cig, int pos, int *n_out) {
  int n = 0;
  for (int k = 0; k < n_cig; ++k) { int op = cig[k] & 0xf; if (op != 5 && op != 6) n += (int)(cig[k] >> 4); }
  pair_t *p = (pair_t *)malloc(sizeof(pair_t) * (size_t)(n + 1));
  int ref = pos, rd = 0, m = 0;
  for (int k = 0; k < n_cig; ++k) {
    int op = cig[k] & 0xf, len = (int)(cig[k] >> 4);
    if (op == 0 || op == 7 || op == 8) for (int i = 0; i < len; ++i) { p[m].q = rd++; p[m].r = ref++; ++m; }
    else if (op == 1 || op == 4) for (int i = 0; i < len; ++i) { p[m].q = rd++; p[m].r = -1; ++m; }
    else if (op == 2 || op == 3) for (int i = 0; i < len; ++i) { p[m].q = -1; p[m].r = ref++; ++m; }
  }
  *n_out = m;
  return p;
}

static int endpos_of(const uint32_t *cig, int n_cig, int pos) {
  int span = 0;
  for (int k = 0; k < n_cig; ++k) { int op = cig[k] & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += (int)(cig[k] >> 4); }
  return pos + (span ? span : 1);
}

/* the k bytes at chrom + r, never past the terminator (a C string in the reference) */
static void kmer_at(const uint8_t *chrom, int64_t len, int r, int k, uint8_t *out) {
  memset(out, 0, 8);
  for (int i = 0; i < k; ++i) { int64_t p = (int64_t)r + i; if (p >= 0 && p < len) out[i] = chrom[p]; else break; }
}

/* clusterer.cpp:350-403 */
static pair_t unique_kmer(const pair_t *al, int n, int k, int from_end, const uint8_t *chrom, int64_t clen) {
  pair_t last = {-1, -1};
  if (n < k) return last;
  uint8_t(*km)[8] = (uint8_t(*)[8])malloc((size_t)(n + 1) * 8);
  int nk = 0, i = 0;
  while (i < n - k + 1) {
    int skip = 0;
    for (int j = i; j < i + k; ++j) if (al[j].q == -1 || al[j].r == -1) { skip = 1; i = j + 1; break; }
    if (skip) continue;
    kmer_at(chrom, clen, al[i].r, k, km[nk++]);
    ++i;
  }
  i = 0;
  while (i < n - k + 1) {
    int off = from_end ? n - k - i : i, skip = 0;
    for (int j = off; j < off + k; ++j) if (al[j].q == -1 || al[j].r == -1) { skip = 1; i += j - off; break; }
    if (skip) { ++i; continue; }
    last = al[off];
    uint8_t me[8];
    kmer_at(chrom, clen, al[off].r, k, me);
    int cnt = 0;
    for (int t = 0; t < nk; ++t) cnt += memcmp(km[t], me, 8) == 0;
    if (cnt == 1) break;
    ++i;
  }
  free(km);
  return last;
}

typedef struct { int aln, rs, re, qs, qe; } ext_t;

/* clusterer.cpp:158-345 for one read; out needs n_sfs slots; returns the number of merged extended SFSs */
static int extend_alignment(const svb_alns_t *A, const svb_ref_t *R, int a, int flank, int ksize, int clipped, ext_t *out, int64_t *cnt,
                            int32_t *clip) {
  const int t = A->tid[a];
  if (t < 0 || t >= R->n_contigs || R->len[t] < 0) return 0;
  const uint8_t *chrom = R->seq + R->start[t];
  const int64_t clen = R->len[t];
  const uint32_t *cig = A->cigar + A->cigar_offs[a];
  const int n_cig = (int)(A->cigar_offs[a + 1] - A->cigar_offs[a]);
  int n_al = 0;
  pair_t *al = aligned_pairs(cig, n_cig, A->pos[a], &n_al);
  int last_pos = 0, n_local = 0, n_out = 0;
  ext_t *local = (ext_t *)malloc(sizeof(ext_t) * (size_t)(A->sfs_offs[a + 1] - A->sfs_offs[a] + 1));
  pair_t *pre = (pair_t *)malloc(sizeof(pair_t) * (size_t)(flank + 1)), *post = (pair_t *)malloc(sizeof(pair_t) * (size_t)(flank + 1));
  for (int64_t x = A->sfs_offs[a]; x < A->sfs_offs[a + 1]; ++x) {
    const int s = A->sfs_qs[x], e = A->sfs_qs[x] + A->sfs_len[x] - 1;
    int aln_start = -1, aln_end = -1, refs = -1, refe = -1;
    for (int i = last_pos; i < n_al; ++i) {
      const int q = al[i].q, r = al[i].r;
      if (q == -1 || r == -1) continue;
      else if (q < s) { last_pos = i; refs = r; aln_start = i; }
      else if (q > e) { refe = r; aln_end = i; break; }
    }
    if (refs == -1 && refe == -1) { ++cnt[0]; continue; }
    else if (refs == -1) {
      if (n_cig && (cig[0] & 0xf) == 4 && clipped) { clip[0] = A->pos[a]; clip[1] = (int)(cig[0] >> 4); } else ++cnt[1];
      continue;
    } else if (refe == -1) {
      if (n_cig && (cig[n_cig - 1] & 0xf) == 4 && clipped) { clip[2] = endpos_of(cig, n_cig, A->pos[a]); clip[3] = (int)(cig[n_cig - 1] >> 4); } else ++cnt[2];
      continue;
    }
    /* local_alpairs (:229-244): only its front and back are used */
    pair_t front = {-1, -1}, back = {-1, -1};
    int have = 0, last_r = refs - 1;
    for (int i = aln_start; i <= aln_end; ++i) {
      const int q = al[i].q, r = al[i].r;
      int push = 0;
      if (r == -1) { if (refs <= last_r && last_r <= refe) push = 1; }
      else { last_r = r; if (refs <= r && r <= refe) push = 1; }
      if (push) { if (!have) { front = al[i]; have = 1; } back = al[i]; }
      if (q != -1 && r != -1 && r >= refe) break;
    }
    int n_pre = 0, n_post = 0;
    for (int i = aln_start - 1; i >= 0; --i) { pre[n_pre++] = al[i]; if (n_pre == flank) break; }
    for (int i = 0; i < n_pre / 2; ++i) { pair_t tmp = pre[i]; pre[i] = pre[n_pre - 1 - i]; pre[n_pre - 1 - i] = tmp; }
    for (int i = aln_end + 1; i < n_al; ++i) { post[n_post++] = al[i]; if (n_post == flank) break; }
    pair_t pk = unique_kmer(pre, n_pre, ksize, 1, chrom, clen), sk = unique_kmer(post, n_post, ksize, 0, chrom, clen);
    if (pk.q == -1 || pk.r == -1) pk = front;
    if (sk.q == -1 || sk.r == -1) sk = back;
    if (pk.q == -1 || pk.r == -1 || sk.q == -1 || sk.r == -1) { ++cnt[3]; continue; }
    if ((unsigned)pk.r > (unsigned)(sk.r + ksize)) continue;
    ext_t v = {a, pk.r, sk.r + ksize, pk.q, sk.q + ksize};
    local[n_local++] = v;
  }
  for (int i = 0; i < n_local; ++i) {   /* :314-337 */
    int j;
    for (j = 0; j < n_out; ++j)
      if ((local[i].rs <= out[j].rs && out[j].rs <= local[i].re) || (out[j].rs <= local[i].rs && local[i].rs <= out[j].re)) break;
    if (j < n_out) {
      if (local[i].rs < out[j].rs) out[j].rs = local[i].rs;
      if (local[i].re > out[j].re) out[j].re = local[i].re;
      if (local[i].qs < out[j].qs) out[j].qs = local[i].qs;
      if (local[i].qe > out[j].qe) out[j].qe = local[i].qe;
    } else out[n_out++] = local[i];
  }
  free(al); free(local); free(pre); free(post);
  return n_out;
}

/* sort keys for the (stable) sort of the extended SFSs and for the per-thread std::map<(low, high)> */
typedef struct { int64_t key; int64_t idx; } kv_t;
static int kv_cmp(const void *a, const void *b) {
  const kv_t *x = (const kv_t *)a, *y = (const kv_t *)b;
  if (x->key != y->key) return x->key < y->key ? -1 : 1;
  return x->idx < y->idx ? -1 : x->idx > y->idx;
}

static void *zalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }

ORC_API void orc_clusters_free(svb_clusters_t *o) {
  free(o->tid); free(o->s); free(o->e); free(o->cov0); free(o->cov1); free(o->cov2); free(o->placed); free(o->sub_offs);
  free(o->sub_aln); free(o->sub_qs); free(o->sub_qe); free(o->sub_hp); free(o->rvec_offs); free(o->rvec); free(o->clip);
  memset(o, 0, sizeof(*o));
}

ORC_API int orc_cluster(const svb_alns_t *A, const svb_ref_t *R, int threads, int min_cluster_weight, int flank, int ksize, int clipped,
                        int omp_threads, svb_clusters_t *out) {
  memset(out, 0, sizeof(*out));
  const int64_t n = A->n_aln;
  const int T = threads > 0 ? threads : 1;
#ifdef _OPENMP
  if (omp_threads <= 0) omp_threads = omp_get_max_threads();
#else
  omp_threads = 1;
#endif
  int64_t n_acc = 0;
  int32_t *acc = (int32_t *)zalloc((size_t)n, 4);
  for (int64_t a = 0; a < n; ++a) if (A->sfs_offs[a + 1] > A->sfs_offs[a]) acc[n_acc++] = (int32_t)a;
  const int64_t n_sfs = n ? A->sfs_offs[n] : 0;
  ext_t *ext_all = (ext_t *)zalloc((size_t)n_sfs, sizeof(ext_t));
  int32_t *n_ext = (int32_t *)zalloc((size_t)n_acc, 4);
  int32_t *endp = (int32_t *)zalloc((size_t)n, 4);
  if (clipped) out->clip = (int32_t *)zalloc((size_t)n * 4, 4);
  int64_t c0 = 0, c1 = 0, c2 = 0, c3 = 0;
#pragma omp parallel for schedule(dynamic, 64) num_threads(omp_threads)
  for (int64_t a = 0; a < n; ++a) endp[a] = endpos_of(A->cigar + A->cigar_offs[a], (int)(A->cigar_offs[a + 1] - A->cigar_offs[a]), A->pos[a]);
#pragma omp parallel for schedule(dynamic, 16) num_threads(omp_threads) reduction(+ : c0, c1, c2, c3)
  for (int64_t i = 0; i < n_acc; ++i) {
    int64_t cnt[4] = {0, 0, 0, 0};
    int32_t cl4[4] = {0, 0, 0, 0};
    n_ext[i] = extend_alignment(A, R, acc[i], flank, ksize, clipped, ext_all + A->sfs_offs[acc[i]], cnt, cl4);
    c0 += cnt[0]; c1 += cnt[1]; c2 += cnt[2]; c3 += cnt[3];
    if (clipped) memcpy(out->clip + (int64_t)acc[i] * 4, cl4, 16);
  }
  out->unplaced = c0; out->s_unplaced = c1; out->e_unplaced = c2; out->unknown = c3;
  /* slots: accepted read i goes to thread slot i % T, slots are concatenated (clusterer.cpp:21-25, 109-133) */
  int64_t n_e = 0;
  for (int64_t i = 0; i < n_acc; ++i) n_e += n_ext[i];
  out->n_extended = n_e;
  ext_t *ext = (ext_t *)zalloc((size_t)n_e, sizeof(ext_t));
  {
    int64_t m = 0;
    for (int t = 0; t < T; ++t)
      for (int64_t i = t; i < n_acc; i += T)
        for (int k = 0; k < n_ext[i]; ++k) ext[m++] = ext_all[A->sfs_offs[acc[i]] + k];
  }
  int64_t nc = 0;
  int64_t *cl_first = NULL, *cl_cnt = NULL;   /* cluster c = members mem[cl_first[c] .. + cl_cnt[c]) */
  ext_t *mem = NULL;
  if (n_e) {
    /* std::sort by (chrom name, rs): stable here */
    kv_t *kv = (kv_t *)zalloc((size_t)n_e, sizeof(kv_t));
    for (int64_t i = 0; i < n_e; ++i) {
      const int t = A->tid[ext[i].aln];
      kv[i].key = ((int64_t)(R->name_rank ? R->name_rank[t] : t) << 32) | (uint32_t)ext[i].rs;
      kv[i].idx = i;
    }
    qsort(kv, (size_t)n_e, sizeof(kv_t), kv_cmp);
    ext_t *srt = (ext_t *)zalloc((size_t)n_e, sizeof(ext_t));
    int32_t *crank = (int32_t *)zalloc((size_t)n_e, 4);
    for (int64_t i = 0; i < n_e; ++i) { srt[i] = ext[kv[i].idx]; crank[i] = (int32_t)(kv[i].key >> 32); }
    free(kv);
    int mx = 0;
    for (int64_t i = 0; i < n_e; ++i) if (srt[i].re - srt[i].rs > mx) mx = srt[i].re - srt[i].rs;
    out->max_ext_len = mx;
    out->dist = (int)((double)mx * 1.1);
    /* :419-441 intervals */
    int64_t *iv = (int64_t *)zalloc((size_t)n_e * 2 + 2, 8);
    int64_t n_iv = 0, prev_i = 0;
    int prev_e = srt[0].re, prev_c = crank[0];
    for (int64_t i = 1; i < n_e; ++i) {
      if (crank[i] != prev_c) { prev_c = crank[i]; iv[n_iv * 2] = prev_i; iv[n_iv * 2 + 1] = i - 1; ++n_iv; prev_i = i; prev_e = srt[i].re; }
      else if (srt[i].rs - prev_e > out->dist) { iv[n_iv * 2] = prev_i; iv[n_iv * 2 + 1] = i - 1; ++n_iv; prev_e = srt[i].re; prev_i = i; }
    }
    iv[n_iv * 2] = prev_i; iv[n_iv * 2 + 1] = n_e - 1; ++n_iv;
    /* :443-474: every interval is swept into runs; run g of thread t is filed under key (low, high) in t's map.
     * A run = (thread, low, high, first, count, order of creation). */
    typedef struct { int t, low, high; int64_t first, cnt, seq; } run_t;
    run_t *runs = (run_t *)zalloc((size_t)n_e, sizeof(run_t));
    int64_t n_runs = 0;
    for (int64_t i = 0; i < n_iv; ++i) {
      int64_t j = iv[i * 2], last_j = j;
      int low = srt[j].rs, high = srt[j].re;
      for (++j; j <= iv[i * 2 + 1]; ++j) {
        if (srt[j].rs <= high) { if (srt[j].rs < low) low = srt[j].rs; if (srt[j].re > high) high = srt[j].re; }
        else {
          run_t r = {(int)(i % T), low, high, last_j, j - last_j, n_runs}; runs[n_runs++] = r;
          low = srt[j].rs; high = srt[j].re; last_j = j;
        }
      }
      run_t r = {(int)(i % T), low, high, last_j, iv[i * 2 + 1] + 1 - last_j, n_runs}; runs[n_runs++] = r;
    }
    /* map order: thread, then (low, high), then creation order (equal keys of one thread merge, :33-36) */
    kv_t *rk = (kv_t *)zalloc((size_t)n_runs, sizeof(kv_t));
    /* composite key does not fit 64 bits (t, low, high): sort three times, least significant first (stable via idx) */
    int64_t *ord = (int64_t *)zalloc((size_t)n_runs, 8);
    for (int64_t a = 0; a < n_runs; ++a) ord[a] = a;
    for (int pass = 0; pass < 3; ++pass) {
      for (int64_t a = 0; a < n_runs; ++a) {
        const run_t *r = &runs[ord[a]];
        rk[a].key = pass == 0 ? r->high : pass == 1 ? r->low : r->t;
        rk[a].idx = a;
      }
      qsort(rk, (size_t)n_runs, sizeof(kv_t), kv_cmp);
      int64_t *nw = (int64_t *)zalloc((size_t)n_runs, 8);
      for (int64_t a = 0; a < n_runs; ++a) nw[a] = ord[rk[a].idx];
      free(ord); ord = nw;
    }
    free(rk);
    mem = (ext_t *)zalloc((size_t)n_e, sizeof(ext_t));
    cl_first = (int64_t *)zalloc((size_t)n_runs, 8); cl_cnt = (int64_t *)zalloc((size_t)n_runs, 8);
    int64_t m = 0;
    for (int64_t a = 0; a < n_runs; ++a) {
      const run_t *r = &runs[ord[a]];
      const int fresh = a == 0 || runs[ord[a - 1]].t != r->t || runs[ord[a - 1]].low != r->low || runs[ord[a - 1]].high != r->high;
      if (fresh) { cl_first[nc] = m; cl_cnt[nc] = 0; ++nc; }
      memcpy(mem + m, srt + r->first, (size_t)r->cnt * sizeof(ext_t));
      m += r->cnt; cl_cnt[nc - 1] += r->cnt;
    }
    free(ord); free(runs); free(iv); free(srt); free(crank);
  }
  /* ---- fill_clusters (:478-610) */
  out->n_clusters = nc;
  out->tid = zalloc((size_t)nc, 4); out->s = zalloc((size_t)nc, 4); out->e = zalloc((size_t)nc, 4);
  out->cov0 = zalloc((size_t)nc, 4); out->cov1 = zalloc((size_t)nc, 4); out->cov2 = zalloc((size_t)nc, 4);
  out->placed = zalloc((size_t)nc, 1);
  out->sub_offs = zalloc((size_t)nc + 1, 8); out->rvec_offs = zalloc((size_t)nc + 1, 8);
  /* per-cluster results, concatenated afterwards */
  int32_t **p_sub = (int32_t **)zalloc((size_t)nc, sizeof(void *));
  uint8_t **p_rv = (uint8_t **)zalloc((size_t)nc, sizeof(void *));
  int32_t *p_nsub = (int32_t *)zalloc((size_t)nc, 4), *p_nrv = (int32_t *)zalloc((size_t)nc, 4);
  int64_t small1 = 0, small2 = 0, unext = 0;
  int max_span = 1;   /* the .bai query of clusterer.cpp:485-492 stands in as: first record of the chromosome that can still overlap */
  for (int64_t a = 0; a < n; ++a) if (endp[a] - A->pos[a] > max_span) max_span = endp[a] - A->pos[a];
#pragma omp parallel for schedule(dynamic, 4) num_threads(omp_threads) reduction(+ : small1, small2, unext)
  for (int64_t c = 0; c < nc; ++c) {
    const ext_t *sf = mem + cl_first[c];
    const int64_t ns = cl_cnt[c];
    out->tid[c] = A->tid[sf[0].aln];
    int min_s = 0x7fffffff, max_e = 0;
    /* the std::set<string> of read names: distinct alignment indices */
    int32_t *rd = (int32_t *)malloc((size_t)ns * 4);
    int n_rd = 0;
    for (int64_t k = 0; k < ns; ++k) {
      if (sf[k].rs < min_s) min_s = sf[k].rs;
      if (sf[k].re > max_e) max_e = sf[k].re;
      int seen = 0;
      for (int q = 0; q < n_rd; ++q) if (rd[q] == sf[k].aln) { seen = 1; break; }
      if (!seen) rd[n_rd++] = sf[k].aln;
    }
    if (n_rd < min_cluster_weight) { ++small1; free(rd); continue; }
    out->placed[c] = 1; out->s[c] = min_s; out->e[c] = max_e;
    const int beg = min_s - 1 < 0 ? 0 : min_s - 1, end = max_e;
    int cov[3] = {0, 0, 0};
    int cap = 64, nsub = 0, nrv = 0;
    int32_t *sub = (int32_t *)malloc((size_t)cap * 16);
    uint8_t *rv = (uint8_t *)malloc((size_t)cap);
    int rvcap = cap;
    /* the region query: every record of the chromosome that overlaps [beg, end), in file order */
    int64_t a0 = 0;
    {
      int64_t lo = 0, hi = n;   /* first record with (tid, pos) >= (cluster tid, beg - max_span) */
      while (lo < hi) {
        const int64_t mid = (lo + hi) / 2;
        if (A->tid[mid] < out->tid[c] || (A->tid[mid] == out->tid[c] && A->pos[mid] < beg - max_span)) lo = mid + 1; else hi = mid;
      }
      a0 = lo;
    }
    for (int64_t a = a0; a < n; ++a) {
      if (A->tid[a] != out->tid[c]) break;
      if (A->pos[a] >= end) break;                 /* coordinate-sorted */
      if (!(endp[a] > beg)) continue;
      const int hp_t = (A->hp[a] == 1 || A->hp[a] == 2) ? A->hp[a] : 0;
      ++cov[hp_t];
      int mine = 0;
      for (int q = 0; q < n_rd; ++q) if (rd[q] == (int32_t)a) { mine = 1; break; }
      if (nrv == rvcap) { rvcap *= 2; rv = (uint8_t *)realloc(rv, (size_t)rvcap); }
      rv[nrv++] = (uint8_t)(mine | ((hp_t == 0 ? 3 : hp_t) << 1));
      if (!mine) continue;
      int n_al = 0;
      pair_t *al = aligned_pairs(A->cigar + A->cigar_offs[a], (int)(A->cigar_offs[a + 1] - A->cigar_offs[a]), A->pos[a], &n_al);
      int qs = -1, qe = -1;
      for (int i = n_al - 1; i >= 0; --i) { if (al[i].q == -1 || al[i].r == -1) continue; if (al[i].r <= min_s) { qs = al[i].q; break; } }
      for (int i = 0; i < n_al; ++i) { if (al[i].q == -1 || al[i].r == -1) continue; if (al[i].r >= max_e) { qe = al[i].q; break; } }
      free(al);
      if (qs == -1 || qe == -1) { ++unext; continue; }
      if (nsub == cap) { cap *= 2; sub = (int32_t *)realloc(sub, (size_t)cap * 16); }
      sub[nsub * 4] = (int32_t)a; sub[nsub * 4 + 1] = qs; sub[nsub * 4 + 2] = qe; sub[nsub * 4 + 3] = hp_t;
      ++nsub;
    }
    free(rd);
    p_sub[c] = sub; p_nsub[c] = nsub;
    if (nsub >= min_cluster_weight) { out->cov0[c] = cov[0]; out->cov1[c] = cov[1]; out->cov2[c] = cov[2]; p_rv[c] = rv; p_nrv[c] = nrv; }
    else { ++small2; free(rv); }
  }
  out->small_clusters = small1; out->small_clusters_2 = small2; out->unextended = unext;
  int64_t ts = 0, tr = 0;
  for (int64_t c = 0; c < nc; ++c) { out->sub_offs[c] = ts; out->rvec_offs[c] = tr; ts += p_nsub[c]; tr += p_nrv[c]; }
  out->sub_offs[nc] = ts; out->rvec_offs[nc] = tr;
  out->sub_aln = zalloc((size_t)ts, 4); out->sub_qs = zalloc((size_t)ts, 4); out->sub_qe = zalloc((size_t)ts, 4); out->sub_hp = zalloc((size_t)ts, 4);
  out->rvec = zalloc((size_t)tr, 1);
  for (int64_t c = 0; c < nc; ++c) {
    for (int k = 0; k < p_nsub[c]; ++k) {
      const int64_t o = out->sub_offs[c] + k;
      out->sub_aln[o] = p_sub[c][k * 4]; out->sub_qs[o] = p_sub[c][k * 4 + 1]; out->sub_qe[o] = p_sub[c][k * 4 + 2]; out->sub_hp[o] = p_sub[c][k * 4 + 3];
    }
    if (p_nrv[c]) memcpy(out->rvec + out->rvec_offs[c], p_rv[c], (size_t)p_nrv[c]);
    free(p_sub[c]); free(p_rv[c]);
  }
  free(p_sub); free(p_rv); free(p_nsub); free(p_nrv); free(mem); free(cl_first); free(cl_cnt);
  free(acc); free(ext_all); free(n_ext); free(endp); free(ext);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ Caller::pcall */

typedef struct { int *subs; int n, cap; int cov, cov0, cov1, cov2; } job_t;

static void job_push(job_t *j, int s) {
  if (j->n == j->cap) { j->cap = j->cap ? j->cap * 2 : 8; j->subs = (int *)realloc(j->subs, (size_t)j->cap * 4); }
  j->subs[j->n++] = s;
}
static job_t job_shell(const job_t *c) { job_t j = {NULL, 0, 0, c->cov, c->cov0, c->cov1, c->cov2}; return j; }
static int job_len(const job_t *j, const int *len) {   /* Cluster::get_len, unsigned integer mean */
  unsigned l = 0, n = 0;
  for (int i = 0; i < j->n; ++i) { ++n; l += (unsigned)len[j->subs[i]]; }
  return (int)(l / n);
}
typedef struct { job_t *v; int n, cap; } jobs_t;
static job_t *jobs_add(jobs_t *J, job_t j) {
  if (J->n == J->cap) { J->cap = J->cap ? J->cap * 2 : 4; J->v = (job_t *)realloc(J->v, (size_t)J->cap * sizeof(job_t)); }
  J->v[J->n] = j;
  return &J->v[J->n++];
}
static void jobs_free(jobs_t *J, int keep) { for (int i = 0; i < J->n; ++i) if (i != keep) free(J->v[i].subs); free(J->v); }
static float fminf_(float a, float b) { return a < b ? a : b; }
static float fmaxf_(float a, float b) { return a > b ? a : b; }

/* caller.cpp:78-97 */
static jobs_t split_by_len(const job_t *in, const int *len, float min_ratio) {
  jobs_t out = {NULL, 0, 0};
  for (int x = 0; x < in->n; ++x) {
    const int s = in->subs[x];
    int i;
    for (i = 0; i < out.n; ++i) {
      const float cl = (float)job_len(&out.v[i], len), sl = (float)len[s];
      if (fminf_(cl, sl) / fmaxf_(cl, sl) >= min_ratio) break;
    }
    if (i == out.n) jobs_add(&out, job_shell(in));
    job_push(&out.v[i], s);
  }
  return out;
}
static int largest(const jobs_t *J) {
  unsigned vmax = 0; int imax = -1;
  for (int i = 0; i < J->n; ++i) if ((unsigned)J->v[i].n > vmax) { vmax = (unsigned)J->v[i].n; imax = i; }
  return imax;
}

/* caller.cpp:100-255; appends the kept sub-clusters to `res` */
static void split_cluster(const job_t *cluster, const int *len, const int32_t *hp, float min_ratio, int useht, jobs_t *res) {
  job_t c0 = job_shell(cluster), c1 = job_shell(cluster), c2 = job_shell(cluster);
  for (int x = 0; x < cluster->n; ++x) {
    const int s = cluster->subs[x];
    if (useht && hp[s] == 1) job_push(&c1, s);
    else if (useht && hp[s] == 2) job_push(&c2, s);
    else job_push(&c0, s);
  }
  c0.cov1 = -1; c0.cov2 = -1; c1.cov0 = -1; c1.cov2 = -1; c2.cov0 = -1; c2.cov1 = -1;
  if (c1.n == 0 && c2.n == 0) {
    jobs_t sub = split_by_len(&c0, len, min_ratio);
    int i1 = -1, i2 = -1;
    unsigned v1 = 0, v2 = 0;
    for (int i = 0; i < sub.n; ++i) {
      if ((unsigned)sub.v[i].n > v1) { v2 = v1; i2 = i1; v1 = (unsigned)sub.v[i].n; i1 = i; }
      else if ((unsigned)sub.v[i].n > v2) { v2 = (unsigned)sub.v[i].n; i2 = i; }
    }
    if (i1 != -1) jobs_add(res, sub.v[i1]);
    if (i2 != -1) jobs_add(res, sub.v[i2]);
    for (int i = 0; i < sub.n; ++i) if (i != i1 && i != i2) free(sub.v[i].subs);
    free(sub.v); free(c0.subs); free(c1.subs); free(c2.subs);
    return;
  }
  const int both = (c1.n ? 1 : 0) + (c2.n ? 2 : 0);
  jobs_t sub1 = split_by_len(&c1, len, min_ratio), sub2 = split_by_len(&c2, len, min_ratio);
  job_t fresh = {NULL, 0, 0, cluster->cov, cluster->cov0, -1, -1};
  for (int x = 0; x < c0.n; ++x) {
    const int s = c0.subs[x];
    const float sl = (float)len[s];
    int best_1 = -1, best_ratio_1 = -1, best_2 = -1, best_ratio_2 = -1;   /* declared int in the reference (caller.cpp:162,172) */
    for (int i = 0; i < sub1.n; ++i) {
      const float cl = (float)job_len(&sub1.v[i], len), r = fminf_(cl, sl) / fmaxf_(cl, sl);
      if (r >= min_ratio && r > (float)best_ratio_1) { best_1 = i; best_ratio_1 = (int)r; }
    }
    for (int i = 0; i < sub2.n; ++i) {
      const float cl = (float)job_len(&sub2.v[i], len), r = fminf_(cl, sl) / fmaxf_(cl, sl);
      if (r >= min_ratio && r > (float)best_ratio_2) { best_2 = i; best_ratio_2 = (int)r; }
    }
    if (both == 1) {
      if (best_1 == -1) job_push(&fresh, s);
      else { job_push(&sub1.v[best_1], s); ++sub1.v[best_1].cov1; --fresh.cov0; }
    } else if (both == 2) {
      if (best_2 == -1) job_push(&fresh, s);
      else { job_push(&sub2.v[best_2], s); ++sub2.v[best_2].cov2; --fresh.cov0; }
    } else {
      if (best_1 != -1 && best_ratio_1 > best_ratio_2) { job_push(&sub1.v[best_1], s); ++sub1.v[best_1].cov1; --fresh.cov0; }
      else if (best_2 != -1 && best_ratio_2 > best_ratio_1) { job_push(&sub2.v[best_2], s); ++sub2.v[best_2].cov2; --fresh.cov0; }
    }
  }
  int k1 = largest(&sub1), k2 = largest(&sub2);
  if (k1 != -1) jobs_add(res, sub1.v[k1]);
  if (k2 != -1) jobs_add(res, sub2.v[k2]);
  jobs_free(&sub1, k1); jobs_free(&sub2, k2);
  if (both != 3) {
    jobs_t subn = split_by_len(&fresh, len, min_ratio);
    const int kn = largest(&subn);
    if (kn != -1) {
      if (both == 1) subn.v[kn].cov1 = -1; else subn.v[kn].cov2 = -1;
      jobs_add(res, subn.v[kn]);
    }
    jobs_free(&subn, kn);
  }
  free(fresh.subs); free(c0.subs); free(c1.subs); free(c2.subs);
}

static uint8_t code_nt6(uint8_t b) { return (b >= 1 && b <= 4) ? (uint8_t)(b - 1) : 4; }
static uint8_t code_nt16(uint8_t b) { return b == 1 ? 0 : b == 2 ? 1 : b == 4 ? 2 : b == 8 ? 3 : 4; }
static uint8_t code_ascii(uint8_t c) {
  switch (c) {
    case 'A': case 'a': case 0: return 0;
    case 'C': case 'c': case 1: return 1;
    case 'G': case 'g': case 2: return 2;
    case 'T': case 't': case 'U': case 'u': case 3: return 3;
    default: return 4;
  }
}

ORC_API void orc_calls_free(svb_calls_t *o) {
  free(o->job_cluster); free(o->job_cov); free(o->job_sub_offs); free(o->job_sub); free(o->cons_offs); free(o->cons); free(o->score);
  free(o->cigar_offs); free(o->cigar); free(o->sv_job); free(o->sv_type); free(o->sv_pos); free(o->sv_len); free(o->sv_cpos); free(o->job_nv);
  memset(o, 0, sizeof(*o));
}

/* reads / ref: HOST memory only */
ORC_API int orc_call(const svb_clusters_t *CL, const svb_seqs_t *RD, const svb_ref_t *R, int min_cluster_weight, int min_sv_length,
                     float min_ratio, int useht, int omp_threads, svb_calls_t *out) {
  memset(out, 0, sizeof(*out));
#ifdef _OPENMP
  if (omp_threads <= 0) omp_threads = omp_get_max_threads();
#else
  omp_threads = 1;
#endif
  const int64_t nc = CL->n_clusters, n_sub = nc ? CL->sub_offs[nc] : 0;
  int *len = (int *)zalloc((size_t)n_sub, 4);
  for (int64_t k = 0; k < n_sub; ++k) len[k] = CL->sub_qe[k] >= CL->sub_qs[k] ? CL->sub_qe[k] - CL->sub_qs[k] + 1 : 0;
  jobs_t J = {NULL, 0, 0};
  int32_t *jc = NULL;
  int jc_cap = 0;
  for (int64_t c = 0; c < nc; ++c) {
    const int64_t a = CL->sub_offs[c], b = CL->sub_offs[c + 1];
    if (!CL->placed[c] || b - a < min_cluster_weight) continue;
    const int t = CL->tid[c];
    if (t < 0 || t >= R->n_contigs || CL->s[c] < 1 || CL->e[c] < CL->s[c] || CL->e[c] >= R->len[t]) { ++out->skipped_outside; continue; }
    job_t cl = {NULL, 0, 0, CL->cov0[c] + CL->cov1[c] + CL->cov2[c], CL->cov0[c], CL->cov1[c], CL->cov2[c]};
    for (int64_t k = a; k < b; ++k) job_push(&cl, (int)k);
    const int before = J.n;
    split_cluster(&cl, len, CL->sub_hp, min_ratio, useht, &J);
    free(cl.subs);
    if (J.n > jc_cap) { jc_cap = J.n * 2 + 8; jc = (int32_t *)realloc(jc, (size_t)jc_cap * 4); }
    for (int j = before; j < J.n; ++j) jc[j] = (int32_t)c;
  }
  const int64_t nj = J.n;
  out->n_jobs = nj;
  out->job_cluster = zalloc((size_t)nj, 4); out->job_cov = zalloc((size_t)nj * 4, 4); out->job_sub_offs = zalloc((size_t)nj + 1, 8);
  out->job_nv = zalloc((size_t)nj, 4); out->cons_offs = zalloc((size_t)nj + 1, 8); out->score = zalloc((size_t)nj, 4);
  out->cigar_offs = zalloc((size_t)nj + 1, 8);
  int64_t njs = 0;
  for (int64_t j = 0; j < nj; ++j) njs += J.v[j].n;
  out->job_sub = zalloc((size_t)njs, 4);
  uint8_t **p_cons = (uint8_t **)zalloc((size_t)nj, sizeof(void *));
  uint32_t **p_cig = (uint32_t **)zalloc((size_t)nj, sizeof(void *));
  int32_t *p_ncons = (int32_t *)zalloc((size_t)nj, 4), *p_ncig = (int32_t *)zalloc((size_t)nj, 4);
  {
    int64_t i = 0;
    for (int64_t j = 0; j < nj; ++j) {
      out->job_cluster[j] = jc[j];
      out->job_cov[j * 4] = J.v[j].cov; out->job_cov[j * 4 + 1] = J.v[j].cov0; out->job_cov[j * 4 + 2] = J.v[j].cov1; out->job_cov[j * 4 + 3] = J.v[j].cov2;
      out->job_sub_offs[j] = i;
      for (int k = 0; k < J.v[j].n; ++k) out->job_sub[i++] = J.v[j].subs[k];
    }
    out->job_sub_offs[nj] = i;
  }
  int64_t poa_cells = 0, ksw_cells = 0;
#pragma omp parallel for schedule(static, 1) num_threads(omp_threads) reduction(+ : poa_cells, ksw_cells)
  for (int64_t j = 0; j < nj; ++j) {
    const int c = out->job_cluster[j];
    /* run_poa (caller.cpp:257-308) */
    const int ns = J.v[j].n;
    int64_t *so = (int64_t *)zalloc((size_t)ns + 1, 8);
    int lmax = 1;
    for (int k = 0; k < ns; ++k) { so[k + 1] = so[k] + len[J.v[j].subs[k]]; if (len[J.v[j].subs[k]] > lmax) lmax = len[J.v[j].subs[k]]; }
    uint8_t *sq = (uint8_t *)zalloc((size_t)so[ns] + 1, 1);
    for (int k = 0; k < ns; ++k) {
      const int s = J.v[j].subs[k], qs = CL->sub_qs[s];
      const uint8_t *base = RD->seq + RD->offs[CL->sub_aln[s]];
      uint8_t *d = sq + so[k];
      for (int i = 0; i < len[s]; ++i) {
        const int p = qs + i;
        d[i] = RD->fmt == SVB_SEQ_BAM4 ? code_nt16((p & 1) ? (base[p >> 1] & 0xf) : (base[p >> 1] >> 4)) : RD->fmt == SVB_SEQ_NT6 ? code_nt6(base[p]) : code_ascii(base[p]);
      }
    }
    const int cap = 2 * lmax + 64;
    uint8_t *cons = (uint8_t *)zalloc((size_t)cap, 1);
    int64_t st[3] = {0, 0, 0};
    int cl = orc_poa(sq, so, ns, 1, 2, 4, 4, 2, 24, 1, 10, 0.01, cons, cap, st);
    if (cl > cap) cl = cap;
    poa_cells += st[0];
    free(sq); free(so);
    /* ksw_extd2_sse against chromosome[s, e] (caller.cpp:329-355) */
    const int tl = CL->e[c] - CL->s[c] + 1;
    uint8_t *tw = (uint8_t *)zalloc((size_t)tl, 1);
    const uint8_t *rp = R->seq + R->start[CL->tid[c]] + CL->s[c];
    for (int i = 0; i < tl; ++i) tw[i] = R->fmt == SVB_SEQ_NT6 ? code_nt6(rp[i]) : code_ascii(rp[i]);
    uint32_t *cg = (uint32_t *)zalloc((size_t)(cl + tl + 2), 4);
    int ncg = 0;
    out->score[j] = orc_ksw_extd2(cl, cons, tl, tw, 1, -9, -1, 16, 2, 41, 1, cg, cl + tl + 2, &ncg);
    ksw_cells += (int64_t)cl * tl;
    free(tw);
    p_cons[j] = cons; p_ncons[j] = cl; p_cig[j] = cg; p_ncig[j] = ncg;
  }
  out->poa_cells = poa_cells; out->ksw_cells = ksw_cells;
  int64_t tc = 0, tg = 0;
  for (int64_t j = 0; j < nj; ++j) { out->cons_offs[j] = tc; out->cigar_offs[j] = tg; tc += p_ncons[j]; tg += p_ncig[j]; }
  out->cons_offs[nj] = tc; out->cigar_offs[nj] = tg;
  out->cons = zalloc((size_t)tc, 1); out->cigar = zalloc((size_t)tg, 4);
  int64_t n_sv = 0;
  for (int64_t j = 0; j < nj; ++j) {
    memcpy(out->cons + out->cons_offs[j], p_cons[j], (size_t)p_ncons[j]);
    memcpy(out->cigar + out->cigar_offs[j], p_cig[j], (size_t)p_ncig[j] * 4);
    for (int i = 0; i < p_ncig[j]; ++i) if ((p_cig[j][i] & 0xf) != 0 && (int)(p_cig[j][i] >> 4) >= min_sv_length) ++n_sv;
    free(p_cons[j]); free(p_cig[j]);
  }
  out->n_svs = n_sv;
  out->sv_job = zalloc((size_t)n_sv, 4); out->sv_type = zalloc((size_t)n_sv, 1); out->sv_pos = zalloc((size_t)n_sv, 4);
  out->sv_len = zalloc((size_t)n_sv, 4); out->sv_cpos = zalloc((size_t)n_sv, 4);
  int64_t k = 0;
  for (int64_t j = 0; j < nj; ++j) {   /* caller.cpp:359-401 */
    unsigned rpos = (unsigned)CL->s[out->job_cluster[j]], cpos = 0;
    int nv = 0;
    for (int64_t i = out->cigar_offs[j]; i < out->cigar_offs[j + 1]; ++i) {
      const unsigned l = out->cigar[i] >> 4, op = out->cigar[i] & 0xf;
      if (op == 0) { rpos += l; cpos += l; }
      else {
        if (l >= (unsigned)min_sv_length) {
          out->sv_job[k] = (int32_t)j; out->sv_type[k] = op == 1 ? 0 : 1; out->sv_pos[k] = (int32_t)rpos; out->sv_len[k] = (int32_t)l; out->sv_cpos[k] = (int32_t)cpos;
          ++k; ++nv;
        }
        if (op == 1) cpos += l; else rpos += l;
      }
    }
    out->job_nv[j] = nv;
  }
  for (int64_t j = 0; j < nj; ++j) free(J.v[j].subs);
  free(J.v); free(jc); free(len); free(p_cons); free(p_cig); free(p_ncons); free(p_ncig);
  return 0;
}

/* ------------------------------------------------------------------------------------------------ batch helpers
 * (the CPU legs of bench.py: the same scalar restatements, one item per OpenMP task) */

ORC_API int64_t orc_poa_batch(const uint8_t *seqs, const int64_t *seq_offs, const int64_t *cluster_offs, int64_t n_clusters, int omp_threads,
                              uint8_t *cons, const int64_t *cons_offs /* n_clusters + 1: capacity slots */, int32_t *cons_len) {
#ifdef _OPENMP
  if (omp_threads <= 0) omp_threads = omp_get_max_threads();
#else
  omp_threads = 1;
#endif
  int64_t cells = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(omp_threads) reduction(+ : cells)
  for (int64_t c = 0; c < n_clusters; ++c) {
    const int64_t a = cluster_offs[c], b = cluster_offs[c + 1];
    int64_t st[3] = {0, 0, 0};
    const int cap = (int)(cons_offs[c + 1] - cons_offs[c]);
    int l = orc_poa(seqs, seq_offs + a, (int)(b - a), 1, 2, 4, 4, 2, 24, 1, 10, 0.01, cons + cons_offs[c], cap, st);
    cons_len[c] = l > cap ? cap : l;
    cells += st[0];
  }
  return cells;
}

ORC_API int64_t orc_ksw_batch(const uint8_t *q, const int64_t *qo, const uint8_t *t, const int64_t *to, int64_t n, int omp_threads, int32_t *score) {
#ifdef _OPENMP
  if (omp_threads <= 0) omp_threads = omp_get_max_threads();
#else
  omp_threads = 1;
#endif
  int64_t cells = 0;
#pragma omp parallel for schedule(dynamic, 1) num_threads(omp_threads) reduction(+ : cells)
  for (int64_t p = 0; p < n; ++p) {
    const int ql = (int)(qo[p + 1] - qo[p]), tl = (int)(to[p + 1] - to[p]);
    uint32_t *cg = (uint32_t *)malloc((size_t)(ql + tl + 2) * 4);
    int ncg = 0;
    score[p] = orc_ksw_extd2(ql, q + qo[p], tl, t + to[p], 1, -9, -1, 16, 2, 41, 1, cg, ql + tl + 2, &ncg);
    free(cg);
    cells += (int64_t)ql * tl;
  }
  return cells;
}

/* Assembler::assemble (assembler.cpp:34-56) for every read of a batch: records of read r = (qs, len)[offs[r] .. offs[r + 1]) in emit
 * order (descending qs); out_* hold the assembled records in ascending qs, out_cnt the count per read */
ORC_API int64_t orc_assemble_batch(const int64_t *offs, const int32_t *qs, const int32_t *len, int64_t n_reads, int32_t *out_qs, int32_t *out_len,
                                   int64_t *out_cnt) {
  int64_t m = 0;
  for (int64_t r = 0; r < n_reads; ++r) {
    const int64_t a = offs[r], b = offs[r + 1];
    int64_t i = b - 1, c = 0;          /* descending input: walk it backwards = ascending qs (stable for equal qs is moot: qs strictly decrease) */
    while (i >= a) {
      int64_t j = i - 1;
      int end = qs[i] + len[i];
      int prev_q = qs[i], prev_l = len[i];
      while (j >= a && prev_q + prev_l > qs[j]) { end = qs[j] + len[j]; prev_q = qs[j]; prev_l = len[j]; --j; }
      out_qs[m] = qs[i]; out_len[m] = end - qs[i]; ++m; ++c;
      i = j;
    }
    out_cnt[r] = c;
  }
  return m;
}
