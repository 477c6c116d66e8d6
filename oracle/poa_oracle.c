/*
 * oracle/poa_oracle.c -- CPU restatement of Caller::run_poa (caller.cpp:257-308): partial-order
 * alignment of a cluster's sub-reads (added in input order) and heaviest-bundling consensus.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/sfs_oracle.c header).
 *
 * PARITY UNPINNED: abPOA (yangao07/abPOA @e6bb6fd, reference CMakeLists.txt:97-99) is not in the
 * reference tree and cannot be fetched; the reference has no test vectors for it.  This file
 * restates the published algorithm with the reference's parameters (caller.cpp:259-271 over
 * abpoa_init_para defaults): global alignment, convex gap min(4+2k, 24+k), match +2,
 * mismatch -4, N scores 0, sequences added in input order with weight 1, adaptive band
 * w = 10 + 0.01*qlen around the best-scoring columns, heaviest-bundling consensus (max_n_cons=1).
 * Where abPOA's source would be needed for tie-breaks (topological order among unrelated nodes,
 * predecessor/op preference on equal scores) this file FIXES a deterministic rule, stated below,
 * and the CUDA kernel follows the same rules, so kernel == orc_poa(band=1) bit for bit while
 * orc_poa(band=0) (exact, un-banded) is the tolerance reference (edit distance <= 1 %).
 *
 * Rules:
 *  graph   node 0 = source, 1 = sink; nodes carry a rank (topological position).  A new node made
 *          while adding a read is ranked right after its anchor = the highest-ranked member of the
 *          aligned group of the last pre-existing node the read's path visited (so aligned groups
 *          stay contiguous in rank); new nodes of one read keep their creation order.
 *  DP      rows in rank order, columns j = 0..qlen.  Hp = max(M, E1, E2):
 *            M (i,j) = max_p H(p,j-1) + s(i,j)           first predecessor (edge order) wins ties
 *            E1(i,j) = max_p max(H(p,j)-o1, E1(p,j)) - e1  open wins ties; first predecessor wins
 *            F1(i,j) = max(Hp(i,j-1)-o1, F1(i,j-1)) - e1   open wins ties      (same for E2/F2)
 *            H = max(M,E1,E2,F1,F2), priority in that order (later only if strictly greater)
 *  band    remain[v] = remain[heaviest out-neighbour (first on ties)] + 1, remain[sink] = 0;
 *          c = qlen - remain[v] + 1; beg = max(0, min(mpl[v], c) - w); end = min(qlen, max(mpr[v], c) + w);
 *          mpl/mpr = min/max over predecessors of (first/last column holding the row maximum) + 1.
 *  end     best sink predecessor (first on ties) at column qlen.
 *  update  abpoa_add_graph_alignment: reuse node on equal base, else the aligned-group member with
 *          that base, else new node joined to the group; insertions make new nodes; every
 *          consecutive pair of used nodes gets edge weight += 1 (source->first, last->sink too).
 *  cons    reverse rank order: pick the heaviest out-edge, ties by larger-or-equal score of the head
 *          (later edge wins a full tie); score[v] = w + score[head]; walk from the source.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <limits.h>
#include <stdio.h>

#define ORC_API __attribute__((visibility("default")))
#define NEG (-(1 << 29))

typedef struct {
  int n, ncap, ne, ecap;
  uint8_t *base;
  int *rank, *first_in, *last_in, *first_out, *last_out, *ring, *nread;
  int *efrom, *eto, *ew, *enin, *enout;
} graph;

static void g_init(graph *g) {
  memset(g, 0, sizeof(*g));
  g->ncap = 1024; g->ecap = 2048;
#define A(f, t, c) g->f = (t *)malloc(sizeof(t) * (size_t)(c))
  A(base, uint8_t, g->ncap); A(rank, int, g->ncap); A(first_in, int, g->ncap); A(last_in, int, g->ncap);
  A(first_out, int, g->ncap); A(last_out, int, g->ncap); A(ring, int, g->ncap); A(nread, int, g->ncap);
  A(efrom, int, g->ecap); A(eto, int, g->ecap); A(ew, int, g->ecap); A(enin, int, g->ecap); A(enout, int, g->ecap);
#undef A
}
static void g_free(graph *g) {
  free(g->base); free(g->rank); free(g->first_in); free(g->last_in); free(g->first_out); free(g->last_out);
  free(g->ring); free(g->nread); free(g->efrom); free(g->eto); free(g->ew); free(g->enin); free(g->enout);
}
static int g_node(graph *g, uint8_t b) {
  if (g->n == g->ncap) {
    g->ncap *= 2;
#define R(f, t) g->f = (t *)realloc(g->f, sizeof(t) * (size_t)g->ncap)
    R(base, uint8_t); R(rank, int); R(first_in, int); R(last_in, int); R(first_out, int); R(last_out, int);
    R(ring, int); R(nread, int);
#undef R
  }
  int v = g->n++;
  g->base[v] = b; g->rank[v] = -1; g->first_in[v] = g->last_in[v] = g->first_out[v] = g->last_out[v] = -1;
  g->ring[v] = v; g->nread[v] = 0;
  return v;
}
static void g_edge(graph *g, int u, int v) { /* weight += 1, create at the tail of both lists */
  for (int e = g->first_out[u]; e >= 0; e = g->enout[e])
    if (g->eto[e] == v) { g->ew[e]++; return; }
  if (g->ne == g->ecap) {
    g->ecap *= 2;
#define R(f) g->f = (int *)realloc(g->f, sizeof(int) * (size_t)g->ecap)
    R(efrom); R(eto); R(ew); R(enin); R(enout);
#undef R
  }
  int e = g->ne++;
  g->efrom[e] = u; g->eto[e] = v; g->ew[e] = 1; g->enin[e] = g->enout[e] = -1;
  if (g->last_out[u] < 0) g->first_out[u] = e; else g->enout[g->last_out[u]] = e;
  g->last_out[u] = e;
  if (g->last_in[v] < 0) g->first_in[v] = e; else g->enin[g->last_in[v]] = e;
  g->last_in[v] = e;
}

typedef struct { int match, mismatch, o1, e1, o2, e2, wb; double wf; } poa_par;

static inline int sc(const poa_par *P, uint8_t a, uint8_t b) {
  return (a >= 4 || b >= 4) ? 0 : (a == b ? P->match : -P->mismatch);
}

/* traceback word: bits 0-2 H state (0 M,1 E1,2 E2,3 F1,4 F2); 3-4 Hp state (0 M,1 E1,2 E2);
 * 5 E1 ext, 6 E2 ext, 7 F1 ext, 8 F2 ext; 12-19 M pred ordinal, 20-25 E1 pred ordinal, 26-31 E2 */
#define TB_H(t) ((t) & 7)
#define TB_HP(t) (((t) >> 3) & 3)

/* order[] = node ids by rank (excluding source/sink), n_ord entries */
static void build_order(const graph *g, int *order, int *n_ord) {
  int m = 0;
  for (int v = 2; v < g->n; ++v) order[g->rank[v]] = v, ++m;
  *n_ord = m;
}

/* align read q (len ql) to g, append to graph. band: 0 = full matrix, 1 = adaptive band */
static void add_read(graph *g, const uint8_t *q, int ql, const poa_par *P, int band, int64_t *cells) {
  int n_ord;
  if (g->n == 2) { /* first read: a chain */
    int prev = 0;
    for (int j = 0; j < ql; ++j) { int v = g_node(g, q[j]); g->rank[v] = j; g->nread[v] = 1; g_edge(g, prev, v); prev = v; }
    g_edge(g, prev, 1);
    return;
  }
  int N = g->n;
  int *order = (int *)malloc(sizeof(int) * (size_t)N);
  build_order(g, order, &n_ord);
  const int W = ql + 1;
  /* remain */
  int *remain = (int *)malloc(sizeof(int) * (size_t)N);
  remain[1] = 0;
  for (int r = n_ord - 1; r >= -1; --r) {
    int v = r >= 0 ? order[r] : 0, bw = -1, bv = 1;
    for (int e = g->first_out[v]; e >= 0; e = g->enout[e])
      if (g->ew[e] > bw) { bw = g->ew[e]; bv = g->eto[e]; }
    remain[v] = remain[bv] + 1;
  }
  int w = P->wb + (int)(P->wf * ql);
  int *mpl = (int *)malloc(sizeof(int) * (size_t)N), *mpr = (int *)malloc(sizeof(int) * (size_t)N);
  int *beg = (int *)malloc(sizeof(int) * (size_t)N), *end = (int *)malloc(sizeof(int) * (size_t)N);
  for (int v = 0; v < N; ++v) { mpl[v] = INT_MAX; mpr[v] = -1; }
  /* rows are indexed by node id here (the kernel indexes by rank; same cells) */
  int32_t *H = (int32_t *)malloc(sizeof(int32_t) * (size_t)N * W);
  int32_t *E1 = (int32_t *)malloc(sizeof(int32_t) * (size_t)N * W);
  int32_t *E2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)N * W);
  uint32_t *TB = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)N * W);
  /* source row */
  {
    beg[0] = 0; end[0] = band ? (w < ql ? w : ql) : ql;
    for (int j = 0; j <= end[0]; ++j) {
      int c1 = P->o1 + j * P->e1, c2 = P->o2 + j * P->e2;
      H[j] = j ? -(c1 < c2 ? c1 : c2) : 0; E1[j] = NEG; E2[j] = NEG;
      TB[j] = j ? (uint32_t)(c1 <= c2 ? 3 : 4) : 0; /* F1 wins ties */
      if (j > 1) TB[j] |= (c1 <= c2) ? (1u << 7) : (1u << 8);
    }
    for (int e = g->first_out[0]; e >= 0; e = g->enout[e]) {
      int o = g->eto[e];
      if (1 < mpl[o]) mpl[o] = 1;
      if (1 > mpr[o]) mpr[o] = 1;
    }
  }
  for (int r = 0; r < n_ord; ++r) {
    int v = order[r];
    int b, en;
    if (band) {
      int c = ql - remain[v] + 1;
      int lo = mpl[v] < c ? mpl[v] : c, hi = mpr[v] > c ? mpr[v] : c;
      b = lo - w; if (b < 0) b = 0;
      en = hi + w; if (en > ql) en = ql;
      if (b > en) b = en;
    } else { b = 0; en = ql; }
    beg[v] = b; end[v] = en;
    int32_t *h = H + (size_t)v * W, *e1r = E1 + (size_t)v * W, *e2r = E2 + (size_t)v * W;
    uint32_t *tb = TB + (size_t)v * W;
    int32_t f1 = NEG, f2 = NEG, hp_prev = NEG;
    int32_t rowmax = NEG; int left = b, right = b;
    for (int j = b; j <= en; ++j) {
      int32_t m = NEG, x1 = NEG, x2 = NEG; int pm = 0, p1 = 0, p2 = 0, x1ext = 0, x2ext = 0, ord = 0;
      for (int e = g->first_in[v]; e >= 0; e = g->enin[e], ++ord) {
        int p = g->efrom[e];
        if (p != 0 && g->rank[p] >= r) abort(); /* ranks must be a topological order */
        const int32_t *ph = H + (size_t)p * W;
        if (j - 1 >= beg[p] && j - 1 <= end[p] && j >= 1) {
          int32_t c = ph[j - 1] + sc(P, g->base[v], q[j - 1]);
          if (c > m) { m = c; pm = ord; }
        }
        if (j >= beg[p] && j <= end[p]) {
          int32_t op = ph[j] - P->o1, ex = E1[(size_t)p * W + j];
          int32_t c = (op >= ex ? op : ex) - P->e1;
          if (c > x1) { x1 = c; p1 = ord; x1ext = ex > op; }
          op = ph[j] - P->o2; ex = E2[(size_t)p * W + j];
          c = (op >= ex ? op : ex) - P->e2;
          if (c > x2) { x2 = c; p2 = ord; x2ext = ex > op; }
        }
      }
      if (m < NEG) m = NEG;
      if (x1 < NEG) x1 = NEG;
      if (x2 < NEG) x2 = NEG;
      int32_t hp = m; uint32_t hps = 0;
      if (x1 > hp) { hp = x1; hps = 1; }
      if (x2 > hp) { hp = x2; hps = 2; }
      int f1ext = 0, f2ext = 0;
      if (j > b) {
        int32_t op = hp_prev - P->o1;
        f1ext = f1 > op; f1 = (f1ext ? f1 : op) - P->e1;
        op = hp_prev - P->o2;
        f2ext = f2 > op; f2 = (f2ext ? f2 : op) - P->e2;
        if (f1 < NEG) f1 = NEG;
        if (f2 < NEG) f2 = NEG;
      }
      int32_t hh = hp; uint32_t hs = hps;
      if (f1 > hh) { hh = f1; hs = 3; }
      if (f2 > hh) { hh = f2; hs = 4; }
      h[j] = hh; e1r[j] = x1; e2r[j] = x2;
      tb[j] = hs | (hps << 3) | ((uint32_t)x1ext << 5) | ((uint32_t)x2ext << 6) | ((uint32_t)f1ext << 7) |
              ((uint32_t)f2ext << 8) | ((uint32_t)(pm & 0xff) << 12) | ((uint32_t)(p1 & 0x3f) << 20) |
              ((uint32_t)(p2 & 0x3f) << 26);
      hp_prev = hp;
      if (hh > rowmax) { rowmax = hh; left = j; right = j; }
      else if (hh == rowmax) right = j;
    }
    if (cells) *cells += en - b + 1;
    for (int e = g->first_out[v]; e >= 0; e = g->enout[e]) {
      int o = g->eto[e];
      if (left + 1 < mpl[o]) mpl[o] = left + 1;
      if (right + 1 > mpr[o]) mpr[o] = right + 1;
    }
  }
  /* best sink predecessor at column ql */
  int best_p = -1; int32_t best = NEG - 1;
  for (int e = g->first_in[1]; e >= 0; e = g->enin[e]) {
    int p = g->efrom[e];
    int32_t val = (ql >= beg[p] && ql <= end[p]) ? H[(size_t)p * W + ql] : NEG;
    if (val > best) { best = val; best_p = p; }
  }
  /* traceback: ops from the end; op = (node id or -1 for insertion, query index or -1 for deletion) */
  int *op_node = (int *)malloc(sizeof(int) * (size_t)(N + ql + 2)), *op_q = (int *)malloc(sizeof(int) * (size_t)(N + ql + 2));
  int nop = 0;
  {
    int v = best_p, j = ql, state = 0; /* 0 H, 5 Hp, 1 E1, 2 E2, 3 F1, 4 F2 */
    while (v != 0 || j > 0) {
      if (v == 0) { /* source row: leading insertion */
        op_node[nop] = -1; op_q[nop] = j - 1; ++nop; --j; continue;
      }
      uint32_t t = TB[(size_t)v * W + j];
      if (state == 0) state = (int)TB_H(t);
      else if (state == 5) state = (int)TB_HP(t);
      /* after resolving H/Hp, state is one of M(0),E1,E2,F1,F2 */
      if (state == 0) { /* match/mismatch: consume node v and query j-1, go to pred */
        int ord = (int)((t >> 12) & 0xff), e = g->first_in[v];
        while (ord--) e = g->enin[e];
        op_node[nop] = v; op_q[nop] = j - 1; ++nop;
        v = g->efrom[e]; --j; state = 0;
      } else if (state == 1 || state == 2) { /* deletion of node v */
        int ord = (int)((t >> (state == 1 ? 20 : 26)) & 0x3f), e = g->first_in[v];
        while (ord--) e = g->enin[e];
        int ext = (int)((t >> (state == 1 ? 5 : 6)) & 1);
        op_node[nop] = v; op_q[nop] = -1; ++nop;
        v = g->efrom[e];
        if (!ext) state = 0; /* else stay in E1/E2 at the predecessor */
        if (v == 0) state = 0;
      } else { /* insertion of query j-1 */
        int ext = (int)((t >> (state == 3 ? 7 : 8)) & 1);
        op_node[nop] = -1; op_q[nop] = j - 1; ++nop;
        --j;
        if (!ext) state = 5; /* opened from Hp(i,j-1) */
      }
    }
  }
  /* graph update in forward order; new nodes are ranked after their anchor */
  int *anchor_cnt = (int *)calloc((size_t)n_ord + 1, sizeof(int)); /* slot 0 = source, r+1 = rank r */
  int *new_anchor = (int *)malloc(sizeof(int) * (size_t)(ql + 1)), *new_id = (int *)malloc(sizeof(int) * (size_t)(ql + 1));
  int n_new = 0, prev = 0, anchor = 0 /* slot */;
  const int n_old = N; /* nodes >= n_old are new in this read and have no rank yet */
  for (int k = nop - 1; k >= 0; --k) {
    int v = op_node[k], qi = op_q[k];
    if (v >= 0) { /* the path passes v's column: anchor after the whole aligned group */
      int mr = g->rank[v];
      for (int u = g->ring[v]; u != v; u = g->ring[u]) if (u < n_old && g->rank[u] > mr) mr = g->rank[u];
      anchor = mr + 1;
    }
    if (qi < 0) continue; /* deletion */
    int use;
    if (v >= 0) {
      uint8_t b = q[qi];
      if (g->base[v] == b) use = v;
      else {
        use = -1;
        for (int u = g->ring[v]; u != v; u = g->ring[u]) if (g->base[u] == b) { use = u; break; }
        if (use < 0) {
          use = g_node(g, b);
          g->ring[use] = g->ring[v]; g->ring[v] = use; /* join the aligned group */
          new_anchor[n_new] = anchor; new_id[n_new] = use; ++n_new; anchor_cnt[anchor]++;
        }
      }
    } else {
      use = g_node(g, q[qi]);
      new_anchor[n_new] = anchor; new_id[n_new] = use; ++n_new; anchor_cnt[anchor]++;
    }
    g->nread[use]++;
    g_edge(g, prev, use);
    prev = use;
  }
  g_edge(g, prev, 1);
  /* re-rank: old node of rank r moves to r + (#new nodes anchored at slots <= r), new nodes follow
   * their anchor in creation order */
  {
    int *shift = (int *)malloc(sizeof(int) * (size_t)(n_ord + 1)), acc = 0;
    for (int s = 0; s <= n_ord; ++s) { shift[s] = acc; acc += anchor_cnt[s]; }
    /* shift[s] = number of new nodes anchored at slots < s */
    for (int r = 0; r < n_ord; ++r) g->rank[order[r]] = r + shift[r + 1];
    int *used = (int *)calloc((size_t)n_ord + 1, sizeof(int));
    for (int k = 0; k < n_new; ++k) {
      int s = new_anchor[k];
      /* slot s holds old rank s-1 (or the source); new nodes go right after it */
      g->rank[new_id[k]] = (s - 1 + shift[s]) + 1 + used[s];
      used[s]++;
    }
    free(shift); free(used);
  }
  free(order); free(remain); free(mpl); free(mpr); free(beg); free(end); free(H); free(E1); free(E2); free(TB);
  free(op_node); free(op_q); free(anchor_cnt); free(new_anchor); free(new_id);
}

static int consensus(const graph *g, uint8_t *out, int cap) {
  int N = g->n, n_ord;
  int *order = (int *)malloc(sizeof(int) * (size_t)N);
  build_order(g, order, &n_ord);
  int *score = (int *)calloc((size_t)N, sizeof(int)), *nxt = (int *)malloc(sizeof(int) * (size_t)N);
  for (int r = n_ord - 1; r >= -1; --r) {
    int v = r >= 0 ? order[r] : 0, mw = -1, mi = -1;
    for (int e = g->first_out[v]; e >= 0; e = g->enout[e]) {
      int o = g->eto[e], w = g->ew[e];
      if (mw < w) { mw = w; mi = o; }
      else if (mw == w && score[mi] <= score[o]) mi = o;
    }
    nxt[v] = mi;
    score[v] = mi >= 0 ? mw + score[mi] : 0;
  }
  int n = 0;
  for (int v = nxt[0]; v > 1; v = nxt[v]) { if (n < cap) out[n] = g->base[v]; ++n; }
  free(order); free(score); free(nxt);
  return n;
}

/* seqs: codes 0..4 (caller.hpp _char26_table), offs[n+1]. Returns consensus length (bases written
 * up to cap). band: 0 exact, 1 adaptive. stats (optional): [0] DP cells, [1] nodes, [2] edges */
ORC_API int orc_poa(const uint8_t *seqs, const int64_t *offs, int n, int band, int match, int mismatch, int o1,
                    int e1, int o2, int e2, int wb, double wf, uint8_t *cons, int cap, int64_t *stats) {
  if (n <= 0) return 0;
  graph g;
  g_init(&g);
  g_node(&g, 0); g_node(&g, 0); /* source, sink */
  poa_par P = {match, mismatch < 0 ? -mismatch : mismatch, o1, e1, o2, e2, wb, wf};
  int64_t cells = 0;
  for (int i = 0; i < n; ++i) {
    int ql = (int)(offs[i + 1] - offs[i]);
    if (ql <= 0) continue;
    add_read(&g, seqs + offs[i], ql, &P, band, &cells);
  }
  int len = g.n > 2 ? consensus(&g, cons, cap) : 0;
  if (stats) { stats[0] = cells; stats[1] = g.n; stats[2] = g.ne; }
  g_free(&g);
  return len;
}

/* plain Levenshtein distance (tolerance checks) */
ORC_API int orc_edit_distance(const uint8_t *a, int n, const uint8_t *b, int m) {
  int *row = (int *)malloc(sizeof(int) * (size_t)(m + 1));
  for (int j = 0; j <= m; ++j) row[j] = j;
  for (int i = 1; i <= n; ++i) {
    int diag = row[0];
    row[0] = i;
    for (int j = 1; j <= m; ++j) {
      int up = row[j], c = diag + (a[i - 1] != b[j - 1]);
      if (up + 1 < c) c = up + 1;
      if (row[j - 1] + 1 < c) c = row[j - 1] + 1;
      diag = up; row[j] = c;
    }
  }
  int d = row[m];
  free(row);
  return d;
}
