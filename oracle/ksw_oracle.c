/*
 * oracle/ksw_oracle.c -- CPU restatement of ksw_extd2_sse as SVDSS calls it (caller.cpp:332-355).
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/sfs_oracle.c header for who may use it).
 *
 * PARITY UNPINNED: ksw2 (lh3/ksw2, unpinned HEAD, reference CMakeLists.txt:116-118) is fetched at
 * build time and is not in the reference tree; the reference has no test vectors for it.  This
 * file restates the published algorithm (Suzuki-Kasahara difference recurrence as implemented in
 * ksw2_extd2_sse.c + ksw_backtrack in ksw2.h; restated in SURVEY.md appendix A.2) in plain absolute
 * int32 arithmetic, which is mathematically the same DP as long as ksw2's int8 differences do not
 * overflow (ksw2 guarantees that for its accepted parameters).  Pinned by: (a) the optimal score is
 * unique, tests check it against an independent Gotoh-style min-cost DP; (b) CIGARs must re-score
 * to exactly that score; (c) left-alignment / tie-break behaviour follows the quoted rules below.
 *
 * Call site parameters (caller.cpp:333-349): m=5, match a=+1, mismatch b=-9, N scores -e2 = -1
 * (mat[24]==0 and no KSW_EZ_GENERIC_SC), q=16 e=2 q2=41 e2=1, w=-1 (full), zdrop=-1, end_bonus=-1,
 * flag=0 (global, score + CIGAR, gaps left-aligned).  i indexes the target, j the query.
 *
 *   H(i,j)   = max{ H(i-1,j-1)+s(i,j), E(i,j), F(i,j), E2(i,j), F2(i,j) }
 *   E(i+1,j) = max{ H(i,j)-q,  E(i,j)  } - e       (deletion: consumes target)
 *   F(i,j+1) = max{ H(i,j)-q,  F(i,j)  } - e       (insertion: consumes query)
 *   E2,F2 likewise with q2,e2.   H(-1,-1)=0, H(-1,j) = H(j,-1) = -min(q+(j+1)e, q2+(j+1)e2).
 * Traceback byte per cell: low 3 bits = arg max with priority H-diag, E, F, E2, F2 (a later state
 * wins only if strictly greater); bit 3..6 = "E/F/E2/F2 leaving this cell is a continuation",
 * i.e. X(i,j) > H(i,j) - open strictly.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))
#define KSW_NEG_INF (-0x40000000)

static inline int gapcost(int k, int q, int e, int q2, int e2) {
  int a = q + k * e, b = q2 + k * e2;
  return a < b ? a : b;
}

/* returns the score; cigar (len<<4|op, op 0=M 1=I 2=D) in forward order, *n_cigar ops.
 * If the cigar needs more than cap entries, *n_cigar is still the true count (nothing beyond cap
 * is written). */
ORC_API int orc_ksw_extd2(int ql, const uint8_t *query, int tl, const uint8_t *target, int a, int b,
                          int sc_n, int q, int e, int q2, int e2, uint32_t *cigar, int cap,
                          int *n_cigar) {
  *n_cigar = 0;
  if (ql <= 0 || tl <= 0) return KSW_NEG_INF; /* ksw_reset_extz + early return */
  if (q2 + e2 < q + e) { int t = q; q = q2; q2 = t; t = e; e = e2; e2 = t; }
  const int NEG = -0x3fffffff / 2;
  int32_t *H = (int32_t *)malloc(sizeof(int32_t) * (size_t)ql);
  int32_t *E = (int32_t *)malloc(sizeof(int32_t) * (size_t)ql);
  int32_t *E2 = (int32_t *)malloc(sizeof(int32_t) * (size_t)ql);
  uint8_t *p = (uint8_t *)malloc((size_t)ql * (size_t)tl);
  for (int j = 0; j < ql; ++j) { H[j] = -gapcost(j + 1, q, e, q2, e2); E[j] = NEG; E2[j] = NEG; }
  for (int i = 0; i < tl; ++i) {
    int32_t hleft = -gapcost(i + 1, q, e, q2, e2);
    int32_t hdiag = i ? -gapcost(i, q, e, q2, e2) : 0;
    int32_t f = hleft - q - e, f2 = hleft - q2 - e2;
    uint8_t ti = target[i];
    uint8_t *pr = p + (size_t)i * ql;
    for (int j = 0; j < ql; ++j) {
      int32_t hup = H[j];
      int32_t ee = (hup - q > E[j] ? hup - q : E[j]) - e;
      int32_t ee2 = (hup - q2 > E2[j] ? hup - q2 : E2[j]) - e2;
      uint8_t qj = query[j];
      int s = (ti == 4 || qj == 4) ? sc_n : (ti == qj ? a : b);
      int32_t h = hdiag + s;
      uint8_t d = 0;
      if (ee > h) { h = ee; d = 1; }
      if (f > h) { h = f; d = 2; }
      if (ee2 > h) { h = ee2; d = 3; }
      if (f2 > h) { h = f2; d = 4; }
      if (ee > h - q) d |= 0x08;
      if (f > h - q) d |= 0x10;
      if (ee2 > h - q2) d |= 0x20;
      if (f2 > h - q2) d |= 0x40;
      pr[j] = d;
      hdiag = hup;
      H[j] = h; E[j] = ee; E2[j] = ee2;
      f = (h - q > f ? h - q : f) - e;
      f2 = (h - q2 > f2 ? h - q2 : f2) - e2;
    }
  }
  int score = H[ql - 1];
  /* ksw_backtrack (ksw2.h), is_rot irrelevant for the result, min_intron_len = 0 */
  int n = 0, i = tl - 1, j = ql - 1, state = 0;
  uint32_t *rev = (uint32_t *)malloc(sizeof(uint32_t) * (size_t)(ql + tl + 2));
#define PUSH(op, len)                                              \
  do {                                                             \
    if (n > 0 && (rev[n - 1] & 0xf) == (uint32_t)(op)) rev[n - 1] += (uint32_t)(len) << 4; \
    else rev[n++] = ((uint32_t)(len) << 4) | (uint32_t)(op);       \
  } while (0)
  while (i >= 0 && j >= 0) {
    uint8_t tmp = p[(size_t)i * ql + j];
    if (state == 0) state = tmp & 7;
    else if (!((tmp >> (state + 2)) & 1)) state = 0;
    if (state == 0) state = tmp & 7;
    if (state == 0) { PUSH(0, 1); --i; --j; }
    else if (state == 1 || state == 3) { PUSH(2, 1); --i; }
    else { PUSH(1, 1); --j; }
  }
  if (i >= 0) PUSH(2, i + 1);
  if (j >= 0) PUSH(1, j + 1);
#undef PUSH
  *n_cigar = n;
  for (int k = 0; k < n && k < cap; ++k) cigar[k] = rev[n - 1 - k];
  free(rev); free(p); free(H); free(E); free(E2);
  return score;
}

/* Independent check of the optimum: minimum-cost formulation (Gotoh with two affine pieces),
 * cost = -score, computed column-major over the query with separate gap-length bookkeeping-free
 * recurrences.  Only the optimal value is compared. */
ORC_API int orc_affine2_score(int ql, const uint8_t *query, int tl, const uint8_t *target, int a,
                              int b, int sc_n, int q, int e, int q2, int e2) {
  if (ql <= 0 || tl <= 0) return KSW_NEG_INF;
  const int64_t INF = (int64_t)1 << 40;
  size_t W = (size_t)tl + 1;
  int64_t *M = (int64_t *)malloc(sizeof(int64_t) * W * 5 * 2);
  /* layers: 0 = ends in match/mismatch or origin, 1/2 = gap in query (piece 1/2), 3/4 = gap in target */
  int64_t *cur = M, *prv = M + W * 5;
  for (int j = 0; j <= ql; ++j) {
    int64_t *t_ = cur; cur = prv; prv = t_;
    for (int i = 0; i <= tl; ++i) {
      int64_t *c = cur + (size_t)i * 5;
      for (int k = 0; k < 5; ++k) c[k] = INF;
      if (i == 0 && j == 0) { c[0] = 0; continue; }
      if (i > 0) { /* consume target only (deletion) */
        const int64_t *u = cur + (size_t)(i - 1) * 5;
        int64_t best = u[0]; for (int k = 1; k < 5; ++k) if (u[k] < best) best = u[k];
        int64_t o1 = best + q + e, x1 = u[1] + e;
        int64_t o2 = best + q2 + e2, x2 = u[2] + e2;
        c[1] = o1 < x1 ? o1 : x1;
        c[2] = o2 < x2 ? o2 : x2;
      }
      if (j > 0) { /* consume query only (insertion) */
        const int64_t *l = prv + (size_t)i * 5;
        int64_t best = l[0]; for (int k = 1; k < 5; ++k) if (l[k] < best) best = l[k];
        int64_t o1 = best + q + e, x1 = l[3] + e;
        int64_t o2 = best + q2 + e2, x2 = l[4] + e2;
        c[3] = o1 < x1 ? o1 : x1;
        c[4] = o2 < x2 ? o2 : x2;
      }
      if (i > 0 && j > 0) {
        const int64_t *dg = prv + (size_t)(i - 1) * 5;
        int64_t best = dg[0]; for (int k = 1; k < 5; ++k) if (dg[k] < best) best = dg[k];
        uint8_t ti = target[i - 1], qj = query[j - 1];
        int s = (ti == 4 || qj == 4) ? sc_n : (ti == qj ? a : b);
        c[0] = best - s;
      }
    }
  }
  const int64_t *c = cur + (size_t)tl * 5;
  int64_t best = c[0]; for (int k = 1; k < 5; ++k) if (c[k] < best) best = c[k];
  free(M);
  return (int)(-best);
}

/* score of a CIGAR under the same model (gap of length k costs min(q+ke, q2+ke2)); returns
 * KSW_NEG_INF if the CIGAR does not consume exactly (ql, tl). */
ORC_API int orc_cigar_score(int ql, const uint8_t *query, int tl, const uint8_t *target, int a, int b,
                            int sc_n, int q, int e, int q2, int e2, const uint32_t *cigar, int n) {
  int i = 0, j = 0, sc = 0;
  for (int k = 0; k < n; ++k) {
    int len = (int)(cigar[k] >> 4), op = (int)(cigar[k] & 0xf);
    if (op == 0) {
      for (int x = 0; x < len; ++x, ++i, ++j) {
        if (i >= tl || j >= ql) return KSW_NEG_INF;
        uint8_t ti = target[i], qj = query[j];
        sc += (ti == 4 || qj == 4) ? sc_n : (ti == qj ? a : b);
      }
    } else if (op == 1) { j += len; sc -= gapcost(len, q, e, q2, e2); }
    else if (op == 2) { i += len; sc -= gapcost(len, q, e, q2, e2); }
    else return KSW_NEG_INF;
  }
  return (i == tl && j == ql) ? sc : KSW_NEG_INF;
}
