// Per-item work of the Clusterer (reference clusterer.cpp:156-403, 478-610) stated on CIGAR ops instead of
// materialised aligned-pair vectors: the reference builds get_aligned_pairs(aln) -- one (read pos, ref pos) pair per
// alignment column, bam.cpp:92-134 -- for every read that carries an SFS and then scans it; here the same columns are
// addressed through the CIGAR (a column index is a position in the concatenation of the M/=/X, I/S and D/N ops), and
// only the <= 2 * flank columns around an SFS are ever materialised.  Every function is __host__ __device__ and free
// of runtime-API calls: cluster.cu runs them one item per thread on the GPU, tests/emul compiles the same source
// for the CPU to check it against the literal Python transcription (tests/cluster_model.py) without a GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define SVB_HD __host__ __device__ __forceinline__
#else
#define SVB_HD inline
#endif

namespace svb {

constexpr int CL_FLANK_MAX = 128;   // config.hpp:85 fixes flank at 100
constexpr int CL_KSIZE_MAX = 8;     // config.hpp:86 fixes ksize at 7; a k-mer is compared as <= 8 packed bytes

// One alignment as the Clusterer sees it (bam1_core_t::pos, the CIGAR) plus the chromosome it lies on.
struct ClAln {
  const uint32_t* cig;   // BAM encoding: len << 4 | op, op = MIDNSHP=X
  int n_cig;
  int pos;
  const uint8_t* chrom;  // chromosome_seqs[chrom]: opaque bytes (ASCII or nt6 codes), equality is all that is used
  int64_t chrom_len;
};

// column kinds of bam.cpp:92-134: 0 = (q, r) match, 1 = (q, -1) I/S, 2 = (-1, r) D/N, 3 = no column (H, P)
SVB_HD int cl_kind(uint32_t op) { return (op == 0 || op == 7 || op == 8) ? 0 : (op == 1 || op == 4) ? 1 : (op == 2 || op == 3) ? 2 : 3; }

struct ClCol { int col, q, r; };

// clusterer.cpp:183-201: the LAST column i >= min_col with q != -1, r != -1, q < s (aln_start), and the FIRST
// column i >= min_col with both set and q > e (aln_end).  Matched columns are monotone in q, so the scan's
// "assign while q < s, stop at the first q > e" is exactly that.  col = -1 if none.
SVB_HD void cl_place(const ClAln& a, int s, int e, int min_col, ClCol& st, ClCol& en) {
  st.col = -1; st.q = -1; st.r = -1; en.col = -1; en.q = -1; en.r = -1;
  int c0 = 0, q0 = 0, r0 = a.pos;
  for (int k = 0; k < a.n_cig; ++k) {
    const int len = (int)(a.cig[k] >> 4), kind = cl_kind(a.cig[k] & 0xf);
    if (kind == 0 && len > 0) {
      // columns j of this op with q0 + j < s
      int m = s - q0; if (m > len) m = len;
      if (m > 0 && c0 + m - 1 >= min_col) { st.col = c0 + m - 1; st.q = q0 + m - 1; st.r = r0 + m - 1; }
      // first column j with q0 + j > e and c0 + j >= min_col
      int j = e + 1 - q0; if (j < 0) j = 0;
      if (min_col - c0 > j) j = min_col - c0;
      if (j < len) { en.col = c0 + j; en.q = q0 + j; en.r = r0 + j; return; }
      c0 += len; q0 += len; r0 += len;
    } else if (kind == 1) { c0 += len; q0 += len; }
    else if (kind == 2) { c0 += len; r0 += len; }
  }
}

// columns [lo, hi) of the aligned-pair list into q[] / r[] (hi - lo <= CL_FLANK_MAX); returns the number written
// (fewer when the list ends first)
SVB_HD int cl_window(const ClAln& a, int lo, int hi, int* q, int* r) {
  int n = 0, c0 = 0, q0 = 0, r0 = a.pos;
  if (lo < 0) lo = 0;
  for (int k = 0; k < a.n_cig && c0 < hi; ++k) {
    const int len = (int)(a.cig[k] >> 4), kind = cl_kind(a.cig[k] & 0xf);
    if (kind == 3) continue;
    int j0 = lo - c0; if (j0 < 0) j0 = 0;
    int j1 = hi - c0; if (j1 > len) j1 = len;
    for (int j = j0; j < j1; ++j) {
      q[n] = kind == 2 ? -1 : q0 + j;
      r[n] = kind == 1 ? -1 : r0 + j;
      ++n;
    }
    c0 += len;
    if (kind != 2) q0 += len;
    if (kind != 1) r0 += len;
  }
  return n;
}

// string(chromosome_seqs[chrom] + r, k) (clusterer.cpp:370, 394) as packed bytes; the reference reads a C string,
// i.e. never past the terminator: missing bytes stay 0
SVB_HD uint64_t cl_kmer(const ClAln& a, int r, int k) {
  uint64_t v = 0;
  for (int i = 0; i < k; ++i) {
    const int64_t p = (int64_t)r + i;
    const uint64_t b = (p >= 0 && p < a.chrom_len) ? a.chrom[p] : 0;
    v |= b << (8 * i);
  }
  return v;
}

// Clusterer::get_unique_kmers (clusterer.cpp:350-403) on a window of n columns: the first k-mer from the inner end
// whose k columns are all matched and that occurs once among the window's clean k-mers; the last clean k-mer looked
// at if none is unique; (-1, -1) if there is no clean k-mer.  km[] is scratch for n entries.
SVB_HD void cl_unique_kmer(const ClAln& a, const int* q, const int* r, int n, int k, bool from_end, uint64_t* km, int& oq, int& orr) {
  oq = -1; orr = -1;
  if (n < k) return;
  int nk = 0;
  int i = 0;
  while (i < n - k + 1) {                      // :359-373, the counting pass (a std::map<string, int> there)
    bool skip = false;
    for (int j = i; j < i + k; ++j)
      if (q[j] == -1 || r[j] == -1) { skip = true; i = j + 1; break; }
    if (skip) continue;
    km[nk++] = cl_kmer(a, r[i], k);
    ++i;
  }
  i = 0;
  while (i < n - k + 1) {                      // :377-401
    const int off = from_end ? n - k - i : i;
    bool skip = false;
    for (int j = off; j < off + k; ++j)
      if (q[j] == -1 || r[j] == -1) { skip = true; i += j - off; break; }
    if (skip) { ++i; continue; }
    oq = q[off]; orr = r[off];
    const uint64_t me = cl_kmer(a, r[off], k);
    int cnt = 0;
    for (int t = 0; t < nk; ++t) cnt += km[t] == me;
    if (cnt == 1) break;
    ++i;
  }
}

struct ClExt { int rs, re, qs, qe; };

// Clusterer::extend_alignment (clusterer.cpp:158-345) for one read: its SFSs (qs, len), in the order of the .sfs
// file, are placed on the reference through the alignment, extended to the unique k-mers of the flanks and merged
// when they overlap (:314-337).  out[] needs room for n_sfs records (it is filled in place); returns how many.
// cnt[4] += unplaced, s_unplaced, e_unplaced, unknown (:206-226, :294-299).  clip[0..1] = left clip (alignment
// start, clipped bases), clip[2..3] = right clip (bam_endpos, clipped bases) of a read whose SFS lies in a soft
// clip -- recorded only when `clipped` (config->clipped), and then the s/e_unplaced counters are not bumped.
SVB_HD int cl_extend_read(const ClAln& a, const int32_t* sfs_qs, const int32_t* sfs_len, int n_sfs, int flank, int ksize, bool clipped,
                          int endpos, ClExt* out, unsigned* cnt, int* clip) {
  int last_pos = 0, n_local = 0;
  int wq[CL_FLANK_MAX], wr[CL_FLANK_MAX];
  uint64_t km[CL_FLANK_MAX];
  for (int x = 0; x < n_sfs; ++x) {
    const int s = sfs_qs[x], e = sfs_qs[x] + sfs_len[x] - 1;
    ClCol st, en;
    cl_place(a, s, e, last_pos, st, en);
    if (st.col >= 0) last_pos = st.col;
    if (st.col < 0 && en.col < 0) { ++cnt[0]; continue; }
    if (st.col < 0) {
      const uint32_t c0 = a.n_cig ? a.cig[0] : 0;
      if ((c0 & 0xf) == 4 && clipped) { clip[0] = a.pos; clip[1] = (int)(c0 >> 4); } else ++cnt[1];
      continue;
    }
    if (en.col < 0) {
      const uint32_t c1 = a.n_cig ? a.cig[a.n_cig - 1] : 0;
      if ((c1 & 0xf) == 4 && clipped) { clip[2] = endpos; clip[3] = (int)(c1 >> 4); } else ++cnt[2];
      continue;
    }
    // local_alpairs (:229-244) is only ever asked for its front and back: the aln_start and aln_end columns
    // themselves (both matched, refs <= r <= refe holds for the first, the loop breaks on the second)
    int pq, pr, sq, sr;
    {
      const int lo = st.col - flank < 0 ? 0 : st.col - flank;
      const int n = cl_window(a, lo, st.col, wq, wr);
      cl_unique_kmer(a, wq, wr, n, ksize, true, km, pq, pr);
    }
    {
      const int n = cl_window(a, en.col + 1, en.col + 1 + flank, wq, wr);
      cl_unique_kmer(a, wq, wr, n, ksize, false, km, sq, sr);
    }
    if (pq == -1 || pr == -1) { pq = st.q; pr = st.r; }
    if (sq == -1 || sr == -1) { sq = en.q; sr = en.r; }
    if (pq == -1 || pr == -1 || sq == -1 || sr == -1) { ++cnt[3]; continue; }
    if ((unsigned)pr > (unsigned)(sr + ksize)) continue;                        // :301-303 (a warning there)
    ClExt v; v.rs = pr; v.re = sr + ksize; v.qs = pq; v.qe = sq + ksize;
    // merge into the read's list (:314-337); out[] doubles as local_extended_sfs: slot n_out <= slot being read
    int j;
    for (j = 0; j < n_local; ++j)
      if ((v.rs <= out[j].rs && out[j].rs <= v.re) || (out[j].rs <= v.rs && v.rs <= out[j].re)) break;
    if (j < n_local) {
      if (v.rs < out[j].rs) out[j].rs = v.rs;
      if (v.re > out[j].re) out[j].re = v.re;
      if (v.qs < out[j].qs) out[j].qs = v.qs;
      if (v.qe > out[j].qe) out[j].qe = v.qe;
    } else out[n_local++] = v;
  }
  return n_local;
}

// bam_endpos: pos + reference span (at least 1)
SVB_HD int cl_endpos(const uint32_t* cig, int n_cig, int pos) {
  int span = 0;
  for (int k = 0; k < n_cig; ++k) { const int kind = cl_kind(cig[k] & 0xf); if (kind == 0 || kind == 2) span += (int)(cig[k] >> 4); }
  return pos + (span ? span : 1);
}

// fill_clusters, clusterer.cpp:558-579: the read base aligned to the last matched column with ref <= min_s (scan
// from the end) and to the first matched column with ref >= max_e; -1 if none
SVB_HD void cl_window_on_read(const uint32_t* cig, int n_cig, int pos, int min_s, int max_e, int& qs, int& qe) {
  qs = -1; qe = -1;
  int ref = pos, rd = 0;
  for (int k = 0; k < n_cig; ++k) {
    const int len = (int)(cig[k] >> 4), kind = cl_kind(cig[k] & 0xf);
    if (kind == 0) {
      if (len > 0) {
        if (ref <= min_s) { int o = min_s - ref; if (o > len - 1) o = len - 1; qs = rd + o; }   // later ops overwrite: the last one wins
        if (qe == -1 && ref + len - 1 >= max_e) { int o = max_e - ref; if (o < 0) o = 0; qe = rd + o; }
      }
      ref += len; rd += len;
    } else if (kind == 1) rd += len;
    else if (kind == 2) ref += len;
  }
}

// fill_clusters (clusterer.cpp:485-607) for one cluster that has enough reads: every alignment of its chromosome
// overlapping chrom:min_s-max_e (htslib region: pos < max_e and bam_endpos > min_s - 1) counts into the coverage of
// its haplotype and into the RVEC list; those that carry one of the cluster's SFSs (`members`, ascending alignment
// indices) are cut to the window.  Alignments [lo, hi) = the candidates (BAM order).  sub_*[] need n_members slots,
// rvec[] hi - lo.  rvec byte = has-SFS | haplotype code << 1 (1, 2, or 3 for untagged: Cluster::reads).
SVB_HD void cl_fill_cluster(const int32_t* pos, const int32_t* endp, const int32_t* hp, const int64_t* cigar_offs, const uint32_t* cigar,
                            int lo, int hi, int min_s, int max_e, const int32_t* members, int n_members,
                            int32_t* sub_aln, int32_t* sub_qs, int32_t* sub_qe, int32_t* sub_hp, int& n_sub,
                            uint8_t* rvec, int& n_rv, int* cov, unsigned& unextended) {
  const int beg = min_s - 1 < 0 ? 0 : min_s - 1, end = max_e;
  n_sub = 0; n_rv = 0; cov[0] = cov[1] = cov[2] = 0;
  int m = 0;   // members[] and the candidates are both ascending: one merge pass
  for (int a = lo; a < hi; ++a) {
    if (!(pos[a] < end && endp[a] > beg)) continue;
    const int hp_t = (hp[a] == 1 || hp[a] == 2) ? hp[a] : 0;   // other values index out of range in the reference
    ++cov[hp_t];
    while (m < n_members && members[m] < a) ++m;
    const bool mine = m < n_members && members[m] == a;
    rvec[n_rv++] = (uint8_t)((mine ? 1 : 0) | ((hp_t == 0 ? 3 : hp_t) << 1));
    if (!mine) continue;
    int qs, qe;
    cl_window_on_read(cigar + cigar_offs[a], (int)(cigar_offs[a + 1] - cigar_offs[a]), pos[a], min_s, max_e, qs, qe);
    if (qs == -1 || qe == -1) { ++unextended; continue; }
    sub_aln[n_sub] = a; sub_qs[n_sub] = qs; sub_qe[n_sub] = qe; sub_hp[n_sub] = hp_t; ++n_sub;
  }
}

}  // namespace svb
