// CPU harness for the device BAM loader: the per-item functions of svdss_b200/csrc/bam_core.cuh compiled as they are,
// with the kernels of bam_stream.cu (k_bam_walk, k_bam_walk_seg + k_bam_walk_link, k_bam_parse) replaced by plain loops
// over the same functions.  tests/test_bam_emul.py holds them against a plain Python BAM parser.
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../svdss_b200/csrc/bam_core.cuh"

using namespace svb;

// res: n, end position, error, segments joined on their guess
extern "C" void emul_bam_walk(const uint8_t* win, int64_t start, int64_t total, int n_seg, int n_ref, int64_t* rec_off, int64_t cap, int64_t* res) {
  if (n_seg <= 0) {   // k_bam_walk
    int64_t end = start;
    int flag = 0;
    const int64_t n = bam_chase(win, start, total, total, rec_off, cap, &end, &flag);
    for (int64_t i = 0; i < n; ++i) rec_off[i] += 4;
    res[0] = n; res[1] = end; res[2] = flag == 2; res[3] = 0;
    return;
  }
  const int64_t seg_len = (total - start + n_seg - 1) / n_seg, seg_cap = seg_len / 36 + 2;
  std::vector<int64_t> seg_pos((size_t)n_seg * (size_t)seg_cap), seg_cnt((size_t)n_seg), seg_end((size_t)n_seg);
  std::vector<int> seg_flag((size_t)n_seg);
  for (int s = 0; s < n_seg; ++s) {   // k_bam_walk_seg
    const int64_t a = start + (int64_t)s * seg_len, b = std::min(total, a + seg_len);
    int64_t p = -1;
    if (s == 0) p = a;
    else
      for (int64_t q = a; q < b; ++q)
        if (bam_guess(win, q, total, n_ref)) { p = q; break; }
    int64_t n = 0, end = p;
    int flag = 0;
    if (p >= 0) n = bam_chase(win, p, b, total, seg_pos.data() + (size_t)s * (size_t)seg_cap, seg_cap, &end, &flag);
    seg_cnt[(size_t)s] = n; seg_end[(size_t)s] = end; seg_flag[(size_t)s] = flag;
  }
  int64_t cur = start, n = 0, joined = 0;   // k_bam_walk_link
  int err = 0;
  for (int s = 0; s < n_seg; ++s) {
    const int64_t a = start + (int64_t)s * seg_len, b = std::min(total, a + seg_len);
    const int64_t* chain = seg_pos.data() + (size_t)s * (size_t)seg_cap;
    const int64_t c = seg_cnt[(size_t)s];
    int64_t join = -1;
    const bool go = bam_link_segment(win, b, total, chain, c, seg_end[(size_t)s], seg_flag[(size_t)s], rec_off, cap, &cur, &n, &join, &err);
    if (join >= 0) {
      for (int64_t k = join; k < c; ++k) rec_off[n++] = chain[k] + 4;
      ++joined;
    }
    if (!go) break;
  }
  res[0] = n; res[1] = cur; res[2] = err; res[3] = joined;
}

// per record: tid, l_qseq, xf, hp, flag, state, name_len, seq offset in the window; returns 1 if any record is malformed
extern "C" int emul_bam_parse(const uint8_t* win, const int64_t* rec_off, int64_t n, int putative, int32_t* tid, int32_t* l_qseq, int32_t* xf,
                              int32_t* hp, int32_t* flag, int32_t* state, int32_t* name_len, int64_t* seq_off) {
  int err = 0;
  for (int64_t i = 0; i < n; ++i) {
    BamMeta m;
    int64_t rel = 0;
    if (!bam_parse_record(win + rec_off[i], putative, &m, &rel)) err = 1;
    tid[i] = m.tid; l_qseq[i] = m.l_qseq; xf[i] = m.xf; hp[i] = m.hp; flag[i] = m.flag; state[i] = m.state; name_len[i] = m.name_len;
    seq_off[i] = rec_off[i] + rel;
  }
  return err;
}

extern "C" void emul_bam_decode(const uint8_t* seq4, int32_t l_qseq, uint8_t* out) {
  for (int32_t k = 0; k < l_qseq; ++k) {
    const uint8_t b = seq4[k >> 1];
    out[k] = bam_nt6_of_nt16((k & 1) ? (b & 0xfu) : (b >> 4));
  }
}

// alignment mode: state (0 / 1), pos, endpos, ops of the CIGAR in force and where they lie in the window
extern "C" int emul_bam_parse_aln(const uint8_t* win, const int64_t* rec_off, int64_t n, int32_t* state, int32_t* name_len, int32_t* pos, int32_t* endpos,
                                  int32_t* n_cigar, int64_t* cigar_off, int32_t* mapq) {
  int err = 0;
  for (int64_t i = 0; i < n; ++i) {
    BamMeta m;
    BamAln al;
    int64_t rel = 0;
    if (!bam_parse_record(win + rec_off[i], 1, &m, &rel, &al)) err = 1;
    state[i] = m.state; name_len[i] = m.name_len; pos[i] = al.pos; endpos[i] = al.endpos; n_cigar[i] = al.n_cigar;
    cigar_off[i] = rec_off[i] + al.cigar_rel; mapq[i] = al.mapq;
  }
  return err;
}
