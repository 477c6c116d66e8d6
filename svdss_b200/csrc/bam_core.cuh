// Per-item work of the device BAM loader (csrc/bam_stream.cu; the reference: PingPong::load_batch_bam and the filters of
// ping_pong.cpp:53-131,196-203 over htslib's sam_read1): record plausibility, the walk of one segment of a window, the
// linking of the segments, and the parse of one record.  Every function is __host__ __device__ and free of runtime-API
// calls: bam_stream.cu runs them on the GPU, tests/emul/bam_emul.cpp compiles the same source for the CPU, where
// tests/test_bam_emul.py holds it against a plain Python BAM parser -- including windows whose payload contains byte
// strings that look like records.
#pragma once
#include <stdint.h>

#ifndef SVB_HD
#ifdef __CUDACC__
#define SVB_HD __host__ __device__ __forceinline__
#else
#define SVB_HD inline
#endif
#endif

namespace svb {

struct BamMeta {   // per record, device and host
  int32_t tid, l_qseq, xf, hp;
  uint16_t flag;
  uint8_t state;     // 0 dropped by the flag filter, 3 dropped for l_qseq < 100, 1 kept but not searched, 2 searched
  uint8_t name_len;  // without the NUL; 0 for dropped records
};

struct BamAln {     // the alignment of a record, for the Clusterer's scan (clusterer.cpp:58-153)
  int32_t pos, endpos;   // bam1_core_t::pos, bam_endpos (pos + reference span of the CIGAR, pos + 1 for an empty span)
  int32_t n_cigar;       // ops of the CIGAR in force: the field's, or the CG:B,I tag's when the field holds the placeholder
  int32_t cigar_rel;     // where those ops lie, relative to the record
  uint8_t mapq, pad[3];
};

SVB_HD uint32_t bam_ld16(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
SVB_HD uint32_t bam_ld32(const uint8_t* p) { return bam_ld16(p) | (bam_ld16(p + 2) << 16); }

// Could a record start at offset q (= the position of its block_size field) of win[0, total)?  Only the WALK'S SPEED
// depends on the answer (bam_link_segment checks every guess against the true chain).
SVB_HD bool bam_plausible(const uint8_t* win, int64_t q, int64_t total, int n_ref) {
  if (q + 36 > total) return false;
  const int64_t bs = (int32_t)bam_ld32(win + q);
  if (bs < 32 || bs > (1 << 28)) return false;
  const int32_t tid = (int32_t)bam_ld32(win + q + 4), pos = (int32_t)bam_ld32(win + q + 8);
  if (tid < -1 || tid >= n_ref || pos < -1) return false;
  const int64_t l_read_name = win[q + 12], n_cigar = (int64_t)bam_ld16(win + q + 16);
  const int64_t l_qseq = (int32_t)bam_ld32(win + q + 20);
  if (l_read_name < 1 || l_qseq < 0) return false;
  if (32 + l_read_name + 4 * n_cigar + (l_qseq + 1) / 2 + l_qseq > bs) return false;
  const int64_t nul = q + 36 + l_read_name - 1;
  if (nul < total && win[nul] != 0) return false;
  if (l_read_name > 1 && q + 36 < total && win[q + 36] < 33) return false;   // a name starts with a printable character
  return true;
}
// the guess of a segment: a plausible record whose successor is plausible too
SVB_HD bool bam_guess(const uint8_t* win, int64_t q, int64_t total, int n_ref) {
  if (!bam_plausible(win, q, total, n_ref)) return false;
  const int64_t q2 = q + 4 + (int64_t)(int32_t)bam_ld32(win + q);
  return q2 + 36 > total || bam_plausible(win, q2, total, n_ref);
}

// Follow the block_size chain from p while it stays below b: positions of the block_size fields into out[0, cap).
// flag: 0 = left the segment, 1 = the window ends inside the record at *end, 2 = a block_size below 32.
SVB_HD int64_t bam_chase(const uint8_t* win, int64_t p, int64_t b, int64_t total, int64_t* out, int64_t cap, int64_t* end, int* flag) {
  int64_t n = 0;
  *flag = 0;
  while (p < b && n < cap) {
    if (p + 4 > total) { *flag = 1; break; }
    const int64_t bs = (int32_t)bam_ld32(win + p);
    if (bs < 32) { *flag = 2; break; }
    if (p + 4 + bs > total) { *flag = 1; break; }
    out[n++] = p;
    p += 4 + bs;
  }
  *end = p;
  return n;
}

// Linking, one segment: enter at the true position *cur with *n records found so far.  Follows the true chain hop by hop
// until it meets the segment's guessed chain (then *join = index of the meeting point: chain[*join, c) is the truth and
// the caller appends it, *cur moves to the chain's end) or leaves the segment (*join = -1).  Returns false when the walk
// is over (window exhausted, capacity reached, or *err set).
SVB_HD bool bam_link_segment(const uint8_t* win, int64_t b, int64_t total, const int64_t* chain, int64_t c, int64_t chain_end, int chain_flag,
                             int64_t* rec_off, int64_t cap, int64_t* cur, int64_t* n, int64_t* join, int* err) {
  *join = -1;
  while (*cur < b) {
    // the usual case first: the true walk enters at the first record of the segment, which is where a right guess starts
    int64_t lo = 0, hi = c;
    if (c > 0 && chain[0] == *cur) hi = 0;
    while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (chain[mid] < *cur) lo = mid + 1; else hi = mid; }
    if (lo < c && chain[lo] == *cur && *n + (c - lo) <= cap) {
      *join = lo;
      *cur = chain_end;
      if (chain_flag == 2) *err = 1;
      return chain_flag == 0;
    }
    if (*cur + 4 > total || *n >= cap) return false;
    const int64_t bs = (int32_t)bam_ld32(win + *cur);
    if (bs < 32) { *err = 1; return false; }
    if (*cur + 4 + bs > total) return false;
    rec_off[(*n)++] = *cur + 4;
    *cur += 4 + bs;
  }
  return true;
}

// One record (p = its first byte after block_size): BAM spec 4.2 core fields, then the aux walk of host/io.hpp's
// BamReader::next; the load filters of ping_pong.cpp:66-75 and the XF rule of :196-203.  Returns false for a record
// whose fields run past its block_size.
SVB_HD bool bam_parse_record(const uint8_t* p, int putative, BamMeta* out, int64_t* seq_rel, BamAln* aln = nullptr) {
  const int64_t bs = (int64_t)(int32_t)bam_ld32(p - 4);
  BamMeta m;
  m.tid = (int32_t)bam_ld32(p);
  const unsigned l_read_name = p[8];
  const unsigned n_cigar = bam_ld16(p + 12);
  m.flag = (uint16_t)bam_ld16(p + 14);
  m.l_qseq = (int32_t)bam_ld32(p + 16);
  m.xf = 0; m.hp = 0;
  m.name_len = (uint8_t)(l_read_name ? l_read_name - 1 : 0);
  int64_t o = 32 + (int64_t)l_read_name + 4 * (int64_t)n_cigar;
  const int64_t seq_bytes = ((int64_t)m.l_qseq + 1) / 2;
  bool bad = m.l_qseq < 0 || o > bs || o + seq_bytes + (int64_t)m.l_qseq > bs;
  *seq_rel = o;
  bool has_xf = false;
  int64_t cg_rel = 32 + (int64_t)l_read_name, cg_n = n_cigar;     // the CIGAR in force
  if (!bad) {
    o += seq_bytes + m.l_qseq;
    while (o + 3 <= bs) {
      const char t0 = (char)p[o], t1 = (char)p[o + 1], ty = (char)p[o + 2];
      o += 3;
      int64_t iv = 0;
      bool is_int = false;
      switch (ty) {
        case 'A': o += 1; break;
        case 'c': iv = (int8_t)p[o]; is_int = true; o += 1; break;
        case 'C': iv = p[o]; is_int = true; o += 1; break;
        case 's': iv = (int16_t)bam_ld16(p + o); is_int = true; o += 2; break;
        case 'S': iv = bam_ld16(p + o); is_int = true; o += 2; break;
        case 'i': iv = (int32_t)bam_ld32(p + o); is_int = true; o += 4; break;
        case 'I': iv = bam_ld32(p + o); is_int = true; o += 4; break;
        case 'f': o += 4; break;
        case 'd': o += 8; break;
        case 'Z': case 'H': while (o < bs && p[o]) ++o; ++o; break;
        case 'B': {
          if (o + 5 > bs) { bad = true; break; }
          const char st = (char)p[o];
          const int32_t cnt = (int32_t)bam_ld32(p + o + 1);
          const int64_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
          const int64_t bytes = es * (int64_t)(cnt < 0 ? 0 : cnt);
          if (o + 5 + bytes > bs) { bad = true; break; }
          // CG:B,I -- the real CIGAR of a record with more than 65535 ops, whose CIGAR field then holds the placeholder
          // <l_seq>S<ref_len>N (SAM spec 4.2.2; htslib's sam_read1 swaps it in, as does host/io.hpp)
          if (t0 == 'C' && t1 == 'G' && st == 'I' && cnt > 0 && n_cigar == 2) {
            const uint32_t c0 = bam_ld32(p + 32 + l_read_name), c1 = bam_ld32(p + 32 + l_read_name + 4);
            if ((c0 & 0xf) == 4 && (int32_t)(c0 >> 4) == m.l_qseq && (c1 & 0xf) == 3) { cg_rel = o + 5; cg_n = cnt; }
          }
          o += 5 + bytes;
          break;
        }
        default: bad = true; break;
      }
      if (bad || o > bs) { bad = true; break; }
      if (is_int && t0 == 'X' && t1 == 'F') { has_xf = true; m.xf = (int32_t)iv; }
      if (is_int && t0 == 'H' && t1 == 'P') m.hp = (int32_t)iv;
    }
  }
  if (bad) m.state = 0;
  else if (m.flag & (0x4 | 0x800 | 0x100)) m.state = 0;                 // ping_pong.cpp:66-69 = clusterer.cpp:116-120
  else if (aln) m.state = 1;                                             // the Clusterer's scan has no length filter and batches nothing
  else if (m.l_qseq < 100) m.state = 3;                                  // :70-75
  else m.state = (putative && has_xf && m.xf != 0) ? 1 : 2;              // :196-203
  if (m.state == 0 || m.state == 3) m.name_len = 0;   // the host hears about them (a warning per short record) but needs no name
  *out = m;
  if (aln) {
    BamAln al;
    al.pos = (int32_t)bam_ld32(p + 4);
    al.mapq = p[9]; al.pad[0] = al.pad[1] = al.pad[2] = 0;
    al.n_cigar = bad ? 0 : (int32_t)cg_n;
    al.cigar_rel = (int32_t)cg_rel;
    int64_t span = 0;
    for (int64_t k = 0; k < al.n_cigar; ++k) {
      const uint32_t c = bam_ld32(p + cg_rel + 4 * k), op = c & 0xf;
      if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += c >> 4;
    }
    al.endpos = (int32_t)(al.pos + (span ? span : 1));
    *aln = al;
  }
  return !bad;
}

// ping_pong.cpp:88-94: seq_nt16_str, then seq_nt6_table -- A C G T -> 1 2 3 4, everything else 5
SVB_HD uint8_t bam_nt6_of_nt16(unsigned c) { return c == 1 ? 1 : c == 2 ? 2 : c == 4 ? 3 : c == 8 ? 4 : 5; }

}  // namespace svb
