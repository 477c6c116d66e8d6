// `SVDSS smooth` (reference smoother.cpp:347-488): rewrites every primary alignment so that it
// differs from the reference only where a structural variant may be -- aligned bases are replaced
// by the reference's, INS/DEL of at most min_indel_length are undone, longer ones and soft clips are
// kept -- and tags it XF:i (0 smoothed / 1 too dirty / 2 nothing of interest / 3 internal error),
// the tag `search` filters on (ping_pong.cpp:196-203).  The smoothed BAM goes to stdout.
//
// Host side by design: per read it is a CIGAR walk and a few memcpy; the cost is BGZF inflate and
// deflate, both done in parallel windows (io.hpp) like the reference's bgzf_mt(.., 8, ..).
// Output order = input order of the accepted records, which is what the reference's batch layout
// (k-major, thread-minor, smoother.cpp:424-450) produces for any --threads.
#pragma once
#include <future>
#include <chrono>
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <set>
#include <string>
#include <unordered_map>
#include <vector>

#include "io.hpp"

namespace svdss {

struct SmoothConfig {
  std::string bam, reference;
  unsigned min_mapq = 20;
  int min_indel_length = 20;   // config.hpp:88
  float accp = 0.98f;          // config.hpp:77
  int threads = 4;
};

struct SmoothOut { int xf; bool rebuilt; std::vector<uint32_t> cigar; std::string seq; std::vector<uint8_t> qual; };

// smoother.cpp:90-239.  read_seq = ASCII of the stored sequence, ref_seq = upper-cased chromosome.
inline SmoothOut smooth_read(const BamRecord& r, const std::string& read_seq, const std::string& ref_seq, int min_indel_length,
                             double al_accuracy) {
  SmoothOut o{2, false, {}, std::string(), {}};
  const uint8_t* qual = r.raw.data() + r.off_qual;
  bool should_ignore = true;
  size_t ref_offset = (size_t)r.pos, ins_offset = 0, match_offset = 0, soft_clip_offset = 0;
  int m_diff = 0;
  double num_match = 0, num_mismatch = 0;
  std::string& seq = o.seq;
  std::vector<uint8_t>& nq = o.qual;
  const size_t lq = (size_t)r.l_qseq;
  // `len` qualities from read offset `at` (the reference memcpy's past the read for a malformed CIGAR;
  // here the last quality is repeated instead)
  auto quals = [&](size_t at, size_t len) { for (size_t j = 0; j < len; ++j) nq.push_back(lq ? qual[std::min(at + j, lq - 1)] : (uint8_t)0xff); };
  auto rd = [&](size_t at, size_t len) {   // copy read bases + qualities
    const size_t before = seq.size();
    if (at < read_seq.size()) seq.append(read_seq, at, len);
    seq.resize(before + len, 'N');
    quals(at, len);
  };
  for (uint32_t c : r.cigar) {
    const uint32_t op = c & 0xf, len = c >> 4;
    const size_t at = soft_clip_offset + match_offset + ins_offset;
    if (op == 0 || op == 7 || op == 8) {                       // M = X: the reference's bases (:120-143)
      { const size_t before = seq.size(); if (ref_offset < ref_seq.size()) seq.append(ref_seq, ref_offset, len); seq.resize(before + len, 'N'); }
      quals(at, len);
      for (uint32_t j = 0; j < len; j++) {
        const char rc = ref_offset + j < ref_seq.size() ? ref_seq[ref_offset + j] : '\0', qc = at + j < read_seq.size() ? read_seq[at + j] : '\1';
        if (rc == qc) ++num_match; else ++num_mismatch;
      }
      ref_offset += len; match_offset += len;
      if (!o.cigar.empty() && (o.cigar.back() & 0xf) == 0) o.cigar.back() += (uint32_t)(len + m_diff) << 4;
      else o.cigar.push_back((uint32_t)(len + m_diff) << 4);
      m_diff = 0;
    } else if (op == 1) {                                      // I (:144-160)
      if ((int)len > min_indel_length) { should_ignore = false; rd(at, len); o.cigar.push_back(c); }
      ins_offset += len;
    } else if (op == 2) {                                      // D (:161-177)
      if ((int)len <= min_indel_length) {
        { const size_t before = seq.size(); if (ref_offset < ref_seq.size()) seq.append(ref_seq, ref_offset, len); seq.resize(before + len, 'N'); }
        quals(at, len);                                        // the reference copies the qualities that follow
        m_diff += (int)len;
      } else { should_ignore = false; o.cigar.push_back(c); }
      ref_offset += len;
    } else if (op == 4) {                                      // S (:178-187)
      should_ignore = false;
      rd(at, len);
      soft_clip_offset += len;
      o.cigar.push_back(c);
    } else break;                                              // H, P, N: stop (:188-192)
  }
  if (num_mismatch / num_match > al_accuracy) o.xf = 1;        // :214-216
  else if (should_ignore) o.xf = 2;                            // :217-218
  else { o.xf = 0; o.rebuilt = true; }                         // :219-231
  return o;
}

// mismatch rate of one alignment as compute_maxaccuracy measures it (smoother.cpp:310-339)
inline double mismatch_rate(const BamRecord& r, const std::string& read_seq, const std::string& ref_seq) {
  size_t ref_offset = (size_t)r.pos, ins_offset = 0, match_offset = 0, soft_clip_offset = 0;
  int num_match = 0, num_mismatch = 0;
  for (uint32_t c : r.cigar) {
    const uint32_t op = c & 0xf, len = c >> 4;
    if (op == 0 || op == 7 || op == 8) {
      for (uint32_t j = 0; j < len; ++j) {
        const size_t a = match_offset + ins_offset + soft_clip_offset + j;
        const char rc = ref_offset + j < ref_seq.size() ? ref_seq[ref_offset + j] : '\0', qc = a < read_seq.size() ? read_seq[a] : '\1';
        if (rc == qc) ++num_match; else ++num_mismatch;
      }
      ref_offset += len; match_offset += len;
    } else if (op == 1) ins_offset += len;
    else if (op == 2) ref_offset += len;
    else if (op == 4) soft_clip_offset += len;
    else break;
  }
  return num_mismatch / (double)num_match;
}

inline double percentile(const std::vector<double>& x, double q) {   // smoother.cpp:254-263
  const size_t n = x.size();
  const double id = (double)(n - 1) * q;
  const double lo = std::floor(id), hi = std::ceil(id);
  return (1.0 - (id - lo)) * x[(size_t)lo] + (id - lo) * x[(size_t)hi];
}

inline std::string decode_seq(const BamRecord& r) {
  std::string s((size_t)r.l_qseq, 'N');
  for (int32_t i = 0; i < r.l_qseq; ++i) { const uint8_t b = r.seq4[(size_t)i >> 1]; s[(size_t)i] = nt16_char((i & 1) ? (b & 0xf) : (b >> 4)); }
  return s;
}

// the record body with XF:i set to v: an existing integer XF is replaced where it stands, else the
// tag is appended (bam_aux_update_int; one unsigned byte is enough for 0..3)
inline void set_xf(std::vector<uint8_t>& body, size_t off_aux, int v) {
  size_t o = off_aux;
  while (o + 3 <= body.size()) {
    const char t0 = (char)body[o], t1 = (char)body[o + 1], ty = (char)body[o + 2];
    size_t sz;
    switch (ty) {
      case 'A': case 'c': case 'C': sz = 1; break;
      case 's': case 'S': sz = 2; break;
      case 'i': case 'I': case 'f': sz = 4; break;
      case 'd': sz = 8; break;
      case 'Z': case 'H': { size_t e = o + 3; while (e < body.size() && body[e]) ++e; sz = e - (o + 3) + 1; break; }
      case 'B': {
        if (o + 8 > body.size()) return;
        const char st = (char)body[o + 3]; int32_t cnt; memcpy(&cnt, &body[o + 4], 4);
        sz = 5 + (size_t)((st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4) * (size_t)(cnt < 0 ? 0 : cnt);
        break;
      }
      default: return;
    }
    if (t0 == 'X' && t1 == 'F' && (ty == 'c' || ty == 'C' || ty == 's' || ty == 'S' || ty == 'i' || ty == 'I')) {
      body.erase(body.begin() + (std::ptrdiff_t)(o + 2), body.begin() + (std::ptrdiff_t)(o + 3 + sz));
      const uint8_t rep[2] = {(uint8_t)'C', (uint8_t)v};
      body.insert(body.begin() + (std::ptrdiff_t)(o + 2), rep, rep + 2);
      return;
    }
    o += 3 + sz;
  }
  const uint8_t tag[4] = {'X', 'F', 'C', (uint8_t)v};
  body.insert(body.end(), tag, tag + 4);
}

// rebuild_bam_entry (smoother.cpp:51-88): core fields with the new n_cigar / l_qseq, qname, new
// CIGAR, re-packed sequence, qualities, the aux block as it was
inline std::vector<uint8_t> rebuild_record(const BamRecord& r, const SmoothOut& s) {
  static uint8_t nt16_of[256];
  static bool init = false;
  if (!init) { memset(nt16_of, 15, sizeof(nt16_of)); const char* t = "=ACMGRSVTWYHKDBN"; for (int i = 0; i < 16; ++i) { nt16_of[(uint8_t)t[i]] = (uint8_t)i; nt16_of[(uint8_t)tolower(t[i])] = (uint8_t)i; } init = true; }
  const size_t l = s.seq.size();
  std::vector<uint8_t> b(r.raw.begin(), r.raw.begin() + (std::ptrdiff_t)r.off_cigar);   // core + qname
  // more than 65535 ops do not fit bam1_core_t::n_cigar: the CIGAR field then holds <l_seq>S<ref_len>N and the real one
  // travels in a CG:B,I tag (SAM spec 4.2.2; what htslib's bam_write1 does)
  const bool long_cigar = s.cigar.size() > 65535;
  uint32_t placeholder[2] = {0, 0};
  if (long_cigar) {
    uint32_t span = 0;
    for (uint32_t c : s.cigar) { const uint32_t op = c & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += c >> 4; }
    placeholder[0] = (uint32_t)l << 4 | 4; placeholder[1] = span << 4 | 3;
  }
  const uint16_t n_cigar = long_cigar ? 2 : (uint16_t)s.cigar.size();
  const int32_t l_qseq = (int32_t)l;
  memcpy(&b[12], &n_cigar, 2);
  memcpy(&b[16], &l_qseq, 4);
  const uint8_t* cp = long_cigar ? reinterpret_cast<const uint8_t*>(placeholder) : reinterpret_cast<const uint8_t*>(s.cigar.data());
  b.insert(b.end(), cp, cp + 4 * (size_t)n_cigar);
  const size_t so = b.size();
  b.resize(so + (l + 1) / 2, 0);
  for (size_t i = 0; i < l; ++i) b[so + (i >> 1)] |= (uint8_t)(nt16_of[(uint8_t)s.seq[i]] << ((i & 1) ? 0 : 4));   // encode_bam_seq, bam.cpp:47-63
  b.insert(b.end(), s.qual.begin(), s.qual.end());
  // the aux block as it was, minus a CG tag the reader resolved (its CIGAR is stale now)
  if (r.cg_cigar && r.cg_off >= r.off_aux && r.cg_off + r.cg_len <= r.raw.size()) {
    b.insert(b.end(), r.raw.begin() + (std::ptrdiff_t)r.off_aux, r.raw.begin() + (std::ptrdiff_t)r.cg_off);
    b.insert(b.end(), r.raw.begin() + (std::ptrdiff_t)(r.cg_off + r.cg_len), r.raw.end());
  } else b.insert(b.end(), r.raw.begin() + (std::ptrdiff_t)r.off_aux, r.raw.end());
  if (long_cigar) {
    const uint8_t hd[4] = {'C', 'G', 'B', 'I'};
    const int32_t cnt = (int32_t)s.cigar.size();
    b.insert(b.end(), hd, hd + 4);
    const uint8_t* q = reinterpret_cast<const uint8_t*>(&cnt);
    b.insert(b.end(), q, q + 4);
    const uint8_t* c2 = reinterpret_cast<const uint8_t*>(s.cigar.data());
    b.insert(b.end(), c2, c2 + 4 * s.cigar.size());
  }
  return b;
}

inline bool smooth_accept(const BamRecord& r, unsigned min_mapq, const std::vector<std::string>& names,
                          const std::unordered_map<std::string, std::string>& seqs, std::set<std::string>* warned,
                          void (*log)(const char*, const std::string&)) {
  if (r.flag & 0x4 || r.flag & 0x800 || r.flag & 0x100) return false;   // smoother.cpp:508-511
  if (r.mapq < min_mapq) return false;                                   // :512-514
  if (r.l_qseq < 2) { if (log) log("warning", "Alignment filtered due to l_qseq. Why are we here? Please check"); return false; }
  if (r.tid < 0 || (size_t)r.tid >= names.size()) { if (log) log("critical", "core.tid < 0. Why are we here? Please check"); exit(1); }
  if (!seqs.count(names[(size_t)r.tid])) {
    if (warned && log && warned->insert(names[(size_t)r.tid]).second)
      log("warning", "Skipping alignment(s) on " + names[(size_t)r.tid] + " since it is not present in the reference");
    return false;
  }
  return true;
}

inline int run_smooth(const SmoothConfig& c, void (*log)(const char*, const std::string&)) {
  std::unordered_map<std::string, std::string> seqs;   // load_chromosomes: upper-cased
  {
    FastxReader fx(c.reference);
    if (!fx.ok()) { log("critical", "cannot open reference " + c.reference); return 1; }
    FastxRecord r;
    while (fx.next(r)) { for (auto& ch : r.seq) ch = (char)toupper((unsigned char)ch); seqs[r.name] = r.seq; }
  }
  // pass 1 (compute_maxaccuracy, :267-345): accp-percentile of the mismatch rate of the first 10000 accepted alignments
  double al_accuracy = 0;
  auto now = []() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t_pass1 = now(), t_read = 0, t_smooth = 0, t_write = 0;
  {
    BamReader bam(c.bam);
    if (!bam.ok()) { log("critical", "cannot read BAM " + c.bam); return 1; }
    bam.want_alignment(true);
    std::vector<BamRecord> first;
    BamRecord r;
    while (first.size() < 10000 && bam.next(r) == 1)
      if (smooth_accept(r, c.min_mapq, bam.ref_names(), seqs, nullptr, nullptr)) first.push_back(r);
    std::vector<double> acc(first.size());
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)first.size(); ++i)
      acc[(size_t)i] = mismatch_rate(first[(size_t)i], decode_seq(first[(size_t)i]), seqs.at(bam.ref_names()[(size_t)first[(size_t)i].tid]));
    if (acc.empty()) { log("critical", "no usable alignment in " + c.bam); return 1; }
    std::sort(acc.begin(), acc.end());
    al_accuracy = percentile(acc, c.accp);
  }
  log("info", "Max allowed alignment accuracy: " + std::to_string(al_accuracy));
  t_pass1 = now() - t_pass1;
  BamReader bam(c.bam);
  if (!bam.ok()) { log("critical", "cannot read BAM " + c.bam); return 1; }
  bam.want_raw(true);
  int level = 6;   // zlib's default, what hts_open("-", "wb") uses (smoother.cpp:362)
  if (const char* e = getenv("SVB_BGZF_LEVEL")) level = std::max(0, std::min(9, atoi(e)));
  BgzfWriter out(stdout, level);
  {  // sam_hdr_write: the header as it came
    std::vector<uint8_t> h;
    auto put32 = [&](int32_t v) { const uint8_t* p = reinterpret_cast<const uint8_t*>(&v); h.insert(h.end(), p, p + 4); };
    h.insert(h.end(), {'B', 'A', 'M', 1});
    put32((int32_t)bam.header_text().size());
    h.insert(h.end(), bam.header_text().begin(), bam.header_text().end());
    put32((int32_t)bam.ref_names().size());
    for (size_t i = 0; i < bam.ref_names().size(); ++i) {
      put32((int32_t)bam.ref_names()[i].size() + 1);
      h.insert(h.end(), bam.ref_names()[i].begin(), bam.ref_names()[i].end());
      h.push_back(0);
      put32(bam.ref_lens()[i]);
    }
    if (!out.write(h.data(), h.size())) { log("critical", "Can't write corrected BAM header, aborting.."); return 1; }
  }
  log("info", "Smoothing alignments on " + std::to_string(c.threads) + " threads..");
  std::set<std::string> warned;
  const size_t BATCH = 4096;
  std::vector<BamRecord> batch, next_batch;
  std::vector<std::vector<uint8_t>> bodies;
  uint64_t processed = 0, written = 0;
  int counts[4] = {0, 0, 0, 0};
  bool eof = false;
  // The record parse is the one serial part of the loop (the reference loads its next batch on a thread of
  // its own, smoother.cpp:412-460): batch k+1 is read by a helper thread while batch k is smoothed and written.
  // Only the helper touches the reader, `warned` and `processed` while it runs.
  auto read_batch = [&](std::vector<BamRecord>* dst) -> int {
    dst->clear();
    BamRecord r;
    int st_ = 1;
    while (dst->size() < BATCH && (st_ = bam.next(r)) == 1) {
      ++processed;
      if (smooth_accept(r, c.min_mapq, bam.ref_names(), seqs, &warned, log)) { dst->emplace_back(std::move(r)); r = BamRecord(); }
    }
    return st_;
  };
  std::future<int> pending = std::async(std::launch::async, read_batch, &next_batch);
  while (!eof) {
    double t0 = now();
    const int st = pending.get();
    batch.swap(next_batch);
    if (st != 1) eof = true;
    if (st < 0) { log("critical", "truncated or corrupt BAM"); return 1; }
    if (!eof) pending = std::async(std::launch::async, read_batch, &next_batch);
    bodies.assign(batch.size(), std::vector<uint8_t>());
    std::vector<int> xf(batch.size(), 0);
    t_read += now() - t0; t0 = now();
#pragma omp parallel for schedule(dynamic, 64)
    for (long long i = 0; i < (long long)batch.size(); ++i) {
      const BamRecord& b = batch[(size_t)i];
      const SmoothOut s = smooth_read(b, decode_seq(b), seqs.at(bam.ref_names()[(size_t)b.tid]), c.min_indel_length, al_accuracy);
      std::vector<uint8_t> body = s.rebuilt ? rebuild_record(b, s) : b.raw;
      // the aux block starts after the (possibly new) qualities
      const size_t off_aux = s.rebuilt ? b.off_cigar + 4 * s.cigar.size() + (s.seq.size() + 1) / 2 + s.seq.size() : b.off_aux;
      set_xf(body, off_aux, s.xf);
      bodies[(size_t)i].swap(body);
      xf[(size_t)i] = s.xf;
    }
    t_smooth += now() - t0; t0 = now();
    for (size_t i = 0; i < bodies.size(); ++i) {
      const int32_t bs = (int32_t)bodies[i].size();
      if (!out.write(&bs, 4) || !out.write(bodies[i].data(), bodies[i].size())) { log("critical", "Can't write corrected BAM record, aborting.."); return 1; }
      ++written; ++counts[xf[i] & 3];
    }
    t_write += now() - t0;
  }
  { const double t0 = now(); const bool okc = out.close(); t_write += now() - t0; if (!okc) { log("critical", "Can't write corrected BAM record, aborting.."); return 1; } }
  log("info", "Alignments processed: " + std::to_string(processed) + ", written: " + std::to_string(written) + " (XF 0/1/2: " +
                  std::to_string(counts[0]) + "/" + std::to_string(counts[1]) + "/" + std::to_string(counts[2]) + ")");
  char tb[160];
  snprintf(tb, sizeof(tb), "Stages: accuracy pass %.2f s, waiting for the reader %.2f s, smooth %.2f s, deflate + write %.2f s", t_pass1, t_read, t_smooth, t_write);
  log("info", tb);
  return 0;
}

}  // namespace svdss
