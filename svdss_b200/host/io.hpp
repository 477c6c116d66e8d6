// Host-side input readers for the SVDSS shell: FASTA/FASTQ (zlib) and a minimal BGZF/BAM reader.
// htslib is not available offline, so the fields the search path reads (SURVEY appendix B:
// flag, l_qseq, tid, 4-bit seq, qname, aux XF:i / HP:i) are parsed here from the BAM spec.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace svdss {

// ping_pong.hpp:46-52 seq_nt6_table ($=0 A=1 C=2 G=3 T=4 other=5); rb3_char2nt6 is the same map
inline const uint8_t* nt6_table() {
  static uint8_t t[256];
  static bool init = false;
  if (!init) {
    memset(t, 5, sizeof(t));
    t[0] = 0;
    t[(int)'A'] = t[(int)'a'] = 1; t[(int)'C'] = t[(int)'c'] = 2;
    t[(int)'G'] = t[(int)'g'] = 3; t[(int)'T'] = t[(int)'t'] = 4;
    init = true;
  }
  return t;
}

// buffered gz line/byte source (works on plain files too; BGZF is multi-member gzip)
class GzSource {
 public:
  explicit GzSource(const std::string& path) : f_(gzopen(path.c_str(), "rb")) {
    if (f_) gzbuffer(f_, 1 << 20);
  }
  ~GzSource() { if (f_) gzclose(f_); }
  bool ok() const { return f_ != nullptr; }
  bool read_exact(void* dst, size_t n) {
    uint8_t* p = static_cast<uint8_t*>(dst);
    while (n) {
      int got = gzread(f_, p, (unsigned)(n > (1u << 30) ? (1u << 30) : n));
      if (got <= 0) return false;
      p += got; n -= (size_t)got;
    }
    return true;
  }
  bool getline(std::string& line) {
    line.clear();
    char buf[1 << 16];
    while (gzgets(f_, buf, sizeof(buf))) {
      size_t l = strlen(buf);
      line.append(buf, l);
      if (l && buf[l - 1] == '\n') break;
    }
    if (line.empty() && gzeof(f_)) return false;
    while (!line.empty() && (line.back() == '\n' || line.back() == '\r')) line.pop_back();
    return true;
  }
 private:
  gzFile f_;
};

struct FastxRecord { std::string name, seq; };

// kseq-style reader (fastq.hpp / kseq.h): multi-line FASTA and FASTQ
class FastxReader {
 public:
  explicit FastxReader(const std::string& path) : src_(path) {}
  bool ok() const { return src_.ok(); }
  bool next(FastxRecord& r) {
    std::string line;
    if (pending_.empty()) {
      do { if (!src_.getline(line)) return false; } while (line.empty() || (line[0] != '>' && line[0] != '@'));
    } else { line.swap(pending_); pending_.clear(); }
    const bool fq = line[0] == '@';
    size_t e = 1;
    while (e < line.size() && !isspace((unsigned char)line[e])) ++e;
    r.name.assign(line, 1, e - 1);
    r.seq.clear();
    while (src_.getline(line)) {
      if (!fq && !line.empty() && line[0] == '>') { pending_ = line; return true; }
      if (fq && !line.empty() && line[0] == '+') {
        size_t need = r.seq.size(), got = 0;
        while (got < need && src_.getline(line)) got += line.size();
        return true;
      }
      r.seq += line;
    }
    return true;
  }
 private:
  GzSource src_;
  std::string pending_;
};

struct BamRecord {
  int32_t tid = -1, pos = 0, l_qseq = 0;
  uint16_t flag = 0;
  uint8_t mapq = 0;
  std::string qname;
  std::vector<uint8_t> nt6;   // decoded sequence, nt6 codes (ping_pong.cpp:90-94)
  std::vector<uint32_t> cigar; // len<<4 | op (MIDNSHP=X), as stored
  std::vector<uint8_t> seq4;   // 4-bit packed sequence, as stored (Clusterer decodes it on demand)
  bool has_xf = false, has_hp = false;
  int64_t xf = 0, hp = 0;
  // bam_endpos: pos + reference span of the CIGAR (M,D,N,=,X); pos+1 for an empty span
  int32_t endpos() const {
    int64_t span = 0;
    for (uint32_t c : cigar) { const uint32_t op = c & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) span += c >> 4; }
    return (int32_t)(pos + (span ? span : 1));
  }
};

// htslib seq_nt16_str: ASCII of one 4-bit code
inline char nt16_char(unsigned c) { return "=ACMGRSVTWYHKDBN"[c & 15]; }

class BamReader {
 public:
  explicit BamReader(const std::string& path) : src_(path) {
    if (!src_.ok()) return;
    char magic[4];
    int32_t l_text = 0, n_ref = 0;
    if (!src_.read_exact(magic, 4) || memcmp(magic, "BAM\1", 4) != 0) return;
    if (!src_.read_exact(&l_text, 4) || l_text < 0) return;
    text_.resize((size_t)l_text);
    if (l_text && !src_.read_exact(&text_[0], (size_t)l_text)) return;
    if (!src_.read_exact(&n_ref, 4) || n_ref < 0) return;
    for (int i = 0; i < n_ref; ++i) {
      int32_t l_name = 0, l_ref = 0;
      if (!src_.read_exact(&l_name, 4) || l_name <= 0) return;
      std::string nm((size_t)l_name, '\0');
      if (!src_.read_exact(&nm[0], (size_t)l_name) || !src_.read_exact(&l_ref, 4)) return;
      nm.resize(strlen(nm.c_str()));
      ref_names_.push_back(nm); ref_lens_.push_back(l_ref);
    }
    ok_ = true;
  }
  bool ok() const { return ok_; }
  const std::vector<std::string>& ref_names() const { return ref_names_; }
  // alignment mode (`call`): keep CIGAR + packed sequence instead of decoding nt6 codes
  void want_alignment(bool on) { want_align_ = on; }
  // 1 = record, 0 = clean EOF, -1 = truncated/corrupt
  int next(BamRecord& r) {
    int32_t bs = 0;
    if (!src_.read_exact(&bs, 4)) return 0;
    if (bs < 32) return -1;
    buf_.resize((size_t)bs);
    if (!src_.read_exact(buf_.data(), (size_t)bs)) return -1;
    const uint8_t* p = buf_.data();
    auto rd32 = [&](size_t o) { int32_t v; memcpy(&v, p + o, 4); return v; };
    auto rd16 = [&](size_t o) { uint16_t v; memcpy(&v, p + o, 2); return v; };
    r.tid = rd32(0); r.pos = rd32(4);
    const uint8_t l_read_name = p[8];
    r.mapq = p[9];
    const uint16_t n_cigar = rd16(12);
    r.flag = rd16(14);
    r.l_qseq = rd32(16);
    size_t o = 32;
    if (o + l_read_name > (size_t)bs) return -1;
    r.qname.assign((const char*)p + o, l_read_name ? l_read_name - 1 : 0);
    o += l_read_name;
    if (o + (size_t)n_cigar * 4 > (size_t)bs) return -1;
    if (want_align_) { r.cigar.resize(n_cigar); if (n_cigar) memcpy(r.cigar.data(), p + o, (size_t)n_cigar * 4); }
    o += (size_t)n_cigar * 4;
    const size_t seq_bytes = ((size_t)r.l_qseq + 1) / 2;
    if (r.l_qseq < 0 || o + seq_bytes + (size_t)r.l_qseq > (size_t)bs) return -1;
    static const char nt16[] = "=ACMGRSVTWYHKDBN";  // htslib seq_nt16_str
    const uint8_t* t6 = nt6_table();
    if (want_align_) {
      r.seq4.assign(p + o, p + o + seq_bytes);
    } else {
      r.nt6.resize((size_t)r.l_qseq);
      for (int32_t i = 0; i < r.l_qseq; ++i) {
        const uint8_t b = p[o + (i >> 1)];
        r.nt6[i] = t6[(int)nt16[(i & 1) ? (b & 0xf) : (b >> 4)]];
      }
    }
    o += seq_bytes + (size_t)r.l_qseq;
    r.has_xf = r.has_hp = false; r.xf = r.hp = 0;
    // aux fields (bam_aux_get + bam_aux2i for XF / HP, ping_pong.cpp:196-201)
    while (o + 3 <= (size_t)bs) {
      const char t0 = (char)p[o], t1 = (char)p[o + 1], ty = (char)p[o + 2];
      o += 3;
      int64_t iv = 0; bool is_int = false;
      switch (ty) {
        case 'A': o += 1; break;
        case 'c': iv = (int8_t)p[o]; is_int = true; o += 1; break;
        case 'C': iv = p[o]; is_int = true; o += 1; break;
        case 's': { int16_t v; memcpy(&v, p + o, 2); iv = v; is_int = true; o += 2; break; }
        case 'S': { uint16_t v; memcpy(&v, p + o, 2); iv = v; is_int = true; o += 2; break; }
        case 'i': { int32_t v; memcpy(&v, p + o, 4); iv = v; is_int = true; o += 4; break; }
        case 'I': { uint32_t v; memcpy(&v, p + o, 4); iv = v; is_int = true; o += 4; break; }
        case 'f': o += 4; break;
        case 'd': o += 8; break;
        case 'Z': case 'H': while (o < (size_t)bs && p[o]) ++o; ++o; break;
        case 'B': {
          if (o + 5 > (size_t)bs) return -1;
          const char st = (char)p[o]; int32_t cnt; memcpy(&cnt, p + o + 1, 4);
          const size_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : 4;
          o += 5 + es * (size_t)(cnt < 0 ? 0 : cnt);
          break;
        }
        default: return -1;
      }
      if (o > (size_t)bs) return -1;
      if (is_int && t0 == 'X' && t1 == 'F') { r.has_xf = true; r.xf = iv; }
      if (is_int && t0 == 'H' && t1 == 'P') { r.has_hp = true; r.hp = iv; }
    }
    return 1;
  }
 private:
  GzSource src_;
  bool ok_ = false, want_align_ = false;
  std::string text_;
  std::vector<std::string> ref_names_;
  std::vector<int32_t> ref_lens_;
  std::vector<uint8_t> buf_;
};

}  // namespace svdss
