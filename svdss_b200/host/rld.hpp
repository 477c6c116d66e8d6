// ropebwt3 `.fmd` files: reader, BWT inversion, writer.  Host side, C++14.
//
// `SVDSS index -d` in the reference is ropebwt3's `build -d`: the BWT of {S_i $, rc(S_i) $} dumped
// in the "fermi delta" format of rld0.c/rld0.h (magic "RLD\3"), which rb3_fmi_restore
// (ping_pong.cpp:244-245) loads and rld_rank2a walks.  This header lets an index a user already
// has be used here: decode the run-length-delta stream to the BWT, invert the BWT to recover the
// indexed sequences, and hand the forward strands to svb_index_build (the GPU index needs the text
// for its located-match tables, and rebuilding takes seconds).  The writer goes the other way: a
// BWT built on the GPU dumped as `.fmd`.
//
// PARITY UNPINNED.  ropebwt3 (pin 0ea3919, CMakeLists.txt:154-156) is fetched by the reference's
// CMake and is absent from /root/reference and from this image, and no `.fmd` file exists here.
// The layout below is the published rld0 format restated from memory; the tests check the writer
// against the reader and against an independent Python decoder of the same statement, which
// catches coding slips but not a misremembered detail.
//
// Layout (little endian):
//   "RLD\3" | u32 asize<<16 | sbits | u64 reserved (0) | u64 n_bytes = 8 * n_words | u64 n_frames | u64 mcnt[asize] (symbols per code)
//   | u64 words[n_words] | u64 frame[n_frames][asize+1]
// words[] is cut into chunks of 2^23 words whose last word is never used, and into blocks of
// 2^sbits words.  A block starts with asize+1 counters -- [0] = symbols, [1+c] = symbols of code c
// encoded in the PREVIOUS block -- as u16 (2 words for asize 6), u32 (4 words) or u64 (7 words)
// depending on the total; the width code (0/1/2) sits in the top two bits of the block's first
// word.  After the counters: (run length, symbol) pairs, most significant bit first, the length
// in Elias delta code followed by the symbol in abits = ilog2(asize)+1 bits; a pair never
// straddles two blocks; the rest of a block is zero, and six zero bits where a pair should start
// end the block.  The stream ends with one block that holds only its counters.
// frame[k] = (word offset of a block that starts at or before symbol k << ibits, per-code symbol
// counts before that block), ibits = ilog2(n / n_blocks) + 4: the reader's entry points for rank.
#pragma once
#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

namespace svdss {

// fn(i) for i in [0, n), dealt in chunks to all hardware threads.  std::thread rather than OpenMP so that
// the same header serves the shell (g++ -fopenmp) and libsvdss_b200 (nvcc host pass, no OpenMP).
template <class F>
inline void rld_parallel_for(uint64_t n, uint64_t chunk, F fn) {
  unsigned nt = std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  const uint64_t n_chunks = (n + chunk - 1) / chunk;
  if (nt > n_chunks) nt = (unsigned)n_chunks;
  if (nt <= 1) { for (uint64_t i = 0; i < n; ++i) fn(i); return; }
  std::atomic<uint64_t> next(0);
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t)
    th.emplace_back([&]() {
      for (;;) {
        const uint64_t c = next.fetch_add(1);
        if (c >= n_chunks) return;
        const uint64_t hi = std::min(n, (c + 1) * chunk);
        for (uint64_t i = c * chunk; i < hi; ++i) fn(i);
      }
    });
  for (auto& x : th) x.join();
}

struct RldFile {
  int asize = 6, sbits = 3;
  uint64_t n_frames = 0;
  std::vector<uint64_t> mcnt;    // symbols per code
  std::vector<uint64_t> words;
  std::vector<uint64_t> frame;   // n_frames * (asize + 1)
  uint64_t n_symbols() const { uint64_t n = 0; for (uint64_t v : mcnt) n += v; return n; }
};

class Rld {
 public:
  static const int LBITS = 23;   // RLD_LBITS

  static bool is_rld(const std::string& path) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) return false;
    char m[4] = {0, 0, 0, 0};
    const bool ok = fread(m, 1, 4, f) == 4 && memcmp(m, "RLD\3", 4) == 0;
    fclose(f);
    return ok;
  }

  // rld_restore_header + rld_restore
  static bool read(const std::string& path, RldFile& r, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "cannot open " + path; return false; }
    char m[4];
    uint32_t a = 0;
    uint64_t n_words = 0;
    // rld_dump (rld0.c): magic, u32 asize << 16 | sbits, 8 bytes reserved for future use (written as 0), n_bytes of the
    // word stream (always a multiple of 8), n_frames; rld_restore_header reads the three 64-bit words in one go and
    // ignores the first.  Round 1 of this file wrote (n_words, n_frames) without the reserved word: a first word that is
    // not 0 marks such a file, which is still read (restated from memory like the rest -- ropebwt3 is not vendored).
    uint64_t h[3] = {0, 0, 0};
    long header_bytes = 32;
    bool ok = fread(m, 1, 4, f) == 4 && memcmp(m, "RLD\3", 4) == 0 && fread(&a, 4, 1, f) == 1 && fread(h, 8, 3, f) == 3;
    if (ok) {
      if (h[0] == 0) { ok = (h[1] & 7) == 0; n_words = h[1] >> 3; r.n_frames = h[2]; }
      else { n_words = h[0]; r.n_frames = h[1]; header_bytes = 24; ok = fseek(f, 24, SEEK_SET) == 0; }
    }
    if (ok) {
      r.asize = (int)(a >> 16); r.sbits = (int)(a & 0xffff);
      ok = r.asize >= 1 && r.asize <= 15 && r.sbits >= 3 && r.sbits <= 16 && n_words < ((uint64_t)1 << 40) && r.n_frames < ((uint64_t)1 << 40);
    }
    if (ok) {   // the sizes the header announces must fit the file (checked before anything is allocated)
      const long at = ftell(f);
      ok = at >= 0 && fseek(f, 0, SEEK_END) == 0;
      const long fsize = ok ? ftell(f) : -1;
      ok = ok && fsize >= 0 && fseek(f, at, SEEK_SET) == 0 &&
           (unsigned long long)fsize >= (unsigned long long)header_bytes + 8ull * (unsigned)r.asize + 8ull * n_words + 8ull * r.n_frames * (unsigned)(r.asize + 1);
    }
    if (ok) {
      r.mcnt.resize((size_t)r.asize);
      ok = fread(r.mcnt.data(), 8, (size_t)r.asize, f) == (size_t)r.asize;
    }
    if (ok) {
      r.words.resize((size_t)n_words + 1);   // one spare zero word: a pair may be composed from words[p] and words[p + 1]
      ok = fread(r.words.data(), 8, (size_t)n_words, f) == (size_t)n_words;
      r.words[(size_t)n_words] = 0;
    }
    if (ok) {
      r.frame.resize((size_t)(r.n_frames * (uint64_t)(r.asize + 1)));
      ok = r.frame.empty() || fread(r.frame.data(), 8, r.frame.size(), f) == r.frame.size();
    }
    fclose(f);
    if (!ok) err = path + " is not a complete RLD\\3 (ropebwt3 .fmd) file";
    return ok;
  }

  // the whole stream as symbols: n_symbols() bytes
  static bool decode_bwt(const RldFile& r, std::vector<uint8_t>& bwt, std::string& err) {
    const uint64_t ssize = (uint64_t)1 << r.sbits, n_words = r.words.size() - 1, last = n_words >> r.sbits << r.sbits;
    const size_t n_blocks = (size_t)(last >> r.sbits);   // blocks that carry pairs; the block at `last` only closes the stream
    // symbols of block b stand in the counters of block b + 1
    std::vector<uint64_t> start(n_blocks + 1, 0);
    for (size_t b = 0; b < n_blocks; ++b) start[b + 1] = start[b] + header_total(r.words[(size_t)(((uint64_t)b + 1) << r.sbits)]);
    const uint64_t n = r.n_symbols();
    if (start[n_blocks] != n) { err = "RLD block counters do not add up to the symbol counts of the header"; return false; }
    if (n > ((uint64_t)1 << 40)) { err = "RLD header announces more than 2^40 symbols"; return false; }
    try { bwt.assign((size_t)n, 0); } catch (const std::bad_alloc&) { err = "not enough host memory for a BWT of " + std::to_string(n) + " symbols"; return false; }
    std::atomic<int> bad(0);
    rld_parallel_for((uint64_t)n_blocks, 4096, [&](uint64_t b) {
      const uint64_t o = ((uint64_t)b) << r.sbits;
      uint64_t out = start[(size_t)b];
      const uint64_t out_end = start[(size_t)b + 1];
      BitPos p = first_pair(r, o, ssize);
      uint64_t per_code[16] = {0};
      bool ok = true;
      while (out < out_end) {
        int c;
        const int64_t l = dec0(r, p, c);
        if (l <= 0 || c >= r.asize || out + (uint64_t)l > out_end) { ok = false; break; }
        memset(bwt.data() + out, c, (size_t)l);
        out += (uint64_t)l;
        per_code[c] += (uint64_t)l;
      }
      for (int c = 0; ok && c < r.asize; ++c) ok = per_code[c] == header_count(r, o + ssize, c + 1);   // the next block's counters
      if (!ok) bad.fetch_add(1);
    });
    if (bad.load()) { err = "corrupt RLD block (pairs do not match the block counters)"; return false; }
    return true;
  }

  // rld_init(asize, sbits) + rld_enc per run + rld_enc_finish + rld_rank_index + rld_dump
  static bool write(const std::string& path, const uint8_t* bwt, uint64_t n, std::string& err, int asize = 6, int sbits = 3) {
    Encoder e(asize, sbits);
    for (uint64_t i = 0; i < n;) {
      uint64_t j = i + 1;
      while (j < n && bwt[j] == bwt[i]) ++j;
      if (bwt[i] >= asize) { err = "symbol out of range"; return false; }
      e.enc0((int64_t)(j - i), bwt[i]);
      i = j;
    }
    e.finish();
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) { err = "cannot write " + path; return false; }
    const uint32_t a = (uint32_t)asize << 16 | (uint32_t)sbits;
    const uint64_t n_words = e.n_words, n_frames = e.n_frames, reserved = 0, n_bytes = 8 * e.n_words;
    bool ok = fwrite("RLD\3", 1, 4, f) == 4 && fwrite(&a, 4, 1, f) == 1 && fwrite(&reserved, 8, 1, f) == 1 && fwrite(&n_bytes, 8, 1, f) == 1 &&
              fwrite(&n_frames, 8, 1, f) == 1 &&
              fwrite(e.mcnt.data() + 1, 8, (size_t)asize, f) == (size_t)asize &&
              fwrite(e.z.data(), 8, (size_t)n_words, f) == (size_t)n_words &&
              fwrite(e.frame.data(), 8, e.frame.size(), f) == e.frame.size();
    ok = (fclose(f) == 0) && ok;
    if (!ok) err = "short write on " + path;
    return ok;
  }

 private:
  struct BitPos { uint64_t p, stail; int r; };   // word, last word of the block that may hold pairs, bits left in words[p]

  static int ilog2(uint64_t v) { int l = -1; while (v) { v >>= 1; ++l; } return l; }
  static int offset0(int asize, int type) { const int a1 = asize + 1; return type == 0 ? (a1 * 16 + 63) / 64 : type == 1 ? (a1 * 32 + 63) / 64 : a1; }
  static uint64_t header_total(uint64_t w0) {
    const int type = (int)(w0 >> 62);
    return type == 0 ? (w0 & 0xffff) : type == 1 ? (w0 & 0xffffffffu) : (w0 & 0x3fffffffffffffffull);
  }
  static uint64_t header_count(const RldFile& r, uint64_t o, int j) {   // counter j of the block at word o
    const int type = (int)(r.words[(size_t)o] >> 62);
    if (type == 0) return r.words[(size_t)(o + (uint64_t)(j >> 2))] >> (16 * (j & 3)) & (j == 3 ? 0x3fffu : 0xffffu);
    if (type == 1) return r.words[(size_t)(o + (uint64_t)(j >> 1))] >> (32 * (j & 1)) & (j == 1 ? 0x3fffffffu : 0xffffffffu);
    return r.words[(size_t)(o + (uint64_t)j)];
  }
  // blocks are contiguous; the last word of a 2^23-word chunk belongs to no block's pairs (rld_get_stail)
  static uint64_t stail_of(uint64_t o, uint64_t ssize) {
    const uint64_t lmask = ((uint64_t)1 << LBITS) - 1;
    return o + ssize - ((((o + ssize) & lmask) == 0) ? 2 : 1);
  }
  static BitPos first_pair(const RldFile& r, uint64_t o, uint64_t ssize) {
    BitPos p;
    p.p = o + (uint64_t)offset0(r.asize, (int)(r.words[(size_t)o] >> 62));
    p.stail = stail_of(o, ssize);
    p.r = 64;
    return p;
  }
  // rld_dec0: one (length, symbol) pair; 0 at the end of a block
  static int64_t dec0(const RldFile& r, BitPos& it, int& c) {
    if (it.p > it.stail) return 0;   // the previous pair filled the block to its last bit
    const int abits = ilog2((uint64_t)r.asize) + 1;
    const uint64_t w0 = r.words[(size_t)it.p];
    const uint64_t x = (it.r == 64 ? w0 : w0 << (64 - it.r)) | ((it.p != it.stail && it.r != 64) ? r.words[(size_t)it.p + 1] >> it.r : 0);
    int w;
    int64_t y;
    if (x >> 63 == 0) {
      w = (int)(0x333333335555779bull >> (x >> 59 << 2) & 0xf);   // width of the gamma-coded bit count: 3, 5, 7, 9 or 11
      if (w == 0xb && x >> 58 == 0) return 0;
      const int l = (int)(x >> (64 - w)) - 1;                      // bits of the length below its leading one
      if (l < 1 || w + l + abits > 64) return -1;
      y = (int64_t)((x << w >> (64 - l)) | (uint64_t)1 << l);
      w += l;
    } else { w = 1; y = 1; }
    c = (int)(x << w >> (64 - abits));
    w += abits;
    if (it.r > w) it.r -= w;
    else { ++it.p; it.r = 64 + it.r - w; }
    return y;
  }

  struct Encoder {
    int asize, sbits, abits;
    uint64_t ssize;
    std::vector<uint64_t> z;          // all chunks back to back
    std::vector<uint64_t> cnt, mcnt;  // [0] = all symbols, [1 + c] = code c; cnt running, mcnt at the start of the open block
    uint64_t shead = 0, p = 0, stail = 0;
    int r = 64;
    uint64_t n_words = 0, n_frames = 0;
    std::vector<uint64_t> frame;
    Encoder(int asize_, int sbits_) : asize(asize_), sbits(sbits_), abits(ilog2((uint64_t)asize_) + 1), ssize((uint64_t)1 << sbits_),
                                      cnt((size_t)asize_ + 1, 0), mcnt((size_t)asize_ + 1, 0) {
      z.assign((size_t)ssize, 0);     // first block: counters all zero, u16 form
      shead = 0; p = (uint64_t)offset0(asize, 0); stail = stail_of(0, ssize); r = 64;
    }
    void next_block() {               // enc_next_block
      shead += ssize;
      z.resize((size_t)(shead + ssize), 0);
      const uint64_t tot = cnt[0] - mcnt[0];
      int type;
      if (tot < 0x4000) {
        type = 0;
        for (int i = 0; i <= asize; ++i) z[(size_t)(shead + (uint64_t)(i >> 2))] |= ((cnt[(size_t)i] - mcnt[(size_t)i]) & 0xffff) << (16 * (i & 3));
      } else if (tot < 0x40000000) {
        type = 1;
        for (int i = 0; i <= asize; ++i) z[(size_t)(shead + (uint64_t)(i >> 1))] |= ((cnt[(size_t)i] - mcnt[(size_t)i]) & 0xffffffffu) << (32 * (i & 1));
      } else {
        type = 2;
        for (int i = 0; i <= asize; ++i) z[(size_t)(shead + (uint64_t)i)] = cnt[(size_t)i] - mcnt[(size_t)i];
      }
      z[(size_t)shead] |= (uint64_t)type << 62;
      p = shead + (uint64_t)offset0(asize, type);
      stail = stail_of(shead, ssize);
      r = 64;
      mcnt = cnt;
    }
    void enc0(int64_t l, uint8_t c) {  // rld_enc0 with rld_delta_enc1
      const int y = ilog2((uint64_t)l), zz = ilog2((uint64_t)y + 1);
      int w = (zz << 1) + 1 + y;
      uint64_t x = (((uint64_t)l ^ (uint64_t)1 << y) | (uint64_t)(y + 1) << y) << abits | c;
      w += abits;
      // a pair that would fill the block to its last bit goes to the next block too: whatever a
      // reader does when its position runs into the next block's counters, it never has to
      if (w >= r && p == stail) next_block();
      if (w > r) {
        w -= r;
        z[(size_t)p++] |= x >> w;      // r == 0 (the word is full): w is the whole pair and x >> w == 0
        r = 64 - w;
        z[(size_t)p] |= x << r;
      } else { r -= w; z[(size_t)p] |= x << r; }
      cnt[0] += (uint64_t)l;
      cnt[(size_t)c + 1] += (uint64_t)l;
    }
    void finish() {                    // rld_enc_finish + rld_rank_index
      next_block();
      n_words = p;
      z.resize((size_t)n_words);
      // frames
      const uint64_t n = mcnt[0], n_blks = n_words * 8 * 8 / 64 / ssize + 1, last = n_words >> sbits << sbits;
      const int ibits = ilog2(n / n_blks) + 4;   // RLD_IBITS_PLUS
      n_frames = ((n + ((uint64_t)1 << ibits) - 1) >> ibits) + 1;
      const size_t a1 = (size_t)asize + 1;
      frame.assign((size_t)n_frames * a1, 0);
      // frame[k] = the block that holds symbol k << ibits (the last block starting at or before it)
      // and the per-code counts before that block; frames at or past the end point at the last
      // block with pairs.  c = counts before block `i`, `nxt` = counts before the block after it.
      std::vector<uint64_t> c((size_t)asize, 0), nxt((size_t)asize, 0);
      uint64_t k = 1;
      for (uint64_t i = 0; i < last; i += ssize) {
        const uint64_t h = i + ssize, w0 = z[(size_t)h];
        const int type = (int)(w0 >> 62);
        uint64_t s_next = 0;
        for (int j = 1; j <= asize; ++j) {
          uint64_t v;
          if (type == 0) v = z[(size_t)(h + (uint64_t)(j >> 2))] >> (16 * (j & 3)) & (j == 3 ? 0x3fffu : 0xffffu);
          else if (type == 1) v = z[(size_t)(h + (uint64_t)(j >> 1))] >> (32 * (j & 1)) & (j == 1 ? 0x3fffffffu : 0xffffffffu);
          else v = z[(size_t)(h + (uint64_t)j)];
          nxt[(size_t)j - 1] = c[(size_t)j - 1] + v;
          s_next += nxt[(size_t)j - 1];
        }
        const bool final_block = h == last;
        while (k < n_frames && ((k << ibits) < s_next || final_block)) {
          frame[(size_t)k * a1] = i;
          for (int j = 0; j < asize; ++j) frame[(size_t)k * a1 + 1 + (size_t)j] = c[(size_t)j];
          ++k;
        }
        c = nxt;
      }
    }
  };
};

// ---- BWT -> sequences -------------------------------------------------------------------------
// The BWT of a string collection with sentinels ordered by sequence (BCR, what ropebwt builds):
// row k < m (m = number of '$') is the suffix "$" that ends sequence k, so walking LF from row k
// until a '$' is read spells sequence k backwards.
class BwtInverter {
 public:
  BwtInverter(const uint8_t* bwt, uint64_t n) : bwt_(bwt), n_(n) {
    const uint64_t nb = (n + STEP - 1) / STEP + 1;
    occ_.assign((size_t)nb * 6, 0);
    // per-block histograms in parallel, then a serial prefix sum over the blocks
    rld_parallel_for(nb - 1, 8192, [&](uint64_t b) {
      uint64_t h[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      const uint64_t lo = (uint64_t)b * STEP, hi = std::min(n, lo + STEP);
      for (uint64_t i = lo; i < hi; ++i) ++h[bwt[i] & 7];
      for (int c = 0; c < 6; ++c) occ_[((size_t)b + 1) * 6 + (size_t)c] = h[c];
    });
    for (uint64_t b = 1; b < nb; ++b) for (int c = 0; c < 6; ++c) occ_[(size_t)b * 6 + (size_t)c] += occ_[(size_t)(b - 1) * 6 + (size_t)c];
    acc_[0] = 0;
    for (int c = 0; c < 6; ++c) acc_[c + 1] = acc_[c] + occ_[(size_t)(nb - 1) * 6 + (size_t)c];
  }
  uint64_t n_sequences() const { return acc_[1]; }
  uint64_t rank(int c, uint64_t i) const {   // occurrences of c in bwt[0, i)
    const uint64_t b = i / STEP;
    uint64_t v = occ_[(size_t)b * 6 + (size_t)c];
    for (uint64_t j = b * STEP; j < i; ++j) v += (bwt_[j] == c);
    return v;
  }
  // sequences as nt6 codes, in collection order; false if a walk does not end (not a BWT of '$'-terminated strings)
  bool sequences(std::vector<std::string>& out) const {
    const uint64_t m = n_sequences();
    out.assign((size_t)m, std::string());
    std::atomic<int> bad(0);
    rld_parallel_for(m, 1, [&](uint64_t k) {
      std::string& s = out[(size_t)k];
      uint64_t i = (uint64_t)k, steps = 0;
      while (true) {
        const int c = bwt_[i];
        if (c == 0) break;
        if (c > 5 || ++steps > n_) { bad.fetch_add(1); break; }
        s.push_back((char)c);
        i = acc_[c] + rank(c, i);
      }
      std::reverse(s.begin(), s.end());
    });
    return bad.load() == 0;
  }
 private:
  static const uint64_t STEP = 128;
  const uint8_t* bwt_;
  uint64_t n_;
  std::vector<uint64_t> occ_;
  uint64_t acc_[7];
};

inline std::string nt6_revcomp(const std::string& s) {
  std::string r(s.rbegin(), s.rend());
  for (auto& ch : r) if (ch >= 1 && ch <= 4) ch = (char)(5 - ch);
  return r;
}

// One strand of every (S, rc(S)) pair of a collection that holds both, which is what ropebwt3 build
// indexes by default: sequence 2i and 2i+1 when they pair up that way, otherwise matched through a
// hash of the reverse complement.  false if some sequence has no partner (index built with -R).
inline bool forward_strands(const std::vector<std::string>& seqs, std::vector<size_t>& keep) {
  keep.clear();
  bool adjacent = seqs.size() % 2 == 0;
  for (size_t i = 0; adjacent && i + 1 < seqs.size(); i += 2) adjacent = seqs[i + 1] == nt6_revcomp(seqs[i]);
  if (adjacent) { for (size_t i = 0; i < seqs.size(); i += 2) keep.push_back(i); return true; }
  std::unordered_multimap<std::string, size_t> open;   // sequences waiting for their reverse complement
  std::vector<char> used(seqs.size(), 0);
  for (size_t i = 0; i < seqs.size(); ++i) {
    auto it = open.find(seqs[i]);
    if (it != open.end()) { used[i] = 1; open.erase(it); continue; }   // i is the partner of an earlier sequence
    open.emplace(nt6_revcomp(seqs[i]), i);
    keep.push_back(i);
  }
  return open.empty();
}

}  // namespace svdss
