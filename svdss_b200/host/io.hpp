// Host-side input readers for the SVDSS shell: FASTA/FASTQ (zlib) and a minimal BGZF/BAM reader.
// htslib is not available offline, so the fields the search path reads (SURVEY appendix B:
// flag, l_qseq, tid, 4-bit seq, qname, aux XF:i / HP:i) are parsed here from the BAM spec.
#pragma once
#include <zlib.h>

#include <cstdint>
#include <cstdio>
#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <future>
#include <memory>
#include <string>
#include <vector>

#include "../../include/svdss_b200.h"

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

// Where BgzfSource inflates: -1 = host threads (zlib), d >= 0 = device d through svb_bgzf_inflate_device
// (`--gpu-inflate` / SVB_BGZF_GPU=<device> in the CLI).  Process-wide; set before the readers are opened.
inline int& bgzf_gpu_device() { static int d = -1; return d; }

// window buffer of BgzfSource: like a byte vector without the zero fill, in pinned host memory when the
// device inflates (the copies run at PCIe speed only from / to pinned buffers)
class ByteBuf {
 public:
  ByteBuf() = default;
  ByteBuf(const ByteBuf&) = delete;
  ByteBuf& operator=(const ByteBuf&) = delete;
  ~ByteBuf() { release(p_, pinned_); }
  void pin(bool on) { if (on != want_pin_ && cap_ == 0) want_pin_ = on; }
  uint8_t* data() { return p_; }
  const uint8_t* data() const { return p_; }
  size_t size() const { return n_; }
  void clear() { n_ = 0; }
  // capacity up front: growing a pinned buffer step by step would pin it again every time
  bool reserve(size_t n) { const size_t keep = n_; if (n > cap_ && !resize(n)) return false; n_ = keep; return true; }
  // keeps the first min(size, n) bytes; false when memory runs out
  bool resize(size_t n) {
    if (n > cap_) {
      const size_t cap = std::max(n, cap_ + cap_ / 2);
      bool pinned = want_pin_;
      uint8_t* q = pinned ? static_cast<uint8_t*>(svb_host_alloc_pinned(cap)) : nullptr;
      if (!q) { pinned = false; q = static_cast<uint8_t*>(malloc(cap)); }   // unpinned still works, the copies are slower
      if (!q) return false;
      if (n_) memcpy(q, p_, n_);
      release(p_, pinned_);
      p_ = q; cap_ = cap; pinned_ = pinned;
    }
    n_ = n;
    return true;
  }
  void swap(ByteBuf& o) { std::swap(p_, o.p_); std::swap(n_, o.n_); std::swap(cap_, o.cap_); std::swap(pinned_, o.pinned_); std::swap(want_pin_, o.want_pin_); }
 private:
  static void release(uint8_t* p, bool pinned) { if (!p) return; if (pinned) svb_host_free_pinned(p); else free(p); }
  uint8_t* p_ = nullptr;
  size_t n_ = 0, cap_ = 0;
  bool pinned_ = false, want_pin_ = false;
};

// BGZF byte source with parallel inflate: BGZF members (<= 64 KiB each, size in the BC extra field)
// are independent deflate streams, so a window of them is read sequentially and inflated by all
// host threads at once -- the reference does the same through htslib's bgzf_mt(.., 8, ..)
// (ping_pong.cpp:249, clusterer.cpp:13).  With bgzf_gpu_device() >= 0 the window goes to the device instead
// (k_bgzf_inflate_warp, one warp per member), pinned window buffers.  Files that are not BGZF go through plain zlib.
class BgzfSource {
 public:
  // window: compressed bytes per window, 0 = the default (a reader that only wants the BAM header asks for a small one)
  explicit BgzfSource(const std::string& path, size_t window = 0) : f_(fopen(path.c_str(), "rb")), gpu_(bgzf_gpu_device()), window_(window) {
    if (!f_) return;
    uint8_t h[18];
    const size_t got = fread(h, 1, 18, f_);
    bgzf_ = got == 18 && h[0] == 31 && h[1] == 139 && h[2] == 8 && (h[3] & 4) && h[12] == 'B' && h[13] == 'C';
    if (bgzf_ && gpu_ >= 0) {
      // pinning the window buffers and the first CUDA call cost ~1.5 s: below a gigabyte of file the host threads win
      long long min_bytes = 1ll << 30;
      if (const char* e = getenv("SVB_BGZF_GPU_MIN_BYTES")) min_bytes = atoll(e);
      fseek(f_, 0, SEEK_END);
      if ((long long)ftell(f_) < min_bytes) gpu_ = -1;
    }
    if (bgzf_) fseek(f_, 0, SEEK_SET);
    else { fclose(f_); f_ = nullptr; plain_.reset(new GzSource(path)); }
  }
  ~BgzfSource() {
    if (pending_.valid()) pending_.wait();
    if (reading_.valid()) reading_.wait();
    if (f_) fclose(f_);
    if (bgzf_ && getenv("SVB_BGZF_STATS"))
      fprintf(stderr, "[svdss] BGZF reader: %llu windows, %.3f s reading members (%.3f s of it pinning the window buffers), %.3f s inflating (%s), %.3f s the consumer waited\n",
              (unsigned long long)n_windows_, t_read_ + t_pin_out_, t_pin_ + t_pin_out_, t_inflate_, gpu_ >= 0 ? "device" : "host threads", t_wait_);
  }
  bool ok() const { return bgzf_ ? f_ != nullptr : (plain_ && plain_->ok()); }
  bool read_exact(void* dst, size_t n) {
    if (!bgzf_) return plain_->read_exact(dst, n);
    uint8_t* p = static_cast<uint8_t*>(dst);
    while (n) {
      if (pos_ == out_.size() && !refill()) return false;
      const size_t k = std::min(n, out_.size() - pos_);
      memcpy(p, out_.data() + pos_, k);
      p += k; pos_ += k; n -= k;
    }
    return true;
  }
 public:
  // Raw mode, for a consumer that inflates elsewhere (svb_bamstream_window: records decoded on the device): the next
  // window of whole BGZF members as they lie in the file -- base[in_offs[m]] is the first deflate byte of member m
  // (trailer and next header ride along), out_offs its place in the window's payload.  The window after it is read
  // by a helper thread meanwhile.  false at EOF or on a malformed file.  Not to be mixed with read_exact / peek.
  bool next_members(const uint8_t*& base, std::vector<int64_t>& in_offs, std::vector<int64_t>& out_offs) {
    if (!bgzf_) return false;
    for (;;) {
      Window& w = win_[seq_ & 1];
      const auto w0 = Clock::now();
      const bool have = reading_.valid() ? reading_.get() : read_window(w);
      t_wait_ += since(w0);
      if (!have) return false;
      ++seq_;
      Window& next = win_[seq_ & 1];
      reading_ = std::async(std::launch::async, [this, &next]() { return read_window(next); });
      if (w.out_total == 0) continue;             // only empty members (the EOF marker)
      const size_t nb = w.blks.size(), b0 = w.blks[0].in_off;
      in_offs.resize(nb + 1); out_offs.resize(nb + 1);
      for (size_t i = 0; i < nb; ++i) { in_offs[i] = (int64_t)(w.blks[i].in_off - b0); out_offs[i] = (int64_t)w.blks[i].out_off; }
      in_offs[nb] = (int64_t)(w.in_end - b0); out_offs[nb] = (int64_t)w.out_total;
      base = w.in.data() + b0;
      return true;
    }
  }
  bool is_bgzf() const { return bgzf_; }
  bool failed() const { return bad_; }              // the last false of next_members was an error, not the end of the file
  bool device_inflate() const { return bgzf_ && gpu_ >= 0; }
  // zero-copy access for record parsers: n contiguous bytes of the current window, or nullptr if the
  // window ends before (then read_exact copies across the boundary)
  const uint8_t* peek(size_t n) const { return bgzf_ && out_.size() - pos_ >= n ? out_.data() + pos_ : nullptr; }
  void skip(size_t n) { pos_ += n; }
 private:
  struct Blk { size_t in_off, in_len, out_off, out_len; };
  // a window of the file as read: whole BGZF members, blks[] = where their deflate data lie and where their payload goes
  struct Window {
    ByteBuf in;
    std::vector<Blk> blks;
    size_t in_end = 0, out_total = 0;
  };
  using Clock = std::chrono::steady_clock;
  static double since(Clock::time_point t) { return std::chrono::duration<double>(Clock::now() - t).count(); }

  // Three stages run side by side: the caller parses window k, a background task inflates window k + 1, and that
  // task's own helper reads window k + 2 from the file.
  bool refill() {
    bool ok;
    const auto w0 = Clock::now();
    if (pending_.valid()) ok = pending_.get();
    else ok = next_window(out_next_);
    t_wait_ += since(w0);
    if (!ok) { out_.clear(); pos_ = 0; return false; }
    out_.swap(out_next_);
    pos_ = 0;
    pending_ = std::async(std::launch::async, [this]() { return next_window(out_next_); });
    return true;
  }
  // the next window with any payload, inflated into `out`.  false at EOF (no payload left) or on a malformed file.
  bool next_window(ByteBuf& out) {
    for (;;) {
      Window& w = win_[seq_ & 1];
      const bool have = reading_.valid() ? reading_.get() : read_window(w);
      if (!have) return false;
      ++seq_;
      // the other buffer held the window before this one, inflated long ago: read the window after this one into it
      Window& next = win_[seq_ & 1];
      reading_ = std::async(std::launch::async, [this, &next]() { return read_window(next); });
      if (!inflate_window(w, out)) return false;
      if (w.out_total > 0) return true;
      // a window holding only empty members (the EOF marker): keep reading
    }
  }
  size_t window_bytes() const {
    size_t window = window_ ? window_ : (size_t)64 << 20;   // compressed bytes per window (~5 k members: one wave of warps on the device)
    if (const char* e = getenv("SVB_BGZF_WINDOW")) { const long long v = atoll(e); if (v > 0) window = (size_t)v; }   // tests: force records across windows
    return window;
  }
  // device: pinned buffers of a fixed size, pinned once (pinning costs ~1 s per GB here); a window ends where its
  // payload would not fit and the rest of the read waits in carry_
  size_t payload_limit(size_t window) const {
    size_t cap = gpu_ >= 0 ? std::max<size_t>(window * 6, (size_t)1 << 20) : ~(size_t)0;   // BAM inflates 3-6x
    if (const char* e = getenv("SVB_BGZF_PAYLOAD")) { const long long v = atoll(e); if (v >= (1 << 16)) cap = (size_t)v; }   // tests: the limit on the host path too
    return cap;
  }
  // one window: up to 64 MiB of the file in one read, its BGZF members found in memory (a member the read cut in
  // two waits in carry_ for the next window).  false at EOF or on a malformed file.  Only one call at a time
  // (next_window waits for a read before it starts the next): carry_, eof_ and the file position are its own.
  bool read_window(Window& w) {
    const auto c0 = Clock::now();
    auto bad = [this]() { bad_ = true; return false; };   // malformed or truncated, as opposed to the end of the file
    ByteBuf& in = w.in;
    const size_t window = window_bytes(), out_cap = payload_limit(window);
    if (gpu_ >= 0) {
      const auto p0 = Clock::now();
      in.pin(true);
      if (!in.reserve(window + ((size_t)4 << 20))) return false;   // a longer carry (payload limit hit early) pins a bigger buffer once
      t_pin_ += since(p0);
    }
    bool starved = false;
    for (;;) {
      w.blks.clear();
      w.out_total = 0;
      size_t have = carry_.size();
      in.clear();                                   // nothing of the last window is kept (a growing buffer would copy it)
      // a carry as long as a window (payload limit hit early): no read this time -- unless it holds no whole member
      const size_t want = starved ? window : (window > have ? window - have : 0);
      starved = false;
      if (!in.resize(have + want + 16)) return false;
      if (have) memcpy(in.data(), carry_.data(), have);
      carry_.clear();
      if (!eof_ && want) {
        const size_t got = fread(in.data() + have, 1, want, f_);
        if (got < want) eof_ = true;
        have += got;
      }
      size_t p = 0;
      bool full = false;                            // the payload limit ended the window, not the read
      while (p + 18 <= have) {
        const uint8_t* h = in.data() + p;
        if (h[0] != 31 || h[1] != 139 || !(h[3] & 4)) return bad();
        // walk the extra field for the BC subfield (SAM spec 4.1)
        const size_t xlen = h[10] | (h[11] << 8);
        if (p + 12 + xlen > have) break;
        const uint8_t* extra = h + 12;
        int bsize = -1;
        for (size_t o = 0; o + 4 <= xlen;) {
          const unsigned sl = extra[o + 2] | (extra[o + 3] << 8);
          if (extra[o] == 'B' && extra[o + 1] == 'C' && sl == 2 && o + 6 <= xlen) bsize = extra[o + 4] | (extra[o + 5] << 8);
          o += 4 + sl;
        }
        if (bsize < 0) return bad();
        const size_t total = (size_t)bsize + 1, head = 12 + xlen;
        if (total < head + 8) return bad();
        if (p + total > have) break;
        uint32_t isize;
        memcpy(&isize, h + total - 4, 4);
        if (isize > (1u << 16)) return bad();
        if (w.out_total + isize > out_cap) { full = true; break; }
        w.blks.push_back(Blk{p + head, total - head - 8, w.out_total, isize});   // deflate data; CRC32 + ISIZE follow
        w.out_total += isize;
        p += total;
      }
      if (p < have) {                               // the window ends inside a member, or at the payload limit
        if (eof_ && !full) return bad();            // truncated file
        carry_.assign(in.data() + p, in.data() + have);
      }
      if (w.blks.empty()) {
        if (eof_) return false;
        starved = true;
        continue;                                   // a window smaller than one member (tests): read on
      }
      w.in_end = p;
      t_read_ += since(c0);
      ++n_windows_;
      return true;
    }
  }
  // the members of a window inflated in parallel: by the device, or by all host threads
  bool inflate_window(Window& w, ByteBuf& out) {
    const auto c0 = Clock::now();
    struct Tick { Clock::time_point a; double& acc; ~Tick() { acc += since(a); } } tick{c0, t_inflate_};
    ByteBuf& in = w.in;
    const std::vector<Blk>& blks = w.blks;
    out.clear();
    if (gpu_ >= 0) {
      const auto p0 = Clock::now();
      out.pin(true);
      if (!out.reserve(payload_limit(window_bytes()))) return false;
      const double pin_s = since(p0);
      t_pin_out_ += pin_s; t_inflate_ -= pin_s;
    }
    if (!out.resize(w.out_total)) return false;
    if (gpu_ >= 0) {
      // the members as they lie in the window, gzip trailer and the next header still behind every deflate stream
      // (the kernel stops at the final block of a stream and checks the payload size)
      std::vector<int64_t> io(blks.size() + 1), oo(blks.size() + 1);
      const size_t base = blks[0].in_off;         // member m: from its deflate data to the next member's (trailer and header ride along)
      for (size_t i = 0; i < blks.size(); ++i) { io[i] = (int64_t)(blks[i].in_off - base); oo[i] = (int64_t)blks[i].out_off; }
      io[blks.size()] = (int64_t)(w.in_end - base); oo[blks.size()] = (int64_t)w.out_total;
      if (svb_bgzf_inflate_device(in.data() + base, io.data(), oo.data(), (int64_t)blks.size(), gpu_, out.data(), nullptr, nullptr) != SVB_OK) {
        fprintf(stderr, "[svdss] BGZF inflate on device %d: %s\n", gpu_, svb_last_error());
        return false;
      }
      return true;
    }
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 8) reduction(+ : bad)
    for (long long i = 0; i < (long long)blks.size(); ++i) {
      const Blk& b = blks[(size_t)i];
      if (b.out_len == 0) continue;
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      if (inflateInit2(&zs, -15) != Z_OK) { ++bad; continue; }
      zs.next_in = in.data() + b.in_off; zs.avail_in = (uInt)b.in_len;
      zs.next_out = out.data() + b.out_off; zs.avail_out = (uInt)b.out_len;
      const int rc = inflate(&zs, Z_FINISH);
      if (rc != Z_STREAM_END || zs.avail_out != 0) ++bad;
      inflateEnd(&zs);
    }
    return bad == 0;
  }
  FILE* f_ = nullptr;
  bool bgzf_ = false;
  std::unique_ptr<GzSource> plain_;
  int gpu_ = -1;
  size_t window_ = 0;
  double t_read_ = 0, t_pin_ = 0, t_inflate_ = 0, t_pin_out_ = 0, t_wait_ = 0;   // SVB_BGZF_STATS=1 (the first two written by the reading task, the next two by the inflating one)
  unsigned long long n_windows_ = 0;
  ByteBuf out_, out_next_;
  Window win_[2];                                   // window being inflated / window being read
  unsigned long long seq_ = 0;                      // windows handed to inflate_window so far
  std::vector<uint8_t> carry_;                      // head of the member the last read cut
  bool eof_ = false, bad_ = false;
  std::future<bool> pending_, reading_;
  size_t pos_ = 0;
};

// BGZF writer with parallel deflate (the reference writes its smoothed BAM through htslib with
// bgzf_mt(.., 8, ..), smoother.cpp:362-363): payload is cut into 0xff00-byte members like htslib,
// a window of them is deflated by all host threads, then written in order; close() adds the EOF member.
class BgzfWriter {
 public:
  explicit BgzfWriter(FILE* f, int level = 6) : f_(f), level_(level) {}
  bool write(const void* p, size_t n) {
    const uint8_t* b = static_cast<const uint8_t*>(p);
    buf_.insert(buf_.end(), b, b + n);
    return buf_.size() < ((size_t)64 << 20) || flush(false);
  }
  bool close() {
    if (!flush(true)) return false;
    static const uint8_t eof[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    return fwrite(eof, 1, 28, f_) == 28 && fflush(f_) == 0;
  }
 private:
  bool flush(bool all) {
    const size_t BS = 0xff00;
    const size_t nblk = all ? (buf_.size() + BS - 1) / BS : buf_.size() / BS;
    if (nblk == 0) return true;
    std::vector<std::vector<uint8_t>> out(nblk);
    int bad = 0;
#pragma omp parallel for schedule(dynamic, 4) reduction(+ : bad)
    for (long long i = 0; i < (long long)nblk; ++i) {
      const size_t o = (size_t)i * BS, len = std::min(BS, buf_.size() - o);
      std::vector<uint8_t>& m = out[(size_t)i];
      m.resize(18 + compressBound((uLong)len) + 8);
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      if (deflateInit2(&zs, level_, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { ++bad; continue; }
      zs.next_in = const_cast<uint8_t*>(buf_.data() + o); zs.avail_in = (uInt)len;
      zs.next_out = m.data() + 18; zs.avail_out = (uInt)(m.size() - 26);
      const int rc = deflate(&zs, Z_FINISH);
      const size_t clen = zs.total_out;
      deflateEnd(&zs);
      if (rc != Z_STREAM_END || 18 + clen + 8 > 65536) { ++bad; continue; }
      const uint8_t head[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};
      memcpy(m.data(), head, 16);
      const uint16_t bsize = (uint16_t)(18 + clen + 8 - 1);
      memcpy(m.data() + 16, &bsize, 2);
      const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), buf_.data() + o, (uInt)len), isz = (uint32_t)len;
      memcpy(m.data() + 18 + clen, &crc, 4);
      memcpy(m.data() + 18 + clen + 4, &isz, 4);
      m.resize(18 + clen + 8);
    }
    if (bad) return false;
    for (const auto& m : out) if (fwrite(m.data(), 1, m.size(), f_) != m.size()) return false;
    buf_.erase(buf_.begin(), buf_.begin() + (std::ptrdiff_t)std::min(buf_.size(), nblk * BS));
    return true;
  }
  FILE* f_;
  int level_;
  std::vector<uint8_t> buf_;
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
  const uint8_t* seq4_view = nullptr;   // view mode (BamReader::want_view): the packed sequence where it lies in the window; valid until the next next()
  bool has_xf = false, has_hp = false;
  bool cg_cigar = false;       // `cigar` came from the CG:B,I tag of a record with more than 65535 ops
  size_t cg_off = 0, cg_len = 0;   // where that tag sits in `raw` (offset of its two-letter name, bytes incl. name and type)
  int64_t xf = 0, hp = 0;
  // raw mode (`smooth`): the record body as stored (without its 4-byte length) and where its parts start
  std::vector<uint8_t> raw;
  size_t off_cigar = 0, off_seq = 0, off_qual = 0, off_aux = 0;
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
  explicit BamReader(const std::string& path, size_t window = 0) : src_(path, window) {
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
      header_bytes_ += 8 + (int64_t)l_name;
    }
    header_bytes_ += 12 + (int64_t)l_text;
    ok_ = true;
  }
  // inflated bytes before the first record (magic, text, reference dictionary)
  int64_t header_bytes() const { return header_bytes_; }
  bool ok() const { return ok_; }
  const std::vector<std::string>& ref_names() const { return ref_names_; }
  // alignment mode (`call`): keep CIGAR + packed sequence instead of decoding nt6 codes
  void want_alignment(bool on) { want_align_ = on; }
  // view mode (`search`): no copy of the packed sequence at all -- BamRecord::seq4_view points into the inflated window
  void want_view(bool on) { want_view_ = on; if (on) want_align_ = true; }
  // raw mode (`smooth`): additionally keep the whole record body
  void want_raw(bool on) { want_raw_ = on; if (on) want_align_ = true; }
  const std::string& header_text() const { return text_; }
  const std::vector<int32_t>& ref_lens() const { return ref_lens_; }
  // 1 = record, 0 = clean EOF, -1 = truncated/corrupt
  int next(BamRecord& r) {
    int32_t bs = 0;
    if (!src_.read_exact(&bs, 4)) return 0;
    if (bs < 32) return -1;
    const uint8_t* p = src_.peek((size_t)bs);   // the record as it lies in the inflated window
    if (p) src_.skip((size_t)bs);
    else {
      buf_.resize((size_t)bs);
      if (!src_.read_exact(buf_.data(), (size_t)bs)) return -1;
      p = buf_.data();
    }
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
    r.off_cigar = o;
    if (o + (size_t)n_cigar * 4 > (size_t)bs) return -1;
    if (want_align_) { r.cigar.resize(n_cigar); if (n_cigar) memcpy(r.cigar.data(), p + o, (size_t)n_cigar * 4); }
    o += (size_t)n_cigar * 4;
    const size_t seq_bytes = ((size_t)r.l_qseq + 1) / 2;
    if (r.l_qseq < 0 || o + seq_bytes + (size_t)r.l_qseq > (size_t)bs) return -1;
    static const char nt16[] = "=ACMGRSVTWYHKDBN";  // htslib seq_nt16_str
    const uint8_t* t6 = nt6_table();
    if (want_view_) {
      r.seq4_view = p + o;            // `search` copies only the reads it submits (1 in 9 of a smoothed BAM)
    } else if (want_align_) {
      r.seq4.assign(p + o, p + o + seq_bytes);
    } else {
      r.nt6.resize((size_t)r.l_qseq);
      for (int32_t i = 0; i < r.l_qseq; ++i) {
        const uint8_t b = p[o + (i >> 1)];
        r.nt6[i] = t6[(int)nt16[(i & 1) ? (b & 0xf) : (b >> 4)]];
      }
    }
    r.off_seq = o; r.off_qual = o + seq_bytes;
    o += seq_bytes + (size_t)r.l_qseq;
    r.off_aux = o;
    if (want_raw_) r.raw.assign(p, p + bs);
    r.has_xf = r.has_hp = false; r.xf = r.hp = 0; r.cg_cigar = false; r.cg_off = r.cg_len = 0;
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
          const size_t bytes = es * (size_t)(cnt < 0 ? 0 : cnt);
          if (o + 5 + bytes > (size_t)bs) return -1;
          // CG:B,I -- the real CIGAR of a record with more than 65535 ops, whose CIGAR field then holds the placeholder
          // <l_seq>S<ref_len>N (SAM spec 4.2.2; htslib's sam_read1 swaps it in transparently, bam_tag2cigar).  ADVICE r1.
          if (want_align_ && t0 == 'C' && t1 == 'G' && st == 'I' && cnt > 0 && r.cigar.size() == 2 && (r.cigar[0] & 0xf) == 4 &&
              (int32_t)(r.cigar[0] >> 4) == r.l_qseq && (r.cigar[1] & 0xf) == 3) {
            r.cigar.resize((size_t)cnt);
            memcpy(r.cigar.data(), p + o + 5, (size_t)cnt * 4);
            r.cg_cigar = true; r.cg_off = o - 3; r.cg_len = 3 + 5 + bytes;
          }
          o += 5 + bytes;
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
  BgzfSource src_;
  bool ok_ = false, want_align_ = false, want_raw_ = false, want_view_ = false;
  std::string text_;
  std::vector<std::string> ref_names_;
  std::vector<int32_t> ref_lens_;
  std::vector<uint8_t> buf_;
  int64_t header_bytes_ = 0;
};

}  // namespace svdss
