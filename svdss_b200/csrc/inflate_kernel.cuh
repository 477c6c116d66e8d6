// BGZF members inflated on the device (second half: the CLI stages are bound by host inflate --
// the reference spreads it over htslib's bgzf_mt threads, ping_pong.cpp:249, clusterer.cpp:13).  A BGZF member
// is an independent raw-deflate stream of at most 64 KiB of payload, so a window of a BAM file is thousands of
// independent jobs: one THREAD per member, canonical-Huffman decode by code-length counts (no look-up tables to
// build: 16 counts + the symbols in code order, 1.3 KB of local memory per thread), output written straight to
// its place in the window, back-references read from there (a member never refers across its own start).
// Round-1 state: a building block, checked on the CPU through tests/emul against zlib and on the GPU by
// svb_bgzf_inflate_device; round 2: BgzfSource uses it with `--gpu-inflate` (host/io.hpp).  Free of host
// code so that tests/emul compiles it for the CPU.
//
// RFC 1951 restated: a stream is a sequence of blocks, each with a 3-bit header (BFINAL, BTYPE).  BTYPE 0:
// skip to a byte boundary, LEN, ~LEN, LEN literal bytes.  BTYPE 1: the fixed code (literal/length lengths 8,9,7,8
// over 0..143, 144..255, 256..279, 280..287; 30 distance codes of 5 bits).  BTYPE 2: HLIT, HDIST, HCLEN, the code
// length code in the order 16 17 18 0 8 7 9 6 10 5 11 4 12 3 13 2 14 1 15, then the run-length coded lengths of
// both codes.  Symbols 257..285 are lengths (base + extra bits), followed by a distance symbol (base + extra bits).
// Bits are taken from the least significant end of each byte; Huffman codes are packed most significant bit first.
#pragma once
#include <stdint.h>

namespace svb {

enum : int { INF_OK = 0, INF_EINPUT = 1, INF_EOUTPUT = 2, INF_ECODE = 3, INF_EDIST = 4, INF_ESIZE = 5 };

struct InfBits {
  const uint8_t* p;
  int64_t n, pos;      // input bytes, next byte
  uint64_t buf;        // bits not yet consumed, least significant first
  int cnt;
  bool over;           // read past the end of the input
};

__device__ __forceinline__ void inf_refill(InfBits& b) {
  while (b.cnt <= 56) {
    uint64_t v = 0;
    if (b.pos < b.n) v = b.p[b.pos];
    else if (b.pos >= b.n + 8) { b.over = true; }   // up to 8 bytes of look-ahead are zero padding, not an error yet
    ++b.pos;
    b.buf |= v << b.cnt;
    b.cnt += 8;
  }
}
__device__ __forceinline__ unsigned inf_bits(InfBits& b, int k) {   // k <= 16
  if (b.cnt < k) inf_refill(b);
  const unsigned v = (unsigned)(b.buf & ((1ull << k) - 1ull));
  b.buf >>= k; b.cnt -= k;
  return v;
}
// bytes of the input actually consumed (whole bytes still buffered are given back)
__device__ __forceinline__ int64_t inf_consumed(const InfBits& b) { return b.pos - (b.cnt >> 3); }

struct InfHuff {
  uint16_t count[16];    // codes of each length
  uint16_t symbol[288];  // symbols ordered by code
};

// canonical code from code lengths; returns false for an over-subscribed set (incomplete sets are allowed where
// RFC 1951 allows them: a single distance code)
__device__ __forceinline__ bool inf_construct(InfHuff& h, const uint8_t* len, int n) {
  for (int l = 0; l < 16; ++l) h.count[l] = 0;
  for (int s = 0; s < n; ++s) ++h.count[len[s]];
  int left = 1;
  for (int l = 1; l < 16; ++l) {
    left <<= 1;
    left -= h.count[l];
    if (left < 0) return false;
  }
  uint16_t offs[16];
  offs[1] = 0;
  for (int l = 1; l < 15; ++l) offs[l + 1] = (uint16_t)(offs[l] + h.count[l]);
  for (int s = 0; s < n; ++s)
    if (len[s]) h.symbol[offs[len[s]]++] = (uint16_t)s;
  return true;
}

// one symbol: walk the code lengths, one bit at a time (codes are sent most significant bit first)
__device__ __forceinline__ int inf_decode(InfBits& b, const InfHuff& h) {
  if (b.cnt < 15) inf_refill(b);
  int code = 0, first = 0, index = 0;
  uint64_t bits = b.buf;
  for (int l = 1; l < 16; ++l) {
    code |= (int)(bits & 1u);
    bits >>= 1;
    const int c = h.count[l];
    if (code - c < first) {
      b.buf >>= l; b.cnt -= l;
      return h.symbol[index + (code - first)];
    }
    index += c;
    first += c;
    first <<= 1;
    code <<= 1;
  }
  return -1;
}

// inflate one raw-deflate stream in[0, in_len) into out[0, out_len); the stream must produce exactly out_len bytes
__device__ __forceinline__ int inflate_member(const uint8_t* __restrict__ in, int64_t in_len, uint8_t* __restrict__ out, int64_t out_len) {
  const uint16_t lbase[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  const uint8_t lext[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  const uint16_t dbase[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  const uint8_t dext[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  InfBits b;
  b.p = in; b.n = in_len; b.pos = 0; b.buf = 0; b.cnt = 0; b.over = false;
  InfHuff lc, dc;
  uint8_t len[320];
  int64_t o = 0;
  for (;;) {
    const unsigned last = inf_bits(b, 1), type = inf_bits(b, 2);
    if (type == 0) {   // stored
      b.buf >>= (b.cnt & 7); b.cnt -= (b.cnt & 7);
      const unsigned n = inf_bits(b, 16), nn = inf_bits(b, 16);
      if (b.over || (n ^ nn) != 0xffffu) return INF_EINPUT;
      // whole bytes only from here: hand the buffered ones back and copy
      int64_t at = inf_consumed(b);
      if (at + (int64_t)n > in_len) return INF_EINPUT;
      if (o + (int64_t)n > out_len) return INF_EOUTPUT;
      for (unsigned i = 0; i < n; ++i) out[o + i] = in[at + i];
      o += n; at += n;
      b.pos = at; b.buf = 0; b.cnt = 0;
    } else if (type == 1 || type == 2) {
      if (type == 1) {
        int s = 0;
        for (; s < 144; ++s) len[s] = 8;
        for (; s < 256; ++s) len[s] = 9;
        for (; s < 280; ++s) len[s] = 7;
        for (; s < 288; ++s) len[s] = 8;
        inf_construct(lc, len, 288);
        for (s = 0; s < 30; ++s) len[s] = 5;
        inf_construct(dc, len, 30);
      } else {
        const int nlen = (int)inf_bits(b, 5) + 257, ndist = (int)inf_bits(b, 5) + 1, ncode = (int)inf_bits(b, 4) + 4;
        if (nlen > 286 || ndist > 30) return INF_ECODE;
        int i = 0;
        for (; i < ncode; ++i) len[order[i]] = (uint8_t)inf_bits(b, 3);
        for (; i < 19; ++i) len[order[i]] = 0;
        if (!inf_construct(lc, len, 19)) return INF_ECODE;   // lc holds the code length code for a moment
        i = 0;
        while (i < nlen + ndist) {
          int sym = inf_decode(b, lc);
          if (sym < 0 || b.over) return INF_ECODE;
          if (sym < 16) len[i++] = (uint8_t)sym;
          else {
            int prev = 0, rep;
            if (sym == 16) {
              if (i == 0) return INF_ECODE;
              prev = len[i - 1];
              rep = 3 + (int)inf_bits(b, 2);
            } else if (sym == 17) rep = 3 + (int)inf_bits(b, 3);
            else rep = 11 + (int)inf_bits(b, 7);
            if (i + rep > nlen + ndist) return INF_ECODE;
            while (rep--) len[i++] = (uint8_t)prev;
          }
        }
        if (len[256] == 0) return INF_ECODE;   // no end-of-block code
        // the distance lengths follow the literal/length ones in len[]: build the distance code first
        if (!inf_construct(dc, len + nlen, ndist)) return INF_ECODE;
        if (!inf_construct(lc, len, nlen)) return INF_ECODE;
      }
      for (;;) {
        const int sym = inf_decode(b, lc);
        if (sym < 0 || b.over) return sym < 0 ? INF_ECODE : INF_EINPUT;
        if (sym < 256) {
          if (o >= out_len) return INF_EOUTPUT;
          out[o++] = (uint8_t)sym;
        } else if (sym == 256) break;
        else {
          const int ls = sym - 257;
          if (ls >= 29) return INF_ECODE;
          const int l = lbase[ls] + (int)inf_bits(b, lext[ls]);
          const int ds = inf_decode(b, dc);
          if (ds < 0 || ds >= 30) return INF_ECODE;
          const int64_t d = dbase[ds] + (int64_t)inf_bits(b, dext[ds]);
          if (d > o) return INF_EDIST;
          if (o + l > out_len) return INF_EOUTPUT;
          for (int k = 0; k < l; ++k) { out[o] = out[o - d]; ++o; }   // overlapping copies repeat, byte by byte
        }
      }
    } else return INF_ECODE;
    if (b.over) return INF_EINPUT;
    if (last) break;
  }
  if (inf_consumed(b) > in_len) return INF_EINPUT;   // the last symbols came out of the padding: truncated input
  return o == out_len ? INF_OK : INF_ESIZE;
}

// member m: deflate payload comp[in_offs[m], in_offs[m+1]) -> out[out_offs[m], out_offs[m+1]); status[m] = INF_*.
// One warp per CTA, `mpw` (1..32) of its lanes take a member each: the lanes of a warp walk different streams and
// diverge, so a window with fewer members than the machine has warp slots spreads over more warps instead
// (svb_bgzf_inflate_device picks mpw).
__global__ void __launch_bounds__(32) k_bgzf_inflate(const uint8_t* __restrict__ comp, const int64_t* __restrict__ in_offs,
                                                     const int64_t* __restrict__ out_offs, int64_t n_members,
                                                     uint8_t* __restrict__ out, int32_t* __restrict__ status, int mpw) {
  if ((int)threadIdx.x >= mpw) return;
  const int64_t m = (int64_t)blockIdx.x * mpw + threadIdx.x;
  if (m >= n_members) return;
  const int64_t a = in_offs[m], b = in_offs[m + 1], oa = out_offs[m], ob = out_offs[m + 1];
  status[m] = (ob == oa && b == a) ? INF_OK : inflate_member(comp + a, b - a, out + oa, ob - oa);
}

}  // namespace svb
