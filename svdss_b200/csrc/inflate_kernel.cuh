// BGZF members inflated on the device (second half: the CLI stages are bound by host inflate --
// the reference spreads it over htslib's bgzf_mt threads, ping_pong.cpp:249, clusterer.cpp:13).  A BGZF member
// is an independent raw-deflate stream of at most 64 KiB of payload, so a window of a BAM file is thousands of
// independent jobs.  Two kernels: k_bgzf_inflate, one THREAD per member (round 1: canonical-Huffman decode by
// code-length counts -- no look-up tables to build: 16 counts + the symbols in code order, 1.3 KB of local memory per
// thread), and k_bgzf_inflate_warp further down, one WARP per member (round 2, the default: look-up tables in shared
// memory, queued copies).  Output is written straight to its place in the window, back-references are read from there
// (a member never refers across its own start).  Checked on the CPU through tests/emul against zlib and on the GPU by
// svb_bgzf_inflate_device; BgzfSource (host/io.hpp) and the device BAM loader (bam_stream.cu) use it with
// `--gpu-inflate`.  Free of host code so that tests/emul compiles it for the CPU.
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

// ---------------------------------------------------------------------------------------------------------------
// Second kernel, one WARP per member (round 2: the thread-per-member kernel above takes ~200 ms for a member whatever
// the window holds -- 32 diverging lanes, a bit-by-bit code walk through local memory, one L2 round trip per copied
// byte).  Lane 0 walks the stream; the warp shares
//   * look-up tables in shared memory: the next INF_LBITS (INF_DBITS) bits of the stream index an entry
//     (symbol << 4 | code length), one shared load per symbol; longer codes (rare) fall back to the canonical walk;
//     all 32 lanes fill the tables of a block from the canonical code lane 0 derived (code of the i-th symbol in
//     code order = first[len] + i - offs[len], bit-reversed because codes are packed most significant bit first);
//   * the copies: lane 0 writes literals itself but only QUEUES matches (decoding never needs the copied bytes) and
//     hands 32 of them to the warp at a time; the short ones whose source lies before the whole batch are copied one
//     per lane, all at once (one round trip to L2 for 32 matches -- the first version did one per match and spent
//     its time waiting: 9.7 ms per member); the rest go one after the other, spread over the lanes: l loads then l
//     stores, source byte k of an overlapping match (d < l) being k mod d.
// The other lanes wait in the shuffle that broadcasts lane 0's next event (queue full / end of block / error).
// ---------------------------------------------------------------------------------------------------------------
constexpr int INF_LBITS = 10, INF_DBITS = 8;

struct InfWarpMem {
  uint16_t ltab[1 << INF_LBITS];
  uint16_t dtab[1 << INF_DBITS];
  InfHuff lc, dc;
  uint16_t first[2][16], offs[2][16];   // first code / first index in code order of every length, literal-length and distance
  uint16_t qo[32], ql[32], qd[32];      // queued matches: place in the member's payload, length, distance
  uint8_t len[320];
};
constexpr int INF_QUEUE = 32;

#ifdef __CUDA_ARCH__
__device__ __forceinline__ unsigned inf_rev(unsigned code, int l) { return __brev(code) >> (32 - l); }
#else
inline unsigned inf_rev(unsigned code, int l) { unsigned r = 0; for (int i = 0; i < l; ++i) r |= ((code >> i) & 1u) << (l - 1 - i); return r; }
#endif

// four more bytes into the bit buffer (cnt <= 32 on entry): independent loads, one latency
__device__ __forceinline__ void inf_refill4(InfBits& b) {
  uint32_t w = 0;
  if (b.pos + 4 <= b.n) w = (uint32_t)b.p[b.pos] | ((uint32_t)b.p[b.pos + 1] << 8) | ((uint32_t)b.p[b.pos + 2] << 16) | ((uint32_t)b.p[b.pos + 3] << 24);
  else {
    for (int i = 0; i < 4; ++i) if (b.pos + i < b.n) w |= (uint32_t)b.p[b.pos + i] << (8 * i);
    if (b.pos >= b.n + 8) b.over = true;
  }
  b.buf |= (uint64_t)w << b.cnt;
  b.cnt += 32; b.pos += 4;
}

// lane 0: first[] / offs[] of a canonical code (counts already in h)
__device__ __forceinline__ void inf_code_starts(const InfHuff& h, uint16_t* first, uint16_t* offs) {
  unsigned f = 0, o = 0;
  first[0] = 0; offs[0] = 0;
  for (int l = 1; l < 16; ++l) {
    first[l] = (uint16_t)f; offs[l] = (uint16_t)o;
    f = (f + h.count[l]) << 1; o += h.count[l];
  }
}

// all lanes: table of the next `bits` bits -> (symbol << 4 | length); 0 where the code is longer or unassigned
__device__ __forceinline__ void inf_fill_table(uint16_t* tab, int bits, const InfHuff& h, const uint8_t* len, const uint16_t* first,
                                               const uint16_t* offs, int lane) {
  for (int i = lane; i < (1 << bits); i += 32) tab[i] = 0;
  __syncwarp();
  int coded = 0;
  for (int l = 1; l < 16; ++l) coded += h.count[l];
  for (int i = lane; i < coded; i += 32) {
    const int s = h.symbol[i], l = len[s];
    if (l > bits) continue;
    const unsigned code = (unsigned)first[l] + (unsigned)(i - offs[l]);
    for (unsigned r = inf_rev(code, l); r < (1u << bits); r += 1u << l) tab[r] = (uint16_t)((s << 4) | l);
  }
  __syncwarp();
}

enum : int { INF_EV_END = 0, INF_EV_MATCH = 1, INF_EV_ERR = 2 };

// the whole warp; returns INF_* (the same on every lane)
__device__ __forceinline__ int inflate_member_warp(const uint8_t* __restrict__ in, int64_t in_len, uint8_t* __restrict__ out, int out_len,
                                                   InfWarpMem& M, int lane) {
  const unsigned FULL = 0xffffffffu;
  const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  InfBits b;
  b.p = in; b.n = in_len; b.pos = 0; b.buf = 0; b.cnt = 0; b.over = false;
  int o = 0;
  for (;;) {
    // ---- block header (lane 0): st = error, type, last; stored blocks: n bytes from `at`; coded blocks: the canonical codes
    int st = INF_OK, type = 0, last = 0, nlen = 288, n = 0, at = 0;
    if (lane == 0) {
      last = (int)inf_bits(b, 1); type = (int)inf_bits(b, 2);
      if (type == 0) {
        b.buf >>= (b.cnt & 7); b.cnt -= (b.cnt & 7);
        const unsigned a = inf_bits(b, 16), na = inf_bits(b, 16);
        n = (int)a; at = (int)inf_consumed(b);
        if (b.over || (a ^ na) != 0xffffu || (int64_t)at + n > in_len) st = INF_EINPUT;
        else if (o + n > out_len) st = INF_EOUTPUT;
      } else if (type == 1) {
        int s = 0;
        for (; s < 144; ++s) M.len[s] = 8;
        for (; s < 256; ++s) M.len[s] = 9;
        for (; s < 280; ++s) M.len[s] = 7;
        for (; s < 288; ++s) M.len[s] = 8;
        for (; s < 318; ++s) M.len[s] = 5;
        inf_construct(M.lc, M.len, 288);
        inf_construct(M.dc, M.len + 288, 30);
      } else if (type == 2) {
        nlen = (int)inf_bits(b, 5) + 257;
        const int ndist = (int)inf_bits(b, 5) + 1, ncode = (int)inf_bits(b, 4) + 4;
        if (nlen > 286 || ndist > 30) st = INF_ECODE;
        else {
          int i = 0;
          for (; i < ncode; ++i) M.len[order[i]] = (uint8_t)inf_bits(b, 3);
          for (; i < 19; ++i) M.len[order[i]] = 0;
          if (!inf_construct(M.lc, M.len, 19)) st = INF_ECODE;   // lc holds the code length code for a moment
          i = 0;
          while (st == INF_OK && i < nlen + ndist) {
            const int sym = inf_decode(b, M.lc);
            if (sym < 0 || b.over) { st = INF_ECODE; break; }
            if (sym < 16) M.len[i++] = (uint8_t)sym;
            else {
              int prev = 0, rep;
              if (sym == 16) {
                if (i == 0) { st = INF_ECODE; break; }
                prev = M.len[i - 1];
                rep = 3 + (int)inf_bits(b, 2);
              } else if (sym == 17) rep = 3 + (int)inf_bits(b, 3);
              else rep = 11 + (int)inf_bits(b, 7);
              if (i + rep > nlen + ndist) { st = INF_ECODE; break; }
              while (rep--) M.len[i++] = (uint8_t)prev;
            }
          }
          if (st == INF_OK && M.len[256] == 0) st = INF_ECODE;   // no end-of-block code
          if (st == INF_OK && !inf_construct(M.dc, M.len + nlen, ndist)) st = INF_ECODE;
          if (st == INF_OK && !inf_construct(M.lc, M.len, nlen)) st = INF_ECODE;
        }
      } else st = INF_ECODE;
      if (st == INF_OK && type != 0) {
        inf_code_starts(M.lc, M.first[0], M.offs[0]);
        inf_code_starts(M.dc, M.first[1], M.offs[1]);
      }
    }
    st = __shfl_sync(FULL, st, 0);
    if (st != INF_OK) return st;
    type = __shfl_sync(FULL, type, 0); last = __shfl_sync(FULL, last, 0);
    __syncwarp();
    if (type == 0) {
      n = __shfl_sync(FULL, n, 0); at = __shfl_sync(FULL, at, 0);
      for (int i = lane; i < n; i += 32) out[o + i] = in[at + i];
      o += n;
      if (lane == 0) { b.pos = (int64_t)at + n; b.buf = 0; b.cnt = 0; }
      __syncwarp();
    } else {
      nlen = __shfl_sync(FULL, nlen, 0);
      inf_fill_table(M.ltab, INF_LBITS, M.lc, M.len, M.first[0], M.offs[0], lane);
      inf_fill_table(M.dtab, INF_DBITS, M.dc, M.len + nlen, M.first[1], M.offs[1], lane);
      for (;;) {
        // lane 0 runs ahead: literals go straight to their place, matches are queued (decoding never needs the bytes a
        // match copies); the warp then does the queued copies together -- one round trip to memory per INF_QUEUE matches
        int ev = INF_EV_END, nq = 0;
        if (lane == 0) {
          for (;;) {
            if (b.cnt < 32) inf_refill4(b);
            unsigned e = M.ltab[(unsigned)b.buf & ((1u << INF_LBITS) - 1u)];
            int sym;
            if (e & 15u) { sym = (int)(e >> 4); b.buf >>= (e & 15u); b.cnt -= (int)(e & 15u); }
            else sym = inf_decode(b, M.lc);
            if (sym < 0 || b.over) { ev = INF_EV_ERR; st = sym < 0 ? INF_ECODE : INF_EINPUT; break; }
            if (sym < 256) {
              if (o >= out_len) { ev = INF_EV_ERR; st = INF_EOUTPUT; break; }
              out[o++] = (uint8_t)sym;
              continue;
            }
            if (sym == 256) break;
            const int ls = sym - 257;
            if (ls >= 29) { ev = INF_EV_ERR; st = INF_ECODE; break; }
            // RFC 1951 3.2.5 by formula: lengths 3..10 one by one, then four codes per extra bit, 258 on its own
            const int le = ls < 8 ? 0 : (ls == 28 ? 0 : (ls - 4) >> 2);
            const int l = (ls < 8 ? 3 + ls : (ls == 28 ? 258 : 3 + ((4 + (ls & 3)) << le))) + (int)inf_bits(b, le);
            if (b.cnt < 32) inf_refill4(b);
            e = M.dtab[(unsigned)b.buf & ((1u << INF_DBITS) - 1u)];
            int ds;
            if (e & 15u) { ds = (int)(e >> 4); b.buf >>= (e & 15u); b.cnt -= (int)(e & 15u); }
            else ds = inf_decode(b, M.dc);
            if (ds < 0 || ds >= 30) { ev = INF_EV_ERR; st = INF_ECODE; break; }
            const int de = ds < 4 ? 0 : (ds - 2) >> 1;
            const int d = (ds < 4 ? 1 + ds : 1 + ((2 + (ds & 1)) << de)) + (int)inf_bits(b, de);
            if (d > o) { ev = INF_EV_ERR; st = INF_EDIST; break; }
            if (o + l > out_len) { ev = INF_EV_ERR; st = INF_EOUTPUT; break; }
            M.qo[nq] = (uint16_t)o; M.ql[nq] = (uint16_t)l; M.qd[nq] = (uint16_t)d;   // o <= 65533, d <= 32768
            ++nq;
            o += l;
            if (nq == INF_QUEUE) { ev = INF_EV_MATCH; break; }
          }
        }
        ev = __shfl_sync(FULL, ev, 0);
        if (ev == INF_EV_ERR) return __shfl_sync(FULL, st, 0);
        o = __shfl_sync(FULL, o, 0);
        nq = __shfl_sync(FULL, nq, 0);
        __syncwarp();                                   // lane 0's literals and the queue are visible to the lanes that copy
        if (nq) {
          // a match whose source lies before the first queued match reads nothing this round writes: those, when
          // short, are done by one lane each, all at once; the others one after the other by the whole warp
          int mo = 0, ml = 0, md = 1;
          bool later = false;
          if (lane < nq) {
            mo = M.qo[lane]; ml = M.ql[lane]; md = M.qd[lane];
            later = ml > 16 || mo - md + (ml < md ? ml : md) > (int)M.qo[0];
            if (!later) {
              uint8_t* dst = out + mo;
              const uint8_t* src = dst - md;
              uint8_t v[16];
#pragma unroll
              for (int k = 0; k < 16; ++k)
                if (k < ml) v[k] = src[md >= ml ? k : k % md];
#pragma unroll
              for (int k = 0; k < 16; ++k)
                if (k < ml) dst[k] = v[k];
            }
          }
          unsigned todo = __ballot_sync(FULL, later);
          __syncwarp();
          while (todo) {
            const int q = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const int l = M.ql[q], d = M.qd[q];
            uint8_t* dst = out + M.qo[q];
            const uint8_t* src = dst - d;
            uint8_t v[9];
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              const int k = lane + 32 * j;
              if (k < l) v[j] = src[d >= l ? k : k % d];
            }
#pragma unroll
            for (int j = 0; j < 9; ++j) {
              const int k = lane + 32 * j;
              if (k < l) dst[k] = v[j];
            }
            __syncwarp();
          }
        }
        if (ev == INF_EV_END) break;
      }
    }
    st = __shfl_sync(FULL, (int)(lane == 0 && b.over), 0);
    if (st) return INF_EINPUT;
    if (last) break;
  }
  int rc = INF_OK;
  if (lane == 0) rc = inf_consumed(b) > in_len ? INF_EINPUT : (o == out_len ? INF_OK : INF_ESIZE);
  return __shfl_sync(FULL, rc, 0);
}

extern __shared__ unsigned long long inf_smem[];

// blockDim.x / 32 warps per CTA, warp w of CTA c takes member c * warps + w
template <int MIN_CTAS>   // 16: 64 registers, 32 warps per SM; 12: 80 registers, no spills
__global__ void __launch_bounds__(64, MIN_CTAS) k_bgzf_inflate_warp(const uint8_t* __restrict__ comp, const int64_t* __restrict__ in_offs,
                                                          const int64_t* __restrict__ out_offs, int64_t n_members,
                                                          uint8_t* __restrict__ out, int32_t* __restrict__ status) {
  const int warp = (int)threadIdx.x >> 5, lane = (int)threadIdx.x & 31, warps = (int)blockDim.x >> 5;
  const int64_t m = (int64_t)blockIdx.x * warps + warp;
  if (m >= n_members) return;
  InfWarpMem& M = reinterpret_cast<InfWarpMem*>(inf_smem)[warp];
  const int64_t a = in_offs[m], b = in_offs[m + 1], oa = out_offs[m], ob = out_offs[m + 1];
  const int rc = (ob == oa && b == a) ? INF_OK : inflate_member_warp(comp + a, b - a, out + oa, (int)(ob - oa), M, lane);
  if (lane == 0) status[m] = rc;
}

}  // namespace svb
