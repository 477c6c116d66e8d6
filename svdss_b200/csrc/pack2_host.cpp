// Host-side re-packing of BAM-native 4-bit reads to 2 bits per base (building block for the next e2e step:
// the search is PCIe-bound end to end -- 7.6 GB of 4-bit reads per million 15 kb reads at ~46 GB/s -- so the
// lever left is fewer bytes on the wire).  Host threads pack chunk k+1 while chunk k is in
// flight; the GPU side (unpack2.cuh, used by the streamed search when SVB_STREAM_PACK2=1) decodes 2 bits -> nt6.
//
// Layout: read r occupies (l + 3) / 4 packed bytes at out_offs[r] (every read starts on an output byte): every
// input byte (two bases, first in the high nibble, htslib nt16 codes) becomes one nibble (c2(first) << 2 |
// c2(second)), two input bytes one output byte, first byte in the high nibble; A C G T (nt16 1 2 4 8) ->
// 0 1 2 3.  A read holding any other code (N, IUPAC, '=') cannot be expressed: it is reported in `exception`
// and must travel in the 4-bit form.
#include <stdint.h>
#include <string.h>

#include <atomic>
#include <thread>
#include <vector>

#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "../../include/svdss_b200.h"

namespace {

// nt16 code -> 2-bit code, 0x80 = not expressible
const uint8_t C2[16] = {0x80, 0, 1, 0x80, 2, 0x80, 0x80, 0x80, 3, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80, 0x80};

// n input bytes (n even) -> n / 2 output bytes; returns true if every base was A, C, G or T
bool pack_scalar(const uint8_t* in, size_t n, uint8_t* out) {
  unsigned bad = 0;
  for (size_t k = 0; k + 1 < n; k += 2) {
    const uint8_t a = in[k], b = in[k + 1];
    const unsigned a1 = C2[a >> 4], a0 = C2[a & 15], b1 = C2[b >> 4], b0 = C2[b & 15];
    bad |= a1 | a0 | b1 | b0;
    out[k >> 1] = (uint8_t)(((a1 & 3) << 6) | ((a0 & 3) << 4) | ((b1 & 3) << 2) | (b0 & 3));
  }
  return !(bad & 0x80);
}

#if defined(__x86_64__)
__attribute__((target("avx2"))) bool pack_avx2(const uint8_t* in, size_t n, uint8_t* out) {
  const __m256i lut = _mm256_broadcastsi128_si256(_mm_loadu_si128(reinterpret_cast<const __m128i*>(C2)));
  const __m256i m0f = _mm256_set1_epi8(0x0f), m03 = _mm256_set1_epi8(0x03);
  const __m256i w = _mm256_set1_epi16(0x0110);   // maddubs: first byte of a pair * 16 + second byte * 1
  __m256i bad = _mm256_setzero_si256();
  size_t k = 0;
  for (; k + 64 <= n; k += 64) {
    const __m256i v0 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + k));
    const __m256i v1 = _mm256_loadu_si256(reinterpret_cast<const __m256i*>(in + k + 32));
    const __m256i h0 = _mm256_shuffle_epi8(lut, _mm256_and_si256(_mm256_srli_epi16(v0, 4), m0f)), l0 = _mm256_shuffle_epi8(lut, _mm256_and_si256(v0, m0f));
    const __m256i h1 = _mm256_shuffle_epi8(lut, _mm256_and_si256(_mm256_srli_epi16(v1, 4), m0f)), l1 = _mm256_shuffle_epi8(lut, _mm256_and_si256(v1, m0f));
    bad = _mm256_or_si256(bad, _mm256_or_si256(_mm256_or_si256(h0, l0), _mm256_or_si256(h1, l1)));
    // nibble of every input byte: c2(first) << 2 | c2(second)
    const __m256i n0 = _mm256_or_si256(_mm256_slli_epi16(_mm256_and_si256(h0, m03), 2), _mm256_and_si256(l0, m03));
    const __m256i n1 = _mm256_or_si256(_mm256_slli_epi16(_mm256_and_si256(h1, m03), 2), _mm256_and_si256(l1, m03));
    // adjacent nibbles -> one byte (16-bit lanes), then 16 -> 8 bits; packus works per 128-bit lane: fix the order
    const __m256i p0 = _mm256_maddubs_epi16(n0, w), p1 = _mm256_maddubs_epi16(n1, w);
    const __m256i pk = _mm256_permute4x64_epi64(_mm256_packus_epi16(p0, p1), 0xd8);
    _mm256_storeu_si256(reinterpret_cast<__m256i*>(out + (k >> 1)), pk);
  }
  const bool head_ok = _mm256_movemask_epi8(bad) == 0;   // bit 7 of any looked-up byte
  return pack_scalar(in + k, n - k, out + (k >> 1)) && head_ok;
}
#endif

bool pack_bytes(const uint8_t* in, size_t n, uint8_t* out) {
#if defined(__x86_64__)
  static const bool have_avx2 = __builtin_cpu_supports("avx2");
  if (have_avx2) return pack_avx2(in, n, out);
#endif
  return pack_scalar(in, n, out);
}

// one read: l bases in (l + 1) / 2 input bytes starting at an even offset; the pad nibble of an odd length and
// the pad byte that makes the byte count even are not bases
bool pack_read(const uint8_t* in, int64_t l, uint8_t* out) {
  const int64_t full = l / 2;               // bytes holding two bases
  const int64_t even = full & ~(int64_t)1;  // of those, the ones that pair up
  bool ok = pack_bytes(in, (size_t)even, out);
  const int64_t rest_bases = l - 2 * even;  // 0..3 bases in up to two more input bytes
  if (rest_bases > 0) {
    unsigned c[4] = {0, 0, 0, 0}, bad = 0;
    for (int64_t b = 0; b < rest_bases; ++b) {
      const uint8_t byte = in[even + (b >> 1)];
      c[b] = C2[(b & 1) ? (byte & 15) : (byte >> 4)];
      bad |= c[b];
    }
    out[even >> 1] = (uint8_t)(((c[0] & 3) << 6) | ((c[1] & 3) << 4) | ((c[2] & 3) << 2) | (c[3] & 3));
    ok = ok && !(bad & 0x80);
  }
  return ok;
}

}  // namespace

extern "C" SVB_API int svb_pack2_host(const uint8_t* seq4, const int64_t* seq4_offs, const int32_t* l_qseq, int64_t n_reads,
                                      uint8_t* out, const int64_t* out_offs, uint8_t* exception, int threads) {
  if (!seq4 || !seq4_offs || !l_qseq || !out || !out_offs || !exception || n_reads < 0) return SVB_EINVAL;
  unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  const int64_t CH = 256;   // reads per work item
  std::atomic<int64_t> next(0);
  auto worker = [&]() {
    for (;;) {
      const int64_t r0 = next.fetch_add(CH);
      if (r0 >= n_reads) return;
      const int64_t r1 = r0 + CH < n_reads ? r0 + CH : n_reads;
      for (int64_t r = r0; r < r1; ++r) exception[r] = pack_read(seq4 + seq4_offs[r], l_qseq[r], out + out_offs[r]) ? 0 : 1;
    }
  };
  if (nt == 1 || n_reads <= CH) { worker(); return SVB_OK; }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < nt; ++t) th.emplace_back(worker);
  for (auto& x : th) x.join();
  return SVB_OK;
}

// The streamed form: one chunk of the batch = the base positions [o, o + nb) of the concatenated reads (the unit
// the search kernel waits for).  Packs, with all host threads, the part of every read r_lo..r_hi that falls into
// the chunk -- widened to whole packed bytes, so neighbouring chunks may both carry a byte they share -- into
// `stage`, which then holds the bytes [*pa2, *pe2) of the batch's packed layout (read r at pk_offs[r]).  Bases
// that are not A, C, G or T have no 2-bit form: their positions (all of them decode to nt6 N) are appended to
// exc_pos, to be patched on the device after the chunk has been decoded; *n_exc receives how many (if it
// exceeds exc_cap the chunk should travel in the 4-bit form instead).
extern "C" SVB_API int svb_pack2_chunk(const uint8_t* seq4, const int64_t* seq4_offs, const int64_t* offs, const int64_t* pk_offs,
                                       int64_t r_lo, int64_t r_hi, int64_t o, int64_t nb, uint8_t* stage, int64_t stage_cap,
                                       int64_t* pa2, int64_t* pe2, int64_t* exc_pos, int64_t exc_cap, int64_t* n_exc, int threads) {
  if (!seq4 || !seq4_offs || !offs || !pk_offs || !stage || !pa2 || !pe2 || !n_exc || r_lo < 0 || r_hi < r_lo || nb <= 0) return SVB_EINVAL;
  auto span = [&](int64_t r, int64_t& q0, int64_t& q1, int64_t& l) {   // packed bytes [q0, q1) of read r hold its bases inside the chunk
    l = offs[r + 1] - offs[r];
    const int64_t j0 = o > offs[r] ? o - offs[r] : 0, j1 = (o + nb < offs[r + 1] ? o + nb : offs[r + 1]) - offs[r];
    if (j1 <= j0) { q0 = q1 = 0; return false; }
    q0 = j0 >> 2; q1 = (j1 + 3) >> 2;
    return true;
  };
  int64_t q0, q1, l;
  *pa2 = *pe2 = pk_offs[r_lo];
  bool any = false;
  for (int64_t r = r_lo; r <= r_hi; ++r)
    if (span(r, q0, q1, l)) { if (!any) *pa2 = pk_offs[r] + q0; *pe2 = pk_offs[r] + q1; any = true; }
  *n_exc = 0;
  if (!any) return SVB_OK;
  if (*pe2 - *pa2 > stage_cap) return SVB_ERANGE;
  const int64_t base = *pa2;
  unsigned nt = threads > 0 ? (unsigned)threads : std::thread::hardware_concurrency();
  if (nt == 0) nt = 1;
  const int64_t CH = 128, n = r_hi - r_lo + 1;
  std::atomic<int64_t> next(0), n_bad(0);
  std::vector<std::vector<int64_t>> bad_reads(nt);
  auto worker = [&](unsigned t) {
    for (;;) {
      const int64_t k0 = next.fetch_add(CH);
      if (k0 >= n) return;
      for (int64_t r = r_lo + k0; r < r_lo + (k0 + CH < n ? k0 + CH : n); ++r) {
        int64_t a, b, len;
        if (!span(r, a, b, len)) continue;
        const int64_t bases = (len < 4 * b ? len : 4 * b) - 4 * a;
        if (!pack_read(seq4 + seq4_offs[r] + 2 * a, bases, stage + (pk_offs[r] + a - base))) bad_reads[t].push_back(r);
      }
    }
  };
  if (nt == 1 || n <= CH) worker(0);
  else {
    std::vector<std::thread> th;
    for (unsigned t = 0; t < nt; ++t) th.emplace_back(worker, t);
    for (auto& x : th) x.join();
  }
  // the rare reads with other codes: their positions inside the chunk, base by base
  int64_t m = 0;
  for (auto& v : bad_reads)
    for (int64_t r : v) {
      const uint8_t* in = seq4 + seq4_offs[r];
      const int64_t j0 = o > offs[r] ? o - offs[r] : 0, j1 = (o + nb < offs[r + 1] ? o + nb : offs[r + 1]) - offs[r];
      for (int64_t j = j0; j < j1; ++j) {
        const unsigned c = (j & 1) ? (in[j >> 1] & 15u) : (unsigned)(in[j >> 1] >> 4);
        if (C2[c] & 0x80) { if (exc_pos && m < exc_cap) exc_pos[m] = offs[r] + j; ++m; }
      }
    }
  *n_exc = m;
  return SVB_OK;
}
