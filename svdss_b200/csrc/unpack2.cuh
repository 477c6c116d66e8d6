// Device side of the 2-bit read transport (host side: pack2_host.cpp, svb_pack2_host): reads packed 2 bits per
// base, first base in the two high bits of a byte, A C G T = 0 1 2 3, every read starting on a byte -- decoded
// to the one-byte-per-base nt6 form the search kernels read (A C G T = 1 2 3 4).  Same shape as the 4-bit
// unpack of sfs_search.cu (unpack16 / unpack4_read / k_unpack4): a warp per read, lanes own 16-byte aligned
// groups of the OUTPUT so that stores are coalesced 16-byte stores, the ragged head and tail of a range go base
// by base, and a range [A, B) of output positions can be requested so that the streamed pipeline can decode
// chunk by chunk (unpack_cta_loop2 of sfs_search.cu does, with SVB_STREAM_PACK2=1; off by default).  Free of host code so that
// tests/emul compiles it for the CPU.
#pragma once
#include <stdint.h>

namespace svb {

#ifdef __CUDA_ARCH__
#define SVB_LDCG(p) __ldcg(p)
#else
#define SVB_LDCG(p) (*(p))
#endif

// bytes [sh/8, sh/8 + 8) of the 16-byte little-endian value hi:lo
__device__ __forceinline__ uint64_t u2_funnel64(uint64_t lo, uint64_t hi, int sh) { return sh ? (lo >> sh) | (hi << (64 - sh)) : lo; }

// 16 bases starting at quarter `q` (0..3) of packed byte `byte0`, as 16 nt6 bytes.  The five packed bytes that can
// hold them come from two aligned 8-byte loads (the buffer is padded by 16 bytes).
__device__ __forceinline__ uint4 unpack2_16(const uint8_t* __restrict__ pk, int64_t byte0, int q) {
  const uint64_t* p = reinterpret_cast<const uint64_t*>(pk);
  const int64_t wi = byte0 >> 3;
  const uint64_t w0 = SVB_LDCG(p + wi), w1 = SVB_LDCG(p + wi + 1);
  const uint64_t x = u2_funnel64(w0, w1, (int)(byte0 & 7) * 8);   // packed bytes byte0 .. byte0 + 7, byte0 in the low bits
  uint32_t o[4];
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    uint32_t v = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const int u = q + 4 * g + t;                                  // quarter index from byte0: byte u >> 2, base u & 3 of it
      const unsigned c = (unsigned)(x >> (8 * (u >> 2) + 2 * (3 - (u & 3)))) & 3u;
      v |= (c + 1u) << (8 * t);
    }
    o[g] = v;
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}

// one warp decodes the bases of read r that fall into output positions [A, B)
__device__ __forceinline__ void unpack2_read(const uint8_t* __restrict__ pk, const int64_t* __restrict__ pk_offs,
                                             const int64_t* __restrict__ offs, int64_t r, int64_t A, int64_t B,
                                             uint8_t* __restrict__ out, int lane) {
  const int64_t o = offs[r], l = offs[r + 1] - o, pb = pk_offs[r];
  const int64_t lo = o > A ? o : A, hi = o + l < B ? o + l : B;
  if (lo >= hi) return;
  auto one = [&](int64_t g) {   // output position g
    const int64_t j = g - o;
    const uint8_t b = SVB_LDCG(pk + pb + (j >> 2));
    out[g] = (uint8_t)(((b >> (2 * (3 - (int)(j & 3)))) & 3u) + 1u);
  };
  const int64_t al = (lo + 15) & ~(int64_t)15;
  const int64_t body0 = hi < al ? hi : al, body1 = body0 + ((hi - body0) & ~(int64_t)15);
  if (lo + lane < body0) one(lo + lane);                       // head: fewer than 16 positions
  for (int64_t g = body0 + 16 * lane; g < body1; g += 512) {
    const int64_t j = g - o;
    *reinterpret_cast<uint4*>(out + g) = unpack2_16(pk, pb + (j >> 2), (int)(j & 3));
  }
  if (body1 + lane < hi) one(body1 + lane);                    // tail: fewer than 16 positions
}

// whole batch at once (no streaming)
__global__ void __launch_bounds__(128) k_unpack2(const uint8_t* __restrict__ pk, const int64_t* __restrict__ pk_offs,
                                                  const int64_t* __restrict__ offs, int64_t n_reads, int64_t total,
                                                  uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += nwarps)
    unpack2_read(pk, pk_offs, offs, r, 0, total, out, lane);
}

}  // namespace svb
