// Shared declarations for libsvdss_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <chrono>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/svdss_b200.h"

namespace svb {

// thread-local last error text, readable through svb_last_error()
void set_error(const char* fmt, ...);

#define SVB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      svb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SVB_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)

#define SVB_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != SVB_OK) return _r; \
  } while (0)

// Stream-ordered allocations from the device's default memory pool with an unbounded release
// threshold: repeated batches reuse the same physical memory instead of paying cudaMalloc/cudaFree
// (tens of ms per call for multi-GB buffers on some hosts) inside every svb_sfs_* call.
inline cudaError_t pmalloc(void** p, size_t bytes, cudaStream_t st) {
  static thread_local int tuned_dev = -1;
  int dev = 0;
  cudaGetDevice(&dev);
  if (tuned_dev != dev) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      uint64_t thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
    tuned_dev = dev;
  }
  return cudaMallocAsync(p, bytes ? bytes : 1, st);
}
inline void pfree(void* p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}
// bytes a pmalloc can still get: what the driver reports free plus what the pool holds without using it.
// cudaMemGetInfo takes driver-wide locks: measured 0.2 ms as a rule but 8-36 ms when something else talks to the driver
// at that moment (an nvidia-smi query, another rank's allocation) -- twice per svb_call_batch.  So the driver's figure is
// kept for a few seconds and moved along with the pool's own reservation, which is the only thing this process changes;
// what other processes take in between is seen at the next refresh (callers budget 80 % of the answer).
inline cudaError_t pool_available(size_t* avail) {
  struct Seen { size_t free_b = 0; uint64_t reserved = 0; std::chrono::steady_clock::time_point at; bool valid = false; };
  static Seen seen[64];
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  cudaMemPool_t pool;
  uint64_t reserved = 0, used = 0;
  const bool have_pool = cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess &&
                         cudaMemPoolGetAttribute(pool, cudaMemPoolAttrReservedMemCurrent, &reserved) == cudaSuccess &&
                         cudaMemPoolGetAttribute(pool, cudaMemPoolAttrUsedMemCurrent, &used) == cudaSuccess;
  if (!have_pool) reserved = used = 0;
  std::lock_guard<std::mutex> lock(mu);
  Seen& c = seen[dev & 63];
  const auto now = std::chrono::steady_clock::now();
  if (!c.valid || now - c.at > std::chrono::seconds(5) || getenv("SVB_POOL_FRESH")) {
    size_t free_b = 0, total_b = 0;
    cudaError_t e = cudaMemGetInfo(&free_b, &total_b);
    if (e != cudaSuccess) return e;
    c.free_b = free_b; c.reserved = reserved; c.at = now; c.valid = true;
  }
  // free now = free then - what the pool has reserved since (+ what it has given back)
  int64_t free_now = (int64_t)c.free_b - ((int64_t)reserved - (int64_t)c.reserved);
  if (free_now < 0) free_now = 0;
  *avail = (size_t)free_now + (reserved > used ? (size_t)(reserved - used) : 0);
  return cudaSuccess;
}

// SVB_STAGE_STATS=1: host-side lap times of a batch call on stderr (where does a stage spend what its kernel does not?).
// A lap synchronises the device, so the lines are a diagnosis, not a measurement of the unperturbed call.
struct StageLog {
  bool on;
  const char* who;
  std::chrono::steady_clock::time_point t0;
  explicit StageLog(const char* w) : on(getenv("SVB_STAGE_STATS") != nullptr), who(w), t0(std::chrono::steady_clock::now()) {}
  void lap(const char* what) {
    if (!on) return;
    cudaDeviceSynchronize();
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[svb-stage] %s: %s %.3f ms\n", who, what, std::chrono::duration<double, std::milli>(t - t0).count());
    t0 = t;
  }
};

// ---------------------------------------------------------------------------------------------
// FM-index block array ("sampled-Occ BWT blocks") in HBM.
//
// A block is G consecutive 16-byte *slices*; slice j of a block is one uint4
//     .x = Occ count slot j at block start, u32, relative to the block's 2^32-symbol superblock
//     .y/.z/.w = bit-planes 0/1/2 of the nt6 codes of BWT symbols [32j, 32j+32) of the block
// so a G-lane group fetches a whole block with ONE coalesced 16-byte load per lane and every lane
// owns the 32 symbols whose planes it loaded.
//   G = 4 : 64-byte blocks, 128 symbols, slots {A,C,G,T};       N from the side array cntN[]
//   G = 8 : 128-byte blocks, 256 symbols, slots {A,C,G,T,N, cum64, cum128, cum192} where cumX packs,
//           one byte per symbol A,C,G,T, the in-block count before symbol X (sub-block sampling)
// Padding symbols past n carry code 7 (matches nothing).  Superblock table:
//     sbase[sb*8 + c] = acc[c] + Occ(c, sb << 32)           (int64, c = 0..5)
// ---------------------------------------------------------------------------------------------
struct IndexDev {
  int device = 0;
  int G = 8;                 // lanes per block: 4 or 8
  int64_t n = 0;             // BWT length (= text length, both strands + sentinels)
  int64_t acc[7] = {0};
  int64_t n_blocks = 0;
  uint4* d_blocks = nullptr; // n_blocks * G slices
  uint32_t* d_cntN = nullptr;// G == 4 only: Occ(N) at block start, relative to superblock
  int64_t* d_sbase = nullptr;
  int n_sb = 0;
  int64_t n_contigs = 0;
  // Located-match acceleration (built by svb_index_build / restored by svb_index_load; absent for
  // svb_index_from_bwt): once a search interval has size 1 and sits on a sampled SA row, the match
  // is a single text position and every further extension is a byte compare against the text.
  //   d_text   T itself, one byte per symbol, sentinels remapped to TEXT_SENTINEL (matches no read
  //            code), TEXT_PAD pad bytes of the same value on both sides
  //   d_ssa    SA[k] for every k with k % (1 << ss_log) == 0
  //   d_tstart text start of contig pair r (S_r $ rc(S_r) $), n_contigs + 1 entries: the mirror of
  //            text position x inside pair r is tstart[r] + tstart[r+1] - 2 - x
  uint8_t* d_text_alloc = nullptr;
  uint8_t* d_text = nullptr;
  uint64_t* d_ssa = nullptr;
  int64_t n_ssa = 0;
  int ss_log = 4;
  int64_t* d_tstart = nullptr;
  // K-mer jump table (128-byte blocks only; rebuilt from the block array at build / load time):
  // kmt[code] = interval of the K-mer whose i-th base (text order) sits in bits 2i..2i+1 (A=0..T=3),
  // packed as start | min(size, KMT_SAT) << 40.  A restart of the ping-pong walk replaces its first
  // K-1 extensions (all of which succeed when the K-mer occurs) by one 8-byte lookup.
  uint64_t* d_kmt = nullptr;
  int kmer_k = 0;
  // longest run of N in the text (-1: unknown, no text): a restart of the walk inside a longer run of N in a READ has a
  // closed form (sfs_search.cu, k_sfs_search_mop)
  int64_t max_nrun = -1;
};
constexpr uint64_t KMT_SAT = (1ull << 24) - 1;

// sfs_search.cu: fills d_kmt / kmer_k from the block array (no-op for 64-byte blocks)
int build_kmer_table(IndexDev* idx);

constexpr uint8_t TEXT_SENTINEL = 0xF0;
// pad bytes (TEXT_SENTINEL) either side of d_text: the warp-cooperative compare of the located-match mode (coop_text_step)
// lets all 32 lanes load 48-byte text windows up to 32 * 31 + 48 bytes past the current match whenever the READ still has
// bases there -- also when the match sits at the very end of the text (64 bytes were not enough)
constexpr int TEXT_PAD = 1152;

}  // namespace svb

struct svb_index {
  svb::IndexDev dev;
};
