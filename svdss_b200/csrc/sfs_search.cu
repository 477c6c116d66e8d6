// SFS extraction on the GPU: the ping-pong FMD search of PingPong::ping_pong_search
// (reference ping_pong.cpp:4-49) for a whole batch of reads, with Assembler::assemble
// (assembler.cpp:34-56) fused as a streaming epilogue.
//
// Only Occ(c, k) and Occ(c, k + size) of ONE symbol are needed per step: every direction switch in
// ping_pong.cpp restarts from rb3_fmd_set_intv (:12, :30), so the bidirectional bookkeeping of
// rb3_fmd_extend is never observed; the forward phase is a backward search with complemented
// characters on the same (strand-closed) BWT.
//
// Three generations of the kernel live here (DESIGN.md 3.1 has the measurements behind each step):
//   k_sfs_search<G,SPL>   one G-lane group per read, each lane loads 16 bytes of the index block; the
//                         only kernel for 64-byte blocks; persistent grid, reads pulled from a counter
//   k_sfs_search_tma      thread per read, 128-byte blocks staged into shared memory by the warp
//                         (cp.async or TMA bulk copies): the pure rank walk at 0.69-0.72 of HBM peak
//   k_sfs_search_mop      the default: thread per read, micro-op pipeline over three ways of answering
//                         an extension (index blocks, located-match text compare done by the whole warp,
//                         K-mer jump table), main launch + tail launch (parked walks, one per warp, with
//                         32-link sprints through novel sequence); also unpacks BAM-native 4-bit reads
//                         streamed from the host in dedicated CTAs of the same launch
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "unpack2.cuh"   // 2-bit read transport, device side (emulator-tested; used by unpack_cta_loop2 when SVB_STREAM_PACK2=1)

struct svb_reads {
  int device = 0;
  int64_t n_reads = 0;
  int64_t total = 0;           // bytes of sequence
  uint8_t* d_seq = nullptr;    // padded to 64 B beyond total
  int64_t* d_offs = nullptr;   // n_reads + 1
  uint32_t* d_order = nullptr; // read indices, longest first
  ulonglong2* d_sched = nullptr; // hand-out order: {offset, len << 32 | read index} (one 16-byte load per read)
  bool owns_seq = true;
  int64_t max_len = 0;
};

namespace svb {

struct SearchParams {
  const uint4* __restrict__ blocks;
  const uint32_t* __restrict__ cntN;
  const int64_t* __restrict__ sbase;
  int64_t acc[7];
  const uint8_t* __restrict__ seq;
  const int64_t* __restrict__ offs;
  const uint32_t* __restrict__ order;
  const ulonglong2* __restrict__ sched;  // hand-out slot w -> {offs[r], len << 32 | r}; nullptr: use order / offs
  int64_t n_reads;
  int overlap;
  int assemble;
  unsigned long long* work;      // next read to hand out
  unsigned long long* out_count; // records appended
  unsigned long long* stats;     // [0] extensions, [1] blocks touched, [2] extensions answered from the text, [3] stream stalled,
                                 // [4] first thread in / [5] queue ran dry / [6] last warp out (globaltimer ns)
  // located-match mode (IndexDev::d_text / d_ssa / d_tstart); text == nullptr: rank mode only
  const uint8_t* __restrict__ text;
  const uint64_t* __restrict__ ssa;
  const int64_t* __restrict__ tstart;
  int64_t n_contigs;
  int ss_log;
  // K-mer jump table (IndexDev::d_kmt); kmt == nullptr: every restart walks from one base
  const uint64_t* __restrict__ kmt;
  int max_nrun;                  // longest run of N in the text, -1 = unknown (closed form of N-run restarts off)
  int kmer_k;
  uint64_t* out_key;             // read << 32 | sort key
  uint32_t* out_len;
  unsigned long long out_cap;
  // streamed batches (svb_sfs_batch): the read bytes arrive chunk by chunk while the kernel runs;
  // ready[c] != 0 once chunk c (chunk_bytes each) is resident.  nullptr = everything resident.
  const unsigned int* ready;
  int64_t chunk_bytes;
  // packed streamed batches (svb_sfs_batch_bam4): the first n_unpack CTAs of the grid do not search;
  // they turn each chunk of 4-bit reads into bytes as soon as the copy engine has delivered it
  // (arrived[c]) and raise ready[c] themselves.  Part of the same launch, so they are resident
  // next to the searching CTAs by construction.
  const uint8_t* seq4;
  const int64_t* seq4_offs;
  const unsigned int* arrived;
  unsigned int* ready_w;
  unsigned int* chunk_done;
  const int64_t* chunk_r0;   // first / last read reaching into chunk c
  const int64_t* chunk_r1;
  uint8_t* seq_w;            // = seq, writable
  int64_t total;
  int n_chunks, n_unpack;
  // 2-bit transport (SVB_STREAM_PACK2, k_sfs_search_mop<.., .., true>): the host re-packs every chunk to 2 bits per
  // base (svb_pack2_chunk) before it crosses PCIe; chunk_mode[c], written before arrived[c], says what arrived:
  // 2 = 2-bit bytes in pk2 (read r at pk2_offs[r]) plus exc_n[c] positions in exc_pos[c * exc_cap ...] to patch to
  // N once the chunk is decoded, 1 = the 4-bit bytes in seq4 as above (a chunk with too many such positions)
  const uint8_t* pk2;
  const int64_t* pk2_offs;
  const unsigned int* chunk_mode;
  const int64_t* exc_pos;
  const unsigned int* exc_n;
  int exc_cap;
  // two-phase launch of k_sfs_search_mop: when the work queue runs dry the main kernel parks every
  // unfinished walk in cont[] and ends; the tail kernel finishes them one per WARP (see Cont)
  struct Cont* cont;
  unsigned long long* n_cont;   // continuations written (main) / to hand out (tail)
  unsigned int* dry;            // raised by the first thread that finds the queue empty
  int stats_on;                    // SVB_SEARCH_STATS: extra counters
  int sprint_budget;               // rank steps a helper lane spends on one link before giving it up
  const unsigned int* last_ready;  // streamed batch: flag of the last chunk (walks are parked only once it is up:
                                   // before that the queue is "empty" merely because every thread holds a
                                   // reservation for a read that has not arrived yet)
};

__device__ __forceinline__ uint4 ldg_slice(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// Occ-based backward extension of [k, k+s) by symbol c (1..5) for one G-lane group.
// A block is NS = G*SPL slices; lane lg owns slices lg, lg+G, .. (SPL of them), so one load
// instruction of the group covers G*16 contiguous bytes and every 32-byte sector is requested once.
// lg = lane within group, gbase = first lane of the group, gmask = group's lane mask.
template <int G, int SPL>
__device__ __forceinline__ void extend_group(const SearchParams& P, int c, uint64_t& k, uint64_t& s,
                                             int lg, int gbase, unsigned gmask, unsigned& nblk) {
  constexpr int NS = G * SPL;
  constexpr int LOGB = (NS == 4) ? 7 : 8;
  constexpr unsigned BMASK = (1u << LOGB) - 1;
  static_assert(NS == 4 || NS == 8, "block is 4 or 8 slices");
  const uint64_t l = k + s;
  const uint64_t bk = k >> LOGB, bl = l >> LOGB;
  uint4 sk[SPL], sl[SPL];
#pragma unroll
  for (int i = 0; i < SPL; ++i) sk[i] = ldg_slice(P.blocks + bk * NS + lg + G * i);
  const bool two = (bl != bk);
#pragma unroll
  for (int i = 0; i < SPL; ++i) sl[i] = sk[i];
  if (two) {
#pragma unroll
    for (int i = 0; i < SPL; ++i) sl[i] = ldg_slice(P.blocks + bl * NS + lg + G * i);
  }
  nblk += two ? 2u : 1u;
  // superblock bases (L1-resident, tiny): acc[c] + Occ(c, sb << 32)
  const int64_t sbk = __ldg(P.sbase + (k >> 32) * 8 + c);
  const int64_t sbl = __ldg(P.sbase + (l >> 32) * 8 + c);
  const unsigned c0 = (c & 1) ? 0u : ~0u, c1 = (c & 2) ? 0u : ~0u, c2 = (c & 4) ? 0u : ~0u;
  const int offk = (int)((unsigned)k & BMASK), offl = (int)((unsigned)l & BMASK);
  unsigned v = 0;
#pragma unroll
  for (int i = 0; i < SPL; ++i) {
    const int j32 = 32 * (lg + G * i);
    const unsigned mk = (sk[i].y ^ c0) & (sk[i].z ^ c1) & (sk[i].w ^ c2);
    const unsigned ml = (sl[i].y ^ c0) & (sl[i].z ^ c1) & (sl[i].w ^ c2);
    const int nk_ = max(0, min(32, offk - j32));
    const int nl_ = max(0, min(32, offl - j32));
    const unsigned maskk = nk_ >= 32 ? ~0u : ((1u << nk_) - 1u);
    const unsigned maskl = nl_ >= 32 ? ~0u : ((1u << nl_) - 1u);
    v += __popc(mk & maskk) | (__popc(ml & maskl) << 16);
  }
#pragma unroll
  for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(gmask, v, o);
  unsigned ck, cl;
  if (NS == 8 || c <= 4) {
    // count slot c-1 lives in slice c-1 = lane (c-1) % G, register (c-1) / G
    const int slot = c - 1;
    // mask-select of the register (keeps sk/sl in registers: an if-chain becomes a local array)
    unsigned xk = 0, xl = 0;
    const int reg = slot / G;
#pragma unroll
    for (int i = 0; i < SPL; ++i) {
      const unsigned m = (reg == i) ? ~0u : 0u;
      xk |= sk[i].x & m;
      xl |= sl[i].x & m;
    }
    if (G > 1) {
      ck = __shfl_sync(gmask, xk, gbase + (slot % G));
      cl = __shfl_sync(gmask, xl, gbase + (slot % G));
    } else {
      ck = xk; cl = xl;
    }
  } else {  // 64-byte blocks, symbol N: side array
    ck = __ldg(P.cntN + bk);
    cl = __ldg(P.cntN + bl);
  }
  const uint64_t nk = (uint64_t)sbk + ck + (v & 0xffffu);
  const uint64_t nl = (uint64_t)sbl + cl + (v >> 16);
  k = nk;
  s = nl - nk;
}

__device__ __forceinline__ int comp6(int c) { return (c >= 1 && c <= 4) ? 5 - c : c; }

// read-character window: each lane of the group keeps 4 consecutive bases (one u32) of a
// 4G-byte aligned window of the concatenated read buffer, plus the prefetched neighbour window.
template <int G>
struct ReadWin {
  uint32_t cur, nxt;
  int64_t id;  // window index of cur (in units of 4G bytes); -1 = none
  __device__ __forceinline__ uint32_t load(const uint8_t* seq, int64_t w, int lg) const {
    return __ldg(reinterpret_cast<const uint32_t*>(seq) + w * G + lg);
  }
  // character at global byte position gp, walking in direction dir (+1 / -1); group-uniform
  __device__ __forceinline__ int get(const uint8_t* seq, int64_t gp, int dir, int lg, int gbase,
                                     unsigned gmask) {
    constexpr int LOGW = (G == 1) ? 2 : (G == 2) ? 3 : (G == 4) ? 4 : 5;
    const int64_t w = gp >> LOGW;
    if (w != id) {
      if (w == id + dir && id >= 0) cur = nxt;
      else cur = load(seq, w, lg);
      id = w;
      const int64_t wn = w + dir;
      nxt = wn >= 0 ? load(seq, wn, lg) : 0u;  // buffer is padded by one window at the end
    }
    const unsigned v = G > 1 ? __shfl_sync(gmask, cur, gbase + (int)((gp >> 2) & (G - 1))) : cur;
    return (int)((v >> ((gp & 3) * 8)) & 0xffu);
  }
};

template <int G, int SPL, int MINB>
__global__ void __launch_bounds__(256, MINB) k_sfs_search(const SearchParams P) {
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = (G == 32 ? ~0u : ((1u << G) - 1u)) << gbase;

  // per-group state (replicated in the G lanes)
  bool alive = true, have = false;
  int phase = 0;  // 0 = backward, 1 = forward
  uint32_t ridx = 0;
  int64_t roff = 0;
  int len = 0, pos = 0, begin = 0;
  uint64_t k = 0, s = 0;
  int chain_qs = -1, chain_end = -1;  // assemble: open superstring (assembler.cpp:34-56)
  unsigned n_ext = 0, n_blk = 0;
  ReadWin<G> win;
  win.id = -1; win.cur = 0; win.nxt = 0;

  auto emit = [&](int qs, int ln) {
    if (lg == 0) {
      unsigned long long slot = atomicAdd(P.out_count, 1ull);
      if (slot < P.out_cap) {
        // sort key: ascending qs when assembling, emit (descending-qs) order otherwise
        uint32_t sk = P.assemble ? (uint32_t)qs : ~(uint32_t)qs;
        P.out_key[slot] = ((uint64_t)ridx << 32) | sk;
        P.out_len[slot] = (uint32_t)ln;
      }
    }
  };
  auto on_sfs = [&](int qs, int ln) {
    if (!P.assemble) { emit(qs, ln); return; }
    // SFSs arrive with strictly decreasing qs; the previous one is (chain_qs, ..)
    if (chain_qs >= 0 && qs + ln > chain_qs) {
      chain_qs = qs;  // overlaps its successor: extend the open superstring to the left
    } else {
      if (chain_qs >= 0) emit(chain_qs, chain_end - chain_qs);
      chain_qs = qs;
      chain_end = qs + ln;
    }
  };
  auto finish_read = [&]() {
    if (P.assemble && chain_qs >= 0) emit(chain_qs, chain_end - chain_qs);
    chain_qs = -1;
    have = false;
    if (lg == 0) {
      atomicAdd(P.stats + 0, (unsigned long long)n_ext);
      atomicAdd(P.stats + 1, (unsigned long long)n_blk);
    }
    n_ext = 0; n_blk = 0;
  };

  while (__any_sync(0xffffffffu, alive)) {
    int c = 0;
    bool do_ext = false;
    if (alive) {
      if (!have) {
        unsigned long long w = 0;
        if (lg == 0) w = atomicAdd(P.work, 1ull);
        if (G > 1) w = __shfl_sync(gmask, w, gbase);
        if (w >= (unsigned long long)P.n_reads) {
          alive = false;
        } else {
          ridx = P.order ? P.order[w] : (uint32_t)w;
          roff = P.offs[ridx];
          len = (int)(P.offs[ridx + 1] - roff);
          if (len > 0) {
            have = true;
            phase = 0;
            pos = len - 1;
            win.id = -1;
            const int c0 = win.get(P.seq, roff + pos, -1, lg, gbase, gmask);
            k = (uint64_t)P.acc[c0];
            s = (uint64_t)(P.acc[c0 + 1] - P.acc[c0]);  // rb3_fmd_set_intv (ping_pong.cpp:12)
          }
        }
      }
      // advance the state machine to the next pending extension (ping_pong.cpp:15-47)
      while (have && !do_ext) {
        if (phase == 0) {
          if (s != 0 && pos > 0) {
            --pos;
            c = win.get(P.seq, roff + pos, -1, lg, gbase, gmask);
            do_ext = true;
          } else if (s != 0) {  // pos == 0 and still matching: done (ping_pong.cpp:24-25)
            finish_read();
          } else {              // mismatch at pos: switch to forward from here (:27-30)
            begin = pos;
            phase = 1;
            win.id = -1;
            const int c0 = comp6(win.get(P.seq, roff + pos, +1, lg, gbase, gmask));
            k = (uint64_t)P.acc[c0];
            s = (uint64_t)(P.acc[c0 + 1] - P.acc[c0]);
          }
        } else {
          if (s != 0 && pos + 1 < len) {
            ++pos;
            c = comp6(win.get(P.seq, roff + pos, +1, lg, gbase, gmask));
            do_ext = true;
          } else {
            // s == 0: P[begin..pos] is the SFS (ping_pong.cpp:39-41).  (s != 0 at the read end
            // cannot happen: P[begin..top] does not occur and pos <= top.)
            if (s != 0) ++pos;  // defensive: mirror the reference running onto P[l] = '$'
            on_sfs(begin, pos - begin + 1);
            if (begin == 0) {
              finish_read();
            } else {
              int nb = (P.overlap == 0) ? begin - 1 : pos + P.overlap;  // :44-47
              if (nb < 0) {
                finish_read();
              } else {
                if (nb > len - 1) nb = len - 1;
                pos = nb;
                phase = 0;
                win.id = -1;
                const int c0 = win.get(P.seq, roff + pos, -1, lg, gbase, gmask);
                k = (uint64_t)P.acc[c0];
                s = (uint64_t)(P.acc[c0 + 1] - P.acc[c0]);
              }
            }
          }
        }
      }
    }
    if (do_ext) {
      extend_group<G, SPL>(P, c, k, s, lg, gbase, gmask, n_blk);
      ++n_ext;
    }
  }
}

// ------------------------------------------------------------------------------ rank kernels
template <int G, int SPL>
__global__ void k_rank2a(const SearchParams P, const int64_t* __restrict__ qk, const int64_t* __restrict__ ql,
                         int64_t nq, int64_t n, int64_t* __restrict__ ok6, int64_t* __restrict__ ol6) {
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = ((1u << G) - 1u) << gbase;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  for (int64_t q = g0; q < nq; q += ngroups) {
    int64_t kk = qk[q], ll = ql[q];
    int64_t sumk = 0, suml = 0;
    for (int c = 1; c <= 5; ++c) {
      uint64_t k = (uint64_t)kk, s = (uint64_t)(ll - kk);
      unsigned nb = 0;
      extend_group<G, SPL>(P, c, k, s, lg, gbase, gmask, nb);
      int64_t okc = (int64_t)k - P.acc[c], olc = okc + (int64_t)s;
      sumk += okc; suml += olc;
      if (lg == 0) { ok6[q * 6 + c] = okc; ol6[q * 6 + c] = olc; }
    }
    if (lg == 0) { ok6[q * 6] = kk - sumk; ol6[q * 6] = ll - suml; }
    (void)n;
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

// one extension per query, queries generated on the fly: k uniform in [0, n - delta)
template <int G, int SPL>
__global__ void __launch_bounds__(256) k_rank_bench(const SearchParams P, int64_t nq, int64_t n, int64_t delta,
                                                    uint64_t seed, unsigned long long* __restrict__ sink,
                                                    unsigned long long* __restrict__ blocks_touched) {
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = ((1u << G) - 1u) << gbase;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  unsigned long long acc = 0;
  unsigned nb = 0;
  for (int64_t q = g0; q < nq; q += ngroups) {
    uint64_t r = splitmix64(seed + (uint64_t)q);
    uint64_t k = __umul64hi(r, (uint64_t)(n - delta));  // uniform in [0, n - delta), no 64-bit division
    uint64_t s = (uint64_t)delta;
    int c = 1 + (int)((r >> 60) & 3);
    extend_group<G, SPL>(P, c, k, s, lg, gbase, gmask, nb);
    acc += k + s;
  }
  if (lg == 0) {
    atomicAdd(blocks_touched, (unsigned long long)nb);
    if (acc == 0x123456789ULL) atomicAdd(sink, acc);
  }
}

// ------------------------------------------------------------------------------ v2 kernel
// Thread-per-read state machine + TMA-staged index blocks (128-byte blocks only).
//
// ncu on the lane-group kernel showed it issue-bound (68 % issue slots, ~49 warp instructions per
// extension: every lane of a group replays the whole state machine) with only 4-8 dependent-load
// chains per warp.  Here every THREAD walks its own read, so a warp carries 32 chains, and the
// 128-byte index blocks are fetched by the TMA unit: one cp.async.bulk (UBLKCP) per block straight
// into the thread's shared-memory slot, completion counted on a per-warp mbarrier.  No register is
// tied up by a load in flight; one request = one full 128-byte line.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
// streamed batch: wait until the chunk holding a read's last base has landed.  The chunks are fed by
// copies (and, for packed input, unpack kernels) queued behind this kernel; if they never arrive --
// a scheduling assumption broken -- give up after 60 s instead of hanging the device: stats[3] tells
// the host, which fails the call.
__device__ __forceinline__ bool wait_chunk(const volatile unsigned int* flag, unsigned long long* stall_flag) {
  if (*flag != 0u) return true;
  const unsigned long long t0 = globaltimer_ns();
  while (*flag == 0u) {
    __nanosleep(500);
    if (globaltimer_ns() - t0 > 60000000000ull || *reinterpret_cast<volatile unsigned long long*>(stall_flag) != 0ull) {
      atomicExch(stall_flag, 1ull);
      return false;
    }
  }
  return true;
}

// per-thread read window: 8 bases in a u64 + the prefetched neighbour in walking direction.
// A direction switch keeps `cur` (the pivot base is in it) and only re-aims the prefetch, so the
// warp never waits on a read-window load except at the first base of a read.
struct ReadWin1 {
  uint64_t cur, nxt;
  int64_t id, nid;  // window ids held in cur / nxt (-1 = none)
  __device__ __forceinline__ void reset() { id = -1; nid = -1; cur = 0; nxt = 0; }
  __device__ __forceinline__ int get(const uint8_t* seq, int64_t gp, int dir) {
    const uint64_t* p = reinterpret_cast<const uint64_t*>(seq);
    const int64_t w = gp >> 3;
    if (w != id) {
      cur = (w == nid) ? nxt : __ldcg(p + w);
      id = w;
    }
    const int64_t wn = w + dir;
    if (wn != nid) {
      nid = wn;
      nxt = wn >= 0 ? __ldcg(p + wn) : 0ull;  // consumed >= 1 step later
    }
    return (int)((cur >> ((gp & 7) * 8)) & 0xffull);
  }
};

// issue the TMA copies of the block(s) needed by the extension of [k, k+s); returns bl != bk
__device__ __forceinline__ bool tma_issue(const SearchParams& P, uint64_t k, uint64_t s, uint32_t slot, uint32_t bar) {
  const uint64_t bk = k >> 8, bl = (k + s) >> 8;
  const bool two = bl != bk;
  mbar_arrive_expect_tx(bar, two ? 256u : 128u);
  bulk_g2s(slot, P.blocks + bk * 8, 128u, bar);
  if (two) bulk_g2s(slot + 128u, P.blocks + bl * 8, 128u, bar);
  return two;
}

// rank of symbol c (1..4) at in-block offset `off` from a staged block: base count slot + in-block
// sample (slots 5..7) + popcount over ONE 64-symbol sub-block.  rot = slice rotation of this thread.
__device__ __forceinline__ unsigned staged_occ(const uint4* blk, int off, int c, int rot) {
  const int sub = off >> 6, r = off & 63;
  const uint4 a = blk[(2 * sub + rot) & 7];
  const uint4 b = blk[(2 * sub + 1 + rot) & 7];
  const unsigned c0 = (c & 1) ? 0u : ~0u, c1 = (c & 2) ? 0u : ~0u, c2 = (c & 4) ? 0u : ~0u;
  const unsigned m0 = (a.y ^ c0) & (a.z ^ c1) & (a.w ^ c2);
  const unsigned m1 = (b.y ^ c0) & (b.z ^ c1) & (b.w ^ c2);
  const unsigned k0 = r >= 32 ? ~0u : ((1u << r) - 1u);
  const unsigned k1 = r > 32 ? ((1u << (r - 32)) - 1u) : 0u;
  const unsigned base = blk[(c - 1 + rot) & 7].x;
  const unsigned cum = sub ? ((blk[(4 + sub + rot) & 7].x >> (8 * (c - 1))) & 0xffu) : 0u;
  return base + cum + __popc(m0 & k0) + __popc(m1 & k1);
}

// finish the extension from the staged blocks (thread-local).  ACGT take the sub-block path; N
// (rare: only reads that carry N) scans the whole block.
__device__ __forceinline__ void tma_consume(const SearchParams& P, const uint4* slot, bool two, int c, uint64_t& k,
                                            uint64_t& s, int rot) {
  const uint64_t l = k + s;
  const int offk = (int)((unsigned)k & 255u), offl = (int)((unsigned)l & 255u);
  const uint4* sl_ptr = two ? slot + 8 : slot;
  unsigned rk, rl;
  if (c <= 4) {
    rk = staged_occ(slot, offk, c, rot);
    rl = staged_occ(sl_ptr, offl, c, rot);
  } else {
    const unsigned c0 = (c & 1) ? 0u : ~0u, c1 = (c & 2) ? 0u : ~0u, c2 = (c & 4) ? 0u : ~0u;
    unsigned pk = 0, pl = 0;
#pragma unroll 1
    for (int j = 0; j < 8; ++j) {
      const uint4 a = slot[(j + rot) & 7];
      const uint4 b = sl_ptr[(j + rot) & 7];
      const int nk_ = max(0, min(32, offk - 32 * j));
      const int nl_ = max(0, min(32, offl - 32 * j));
      pk += __popc((a.y ^ c0) & (a.z ^ c1) & (a.w ^ c2) & (nk_ >= 32 ? ~0u : ((1u << nk_) - 1u)));
      pl += __popc((b.y ^ c0) & (b.z ^ c1) & (b.w ^ c2) & (nl_ >= 32 ? ~0u : ((1u << nl_) - 1u)));
    }
    rk = slot[(4 + rot) & 7].x + pk;    // count slot 4 = N
    rl = sl_ptr[(4 + rot) & 7].x + pl;
  }
  const int64_t sbk = __ldg(P.sbase + (k >> 32) * 8 + c);
  const int64_t sbl = __ldg(P.sbase + (l >> 32) * 8 + c);
  const uint64_t nk = (uint64_t)sbk + rk;
  const uint64_t nl = (uint64_t)sbl + rl;
  k = nk;
  s = nl - nk;
}

// ---- LDGSTS variant of the staging: the warp copies its 32 threads' blocks cooperatively, 8 lanes
// x 16 bytes per block and 4 blocks per round, so every request is one coalesced 128-byte line and
// no per-lane TMA issue loop (UBLKCP takes warp-uniform operands) is needed.
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
// bk/bl: block ids of this thread's pending extension (bl == bk if one block, both NOBLK if none)
constexpr uint32_t NOBLK = 0xffffffffu;
__device__ __forceinline__ void cpa_fetch(const SearchParams& P, uint32_t bk, uint32_t bl, uint32_t warp_stage_s,
                                          int lane) {
  const int sub = lane >> 3, j = lane & 7;
#pragma unroll
  for (int r = 0; r < 8; ++r) {
    const int t = 4 * r + sub;
    const uint32_t xk = __shfl_sync(0xffffffffu, bk, t);
    const uint32_t xl = __shfl_sync(0xffffffffu, bl, t);
    // slice j of thread t's block sits at physical slot (j + t) & 7: threads reading the same logical
    // slice then hit different shared-memory banks
    const uint32_t dst = warp_stage_s + (uint32_t)(t * 256 + ((j + t) & 7) * 16);
    if (xk != NOBLK) cp_async16(dst, P.blocks + (uint64_t)xk * 8 + j);
    if (xl != xk) cp_async16(dst + 128u, P.blocks + (uint64_t)xl * 8 + j);
  }
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}

// ---- located-match mode -------------------------------------------------------------------
// Once the interval of the current match has size 1 and its row carries an SA sample, the match is
// ONE text position p and "does cW occur" is "T[p-1] == c".  The ping-pong walk then turns into a
// byte compare of the read against the text: the backward phase walks both downwards; the forward
// phase (a backward search of rc(P[b..e])) is mirrored onto the other strand, where the same
// occurrence reads P[b..e] left to right, so it walks both upwards.  Results are identical to the
// rank walk by construction (same occurrence set); 16 bases cost ~30 instructions and two
// sequential 24-byte reads instead of 16 random 128-byte block fetches.
__device__ __forceinline__ uint64_t funnel64(uint64_t lo, uint64_t hi, int sh) {  // bytes [sh/8, sh/8+8) of hi:lo
  return sh ? ((lo >> sh) | (hi << (64 - sh))) : lo;
}
struct Win16 { uint64_t w0, w1, w2; int sh; };
// the 16 bytes at [a, a+16) of `base` (a may be < 0: those bytes are never looked at)
__device__ __forceinline__ void load16(const uint8_t* base, int64_t a, Win16& w, int64_t min_word) {
  const uint64_t* p = reinterpret_cast<const uint64_t*>(base);
  const int64_t i = a >> 3;
  w.sh = (int)(a & 7) * 8;
  w.w0 = __ldcg(p + max(i, min_word));
  w.w1 = __ldcg(p + max(i + 1, min_word));
  w.w2 = __ldcg(p + max(i + 2, min_word));
}
// text position of the mirror image (other strand) of text position x
__device__ __forceinline__ int64_t mirror_pos(const SearchParams& P, int64_t x) {
  int64_t lo = 0, hi = P.n_contigs - 1;
  while (lo < hi) {
    const int64_t mid = (lo + hi + 1) >> 1;
    if (__ldg(P.tstart + mid) <= x) lo = mid; else hi = mid - 1;
  }
  return __ldg(P.tstart + lo) + __ldg(P.tstart + lo + 1) - 2 - x;
}

// ---- K-mer jump table ----------------------------------------------------------------------
// 2-bit codes of the first K (<= 16) bytes of a 16-byte window; false if one of them is not A,C,G,T
__device__ __forceinline__ bool words_kmer(uint64_t lo, uint64_t hi, int K, uint32_t& code);
__device__ __forceinline__ bool window_kmer(const Win16& w, int K, uint32_t& code) {
  return words_kmer(funnel64(w.w0, w.w1, w.sh), funnel64(w.w1, w.w2, w.sh), K, code);
}
// lo = bytes 0..7, hi = bytes 8..15 of the window
__device__ __forceinline__ bool words_kmer(uint64_t lo, uint64_t hi, int K, uint32_t& code) {
  const uint64_t H = 0x8080808080808080ull, L1 = 0x0101010101010101ull;
  const uint64_t rlo = (lo | H) - L1, rhi = (hi | H) - L1;   // per byte: 0x80 + (b - 1), no borrow across bytes
  uint64_t blo = ((rlo ^ H) & 0xFCFCFCFCFCFCFCFCull) | (lo & 0xF8F8F8F8F8F8F8F8ull);
  uint64_t bhi = ((rhi ^ H) & 0xFCFCFCFCFCFCFCFCull) | (hi & 0xF8F8F8F8F8F8F8F8ull);
  if (K < 8) blo &= (1ull << (8 * K)) - 1ull;
  if (K <= 8) bhi = 0; else if (K < 16) bhi &= (1ull << (8 * (K - 8))) - 1ull;
  if (blo | bhi) return false;
  auto squeeze = [](uint64_t x) {   // bits 8i..8i+1 of x -> bits 2i..2i+1
    x &= 0x0303030303030303ull;
    x = (x | (x >> 6)) & 0x000F000F000F000Full;
    x = (x | (x >> 12)) & 0x000000FF000000FFull;
    x = (x | (x >> 24)) & 0xFFFFull;
    return (uint32_t)x;
  };
  code = squeeze(rlo) | (squeeze(rhi) << 16);
  if (K < 16) code &= (1u << (2 * K)) - 1u;
  return true;
}
// code of the reverse complement of a K-mer code
__device__ __forceinline__ uint32_t kmer_rc(uint32_t code, int K) {
  uint32_t r = __brev(~code);                                             // complement, then reverse bits
  r = ((r & 0xAAAAAAAAu) >> 1) | ((r & 0x55555555u) << 1);                // un-reverse inside each 2-bit group
  return K < 16 ? r >> (32 - 2 * K) : r;
}
// acc[c] + Occ(c, pos) straight from global memory (128-byte blocks with in-block samples)
__device__ __forceinline__ uint64_t occ_global(const SearchParams& P, int c, uint64_t pos) {
  return (uint64_t)__ldg(P.sbase + (pos >> 32) * 8 + c) + staged_occ(P.blocks + (pos >> 8) * 8, (int)((unsigned)pos & 255u), c, 0);
}
// level j of the table from level j-1: interval(cX) = extend(interval(X), c); first base in the low bits.
// Intermediate levels keep exact (start, size) pairs; the last level is packed into the table.
__global__ void k_kmer_level(const SearchParams P, const ulonglong2* __restrict__ prev, ulonglong2* __restrict__ cur,
                             uint64_t* __restrict__ packed, int j, const int64_t n) {
  const uint64_t total = 1ull << (2 * j), stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t y = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; y < total; y += stride) {
    ulonglong2 e = make_ulonglong2(0ull, (unsigned long long)n);
    if (j > 1) e = prev[y >> 2];
    ulonglong2 o = make_ulonglong2(0ull, 0ull);
    if (e.y != 0) {
      const int c = (int)(y & 3) + 1;
      const uint64_t nk = occ_global(P, c, e.x), nl = occ_global(P, c, e.x + e.y);
      o = make_ulonglong2(nk, nl - nk);
    }
    if (packed) packed[y] = o.x | (min((uint64_t)o.y, KMT_SAT) << 40);
    else cur[y] = o;
  }
}

constexpr int TMA_WARPS = 3;  // warps per CTA; 8 KB of staging per warp -> 9 CTAs = 27 warps per SM

template <int MINB, int MODE>  // MODE 0: TMA bulk copies (UBLKCP), 1: cooperative cp.async (LDGSTS)
__global__ void __launch_bounds__(TMA_WARPS * 32, MINB) k_sfs_search_tma(const SearchParams P) {
  __shared__ __align__(128) uint4 stage[TMA_WARPS * 32 * 16];  // [warp][lane][2 blocks][8 slices]
  __shared__ __align__(8) uint64_t bars[TMA_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar = smem_u32(&bars[warp]);
  uint4* my = stage + (warp * 32 + lane) * 16;
  const uint32_t my_s = smem_u32(my);
  if (lane == 0) {
    mbar_init(bar, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  uint32_t parity = 0;

  bool alive = true, have = false;
  int phase = 0;
  uint32_t ridx = 0;
  int64_t roff = 0;
  int len = 0, pos = 0, begin = 0;
  uint64_t k = 0, s = 0;
  int chain_qs = -1, chain_end = -1;
  unsigned n_ext = 0, n_blk = 0, n_txt = 0;
  bool tmode = false;   // located-match mode: T[g + delta] is aligned with read byte g
  int64_t delta = 0;
  const unsigned ss_mask = (1u << P.ss_log) - 1u;
  ReadWin1 win;
  win.reset();

  auto emit = [&](int qs, int ln) {
    unsigned long long o = atomicAdd(P.out_count, 1ull);
    if (o < P.out_cap) {
      uint32_t sk = P.assemble ? (uint32_t)qs : ~(uint32_t)qs;
      P.out_key[o] = ((uint64_t)ridx << 32) | sk;
      P.out_len[o] = (uint32_t)ln;
    }
  };
  auto on_sfs = [&](int qs, int ln) {
    if (!P.assemble) { emit(qs, ln); return; }
    if (chain_qs >= 0 && qs + ln > chain_qs) {
      chain_qs = qs;
    } else {
      if (chain_qs >= 0) emit(chain_qs, chain_end - chain_qs);
      chain_qs = qs;
      chain_end = qs + ln;
    }
  };
  auto finish_read = [&]() {
    if (P.assemble && chain_qs >= 0) emit(chain_qs, chain_end - chain_qs);
    chain_qs = -1;
    have = false;
    tmode = false;
    atomicAdd(P.stats + 0, (unsigned long long)n_ext + n_txt);
    atomicAdd(P.stats + 1, (unsigned long long)n_blk);
    if (n_txt) atomicAdd(P.stats + 2, (unsigned long long)n_txt);
    n_ext = 0; n_blk = 0; n_txt = 0;
  };
  auto set_intv = [&](int c0) {
    k = (uint64_t)P.acc[c0];
    s = (uint64_t)(P.acc[c0 + 1] - P.acc[c0]);  // rb3_fmd_set_intv (ping_pong.cpp:12,30)
    tmode = false;
  };
  // (re)start the walk at pivot `pos` in the direction of `phase`: rb3_fmd_set_intv + the first K-1
  // extensions in one table lookup when the K-mer at the pivot occurs, else from the pivot base alone
  auto start_walk = [&]() {
    const int K = P.kmer_k;
    if (P.kmt != nullptr && (phase ? pos + K <= len : pos + 1 >= K)) {
      Win16 w;
      load16(P.seq, roff + (phase ? pos : pos - K + 1), w, 0);
      uint32_t code;
      if (window_kmer(w, K, code)) {
        const uint64_t e = __ldg(P.kmt + (phase ? kmer_rc(code, K) : code));
        const uint64_t sz = e >> 40;
        if (sz != 0 && sz != KMT_SAT) {
          k = e & ((1ull << 40) - 1ull);
          s = sz;
          pos += phase ? K - 1 : -(K - 1);
          n_ext += (unsigned)(K - 1);
          tmode = false;
          win.id = -1; win.nid = -1;
          return;
        }
      }
    }
    const int ch = win.get(P.seq, roff + pos, phase ? 1 : -1);
    set_intv(phase ? comp6(ch) : ch);
  };

  while (__any_sync(0xffffffffu, alive)) {
    int c = 0;
    bool do_ext = false, do_txt = false;
    if (alive) {
      if (!have) {
        const unsigned long long w = atomicAdd(P.work, 1ull);
        if (w >= (unsigned long long)P.n_reads) {
          alive = false;
        } else {
          ridx = P.order ? P.order[w] : (uint32_t)w;
          roff = P.offs[ridx];
          len = (int)(P.offs[ridx + 1] - roff);
          if (P.ready && len > 0) {  // streamed batch: wait until the chunk holding the last base landed
            if (!wait_chunk(P.ready + (roff + len - 1) / P.chunk_bytes, P.stats + 3)) { alive = false; len = 0; }
            __threadfence();
          }
          if (len > 0) {
            have = true;
            phase = 0;
            pos = len - 1;
            win.reset();
            start_walk();
          }
        }
      }
      // state machine of ping_pong.cpp:15-47, advanced to the next pending extension.  Both
      // directions share ONE fast path (dir = -1 backward, +1 forward) so that threads in
      // different phases do not serialise the warp.
      while (have && !do_ext && !do_txt) {
        const int dir = phase ? 1 : -1;
        const bool room = phase ? (pos + 1 < len) : (pos > 0);
        if (s != 0 && room) {
          if (!tmode && P.text != nullptr && s == 1 && ((unsigned)k & ss_mask) == 0u) {
            // unique and on a sampled row: locate.  Backward: P[pos..] starts at T[p].  Forward:
            // rc(P[begin..pos]) starts at T[p]; its mirror image is P[begin..pos] read left to right.
            const int64_t p = (int64_t)__ldg(P.ssa + (k >> P.ss_log));
            delta = phase ? mirror_pos(P, p + (pos - begin)) - (roff + begin) : p - (roff + pos);
            tmode = true;
          }
          if (tmode) {
            do_txt = true;
          } else {
            pos += dir;
            const int ch = win.get(P.seq, roff + pos, dir);
            c = phase ? comp6(ch) : ch;
            do_ext = true;
          }
        } else if (phase == 0) {
          if (s != 0) {
            finish_read();  // reached the read start still matching (ping_pong.cpp:24-25)
          } else {          // mismatch at pos: forward search from here (ping_pong.cpp:27-30)
            begin = pos;
            phase = 1;
            start_walk();
          }
        } else {
          if (s != 0) ++pos;               // defensive: cannot happen (see k_sfs_search)
          on_sfs(begin, pos - begin + 1);  // ping_pong.cpp:39-41
          if (begin == 0) {
            finish_read();
          } else {
            int nb = (P.overlap == 0) ? begin - 1 : pos + P.overlap;  // ping_pong.cpp:44-47
            if (nb < 0) {
              finish_read();
            } else {
              if (nb > len - 1) nb = len - 1;
              pos = nb;
              phase = 0;
              start_walk();
            }
          }
        }
      }
    }
    // located threads: issue the loads of the next 16 read / text bytes now, so that they are in
    // flight together with the index blocks of the warp's rank-mode threads
    Win16 rw, tw;
    int64_t a0 = 0;
    if (do_txt) {
      a0 = roff + pos + (phase ? 1 : -16);   // window [a0, a0+16): above pos (forward) / below pos (backward)
      load16(P.seq, a0, rw, 0);
      load16(P.text, a0 + delta, tw, -(int64_t)(TEXT_PAD / 8));
    }
    bool two = false;
    if (MODE == 0) {
      if (do_ext) two = tma_issue(P, k, s, my_s, bar);
      else mbar_arrive(bar);
      while (!mbar_try_wait(bar, parity)) {}
      parity ^= 1u;
    } else {
      const uint32_t bk = do_ext ? (uint32_t)(k >> 8) : NOBLK;
      const uint32_t bl = do_ext ? (uint32_t)((k + s) >> 8) : NOBLK;
      two = bl != bk;
      cpa_fetch(P, bk, bl, smem_u32(stage + warp * 32 * 16), lane);
    }
    if (do_ext) {
      tma_consume(P, my, two, c, k, s, MODE == 1 ? lane : 0);
      ++n_ext;
      n_blk += two ? 2u : 1u;
    }
    if (do_txt) {
      const uint64_t xlo = funnel64(rw.w0, rw.w1, rw.sh) ^ funnel64(tw.w0, tw.w1, tw.sh);   // bytes a0 .. a0+7
      const uint64_t xhi = funnel64(rw.w1, rw.w2, rw.sh) ^ funnel64(tw.w1, tw.w2, tw.sh);   // bytes a0+8 .. a0+15
      int nb, mt;  // bases available in walking direction (<= 16), leading bases that match
      if (phase) {
        nb = min(16, len - 1 - pos);
        mt = xlo ? (__ffsll((long long)xlo) - 1) >> 3 : 8 + (xhi ? (__ffsll((long long)xhi) - 1) >> 3 : 8);
      } else {
        nb = min(16, pos);
        mt = xhi ? __clzll((long long)xhi) >> 3 : 8 + (xlo ? __clzll((long long)xlo) >> 3 : 8);
      }
      const int dir = phase ? 1 : -1;
      if (mt >= nb) {          // all nb extensions succeed (interval stays of size 1)
        pos += dir * nb;
        n_txt += (unsigned)nb;
      } else {                 // extension number mt+1 fails: same state as a rank walk ending in s == 0
        pos += dir * (mt + 1);
        n_txt += (unsigned)(mt + 1);
        s = 0;
        tmode = false;
      }
      win.id = -1; win.nid = -1;   // the byte window no longer follows pos
    }
    if (MODE == 1) __syncwarp();  // staging slots are rewritten by other lanes next iteration
  }
}

// ping_pong.cpp:90-94 on the device: htslib nt16 code -> ASCII (seq_nt16_str) -> nt6 (seq_nt6_table):
// A=1 C=2 G=4 T=8 map to 1..4, every other code (=, IUPAC, N) to 5
__device__ __forceinline__ uint8_t nt16_to_nt6(unsigned c) {
  return c == 1u ? 1 : c == 2u ? 2 : c == 4u ? 3 : c == 8u ? 4 : 5;
}
// 16 bases of one read, starting at base j (any parity), as 16 nt6 bytes.  The nine packed bytes that
// can hold them are taken from two aligned 8-byte loads.
__device__ __forceinline__ uint4 unpack16(const uint8_t* __restrict__ seq4, int64_t byte0, int odd) {
  const uint64_t* p = reinterpret_cast<const uint64_t*>(seq4);
  const int64_t wi = byte0 >> 3;
  const int sh = (int)(byte0 & 7) * 8;
  const uint64_t w0 = __ldcg(p + wi), w1 = __ldcg(p + wi + 1);
  uint64_t x = funnel64(w0, w1, sh);                              // packed bytes byte0 .. byte0+7
  const uint64_t ninth = (w1 >> sh) & 0xffull;                    // packed byte byte0+8
  // BAM keeps the first base of a byte in the HIGH nibble: swap so that nibble k of x is base k
  x = ((x & 0x0F0F0F0F0F0F0F0Full) << 4) | ((x >> 4) & 0x0F0F0F0F0F0F0F0Full);
  if (odd) x = (x >> 4) | ((ninth >> 4) << 60);
  // nt16 -> nt6 through a 16-entry nibble table: 1 -> 1, 2 -> 2, 4 -> 3, 8 -> 4, everything else -> 5
  const uint64_t LUT = 0x5555555455535215ull;
  uint32_t o[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t v = 0;
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const unsigned n = (unsigned)(x >> (4 * (4 * q + t))) & 15u;
      v |= (uint32_t)((LUT >> (4 * n)) & 15ull) << (8 * t);
    }
    o[q] = v;
  }
  return make_uint4(o[0], o[1], o[2], o[3]);
}
// one warp unpacks the bases of read r that fall into base positions [A, B).  Lanes own 16-byte
// aligned groups of the OUTPUT (coalesced 16-byte stores, two groups per lane and step); the ragged
// head and tail of the range go base by base.
__device__ __forceinline__ void unpack4_read(const uint8_t* __restrict__ seq4, const int64_t* __restrict__ seq4_offs,
                                             const int64_t* __restrict__ offs, int64_t r, int64_t A, int64_t B,
                                             uint8_t* __restrict__ out, int lane) {
  const int64_t o = offs[r], l = offs[r + 1] - o, pb = seq4_offs[r];
  const int64_t lo = max(o, A), hi = min(o + l, B);
  if (lo >= hi) return;
  auto one = [&](int64_t g) {   // output position g
    const int64_t j = g - o;
    const uint8_t b = __ldcg(seq4 + pb + (j >> 1));
    out[g] = nt16_to_nt6((j & 1) ? (b & 15u) : (unsigned)(b >> 4));
  };
  const int64_t body0 = min(hi, (int64_t)((lo + 15) & ~15LL)), body1 = body0 + ((hi - body0) & ~15LL);
  if (lo + lane < body0) one(lo + lane);                       // head: fewer than 16 positions
  for (int64_t g = body0 + 16 * lane; g < body1; g += 1024) {
    const int64_t g2 = g + 512;
    const int64_t j = g - o, j2 = g2 - o;
    const uint4 v = unpack16(seq4, pb + (j >> 1), (int)(j & 1));
    uint4 v2 = make_uint4(0, 0, 0, 0);
    if (g2 < body1) v2 = unpack16(seq4, pb + (j2 >> 1), (int)(j2 & 1));
    *reinterpret_cast<uint4*>(out + g) = v;
    if (g2 < body1) *reinterpret_cast<uint4*>(out + g2) = v2;
  }
  if (body1 + lane < hi) one(body1 + lane);                    // tail: fewer than 16 positions
}
// whole batch at once (small batches, no streaming)
__global__ void __launch_bounds__(128) k_unpack4(const uint8_t* __restrict__ seq4, const int64_t* __restrict__ seq4_offs,
                                                  const int64_t* __restrict__ offs, int64_t n_reads, int64_t total,
                                                  uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n_reads; r += nwarps)
    unpack4_read(seq4, seq4_offs, offs, r, 0, total, out, lane);
}
// the unpacking CTAs of a packed streamed launch (SearchParams::n_unpack of them): chunk after chunk,
// wait for the copy engine, unpack this CTA's share of the chunk's reads, count in; the last CTA to
// finish a chunk raises its ready flag for the searching CTAs
__device__ void unpack_cta_loop(const SearchParams& P) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)P.n_unpack * (blockDim.x >> 5);
  const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int c = 0; c < P.n_chunks; ++c) {
    if (!wait_chunk(P.arrived + c, P.stats + 3)) return;
    __threadfence();
    const int64_t A = (int64_t)c * P.chunk_bytes, B = min(A + P.chunk_bytes, P.total);
    for (int64_t r = P.chunk_r0[c] + w0; r <= P.chunk_r1[c]; r += nwarps)
      unpack4_read(P.seq4, P.seq4_offs, P.offs, r, A, B, P.seq_w, lane);
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      if (atomicAdd(P.chunk_done + c, 1u) == (unsigned)P.n_unpack - 1u) {
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(P.ready_w + c) = 1u;
      }
    }
  }
}

// the same for the 2-bit transport: per chunk the decoder the host chose, and the last CTA to finish a chunk
// patches the positions that have no 2-bit form before it raises the flag
__device__ void unpack_cta_loop2(const SearchParams& P) {
  __shared__ unsigned int s_last;
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)P.n_unpack * (blockDim.x >> 5);
  const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  for (int c = 0; c < P.n_chunks; ++c) {
    if (!wait_chunk(P.arrived + c, P.stats + 3)) return;
    __threadfence();
    const unsigned int mode = __ldcg(P.chunk_mode + c);
    const int64_t A = (int64_t)c * P.chunk_bytes, B = min(A + P.chunk_bytes, P.total);
    for (int64_t r = P.chunk_r0[c] + w0; r <= P.chunk_r1[c]; r += nwarps) {
      if (mode == 2u) unpack2_read(P.pk2, P.pk2_offs, P.offs, r, A, B, P.seq_w, lane);
      else unpack4_read(P.seq4, P.seq4_offs, P.offs, r, A, B, P.seq_w, lane);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      s_last = atomicAdd(P.chunk_done + c, 1u) == (unsigned)P.n_unpack - 1u ? 1u : 0u;
    }
    __syncthreads();
    if (s_last) {   // every other CTA has stored and fenced its share of the chunk
      __threadfence();
      if (mode == 2u) {
        const unsigned int n = __ldcg(P.exc_n + c);
        const int64_t* ep = P.exc_pos + (int64_t)c * P.exc_cap;
        for (unsigned int i = threadIdx.x; i < n; i += blockDim.x) P.seq_w[__ldcg(ep + i)] = 5;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        __threadfence();
        *reinterpret_cast<volatile unsigned int*>(P.ready_w + c) = 1u;
      }
    }
    __syncthreads();
  }
}

// 32-byte window at an arbitrary byte address: three aligned 16-byte loads
struct Win32 { uint4 q0, q1, q2; int sh; };   // sh = byte offset of the window inside q0 (0..15)
__device__ __forceinline__ void load32(const uint8_t* base, int64_t a, Win32& w, int64_t min_q) {
  const uint4* p = reinterpret_cast<const uint4*>(base);
  const int64_t i = a >> 4;
  w.sh = (int)(a & 15);
  w.q0 = __ldcg(p + max(i, min_q));
  w.q1 = __ldcg(p + max(i + 1, min_q));
  w.q2 = __ldcg(p + max(i + 2, min_q));
}
// the four 8-byte words of the window, lowest address first
__device__ __forceinline__ void win32_words(const Win32& w, uint64_t o[4]) {
  const uint64_t v0 = (uint64_t)w.q0.x | ((uint64_t)w.q0.y << 32), v1 = (uint64_t)w.q0.z | ((uint64_t)w.q0.w << 32);
  const uint64_t v2 = (uint64_t)w.q1.x | ((uint64_t)w.q1.y << 32), v3 = (uint64_t)w.q1.z | ((uint64_t)w.q1.w << 32);
  const uint64_t v4 = (uint64_t)w.q2.x | ((uint64_t)w.q2.y << 32), v5 = (uint64_t)w.q2.z | ((uint64_t)w.q2.w << 32);
  const bool hi = w.sh >= 8;
  const int sh = (w.sh & 7) * 8;
  const uint64_t a0 = hi ? v1 : v0, a1 = hi ? v2 : v1, a2 = hi ? v3 : v2, a3 = hi ? v4 : v3, a4 = hi ? v5 : v4;
  o[0] = funnel64(a0, a1, sh); o[1] = funnel64(a1, a2, sh); o[2] = funnel64(a2, a3, sh); o[3] = funnel64(a3, a4, sh);
}
// staging for the lanes that have an extension pending (a minority once located matches and the jump
// table carry most of the walk): four of them per round, eight lanes x 16 bytes per block
__device__ __forceinline__ void cpa_issue_sparse(const SearchParams& P, uint32_t bk, uint32_t bl, uint32_t warp_stage_s,
                                                 int lane, uint8_t* pend /* 32 bytes of shared memory per warp */) {
  // compact the pending lanes into pend[0..n): rank among the pending = popcount of the lower lanes
  const unsigned m = __ballot_sync(0xffffffffu, bk != NOBLK);
  const int n = __popc(m);
  if (bk != NOBLK) pend[__popc(m & ((1u << lane) - 1u))] = (uint8_t)lane;
  __syncwarp();
  const int sub = lane >> 3, j = lane & 7;
  for (int i = sub; i < ((n + 3) & ~3); i += 4) {   // warp-uniform trip count; 8-lane group `sub` takes entry i
    const bool on = i < n;
    const unsigned t = on ? pend[i] : 0u;
    const uint32_t xk = __shfl_sync(0xffffffffu, bk, (int)t);
    const uint32_t xl = __shfl_sync(0xffffffffu, bl, (int)t);
    if (on) {
      const uint32_t dst = warp_stage_s + (uint32_t)(t * 256u + ((j + t) & 7u) * 16u);
      cp_async16(dst, P.blocks + (uint64_t)xk * 8 + j);
      if (xl != xk) cp_async16(dst + 128u, P.blocks + (uint64_t)xl * 8 + j);
    }
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void cpa_wait() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}

// Located match, warp-cooperative: all 32 lanes compare 32 bytes each of ONE lane's read against the
// text -- 1024 bases per step from coalesced loads (the read window of lane L is the L-th 32-byte
// group on the read's 16-byte grid in walking direction; the text side is funnel-shifted).
// g = global position of the last matched base, lim = bases left in walking direction (>= 1).
// Returns the number of extensions that succeed (<= lim); *mismatch says whether the next one fails.
__device__ __forceinline__ int coop_text_step(const SearchParams& P, int64_t g, int64_t delta, int fwd, int lim, int lane,
                                              bool* mismatch) {
  const int64_t a0 = fwd ? ((g + 1) & ~15LL) : ((g + 15) & ~15LL);   // forward: first window starts here; backward: ends here
  const int skip = fwd ? (int)(g + 1 - a0) : (int)(a0 - g);          // bytes of the first window behind the walk
  const int64_t aw = fwd ? a0 + 32 * lane : a0 - 32 * (lane + 1);    // this lane's window [aw, aw + 32)
  const int nvalid = max(0, min(32, lim + skip - 32 * lane));        // its leading bytes (walking order) inside the limit
  int mi = 32;                                                        // leading bytes that match (walking order)
  if (nvalid > 0) {
    const uint4* rp = reinterpret_cast<const uint4*>(P.seq);
    const int64_t qi = aw >> 4;
    const uint4 r0 = __ldcg(rp + max(qi, (int64_t)0)), r1 = __ldcg(rp + max(qi + 1, (int64_t)0));
    Win32 tw;
    load32(P.text, aw + delta, tw, -(int64_t)(TEXT_PAD / 16));
    uint64_t t[4];
    win32_words(tw, t);
    uint64_t x0 = ((uint64_t)r0.x | ((uint64_t)r0.y << 32)) ^ t[0], x1 = ((uint64_t)r0.z | ((uint64_t)r0.w << 32)) ^ t[1];
    uint64_t x2 = ((uint64_t)r1.x | ((uint64_t)r1.y << 32)) ^ t[2], x3 = ((uint64_t)r1.z | ((uint64_t)r1.w << 32)) ^ t[3];
    const int sk = lane == 0 ? skip : 0;
    if (fwd) {             // walking order = ascending bytes; bytes 0 .. sk-1 are behind the walk
      if (sk >= 8) { x0 = 0; x1 &= ~0ull << (8 * (sk - 8)); } else { x0 &= ~0ull << (8 * sk); }
      mi = x0 ? (__ffsll((long long)x0) - 1) >> 3
         : x1 ? 8 + ((__ffsll((long long)x1) - 1) >> 3)
         : x2 ? 16 + ((__ffsll((long long)x2) - 1) >> 3)
         : x3 ? 24 + ((__ffsll((long long)x3) - 1) >> 3) : 32;
    } else {               // walking order = descending bytes; the top sk bytes are behind the walk
      if (sk >= 8) { x3 = 0; x2 &= ~0ull >> (8 * (sk - 8)); } else { x3 &= ~0ull >> (8 * sk); }
      mi = x3 ? __clzll((long long)x3) >> 3
         : x2 ? 8 + (__clzll((long long)x2) >> 3)
         : x1 ? 16 + (__clzll((long long)x1) >> 3)
         : x0 ? 24 + (__clzll((long long)x0) >> 3) : 32;
    }
  }
  const int got = min(mi, nvalid);                    // window bytes consumed before a mismatch or the limit
  const unsigned stop = __ballot_sync(0xffffffffu, got < 32);
  const int first = stop ? __ffs((int)stop) - 1 : 32;
  const int got_f = __shfl_sync(0xffffffffu, got, first & 31);
  const int mi_f = __shfl_sync(0xffffffffu, mi, first & 31);
  const int nv_f = __shfl_sync(0xffffffffu, nvalid, first & 31);
  if (first == 32) { *mismatch = false; return 1024 - skip; }
  *mismatch = mi_f < nv_f;
  return 32 * first + got_f - skip;
}

// A parked walk: everything k_sfs_search_mop keeps per thread between two iterations.
struct Cont {
  int64_t roff, delta;
  uint64_t k, s;
  uint32_t ridx, hist, kcode;
  int len, pos, begin, chain_qs, chain_end, hv;
  unsigned n_ext, n_blk, n_txt;
  uint8_t phase, st, tmode, spr;
};

// ---- tail sprint ---------------------------------------------------------------------------------
// One read's walk is a serial chain: restart s -> (b(s), e(b)) -> next restart e - 1.  Inside novel
// sequence every link costs ~7 dependent memory round trips and moves the restart by about one base,
// and in a warp whose 32 lanes all carry reads a lane gets one micro-op per ~5 us iteration: a read
// with a kilobase of clipped or inserted sequence takes tens of milliseconds, which is what is left
// running when the work queue is empty.  The tail kernel therefore gives a whole warp to one parked
// walk: lane j computes the link of restart s - j on its own (same table jump + rank steps, straight
// from global memory), then the chain is resolved over the 32 precomputed links with shuffles.  Links
// are pure functions of the restart position, so the SFSs and the extension count are those of the
// serial walk; a link that runs long (the restart sits in matching sequence) is abandoned and the
// owner carries on from there by itself.
struct Link { int b, e, cnt, blk, status; };   // status 0: SFS [b, e]; 1: read start reached still matching; 2: abandoned
// All 32 lanes walk their links in lockstep: per round every lane names the index blocks of its next
// rank extension, the warp stages them together (cpa_fetch: coalesced 128-byte lines, as in the rank
// walk kernel) and each lane finishes from shared memory.  Per-lane global loads of the blocks made
// the first version of the sprint L1-request bound (profiles/r01i_sfs_tail_src.txt).
__device__ __forceinline__ Link sprint_links(const SearchParams& P, int64_t roff, int len, int s_lane, int K, int budget,
                                             int lane, const uint4* my, uint32_t warp_stage_s) {
  Link r;
  r.b = r.e = r.cnt = r.blk = 0; r.status = 2;
  const uint8_t* rd = P.seq + roff;
  uint64_t k = 0, sz = 0;
  int pos = s_lane;
  int phs = (s_lane < 0 || s_lane >= len) ? 4 : 0;   // 0/2: start backward/forward, 1/3: walking backward/forward, 4: done
  auto base = [&](int i) { return (int)__ldcg(rd + i); };
  auto start = [&](bool fwd) {   // rb3_fmd_set_intv at `pos` (+ K-1 extensions through the jump table); false: N in the way
    if (fwd ? pos + K <= len : pos + 1 >= K) {
      Win16 w;   // the K bases in one round trip (three aligned 8-byte loads)
      load16(P.seq, roff + (fwd ? pos : pos - K + 1), w, 0);
      uint32_t code;
      if (!window_kmer(w, K, code)) return false;
      const uint64_t e = __ldg(P.kmt + (fwd ? kmer_rc(code, K) : code));
      const uint64_t z = e >> 40;
      if (z != 0 && z != KMT_SAT) { k = e & ((1ull << 40) - 1ull); sz = z; pos += fwd ? K - 1 : -(K - 1); r.cnt += K - 1; return true; }
    }
    const int c = base(pos);
    if (c < 1 || c > 4) return false;
    const int cc = fwd ? 5 - c : c;
    k = (uint64_t)P.acc[cc]; sz = (uint64_t)(P.acc[cc + 1] - P.acc[cc]);
    return true;
  };
  for (;;) {
    if (phs == 0 || phs == 2) phs = start(phs == 2) ? phs + 1 : 4;
    bool ext = false;
    int c = 0;
    if (phs == 1) {                               // backward (ping_pong.cpp:15-25)
      if (sz != 0 && pos > 0) {
        c = --budget < 0 ? 0 : base(--pos);
        if (c < 1 || c > 4) phs = 4; else ext = true;
      } else if (sz != 0) { r.status = 1; phs = 4; }
      else { r.b = pos; phs = 2; }
    } else if (phs == 3) {                        // forward from begin (ping_pong.cpp:27-41)
      if (sz != 0 && pos + 1 < len) {
        c = --budget < 0 ? 0 : base(++pos);
        if (c < 1 || c > 4) phs = 4; else { c = 5 - c; ext = true; }
      } else if (sz != 0) { phs = 4; }            // cannot happen (see k_sfs_search); leave it to the owner
      else { r.e = pos; r.status = 0; phs = 4; }
    }
    if (!__any_sync(0xffffffffu, phs != 4)) break;
    const uint32_t bk = ext ? (uint32_t)(k >> 8) : NOBLK, bl = ext ? (uint32_t)((k + sz) >> 8) : NOBLK;
    cpa_fetch(P, bk, bl, warp_stage_s, lane);
    if (ext) {
      const bool two = bl != bk;
      tma_consume(P, my, two, c, k, sz, lane);
      ++r.cnt;
      r.blk += two ? 2 : 1;
    }
    __syncwarp();
  }
  return r;
}
constexpr int TAIL_OWNERS = 1;        // lanes of a tail-kernel warp that own a parked walk (the rest only help)
constexpr int SPRINT_BUDGET = 16;     // rank steps a helper spends on one link before giving it up

// ------------------------------------------------------------------------------ v3 kernel
// Micro-op pipeline.  ncu on the hybrid version of k_sfs_search_tma (profiles/r01e_*) showed a warp
// iteration lasting ~12 us at 31 % issue utilisation: every divergent path of the state machine
// (K-mer window load -> table lookup, SA sample, read window refill, text compare, index blocks) did
// its own blocking round trip to memory, one after the other.  Here a thread's walk is cut into
// micro-ops that each need ONE batch of loads whose addresses are known up front; per iteration
// every lane (1) advances its state machine without touching memory, (2) issues the loads of its
// micro-op from code all lanes share, (3) waits once together with the warp's cp.async staging,
// (4) consumes.  A warp iteration is then one memory round trip whatever mix of ops its lanes hold.
//   OP_EXT   rank extension: 1-2 index blocks staged by the warp (cp.async, as in k_sfs_search_tma)
//   OP_TXT   located match: 32 read bytes against 32 text bytes
//   OP_KMER  restart, step 1: the K read bytes at the pivot (skipped when the bases just walked are
//            still in the thread's 16-base history register, which is the rule inside novel sequence)
//   OP_KMT   restart, step 2: the jump-table entry of that K-mer
//   OP_SSA   locate: the SA sample of the row (forward phase: + contig lookup for the mirror image)
enum : int { OP_NONE = 0, OP_EXT = 1, OP_TXT = 2, OP_KMER = 3, OP_KMT = 4, OP_SSA = 5, OP_SPRINT = 6 };
enum : int { ST_START = 0, ST_WALK = 1, ST_KMT = 2 };

// TAIL = false: the main kernel, thread per read; parks unfinished walks once the queue is dry (if P.cont).
// TAIL = true: the tail kernel: TAIL_OWNERS lanes of a warp own one parked walk each, the other lanes
// only serve the cooperative steps (located-match compare, sprint, block staging) -- few walks per warp
// keep an iteration short, which is what a long serial walk needs.
template <int MINB, bool TAIL, bool PK2 = false>   // PK2: the unpacking CTAs speak the 2-bit transport (a separate instantiation)
__global__ void __launch_bounds__(TMA_WARPS * 32, MINB) k_sfs_search_mop(const SearchParams P) {
  __shared__ __align__(128) uint4 stage[TMA_WARPS * 32 * 16];  // [warp][lane][2 blocks][8 slices]
  __shared__ uint8_t pend_all[TMA_WARPS * 32];                  // per warp: lanes with an extension pending
  int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  uint4* my = stage + (warp * 32 + lane) * 16;
  uint32_t warp_stage_s = smem_u32(stage + warp * 32 * 16);
  // pin these in registers: under the register cap the compiler otherwise re-derives them (S2R + cvta)
  // inside the hot loops -- 31 % of the tail kernel's instructions in profiles/r01k_sfs_tail_lines.txt
  if (TAIL) asm volatile("" : "+r"(lane), "+r"(warp_stage_s), "+l"(my));
  if (P.n_unpack > 0 && (int)blockIdx.x < P.n_unpack) {   // packed streamed batch: this CTA feeds the others
    if (PK2) unpack_cta_loop2(P); else unpack_cta_loop(P);
    return;
  }
  if (!TAIL && threadIdx.x == 0) atomicMin(P.stats + 4, globaltimer_ns());
  const unsigned long long t_begin = TAIL ? globaltimer_ns() : 0ull;

  bool alive = !TAIL || (lane & (32 / TAIL_OWNERS - 1)) == 0, have = false;
  bool spr = false;   // the pending backward restart follows an SFS found by rank steps alone (novel sequence)
  unsigned it = 0;
  int st = ST_START, phase = 0;
  uint32_t ridx = 0, kcode = 0;
  int64_t roff = 0;
  int len = 0, pos = 0, begin = 0;
  uint64_t k = 0, s = 0;
  int chain_qs = -1, chain_end = -1;
  unsigned n_ext = 0, n_blk = 0, n_txt = 0;
  bool tmode = false;
  int64_t delta = 0;
  int nr_lo = 0, nr_hi = -1;     // read positions known to be one run of N (closed form of N-run restarts)
  bool nr_closed = false;
  // the last bases walked, 2 bits each, in read order: a backward walk keeps the base at `pos` in the
  // low bits (higher positions above it), a forward walk keeps the base at `pos` in the top bits;
  // hv = how many of them are valid (0 after an N, a text-mode step or a fresh read)
  uint32_t hist = 0;
  int hv = 0;
  const unsigned ss_mask = (1u << P.ss_log) - 1u;
  const int K = P.kmer_k;
  ReadWin1 win;
  win.reset();

  auto emit = [&](int qs, int ln) {
    unsigned long long o = atomicAdd(P.out_count, 1ull);
    if (o < P.out_cap) {
      uint32_t sk = P.assemble ? (uint32_t)qs : ~(uint32_t)qs;
      P.out_key[o] = ((uint64_t)ridx << 32) | sk;
      P.out_len[o] = (uint32_t)ln;
    }
  };
  auto on_sfs = [&](int qs, int ln) {
    if (!P.assemble) { emit(qs, ln); return; }
    if (chain_qs >= 0 && qs + ln > chain_qs) {
      chain_qs = qs;
    } else {
      if (chain_qs >= 0) emit(chain_qs, chain_end - chain_qs);
      chain_qs = qs;
      chain_end = qs + ln;
    }
  };
  auto finish_read = [&]() {
    if (P.assemble && chain_qs >= 0) emit(chain_qs, chain_end - chain_qs);
    chain_qs = -1;
    have = false;
    tmode = false;
    atomicAdd(P.stats + 0, (unsigned long long)n_ext + n_txt);
    atomicAdd(P.stats + 1, (unsigned long long)n_blk);
    if (n_txt) atomicAdd(P.stats + 2, (unsigned long long)n_txt);
    n_ext = 0; n_blk = 0; n_txt = 0;
  };
  auto set_intv = [&](int c0) {
    k = (uint64_t)P.acc[c0];
    s = (uint64_t)(P.acc[c0 + 1] - P.acc[c0]);  // rb3_fmd_set_intv (ping_pong.cpp:12,30)
    tmode = false;
    st = ST_WALK;
  };
  auto push_hist = [&](int ch) {   // ch = read base (nt6 code) at the new `pos`
    if (ch >= 1 && ch <= 4) {
      hist = phase ? (hist >> 2) | ((uint32_t)(ch - 1) << 30) : (hist << 2) | (uint32_t)(ch - 1);
      hv = min(16, hv + 1);
    } else {
      hv = 0;
    }
  };
  auto set_intv_at_pivot = [&]() {
    const int ch = win.get(P.seq, roff + pos, phase ? 1 : -1);
    set_intv(phase ? comp6(ch) : ch);
    hv = 0;
    push_hist(ch);
  };

  while (__any_sync(0xffffffffu, alive)) {
    // ---- (1) advance to the next micro-op (ping_pong.cpp:15-47); no waits on memory except the
    //          hand-out of a new read and the rare walk that starts without the jump table
    int op = OP_NONE, c = 0;
    if (!TAIL && P.cont != nullptr && (++it & 7u) == 0u && *reinterpret_cast<volatile unsigned int*>(P.dry) != 0u &&
        (P.last_ready == nullptr || *reinterpret_cast<const volatile unsigned int*>(P.last_ready) != 0u)) {
      // the queue is dry: park the walk for the tail kernel instead of finishing it at one micro-op
      // per (slow) iteration of a warp that is emptying
      if (alive && have) {
        Cont ct;
        ct.roff = roff; ct.delta = delta; ct.k = k; ct.s = s; ct.ridx = ridx; ct.hist = hist; ct.kcode = kcode;
        ct.len = len; ct.pos = pos; ct.begin = begin; ct.chain_qs = chain_qs; ct.chain_end = chain_end; ct.hv = hv;
        ct.n_ext = n_ext; ct.n_blk = n_blk; ct.n_txt = n_txt;
        ct.phase = (uint8_t)phase; ct.st = (uint8_t)st; ct.tmode = tmode ? 1 : 0; ct.spr = spr ? 1 : 0;
        P.cont[atomicAdd(P.n_cont, 1ull)] = ct;
        have = false;
      }
      alive = false;
    }
    while (alive && op == OP_NONE) {
      if (TAIL && !have) {
        const unsigned long long w = atomicAdd(P.work, 1ull);
        if (w >= *P.n_cont) {
          if (P.stats_on) atomicAdd(P.stats + 7, globaltimer_ns() - t_begin);   // this warp's busy time
          alive = false;
          break;
        }
        const Cont ct = P.cont[w];
        roff = ct.roff; delta = ct.delta; k = ct.k; s = ct.s; ridx = ct.ridx; hist = ct.hist; kcode = ct.kcode;
        len = ct.len; pos = ct.pos; begin = ct.begin; chain_qs = ct.chain_qs; chain_end = ct.chain_end; hv = ct.hv;
        n_ext = ct.n_ext; n_blk = ct.n_blk; n_txt = ct.n_txt;
        phase = ct.phase; st = ct.st; tmode = ct.tmode != 0; spr = ct.spr != 0;
        have = true;
        win.reset();
        nr_hi = -1;
        continue;
      }
      if (!have) {
        const unsigned long long w = atomicAdd(P.work, 1ull);
        if (w >= (unsigned long long)P.n_reads) {
          if (w == (unsigned long long)P.n_reads) atomicMin(P.stats + 5, globaltimer_ns());   // the queue just ran dry
          if (P.dry) *reinterpret_cast<volatile unsigned int*>(P.dry) = 1u;
          alive = false;
          break;
        }
        if (P.sched) {
          const ulonglong2 e = __ldg(P.sched + w);
          roff = (int64_t)e.x; ridx = (uint32_t)e.y; len = (int)(e.y >> 32);
        } else {
          ridx = P.order ? P.order[w] : (uint32_t)w;
          roff = P.offs[ridx];
          len = (int)(P.offs[ridx + 1] - roff);
        }
        if (len <= 0) continue;
        if (P.ready) {  // streamed batch: wait until the chunk holding the last base landed
          if (!wait_chunk(P.ready + (roff + len - 1) / P.chunk_bytes, P.stats + 3)) { alive = false; break; }
          __threadfence();
        }
        have = true;
        phase = 0;
        pos = len - 1;
        st = ST_START;
        tmode = false;
        hv = 0;
        spr = false;
        win.reset();
        nr_hi = -1;
      }
      if (st == ST_START) {   // (re)start at pivot `pos`: jump table if K bases are there, else one base
        if (!phase && P.max_nrun >= 0 && P.overlap == -1 && pos >= P.max_nrun && P.seq[roff + pos] == 5) {
          // A backward restart inside a run of N.  With L = the longest run of N in the text, N^(L+1) occurs nowhere: if
          // P[pos-L .. pos] is all N the backward walk fails after exactly L extensions at begin = pos - L, the forward
          // walk from there fails after L more at end = pos, the SFS is (pos - L, L + 1) and the next restart is pos - 1
          // (ping_pong.cpp:12-47) -- no index access at all.  Without this a read that is one long run of N costs
          // 2 L serial rank extensions per base on one lane.  [nr_lo, nr_hi] = bases already known to be N.
          const int L = P.max_nrun;
          if (!(nr_lo <= pos && pos <= nr_hi)) { nr_hi = pos; nr_lo = pos; nr_closed = pos == 0; }
          if (nr_lo > pos - L) {   // not known yet: find the start of the run (once per run: later restarts inside it reuse [nr_lo, nr_hi])
            while (!nr_closed) {
              if (P.seq[roff + nr_lo - 1] == 5) { --nr_lo; if (nr_lo == 0) nr_closed = true; } else nr_closed = true;
            }
          }
          if (nr_lo <= pos - L) {
            // restarts pos, pos - 1, .., nr_lo + L all have the closed form: n SFSs (p - L, L + 1), each overlapping the next
            const int n = pos - L - nr_lo + 1;
            n_ext += 2u * (unsigned)L * (unsigned)n;
            if (P.assemble) {
              on_sfs(pos - L, L + 1);
              chain_qs = nr_lo;                       // what on_sfs would leave after the other n - 1 (each starts one base lower)
            } else {
              const unsigned long long o0 = atomicAdd(P.out_count, (unsigned long long)n);
              for (int i = 0; i < n; ++i) {
                const unsigned long long o = o0 + (unsigned long long)i;
                if (o < P.out_cap) { P.out_key[o] = ((uint64_t)ridx << 32) | (uint32_t)~(uint32_t)(pos - L - i); P.out_len[o] = (uint32_t)(L + 1); }
              }
            }
            hv = 0; tmode = false; spr = false;
            if (nr_lo == 0) finish_read(); else pos = nr_lo + L - 1;
            continue;
          }
        }
        if (TAIL && !phase && spr && P.kmt != nullptr && P.overlap == -1) { op = OP_SPRINT; continue; }
        if (P.kmt != nullptr && (phase ? pos + K <= len : pos + 1 >= K)) {
          // hv > 0 here means the history was left by the walk that ended at the pivot's neighbour:
          //   forward restart at begin (= the base the backward walk failed on): low bits = P[begin..]
          //   backward restart at the base below the forward mismatch: top bits = P[..pos+1]
          if (phase && hv >= K) {
            kcode = kmer_rc(hist & ((1u << (2 * K)) - 1u), K);
            st = ST_KMT;
          } else if (!phase && hv >= K + 1 && P.overlap == -1) {
            kcode = (hist >> (30 - 2 * K)) & ((1u << (2 * K)) - 1u);
            st = ST_KMT;
          } else {
            op = OP_KMER;
          }
        } else {
          set_intv_at_pivot();
        }
        continue;
      }
      if (st == ST_KMT) { op = OP_KMT; continue; }
      const int dir = phase ? 1 : -1;
      const bool room = phase ? (pos + 1 < len) : (pos > 0);
      if (s != 0 && room) {
        if (tmode) {
          op = OP_TXT;
        } else if (P.text != nullptr && s == 1 && ((unsigned)k & ss_mask) == 0u) {
          op = OP_SSA;
        } else {
          pos += dir;
          const int ch = win.get(P.seq, roff + pos, dir);
          push_hist(ch);
          c = phase ? comp6(ch) : ch;
          op = OP_EXT;
        }
      } else if (phase == 0) {
        if (s != 0) {
          finish_read();  // reached the read start still matching (ping_pong.cpp:24-25)
        } else {          // mismatch at pos: forward search from here (ping_pong.cpp:27-30)
          begin = pos;
          phase = 1;
          st = ST_START;
        }
      } else {
        if (s != 0) ++pos;               // defensive: cannot happen (see k_sfs_search)
        on_sfs(begin, pos - begin + 1);  // ping_pong.cpp:39-41
        if (begin == 0) {
          finish_read();
        } else {
          int nb = (P.overlap == 0) ? begin - 1 : pos + P.overlap;  // ping_pong.cpp:44-47
          if (nb < 0) {
            finish_read();
          } else {
            if (nb > len - 1) nb = len - 1;
            pos = nb;
            phase = 0;
            st = ST_START;
            spr = hv >= K + 1;
          }
        }
      }
    }
    // ---- tail sprint: the links of 32 consecutive restarts in one go
    if (TAIL) {
      for (unsigned smask = __ballot_sync(0xffffffffu, op == OP_SPRINT); smask; smask &= smask - 1) {
        const int src = __ffs((int)smask) - 1;
        const int64_t ro = __shfl_sync(0xffffffffu, roff, src);
        const int ln = __shfl_sync(0xffffffffu, len, src);
        const int s0 = __shfl_sync(0xffffffffu, pos, src);
        const Link lk = sprint_links(P, ro, ln, s0 - lane, K, P.sprint_budget, lane, my, warp_stage_s);
        unsigned blk_used = 0;   // index blocks of the links the chain actually takes (the others are waste, not work)
        int cur = s0, n_links = 0;
        bool done = false, gave_up = false;
        while (s0 - cur < 32 && cur >= 0) {       // warp-uniform: every lane follows the same chain
          const int j = s0 - cur;
          const int st_j = __shfl_sync(0xffffffffu, lk.status, j);
          if (st_j == 2) { gave_up = true; break; }   // abandoned link: the owner walks on from `cur`
          ++n_links;
          const int b_j = __shfl_sync(0xffffffffu, lk.b, j), e_j = __shfl_sync(0xffffffffu, lk.e, j);
          const int cnt_j = __shfl_sync(0xffffffffu, lk.cnt, j);
          blk_used += (unsigned)__shfl_sync(0xffffffffu, lk.blk, j);
          if (lane == src) n_ext += (unsigned)cnt_j;
          if (st_j == 1) { done = true; break; }                // reached the read start still matching
          if (lane == src) on_sfs(b_j, e_j - b_j + 1);          // ping_pong.cpp:39-41
          if (b_j == 0 || e_j - 1 < 0) { done = true; break; }  // ping_pong.cpp:42-47 with overlap = -1
          cur = e_j - 1;
        }
        if (lane == src) {
          n_blk += blk_used;
          if (P.stats_on) {
            atomicAdd(P.stats + 11, 1ull); atomicAdd(P.stats + 12, (unsigned long long)n_links);
            if (gave_up) atomicAdd(P.stats + 13, 1ull);
          }
          if (done) finish_read();
          else { pos = cur; phase = 0; st = ST_START; }
          spr = !done && cur < s0;                 // no progress (first link abandoned): walk the next one alone
          hv = 0;                                  // the history register does not follow the sprint
          tmode = false;
          win.id = -1; win.nid = -1;
          op = OP_NONE;
        }
      }
    }
    // ---- (2) issue: every load of this iteration leaves from here
    Win32 rw;
    uint64_t one = 0;
    if (op == OP_TXT) {
      // served below, one lane at a time, by the whole warp
    } else if (op == OP_KMER) {
      load32(P.seq, roff + (phase ? pos : pos - K + 1), rw, 0);
    } else if (op == OP_KMT) {
      one = __ldg(P.kmt + kcode);
    } else if (op == OP_SSA) {
      one = __ldg(P.ssa + (k >> P.ss_log));
    }
    const uint32_t bk = op == OP_EXT ? (uint32_t)(k >> 8) : NOBLK;
    const uint32_t bl = op == OP_EXT ? (uint32_t)((k + s) >> 8) : NOBLK;
    cpa_issue_sparse(P, bk, bl, warp_stage_s, lane, pend_all + warp * 32);
    // located matches: the warp walks up to 1024 bases for each lane that is in that mode, while
    // the index blocks of the other lanes are on their way
    for (unsigned tmask = __ballot_sync(0xffffffffu, op == OP_TXT); tmask; tmask &= tmask - 1) {
      const int src = __ffs((int)tmask) - 1;
      const int64_t g = __shfl_sync(0xffffffffu, roff + pos, src);
      const int64_t dl = __shfl_sync(0xffffffffu, delta, src);
      const int fwd = __shfl_sync(0xffffffffu, phase, src);
      const int lim = __shfl_sync(0xffffffffu, phase ? len - 1 - pos : pos, src);
      bool mism;
      const int adv = coop_text_step(P, g, dl, fwd, lim, lane, &mism);
      if (lane == src) {
        const int dir = phase ? 1 : -1;
        if (!mism) {             // all adv extensions succeed (the interval stays of size 1)
          pos += dir * adv;
          n_txt += (unsigned)adv;
        } else {                 // extension adv+1 fails: the state a rank walk ends in with s == 0
          pos += dir * (adv + 1);
          n_txt += (unsigned)(adv + 1);
          s = 0;
          tmode = false;
        }
        hv = 0;
        win.id = -1; win.nid = -1;
      }
    }
    // ---- (3) one wait for the warp
    cpa_wait();
    // ---- (4) consume
    if (op == OP_EXT) {
      const bool two = bl != bk;
      tma_consume(P, my, two, c, k, s, lane);
      ++n_ext;
      n_blk += two ? 2u : 1u;
    } else if (op == OP_KMER) {
      uint64_t r[4];
      win32_words(rw, r);
      uint32_t code;
      if (words_kmer(r[0], r[1], K, code)) {
        kcode = phase ? kmer_rc(code, K) : code;
        // the K bases become the history of the walk that continues from the jump
        hist = phase ? code << (32 - 2 * K) : code;
        hv = -K;               // negative: valid only if the jump succeeds (flipped in OP_KMT)
        st = ST_KMT;
      } else {                 // an N among the K bases: walk from the pivot base
        set_intv_at_pivot();
      }
    } else if (op == OP_KMT) {
      const uint64_t sz = one >> 40;
      if (sz != 0 && sz != KMT_SAT) {   // the K-mer occurs: rb3_fmd_set_intv + K-1 successful extensions
        k = one & ((1ull << 40) - 1ull);
        s = sz;
        pos += phase ? K - 1 : -(K - 1);
        n_ext += (unsigned)(K - 1);
        tmode = false;
        st = ST_WALK;
        win.id = -1; win.nid = -1;
        if (hv < 0) {
          hv = -hv;                     // history = the K-mer read by OP_KMER
        } else {                        // the K-mer came out of the history register: re-aim it
          const uint32_t code = phase ? kmer_rc(kcode, K) : kcode;   // bases in read order, lowest position first
          hist = phase ? code << (32 - 2 * K) : code;
          hv = K;
        }
      } else {                          // absent (the walk fails within K bases) or saturated entry
        set_intv_at_pivot();
      }
    } else if (op == OP_SSA) {
      // Backward: P[pos..] starts at T[p].  Forward: rc(P[begin..pos]) starts at T[p]; its mirror
      // image on the other strand is P[begin..pos] read left to right.
      const int64_t p = (int64_t)one;
      delta = phase ? mirror_pos(P, p + (pos - begin)) - (roff + begin) : p - (roff + pos);
      tmode = true;
    }
    __syncwarp();  // staging slots are rewritten by other lanes next iteration
  }
  if (lane == 0) atomicMax(P.stats + 6, globaltimer_ns());
}

template <int MODE>
__global__ void __launch_bounds__(TMA_WARPS * 32) k_rank_bench_tma(const SearchParams P, int64_t nq, int64_t n,
                                                                   int64_t delta, uint64_t seed,
                                                                   unsigned long long* __restrict__ sink,
                                                                   unsigned long long* __restrict__ blocks_touched) {
  __shared__ __align__(128) uint4 stage[TMA_WARPS * 32 * 16];
  __shared__ __align__(8) uint64_t bars[TMA_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t bar = smem_u32(&bars[warp]);
  uint4* my = stage + (warp * 32 + lane) * 16;
  const uint32_t my_s = smem_u32(my);
  if (lane == 0) {
    mbar_init(bar, 32);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  uint32_t parity = 0;
  const int64_t nthreads = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t iters = (nq + nthreads - 1) / nthreads;
  unsigned long long acc = 0;
  unsigned nb = 0;
  for (int64_t it = 0; it < iters; ++it) {
    const int64_t q = t0 + it * nthreads;
    const bool act = q < nq;
    uint64_t r = splitmix64(seed + (uint64_t)q);
    uint64_t k = __umul64hi(r, (uint64_t)(n - delta));
    uint64_t s = (uint64_t)delta;
    int c = 1 + (int)((r >> 60) & 3);
    bool two = false;
    if (MODE == 0) {
      if (act) two = tma_issue(P, k, s, my_s, bar);
      else mbar_arrive(bar);
      while (!mbar_try_wait(bar, parity)) {}
      parity ^= 1u;
    } else {
      const uint32_t bk = act ? (uint32_t)(k >> 8) : NOBLK;
      const uint32_t bl = act ? (uint32_t)((k + s) >> 8) : NOBLK;
      two = bl != bk;
      cpa_fetch(P, bk, bl, smem_u32(stage + warp * 32 * 16), lane);
    }
    if (act) {
      tma_consume(P, my, two, c, k, s, MODE == 1 ? lane : 0);
      nb += two ? 2u : 1u;
      acc += k + s;
    }
    if (MODE == 1) __syncwarp();
  }
  atomicAdd(blocks_touched, (unsigned long long)nb);
  if (acc == 0x123456789ULL) atomicAdd(sink, acc);
}

// ------------------------------------------------------------------------------ host side
static void fill_params(SearchParams& P, const IndexDev& d) {
  memset(&P, 0, sizeof(P));
  P.blocks = d.d_blocks;
  P.cntN = d.d_cntN;
  P.sbase = d.d_sbase;
  memcpy(P.acc, d.acc, sizeof(d.acc));
  const char* e = getenv("SVB_SEARCH_TEXT");   // SVB_SEARCH_TEXT=0: rank walk only (the pure FMD kernel)
  if (d.d_text && !(e && *e == '0')) {
    P.text = d.d_text; P.ssa = d.d_ssa; P.tstart = d.d_tstart; P.n_contigs = d.n_contigs; P.ss_log = d.ss_log;
  }
  // a property of the indexed text, whichever way the extensions are answered; -1 (an index built from a bare BWT) switches the closed form off
  P.max_nrun = (d.max_nrun >= 0 && d.max_nrun < 0x3fffffff && !getenv("SVB_SEARCH_NO_NRUN")) ? (int)d.max_nrun : -1;
  P.stats_on = getenv("SVB_SEARCH_STATS") != nullptr;
  P.sprint_budget = SPRINT_BUDGET;
  if (const char* eb = getenv("SVB_SPRINT_BUDGET")) P.sprint_budget = atoi(eb);
  const char* ej = getenv("SVB_SEARCH_JUMP");   // SVB_SEARCH_JUMP=0: restarts walk from one base
  if (d.d_kmt && !(ej && *ej == '0')) { P.kmt = d.d_kmt; P.kmer_k = d.kmer_k; }
}

// sort key of a read: (chunk of its last base) << 32 | ~length  -> chunk-major, longest first
__global__ void k_read_keys(const int64_t* __restrict__ offs, int64_t n, int64_t chunk_bytes,
                            uint64_t* __restrict__ keys, uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t b = offs[i], l = offs[i + 1] - b;
  const uint64_t chunk = (chunk_bytes > 0 && l > 0) ? (uint64_t)((b + l - 1) / chunk_bytes) : 0ull;
  keys[i] = (chunk << 32) | (uint32_t)~(uint32_t)(l > 0xffffffffLL ? 0xffffffffLL : l);
  vals[i] = (uint32_t)i;
}

__global__ void k_make_sched(const uint32_t* __restrict__ order, const int64_t* __restrict__ offs, int64_t n,
                             ulonglong2* __restrict__ sched) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint32_t r = order[i];
  const int64_t b = offs[r], l = offs[r + 1] - b;
  sched[i] = make_ulonglong2((unsigned long long)b, ((unsigned long long)(l > 0x7fffffffLL ? 0x7fffffffLL : l) << 32) | r);
}

// read indices in hand-out order: longest first (within a chunk when the batch is streamed)
static int make_order(svb_reads* R, int64_t chunk_bytes, cudaStream_t st) {
  int64_t n = R->n_reads;
  if (n == 0) return SVB_OK;
  uint64_t *k1 = nullptr, *k2 = nullptr;
  uint32_t* v1 = nullptr;
  SVB_CUDA(pmalloc((void**)&k1, n * 8, st));
  SVB_CUDA(pmalloc((void**)&k2, n * 8, st));
  SVB_CUDA(pmalloc((void**)&v1, n * 4, st));
  SVB_CUDA(pmalloc((void**)&R->d_order, n * 4, st));
  k_read_keys<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R->d_offs, n, chunk_bytes, k1, v1);
  size_t bytes = 0;
  void* tmp = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k1, k2, v1, R->d_order, n, 0, 64, st);
  SVB_CUDA(pmalloc(&tmp, bytes, st));
  SVB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, k1, k2, v1, R->d_order, n, 0, 64, st));
  SVB_CUDA(pmalloc((void**)&R->d_sched, n * 16, st));
  k_make_sched<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R->d_order, R->d_offs, n, R->d_sched);
  SVB_CUDA(cudaGetLastError());
  pfree(tmp, st); pfree(k1, st); pfree(k2, st); pfree(v1, st);
  SVB_CUDA(cudaStreamSynchronize(st));
  return SVB_OK;
}

int check_device(int device);

}  // namespace svb

using namespace svb;

namespace svb {
// K = floor(log4(n)) - 1 keeps the expected number of occurrences of a random K-mer in [4, 16): the
// jump almost always lands on a non-empty interval a few extensions away from a unique match.
// SVB_KMER_K overrides (0 disables the table).  8 bytes per entry: K = 15 is 8.6 GB next to a
// 6.2 G-symbol index.
int build_kmer_table(IndexDev* idx) {
  if (idx->G != 8 || idx->d_kmt) return SVB_OK;
  int K = 0;
  while (K < 16 && (idx->n >> (2 * (K + 1))) > 0) ++K;   // floor(log4 n)
  K = std::min(15, K - 1);
  if (const char* e = getenv("SVB_KMER_K")) K = std::min(15, atoi(e));
  if (K < 2) return SVB_OK;
  SearchParams P;
  fill_params(P, *idx);
  const size_t entries = (size_t)1 << (2 * K);
  ulonglong2 *a = nullptr, *b = nullptr;
  uint64_t* kmt = nullptr;
  cudaError_t e1 = cudaMalloc((void**)&kmt, entries * 8);
  cudaError_t e2 = K > 1 ? cudaMalloc((void**)&a, (entries / 4) * 16) : cudaSuccess;
  cudaError_t e3 = K > 2 ? cudaMalloc((void**)&b, (entries / 16) * 16) : cudaSuccess;
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    cudaFree(kmt); cudaFree(a); cudaFree(b);
    cudaGetLastError();
    return SVB_OK;   // not enough memory for the table: restarts simply walk from one base
  }
  // levels alternate between the two scratch buffers so that level K-1 ends in `a` (the larger one)
  ulonglong2* bufs[2] = {a, b};
  for (int j = 1; j <= K; ++j) {
    const uint64_t total = 1ull << (2 * j);
    const unsigned grid = (unsigned)std::min<uint64_t>((total + 255) / 256, 148 * 32);
    const ulonglong2* prev = j > 1 ? bufs[(K - j) & 1] : nullptr;      // level j-1 lives in bufs[(K-1-(j-1)) & 1]
    ulonglong2* cur = j < K ? bufs[(K - 1 - j) & 1] : nullptr;
    k_kmer_level<<<grid, 256>>>(P, prev, cur, j == K ? kmt : nullptr, j, idx->n);
  }
  cudaError_t e = cudaDeviceSynchronize();
  cudaFree(a); cudaFree(b);
  if (e != cudaSuccess) { cudaFree(kmt); set_error("k_kmer_level failed: %s", cudaGetErrorString(e)); return SVB_ECUDA; }
  idx->d_kmt = kmt;
  idx->kmer_k = K;
  return SVB_OK;
}
}  // namespace svb

namespace svb {
int check_device(int device) {
  int cnt = 0;
  cudaError_t e = cudaGetDeviceCount(&cnt);
  if (e != cudaSuccess || cnt == 0) {
    set_error("no CUDA device available (%s); libsvdss_b200 has no CPU fallback",
              e == cudaSuccess ? "device count is 0" : cudaGetErrorString(e));
    return SVB_ECUDA;
  }
  if (device < 0 || device >= cnt) { set_error("device %d out of range (%d present)", device, cnt); return SVB_EINVAL; }
  SVB_CUDA(cudaSetDevice(device));
  return SVB_OK;
}

// occupancy-sized persistent grid
template <typename K>
static int persistent_grid(K kernel, int threads, int device, int* grid) {
  int per_sm = 0, sms = 0;
  SVB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, 0));
  SVB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  if (per_sm < 1) per_sm = 1;
  *grid = per_sm * sms;
  return SVB_OK;
}

// ---- lane-group configurations: G lanes x SPL slices per lane cover one block (G*SPL = 4 or 8
// slices).  Fewer lanes per read = more independent chains per warp = more loads in flight.
// Selected by SVB_SEARCH_CFG ("8x1", "4x2", "2x4", "1x8" for 128-byte blocks; "4x1", "2x2", "1x4"
// for 64-byte blocks; "cpa" / "tma" = the thread-per-read staged kernels); the default is the
// fastest measured on B200.
static int pick_cfg(int slices, int* G) {
  const char* e = getenv("SVB_SEARCH_CFG");
  // defaults: 128-byte blocks -> thread-per-read kernel with cooperative cp.async staging ("cpa");
  //           64-byte blocks  -> 4 lanes per read
  // ("mop" = micro-op pipeline kernel, the default for 128-byte blocks; "cpa" = its predecessor)
  int g = (slices == 8) ? -2 : 4;
  if (e && (strcmp(e, "tma") == 0 || strcmp(e, "cpa") == 0 || strcmp(e, "mop") == 0)) {  // staged kernels: 128-byte blocks only
    *G = (slices == 8) ? (e[0] == 't' ? 0 : e[0] == 'c' ? -1 : -2) : g;
    return SVB_OK;
  }
  if (e && *e) {
    int a = 0, b = 0;
    if (sscanf(e, "%dx%d", &a, &b) == 2 && a * b == slices && (a == 1 || a == 2 || a == 4 || a == 8)) g = a;
    else if (sscanf(e, "%dx%d", &a, &b) == 2 && (a * b == 4 || a * b == 8)) { /* other block size: keep default */ }
    else { set_error("SVB_SEARCH_CFG=%s is not one of tma, cpa, GxS", e); return SVB_EINVAL; }
  }
  *G = g;
  return SVB_OK;
}

#define SVB_DISPATCH_CFG(slices, G, CALL)                                         \
  do {                                                                            \
    if ((slices) == 8) {                                                          \
      if ((G) == 8) { CALL(8, 1, 5); } else if ((G) == 4) { CALL(4, 2, 5); }      \
      else if ((G) == 2) { CALL(2, 4, 3); } else { CALL(1, 8, 2); }               \
    } else {                                                                      \
      if ((G) == 4) { CALL(4, 1, 5); } else if ((G) == 2) { CALL(2, 2, 4); }      \
      else { CALL(1, 4, 3); }                                                     \
    }                                                                             \
  } while (0)

// a pair of timing events that early returns cannot leak
struct EventPair {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaError_t create() { cudaError_t e = cudaEventCreate(&a); return e != cudaSuccess ? e : cudaEventCreate(&b); }
  ~EventPair() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
};

struct SearchScratch {
  unsigned long long* d_ctr = nullptr;  // [0] work [1] out_count [2] ext [3] blocks [4] text extensions [5] stream stalled
  uint64_t* d_key = nullptr; uint64_t* d_key2 = nullptr;
  uint32_t* d_len = nullptr; uint32_t* d_len2 = nullptr;
  void* d_tmp = nullptr;
  void* d_cont = nullptr;   // parked walks between the main and the tail kernel
  cudaStream_t st = nullptr;
  ~SearchScratch() {
    pfree(d_cont, st);
    pfree(d_ctr, st); pfree(d_key, st); pfree(d_key2, st); pfree(d_len, st); pfree(d_len2, st); pfree(d_tmp, st);
  }
};

// runs the search kernel (+ sort) on reads resident on the device; fills `out` host arrays
// host source of a streamed batch: bytes are copied chunk by chunk on their own stream while the
// search kernel (already launched) consumes them
struct StreamSrc {
  const uint8_t* host = nullptr;
  int64_t total = 0, chunk_bytes = 0, n_chunks = 0;
  unsigned int* d_ready = nullptr;
  unsigned int* h_one = nullptr;  // pinned word holding 1
  cudaStream_t copy_stream = nullptr;
  // BAM-native input (svb_sfs_batch_bam4): `host` holds 4-bit packed reads; a chunk is still a range
  // of UNPACKED base positions, fed by copying the packed bytes that cover it and unpacking them on
  // the device (k_unpack4) before its flag goes up
  bool packed = false;
  uint8_t* d_seq4 = nullptr;            // device copy of the packed bytes
  const int64_t* d_seq4_offs = nullptr; // n_reads + 1 byte offsets (device)
  const int64_t* h_offs = nullptr;      // n_reads + 1 base offsets (host), for the chunk -> read ranges
  const int64_t* h_seq4_offs = nullptr; // n_reads + 1 byte offsets (host)
  const int64_t* h_chunk_r = nullptr;   // [2][n_chunks] (host)
  int64_t n_reads = 0;
  unsigned int* d_arrived = nullptr;    // per chunk: packed bytes delivered (copy engine)
  unsigned int* d_done = nullptr;       // per chunk: unpacking CTAs finished
  int64_t* d_chunk_r = nullptr;         // [2][n_chunks] first / last read of each chunk
  int n_unpack = 0;
  // 2-bit transport (SVB_STREAM_PACK2=1): every chunk is re-packed on the host (svb_pack2_chunk) into one of two
  // pinned staging buffers while the previous chunk is on the wire
  bool pack2 = false;
  uint8_t* d_pk2 = nullptr;             // device: 2-bit bytes of the whole batch
  int64_t* d_pk2_offs = nullptr;
  const int64_t* h_pk2_offs = nullptr;  // n_reads + 1 (host)
  unsigned int* d_mode = nullptr;       // per chunk: 2 = 2-bit bytes arrived, 1 = 4-bit bytes arrived
  int64_t* d_exc_pos = nullptr;         // [n_chunks][exc_cap]
  unsigned int* d_exc_n = nullptr;
  int exc_cap = 0;
  uint8_t* h_stage[2] = {nullptr, nullptr};   // pinned
  int64_t stage_cap = 0;
  int64_t* h_exc[2] = {nullptr, nullptr};     // pinned, exc_cap positions each
  unsigned int* h_mode = nullptr;             // pinned, per chunk (the copy engine reads them later)
  unsigned int* h_exc_n = nullptr;
  cudaEvent_t ev[2] = {nullptr, nullptr};     // staging buffer k is free again
  mutable int64_t sent_bytes = 0;             // payload that crossed PCIe in this mode
};

extern "C" int svb_pack2_chunk(const uint8_t* seq4, const int64_t* seq4_offs, const int64_t* offs, const int64_t* pk_offs,
                               int64_t r_lo, int64_t r_hi, int64_t o, int64_t nb, uint8_t* stage, int64_t stage_cap,
                               int64_t* pa2, int64_t* pe2, int64_t* exc_pos, int64_t exc_cap, int64_t* n_exc, int threads);

static int run_search(const IndexDev& d, const svb_reads* R, int overlap, int assemble, svb_sfs_out_t* out,
                      cudaStream_t st, const StreamSrc* src = nullptr) {
  if (overlap > 0) { set_error("overlap must be <= 0 (config.hpp:82 fixes it at -1)"); return SVB_EINVAL; }
  const int64_t n_reads = R->n_reads;
  out->n_reads = n_reads;
  out->block_bytes = d.G * 16;
  out->offs = (int64_t*)calloc((size_t)n_reads + 1, sizeof(int64_t));
  if (!out->offs) { set_error("out of host memory"); return SVB_ENOMEM; }
  if (n_reads == 0) return SVB_OK;
  if (R->total >= (1LL << 40)) { set_error("batch too large"); return SVB_ERANGE; }

  SearchParams P;
  fill_params(P, d);
  P.seq = R->d_seq; P.offs = R->d_offs; P.order = R->d_order; P.sched = R->d_sched; P.n_reads = n_reads;
  P.overlap = overlap; P.assemble = assemble;
  SearchScratch S;
  S.st = st;
  SVB_CUDA(pmalloc((void**)&S.d_ctr, 16 * sizeof(unsigned long long), st));
  // first guess of output capacity; exact count is known after the run, rerun once if it overflowed
  unsigned long long cap = assemble ? (unsigned long long)(4 * n_reads + 1024)
                                    : (unsigned long long)(R->total / 8 + 64 * n_reads + 1024);
  int grid = 0, cfgG = 0;
  SVB_TRY(pick_cfg(d.G, &cfgG));
  const int tma_minb = 9;
  const int tail_minb = 6;   // the tail kernel carries the sprint: more registers, 18 warps per SM (8 CTAs per SM = 80 registers with spills: 27.3 vs 27.5 ms on the config-3 step, 70.2 vs 69.8 ms on configs[1] -- occupancy is not what bounds it)
  int tail_grid = 0;
  const char* e2p = getenv("SVB_SEARCH_TAIL");   // SVB_SEARCH_TAIL=0: one kernel, every walk finished where it started
  const bool two_phase = !(e2p && *e2p == '0');
  if (cfgG == -2) {
    SVB_TRY(persistent_grid(k_sfs_search_mop<tma_minb, false>, TMA_WARPS * 32, d.device, &grid));
    if (src && src->pack2) {   // its own instantiation: its own occupancy
      int grid2 = 0;
      SVB_TRY(persistent_grid(k_sfs_search_mop<tma_minb, false, true>, TMA_WARPS * 32, d.device, &grid2));
      grid = std::min(grid, grid2);
    }
    SVB_TRY(persistent_grid(k_sfs_search_mop<tail_minb, true>, TMA_WARPS * 32, d.device, &tail_grid));
  } else if (cfgG == 0) {
    SVB_TRY(persistent_grid(k_sfs_search_tma<tma_minb, 0>, TMA_WARPS * 32, d.device, &grid));
  } else if (cfgG == -1) {
    SVB_TRY(persistent_grid(k_sfs_search_tma<tma_minb, 1>, TMA_WARPS * 32, d.device, &grid));
  } else {
#define SVB_GRID(g, spl, minb) SVB_TRY(persistent_grid(k_sfs_search<g, spl, minb>, 256, d.device, &grid))
    SVB_DISPATCH_CFG(d.G, cfgG, SVB_GRID);
#undef SVB_GRID
  }
  if (src && src->packed && cfgG != -2) { set_error("packed streamed batches need the default search kernel"); return SVB_EINVAL; }
  cudaEvent_t e0, e1;
  EventPair evp_;
  SVB_CUDA(evp_.create());
  e0 = evp_.a; e1 = evp_.b;
  unsigned long long ctr[16] = {0};
  float kms = 0.f;
  for (int attempt = 0; attempt < 2; ++attempt) {
    pfree(S.d_key, st); pfree(S.d_len, st); S.d_key = nullptr; S.d_len = nullptr;
    SVB_CUDA(pmalloc((void**)&S.d_key, cap * 8, st));
    SVB_CUDA(pmalloc((void**)&S.d_len, cap * 4, st));
    SVB_CUDA(cudaMemsetAsync(S.d_ctr, 0, 16 * sizeof(unsigned long long), st));
    SVB_CUDA(cudaMemsetAsync(S.d_ctr + 6, 0xff, 2 * sizeof(unsigned long long), st));   // [6] first thread in, [7] queue empty: minima
    P.work = S.d_ctr + 0; P.out_count = S.d_ctr + 1; P.stats = S.d_ctr + 2;
    P.out_key = S.d_key; P.out_len = S.d_len; P.out_cap = cap;
    P.ready = nullptr; P.chunk_bytes = 0; P.n_unpack = 0;
    if (src && attempt == 0) {
      P.ready = src->d_ready; P.chunk_bytes = src->chunk_bytes;
      if (src->packed) {
        P.seq4 = src->d_seq4; P.seq4_offs = src->d_seq4_offs; P.arrived = src->d_arrived; P.ready_w = src->d_ready;
        P.chunk_done = src->d_done; P.chunk_r0 = src->d_chunk_r; P.chunk_r1 = src->d_chunk_r + src->n_chunks;
        P.seq_w = R->d_seq; P.total = src->total; P.n_chunks = (int)src->n_chunks; P.n_unpack = src->n_unpack;
        if (src->pack2) {
          P.pk2 = src->d_pk2; P.pk2_offs = src->d_pk2_offs; P.chunk_mode = src->d_mode; P.exc_pos = src->d_exc_pos;
          P.exc_n = src->d_exc_n; P.exc_cap = src->exc_cap;
        }
      }
    }
    SVB_CUDA(cudaEventRecord(e0, st));
    if (cfgG == -2) {
      if (two_phase) {
        if (!S.d_cont) SVB_CUDA(pmalloc(&S.d_cont, (size_t)grid * TMA_WARPS * 32 * sizeof(Cont), st));
        P.cont = static_cast<Cont*>(S.d_cont); P.n_cont = S.d_ctr + 10; P.dry = reinterpret_cast<unsigned int*>(S.d_ctr + 11);
        P.last_ready = (src && attempt == 0) ? src->d_ready + (src->n_chunks - 1) : nullptr;
      }
      if (src && attempt == 0 && src->pack2) k_sfs_search_mop<tma_minb, false, true><<<grid, TMA_WARPS * 32, 0, st>>>(P);
      else k_sfs_search_mop<tma_minb, false><<<grid, TMA_WARPS * 32, 0, st>>>(P);
      if (two_phase) {
        SearchParams Q = P;
        Q.work = S.d_ctr + 12; Q.ready = nullptr; Q.n_unpack = 0; Q.dry = nullptr;
        k_sfs_search_mop<tail_minb, true><<<tail_grid, TMA_WARPS * 32, 0, st>>>(Q);
        out->launches += 1;
      }
    } else if (cfgG == 0) {
      k_sfs_search_tma<tma_minb, 0><<<grid, TMA_WARPS * 32, 0, st>>>(P);
    } else if (cfgG == -1) {
      k_sfs_search_tma<tma_minb, 1><<<grid, TMA_WARPS * 32, 0, st>>>(P);
    } else {
#define SVB_LAUNCH(g, spl, minb) k_sfs_search<g, spl, minb><<<grid, 256, 0, st>>>(P)
      SVB_DISPATCH_CFG(d.G, cfgG, SVB_LAUNCH);
#undef SVB_LAUNCH
    }
    SVB_CUDA(cudaGetLastError());
    SVB_CUDA(cudaEventRecord(e1, st));
    if (src && attempt == 0) {
      // feed the running kernel: chunk copy, then its ready flag, in order on the copy stream
      for (int64_t c = 0; c < src->n_chunks; ++c) {
        const int64_t o = c * src->chunk_bytes, nb = std::min(src->chunk_bytes, src->total - o);
        if (!src->packed) {
          SVB_CUDA(cudaMemcpyAsync(R->d_seq + o, src->host + o, nb, cudaMemcpyHostToDevice, src->copy_stream));
          // flag written by the copy engine too (a memset could need an SM the persistent kernel holds)
          SVB_CUDA(cudaMemcpyAsync(src->d_ready + c, src->h_one, sizeof(unsigned int), cudaMemcpyHostToDevice, src->copy_stream));
        } else {
          // the packed bytes holding the chunk's bases (neighbouring chunks may share a byte: copied
          // twice, harmless); the unpacking CTAs of the running kernel take it from there
          const int64_t r_lo = src->h_chunk_r[c], r_hi = src->h_chunk_r[src->n_chunks + c];
          const int64_t* ho = src->h_offs;
          const int64_t pa = src->h_seq4_offs[r_lo] + (std::max<int64_t>(o - ho[r_lo], 0) >> 1);
          const int64_t last = std::min(o + nb, ho[r_hi + 1]) - 1 - ho[r_hi];   // last base of r_hi in the chunk
          const int64_t pe = last >= 0 ? src->h_seq4_offs[r_hi] + (last >> 1) + 1 : src->h_seq4_offs[r_hi];
          bool sent = false;
          if (src->pack2) {
            // re-pack this chunk to 2 bits per base into the staging buffer the copy before last has released
            const int k = (int)(c & 1);
            SVB_CUDA(cudaEventSynchronize(src->ev[k]));
            int64_t pa2 = 0, pe2 = 0, n_exc = 0;
            const int prc = svb_pack2_chunk(src->host, src->h_seq4_offs, ho, src->h_pk2_offs, r_lo, r_hi, o, nb, src->h_stage[k], src->stage_cap,
                                            &pa2, &pe2, src->h_exc[k], src->exc_cap, &n_exc, 0);
            if (prc == SVB_OK && n_exc <= src->exc_cap) {
              src->h_mode[c] = 2u; src->h_exc_n[c] = (unsigned int)n_exc;
              if (pe2 > pa2) SVB_CUDA(cudaMemcpyAsync(src->d_pk2 + pa2, src->h_stage[k], pe2 - pa2, cudaMemcpyHostToDevice, src->copy_stream));
              if (n_exc) SVB_CUDA(cudaMemcpyAsync(src->d_exc_pos + c * src->exc_cap, src->h_exc[k], n_exc * 8, cudaMemcpyHostToDevice, src->copy_stream));
              SVB_CUDA(cudaEventRecord(src->ev[k], src->copy_stream));
              src->sent_bytes += (pe2 - pa2) + n_exc * 8 + 12;
              sent = true;
            } else {   // too many such positions (or a chunk larger than the staging buffer): this chunk travels as it is
              src->h_mode[c] = 1u; src->h_exc_n[c] = 0u;
            }
          }
          if (!sent) {
            if (pe > pa) SVB_CUDA(cudaMemcpyAsync(src->d_seq4 + pa, src->host + pa, pe - pa, cudaMemcpyHostToDevice, src->copy_stream));
            src->sent_bytes += (pe - pa) + 12;
          }
          if (src->pack2) {
            SVB_CUDA(cudaMemcpyAsync(src->d_mode + c, src->h_mode + c, sizeof(unsigned int), cudaMemcpyHostToDevice, src->copy_stream));
            SVB_CUDA(cudaMemcpyAsync(src->d_exc_n + c, src->h_exc_n + c, sizeof(unsigned int), cudaMemcpyHostToDevice, src->copy_stream));
          }
          SVB_CUDA(cudaMemcpyAsync(src->d_arrived + c, src->h_one, sizeof(unsigned int), cudaMemcpyHostToDevice, src->copy_stream));
        }
      }
    }
    SVB_CUDA(cudaMemcpyAsync(ctr, S.d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    SVB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    kms += ms;
    out->launches += 1;
    if (ctr[5]) {
      // copies queued on the copy stream still read the caller's staging buffers: let them drain before the caller gets its buffers back
      if (src && src->copy_stream) cudaStreamSynchronize(src->copy_stream);
      set_error("the read stream stalled: chunks queued behind the search kernel never arrived");
      return SVB_ECUDA;
    }
    if (ctr[1] <= cap) break;
    cap = ctr[1];
    if (attempt == 1) { set_error("output overflow persisted"); return SVB_ERANGE; }
  }
  out->kernel_ms = kms;
  out->n_ext = (int64_t)ctr[2];
  out->n_blocks_touched = (int64_t)ctr[3];
  out->n_text_ext = (int64_t)ctr[4];
  if (getenv("SVB_SEARCH_STATS") && ctr[8])   // k_sfs_search_mop only
    fprintf(stderr, "[k_sfs_search_mop] ext %llu blocks %llu text %llu | work queue empty after %.1f ms, last warp done after %.1f ms, %llu walks finished by the tail kernel\n",
            ctr[2], ctr[3], ctr[4], (ctr[7] - ctr[6]) * 1e-6, (ctr[8] - ctr[6]) * 1e-6, ctr[10]);
  if (getenv("SVB_SEARCH_STATS") && ctr[13])
    fprintf(stderr, "[k_sfs_search_mop tail] sprints %llu, links resolved by them %llu (%.1f per sprint), sprints ended by an abandoned link %llu; mean busy time of a warp %.1f ms\n",
            ctr[13], ctr[14], ctr[14] / (double)ctr[13], ctr[15], ctr[9] * 1e-6 / (double)(tail_grid * TMA_WARPS));
  const int64_t m = (int64_t)ctr[1];
  out->n_sfs = m;
  if (m == 0) return SVB_OK;
  // order records by (read, qs asc | emit order)
  SVB_CUDA(pmalloc((void**)&S.d_key2, m * 8, st));
  SVB_CUDA(pmalloc((void**)&S.d_len2, m * 4, st));
  cub::DoubleBuffer<uint64_t> dk(S.d_key, S.d_key2);
  cub::DoubleBuffer<uint32_t> dv(S.d_len, S.d_len2);
  int rbits = 1;
  while (rbits < 32 && ((uint64_t)n_reads >> rbits)) ++rbits;
  size_t bytes = 0;
  SVB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, m, 0, 32 + rbits, st));
  SVB_CUDA(pmalloc(&S.d_tmp, bytes, st));
  SVB_CUDA(cub::DeviceRadixSort::SortPairs(S.d_tmp, bytes, dk, dv, m, 0, 32 + rbits, st));
  out->launches += 1;
  std::vector<uint64_t> hkey((size_t)m);
  out->qs = (int32_t*)malloc((size_t)m * 4);
  out->len = (int32_t*)malloc((size_t)m * 4);
  if (!out->qs || !out->len) { set_error("out of host memory"); return SVB_ENOMEM; }
  SVB_CUDA(cudaMemcpyAsync(hkey.data(), dk.Current(), m * 8, cudaMemcpyDeviceToHost, st));
  SVB_CUDA(cudaMemcpyAsync(out->len, dv.Current(), m * 4, cudaMemcpyDeviceToHost, st));
  SVB_CUDA(cudaStreamSynchronize(st));
  out->d2h_bytes += m * 12 + (int64_t)sizeof(ctr);
  for (int64_t i = 0; i < m; ++i) {
    uint32_t r = (uint32_t)(hkey[i] >> 32);
    uint32_t sk = (uint32_t)hkey[i];
    out->qs[i] = (int32_t)(assemble ? sk : ~sk);
    out->offs[r + 1]++;
  }
  for (int64_t r = 0; r < n_reads; ++r) out->offs[r + 1] += out->offs[r];
  return SVB_OK;
}

}  // namespace svb

extern "C" {

int svb_reads_upload(const uint8_t* seq, const int64_t* offs, int64_t n_reads, int mem, int device,
                     svb_reads_t** out) {
  SVB_TRY(check_device(device));
  if (!offs || n_reads < 0 || !out || (mem != SVB_MEM_HOST && mem != SVB_MEM_DEVICE)) {
    set_error("svb_reads_upload: bad arguments");
    return SVB_EINVAL;
  }
  svb_reads* R = new svb_reads();
  R->device = device;
  R->n_reads = n_reads;
  int64_t first = 0, last = 0;
  if (mem == SVB_MEM_HOST) { first = offs[0]; last = offs[n_reads]; }
  else {
    cudaMemcpy(&first, offs, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&last, offs + n_reads, 8, cudaMemcpyDeviceToHost);
  }
  R->total = last - first;
  int rc = SVB_OK;
  do {
    if (R->total < 0 || (R->total > 0 && !seq)) { set_error("svb_reads_upload: bad offsets"); rc = SVB_EINVAL; break; }
    if (pmalloc((void**)&R->d_offs, (n_reads + 1) * 8, 0) != cudaSuccess) { rc = SVB_ENOMEM; break; }
    if (mem == SVB_MEM_HOST) {
      std::vector<int64_t> rb((size_t)n_reads + 1);
      for (int64_t i = 0; i <= n_reads; ++i) {
        rb[i] = offs[i] - first;
        if (i && rb[i] < rb[i - 1]) { set_error("read offsets must be non-decreasing"); rc = SVB_EINVAL; break; }
      }
      if (rc) break;
      // one spare window (<= 32 B) after the last base, rounded to 64 B
      size_t padded = ((size_t)R->total + 64 + 63) & ~(size_t)63;
      if (pmalloc((void**)&R->d_seq, padded, 0) != cudaSuccess) { rc = SVB_ENOMEM; break; }
      cudaMemcpy(R->d_offs, rb.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice);
      if (R->total) cudaMemcpy(R->d_seq, seq + first, R->total, cudaMemcpyHostToDevice);
      cudaMemset(R->d_seq + R->total, 0, padded - R->total);
    } else {
      // device buffers are adopted as they are: the caller guarantees offs[0] == 0, 32-byte
      // alignment of seq and at least 64 readable bytes after the last base
      if (first != 0) { set_error("device-resident batches must start at offset 0"); rc = SVB_EINVAL; break; }
      R->owns_seq = false;
      R->d_seq = const_cast<uint8_t*>(seq);
      cudaMemcpy(R->d_offs, offs, (n_reads + 1) * 8, cudaMemcpyDeviceToDevice);
    }
    rc = make_order(R, 0, 0);
  } while (0);
  if (rc == SVB_ENOMEM) set_error("out of device memory uploading reads");
  if (rc != SVB_OK) { svb_reads_free(R); return rc; }
  if (cudaDeviceSynchronize() != cudaSuccess) { svb_reads_free(R); set_error("upload failed"); return SVB_ECUDA; }
  *out = R;
  return SVB_OK;
}

void svb_reads_free(svb_reads_t* R) {
  if (!R) return;
  cudaSetDevice(R->device);
  if (R->owns_seq && R->d_seq) pfree(R->d_seq, 0);
  pfree(R->d_offs, 0);
  pfree(R->d_order, 0);
  pfree(R->d_sched, 0);
  delete R;
}

int svb_sfs_resident(const svb_index_t* idx, const svb_reads_t* reads, int overlap, int assemble,
                     svb_sfs_out_t* out) {
  if (!idx || !reads || !out) { set_error("svb_sfs_resident: null argument"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  SVB_TRY(check_device(idx->dev.device));
  if (reads->device != idx->dev.device) { set_error("reads and index live on different devices"); return SVB_EINVAL; }
  cudaEvent_t e0, e1;
  EventPair evp_;
  SVB_CUDA(evp_.create());
  e0 = evp_.a; e1 = evp_.b;
  SVB_CUDA(cudaEventRecord(e0, 0));
  int rc = run_search(idx->dev, reads, overlap, assemble, out, 0);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&out->device_ms, e0, e1);
  if (rc != SVB_OK) svb_sfs_out_free(out);
  return rc;
}

// streamed variant of svb_sfs_batch: offsets first, kernel launched at once, read bytes follow in
// 128 MiB chunks on a copy stream while the kernel works (chunk-major hand-out order)
// packed mode (seq4_offs != nullptr): `seq` holds BAM-native 4-bit reads, `offs` their base offsets
// (offs[0] == 0) and seq4_offs their byte offsets; with stream == false everything is copied and
// unpacked before the kernel starts (small batches)
static int sfs_batch_streamed(const svb_index_t* idx, const uint8_t* seq, const int64_t* offs, int64_t n_reads,
                              int overlap, int assemble, svb_sfs_out_t* out, const int64_t* seq4_offs = nullptr,
                              bool stream = true) {
  const IndexDev& d = idx->dev;
  svb_reads R;
  R.device = d.device;
  R.n_reads = n_reads;
  const int64_t first = offs[0];
  R.total = offs[n_reads] - first;
  StreamSrc src;
  src.host = seq4_offs ? seq + seq4_offs[0] : seq + first;
  src.total = R.total;
  src.packed = seq4_offs != nullptr;
  src.n_reads = n_reads;
  int64_t* d_s4o = nullptr;
  std::vector<int64_t> s4rb, chunk_r, pk2o;
  const int64_t packed_total = seq4_offs ? seq4_offs[n_reads] - seq4_offs[0] : 0;
  src.chunk_bytes = (int64_t)128 << 20;
  if (const char* e = getenv("SVB_STREAM_CHUNK_BYTES")) { long long v = atoll(e); if (v >= 4096) src.chunk_bytes = (v + 63) & ~63LL; }
  src.n_chunks = (R.total + src.chunk_bytes - 1) / src.chunk_bytes;
  cudaStream_t comp = nullptr;
  int rc = SVB_OK;
  std::vector<int64_t> rb((size_t)n_reads + 1);
  for (int64_t i = 0; i <= n_reads; ++i) {
    rb[i] = offs[i] - first;
    if (i && rb[i] < rb[i - 1]) { set_error("read offsets must be non-decreasing"); return SVB_EINVAL; }
  }
  const size_t padded = ((size_t)R.total + 64 + 63) & ~(size_t)63;
#define SCHECK(expr)                                                                              \
  do {                                                                                            \
    cudaError_t _e = (expr);                                                                      \
    if (_e != cudaSuccess) {                                                                      \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));       \
      rc = SVB_ECUDA;                                                                             \
      goto done;                                                                                  \
    }                                                                                             \
  } while (0)
  SCHECK(cudaStreamCreateWithFlags(&comp, cudaStreamNonBlocking));
  SCHECK(cudaStreamCreateWithFlags(&src.copy_stream, cudaStreamNonBlocking));
  SCHECK(pmalloc((void**)&R.d_seq, padded, comp));
  SCHECK(pmalloc((void**)&R.d_offs, (n_reads + 1) * 8, comp));
  SCHECK(pmalloc((void**)&src.d_ready, src.n_chunks * sizeof(unsigned int), comp));
  SCHECK(cudaHostAlloc((void**)&src.h_one, sizeof(unsigned int), cudaHostAllocDefault));
  *src.h_one = 1u;
  SCHECK(cudaMemsetAsync(src.d_ready, 0, src.n_chunks * sizeof(unsigned int), comp));
  SCHECK(cudaMemsetAsync(R.d_seq + R.total, 0, padded - R.total, comp));
  SCHECK(cudaMemcpyAsync(R.d_offs, rb.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, comp));
  if (src.packed) {
    s4rb.resize((size_t)n_reads + 1);
    for (int64_t i = 0; i <= n_reads; ++i) s4rb[i] = seq4_offs[i] - seq4_offs[0];
    SCHECK(pmalloc((void**)&src.d_seq4, (size_t)packed_total + 16, comp));
    SCHECK(pmalloc((void**)&d_s4o, (n_reads + 1) * 8, comp));
    SCHECK(cudaMemcpyAsync(d_s4o, s4rb.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, comp));
    src.d_seq4_offs = d_s4o;
    src.h_offs = rb.data();
    src.h_seq4_offs = s4rb.data();
    if (stream) {
      // reads reaching into each chunk of base positions; flags; how many CTAs unpack
      chunk_r.assign((size_t)src.n_chunks * 2, 0);
      int64_t r_lo = 0;
      for (int64_t c = 0; c < src.n_chunks; ++c) {
        const int64_t o = c * src.chunk_bytes, nb = std::min(src.chunk_bytes, src.total - o);
        while (r_lo + 1 < n_reads && rb[r_lo + 1] <= o) ++r_lo;
        int64_t r_hi = r_lo;
        while (r_hi + 1 < n_reads && rb[r_hi + 1] < o + nb) ++r_hi;
        chunk_r[c] = r_lo; chunk_r[src.n_chunks + c] = r_hi;
      }
      src.h_chunk_r = chunk_r.data();
      SCHECK(pmalloc((void**)&src.d_chunk_r, src.n_chunks * 16, comp));
      SCHECK(pmalloc((void**)&src.d_arrived, src.n_chunks * 4, comp));
      SCHECK(pmalloc((void**)&src.d_done, src.n_chunks * 4, comp));
      SCHECK(cudaMemcpyAsync(src.d_chunk_r, chunk_r.data(), src.n_chunks * 16, cudaMemcpyHostToDevice, comp));
      SCHECK(cudaMemsetAsync(src.d_arrived, 0, src.n_chunks * 4, comp));
      SCHECK(cudaMemsetAsync(src.d_done, 0, src.n_chunks * 4, comp));
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, d.device);
      src.n_unpack = sms;   // one CTA in nine: ~23 GB of streaming traffic per 1 M reads, far below their share
      if (const char* e = getenv("SVB_UNPACK_CTAS")) src.n_unpack = std::max(1, atoi(e));
      if (const char* e = getenv("SVB_STREAM_PACK2")) src.pack2 = atoi(e) != 0;
      if (src.pack2) {
        // 2-bit transport: packed layout (read r at pk2_offs[r], (l + 3) / 4 bytes), staging, per-chunk mode / patch lists
        pk2o.resize((size_t)n_reads + 1);
        pk2o[0] = 0;
        int64_t lmax = 0;
        for (int64_t i = 0; i < n_reads; ++i) { const int64_t l = rb[i + 1] - rb[i]; pk2o[i + 1] = pk2o[i] + (l + 3) / 4; lmax = std::max(lmax, l); }
        src.h_pk2_offs = pk2o.data();
        src.exc_cap = 1 << 16;   // positions a chunk may patch before it falls back to the 4-bit form (SVB_PACK2_EXC_CAP: tests)
        if (const char* e = getenv("SVB_PACK2_EXC_CAP")) src.exc_cap = std::max(1, atoi(e));
        src.stage_cap = src.chunk_bytes / 4 + 2 * ((lmax + 3) / 4) + 4096 + (chunk_r.empty() ? 0 : [&]() { int64_t m = 0; for (int64_t c = 0; c < src.n_chunks; ++c) m = std::max(m, chunk_r[src.n_chunks + c] - chunk_r[c] + 1); return m; }());
        SCHECK(pmalloc((void**)&src.d_pk2, (size_t)pk2o[n_reads] + 16, comp));
        SCHECK(pmalloc((void**)&src.d_pk2_offs, (n_reads + 1) * 8, comp));
        SCHECK(pmalloc((void**)&src.d_mode, src.n_chunks * 4, comp));
        SCHECK(pmalloc((void**)&src.d_exc_n, src.n_chunks * 4, comp));
        SCHECK(pmalloc((void**)&src.d_exc_pos, (size_t)src.n_chunks * src.exc_cap * 8, comp));
        SCHECK(cudaMemcpyAsync(src.d_pk2_offs, pk2o.data(), (n_reads + 1) * 8, cudaMemcpyHostToDevice, comp));
        SCHECK(cudaMemsetAsync(src.d_pk2 + pk2o[n_reads], 0, 16, comp));
        for (int k = 0; k < 2; ++k) {
          SCHECK(cudaHostAlloc((void**)&src.h_stage[k], (size_t)src.stage_cap, cudaHostAllocDefault));
          SCHECK(cudaHostAlloc((void**)&src.h_exc[k], (size_t)src.exc_cap * 8, cudaHostAllocDefault));
          SCHECK(cudaEventCreateWithFlags(&src.ev[k], cudaEventDisableTiming));
        }
        SCHECK(cudaHostAlloc((void**)&src.h_mode, (size_t)src.n_chunks * 4, cudaHostAllocDefault));
        SCHECK(cudaHostAlloc((void**)&src.h_exc_n, (size_t)src.n_chunks * 4, cudaHostAllocDefault));
      }
    }
  }
  if (!stream) {   // small packed batch: copy, unpack, then search the resident reads
    if (packed_total) SCHECK(cudaMemcpyAsync(src.d_seq4, src.host, packed_total, cudaMemcpyHostToDevice, comp));
    if (n_reads && R.total) {
      k_unpack4<<<(unsigned)std::min<int64_t>((n_reads + 3) / 4, 148 * 8), 128, 0, comp>>>(src.d_seq4, d_s4o, R.d_offs, n_reads, R.total, R.d_seq);
      SCHECK(cudaGetLastError());
    }
    rc = make_order(&R, 0, comp);
    if (rc == SVB_OK) rc = run_search(d, &R, overlap, assemble, out, comp, nullptr);
    if (rc == SVB_OK) { out->h2d_bytes = packed_total + (n_reads + 1) * 16; out->launches += 3; }
    goto done;
  }
  rc = make_order(&R, src.chunk_bytes, comp);  // synchronises comp: the allocations above are usable on copy_stream
  if (rc == SVB_OK) rc = run_search(d, &R, overlap, assemble, out, comp, &src);
  if (rc == SVB_OK) {
    SCHECK(cudaStreamSynchronize(src.copy_stream));
    out->h2d_bytes = (src.pack2 ? src.sent_bytes + (n_reads + 1) * 16 : src.packed ? packed_total + (n_reads + 1) * 8 : R.total) + (n_reads + 1) * 8 + src.n_chunks * 4;
    out->launches += 2;  // read keys + order sort
  }
done:
#undef SCHECK
  if (src.copy_stream) { cudaStreamSynchronize(src.copy_stream); cudaStreamDestroy(src.copy_stream); }
  if (comp) {
    pfree(R.d_seq, comp); pfree(R.d_offs, comp); pfree(R.d_order, comp); pfree(R.d_sched, comp); pfree(src.d_ready, comp);
    pfree(src.d_seq4, comp); pfree(d_s4o, comp); pfree(src.d_chunk_r, comp); pfree(src.d_arrived, comp); pfree(src.d_done, comp);
    pfree(src.d_pk2, comp); pfree(src.d_pk2_offs, comp); pfree(src.d_mode, comp); pfree(src.d_exc_n, comp); pfree(src.d_exc_pos, comp);
    cudaStreamSynchronize(comp);
    cudaStreamDestroy(comp);
  }
  if (src.h_one) cudaFreeHost(src.h_one);
  for (int k = 0; k < 2; ++k) {
    if (src.h_stage[k]) cudaFreeHost(src.h_stage[k]);
    if (src.h_exc[k]) cudaFreeHost(src.h_exc[k]);
    if (src.ev[k]) cudaEventDestroy(src.ev[k]);
  }
  if (src.h_mode) cudaFreeHost(src.h_mode);
  if (src.h_exc_n) cudaFreeHost(src.h_exc_n);
  return rc;
}

int svb_sfs_batch(const svb_index_t* idx, const uint8_t* seq, const int64_t* offs, int64_t n_reads,
                  int overlap, int assemble, svb_sfs_out_t* out) {
  if (!idx || !offs || !out || n_reads < 0) { set_error("svb_sfs_batch: bad arguments"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  SVB_TRY(check_device(idx->dev.device));
  cudaEvent_t e0, e1;
  EventPair evp_;
  SVB_CUDA(evp_.create());
  e0 = evp_.a; e1 = evp_.b;
  SVB_CUDA(cudaEventRecord(e0, 0));
  int rc = SVB_OK, cfgG = 0;
  rc = pick_cfg(idx->dev.G, &cfgG);
  const int64_t total = n_reads > 0 ? offs[n_reads] - offs[0] : 0;
  const char* ns = getenv("SVB_NO_STREAM");
  int64_t stream_min = (int64_t)32 << 20;  // below this a plain upload is as fast
  if (const char* e = getenv("SVB_STREAM_MIN_BYTES")) stream_min = atoll(e);
  if (rc == SVB_OK && cfgG <= 0 && seq && total >= stream_min && total > 0 && !(ns && *ns == '1')) {
    rc = sfs_batch_streamed(idx, seq, offs, n_reads, overlap, assemble, out);
  } else if (rc == SVB_OK) {
    svb_reads_t* R = nullptr;
    rc = svb_reads_upload(seq, offs, n_reads, SVB_MEM_HOST, idx->dev.device, &R);
    if (rc == SVB_OK) {
      rc = run_search(idx->dev, R, overlap, assemble, out, 0);
      out->h2d_bytes = R->total + (n_reads + 1) * 8;
      out->launches += 2;  // read keys + order sort
    }
    svb_reads_free(R);
  }
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&out->device_ms, e0, e1);
  if (rc != SVB_OK) svb_sfs_out_free(out);
  return rc;
}

int svb_sfs_batch_bam4(const svb_index_t* idx, const uint8_t* seq4, const int64_t* seq4_offs, const int32_t* l_qseq,
                       int64_t n_reads, int overlap, int assemble, svb_sfs_out_t* out) {
  if (!idx || !seq4_offs || !out || n_reads < 0 || (n_reads > 0 && !l_qseq)) { set_error("svb_sfs_batch_bam4: bad arguments"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  SVB_TRY(check_device(idx->dev.device));
  int cfgG = 0;
  SVB_TRY(pick_cfg(idx->dev.G, &cfgG));
  if (cfgG > 0) { set_error("svb_sfs_batch_bam4 needs a 128-byte-block index (the staged search kernels)"); return SVB_EINVAL; }
  std::vector<int64_t> offs((size_t)n_reads + 1, 0);
  for (int64_t i = 0; i < n_reads; ++i) {
    const int64_t l = l_qseq[i], pb = seq4_offs[i + 1] - seq4_offs[i];
    if (l < 0 || pb < 0 || (l + 1) / 2 > pb) { set_error("read %lld: l_qseq %lld does not fit its %lld packed bytes", (long long)i, (long long)l, (long long)pb); return SVB_EINVAL; }
    offs[i + 1] = offs[i] + l;
  }
  if (offs[n_reads] > 0 && !seq4) { set_error("svb_sfs_batch_bam4: null sequence buffer"); return SVB_EINVAL; }
  cudaEvent_t e0, e1;
  EventPair evp_;
  SVB_CUDA(evp_.create());
  e0 = evp_.a; e1 = evp_.b;
  SVB_CUDA(cudaEventRecord(e0, 0));
  const char* ns = getenv("SVB_NO_STREAM");
  int64_t stream_min = (int64_t)32 << 20;
  if (const char* e = getenv("SVB_STREAM_MIN_BYTES")) stream_min = atoll(e);
  const bool stream = offs[n_reads] >= stream_min && !(ns && *ns == '1');
  int rc = SVB_OK;
  if (n_reads == 0 || offs[n_reads] == 0) {
    out->n_reads = n_reads;
    out->offs = (int64_t*)calloc((size_t)n_reads + 1, sizeof(int64_t));
    if (!out->offs) { set_error("out of host memory"); rc = SVB_ENOMEM; }
  } else {
    rc = sfs_batch_streamed(idx, seq4, offs.data(), n_reads, overlap, assemble, out, seq4_offs, stream);
  }
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&out->device_ms, e0, e1);
  if (rc != SVB_OK) svb_sfs_out_free(out);
  return rc;
}

// test / bench utility: nt6 bytes (device) -> BAM-native 4-bit reads (device), every read starting
// on a byte boundary; out_offs[r] = sum over earlier reads of (len + 1) / 2
__global__ void k_pack4(const uint8_t* __restrict__ seq, const int64_t* __restrict__ offs, const int64_t* __restrict__ poffs,
                        int64_t n, uint8_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; r < n; r += nwarps) {
    const int64_t o = offs[r], l = offs[r + 1] - o, pb = poffs[r];
    for (int64_t b = lane; b < (l + 1) / 2; b += 32) {
      auto enc = [](uint8_t c) -> unsigned { return c == 1 ? 1u : c == 2 ? 2u : c == 3 ? 4u : c == 4 ? 8u : 15u; };
      const unsigned hi = enc(seq[o + 2 * b]), lo = (2 * b + 1 < l) ? enc(seq[o + 2 * b + 1]) : 0u;
      out[pb + b] = (uint8_t)((hi << 4) | lo);
    }
  }
}
int svb_pack4_device(const uint8_t* d_seq, const int64_t* d_offs, const int64_t* d_seq4_offs, int64_t n_reads, int device,
                     uint8_t* d_out) {
  SVB_TRY(check_device(device));
  if (n_reads <= 0) return SVB_OK;
  k_pack4<<<(unsigned)std::min<int64_t>((n_reads + 3) / 4, 148 * 16), 128>>>(d_seq, d_offs, d_seq4_offs, n_reads, d_out);
  SVB_CUDA(cudaGetLastError());
  SVB_CUDA(cudaDeviceSynchronize());
  return SVB_OK;
}

// Test / bench utility: decode a batch packed 2 bits per base by svb_pack2_host (HOST buffers: packed bytes with
// their n_reads + 1 byte offsets, the n_reads + 1 base offsets) on `device` and return the nt6 bytes (host,
// offs[n_reads] bytes).  Reads flagged as exceptions by the packer have no 2-bit form and must not be passed.
int svb_unpack2_device(const uint8_t* pk, const int64_t* pk_offs, const int64_t* offs, int64_t n_reads, int device, uint8_t* out_host) {
  if (!pk_offs || !offs || !out_host || n_reads < 0) { set_error("svb_unpack2_device: bad arguments"); return SVB_EINVAL; }
  SVB_TRY(check_device(device));
  if (n_reads == 0 || offs[n_reads] - offs[0] <= 0) return SVB_OK;
  if (!pk || offs[0] != 0 || pk_offs[0] != 0) { set_error("svb_unpack2_device: offsets must start at 0"); return SVB_EINVAL; }
  const int64_t total = offs[n_reads], pbytes = pk_offs[n_reads];
  uint8_t *d_pk = nullptr, *d_out = nullptr;
  int64_t *d_po = nullptr, *d_o = nullptr;
  int rc = SVB_OK;
  auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == SVB_OK) { set_error("svb_unpack2_device: %s", cudaGetErrorString(e)); rc = SVB_ECUDA; } };
  fail(cudaMalloc((void**)&d_pk, (size_t)pbytes + 16));
  fail(cudaMalloc((void**)&d_out, (size_t)total + 64));
  fail(cudaMalloc((void**)&d_po, (size_t)(n_reads + 1) * 8));
  fail(cudaMalloc((void**)&d_o, (size_t)(n_reads + 1) * 8));
  if (rc == SVB_OK) {
    fail(cudaMemset(d_pk + pbytes, 0, 16));
    fail(cudaMemcpy(d_pk, pk, (size_t)pbytes, cudaMemcpyHostToDevice));
    fail(cudaMemcpy(d_po, pk_offs, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice));
    fail(cudaMemcpy(d_o, offs, (size_t)(n_reads + 1) * 8, cudaMemcpyHostToDevice));
  }
  if (rc == SVB_OK) {
    k_unpack2<<<(unsigned)std::min<int64_t>((n_reads + 3) / 4, 148 * 8), 128>>>(d_pk, d_po, d_o, n_reads, total, d_out);
    fail(cudaGetLastError());
    fail(cudaDeviceSynchronize());
    fail(cudaMemcpy(out_host, d_out, (size_t)total, cudaMemcpyDeviceToHost));
  }
  cudaFree(d_pk); cudaFree(d_out); cudaFree(d_po); cudaFree(d_o);
  return rc;
}

void svb_sfs_out_free(svb_sfs_out_t* out) {
  if (!out) return;
  free(out->offs); free(out->qs); free(out->len);
  out->offs = nullptr; out->qs = nullptr; out->len = nullptr;
}

int svb_rank2a(const svb_index_t* idx, const int64_t* k, const int64_t* l, int64_t n, int64_t* ok6, int64_t* ol6) {
  if (!idx || !k || !l || !ok6 || !ol6 || n < 0) { set_error("svb_rank2a: bad arguments"); return SVB_EINVAL; }
  SVB_TRY(check_device(idx->dev.device));
  if (n == 0) return SVB_OK;
  const IndexDev& d = idx->dev;
  for (int64_t i = 0; i < n; ++i)
    if (k[i] < 0 || l[i] < k[i] || l[i] > d.n) { set_error("rank query %lld out of range", (long long)i); return SVB_ERANGE; }
  int64_t *dk = nullptr, *dl = nullptr, *dok = nullptr, *dol = nullptr;
  SVB_CUDA(cudaMalloc((void**)&dk, n * 8));
  SVB_CUDA(cudaMalloc((void**)&dl, n * 8));
  SVB_CUDA(cudaMalloc((void**)&dok, n * 48));
  SVB_CUDA(cudaMalloc((void**)&dol, n * 48));
  SVB_CUDA(cudaMemcpy(dk, k, n * 8, cudaMemcpyHostToDevice));
  SVB_CUDA(cudaMemcpy(dl, l, n * 8, cudaMemcpyHostToDevice));
  SearchParams P;
  fill_params(P, d);
  int64_t groups = std::min<int64_t>(n, 148 * 64);
  int cfgG = 0;
  SVB_TRY(pick_cfg(d.G, &cfgG));
  if (cfgG <= 0) cfgG = 8;  // the all-symbol rank entry uses the lane-group kernel
  unsigned grid = (unsigned)((groups * cfgG + 255) / 256);
#define SVB_LAUNCH(g, spl, minb) k_rank2a<g, spl><<<grid, 256>>>(P, dk, dl, n, d.n, dok, dol)
  SVB_DISPATCH_CFG(d.G, cfgG, SVB_LAUNCH);
#undef SVB_LAUNCH
  SVB_CUDA(cudaGetLastError());
  SVB_CUDA(cudaMemcpy(ok6, dok, n * 48, cudaMemcpyDeviceToHost));
  SVB_CUDA(cudaMemcpy(ol6, dol, n * 48, cudaMemcpyDeviceToHost));
  cudaFree(dk); cudaFree(dl); cudaFree(dok); cudaFree(dol);
  return SVB_OK;
}

int svb_rank_bench(const svb_index_t* idx, int64_t nq, int64_t delta, uint64_t seed, int iters,
                   float* ms_per_iter, int64_t* blocks_touched) {
  if (!idx || nq <= 0 || iters <= 0 || !ms_per_iter) { set_error("svb_rank_bench: bad arguments"); return SVB_EINVAL; }
  SVB_TRY(check_device(idx->dev.device));
  const IndexDev& d = idx->dev;
  if (delta < 0 || delta >= d.n) { set_error("delta out of range"); return SVB_ERANGE; }
  unsigned long long* dctr = nullptr;
  SVB_CUDA(cudaMalloc((void**)&dctr, 16));
  SVB_CUDA(cudaMemset(dctr, 0, 16));
  SearchParams P;
  fill_params(P, d);
  int grid = 0, cfgG = 0;
  SVB_TRY(pick_cfg(d.G, &cfgG));
  if (cfgG == -2) cfgG = -1;  // the microbenchmark has one cp.async-staged variant
  if (cfgG == 0) {
    SVB_TRY(persistent_grid(k_rank_bench_tma<0>, TMA_WARPS * 32, d.device, &grid));
  } else if (cfgG == -1) {
    SVB_TRY(persistent_grid(k_rank_bench_tma<1>, TMA_WARPS * 32, d.device, &grid));
  } else {
#define SVB_GRID(g, spl, minb) SVB_TRY(persistent_grid(k_rank_bench<g, spl>, 256, d.device, &grid))
    SVB_DISPATCH_CFG(d.G, cfgG, SVB_GRID);
#undef SVB_GRID
  }
  cudaEvent_t e0, e1;
  EventPair evp_;
  SVB_CUDA(evp_.create());
  e0 = evp_.a; e1 = evp_.b;
  // one untimed warm-up launch, then `iters` timed launches with distinct seeds
  for (int it = -1; it < iters; ++it) {
    if (it == 0) { SVB_CUDA(cudaMemset(dctr, 0, 16)); SVB_CUDA(cudaEventRecord(e0, 0)); }
    uint64_t sd = seed + 0x1000003ULL * (uint64_t)(it + 1);
    if (cfgG == 0) {
      k_rank_bench_tma<0><<<grid, TMA_WARPS * 32>>>(P, nq, d.n, delta, sd, dctr, dctr + 1);
    } else if (cfgG == -1) {
      k_rank_bench_tma<1><<<grid, TMA_WARPS * 32>>>(P, nq, d.n, delta, sd, dctr, dctr + 1);
    } else {
#define SVB_LAUNCH(g, spl, minb) k_rank_bench<g, spl><<<grid, 256>>>(P, nq, d.n, delta, sd, dctr, dctr + 1)
      SVB_DISPATCH_CFG(d.G, cfgG, SVB_LAUNCH);
#undef SVB_LAUNCH
    }
  }
  SVB_CUDA(cudaGetLastError());
  SVB_CUDA(cudaEventRecord(e1, 0));
  SVB_CUDA(cudaEventSynchronize(e1));
  float ms = 0.f;
  SVB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
  *ms_per_iter = ms / iters;
  unsigned long long ctr[2];
  SVB_CUDA(cudaMemcpy(ctr, dctr, 16, cudaMemcpyDeviceToHost));
  if (blocks_touched) *blocks_touched = (int64_t)(ctr[1] / (unsigned long long)iters);
  cudaFree(dctr);
  return SVB_OK;
}

}  // extern "C"
