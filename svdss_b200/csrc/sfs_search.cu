// SFS extraction on the GPU: the ping-pong FMD search of PingPong::ping_pong_search
// (reference ping_pong.cpp:4-49) for a whole batch of reads, with Assembler::assemble
// (assembler.cpp:34-56) fused as a streaming epilogue.
//
// Work decomposition: one G-lane *group* (G = 4 for 64-byte index blocks, 8 for 128-byte blocks)
// walks one read; a warp therefore carries 32/G independent dependent-load chains.  Per backward
// extension each lane issues ONE 16-byte load per index block (the group's G loads coalesce into
// one 64/128-byte request), counts matches in its own 32-symbol slice with LOP3+POPC and the group
// reduces with warp shuffles.  Groups pull reads (longest first) from a global work counter, so
// the grid is persistent: #CTAs = #SMs x occupancy.
//
// Only Occ(c, k) and Occ(c, k + size) of ONE symbol are needed per step: every direction switch in
// ping_pong.cpp restarts from rb3_fmd_set_intv (:12, :30), so the bidirectional bookkeeping of
// rb3_fmd_extend is never observed; the forward phase is a backward search with complemented
// characters on the same (strand-closed) BWT.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>

#include "common.cuh"

struct svb_reads {
  int device = 0;
  int64_t n_reads = 0;
  int64_t total = 0;           // bytes of sequence
  uint8_t* d_seq = nullptr;    // padded to 64 B beyond total
  int64_t* d_offs = nullptr;   // n_reads + 1
  uint32_t* d_order = nullptr; // read indices, longest first
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
  int64_t n_reads;
  int overlap;
  int assemble;
  unsigned long long* work;      // next read to hand out
  unsigned long long* out_count; // records appended
  unsigned long long* stats;     // [0] extensions, [1] blocks touched
  uint64_t* out_key;             // read << 32 | sort key
  uint32_t* out_len;
  unsigned long long out_cap;
};

__device__ __forceinline__ uint4 ldg_slice(const uint4* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// Occ-based backward extension of [k, k+s) by symbol c (1..5) for one G-lane group.
// lg = lane within group, gbase = first lane of the group, gmask = group's lane mask.
template <int G>
__device__ __forceinline__ void extend_group(const SearchParams& P, int c, uint64_t& k, uint64_t& s,
                                             int lg, int gbase, unsigned gmask, unsigned& nblk) {
  constexpr int LOGB = (G == 4) ? 7 : 8;
  constexpr unsigned BMASK = (1u << LOGB) - 1;
  const uint64_t l = k + s;
  const uint64_t bk = k >> LOGB, bl = l >> LOGB;
  const uint4 sk = ldg_slice(P.blocks + bk * G + lg);
  uint4 sl = sk;
  const bool two = (bl != bk);
  if (two) sl = ldg_slice(P.blocks + bl * G + lg);
  nblk += two ? 2u : 1u;
  // superblock bases (L1-resident, tiny): acc[c] + Occ(c, sb << 32)
  const int64_t sbk = __ldg(P.sbase + (k >> 32) * 8 + c);
  const int64_t sbl = __ldg(P.sbase + (l >> 32) * 8 + c);
  const unsigned c0 = (c & 1) ? 0u : ~0u, c1 = (c & 2) ? 0u : ~0u, c2 = (c & 4) ? 0u : ~0u;
  const unsigned mk = (sk.y ^ c0) & (sk.z ^ c1) & (sk.w ^ c2);
  const unsigned ml = (sl.y ^ c0) & (sl.z ^ c1) & (sl.w ^ c2);
  int ok_ = (int)((unsigned)k & BMASK) - 32 * lg;
  int ol_ = (int)((unsigned)l & BMASK) - 32 * lg;
  ok_ = max(0, min(32, ok_));
  ol_ = max(0, min(32, ol_));
  const unsigned maskk = ok_ >= 32 ? ~0u : ((1u << ok_) - 1u);
  const unsigned maskl = ol_ >= 32 ? ~0u : ((1u << ol_) - 1u);
  unsigned v = __popc(mk & maskk) | (__popc(ml & maskl) << 16);
#pragma unroll
  for (int o = 1; o < G; o <<= 1) v += __shfl_xor_sync(gmask, v, o);
  unsigned ck, cl;
  if (G == 8 || c <= 4) {
    ck = __shfl_sync(gmask, sk.x, gbase + c - 1);
    cl = __shfl_sync(gmask, sl.x, gbase + c - 1);
  } else {  // G == 4, symbol N: side array
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
    constexpr int LOGW = (G == 4) ? 4 : 5;
    const int64_t w = gp >> LOGW;
    if (w != id) {
      if (w == id + dir && id >= 0) cur = nxt;
      else cur = load(seq, w, lg);
      id = w;
      const int64_t wn = w + dir;
      nxt = wn >= 0 ? load(seq, wn, lg) : 0u;  // buffer is padded by one window at the end
    }
    const unsigned v = __shfl_sync(gmask, cur, gbase + (int)((gp >> 2) & (G - 1)));
    return (int)((v >> ((gp & 3) * 8)) & 0xffu);
  }
};

template <int G>
__global__ void __launch_bounds__(256) k_sfs_search(const SearchParams P) {
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
        w = __shfl_sync(gmask, w, gbase);
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
      extend_group<G>(P, c, k, s, lg, gbase, gmask, n_blk);
      ++n_ext;
    }
  }
}

// ------------------------------------------------------------------------------ rank kernels
template <int G>
__global__ void k_rank2a(const SearchParams P, const int64_t* __restrict__ qk, const int64_t* __restrict__ ql,
                         int64_t nq, int64_t n, int64_t* __restrict__ ok6, int64_t* __restrict__ ol6) {
  const int lane = threadIdx.x & 31;
  const int lg = lane & (G - 1);
  const int gbase = lane & ~(G - 1);
  const unsigned gmask = ((1u << G) - 1u) << gbase;
  const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) / G;
  const int64_t g0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  constexpr int LOGB = (G == 4) ? 7 : 8;
  for (int64_t q = g0; q < nq; q += ngroups) {
    int64_t kk = qk[q], ll = ql[q];
    int64_t sumk = 0, suml = 0;
    for (int c = 1; c <= 5; ++c) {
      uint64_t k = (uint64_t)kk, s = (uint64_t)(ll - kk);
      unsigned nb = 0;
      extend_group<G>(P, c, k, s, lg, gbase, gmask, nb);
      int64_t okc = (int64_t)k - P.acc[c], olc = okc + (int64_t)s;
      sumk += okc; suml += olc;
      if (lg == 0) { ok6[q * 6 + c] = okc; ol6[q * 6 + c] = olc; }
    }
    if (lg == 0) { ok6[q * 6] = kk - sumk; ol6[q * 6] = ll - suml; }
    (void)LOGB; (void)n;
  }
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

// one extension per query, queries generated on the fly: k uniform in [0, n - delta)
template <int G>
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
    extend_group<G>(P, c, k, s, lg, gbase, gmask, nb);
    acc += k + s;
  }
  if (lg == 0) {
    atomicAdd(blocks_touched, (unsigned long long)nb);
    if (acc == 0x123456789ULL) atomicAdd(sink, acc);
  }
}

// ------------------------------------------------------------------------------ host side
static void fill_params(SearchParams& P, const IndexDev& d) {
  memset(&P, 0, sizeof(P));
  P.blocks = d.d_blocks;
  P.cntN = d.d_cntN;
  P.sbase = d.d_sbase;
  memcpy(P.acc, d.acc, sizeof(d.acc));
}

__global__ void k_read_lengths(const int64_t* __restrict__ offs, int64_t n, uint32_t* __restrict__ keys,
                               uint32_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int64_t l = offs[i + 1] - offs[i];
  keys[i] = ~(uint32_t)(l > 0xffffffffLL ? 0xffffffffLL : l);  // descending length
  vals[i] = (uint32_t)i;
}

static int make_order(svb_reads* R, cudaStream_t st) {
  int64_t n = R->n_reads;
  if (n == 0) return SVB_OK;
  uint32_t *k1 = nullptr, *k2 = nullptr, *v1 = nullptr;
  SVB_CUDA(cudaMalloc((void**)&k1, n * 4));
  SVB_CUDA(cudaMalloc((void**)&k2, n * 4));
  SVB_CUDA(cudaMalloc((void**)&v1, n * 4));
  SVB_CUDA(cudaMalloc((void**)&R->d_order, n * 4));
  k_read_lengths<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(R->d_offs, n, k1, v1);
  size_t bytes = 0;
  void* tmp = nullptr;
  cub::DeviceRadixSort::SortPairs(nullptr, bytes, k1, k2, v1, R->d_order, n, 0, 32, st);
  SVB_CUDA(cudaMalloc(&tmp, bytes));
  SVB_CUDA(cub::DeviceRadixSort::SortPairs(tmp, bytes, k1, k2, v1, R->d_order, n, 0, 32, st));
  SVB_CUDA(cudaStreamSynchronize(st));
  cudaFree(tmp); cudaFree(k1); cudaFree(k2); cudaFree(v1);
  return SVB_OK;
}

int check_device(int device);

}  // namespace svb

using namespace svb;

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

struct SearchScratch {
  unsigned long long* d_ctr = nullptr;  // [0] work [1] out_count [2] ext [3] blocks
  uint64_t* d_key = nullptr; uint64_t* d_key2 = nullptr;
  uint32_t* d_len = nullptr; uint32_t* d_len2 = nullptr;
  void* d_tmp = nullptr;
  ~SearchScratch() {
    cudaFree(d_ctr); cudaFree(d_key); cudaFree(d_key2); cudaFree(d_len); cudaFree(d_len2); cudaFree(d_tmp);
  }
};

// runs the search kernel (+ sort) on reads resident on the device; fills `out` host arrays
static int run_search(const IndexDev& d, const svb_reads* R, int overlap, int assemble, svb_sfs_out_t* out,
                      cudaStream_t st) {
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
  P.seq = R->d_seq; P.offs = R->d_offs; P.order = R->d_order; P.n_reads = n_reads;
  P.overlap = overlap; P.assemble = assemble;
  SearchScratch S;
  SVB_CUDA(cudaMalloc((void**)&S.d_ctr, 4 * sizeof(unsigned long long)));
  // first guess of output capacity; exact count is known after the run, rerun once if it overflowed
  unsigned long long cap = assemble ? (unsigned long long)(4 * n_reads + 1024)
                                    : (unsigned long long)(R->total / 8 + 64 * n_reads + 1024);
  int grid = 0;
  if (d.G == 4) SVB_TRY(persistent_grid(k_sfs_search<4>, 256, d.device, &grid));
  else SVB_TRY(persistent_grid(k_sfs_search<8>, 256, d.device, &grid));
  cudaEvent_t e0, e1;
  SVB_CUDA(cudaEventCreate(&e0));
  SVB_CUDA(cudaEventCreate(&e1));
  unsigned long long ctr[4] = {0, 0, 0, 0};
  float kms = 0.f;
  for (int attempt = 0; attempt < 2; ++attempt) {
    cudaFree(S.d_key); cudaFree(S.d_len); S.d_key = nullptr; S.d_len = nullptr;
    SVB_CUDA(cudaMalloc((void**)&S.d_key, cap * 8));
    SVB_CUDA(cudaMalloc((void**)&S.d_len, cap * 4));
    SVB_CUDA(cudaMemsetAsync(S.d_ctr, 0, 4 * sizeof(unsigned long long), st));
    P.work = S.d_ctr + 0; P.out_count = S.d_ctr + 1; P.stats = S.d_ctr + 2;
    P.out_key = S.d_key; P.out_len = S.d_len; P.out_cap = cap;
    SVB_CUDA(cudaEventRecord(e0, st));
    if (d.G == 4) k_sfs_search<4><<<grid, 256, 0, st>>>(P);
    else k_sfs_search<8><<<grid, 256, 0, st>>>(P);
    SVB_CUDA(cudaGetLastError());
    SVB_CUDA(cudaEventRecord(e1, st));
    SVB_CUDA(cudaMemcpyAsync(ctr, S.d_ctr, sizeof(ctr), cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    float ms = 0.f;
    SVB_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    kms += ms;
    out->launches += 1;
    if (ctr[1] <= cap) break;
    cap = ctr[1];
    if (attempt == 1) { set_error("output overflow persisted"); return SVB_ERANGE; }
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  out->kernel_ms = kms;
  out->n_ext = (int64_t)ctr[2];
  out->n_blocks_touched = (int64_t)ctr[3];
  const int64_t m = (int64_t)ctr[1];
  out->n_sfs = m;
  if (m == 0) return SVB_OK;
  // order records by (read, qs asc | emit order)
  SVB_CUDA(cudaMalloc((void**)&S.d_key2, m * 8));
  SVB_CUDA(cudaMalloc((void**)&S.d_len2, m * 4));
  cub::DoubleBuffer<uint64_t> dk(S.d_key, S.d_key2);
  cub::DoubleBuffer<uint32_t> dv(S.d_len, S.d_len2);
  int rbits = 1;
  while (rbits < 32 && ((uint64_t)n_reads >> rbits)) ++rbits;
  size_t bytes = 0;
  SVB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, m, 0, 32 + rbits, st));
  SVB_CUDA(cudaMalloc(&S.d_tmp, bytes));
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
    if (cudaMalloc((void**)&R->d_offs, (n_reads + 1) * 8) != cudaSuccess) { rc = SVB_ENOMEM; break; }
    if (mem == SVB_MEM_HOST) {
      std::vector<int64_t> rb((size_t)n_reads + 1);
      for (int64_t i = 0; i <= n_reads; ++i) {
        rb[i] = offs[i] - first;
        if (i && rb[i] < rb[i - 1]) { set_error("read offsets must be non-decreasing"); rc = SVB_EINVAL; break; }
      }
      if (rc) break;
      // one spare window (<= 32 B) after the last base, rounded to 64 B
      size_t padded = ((size_t)R->total + 64 + 63) & ~(size_t)63;
      if (cudaMalloc((void**)&R->d_seq, padded) != cudaSuccess) { rc = SVB_ENOMEM; break; }
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
    rc = make_order(R, 0);
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
  if (R->owns_seq && R->d_seq) cudaFree(R->d_seq);
  if (R->d_offs) cudaFree(R->d_offs);
  if (R->d_order) cudaFree(R->d_order);
  delete R;
}

int svb_sfs_resident(const svb_index_t* idx, const svb_reads_t* reads, int overlap, int assemble,
                     svb_sfs_out_t* out) {
  if (!idx || !reads || !out) { set_error("svb_sfs_resident: null argument"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  SVB_TRY(check_device(idx->dev.device));
  if (reads->device != idx->dev.device) { set_error("reads and index live on different devices"); return SVB_EINVAL; }
  cudaEvent_t e0, e1;
  SVB_CUDA(cudaEventCreate(&e0));
  SVB_CUDA(cudaEventCreate(&e1));
  SVB_CUDA(cudaEventRecord(e0, 0));
  int rc = run_search(idx->dev, reads, overlap, assemble, out, 0);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&out->device_ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc != SVB_OK) svb_sfs_out_free(out);
  return rc;
}

int svb_sfs_batch(const svb_index_t* idx, const uint8_t* seq, const int64_t* offs, int64_t n_reads,
                  int overlap, int assemble, svb_sfs_out_t* out) {
  if (!idx || !offs || !out || n_reads < 0) { set_error("svb_sfs_batch: bad arguments"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  SVB_TRY(check_device(idx->dev.device));
  cudaEvent_t e0, e1;
  SVB_CUDA(cudaEventCreate(&e0));
  SVB_CUDA(cudaEventCreate(&e1));
  SVB_CUDA(cudaEventRecord(e0, 0));
  svb_reads_t* R = nullptr;
  int rc = svb_reads_upload(seq, offs, n_reads, SVB_MEM_HOST, idx->dev.device, &R);
  if (rc == SVB_OK) {
    rc = run_search(idx->dev, R, overlap, assemble, out, 0);
    out->h2d_bytes = R->total + (n_reads + 1) * 8;
    out->launches += 2;  // length keys + order sort
  }
  svb_reads_free(R);
  cudaEventRecord(e1, 0);
  cudaEventSynchronize(e1);
  cudaEventElapsedTime(&out->device_ms, e0, e1);
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (rc != SVB_OK) svb_sfs_out_free(out);
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
  unsigned grid = (unsigned)((groups * d.G + 255) / 256);
  if (d.G == 4) k_rank2a<4><<<grid, 256>>>(P, dk, dl, n, d.n, dok, dol);
  else k_rank2a<8><<<grid, 256>>>(P, dk, dl, n, d.n, dok, dol);
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
  int grid = 0;
  if (d.G == 4) SVB_TRY(persistent_grid(k_rank_bench<4>, 256, d.device, &grid));
  else SVB_TRY(persistent_grid(k_rank_bench<8>, 256, d.device, &grid));
  cudaEvent_t e0, e1;
  SVB_CUDA(cudaEventCreate(&e0));
  SVB_CUDA(cudaEventCreate(&e1));
  // one untimed warm-up launch, then `iters` timed launches with distinct seeds
  for (int it = -1; it < iters; ++it) {
    if (it == 0) { SVB_CUDA(cudaMemset(dctr, 0, 16)); SVB_CUDA(cudaEventRecord(e0, 0)); }
    uint64_t sd = seed + 0x1000003ULL * (uint64_t)(it + 1);
    if (d.G == 4) k_rank_bench<4><<<grid, 256>>>(P, nq, d.n, delta, sd, dctr, dctr + 1);
    else k_rank_bench<8><<<grid, 256>>>(P, nq, d.n, delta, sd, dctr, dctr + 1);
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
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(dctr);
  return SVB_OK;
}

}  // extern "C"
