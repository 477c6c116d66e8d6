// FMD index construction on the GPU: text -> suffix array -> BWT -> sampled-Occ block array.
//
// Stands in for `SVDSS index` (= ropebwt3 main_build, reference main.cpp:15-17,34-37) and for the
// in-memory index that rb3_fmi_restore hands to PingPong::search (ping_pong.cpp:244-245).
// The text model is ropebwt3's default (both strands):  T = S_0 $ rc(S_0) $ S_1 $ ...
// Sentinels are made distinct by text position; any fixed sentinel order gives the same answers to
// '$'-free queries, which are the only ones the search path issues (ping_pong.cpp:12-36).
//
// Suffix sorting (all on the device, CUB for the radix passes):
//   1. partition suffixes by their first K symbols into bucket groups that fit the sort scratch
//   2. per group: 63-bit key = next 21 symbols (3 bits each, zero-filled after a '$'), stable LSD
//      radix sort of (key, position); equal keys containing a '$' are already in final order
//      (stable sort over ascending positions), everything else forms h=21 groups
//   3. prefix doubling restricted to unresolved groups: key = (dense group id, ISA[pos+h])
#include <cub/cub.cuh>

#include <algorithm>
#include <cstdlib>
#include <cstdarg>
#include <cstring>

#include "common.cuh"
#include "../host/rld.hpp"   // ropebwt3 .fmd reader + BWT inversion (host code, header only)

namespace svb {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* last_error() { return g_err; }

// ------------------------------------------------------------------------------ small helpers
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return cudaMalloc((void**)&p, count * sizeof(T));
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
  }
};

static inline unsigned grid_for(int64_t n, int threads, int64_t per_thread = 1) {
  int64_t b = (n + (int64_t)threads * per_thread - 1) / ((int64_t)threads * per_thread);
  if (b < 1) b = 1;
  if (b > 0x7fffffff) b = 0x7fffffff;
  return (unsigned)b;
}

__host__ __device__ inline uint8_t comp6(uint8_t c) { return (c >= 1 && c <= 4) ? (uint8_t)(5 - c) : c; }

// ------------------------------------------------------------------------------ text
// one thread per output symbol; contig of an output position found by binary search on tstart[]
__global__ void k_build_text(const uint8_t* __restrict__ seqs, const int64_t* __restrict__ offs,
                             int64_t m, int64_t n, uint8_t* __restrict__ T) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    // text start of contig r: 2*(offs[r]-offs[0]) + 2r
    int64_t lo = 0, hi = m - 1;
    int64_t o0 = offs[0];
    while (lo < hi) {
      int64_t mid = (lo + hi + 1) >> 1;
      int64_t ts = 2 * (offs[mid] - o0) + 2 * mid;
      if (ts <= i) lo = mid; else hi = mid - 1;
    }
    int64_t b = offs[lo], e = offs[lo + 1], L = e - b;
    int64_t j = i - (2 * (b - o0) + 2 * lo);
    uint8_t v;
    if (j < L) v = seqs[b + j];
    else if (j == L) v = 0;
    else if (j < 2 * L + 1) v = comp6(seqs[e - 1 - (j - L - 1)]);
    else v = 0;
    T[i] = v;
  }
}

// ------------------------------------------------------------------------------ K-mer partition
// code of the first K symbols in base 6, symbols after the first '$' count as 0
__device__ inline uint32_t kmer_code(const uint8_t* __restrict__ T, int64_t n, int64_t i, int K) {
  uint32_t code = 0;
  bool dead = false;
  for (int j = 0; j < K; ++j) {
    uint8_t c = 0;
    if (!dead && i + j < n) {
      c = T[i + j];
      if (c == 0) dead = true;
    }
    code = code * 6 + c;
  }
  return code;
}

__global__ void k_kmer_hist(const uint8_t* __restrict__ T, int64_t n, int K, int nbins,
                            unsigned long long* __restrict__ hist) {
  extern __shared__ unsigned int sh[];
  bool use_sh = nbins <= 4096;
  if (use_sh) {
    for (int b = threadIdx.x; b < nbins; b += blockDim.x) sh[b] = 0;
    __syncthreads();
  }
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint32_t c = kmer_code(T, n, i, K);
    if (use_sh) atomicAdd(&sh[c], 1u);
    else atomicAdd(&hist[c], 1ull);
  }
  if (use_sh) {
    __syncthreads();
    for (int b = threadIdx.x; b < nbins; b += blockDim.x)
      if (sh[b]) atomicAdd(&hist[b], (unsigned long long)sh[b]);
  }
}

struct InCodeRange {
  const uint8_t* T;
  int64_t n;
  int K;
  uint32_t lo, hi;
  __device__ bool operator()(int64_t i) const {
    if (K == 0) return true;
    uint32_t c = kmer_code(T, n, i, K);
    return c >= lo && c <= hi;
  }
};

// 21 symbols x 3 bits, first symbol most significant, zero-filled after '$'
__global__ void k_make_keys(const uint8_t* __restrict__ T, int64_t n, const uint64_t* __restrict__ pos,
                            int64_t m, uint64_t* __restrict__ keys) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  int64_t p = (int64_t)pos[j];
  uint64_t key = 0;
  bool dead = false;
#pragma unroll
  for (int t = 0; t < 21; ++t) {
    uint64_t c = 0;
    if (!dead && p + t < n) {
      c = T[p + t];
      if (c == 0) dead = true;
    }
    key = (key << 3) | c;
  }
  keys[j] = key;
}

__device__ inline bool key_has_dollar(uint64_t key) {
  // any of the 21 3-bit fields zero?
  uint64_t x = key | (key >> 1) | (key >> 2);
  return (x & 0x1249249249249249ULL) != 0x1249249249249249ULL;
}

// head flag + (head ? j : 0) for the max-scan
__global__ void k_first_heads(const uint64_t* __restrict__ keys, int64_t m, uint8_t* __restrict__ head,
                              uint32_t* __restrict__ headidx) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  uint64_t k = keys[j];
  bool h = (j == 0) || (k != keys[j - 1]) || key_has_dollar(k);
  head[j] = h;
  headidx[j] = h ? (uint32_t)j : 0u;
}

// writes SA and ISA for one bucket group, flags unresolved elements
__global__ void k_first_commit(const uint64_t* __restrict__ pos, const uint8_t* __restrict__ head,
                               const uint32_t* __restrict__ headidx_scanned, int64_t m, int64_t off,
                               uint64_t* __restrict__ SA, uint64_t* __restrict__ ISA,
                               uint8_t* __restrict__ unresolved) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  uint64_t p = pos[j];
  SA[off + j] = p;
  ISA[p] = (uint64_t)off + headidx_scanned[j];
  bool single = head[j] && (j == m - 1 || head[j + 1]);
  unresolved[j] = !single;
}

__global__ void k_gather_unres(const uint8_t* __restrict__ unresolved, const uint32_t* __restrict__ headidx_scanned,
                               const uint32_t* __restrict__ slot /*exclusive scan of unresolved*/,
                               int64_t m, int64_t off, uint64_t* __restrict__ U, uint64_t* __restrict__ Ug) {
  int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= m) return;
  if (unresolved[j]) {
    uint32_t s = slot[j];
    U[s] = (uint64_t)(off + j);
    Ug[s] = (uint64_t)off + headidx_scanned[j];
  }
}

// ------------------------------------------------------------------------------ doubling
// gid: dense group id within U (groups are runs of equal Ug)
__global__ void k_group_flags(const uint64_t* __restrict__ Ug, int64_t u, uint32_t* __restrict__ flag) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= u) return;
  flag[i] = (i == 0 || Ug[i] != Ug[i - 1]) ? 1u : 0u;
}

__global__ void k_doubling_keys(const uint64_t* __restrict__ U, const uint32_t* __restrict__ gid_incl,
                                const uint64_t* __restrict__ SA, const uint64_t* __restrict__ ISA,
                                int64_t u, int64_t h, int64_t n, uint64_t* __restrict__ keys,
                                uint64_t* __restrict__ vals) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= u) return;
  uint64_t p = SA[U[i]];
  uint64_t q = p + (uint64_t)h;
  uint64_t r2 = q < (uint64_t)n ? ISA[q] : 0;  // q < n always holds for unresolved suffixes
  keys[i] = ((uint64_t)(gid_incl[i] - 1) << 33) | r2;
  vals[i] = p;
}

__global__ void k_doubling_heads(const uint64_t* __restrict__ keys, int64_t u, uint8_t* __restrict__ head,
                                 uint32_t* __restrict__ headidx) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= u) return;
  bool h = (i == 0) || keys[i] != keys[i - 1];
  head[i] = h;
  headidx[i] = h ? (uint32_t)i : 0u;
}

__global__ void k_doubling_commit(const uint64_t* __restrict__ U, const uint64_t* __restrict__ vals,
                                  const uint8_t* __restrict__ head, const uint32_t* __restrict__ headidx_scanned,
                                  int64_t u, uint64_t* __restrict__ SA, uint64_t* __restrict__ ISA,
                                  uint64_t* __restrict__ Ug_new, uint8_t* __restrict__ unresolved) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= u) return;
  uint64_t p = vals[i];
  uint64_t rank = U[headidx_scanned[i]];
  SA[U[i]] = p;
  ISA[p] = rank;
  Ug_new[i] = rank;
  bool single = head[i] && (i == u - 1 || head[i + 1]);
  unresolved[i] = !single;
}

__global__ void k_compact2(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b,
                           const uint8_t* __restrict__ keep, const uint32_t* __restrict__ slot, int64_t u,
                           uint64_t* __restrict__ oa, uint64_t* __restrict__ ob) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= u) return;
  if (keep[i]) {
    uint32_t s = slot[i];
    oa[s] = a[i];
    ob[s] = b[i];
  }
}

struct U8ToU32 {
  __host__ __device__ uint32_t operator()(uint8_t v) const { return v; }
};

static int exclusive_count(const uint8_t* d_flags, int64_t m, uint32_t* d_slot, DevBuf<uint8_t>& tmp,
                           int64_t* total, cudaStream_t st) {
  // slot = exclusive prefix sum of flags (as u32); total = number of set flags
  cub::TransformInputIterator<uint32_t, U8ToU32, const uint8_t*> it(d_flags, U8ToU32());
  size_t bytes = 0;
  SVB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, d_slot, m, st));
  if (bytes > tmp.n) SVB_CUDA(tmp.alloc(bytes + (bytes >> 2)));
  SVB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, it, d_slot, m, st));
  uint32_t last_slot = 0;
  uint8_t last_flag = 0;
  SVB_CUDA(cudaMemcpyAsync(&last_slot, d_slot + (m - 1), 4, cudaMemcpyDeviceToHost, st));
  SVB_CUDA(cudaMemcpyAsync(&last_flag, d_flags + (m - 1), 1, cudaMemcpyDeviceToHost, st));
  SVB_CUDA(cudaStreamSynchronize(st));
  *total = (int64_t)last_slot + (last_flag ? 1 : 0);
  return SVB_OK;
}

static int scan_max_u32(uint32_t* d, int64_t m, DevBuf<uint8_t>& tmp, cudaStream_t st) {
  size_t bytes = 0;
  SVB_CUDA(cub::DeviceScan::InclusiveScan(nullptr, bytes, d, d, cub::Max(), m, st));
  if (bytes > tmp.n) SVB_CUDA(tmp.alloc(bytes + (bytes >> 2)));
  SVB_CUDA(cub::DeviceScan::InclusiveScan(tmp.p, bytes, d, d, cub::Max(), m, st));
  return SVB_OK;
}

static int scan_sum_u32(uint32_t* d, int64_t m, DevBuf<uint8_t>& tmp, cudaStream_t st) {
  size_t bytes = 0;
  SVB_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, d, d, m, st));
  if (bytes > tmp.n) SVB_CUDA(tmp.alloc(bytes + (bytes >> 2)));
  SVB_CUDA(cub::DeviceScan::InclusiveSum(tmp.p, bytes, d, d, m, st));
  return SVB_OK;
}

static int sort_pairs_u64(uint64_t*& keys, uint64_t*& vals, uint64_t* keys_alt, uint64_t* vals_alt,
                          int64_t m, int end_bit, DevBuf<uint8_t>& tmp, cudaStream_t st) {
  cub::DoubleBuffer<uint64_t> dk(keys, keys_alt), dv(vals, vals_alt);
  size_t bytes = 0;
  SVB_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, bytes, dk, dv, m, 0, end_bit, st));
  if (bytes > tmp.n) SVB_CUDA(tmp.alloc(bytes + (bytes >> 2)));
  SVB_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, bytes, dk, dv, m, 0, end_bit, st));
  keys = dk.Current();
  vals = dv.Current();
  return SVB_OK;
}

static int bits_for(uint64_t v) {
  int b = 1;
  while (b < 64 && (v >> b)) ++b;
  return b;
}

// Suffix array of device text T[0..n), T[n-1] == 0. d_SA receives n u64.
int build_suffix_array(const uint8_t* d_T, int64_t n, uint64_t* d_SA, cudaStream_t st) {
  if (n <= 0) return SVB_OK;
  if (n > (1LL << 33)) {
    set_error("text of %lld symbols exceeds the 2^33 limit of the suffix sorter", (long long)n);
    return SVB_ERANGE;
  }
  const int TPB = 256;
  DevBuf<uint64_t> ISA;
  SVB_CUDA(ISA.alloc((size_t)n));
  DevBuf<uint8_t> tmp;

  // ---- 1. choose K and bucket groups
  size_t free_b = 0, total_b = 0;
  SVB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  // scratch per element of a group: pos,key x2 (double buffers) + head/unres + headidx/slot + CUB temp
  int64_t cap = (int64_t)std::min<size_t>((size_t)1 << 28, std::max<size_t>(free_b / 2 / 48, (size_t)1 << 16));
  int K = 0;
  std::vector<unsigned long long> hist(1, (unsigned long long)n);
  DevBuf<unsigned long long> d_hist;
  while (true) {
    unsigned long long mx = *std::max_element(hist.begin(), hist.end());
    if ((int64_t)mx <= cap || K >= 6) break;
    ++K;
    int nbins = 1;
    for (int j = 0; j < K; ++j) nbins *= 6;
    SVB_CUDA(d_hist.alloc(nbins));
    SVB_CUDA(cudaMemsetAsync(d_hist.p, 0, sizeof(unsigned long long) * nbins, st));
    size_t sh = nbins <= 4096 ? sizeof(unsigned int) * nbins : 0;
    k_kmer_hist<<<148 * 8, TPB, sh, st>>>(d_T, n, K, nbins, d_hist.p);
    SVB_CUDA(cudaGetLastError());
    hist.resize(nbins);
    SVB_CUDA(cudaMemcpyAsync(hist.data(), d_hist.p, sizeof(unsigned long long) * nbins, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
  }
  struct Group { uint32_t lo, hi; int64_t m, off; };
  std::vector<Group> groups;
  {
    int64_t off = 0;
    size_t b = 0;
    while (b < hist.size()) {
      if (hist[b] == 0) { ++b; continue; }
      Group g{(uint32_t)b, (uint32_t)b, (int64_t)hist[b], off};
      size_t e = b + 1;
      while (e < hist.size() && g.m + (int64_t)hist[e] <= cap) { g.m += (int64_t)hist[e]; g.hi = (uint32_t)e; ++e; }
      // trailing empty bins may be skipped freely
      groups.push_back(g);
      off += g.m;
      b = e;
    }
  }
  int64_t max_m = 0;
  for (auto& g : groups) max_m = std::max(max_m, g.m);
  if (max_m >= (1LL << 32)) {
    set_error("suffix bucket of %lld elements is too large", (long long)max_m);
    return SVB_ERANGE;
  }

  // ---- 2. first pass per group
  DevBuf<uint64_t> pos_a, pos_b, key_a, key_b;
  DevBuf<uint8_t> head, unres;
  DevBuf<uint32_t> hidx, slot;
  DevBuf<int64_t> d_nsel;
  SVB_CUDA(pos_a.alloc(max_m)); SVB_CUDA(pos_b.alloc(max_m));
  SVB_CUDA(key_a.alloc(max_m)); SVB_CUDA(key_b.alloc(max_m));
  SVB_CUDA(head.alloc(max_m)); SVB_CUDA(unres.alloc(max_m));
  SVB_CUDA(hidx.alloc(max_m)); SVB_CUDA(slot.alloc(max_m));
  SVB_CUDA(d_nsel.alloc(1));

  // unresolved list, grown geometrically
  DevBuf<uint64_t> U, Ug;
  int64_t u = 0;
  auto ensure_u = [&](int64_t need) -> int {
    if ((size_t)need <= U.n) return SVB_OK;
    size_t ncap = std::max<size_t>((size_t)need, U.n * 2 + 1024);
    DevBuf<uint64_t> nU, nUg;
    SVB_CUDA(nU.alloc(ncap)); SVB_CUDA(nUg.alloc(ncap));
    if (u) {
      SVB_CUDA(cudaMemcpyAsync(nU.p, U.p, u * 8, cudaMemcpyDeviceToDevice, st));
      SVB_CUDA(cudaMemcpyAsync(nUg.p, Ug.p, u * 8, cudaMemcpyDeviceToDevice, st));
      SVB_CUDA(cudaStreamSynchronize(st));
    }
    std::swap(U.p, nU.p); std::swap(U.n, nU.n);
    std::swap(Ug.p, nUg.p); std::swap(Ug.n, nUg.n);
    return SVB_OK;
  };

  const int64_t CHUNK = 1LL << 30;  // DeviceSelect item-count safety
  for (auto& g : groups) {
    // stable selection of the group's suffix positions in ascending order
    int64_t got = 0;
    InCodeRange pred{d_T, n, K, g.lo, g.hi};
    for (int64_t c0 = 0; c0 < n; c0 += CHUNK) {
      int64_t cn = std::min(CHUNK, n - c0);
      cub::CountingInputIterator<int64_t> it(c0);
      size_t bytes = 0;
      SVB_CUDA(cub::DeviceSelect::If(nullptr, bytes, it, (int64_t*)pos_a.p + got, d_nsel.p, (int)cn, pred, st));
      if (bytes > tmp.n) SVB_CUDA(tmp.alloc(bytes + (bytes >> 2)));
      SVB_CUDA(cub::DeviceSelect::If(tmp.p, bytes, it, (int64_t*)pos_a.p + got, d_nsel.p, (int)cn, pred, st));
      int64_t ns = 0;
      SVB_CUDA(cudaMemcpyAsync(&ns, d_nsel.p, 8, cudaMemcpyDeviceToHost, st));
      SVB_CUDA(cudaStreamSynchronize(st));
      got += ns;
    }
    if (got != g.m) {
      set_error("suffix partition mismatch: selected %lld, histogram says %lld", (long long)got, (long long)g.m);
      return SVB_EINVAL;
    }
    int64_t m = g.m;
    k_make_keys<<<grid_for(m, TPB), TPB, 0, st>>>(d_T, n, pos_a.p, m, key_a.p);
    SVB_CUDA(cudaGetLastError());
    uint64_t *kc = key_a.p, *vc = pos_a.p;
    SVB_TRY(sort_pairs_u64(kc, vc, kc == key_a.p ? key_b.p : key_a.p, vc == pos_a.p ? pos_b.p : pos_a.p, m, 63, tmp, st));
    k_first_heads<<<grid_for(m, TPB), TPB, 0, st>>>(kc, m, head.p, hidx.p);
    SVB_CUDA(cudaGetLastError());
    SVB_TRY(scan_max_u32(hidx.p, m, tmp, st));
    k_first_commit<<<grid_for(m, TPB), TPB, 0, st>>>(vc, head.p, hidx.p, m, g.off, d_SA, ISA.p, unres.p);
    SVB_CUDA(cudaGetLastError());
    int64_t nun = 0;
    SVB_TRY(exclusive_count(unres.p, m, slot.p, tmp, &nun, st));
    if (nun) {
      SVB_TRY(ensure_u(u + nun));
      k_gather_unres<<<grid_for(m, TPB), TPB, 0, st>>>(unres.p, hidx.p, slot.p, m, g.off, U.p + u, Ug.p + u);
      SVB_CUDA(cudaGetLastError());
      u += nun;
    }
  }
  pos_a.release(); pos_b.release(); key_a.release(); key_b.release();
  head.release(); unres.release(); hidx.release(); slot.release();

  // ---- 3. prefix doubling on the unresolved set
  int64_t h = 21;
  const bool stats = getenv("SVB_INDEX_STATS") != nullptr;
  if (stats) fprintf(stderr, "[index] first pass (21-symbol keys): %lld of %lld suffixes unresolved (limit 2^31 = %lld)\n", (long long)u, (long long)n, 1LL << 31);
  if (u >= (1LL << 31)) {
    set_error("%lld unresolved suffixes after the first pass exceed the 2^31 limit", (long long)u);
    return SVB_ERANGE;
  }
  if (u > 0) {
    DevBuf<uint64_t> ka, kb, va, vb, U2, Ug2;
    DevBuf<uint32_t> gid, hidx2, slot2;
    DevBuf<uint8_t> head2, unres2;
    SVB_CUDA(ka.alloc(u)); SVB_CUDA(kb.alloc(u)); SVB_CUDA(va.alloc(u)); SVB_CUDA(vb.alloc(u));
    SVB_CUDA(U2.alloc(u)); SVB_CUDA(Ug2.alloc(u));
    SVB_CUDA(gid.alloc(u)); SVB_CUDA(hidx2.alloc(u)); SVB_CUDA(slot2.alloc(u));
    SVB_CUDA(head2.alloc(u)); SVB_CUDA(unres2.alloc(u));
    uint64_t *pU = U.p, *pUg = Ug.p, *pU2 = U2.p, *pUg2 = Ug2.p;
    int iter = 0;
    while (u > 0) {
      if (++iter > 64) { set_error("prefix doubling did not converge"); return SVB_EINVAL; }
      unsigned gr = grid_for(u, TPB);
      k_group_flags<<<gr, TPB, 0, st>>>(pUg, u, gid.p);
      SVB_CUDA(cudaGetLastError());
      SVB_TRY(scan_sum_u32(gid.p, u, tmp, st));
      uint32_t ngroups = 0;
      SVB_CUDA(cudaMemcpyAsync(&ngroups, gid.p + (u - 1), 4, cudaMemcpyDeviceToHost, st));
      SVB_CUDA(cudaStreamSynchronize(st));
      k_doubling_keys<<<gr, TPB, 0, st>>>(pU, gid.p, d_SA, ISA.p, u, h, n, ka.p, va.p);
      SVB_CUDA(cudaGetLastError());
      uint64_t *kc = ka.p, *vc = va.p;
      int end_bit = std::min(64, 33 + bits_for(ngroups));
      SVB_TRY(sort_pairs_u64(kc, vc, kb.p, vb.p, u, end_bit, tmp, st));
      k_doubling_heads<<<gr, TPB, 0, st>>>(kc, u, head2.p, hidx2.p);
      SVB_CUDA(cudaGetLastError());
      SVB_TRY(scan_max_u32(hidx2.p, u, tmp, st));
      k_doubling_commit<<<gr, TPB, 0, st>>>(pU, vc, head2.p, hidx2.p, u, d_SA, ISA.p, pUg2 /*new ranks*/, unres2.p);
      SVB_CUDA(cudaGetLastError());
      int64_t nun = 0;
      SVB_TRY(exclusive_count(unres2.p, u, slot2.p, tmp, &nun, st));
      if (nun) {
        // compact (U, new ranks) -> (U2', Ug')   [pUg2 holds new ranks; write into pUg / pU2]
        k_compact2<<<gr, TPB, 0, st>>>(pU, pUg2, unres2.p, slot2.p, u, pU2, pUg);
        SVB_CUDA(cudaGetLastError());
        std::swap(pU, pU2);
        // pUg now holds compacted ranks
      }
      if (stats) fprintf(stderr, "[index] doubling round %d (h = %lld): %lld -> %lld unresolved, %u groups\n", iter, (long long)h, (long long)u, (long long)nun, ngroups);
      u = nun;
      h *= 2;
    }
  }
  SVB_CUDA(cudaStreamSynchronize(st));
  return SVB_OK;
}

// ------------------------------------------------------------------------------ block array
// one warp builds one 32-symbol slice per lane-iteration; symbol source is either a BWT array or
// (SA, T) so the BWT is never materialised for big texts.
template <int G>
__global__ void k_block_planes(const uint8_t* __restrict__ bwt, const uint64_t* __restrict__ SA,
                               const uint8_t* __restrict__ T, int64_t n, int64_t n_blocks,
                               uint4* __restrict__ blocks, uint32_t* __restrict__ blkcnt /* [6][n_blocks] */) {
  // one thread per slice: 32 symbols
  int64_t total_slices = n_blocks * G;
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t sidx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; sidx < total_slices; sidx += stride) {
    int64_t base = sidx * 32;
    uint32_t p0 = 0, p1 = 0, p2 = 0;
    uint32_t c[6] = {0, 0, 0, 0, 0, 0};
    for (int j = 0; j < 32; ++j) {
      int64_t i = base + j;
      uint32_t s = 7;
      if (i < n) {
        if (bwt) s = bwt[i];
        else { uint64_t p = SA[i]; s = T[p ? p - 1 : n - 1]; }
#pragma unroll
        for (int q = 0; q < 6; ++q) c[q] += (s == (uint32_t)q);
      }
      p0 |= (s & 1u) << j;
      p1 |= ((s >> 1) & 1u) << j;
      p2 |= ((s >> 2) & 1u) << j;
    }
    blocks[sidx] = make_uint4(0u, p0, p1, p2);
    int64_t b = sidx / G;
#pragma unroll
    for (int q = 0; q < 6; ++q)
      if (c[q]) atomicAdd(&blkcnt[(int64_t)q * n_blocks + b], c[q]);
  }
}

struct U32ToU64 {
  __host__ __device__ uint64_t operator()(uint32_t v) const { return v; }
};

template <int G>
__global__ void k_block_headers(uint4* __restrict__ blocks, const uint64_t* __restrict__ occ /* [6][n_blocks] exclusive */,
                                int64_t n_blocks, uint32_t* __restrict__ cntN) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int LOGB = (G == 4) ? 7 : 8;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < n_blocks; b += stride) {
    // superblock of this block: (b << LOGB) >> 32 ; its first block:
    int64_t sb = (b << LOGB) >> 32;
    int64_t b0 = (sb << 32) >> LOGB;
    // slots: A,C,G,T (,N,$ for G == 8)
#pragma unroll
    for (int j = 0; j < G; ++j) {
      uint32_t v = 0;
      int sym = (j < 4) ? j + 1 : (j == 4 ? 5 : (j == 5 ? 0 : -1));
      if (sym >= 0) v = (uint32_t)(occ[(int64_t)sym * n_blocks + b] - occ[(int64_t)sym * n_blocks + b0]);
      reinterpret_cast<uint32_t*>(&blocks[b * G + j])[0] = v;
    }
    if (G == 4) cntN[b] = (uint32_t)(occ[5LL * n_blocks + b] - occ[5LL * n_blocks + b0]);
    if (G == 8) {
      // in-block sampling for the thread-per-read kernel: slots 5,6,7 hold, one byte per symbol
      // A,C,G,T, the number of occurrences before symbol 64, 128, 192 of the block (<= 192), so a
      // rank needs the planes of ONE 64-symbol sub-block instead of the whole block.
      // (The '$' count that used to sit in slot 5 is never read: rank2a derives it.)
      unsigned cum[4] = {0, 0, 0, 0};
      for (int sub = 0; sub < 3; ++sub) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const uint4 sl = blocks[b * G + 2 * sub + h];
#pragma unroll
          for (int sym = 1; sym <= 4; ++sym) {
            const unsigned c0 = (sym & 1) ? 0u : ~0u, c1 = (sym & 2) ? 0u : ~0u, c2 = (sym & 4) ? 0u : ~0u;
            cum[sym - 1] += __popc((sl.y ^ c0) & (sl.z ^ c1) & (sl.w ^ c2));
          }
        }
        reinterpret_cast<uint32_t*>(&blocks[b * G + 5 + sub])[0] = cum[0] | (cum[1] << 8) | (cum[2] << 16) | (cum[3] << 24);
      }
    }
  }
}

__global__ void k_sbase(const uint64_t* __restrict__ occ, int64_t n_blocks, int logb, int n_sb,
                        const int64_t* __restrict__ acc, int64_t* __restrict__ sbase) {
  int t = threadIdx.x;
  if (t >= n_sb * 8) return;
  int sb = t >> 3, c = t & 7;
  int64_t v = 0;
  if (c < 6) {
    int64_t b0 = ((int64_t)sb << 32) >> logb;
    v = acc[c] + (b0 < n_blocks ? (int64_t)occ[(int64_t)c * n_blocks + b0] : 0);
  }
  sbase[t] = v;
}

// builds idx->d_blocks etc. from either a device BWT or (SA, T)
int build_blocks(IndexDev* idx, const uint8_t* d_bwt, const uint64_t* d_SA, const uint8_t* d_T,
                 cudaStream_t st) {
  const int G = idx->G;
  const int logb = (G == 4) ? 7 : 8;
  int64_t n = idx->n;
  int64_t n_blocks = (n >> logb) + 1;  // always one block past position n
  idx->n_blocks = n_blocks;
  idx->n_sb = (int)(n >> 32) + 1;
  if (idx->n_sb * 8 > 1024) { set_error("too many superblocks"); return SVB_ERANGE; }
  SVB_CUDA(cudaMalloc((void**)&idx->d_blocks, (size_t)n_blocks * G * sizeof(uint4)));
  if (G == 4) SVB_CUDA(cudaMalloc((void**)&idx->d_cntN, (size_t)n_blocks * 4));
  SVB_CUDA(cudaMalloc((void**)&idx->d_sbase, (size_t)idx->n_sb * 8 * sizeof(int64_t)));
  DevBuf<uint32_t> blkcnt;
  DevBuf<uint64_t> occ;
  DevBuf<uint8_t> tmp;
  SVB_CUDA(blkcnt.alloc((size_t)6 * n_blocks));
  SVB_CUDA(occ.alloc((size_t)6 * n_blocks));
  SVB_CUDA(cudaMemsetAsync(blkcnt.p, 0, (size_t)6 * n_blocks * 4, st));
  if (G == 4) k_block_planes<4><<<grid_for(n_blocks * G, 256), 256, 0, st>>>(d_bwt, d_SA, d_T, n, n_blocks, idx->d_blocks, blkcnt.p);
  else k_block_planes<8><<<grid_for(n_blocks * G, 256), 256, 0, st>>>(d_bwt, d_SA, d_T, n, n_blocks, idx->d_blocks, blkcnt.p);
  SVB_CUDA(cudaGetLastError());
  uint64_t totals[6];
  for (int c = 0; c < 6; ++c) {
    cub::TransformInputIterator<uint64_t, U32ToU64, const uint32_t*> it(blkcnt.p + (size_t)c * n_blocks, U32ToU64());
    size_t bytes = 0;
    SVB_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, it, occ.p + (size_t)c * n_blocks, n_blocks, st));
    if (bytes > tmp.n) SVB_CUDA(tmp.alloc(bytes + (bytes >> 2)));
    SVB_CUDA(cub::DeviceScan::ExclusiveSum(tmp.p, bytes, it, occ.p + (size_t)c * n_blocks, n_blocks, st));
    uint64_t last_occ = 0;
    uint32_t last_cnt = 0;
    SVB_CUDA(cudaMemcpyAsync(&last_occ, occ.p + (size_t)c * n_blocks + (n_blocks - 1), 8, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaMemcpyAsync(&last_cnt, blkcnt.p + (size_t)c * n_blocks + (n_blocks - 1), 4, cudaMemcpyDeviceToHost, st));
    SVB_CUDA(cudaStreamSynchronize(st));
    totals[c] = last_occ + last_cnt;
  }
  idx->acc[0] = 0;
  for (int c = 0; c < 6; ++c) idx->acc[c + 1] = idx->acc[c] + (int64_t)totals[c];
  if (idx->acc[6] != n) { set_error("symbol totals %lld != n %lld", (long long)idx->acc[6], (long long)n); return SVB_EINVAL; }
  DevBuf<int64_t> d_acc;
  SVB_CUDA(d_acc.alloc(7));
  SVB_CUDA(cudaMemcpyAsync(d_acc.p, idx->acc, 7 * 8, cudaMemcpyHostToDevice, st));
  if (G == 4) k_block_headers<4><<<grid_for(n_blocks, 256), 256, 0, st>>>(idx->d_blocks, occ.p, n_blocks, idx->d_cntN);
  else k_block_headers<8><<<grid_for(n_blocks, 256), 256, 0, st>>>(idx->d_blocks, occ.p, n_blocks, nullptr);
  SVB_CUDA(cudaGetLastError());
  k_sbase<<<1, 1024, 0, st>>>(occ.p, n_blocks, logb, idx->n_sb, d_acc.p, idx->d_sbase);
  SVB_CUDA(cudaGetLastError());
  SVB_CUDA(cudaStreamSynchronize(st));
  return SVB_OK;
}

// ---- located-match tables (IndexDev::d_text / d_ssa / d_tstart)
__global__ void k_text_device(const uint8_t* __restrict__ T, int64_t n, uint8_t* __restrict__ out /* n + 2*TEXT_PAD */) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, tot = n + 2 * TEXT_PAD;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += stride) {
    const int64_t j = i - TEXT_PAD;
    uint8_t v = TEXT_SENTINEL;
    if (j >= 0 && j < n) { v = T[j]; if (v == 0) v = TEXT_SENTINEL; }
    out[i] = v;
  }
}
__global__ void k_sample_sa(const uint64_t* __restrict__ SA, int64_t n_ssa, int ss_log, uint64_t* __restrict__ ssa) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_ssa; i += stride) ssa[i] = SA[i << ss_log];
}
// forward strands of all contigs, two nt6 codes per byte (low nibble first): the on-disk form of T
__global__ void k_pack_fwd(const uint8_t* __restrict__ text /* d_text */, const int64_t* __restrict__ tstart, int64_t m,
                           int64_t tot, uint8_t* __restrict__ packed) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x, nb = (tot + 1) / 2;
  for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < nb; b += stride) {
    uint8_t v = 0;
    for (int h = 0; h < 2; ++h) {
      const int64_t i = 2 * b + h;
      if (i >= tot) break;
      // contig r holds forward bases [ (tstart[r]-2r)/2, (tstart[r+1]-2(r+1))/2 )
      int64_t lo = 0, hi = m - 1;
      while (lo < hi) { const int64_t mid = (lo + hi + 1) >> 1; if ((tstart[mid] - 2 * mid) / 2 <= i) lo = mid; else hi = mid - 1; }
      const uint8_t c = text[tstart[lo] + (i - (tstart[lo] - 2 * lo) / 2)];
      v |= (uint8_t)((c & 15) << (4 * h));
    }
    packed[b] = v;
  }
}
__global__ void k_unpack_fwd(const uint8_t* __restrict__ packed, int64_t tot, uint8_t* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < tot; i += stride) out[i] = (packed[i >> 1] >> (4 * (i & 1))) & 15;
}

// longest run of `code` in T: per chunk the run touching its left end, the run touching its right end, the best run and
// whether the chunk is all `code`; the host stitches the chunks
constexpr int RUN_CHUNK = 4096;
__global__ void k_chunk_runs(const uint8_t* __restrict__ T, int64_t n, uint8_t code, int64_t n_chunks, int32_t* __restrict__ out /* 3 per chunk */) {
  const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chunks) return;
  const int64_t a = c * RUN_CHUNK, b = min(n, a + RUN_CHUNK);
  int lead = -1, cur = 0, best = 0;
  for (int64_t i = a; i < b; ++i) {
    if (T[i] == code) { ++cur; if (cur > best) best = cur; }
    else { if (lead < 0) lead = cur; cur = 0; }
  }
  if (lead < 0) lead = (int)(b - a);      // the whole chunk
  out[c * 3] = lead; out[c * 3 + 1] = cur; out[c * 3 + 2] = best;
}
static int longest_run(const uint8_t* d_T, int64_t n, uint8_t code, int64_t* out) {
  const int64_t nc = (n + RUN_CHUNK - 1) / RUN_CHUNK;
  *out = 0;
  if (nc == 0) return SVB_OK;
  int32_t* d = nullptr;
  SVB_CUDA(cudaMalloc((void**)&d, (size_t)nc * 12));
  k_chunk_runs<<<(unsigned)((nc + 127) / 128), 128>>>(d_T, n, code, nc, d);
  std::vector<int32_t> h((size_t)nc * 3);
  cudaError_t e = cudaMemcpy(h.data(), d, (size_t)nc * 12, cudaMemcpyDeviceToHost);
  cudaFree(d);
  SVB_CUDA(e);
  int64_t best = 0, open = 0;             // open = run reaching the end of the previous chunk
  for (int64_t c = 0; c < nc; ++c) {
    const int64_t len = std::min<int64_t>(RUN_CHUNK, n - c * RUN_CHUNK);
    const int64_t lead = h[(size_t)c * 3], trail = h[(size_t)c * 3 + 1], inner = h[(size_t)c * 3 + 2];
    best = std::max(best, std::max(inner, open + lead));
    open = lead == len ? open + len : trail;
  }
  *out = best;
  return SVB_OK;
}

// keeps T (device, n bytes) as IndexDev::d_text and uploads the contig-pair starts
static int attach_text(IndexDev* idx, const uint8_t* d_T, const std::vector<int64_t>& offs0 /* m+1, offs0[0] == 0 */) {
  const int64_t n = idx->n, m = (int64_t)offs0.size() - 1;
  SVB_CUDA(cudaMalloc((void**)&idx->d_text_alloc, (size_t)n + 2 * TEXT_PAD));
  idx->d_text = idx->d_text_alloc + TEXT_PAD;
  k_text_device<<<grid_for(n + 2 * TEXT_PAD, 256, 8), 256>>>(d_T, n, idx->d_text_alloc);
  SVB_CUDA(cudaGetLastError());
  std::vector<int64_t> ts((size_t)m + 1);
  for (int64_t r = 0; r <= m; ++r) ts[r] = 2 * offs0[r] + 2 * r;
  SVB_CUDA(cudaMalloc((void**)&idx->d_tstart, (size_t)(m + 1) * 8));
  SVB_CUDA(cudaMemcpy(idx->d_tstart, ts.data(), (size_t)(m + 1) * 8, cudaMemcpyHostToDevice));
  SVB_CUDA(cudaDeviceSynchronize());
  SVB_TRY(longest_run(d_T, n, 5, &idx->max_nrun));
  return SVB_OK;
}

template <int G>
__global__ void k_decode_bwt(const uint4* __restrict__ blocks, int64_t n, uint8_t* __restrict__ out) {
  int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    uint4 s = blocks[i >> 5];
    int j = (int)(i & 31);
    out[i] = (uint8_t)(((s.y >> j) & 1u) | (((s.z >> j) & 1u) << 1) | (((s.w >> j) & 1u) << 2));
  }
}

static int pick_G(int block_bytes, int* G) {
  if (block_bytes == 0 || block_bytes == 128) { *G = 8; return SVB_OK; }
  if (block_bytes == 64) { *G = 4; return SVB_OK; }
  set_error("block_bytes must be 0, 64 or 128 (got %d)", block_bytes);
  return SVB_EINVAL;
}

static int need_device(int device) {
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

}  // namespace svb

using namespace svb;

extern "C" {

const char* svb_last_error(void) { return svb::last_error(); }
const char* svb_version(void) { return "svdss_b200 0.1 (sm_100a)"; }
int svb_device_count(void) {
  int cnt = 0;
  if (cudaGetDeviceCount(&cnt) != cudaSuccess) return 0;
  return cnt;
}

int svb_suffix_array(const uint8_t* text, int64_t n, int mem, int device, int64_t* sa_host) {
  SVB_TRY(need_device(device));
  if (!text || n <= 0 || !sa_host) { set_error("svb_suffix_array: bad arguments"); return SVB_EINVAL; }
  DevBuf<uint8_t> T;
  const uint8_t* d_T = text;
  if (mem == SVB_MEM_HOST) {
    if (text[n - 1] != 0) { set_error("text must end with the sentinel 0"); return SVB_EINVAL; }
    SVB_CUDA(T.alloc(n));
    SVB_CUDA(cudaMemcpy(T.p, text, n, cudaMemcpyHostToDevice));
    d_T = T.p;
  }
  DevBuf<uint64_t> SA;
  SVB_CUDA(SA.alloc(n));
  SVB_TRY(build_suffix_array(d_T, n, SA.p, 0));
  SVB_CUDA(cudaMemcpy(sa_host, SA.p, n * 8, cudaMemcpyDeviceToHost));
  return SVB_OK;
}

int svb_index_build(const uint8_t* contigs, const int64_t* offs, int64_t m, int mem, int device,
                    int block_bytes, svb_index_t** out) {
  SVB_TRY(need_device(device));
  if (!contigs || !offs || m <= 0 || !out) { set_error("svb_index_build: bad arguments"); return SVB_EINVAL; }
  int G;
  SVB_TRY(pick_G(block_bytes, &G));
  std::vector<int64_t> hoffs(m + 1);
  DevBuf<int64_t> d_offs;
  DevBuf<uint8_t> d_seqs_own;
  const uint8_t* d_seqs = contigs;
  const int64_t* d_offs_p = offs;
  if (mem == SVB_MEM_HOST) {
    memcpy(hoffs.data(), offs, (m + 1) * 8);
    int64_t tot = hoffs[m] - hoffs[0];
    SVB_CUDA(d_seqs_own.alloc(std::max<int64_t>(tot, 1)));
    SVB_CUDA(cudaMemcpy(d_seqs_own.p, contigs + hoffs[0], tot, cudaMemcpyHostToDevice));
    // rebase offsets to the uploaded slice
    std::vector<int64_t> rb(m + 1);
    for (int64_t i = 0; i <= m; ++i) rb[i] = hoffs[i] - hoffs[0];
    hoffs = rb;
    SVB_CUDA(d_offs.alloc(m + 1));
    SVB_CUDA(cudaMemcpy(d_offs.p, hoffs.data(), (m + 1) * 8, cudaMemcpyHostToDevice));
    d_seqs = d_seqs_own.p;
    d_offs_p = d_offs.p;
  } else {
    SVB_CUDA(cudaMemcpy(hoffs.data(), offs, (m + 1) * 8, cudaMemcpyDeviceToHost));
  }
  for (int64_t i = 0; i < m; ++i)
    if (hoffs[i + 1] < hoffs[i]) { set_error("contig offsets must be non-decreasing"); return SVB_EINVAL; }
  int64_t n = 2 * (hoffs[m] - hoffs[0] + m);
  for (int64_t i = m; i >= 0; --i) hoffs[i] -= hoffs[0];   // k_build_text lays the text out from offs[0]
  DevBuf<uint8_t> T;
  SVB_CUDA(T.alloc(n));
  k_build_text<<<grid_for(n, 256, 4), 256>>>(d_seqs, d_offs_p, m, n, T.p);
  SVB_CUDA(cudaGetLastError());
  SVB_CUDA(cudaDeviceSynchronize());
  d_seqs_own.release();
  svb_index* idx = new svb_index();
  idx->dev.device = device;
  idx->dev.G = G;
  idx->dev.n = n;
  idx->dev.n_contigs = m;
  int rc;
  {
    DevBuf<uint64_t> SA;
    cudaError_t e = SA.alloc(n);
    if (e != cudaSuccess) { set_error("cudaMalloc SA failed: %s", cudaGetErrorString(e)); delete idx; return SVB_ENOMEM; }
    rc = build_suffix_array(T.p, n, SA.p, 0);
    if (rc == SVB_OK) rc = build_blocks(&idx->dev, nullptr, SA.p, T.p, 0);
    if (rc == SVB_OK) {
      IndexDev& d = idx->dev;
      d.n_ssa = ((n - 1) >> d.ss_log) + 1;
      if (cudaMalloc((void**)&d.d_ssa, (size_t)d.n_ssa * 8) != cudaSuccess) { set_error("cudaMalloc of the SA samples failed"); rc = SVB_ENOMEM; }
      else {
        k_sample_sa<<<grid_for(d.n_ssa, 256, 4), 256>>>(SA.p, d.n_ssa, d.ss_log, d.d_ssa);
        if (cudaDeviceSynchronize() != cudaSuccess) { set_error("k_sample_sa failed"); rc = SVB_ECUDA; }
      }
    }
  }
  if (rc == SVB_OK) rc = attach_text(&idx->dev, T.p, hoffs);
  T.release();
  if (rc == SVB_OK) rc = build_kmer_table(&idx->dev);
  if (rc != SVB_OK) { svb_index_free(idx); return rc; }
  *out = idx;
  return SVB_OK;
}

int svb_index_from_bwt(const uint8_t* bwt, int64_t n, int mem, int device, int block_bytes, svb_index_t** out) {
  SVB_TRY(need_device(device));
  if (!bwt || n <= 0 || !out) { set_error("svb_index_from_bwt: bad arguments"); return SVB_EINVAL; }
  int G;
  SVB_TRY(pick_G(block_bytes, &G));
  DevBuf<uint8_t> own;
  const uint8_t* d_bwt = bwt;
  if (mem == SVB_MEM_HOST) {
    SVB_CUDA(own.alloc(n));
    SVB_CUDA(cudaMemcpy(own.p, bwt, n, cudaMemcpyHostToDevice));
    d_bwt = own.p;
  }
  svb_index* idx = new svb_index();
  idx->dev.device = device;
  idx->dev.G = G;
  idx->dev.n = n;
  int rc = build_blocks(&idx->dev, d_bwt, nullptr, nullptr, 0);
  if (rc != SVB_OK) { svb_index_free(idx); return rc; }
  *out = idx;
  return SVB_OK;
}

void svb_index_free(svb_index_t* idx) {
  if (!idx) return;
  cudaSetDevice(idx->dev.device);
  if (idx->dev.d_blocks) cudaFree(idx->dev.d_blocks);
  if (idx->dev.d_cntN) cudaFree(idx->dev.d_cntN);
  if (idx->dev.d_sbase) cudaFree(idx->dev.d_sbase);
  if (idx->dev.d_text_alloc) cudaFree(idx->dev.d_text_alloc);
  if (idx->dev.d_ssa) cudaFree(idx->dev.d_ssa);
  if (idx->dev.d_tstart) cudaFree(idx->dev.d_tstart);
  if (idx->dev.d_kmt) cudaFree(idx->dev.d_kmt);
  delete idx;
}

int svb_index_info(const svb_index_t* idx, svb_index_info_t* info) {
  if (!idx || !info) { set_error("svb_index_info: null argument"); return SVB_EINVAL; }
  const IndexDev& d = idx->dev;
  info->n = d.n;
  memcpy(info->acc, d.acc, sizeof(d.acc));
  info->n_blocks = d.n_blocks;
  info->block_bytes = d.G * 16;
  info->block_syms = d.G * 32;
  info->n_contigs = d.n_contigs;
  info->device_bytes = d.n_blocks * d.G * 16 + (d.G == 4 ? d.n_blocks * 4 : 0) + (int64_t)d.n_sb * 64 +
                       (d.d_text ? d.n + 2 * TEXT_PAD + d.n_ssa * 8 + (d.n_contigs + 1) * 8 : 0) +
                       (d.d_kmt ? (int64_t)8 << (2 * d.kmer_k) : 0);
  info->device = d.device;
  return SVB_OK;
}

int svb_index_get_bwt(const svb_index_t* idx, uint8_t* bwt_host) {
  if (!idx || !bwt_host) { set_error("svb_index_get_bwt: null argument"); return SVB_EINVAL; }
  SVB_TRY(need_device(idx->dev.device));
  int64_t n = idx->dev.n;
  const int64_t CH = 1LL << 30;
  DevBuf<uint8_t> buf;
  SVB_CUDA(buf.alloc(std::min(n, CH)));
  for (int64_t c0 = 0; c0 < n; c0 += CH) {
    int64_t cn = std::min(CH, n - c0);
    // chunk start is a multiple of 32 symbols so slices line up
    if (idx->dev.G == 4) k_decode_bwt<4><<<grid_for(cn, 256, 4), 256>>>(idx->dev.d_blocks + (c0 >> 5), cn, buf.p);
    else k_decode_bwt<8><<<grid_for(cn, 256, 4), 256>>>(idx->dev.d_blocks + (c0 >> 5), cn, buf.p);
    SVB_CUDA(cudaGetLastError());
    SVB_CUDA(cudaMemcpy(bwt_host + c0, buf.p, cn, cudaMemcpyDeviceToHost));
  }
  return SVB_OK;
}

// ---- index file: private layout (the reference treats the index file as opaque, run_svdss:137-164)
struct FileHeader {
  char magic[8];  // "SVB200I\3"
  int32_t G;
  int32_t n_sb;
  int64_t n;
  int64_t acc[7];
  int64_t n_blocks;
  int64_t n_contigs;
  int32_t has_text;  // 1: tstart[n_contigs+1], 4-bit packed forward strands, SA samples follow the blocks
  int32_t ss_log;
  int64_t n_ssa;
};

int svb_index_save(const svb_index_t* idx, const char* path) {
  if (!idx || !path) { set_error("svb_index_save: null argument"); return SVB_EINVAL; }
  SVB_TRY(need_device(idx->dev.device));
  const IndexDev& d = idx->dev;
  FILE* f = fopen(path, "wb");
  if (!f) { set_error("cannot open %s for writing", path); return SVB_EIO; }
  FileHeader h;
  memset(&h, 0, sizeof(h));
  memcpy(h.magic, "SVB200I\3", 8);
  h.G = d.G; h.n_sb = d.n_sb; h.n = d.n; memcpy(h.acc, d.acc, sizeof(d.acc));
  h.n_blocks = d.n_blocks; h.n_contigs = d.n_contigs;
  h.has_text = d.d_text ? 1 : 0; h.ss_log = d.ss_log; h.n_ssa = d.n_ssa;
  bool ok = fwrite(&h, sizeof(h), 1, f) == 1;
  const size_t CH = (size_t)256 << 20;
  std::vector<uint8_t> buf(CH);
  auto dump = [&](const void* dptr, size_t bytes) {
    for (size_t o = 0; ok && o < bytes; o += CH) {
      size_t c = std::min(CH, bytes - o);
      if (cudaMemcpy(buf.data(), (const uint8_t*)dptr + o, c, cudaMemcpyDeviceToHost) != cudaSuccess) { ok = false; break; }
      ok = fwrite(buf.data(), 1, c, f) == c;
    }
  };
  dump(d.d_sbase, (size_t)d.n_sb * 64);
  dump(d.d_blocks, (size_t)d.n_blocks * d.G * 16);
  if (d.G == 4) dump(d.d_cntN, (size_t)d.n_blocks * 4);
  if (d.d_text) {
    const int64_t tot = (d.n - 2 * d.n_contigs) / 2;
    DevBuf<uint8_t> packed;
    if (packed.alloc((size_t)std::max<int64_t>((tot + 1) / 2, 1)) != cudaSuccess) ok = false;
    else {
      k_pack_fwd<<<grid_for((tot + 1) / 2, 256, 4), 256>>>(d.d_text, d.d_tstart, d.n_contigs, tot, packed.p);
      ok = ok && cudaDeviceSynchronize() == cudaSuccess;
      dump(d.d_tstart, (size_t)(d.n_contigs + 1) * 8);
      dump(packed.p, (size_t)((tot + 1) / 2));
      dump(d.d_ssa, (size_t)d.n_ssa * 8);
    }
  }
  ok = (fclose(f) == 0) && ok;
  if (!ok) { set_error("short write to %s", path); return SVB_EIO; }
  return SVB_OK;
}

int svb_index_load(const char* path, int device, svb_index_t** out) {
  if (!path || !out) { set_error("svb_index_load: null argument"); return SVB_EINVAL; }
  SVB_TRY(need_device(device));
  FILE* f = fopen(path, "rb");
  if (!f) { set_error("cannot open index %s", path); return SVB_EIO; }
  FileHeader h;
  memset(&h, 0, sizeof(h));
  const bool got = fread(&h, sizeof(h), 1, f) == 1;
  if (memcmp(h.magic, "RLD\3", 4) == 0) {
    // An index written by the reference itself (`SVDSS index -d` = ropebwt3 build -d): decode the
    // run-length-delta stream, invert the BWT, keep one strand of every (S, rc S) pair and index that here --
    // what rb3_fmi_restore(&index, path, 0) at ping_pong.cpp:244-245 is handed.  Host work of 1-2 minutes for
    // a human genome; `SVDSS index --from-fmd` converts once.  Layout restated from memory (host/rld.hpp).
    fclose(f);
    svdss::RldFile rf;
    std::string err;
    std::vector<uint8_t> bwt;
    if (!svdss::Rld::read(path, rf, err) || !svdss::Rld::decode_bwt(rf, bwt, err)) { set_error("%s", err.c_str()); return SVB_EIO; }
    if (rf.asize != 6) { set_error("%s: alphabet of %d symbols, expected 6 ($ACGTN)", path, rf.asize); return SVB_EINVAL; }
    rf.words.clear(); rf.words.shrink_to_fit();
    std::vector<std::string> seqs;
    {
      svdss::BwtInverter inv(bwt.data(), bwt.size());
      if (!inv.sequences(seqs)) { set_error("%s: the BWT does not invert to '$'-terminated sequences", path); return SVB_EINVAL; }
    }
    bwt.clear(); bwt.shrink_to_fit();
    std::vector<size_t> keep;
    if (!svdss::forward_strands(seqs, keep) || keep.empty()) {
      set_error("%s: sequences without their reverse complement (ropebwt3 build -R?); the search needs both strands", path);
      return SVB_EINVAL;
    }
    std::vector<uint8_t> cat;
    std::vector<int64_t> offs(1, 0);
    for (size_t k : keep) {
      cat.insert(cat.end(), seqs[k].begin(), seqs[k].end());
      offs.push_back((int64_t)cat.size());
      std::string().swap(seqs[k]);
    }
    if (cat.empty()) cat.push_back(0);
    return svb_index_build(cat.data(), offs.data(), (int64_t)offs.size() - 1, SVB_MEM_HOST, device, 0, out);
  }
  if (!got || memcmp(h.magic, "SVB200I\3", 8) != 0 || (h.G != 4 && h.G != 8)) {
    fclose(f);
    set_error("%s is neither a svdss_b200 index nor a ropebwt3 .fmd (RLD\\3) file", path);
    return SVB_EINVAL;
  }
  svb_index* idx = new svb_index();
  IndexDev& d = idx->dev;
  d.device = device; d.G = h.G; d.n_sb = h.n_sb; d.n = h.n; memcpy(d.acc, h.acc, sizeof(d.acc));
  d.n_blocks = h.n_blocks; d.n_contigs = h.n_contigs;
  bool ok = cudaMalloc((void**)&d.d_sbase, (size_t)d.n_sb * 64) == cudaSuccess &&
            cudaMalloc((void**)&d.d_blocks, (size_t)d.n_blocks * d.G * 16) == cudaSuccess &&
            (d.G != 4 || cudaMalloc((void**)&d.d_cntN, (size_t)d.n_blocks * 4) == cudaSuccess);
  const size_t CH = (size_t)256 << 20;
  std::vector<uint8_t> buf(CH);
  auto slurp = [&](void* dptr, size_t bytes) {
    for (size_t o = 0; ok && o < bytes; o += CH) {
      size_t c = std::min(CH, bytes - o);
      if (fread(buf.data(), 1, c, f) != c) { ok = false; break; }
      ok = cudaMemcpy((uint8_t*)dptr + o, buf.data(), c, cudaMemcpyHostToDevice) == cudaSuccess;
    }
  };
  slurp(d.d_sbase, (size_t)d.n_sb * 64);
  slurp(d.d_blocks, (size_t)d.n_blocks * d.G * 16);
  if (d.G == 4) slurp(d.d_cntN, (size_t)d.n_blocks * 4);
  if (ok && h.has_text) {
    const int64_t m = d.n_contigs, tot = (d.n - 2 * m) / 2;
    d.ss_log = h.ss_log; d.n_ssa = h.n_ssa;
    std::vector<int64_t> ts((size_t)m + 1), offs0((size_t)m + 1);
    ok = fread(ts.data(), 8, (size_t)m + 1, f) == (size_t)m + 1;
    for (int64_t r = 0; ok && r <= m; ++r) offs0[r] = (ts[r] - 2 * r) / 2;
    DevBuf<uint8_t> packed, fwd, T;
    DevBuf<int64_t> d_offs;
    ok = ok && tot >= 0 && offs0[m] == tot && packed.alloc((size_t)std::max<int64_t>((tot + 1) / 2, 1)) == cudaSuccess &&
         fwd.alloc((size_t)std::max<int64_t>(tot, 1)) == cudaSuccess && T.alloc((size_t)d.n) == cudaSuccess &&
         d_offs.alloc((size_t)m + 1) == cudaSuccess && cudaMalloc((void**)&d.d_ssa, (size_t)d.n_ssa * 8) == cudaSuccess;
    slurp(packed.p, (size_t)((tot + 1) / 2));
    slurp(d.d_ssa, (size_t)d.n_ssa * 8);
    if (ok) {
      k_unpack_fwd<<<grid_for(tot, 256, 8), 256>>>(packed.p, tot, fwd.p);
      ok = cudaMemcpy(d_offs.p, offs0.data(), (size_t)(m + 1) * 8, cudaMemcpyHostToDevice) == cudaSuccess;
      k_build_text<<<grid_for(d.n, 256, 4), 256>>>(fwd.p, d_offs.p, m, d.n, T.p);
      ok = ok && cudaDeviceSynchronize() == cudaSuccess && attach_text(&d, T.p, offs0) == SVB_OK;
    }
  }
  fclose(f);
  if (!ok) { svb_index_free(idx); set_error("failed to load index %s (truncated file or out of device memory)", path); return SVB_EIO; }
  { const int rc = build_kmer_table(&d); if (rc != SVB_OK) { svb_index_free(idx); return rc; } }
  *out = idx;
  return SVB_OK;
}

}  // extern "C"
