// Global two-piece affine-gap alignment with CIGAR on the GPU: the batch equivalent of
//   ksw_extd2_sse(0, ql, qs, tl, ts, 5, mat, 16, 2, 41, 1, -1, -1, -1, 0, &ez)
// as called per sub-cluster by Caller::pcall (reference caller.cpp:332-355; ksw2 is an un-vendored
// dependency, its recurrence and traceback rules are restated in SURVEY.md A.2 and followed here).
//
// Mapping: one warp per (consensus, reference window) pair.  The DP matrix (i = target row,
// j = query column) is swept in bands of 32*R target rows; lane l owns R consecutive rows and
// walks the columns with a skew of l steps (wavefront), so every step the lane above hands down
// (H, E, E2) of its last row with three shuffles.  Band boundaries (the last row of a band) go
// through a small per-pair global buffer.  Per cell one traceback byte in ksw2's encoding (low 3
// bits arg-max state with priority H,E,F,E2,F2; bits 3-6 continuation flags) is stored
// wavefront-major -- [band][step][lane][row] -- so each warp step writes 32*R contiguous bytes.
// Lane 0 then replays ksw_backtrack over those bytes.  Integer SIMT, not tensor cores: the
// recurrence is a max-plus scan with data-dependent gaps, not a dense contraction.
#include <cub/cub.cuh>

#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace svb {

int check_device(int device);

constexpr int KSW_NEG_INF = -0x40000000;
constexpr int KR = 4;           // rows per lane
constexpr int KBAND = 32 * KR;  // target rows per band

struct KswParams {
  const uint8_t* __restrict__ q;
  const int64_t* __restrict__ qoff;
  const uint8_t* __restrict__ t;
  const int64_t* __restrict__ toff;
  const uint32_t* __restrict__ order;  // pair indices of this wave, biggest first
  int n;                               // pairs in this wave
  const int64_t* __restrict__ tb_off;  // per wave slot: offset into tb
  const int64_t* __restrict__ bnd_off; // per wave slot: offset into bnd (int32 triples per column)
  const int64_t* __restrict__ cg_off;  // per wave slot: offset into cigar scratch (capacity ql+tl+2)
  uint8_t* tb;
  int32_t* bnd;
  uint32_t* cg;      // reverse-order ops per pair
  int32_t* cg_n;     // per wave slot: number of ops
  int32_t* score;    // per pair (global index)
  unsigned int* work;
  int a, b, sc_n, q1, e1, q2, e2;
};

__device__ __forceinline__ int gapcost(int k, int q1, int e1, int q2, int e2) { return min(q1 + k * e1, q2 + k * e2); }

__global__ void __launch_bounds__(128) k_ksw_extd2(const KswParams P) {
  const int lane = threadIdx.x & 31;
  const int NEG = -0x1fffffff;
  int q1 = P.q1, e1 = P.e1, q2 = P.q2, e2 = P.e2;
  if (q2 + e2 < q1 + e1) { int x = q1; q1 = q2; q2 = x; x = e1; e1 = e2; e2 = x; }  // ksw2 swaps the pieces
  for (;;) {
    unsigned w = 0;
    if (lane == 0) w = atomicAdd(P.work, 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= (unsigned)P.n) break;
    const uint32_t pid = P.order[w];
    const uint8_t* Q = P.q + P.qoff[pid];
    const uint8_t* T = P.t + P.toff[pid];
    const int ql = (int)(P.qoff[pid + 1] - P.qoff[pid]);
    const int tl = (int)(P.toff[pid + 1] - P.toff[pid]);
    if (ql <= 0 || tl <= 0) {  // ksw_reset_extz + early return
      if (lane == 0) { P.score[pid] = KSW_NEG_INF; P.cg_n[w] = 0; }
      continue;
    }
    uint8_t* tb = P.tb + P.tb_off[w];
    int32_t* bnd = P.bnd + P.bnd_off[w];  // [3][ql]: H, E, E2 of the row above the current band
    const int nbands = (tl + KBAND - 1) / KBAND;
    const int nsteps = ql + 31;
    int final_score = 0;
    for (int band = 0; band < nbands; ++band) {
      const int i0 = band * KBAND + lane * KR;  // first row of this lane
      uint8_t tc[KR];
      int hl[KR], f[KR], f2[KR];
#pragma unroll
      for (int r = 0; r < KR; ++r) {
        const int i = i0 + r;
        tc[r] = i < tl ? T[i] : 4;
        hl[r] = -gapcost(i + 1, q1, e1, q2, e2);  // H(i,-1)
        f[r] = hl[r] - q1 - e1;                   // F(i,0)
        f2[r] = hl[r] - q2 - e2;
      }
      int hup_prev = i0 ? -gapcost(i0, q1, e1, q2, e2) : 0;  // H(i0-1,-1)
      // values handed down from the row above at this lane's current column
      int in_h = 0, in_e = NEG, in_e2 = NEG;
      int out_h = 0, out_e = NEG, out_e2 = NEG;  // this lane's last row at its previous column
      int qc = 4, qnext = 4;
      int bh = 0, be = NEG, be2 = NEG;  // lane 0's boundary inputs, prefetched 32 columns at a time
      int wh = 0, we = NEG, we2 = NEG;  // lane 31's boundary outputs, flushed 32 columns at a time
      uint8_t* tbb = tb + (size_t)band * nsteps * KBAND;
      for (int t = 0; t < nsteps; ++t) {
        // ---- inputs for this step
        if ((t & 31) == 0) {
          const int jj = t + lane;  // cooperative prefetch of 32 columns of query + boundary
          qnext = jj < ql ? Q[jj] : 4;
          if (band == 0) {
            bh = jj < ql ? -gapcost(jj + 1, q1, e1, q2, e2) : 0;  // H(-1,j)
            be = NEG; be2 = NEG;
          } else if (jj < ql) {
            bh = bnd[jj]; be = bnd[ql + jj]; be2 = bnd[2 * ql + jj];
          }
        }
        // query char: lane 0 takes column t, others inherit from the lane above (one step later)
        const int q_in = __shfl_sync(0xffffffffu, qnext, t & 31);
        const int q_up = __shfl_up_sync(0xffffffffu, qc, 1);
        qc = lane == 0 ? q_in : q_up;
        const int b_h = __shfl_sync(0xffffffffu, bh, t & 31);
        const int b_e = __shfl_sync(0xffffffffu, be, t & 31);
        const int b_e2 = __shfl_sync(0xffffffffu, be2, t & 31);
        const int u_h = __shfl_up_sync(0xffffffffu, out_h, 1);
        const int u_e = __shfl_up_sync(0xffffffffu, out_e, 1);
        const int u_e2 = __shfl_up_sync(0xffffffffu, out_e2, 1);
        in_h = lane == 0 ? b_h : u_h;
        in_e = lane == 0 ? b_e : u_e;
        in_e2 = lane == 0 ? b_e2 : u_e2;
        const int j = t - lane;
        const bool act = j >= 0 && j < ql;
        unsigned tbw = 0;
        if (act) {
          int hup = in_h, eup = in_e, e2up = in_e2;
          int hdiag = hup_prev;
          hup_prev = in_h;
#pragma unroll
          for (int r = 0; r < KR; ++r) {
            const int ee = max(hup - q1, eup) - e1;
            const int ee2 = max(hup - q2, e2up) - e2;
            const int sc = (tc[r] == 4 || qc == 4) ? P.sc_n : (tc[r] == qc ? P.a : P.b);
            int h = hdiag + sc;
            unsigned d = 0;
            if (ee > h) { h = ee; d = 1; }
            if (f[r] > h) { h = f[r]; d = 2; }
            if (ee2 > h) { h = ee2; d = 3; }
            if (f2[r] > h) { h = f2[r]; d = 4; }
            const int ho1 = h - q1, ho2 = h - q2;
            d |= (ee > ho1) ? 0x08u : 0u;
            d |= (f[r] > ho1) ? 0x10u : 0u;
            d |= (ee2 > ho2) ? 0x20u : 0u;
            d |= (f2[r] > ho2) ? 0x40u : 0u;
            tbw |= d << (8 * r);
            f[r] = max(ho1, f[r]) - e1;
            f2[r] = max(ho2, f2[r]) - e2;
            hdiag = hl[r];
            hl[r] = h;
            hup = h; eup = ee; e2up = ee2;
            if (j == ql - 1 && i0 + r == tl - 1) final_score = h;
          }
          out_h = hup; out_e = eup; out_e2 = e2up;
          *reinterpret_cast<unsigned*>(tbb + (size_t)t * KBAND + lane * KR) = tbw;
        }
        // ---- lane 31 hands its last row to the next band: collect 32 columns, flush coalesced
        if (band + 1 < nbands) {
          const int j31 = t - 31;  // column lane 31 just finished
          const int s_h = __shfl_sync(0xffffffffu, out_h, 31);
          const int s_e = __shfl_sync(0xffffffffu, out_e, 31);
          const int s_e2 = __shfl_sync(0xffffffffu, out_e2, 31);
          if (j31 >= 0 && j31 < ql) {
            if ((j31 & 31) == lane) { wh = s_h; we = s_e; we2 = s_e2; }
            if ((j31 & 31) == 31 || j31 == ql - 1) {
              const int jj = (j31 & ~31) + lane;
              if (jj <= j31) { bnd[jj] = wh; bnd[ql + jj] = we; bnd[2 * ql + jj] = we2; }
            }
          }
        }
      }
      __syncwarp();
    }
    // score lives in the lane that owned row tl-1
    {
      const int owner = ((tl - 1) % KBAND) / KR;
      final_score = __shfl_sync(0xffffffffu, final_score, owner);
    }
    __syncwarp();
    // ---- ksw_backtrack (ksw2.h) by lane 0
    if (lane == 0) {
      uint32_t* cg = P.cg + P.cg_off[w];
      int n = 0, i = tl - 1, j = ql - 1, state = 0;
      uint32_t last = 0;  // open run: len << 4 | op, 0 = none
      auto push = [&](unsigned op, unsigned len) {
        if (last && (last & 0xfu) == op) last += len << 4;
        else { if (last) cg[n++] = last; last = (len << 4) | op; }
      };
      while (i >= 0 && j >= 0) {
        const int bnd_ = i / KBAND, l = (i % KBAND) / KR, r = i % KR;
        const unsigned tmp = tb[(size_t)bnd_ * nsteps * KBAND + (size_t)(j + l) * KBAND + l * KR + r];
        if (state == 0) state = tmp & 7;
        else if (!((tmp >> (state + 2)) & 1)) state = 0;
        if (state == 0) state = tmp & 7;
        if (state == 0) { push(0, 1); --i; --j; }
        else if (state == 1 || state == 3) { push(2, 1); --i; }
        else { push(1, 1); --j; }
      }
      if (i >= 0) push(2, (unsigned)(i + 1));
      if (j >= 0) push(1, (unsigned)(j + 1));
      if (last) cg[n++] = last;
      P.cg_n[w] = n;
      P.score[pid] = final_score;
    }
    __syncwarp();
  }
}

// reverse the per-pair op lists of one wave into a dense wave-local table
__global__ void k_ksw_gather(const uint32_t* __restrict__ cg, const int64_t* __restrict__ cg_off, int n_wave,
                             const int32_t* __restrict__ cg_n, const int64_t* __restrict__ dense_off,
                             uint32_t* __restrict__ dense) {
  const int w = blockIdx.x;
  if (w >= n_wave) return;
  const int n = cg_n[w];
  const uint32_t* src = cg + cg_off[w];
  uint32_t* dst = dense + dense_off[w];
  for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[n - 1 - k];
}

}  // namespace svb

using namespace svb;

extern "C" int svb_ksw_extd2_batch(const uint8_t* q_concat, const int64_t* q_offs, const uint8_t* t_concat,
                                   const int64_t* t_offs, int64_t n_pairs, int match, int mismatch, int sc_n,
                                   int gapo, int gape, int gapo2, int gape2, int device, svb_ksw_out_t* out) {
  if (!out) { set_error("svb_ksw_extd2_batch: null out"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  if (!q_offs || !t_offs || n_pairs < 0 || n_pairs > 0x7fffffff) { set_error("svb_ksw_extd2_batch: bad arguments"); return SVB_EINVAL; }
  SVB_TRY(check_device(device));
  out->n_pairs = n_pairs;
  out->score = (int32_t*)calloc((size_t)n_pairs + 1, 4);
  out->cigar_offs = (int64_t*)calloc((size_t)n_pairs + 1, 8);
  if (!out->score || !out->cigar_offs) { set_error("out of host memory"); return SVB_ENOMEM; }
  if (n_pairs == 0) return SVB_OK;
  const int64_t qtot = q_offs[n_pairs] - q_offs[0], ttot = t_offs[n_pairs] - t_offs[0];
  // order pairs by DP size, biggest first; form waves under the traceback memory budget
  std::vector<uint32_t> order((size_t)n_pairs);
  std::vector<int64_t> cells((size_t)n_pairs);
  double total_cells = 0;
  for (int64_t p = 0; p < n_pairs; ++p) {
    int64_t ql = q_offs[p + 1] - q_offs[p], tl = t_offs[p + 1] - t_offs[p];
    if (ql < 0 || tl < 0 || ql > 0x3fffffff || tl > 0x3fffffff) { set_error("pair %lld has a bad length", (long long)p); return SVB_EINVAL; }
    order[p] = (uint32_t)p;
    cells[p] = ql * tl;
    total_cells += (double)cells[p];
  }
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cells[x] > cells[y]; });
  size_t free_b = 0, total_b = 0;
  SVB_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const char* eb = getenv("SVB_KSW_TB_BYTES");
  int64_t budget = eb ? atoll(eb) : (int64_t)std::min<size_t>(free_b / 2, (size_t)48 << 30);
  auto tb_bytes = [&](uint32_t p) -> int64_t {
    int64_t ql = q_offs[p + 1] - q_offs[p], tl = t_offs[p + 1] - t_offs[p];
    if (ql <= 0 || tl <= 0) return 0;
    return ((tl + KBAND - 1) / KBAND) * (ql + 31) * KBAND;
  };
  if (tb_bytes(order[0]) > budget) budget = tb_bytes(order[0]);

  uint8_t *d_q = nullptr, *d_t = nullptr, *d_tb = nullptr;
  int64_t *d_qoff = nullptr, *d_toff = nullptr, *d_woff = nullptr, *d_outoff = nullptr;
  uint32_t *d_order = nullptr, *d_cg = nullptr, *d_out = nullptr;
  int32_t *d_bnd = nullptr, *d_cgn = nullptr, *d_score = nullptr;
  unsigned int* d_work = nullptr;
  int rc = SVB_OK;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  std::vector<int64_t> qo((size_t)n_pairs + 1), to((size_t)n_pairs + 1);
  for (int64_t p = 0; p <= n_pairs; ++p) { qo[p] = q_offs[p] - q_offs[0]; to[p] = t_offs[p] - t_offs[0]; }
#define KCHECK(expr)                                                                                   \
  do {                                                                                                 \
    cudaError_t _e = (expr);                                                                           \
    if (_e != cudaSuccess) {                                                                           \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e));            \
      rc = SVB_ECUDA;                                                                                  \
      goto done;                                                                                       \
    }                                                                                                  \
  } while (0)
  {
    KCHECK(cudaEventCreate(&e0));
    KCHECK(cudaEventCreate(&e1));
    KCHECK(cudaMalloc((void**)&d_q, std::max<int64_t>(qtot, 1)));
    KCHECK(cudaMalloc((void**)&d_t, std::max<int64_t>(ttot, 1)));
    KCHECK(cudaMalloc((void**)&d_qoff, (n_pairs + 1) * 8));
    KCHECK(cudaMalloc((void**)&d_toff, (n_pairs + 1) * 8));
    KCHECK(cudaMalloc((void**)&d_order, n_pairs * 4));
    KCHECK(cudaMalloc((void**)&d_cgn, n_pairs * 4));
    KCHECK(cudaMalloc((void**)&d_score, n_pairs * 4));
    KCHECK(cudaMalloc((void**)&d_work, 4));
    KCHECK(cudaEventRecord(e0, 0));
    if (qtot) KCHECK(cudaMemcpy(d_q, q_concat + q_offs[0], qtot, cudaMemcpyHostToDevice));
    if (ttot) KCHECK(cudaMemcpy(d_t, t_concat + t_offs[0], ttot, cudaMemcpyHostToDevice));
    KCHECK(cudaMemcpy(d_qoff, qo.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice));
    KCHECK(cudaMemcpy(d_toff, to.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice));
    KCHECK(cudaMemcpy(d_order, order.data(), n_pairs * 4, cudaMemcpyHostToDevice));
    out->h2d_bytes = qtot + ttot + (n_pairs + 1) * 16 + n_pairs * 4;
    // wave plan
    struct Wave { int64_t first, count, tb, bnd, cg; };
    std::vector<Wave> waves;
    std::vector<int64_t> woff((size_t)n_pairs * 3);  // per slot: tb, bnd, cg offsets (wave-local)
    {
      int64_t p = 0;
      while (p < n_pairs) {
        Wave wv{p, 0, 0, 0, 0};
        while (p < n_pairs) {
          uint32_t id = order[p];
          int64_t tbb = tb_bytes(id);
          if (wv.count && wv.tb + tbb > budget) break;
          int64_t ql = qo[id + 1] - qo[id], tl = to[id + 1] - to[id];
          woff[p * 3 + 0] = wv.tb; woff[p * 3 + 1] = wv.bnd; woff[p * 3 + 2] = wv.cg;
          wv.tb += (tbb + 127) & ~127LL;
          wv.bnd += 3 * ql;
          wv.cg += ql + tl + 2;
          ++wv.count; ++p;
        }
        waves.push_back(wv);
      }
    }
    int64_t max_tb = 1, max_bnd = 1, max_cg = 1, max_cnt = 1;
    for (auto& wv : waves) { max_tb = std::max(max_tb, wv.tb); max_bnd = std::max(max_bnd, wv.bnd); max_cg = std::max(max_cg, wv.cg); max_cnt = std::max(max_cnt, wv.count); }
    max_tb = std::max<int64_t>(max_tb, max_cg * 4 + (max_cnt + 1) * 8 + 512);
    KCHECK(cudaMalloc((void**)&d_tb, max_tb));
    KCHECK(cudaMalloc((void**)&d_bnd, max_bnd * 4));
    KCHECK(cudaMalloc((void**)&d_cg, max_cg * 4));
    KCHECK(cudaMalloc((void**)&d_woff, max_cnt * 3 * 8));
    // cigar ops are first collected per wave on the device (reverse order), sizes come back to the
    // host, which lays out the final dense table wave by wave
    std::vector<std::vector<uint32_t>> wave_ops(waves.size());
    std::vector<int32_t> cgn((size_t)n_pairs);
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    float kms_total = 0.f;
    std::vector<int64_t> tmp3;
    for (size_t wi = 0; wi < waves.size(); ++wi) {
      const Wave& wv = waves[wi];
      // slot-major offset arrays for this wave
      tmp3.assign((size_t)wv.count * 3, 0);
      for (int64_t s = 0; s < wv.count; ++s) {
        tmp3[s] = woff[(wv.first + s) * 3 + 0];
        tmp3[wv.count + s] = woff[(wv.first + s) * 3 + 1];
        tmp3[2 * wv.count + s] = woff[(wv.first + s) * 3 + 2];
      }
      KCHECK(cudaMemcpy(d_woff, tmp3.data(), wv.count * 3 * 8, cudaMemcpyHostToDevice));
      KCHECK(cudaMemset(d_work, 0, 4));
      KswParams P;
      P.q = d_q; P.qoff = d_qoff; P.t = d_t; P.toff = d_toff;
      P.order = d_order + wv.first; P.n = (int)wv.count;
      P.tb_off = d_woff; P.bnd_off = d_woff + wv.count; P.cg_off = d_woff + 2 * wv.count;
      P.tb = d_tb; P.bnd = d_bnd; P.cg = d_cg; P.cg_n = d_cgn + wv.first; P.score = d_score; P.work = d_work;
      P.a = match; P.b = mismatch; P.sc_n = sc_n; P.q1 = gapo; P.e1 = gape; P.q2 = gapo2; P.e2 = gape2;
      int64_t warps = std::min<int64_t>(wv.count, (int64_t)sms * 32);
      unsigned grid = (unsigned)((warps + 3) / 4);
      cudaEvent_t k0, k1;
      KCHECK(cudaEventCreate(&k0)); KCHECK(cudaEventCreate(&k1));
      KCHECK(cudaEventRecord(k0, 0));
      k_ksw_extd2<<<grid, 128>>>(P);
      KCHECK(cudaGetLastError());
      KCHECK(cudaEventRecord(k1, 0));
      KCHECK(cudaEventSynchronize(k1));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, k0, k1);
      cudaEventDestroy(k0); cudaEventDestroy(k1);
      kms_total += ms;
      out->launches += 1;
      // op counts of this wave -> dense layout -> gather on the device -> one dense D2H
      KCHECK(cudaMemcpy(cgn.data() + wv.first, d_cgn + wv.first, wv.count * 4, cudaMemcpyDeviceToHost));
      std::vector<int64_t> dense_off((size_t)wv.count + 1, 0);
      for (int64_t s = 0; s < wv.count; ++s) dense_off[s + 1] = dense_off[s] + cgn[wv.first + s];
      const int64_t nd = dense_off[wv.count];
      std::vector<uint32_t>& ops = wave_ops[wi];
      ops.resize((size_t)nd);
      if (nd) {
        // the traceback buffer is free again: reuse its head for the dense table and its offsets
        if (max_tb < nd * 4 + (wv.count + 1) * 8 + 256) { set_error("internal: tb buffer too small for the cigar table"); rc = SVB_ERANGE; goto done; }
        int64_t* d_doff = reinterpret_cast<int64_t*>(d_tb);
        uint32_t* d_dense = reinterpret_cast<uint32_t*>(d_tb + (((wv.count + 1) * 8 + 255) & ~255LL));
        KCHECK(cudaMemcpy(d_doff, dense_off.data(), (wv.count + 1) * 8, cudaMemcpyHostToDevice));
        k_ksw_gather<<<(unsigned)wv.count, 64>>>(d_cg, P.cg_off, (int)wv.count, d_cgn + wv.first, d_doff, d_dense);
        KCHECK(cudaGetLastError());
        KCHECK(cudaMemcpy(ops.data(), d_dense, nd * 4, cudaMemcpyDeviceToHost));
        out->launches += 1;
      }
      out->d2h_bytes += nd * 4 + wv.count * 4;
    }
    KCHECK(cudaMemcpy(out->score, d_score, n_pairs * 4, cudaMemcpyDeviceToHost));
    out->d2h_bytes += n_pairs * 4;
    // cgn is indexed by sorted position; cigar_offs by pair id
    {
      std::vector<int32_t> by_pid((size_t)n_pairs);
      for (int64_t s = 0; s < n_pairs; ++s) by_pid[order[s]] = cgn[s];
      for (int64_t p = 0; p < n_pairs; ++p) out->cigar_offs[p + 1] = out->cigar_offs[p] + by_pid[p];
    }
    out->n_cigar = out->cigar_offs[n_pairs];
    out->cigar = (uint32_t*)malloc(std::max<int64_t>(out->n_cigar, 1) * 4);
    if (!out->cigar) { set_error("out of host memory"); rc = SVB_ENOMEM; goto done; }
    for (size_t wi = 0; wi < waves.size(); ++wi) {
      const Wave& wv = waves[wi];
      int64_t o = 0;
      for (int64_t s = 0; s < wv.count; ++s) {
        uint32_t id = order[wv.first + s];
        int n = cgn[wv.first + s];
        if (n) memcpy(out->cigar + out->cigar_offs[id], wave_ops[wi].data() + o, (size_t)n * 4);
        o += n;
      }
    }
    KCHECK(cudaEventRecord(e1, 0));
    KCHECK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&out->device_ms, e0, e1);
    out->kernel_ms = kms_total;
    out->cells = (int64_t)total_cells;
    out->waves = (int32_t)waves.size();
  }
done:
#undef KCHECK
  cudaFree(d_q); cudaFree(d_t); cudaFree(d_tb); cudaFree(d_qoff); cudaFree(d_toff); cudaFree(d_woff);
  cudaFree(d_outoff); cudaFree(d_order); cudaFree(d_cg); cudaFree(d_out); cudaFree(d_bnd); cudaFree(d_cgn);
  cudaFree(d_score); cudaFree(d_work);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (rc != SVB_OK) svb_ksw_out_free(out);
  return rc;
}

extern "C" void svb_ksw_out_free(svb_ksw_out_t* out) {
  if (!out) return;
  free(out->score); free(out->cigar_offs); free(out->cigar);
  out->score = nullptr; out->cigar_offs = nullptr; out->cigar = nullptr;
}
