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

#include "ksw_kernel.cuh"

namespace svb {

int check_device(int device);

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
  StageLog slog("ksw");
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) { return cells[x] > cells[y]; });
  slog.lap("sort pairs");
  size_t free_b = 0;
  SVB_CUDA(pool_available(&free_b));
  slog.lap("pool_available");
  const char* eb = getenv("SVB_KSW_TB_BYTES");
  // one traceback byte per cell: the budget bounds the cells in flight, hence -- for 10 kb x 10 kb pairs of
  // 100 MB each -- the number of warps that have work.  Most of the free HBM by default (round 1 used
  // min(free / 2, 48 GB): 480 such pairs for 2 960 warp slots).
  int64_t budget = eb ? atoll(eb) : (int64_t)std::min<size_t>((size_t)(free_b * 0.8), (size_t)128 << 30);
  // kernel variant (ksw_kernel.cuh): bit 1 = windowed backtrack with prefetch, bit 2 = checkpointed traceback for
  // pairs of several bands; default 0 = the kernel measured in round 1
  const char* evar = getenv("SVB_KSW_VARIANT");
  const int variant = evar ? atoi(evar) : 0;
  if (variant < 0 || variant > 3) { set_error("SVB_KSW_VARIANT must be 0..3"); return SVB_EINVAL; }
  auto tb_bytes = [&](uint32_t p) -> int64_t { return ksw_tb_bytes(variant, q_offs[p + 1] - q_offs[p], t_offs[p + 1] - t_offs[p]); };
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
    KCHECK(pmalloc((void**)&d_q, (size_t)(std::max<int64_t>(qtot, 1)), 0));
    KCHECK(pmalloc((void**)&d_t, (size_t)(std::max<int64_t>(ttot, 1)), 0));
    KCHECK(pmalloc((void**)&d_qoff, (size_t)((n_pairs + 1) * 8), 0));
    KCHECK(pmalloc((void**)&d_toff, (size_t)((n_pairs + 1) * 8), 0));
    KCHECK(pmalloc((void**)&d_order, (size_t)(n_pairs * 4), 0));
    KCHECK(pmalloc((void**)&d_cgn, (size_t)(n_pairs * 4), 0));
    KCHECK(pmalloc((void**)&d_score, (size_t)(n_pairs * 4), 0));
    KCHECK(pmalloc((void**)&d_work, (size_t)(4), 0));
    KCHECK(cudaEventRecord(e0, 0));
    if (qtot) KCHECK(cudaMemcpy(d_q, q_concat + q_offs[0], qtot, cudaMemcpyHostToDevice));
    if (ttot) KCHECK(cudaMemcpy(d_t, t_concat + t_offs[0], ttot, cudaMemcpyHostToDevice));
    KCHECK(cudaMemcpy(d_qoff, qo.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice));
    KCHECK(cudaMemcpy(d_toff, to.data(), (n_pairs + 1) * 8, cudaMemcpyHostToDevice));
    KCHECK(cudaMemcpy(d_order, order.data(), n_pairs * 4, cudaMemcpyHostToDevice));
    out->h2d_bytes = qtot + ttot + (n_pairs + 1) * 16 + n_pairs * 4;
    slog.lap("buffers + H2D");
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
          wv.bnd += ksw_bnd_ints(variant, ql, tl);
          wv.cg += ql + tl + 2;
          ++wv.count; ++p;
        }
        waves.push_back(wv);
      }
    }
    int64_t max_tb = 1, max_bnd = 1, max_cg = 1, max_cnt = 1;
    for (auto& wv : waves) { max_tb = std::max(max_tb, wv.tb); max_bnd = std::max(max_bnd, wv.bnd); max_cg = std::max(max_cg, wv.cg); max_cnt = std::max(max_cnt, wv.count); }
    max_tb = std::max<int64_t>(max_tb, max_cg * 4 + (max_cnt + 1) * 8 + 512);
    KCHECK(pmalloc((void**)&d_tb, (size_t)(max_tb), 0));
    KCHECK(pmalloc((void**)&d_bnd, (size_t)(max_bnd * 4), 0));
    KCHECK(pmalloc((void**)&d_cg, (size_t)(max_cg * 4), 0));
    KCHECK(pmalloc((void**)&d_woff, (size_t)(max_cnt * 3 * 8), 0));
    slog.lap("wave plan + traceback buffers");
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
      switch (variant) {
        case 1: k_ksw_extd2<1><<<grid, 128>>>(P); break;
        case 2: k_ksw_extd2<2><<<grid, 128>>>(P); break;
        case 3: k_ksw_extd2<3><<<grid, 128>>>(P); break;
        default: k_ksw_extd2<0><<<grid, 128>>>(P); break;
      }
      KCHECK(cudaGetLastError());
      KCHECK(cudaEventRecord(k1, 0));
      KCHECK(cudaEventSynchronize(k1));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, k0, k1);
      cudaEventDestroy(k0); cudaEventDestroy(k1);
      kms_total += ms;
      out->launches += 1;
      slog.lap("wave kernel");
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
      slog.lap("wave cigar gather + D2H");
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
    slog.lap("cigar table on the host");
    KCHECK(cudaEventRecord(e1, 0));
    KCHECK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&out->device_ms, e0, e1);
    out->kernel_ms = kms_total;
    out->cells = (int64_t)total_cells;
    out->waves = (int32_t)waves.size();
  }
done:
#undef KCHECK
  pfree(d_q, 0); pfree(d_t, 0); pfree(d_tb, 0); pfree(d_qoff, 0); pfree(d_toff, 0); pfree(d_woff, 0);
  pfree(d_outoff, 0); pfree(d_order, 0); pfree(d_cg, 0); pfree(d_out, 0); pfree(d_bnd, 0); pfree(d_cgn, 0);
  pfree(d_score, 0); pfree(d_work, 0);
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
