// Cluster-level partial-order alignment + heaviest-bundling consensus on the GPU: the batch
// equivalent of Caller::run_poa (reference caller.cpp:257-308), which hands a cluster's sub-reads
// to abPOA (abpoa_msa, global mode, convex gap, adaptive band, reads added in input order) and
// takes cons_base[0].  abPOA is an un-vendored dependency; its algorithm is restated in SURVEY.md
// A.3 and, with every tie-break fixed, in oracle/poa_oracle.c's header -- this kernel follows
// those rules so that it is bit-identical to the banded oracle.
//
// Mapping: one warp per cluster (clusters are independent, caller.cpp:312-313; 10^4..10^5 of them).
// The graph lives in a per-warp global-memory workspace (L2-resident, a few hundred KB).  Reads
// are added sequentially; for one read the DP runs over graph rows in rank (topological) order and
// the warp parallelises over the query columns of the row's adaptive band, 32 columns per step:
// M/E from the predecessor rows are independent per column, the in-row insertion states F1/F2
// are a max-plus prefix scan (5 shuffle steps each).  Traceback, graph update and consensus are
// short sequential walks done by lane 0.  Integer SIMT: a max-plus recurrence over a DAG with
// data-dependent band and predecessors is not a dense contraction, so no tensor cores.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "common.cuh"

#include "poa_kernel.cuh"


namespace svb {

int check_device(int device);
int poa_batch_impl(const uint8_t* seqs, int seqs_mem, const int64_t* seq_offs, const int64_t* cluster_offs,
                   int64_t n_clusters, int device, svb_poa_out_t* out);

}  // namespace svb

using namespace svb;

// svb_poa_batch with the sequences either in host memory or already on `device` (the call pipeline gathers the
// sub-reads of all clusters on the GPU, call_batch.cu); the offset arrays are always host arrays
int svb::poa_batch_impl(const uint8_t* seqs, int seqs_mem, const int64_t* seq_offs, const int64_t* cluster_offs,
                        int64_t n_clusters, int device, svb_poa_out_t* out) {
  if (!out) { set_error("svb_poa_batch: null out"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  if (!seq_offs || !cluster_offs || n_clusters < 0 || n_clusters > 0x7fffffff) { set_error("svb_poa_batch: bad arguments"); return SVB_EINVAL; }
  SVB_TRY(check_device(device));
  out->n_clusters = n_clusters;
  out->cons_offs = (int64_t*)calloc((size_t)n_clusters + 1, 8);
  out->status = (int32_t*)calloc((size_t)n_clusters + 1, 4);
  if (!out->cons_offs || !out->status) { set_error("out of host memory"); return SVB_ENOMEM; }
  if (n_clusters == 0) return SVB_OK;
  const int64_t n_seqs = cluster_offs[n_clusters];
  if (cluster_offs[0] != 0 || n_seqs < 0) { set_error("cluster offsets must start at 0"); return SVB_EINVAL; }
  const int64_t s_first = seq_offs[0], s_total = seq_offs[n_seqs] - s_first;
  // per-cluster shape
  struct Shape { int64_t sum; int lmax, lmin, nreads; double cost; };
  std::vector<Shape> shp((size_t)n_clusters);
  std::vector<int64_t> cap_off((size_t)n_clusters + 1, 0);
  for (int64_t c = 0; c < n_clusters; ++c) {
    Shape s{0, 0, 0x7fffffff, 0, 0};
    if (cluster_offs[c + 1] < cluster_offs[c]) { set_error("cluster offsets must be non-decreasing"); return SVB_EINVAL; }
    for (int64_t i = cluster_offs[c]; i < cluster_offs[c + 1]; ++i) {
      int64_t l = seq_offs[i + 1] - seq_offs[i];
      if (l < 0 || l > (1 << 24)) { set_error("sequence %lld has a bad length", (long long)i); return SVB_EINVAL; }
      if (l == 0) continue;
      s.sum += l; s.lmax = std::max(s.lmax, (int)l); s.lmin = std::min(s.lmin, (int)l); s.nreads++;
    }
    if (!s.nreads) s.lmin = 0;
    s.cost = (double)s.sum * (s.lmax + 1);
    shp[c] = s;
    cap_off[c + 1] = cap_off[c] + ((2 * (int64_t)s.lmax + 64 + 15) & ~15LL);
  }
  uint8_t *d_seqs = nullptr, *d_ws = nullptr, *d_cons = nullptr;
  int64_t *d_soff = nullptr, *d_coff = nullptr, *d_capoff = nullptr, *d_slotoff = nullptr;
  uint32_t* d_order = nullptr;
  int4* d_dims = nullptr;
  int32_t *d_len = nullptr, *d_status = nullptr;
  unsigned int* d_work = nullptr;
  unsigned long long* d_cells = nullptr;
  unsigned long long* d_phase = nullptr;
  int rc = SVB_OK;
  StageLog slog("poa");
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  std::vector<int64_t> so((size_t)n_seqs + 1);
  for (int64_t i = 0; i <= n_seqs; ++i) so[i] = seq_offs[i] - s_first;
  std::vector<int32_t> h_len((size_t)n_clusters), h_status((size_t)n_clusters);
  std::vector<uint8_t> h_cons((size_t)std::max<int64_t>(cap_off[n_clusters], 1));
#define PCHECK(expr)                                                                        \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      rc = SVB_ECUDA;                                                                       \
      goto done;                                                                            \
    }                                                                                       \
  } while (0)
  {
    // every buffer comes from the device's stream-ordered pool (common.cuh pmalloc): a caller that works through batch
    // after batch pays for the workspace once, not cudaMalloc + cudaFree of tens of GB in every call
    PCHECK(cudaEventCreate(&e0));
    PCHECK(cudaEventCreate(&e1));
    PCHECK(cudaEventRecord(e0, 0));
    if (seqs_mem == SVB_MEM_DEVICE) d_seqs = const_cast<uint8_t*>(seqs) + s_first;
    else PCHECK(pmalloc((void**)&d_seqs, (size_t)std::max<int64_t>(s_total, 1), 0));
    PCHECK(pmalloc((void**)&d_soff, (size_t)(n_seqs + 1) * 8, 0));
    PCHECK(pmalloc((void**)&d_coff, (size_t)(n_clusters + 1) * 8, 0));
    PCHECK(pmalloc((void**)&d_capoff, (size_t)(n_clusters + 1) * 8, 0));
    PCHECK(pmalloc((void**)&d_order, (size_t)n_clusters * 4, 0));
    PCHECK(pmalloc((void**)&d_dims, (size_t)n_clusters * sizeof(int4), 0));
    PCHECK(pmalloc((void**)&d_len, (size_t)n_clusters * 4, 0));
    PCHECK(pmalloc((void**)&d_status, (size_t)n_clusters * 4, 0));
    PCHECK(pmalloc((void**)&d_cons, (size_t)std::max<int64_t>(cap_off[n_clusters], 1), 0));
    PCHECK(pmalloc((void**)&d_work, 4, 0));
    PCHECK(pmalloc((void**)&d_cells, 8, 0));
    PCHECK(cudaMemsetAsync(d_cells, 0, 8, 0));
    if (getenv("SVB_POA_TIMING")) { PCHECK(pmalloc((void**)&d_phase, 40, 0)); PCHECK(cudaMemsetAsync(d_phase, 0, 40, 0)); }
    slog.lap("small buffers");
    if (s_total && seqs_mem != SVB_MEM_DEVICE) PCHECK(cudaMemcpy(d_seqs, seqs + s_first, s_total, cudaMemcpyHostToDevice));
    PCHECK(cudaMemcpy(d_soff, so.data(), (n_seqs + 1) * 8, cudaMemcpyHostToDevice));
    PCHECK(cudaMemcpy(d_coff, cluster_offs, (n_clusters + 1) * 8, cudaMemcpyHostToDevice));
    PCHECK(cudaMemcpy(d_capoff, cap_off.data(), (n_clusters + 1) * 8, cudaMemcpyHostToDevice));
    slog.lap("sequences + offsets H2D");
    out->h2d_bytes = (seqs_mem == SVB_MEM_DEVICE ? 0 : s_total) + (n_seqs + 1) * 8 + (n_clusters + 1) * 16;
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const char* eg = getenv("SVB_POA_GROUP");
    if (eg && atoi(eg) != 32) { set_error("SVB_POA_GROUP: only 32 lanes per cluster are built (16 and 8 were measured slower, profiles/r02a_variants_sweep.txt)"); rc = SVB_EINVAL; goto done; }
    // kernel variant (poa_kernel.cuh): SVB_POA_VARIANT = bit mask; 0 = the kernel measured in round 1
    const char* ev = getenv("SVB_POA_VARIANT");
    int variant = ev ? atoi(ev) : -1;
    if (variant < 0) {
      // Two builds of the same kernel (identical results): 455 = a warp per cluster, 4 CTAs per SM -- the most clusters in
      // flight, 37 GCUPS on 12 000 clusters of 20-60 reads (profiles/r02d_poa_variants.txt); 6599 = a CTA per cluster, its DP
      // rows cut across the four warps, 2 CTAs per SM (217 registers, no spills) -- 2.7x shorter rows for a cluster that
      // runs alone (profiles/r02f_*).  A batch whose biggest cluster is a longer chain of rows (reads x nodes, ~1 us each)
      // than the whole batch is work (cells at ~37 GCUPS) is bound by that chain.
      double chain = 0, cells_est = 0;
      for (int64_t c = 0; c < n_clusters; ++c) {
        const Shape& s = shp[c];
        if (!s.nreads) continue;
        chain = std::max(chain, (double)(s.nreads - 1) * (double)s.lmax);
        cells_est += (double)(s.sum - s.lmax) * (2.0 * (10 + 0.01 * s.lmax) + 1.0);
      }
      variant = (chain * 1.0e-6 > cells_est / 37e9) ? 4096 + 2048 + 455 : 455;
    }
    if (slog.on) {
      // SVB_STAGE_STATS: how the row chains (reads x nodes) of the batch are spread -- the config-3 slice of bench.py has its
      // longest chain at 101 k rows and 265 x that in total: with 296 CTA slots the hand-out (biggest first) packs well and the
      // launch is bound by row throughput (1.24 us per row and CTA at 2 CTAs per SM; 1.03 alone, 2.06 at 3 per SM), not by
      // the longest chain (profiles/r03a_poa_occupancy.txt)
      double chain = 0, rows = 0;
      for (int64_t c = 0; c < n_clusters; ++c) {
        const Shape& s = shp[c];
        if (s.nreads < 2) continue;
        const double r = (double)(s.nreads - 1) * (double)s.lmax;
        chain = std::max(chain, r); rows += r;
      }
      int64_t cnt[10] = {0}; double sum[10] = {0};
      for (int64_t c = 0; c < n_clusters; ++c) {
        const Shape& s = shp[c];
        if (s.nreads < 2) continue;
        const double r = (double)(s.nreads - 1) * (double)s.lmax;
        const int b = std::min(9, (int)(10.0 * r / std::max(chain, 1.0)));
        cnt[b]++; sum[b] += r;
      }
      for (int b = 0; b < 10; ++b) fprintf(stderr, "[svb-stage] poa: chains of %d0-%d0 %% of the longest: %lld clusters, %.1f %% of all rows\n", b, b + 1, (long long)cnt[b], 100.0 * sum[b] / std::max(rows, 1.0));
      fprintf(stderr, "[svb-stage] poa: row chains: longest %.0f, all %.0f (%.1f x)\n", chain, rows, rows / std::max(chain, 1.0));
    }
    // pass 0: heuristic capacities; pass 1: worst-case capacities for the clusters that overflowed
    std::vector<uint32_t> todo((size_t)n_clusters);
    for (int64_t c = 0; c < n_clusters; ++c) todo[c] = (uint32_t)c;
    std::vector<int4> dims((size_t)n_clusters);
    std::vector<int64_t> foot((size_t)n_clusters);
    float kms = 0.f;
    for (int pass = 0; pass < 2 && !todo.empty(); ++pass) {
      // capacities of every cluster of this pass (nodes, edges, band columns, longest read) and the bytes they carve
      int wmax = 4;
      for (uint32_t c : todo) {
        const Shape& s = shp[c];
        int64_t nc = 4, wc = 4;
        if (s.nreads) {
          const int w = 10 + (int)(0.01 * s.lmax);
          const int diff = s.lmax - s.lmin;
          if (pass == 0) {
            nc = std::min<int64_t>(s.sum + 2, 2 * (int64_t)s.lmax + 32 * s.nreads + 64);
            wc = std::min<int64_t>(s.lmax + 1, 2 * w + 1 + 4 * diff + 96);
          } else {
            nc = s.sum + 2;
            wc = s.lmax + 1;
          }
        }
        const int wcap = (int)((wc + 31) & ~31LL);
        const int64_t ecap = 3 * nc + 64;
        if (nc > 0x3fffffff || ecap > 0x7fffffff) { set_error("cluster %u is too large for the POA workspace", c); rc = SVB_ERANGE; goto done; }
        dims[c] = make_int4((int)nc, (int)ecap, wcap, std::max(1, s.lmax));
        foot[c] = poa_ws_carve(nullptr, (int)nc, (int)ecap, wcap, std::max(1, s.lmax), nullptr);
        wmax = std::max(wmax, wcap);
      }
      // Hand-out order = footprint, largest first (footprint ~ nodes x band ~ work: also the longest-processing-time-first
      // order).  Slot s is sized for the s-th cluster handed out; every cluster handed out later is no bigger than the
      // smallest slot, so it fits whichever warp picks it up -- and the workspace is the sum of the biggest `slots`
      // footprints instead of slots x the biggest one.
      std::stable_sort(todo.begin(), todo.end(), [&](uint32_t a, uint32_t b) { return foot[a] > foot[b]; });
      slog.lap("footprints + order");
      size_t free_b = 0;
      PCHECK(pool_available(&free_b));
      slog.lap("pool_available");
      const char* eb = getenv("SVB_POA_WS_BYTES");
      const int64_t budget = eb ? atoll(eb) : (int64_t)(free_b * 0.8);
      // bit 2048 of the variant: the build with 2 CTAs per SM (255 registers, no spills) -- for batches with fewer clusters than
      // warp slots, where the time is the chain of rows of the biggest cluster and occupancy buys nothing
      int mb = (variant & 8192) ? 3 : (variant & 2048) ? 2 : SVB_POA_MINB;   // bit 8192: 3 CTAs per SM (168 registers), CTA rows only
      // experiment knob: fewer resident CTAs per SM than the build allows (the launch pads its shared memory request)
      int ctas_per_sm = 0;
      if (const char* ec = getenv("SVB_POA_CTAS_PER_SM")) { ctas_per_sm = atoi(ec); if (ctas_per_sm >= 1 && ctas_per_sm < mb) mb = ctas_per_sm; else ctas_per_sm = 0; }
      const bool cta = (variant & 4096) != 0;          // bit 4096: a CTA (four warps) per cluster, the DP rows cut across its warps
      const int per_cta = cta ? 1 : 4;                 // clusters in flight per CTA
      int64_t slots = std::min<int64_t>((int64_t)todo.size(), (int64_t)sms * per_cta * mb);
      std::vector<int64_t> slot_off((size_t)slots + 1, 0);
      {
        int64_t s_ok = 0;
        for (int64_t s_ = 0; s_ < slots; ++s_) {
          if (s_ > 0 && slot_off[(size_t)s_] + foot[todo[(size_t)s_]] > budget) break;
          slot_off[(size_t)s_ + 1] = slot_off[(size_t)s_] + foot[todo[(size_t)s_]];
          s_ok = s_ + 1;
        }
        slots = s_ok;
      }
      // a CTA is four warps: round the slot count up to whole CTAs by repeating the smallest slot size
      while (slots % per_cta) { slot_off.resize((size_t)slots + 2); slot_off[(size_t)slots + 1] = slot_off[(size_t)slots] + foot[todo[(size_t)std::min<int64_t>(slots, (int64_t)todo.size() - 1)]]; ++slots; }
      slot_off.resize((size_t)slots + 1);
      if (slot_off[(size_t)slots] > (int64_t)free_b) { set_error("POA workspace of %lld bytes does not fit", (long long)slot_off[(size_t)slots]); rc = SVB_ENOMEM; goto done; }
      pfree(d_ws, 0); d_ws = nullptr;
      pfree(d_slotoff, 0); d_slotoff = nullptr;
      PCHECK(pmalloc((void**)&d_ws, (size_t)slot_off[(size_t)slots], 0));
      slog.lap("workspace from the pool");
      PCHECK(pmalloc((void**)&d_slotoff, (size_t)(slots + 1) * 8, 0));
      PCHECK(cudaMemcpy(d_slotoff, slot_off.data(), (size_t)(slots + 1) * 8, cudaMemcpyHostToDevice));
      PCHECK(cudaMemcpy(d_order, todo.data(), todo.size() * 4, cudaMemcpyHostToDevice));
      PCHECK(cudaMemcpy(d_dims, dims.data(), (size_t)n_clusters * sizeof(int4), cudaMemcpyHostToDevice));
      PCHECK(cudaMemsetAsync(d_work, 0, 4, 0));
      slog.lap("slot table + order + dims H2D");
      PoaParams P;
      memset(&P, 0, sizeof(P));
      P.seqs = d_seqs; P.seq_offs = d_soff; P.cluster_offs = d_coff; P.order = d_order; P.n = (int)todo.size();
      P.work = d_work; P.ws = d_ws; P.slot_off = d_slotoff; P.dims = d_dims; P.n_slots = (int)slots;
      P.cons = d_cons; P.cons_off = d_capoff; P.cons_len = d_len; P.status = d_status; P.cells = d_cells; P.phase = d_phase;
      P.match = 2; P.mismatch = 4; P.o1 = 4; P.e1 = 2; P.o2 = 24; P.e2 = 1; P.wb = 10; P.wf = 0.01f;  // abpoa_init_para
      cudaEvent_t k0, k1;
      PCHECK(cudaEventCreate(&k0)); PCHECK(cudaEventCreate(&k1));
      PCHECK(cudaEventRecord(k0, 0));
      // variants with the shared-memory copy of the previous row use 2 buffers x 3 arrays x swcap ints per warp
      P.swcap = std::min(wmax, 128);
      const size_t smem = cta ? ((size_t)6 * (size_t)P.swcap + 9 * 4 * 2 + 8) * sizeof(int)
                              : (variant & POA_V_SMEM) ? (size_t)4 * 6 * (size_t)P.swcap * sizeof(int) : 0;
      size_t smem_pad = smem;
      if (ctas_per_sm) smem_pad = std::max(smem, (size_t)(227 * 1024) / (size_t)(ctas_per_sm + 1) + 1024);   // ctas_per_sm fit, one more does not
      const unsigned grid = (unsigned)(slots / per_cta);
#define POA_LAUNCH(VV)                                                                                              \
  case (VV):                                                                                                        \
    if (smem > 48 * 1024) PCHECK(cudaFuncSetAttribute(k_poa<VV, 32>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    k_poa<VV, 32><<<grid, 128, smem>>>(P);                                                                          \
    break;
      switch (variant) {
        POA_LAUNCH(0) POA_LAUNCH(455)
        case 2048 + 1479:
          if (smem > 48 * 1024) PCHECK(cudaFuncSetAttribute(k_poa<1479, 32, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_poa<1479, 32, 2><<<grid, 128, smem>>>(P);
          break;
        case 4096 + 455:
          if (smem > 48 * 1024) PCHECK(cudaFuncSetAttribute(k_poa<455, 32, 4, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_poa<455, 32, 4, 4><<<grid, 128, smem>>>(P);
          break;
        case 8192 + 4096 + 455:
          if (smem > 48 * 1024) PCHECK(cudaFuncSetAttribute(k_poa<455, 32, 3, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
          k_poa<455, 32, 3, 4><<<grid, 128, smem>>>(P);
          break;
        case 4096 + 2048 + 455:
          if (smem_pad > 48 * 1024) PCHECK(cudaFuncSetAttribute(k_poa<455, 32, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pad));
          k_poa<455, 32, 2, 4><<<grid, 128, smem_pad>>>(P);
          break;
        default:
          set_error("SVB_POA_VARIANT=%d is not built (0 = round 1, 455 = a warp per cluster, 3527 = the same with four column groups per step at 2 CTAs per SM, 4551 / 6599 = a CTA per cluster at 4 / 2 CTAs per SM)", variant);
          rc = SVB_EINVAL;
          goto done;
      }
#undef POA_LAUNCH
      PCHECK(cudaGetLastError());
      PCHECK(cudaEventRecord(k1, 0));
      PCHECK(cudaEventSynchronize(k1));
      float ms = 0.f;
      cudaEventElapsedTime(&ms, k0, k1);
      cudaEventDestroy(k0); cudaEventDestroy(k1);
      kms += ms;
      out->launches += 1;
      PCHECK(cudaMemcpy(h_status.data(), d_status, n_clusters * 4, cudaMemcpyDeviceToHost));
      std::vector<uint32_t> again;
      for (uint32_t c : todo) {
        // clamped band (pass 0 only) or capacity overflow: redo with worst-case capacities
        if ((h_status[c] & POA_OVERFLOW) || (pass == 0 && (h_status[c] & POA_CLAMPED))) again.push_back(c);
      }
      if (pass == 1 && !again.empty()) { set_error("POA workspace overflow persisted for %zu clusters", again.size()); rc = SVB_ERANGE; goto done; }
      todo.swap(again);
      out->reruns += (int32_t)todo.size();
    }
    slog.lap("kernel launches + status checks");
    PCHECK(cudaMemcpy(h_len.data(), d_len, n_clusters * 4, cudaMemcpyDeviceToHost));
    PCHECK(cudaMemcpy(h_status.data(), d_status, n_clusters * 4, cudaMemcpyDeviceToHost));
    PCHECK(cudaMemcpy(h_cons.data(), d_cons, cap_off[n_clusters], cudaMemcpyDeviceToHost));
    unsigned long long cells = 0;
    PCHECK(cudaMemcpy(&cells, d_cells, 8, cudaMemcpyDeviceToHost));
    slog.lap("consensus D2H");
    out->cells = (int64_t)cells;
    if (d_phase) {
      unsigned long long ph[5];
      PCHECK(cudaMemcpy(ph, d_phase, 40, cudaMemcpyDeviceToHost));
      const double tot = (double)(ph[0] + ph[1] + ph[2] + ph[3] + ph[4]);
      fprintf(stderr, "[k_poa phases, %% of warp cycles] setup %.1f  dp rows %.1f  traceback %.1f  graph update + re-rank %.1f  consensus %.1f\n",
              100 * ph[0] / tot, 100 * ph[1] / tot, 100 * ph[2] / tot, 100 * ph[3] / tot, 100 * ph[4] / tot);
    }
    out->d2h_bytes = cap_off[n_clusters] + n_clusters * 8;
    for (int64_t c = 0; c < n_clusters; ++c) out->cons_offs[c + 1] = out->cons_offs[c] + h_len[c];
    out->cons = (uint8_t*)malloc((size_t)std::max<int64_t>(out->cons_offs[n_clusters], 1));
    if (!out->cons) { set_error("out of host memory"); rc = SVB_ENOMEM; goto done; }
    for (int64_t c = 0; c < n_clusters; ++c) {
      memcpy(out->cons + out->cons_offs[c], h_cons.data() + cap_off[c], (size_t)h_len[c]);
      out->status[c] = h_status[c];
    }
    PCHECK(cudaEventRecord(e1, 0));
    PCHECK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&out->device_ms, e0, e1);
    out->kernel_ms = kms;
  }
done:
#undef PCHECK
  if (seqs_mem != SVB_MEM_DEVICE) pfree(d_seqs, 0);
  pfree(d_ws, 0); pfree(d_cons, 0); pfree(d_soff, 0); pfree(d_coff, 0); pfree(d_capoff, 0); pfree(d_slotoff, 0); pfree(d_dims, 0);
  pfree(d_order, 0); pfree(d_len, 0); pfree(d_status, 0); pfree(d_work, 0); pfree(d_cells, 0); pfree(d_phase, 0);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (rc != SVB_OK) svb_poa_out_free(out);
  return rc;
}

extern "C" int svb_poa_batch(const uint8_t* seqs, const int64_t* seq_offs, const int64_t* cluster_offs,
                             int64_t n_clusters, int device, svb_poa_out_t* out) {
  return poa_batch_impl(seqs, SVB_MEM_HOST, seq_offs, cluster_offs, n_clusters, device, out);
}

extern "C" void svb_poa_out_free(svb_poa_out_t* out) {
  if (!out) return;
  free(out->cons_offs); free(out->cons); free(out->status);
  out->cons_offs = nullptr; out->cons = nullptr; out->status = nullptr;
}
