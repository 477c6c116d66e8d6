// Caller::pcall (reference caller.cpp:311-406) behind svb_call_batch: clusters in, SV records out.  split_cluster
// (caller.cpp:78-255) looks at sub-read lengths and haplotype tags only, so it runs on the host over the arrays of
// svb_clusters_t; the sub-reads of every job are then cut out of the reads where they live (k_gather_subreads when
// the batch is resident in HBM -- one warp per sub-read, 16 bases per lane and step -- or a host loop for host
// buffers), k_poa builds the consensus of every job, k_ksw_extd2 aligns it to chromosome[s, e], and the CIGAR walk
// of caller.cpp:359-401 (a few ops per job) produces the records.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <thread>
#include <vector>

#include "common.cuh"

namespace svb {

int check_device(int device);
int poa_batch_impl(const uint8_t* seqs, int seqs_mem, const int64_t* seq_offs, const int64_t* cluster_offs, int64_t n_clusters,
                   int device, svb_poa_out_t* out);

// _char26_table (caller.hpp:25-37) restricted to what each input format can hold: codes 0..3 = ACGT, 4 = anything else
__host__ __device__ inline uint8_t code_of_nt6(uint8_t b) { return (b >= 1 && b <= 4) ? (uint8_t)(b - 1) : (uint8_t)4; }
__host__ __device__ inline uint8_t code_of_nt16(uint8_t b) { return b == 1 ? 0 : b == 2 ? 1 : b == 4 ? 2 : b == 8 ? 3 : 4; }
__host__ __device__ inline uint8_t code_of_ascii(uint8_t c) {
  switch (c) {
    case 'A': case 'a': case 0: return 0;
    case 'C': case 'c': case 1: return 1;
    case 'G': case 'g': case 2: return 2;
    case 'T': case 't': case 'U': case 'u': case 3: return 3;
    default: return 4;
  }
}

// sub-read k = reads[src[k], src[k] + len[k]) (nt6 bytes) -> dst[dst_off[k] ...) as codes; one warp per sub-read
template <bool TO_CODE>
__global__ void k_gather_subreads(const uint8_t* __restrict__ reads, const int64_t* __restrict__ src, const int64_t* __restrict__ dst_off,
                                  int64_t n, uint8_t* __restrict__ dst) {
  const int lane = threadIdx.x & 31;
  for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += ((int64_t)gridDim.x * blockDim.x) >> 5) {
    const uint8_t* s = reads + src[k];
    uint8_t* d = dst + dst_off[k];
    const int64_t len = dst_off[k + 1] - dst_off[k];
    for (int64_t i = lane; i < len; i += 32) d[i] = TO_CODE ? code_of_nt6(s[i]) : s[i];
  }
}

namespace {

struct Job {
  std::vector<int> subs;   // global sub-read indices
  int cov, cov0, cov1, cov2;
};

// Cluster::get_len (clusterer.hpp:103-111): unsigned integer mean of the sub-read lengths
inline int job_len(const Job& j, const std::vector<int>& sub_len) {
  unsigned l = 0, n = 0;
  for (int s : j.subs) { ++n; l += (unsigned)sub_len[(size_t)s]; }
  return (int)(l / n);
}

// caller.cpp:78-97
std::vector<Job> split_by_len(const Job& in, const std::vector<int>& sub_len, float min_ratio) {
  std::vector<Job> out;
  for (int s : in.subs) {
    size_t i;
    for (i = 0; i < out.size(); ++i) {
      const float cl = (float)job_len(out[i], sub_len), sl = (float)sub_len[(size_t)s];
      if (std::min(cl, sl) / std::max(cl, sl) >= min_ratio) break;
    }
    if (i == out.size()) { Job j; j.cov = in.cov; j.cov0 = in.cov0; j.cov1 = in.cov1; j.cov2 = in.cov2; out.push_back(j); }
    out[i].subs.push_back(s);
  }
  return out;
}

int largest(const std::vector<Job>& v) {
  unsigned v_max = 0; int i_max = -1;
  for (unsigned i = 0; i < v.size(); ++i) if (v[i].subs.size() > v_max) { v_max = (unsigned)v[i].subs.size(); i_max = (int)i; }
  return i_max;
}

// caller.cpp:100-255: by haplotype tag first (unless --noht), by length inside each haplotype; at most two (three
// never happens: the untagged remainder is only kept when one haplotype is empty) sub-clusters come back
std::vector<Job> split_cluster(const Job& cluster, const std::vector<int>& sub_len, const int32_t* sub_hp, float min_ratio, bool useht) {
  Job c0 = cluster, c1 = cluster, c2 = cluster;
  c0.subs.clear(); c1.subs.clear(); c2.subs.clear();
  for (int s : cluster.subs) {
    if (useht && sub_hp[s] == 1) c1.subs.push_back(s);
    else if (useht && sub_hp[s] == 2) c2.subs.push_back(s);
    else c0.subs.push_back(s);
  }
  c0.cov1 = -1; c0.cov2 = -1; c1.cov0 = -1; c1.cov2 = -1; c2.cov0 = -1; c2.cov1 = -1;
  std::vector<Job> out;
  if (c1.subs.empty() && c2.subs.empty()) {   // no alignment is tagged: the two largest length groups
    std::vector<Job> sub = split_by_len(c0, sub_len, min_ratio);
    int i1 = -1, i2 = -1;
    unsigned v1 = 0, v2 = 0;
    for (unsigned i = 0; i < sub.size(); ++i) {
      if (sub[i].subs.size() > v1) { v2 = v1; i2 = i1; v1 = (unsigned)sub[i].subs.size(); i1 = (int)i; }
      else if (sub[i].subs.size() > v2) { v2 = (unsigned)sub[i].subs.size(); i2 = (int)i; }
    }
    if (i1 != -1) out.push_back(sub[(size_t)i1]);
    if (i2 != -1) out.push_back(sub[(size_t)i2]);
    return out;
  }
  const int both = (c1.subs.empty() ? 0 : 1) + (c2.subs.empty() ? 0 : 2);
  std::vector<Job> sub1 = split_by_len(c1, sub_len, min_ratio), sub2 = split_by_len(c2, sub_len, min_ratio);
  Job fresh; fresh.cov = cluster.cov; fresh.cov0 = cluster.cov0; fresh.cov1 = -1; fresh.cov2 = -1;
  for (int s : c0.subs) {
    const float sl = (float)sub_len[(size_t)s];
    // best_ratio_* are declared int in the reference (caller.cpp:162,172): the ratio truncates to 0 (1 for equal
    // lengths) when stored, and `r > best_ratio` compares against that
    int best_1 = -1, best_ratio_1 = -1, best_2 = -1, best_ratio_2 = -1;
    for (unsigned i = 0; i < sub1.size(); ++i) {
      const float cl = (float)job_len(sub1[i], sub_len), r = std::min(cl, sl) / std::max(cl, sl);
      if (r >= min_ratio && r > (float)best_ratio_1) { best_1 = (int)i; best_ratio_1 = (int)r; }
    }
    for (unsigned i = 0; i < sub2.size(); ++i) {
      const float cl = (float)job_len(sub2[i], sub_len), r = std::min(cl, sl) / std::max(cl, sl);
      if (r >= min_ratio && r > (float)best_ratio_2) { best_2 = (int)i; best_ratio_2 = (int)r; }
    }
    if (both == 1) {
      if (best_1 == -1) fresh.subs.push_back(s);
      else { sub1[(size_t)best_1].subs.push_back(s); ++sub1[(size_t)best_1].cov1; --fresh.cov0; }
    } else if (both == 2) {
      if (best_2 == -1) fresh.subs.push_back(s);
      else { sub2[(size_t)best_2].subs.push_back(s); ++sub2[(size_t)best_2].cov2; --fresh.cov0; }
    } else {
      if (best_1 != -1 && best_ratio_1 > best_ratio_2) { sub1[(size_t)best_1].subs.push_back(s); ++sub1[(size_t)best_1].cov1; --fresh.cov0; }
      else if (best_2 != -1 && best_ratio_2 > best_ratio_1) { sub2[(size_t)best_2].subs.push_back(s); ++sub2[(size_t)best_2].cov2; --fresh.cov0; }
    }
  }
  int i_max = largest(sub1);
  if (i_max != -1) out.push_back(sub1[(size_t)i_max]);
  i_max = largest(sub2);
  if (i_max != -1) out.push_back(sub2[(size_t)i_max]);
  if (both != 3) {
    std::vector<Job> subn = split_by_len(fresh, sub_len, min_ratio);
    i_max = largest(subn);
    if (i_max != -1) {
      if (both == 1) subn[(size_t)i_max].cov1 = -1; else subn[(size_t)i_max].cov2 = -1;
      out.push_back(subn[(size_t)i_max]);
    }
  }
  return out;
}

template <class T>
T* halloc(size_t n) { return (T*)calloc(std::max<size_t>(n, 1), sizeof(T)); }

}  // namespace
}  // namespace svb

using namespace svb;

extern "C" void svb_calls_free(svb_calls_t* o) {
  if (!o) return;
  free(o->job_cluster); free(o->job_cov); free(o->job_sub_offs); free(o->job_sub); free(o->cons_offs); free(o->cons); free(o->score);
  free(o->cigar_offs); free(o->cigar); free(o->sv_job); free(o->sv_type); free(o->sv_pos); free(o->sv_len); free(o->sv_cpos); free(o->job_nv);
  memset(o, 0, sizeof(*o));
}

extern "C" int svb_call_batch(const svb_clusters_t* CL, const svb_seqs_t* RD, const svb_ref_t* R, int min_cluster_weight, int min_sv_length,
                              float min_ratio, int useht, int device, svb_calls_t* out) {
  if (!out) { set_error("svb_call_batch: null out"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  if (!CL || !RD || !R || CL->n_clusters < 0) { set_error("svb_call_batch: bad arguments"); return SVB_EINVAL; }
  if (RD->mem == SVB_MEM_DEVICE && RD->fmt != SVB_SEQ_NT6) { set_error("svb_call_batch: reads resident on the device must be nt6 bytes"); return SVB_EINVAL; }
  if (RD->fmt != SVB_SEQ_NT6 && RD->fmt != SVB_SEQ_BAM4 && RD->fmt != SVB_SEQ_ASCII) { set_error("svb_call_batch: bad read format"); return SVB_EINVAL; }
  SVB_TRY(check_device(device));
  const auto t_host0 = std::chrono::steady_clock::now();
  const int64_t nc = CL->n_clusters;
  const int64_t n_sub = nc ? CL->sub_offs[nc] : 0;
  std::vector<int> sub_len((size_t)n_sub);
  for (int64_t k = 0; k < n_sub; ++k) {
    const int a = CL->sub_aln[k];
    if (a < 0 || a >= RD->n || RD->offs[a] < 0) { set_error("svb_call_batch: sub-read %lld refers to a read without sequence", (long long)k); return SVB_EINVAL; }
    sub_len[(size_t)k] = CL->sub_qe[k] >= CL->sub_qs[k] ? CL->sub_qe[k] - CL->sub_qs[k] + 1 : 0;
  }
  // ---- jobs (caller.cpp:311-330)
  std::vector<Job> jobs;
  std::vector<int32_t> job_cluster;
  for (int64_t c = 0; c < nc; ++c) {
    const int64_t a = CL->sub_offs[c], b = CL->sub_offs[c + 1];
    if (!CL->placed[c] || b - a < min_cluster_weight) continue;
    const int t = CL->tid[c];
    if (t < 0 || t >= R->n_contigs || CL->s[c] < 1 || CL->e[c] < CL->s[c] || CL->e[c] >= R->len[t]) { ++out->skipped_outside; continue; }
    Job cl;
    cl.cov0 = CL->cov0[c]; cl.cov1 = CL->cov1[c]; cl.cov2 = CL->cov2[c]; cl.cov = cl.cov0 + cl.cov1 + cl.cov2;
    for (int64_t k = a; k < b; ++k) cl.subs.push_back((int)k);
    for (Job& j : split_cluster(cl, sub_len, CL->sub_hp, min_ratio, useht != 0)) { jobs.push_back(std::move(j)); job_cluster.push_back((int32_t)c); }
  }
  const int64_t nj = (int64_t)jobs.size();
  out->n_jobs = nj;
  out->job_cluster = halloc<int32_t>((size_t)nj); out->job_cov = halloc<int32_t>((size_t)nj * 4);
  out->job_sub_offs = halloc<int64_t>((size_t)nj + 1); out->job_nv = halloc<int32_t>((size_t)nj);
  out->cons_offs = halloc<int64_t>((size_t)nj + 1); out->score = halloc<int32_t>((size_t)nj); out->cigar_offs = halloc<int64_t>((size_t)nj + 1);
  int64_t n_js = 0;
  for (const Job& j : jobs) n_js += (int64_t)j.subs.size();
  out->job_sub = halloc<int32_t>((size_t)n_js);
  if (!out->job_cluster || !out->job_cov || !out->job_sub_offs || !out->job_nv || !out->cons_offs || !out->score || !out->cigar_offs || !out->job_sub) {
    set_error("out of host memory"); svb_calls_free(out); return SVB_ENOMEM;
  }
  // POA input layout: sequence i of the batch = i-th sub-read in job order
  std::vector<int64_t> seq_offs((size_t)n_js + 1, 0), coffs((size_t)nj + 1, 0), src((size_t)n_js);
  {
    int64_t i = 0;
    for (int64_t j = 0; j < nj; ++j) {
      out->job_cluster[j] = job_cluster[(size_t)j];
      out->job_cov[j * 4] = jobs[(size_t)j].cov; out->job_cov[j * 4 + 1] = jobs[(size_t)j].cov0;
      out->job_cov[j * 4 + 2] = jobs[(size_t)j].cov1; out->job_cov[j * 4 + 3] = jobs[(size_t)j].cov2;
      out->job_sub_offs[j] = i;
      for (int s : jobs[(size_t)j].subs) {
        out->job_sub[i] = s;
        seq_offs[(size_t)i + 1] = seq_offs[(size_t)i] + sub_len[(size_t)s];
        src[(size_t)i] = RD->offs[CL->sub_aln[s]];   // + qs in the format's unit, below
        ++i;
      }
      coffs[(size_t)j + 1] = i;
    }
    out->job_sub_offs[nj] = i;
  }
  float host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
  StageLog slog("call");
  slog.lap("(since the lap clock started)");
  if (nj == 0) { out->host_ms = host_ms; return SVB_OK; }
  const int64_t tot = seq_offs[(size_t)n_js];
  // ---- gather the sub-reads
  uint8_t* d_sub = nullptr;
  std::vector<uint8_t> h_sub;
  cudaEvent_t e0 = nullptr, e1 = nullptr, e2 = nullptr;
  int rc = SVB_OK;
  svb_poa_out_t poa;
  svb_ksw_out_t ez;
  memset(&poa, 0, sizeof(poa)); memset(&ez, 0, sizeof(ez));
#define QCHECK(expr)                                                                        \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      rc = SVB_ECUDA;                                                                       \
      goto done;                                                                            \
    }                                                                                       \
  } while (0)
  {
    QCHECK(cudaEventCreate(&e0)); QCHECK(cudaEventCreate(&e1)); QCHECK(cudaEventCreate(&e2));
    QCHECK(cudaEventRecord(e0, 0));
    if (RD->mem == SVB_MEM_DEVICE) {
      int64_t *d_src = nullptr, *d_off = nullptr;
      for (int64_t i = 0; i < n_js; ++i) src[(size_t)i] += CL->sub_qs[out->job_sub[i]];
      QCHECK(pmalloc((void**)&d_sub, (size_t)std::max<int64_t>(tot, 1), 0));
      cudaError_t e = pmalloc((void**)&d_src, (size_t)n_js * 8, 0);
      if (e == cudaSuccess) e = pmalloc((void**)&d_off, ((size_t)n_js + 1) * 8, 0);
      if (e == cudaSuccess) e = cudaMemcpy(d_src, src.data(), (size_t)n_js * 8, cudaMemcpyHostToDevice);
      if (e == cudaSuccess) e = cudaMemcpy(d_off, seq_offs.data(), ((size_t)n_js + 1) * 8, cudaMemcpyHostToDevice);
      if (e == cudaSuccess) {
        const unsigned grid = (unsigned)std::min<int64_t>((n_js + 7) / 8, 148 * 16);
        k_gather_subreads<true><<<grid, 256>>>(RD->seq, d_src, d_off, n_js, d_sub);
        e = cudaGetLastError();
        out->launches += 1;
      }
      pfree(d_src, 0); pfree(d_off, 0);
      QCHECK(e);
      out->h2d_bytes += n_js * 16 + 8;
    } else {
      h_sub.resize((size_t)std::max<int64_t>(tot, 1));
      // the cut + decode of ~10^4..10^5 sub-reads from the caller's buffer: host threads over contiguous ranges of them
      auto cut = [&](int64_t i0, int64_t i1) {
        for (int64_t i = i0; i < i1; ++i) {
          const int s = out->job_sub[i];
          const int qs = CL->sub_qs[s], len = sub_len[(size_t)s];
          uint8_t* d = h_sub.data() + seq_offs[(size_t)i];
          const uint8_t* base = RD->seq + src[(size_t)i];
          if (RD->fmt == SVB_SEQ_BAM4) for (int k = 0; k < len; ++k) { const int p = qs + k; const uint8_t b = base[p >> 1]; d[k] = code_of_nt16((p & 1) ? (b & 0xf) : (b >> 4)); }
          else if (RD->fmt == SVB_SEQ_NT6) for (int k = 0; k < len; ++k) d[k] = code_of_nt6(base[qs + k]);
          else for (int k = 0; k < len; ++k) d[k] = code_of_ascii(base[qs + k]);
        }
      };
      const int nth = (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)std::thread::hardware_concurrency(), (int64_t)16, tot >> 20}));
      if (nth == 1) cut(0, n_js);
      else {
        std::vector<std::thread> th;
        int64_t a = 0;
        for (int t = 0; t < nth; ++t) {      // ranges of equal bases
          int64_t b = t + 1 == nth ? n_js : (int64_t)(std::upper_bound(seq_offs.begin(), seq_offs.begin() + n_js, tot * (t + 1) / nth) - seq_offs.begin());
          b = std::max(a, std::min(b, n_js));
          th.emplace_back(cut, a, b);
          a = b;
        }
        for (auto& x : th) x.join();
      }
    }
    slog.lap("sub-read gather");
    QCHECK(cudaEventRecord(e1, 0));
    // ---- run_poa for every job (caller.cpp:257-308)
    const auto t_poa0 = std::chrono::steady_clock::now();
    rc = poa_batch_impl(RD->mem == SVB_MEM_DEVICE ? d_sub : h_sub.data(), RD->mem == SVB_MEM_DEVICE ? SVB_MEM_DEVICE : SVB_MEM_HOST, seq_offs.data(),
                        coffs.data(), nj, device, &poa);
    if (rc != SVB_OK) goto done;
    out->poa_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_poa0).count();
    slog.lap("poa_batch_impl");
    out->poa_cells = poa.cells; out->poa_kernel_ms = poa.kernel_ms; out->poa_reruns = poa.reruns; out->launches += poa.launches;
    out->h2d_bytes += poa.h2d_bytes; out->d2h_bytes += poa.d2h_bytes;
    // ---- ksw_extd2 of every consensus against its reference window (caller.cpp:329-355)
    {
      std::vector<int64_t> to((size_t)nj + 1, 0);
      for (int64_t j = 0; j < nj; ++j) { const int c = out->job_cluster[j]; to[(size_t)j + 1] = to[(size_t)j] + (CL->e[c] - CL->s[c] + 1); }
      std::vector<uint8_t> tw((size_t)std::max<int64_t>(to[(size_t)nj], 1));
      if (R->mem == SVB_MEM_DEVICE) {
        // the windows are gathered on the device (the same kernel, bytes as they are) and come back in one copy
        std::vector<int64_t> wsrc((size_t)nj);
        for (int64_t j = 0; j < nj; ++j) { const int c = out->job_cluster[j]; wsrc[(size_t)j] = R->start[CL->tid[c]] + CL->s[c]; }
        int64_t *d_src = nullptr, *d_off = nullptr;
        uint8_t* d_tw = nullptr;
        cudaError_t e = pmalloc((void**)&d_src, (size_t)nj * 8, 0);
        if (e == cudaSuccess) e = pmalloc((void**)&d_off, ((size_t)nj + 1) * 8, 0);
        if (e == cudaSuccess) e = pmalloc((void**)&d_tw, tw.size(), 0);
        if (e == cudaSuccess) e = cudaMemcpy(d_src, wsrc.data(), (size_t)nj * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) e = cudaMemcpy(d_off, to.data(), ((size_t)nj + 1) * 8, cudaMemcpyHostToDevice);
        if (e == cudaSuccess) {
          k_gather_subreads<false><<<(unsigned)std::min<int64_t>((nj + 7) / 8, 148 * 16), 256>>>(R->seq, d_src, d_off, nj, d_tw);
          e = cudaGetLastError();
          out->launches += 1;
        }
        if (e == cudaSuccess) e = cudaMemcpy(tw.data(), d_tw, (size_t)to[(size_t)nj], cudaMemcpyDeviceToHost);
        pfree(d_src, 0); pfree(d_off, 0); pfree(d_tw, 0);
        QCHECK(e);
        out->d2h_bytes += to[(size_t)nj];
      } else {
        for (int64_t j = 0; j < nj; ++j) {
          const int c = out->job_cluster[j];
          memcpy(tw.data() + to[(size_t)j], R->seq + R->start[CL->tid[c]] + CL->s[c], (size_t)(to[(size_t)j + 1] - to[(size_t)j]));
        }
      }
      if (R->fmt == SVB_SEQ_NT6) for (auto& b : tw) b = code_of_nt6(b);
      else for (auto& b : tw) b = code_of_ascii(b);
      std::vector<uint8_t> q((size_t)std::max<int64_t>(poa.cons_offs[nj], 1));
      memcpy(q.data(), poa.cons, (size_t)poa.cons_offs[nj]);
      slog.lap("reference windows + consensus copy");
      const auto t_ksw0 = std::chrono::steady_clock::now();
      rc = svb_ksw_extd2_batch(q.data(), poa.cons_offs, tw.data(), to.data(), nj, 1, -9, -1, 16, 2, 41, 1, device, &ez);   // caller.cpp:333-349
      if (rc != SVB_OK) goto done;
      out->ksw_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_ksw0).count();
      slog.lap("svb_ksw_extd2_batch");
      out->ksw_cells = ez.cells; out->ksw_kernel_ms = ez.kernel_ms; out->ksw_waves = ez.waves; out->launches += ez.launches;
      out->h2d_bytes += ez.h2d_bytes; out->d2h_bytes += ez.d2h_bytes;
    }
    QCHECK(cudaEventRecord(e2, 0));
    QCHECK(cudaEventSynchronize(e2));
    cudaEventElapsedTime(&out->gather_ms, e0, e1);
    cudaEventElapsedTime(&out->device_ms, e0, e2);
  }
  {
    // ---- outputs + the CIGAR walk (caller.cpp:359-401)
    const auto t1 = std::chrono::steady_clock::now();
    out->cons = halloc<uint8_t>((size_t)poa.cons_offs[nj]);
    out->cigar = halloc<uint32_t>((size_t)ez.n_cigar);
    if (!out->cons || !out->cigar) { set_error("out of host memory"); rc = SVB_ENOMEM; goto done; }
    memcpy(out->cons_offs, poa.cons_offs, ((size_t)nj + 1) * 8);
    memcpy(out->cons, poa.cons, (size_t)poa.cons_offs[nj]);
    memcpy(out->score, ez.score, (size_t)nj * 4);
    memcpy(out->cigar_offs, ez.cigar_offs, ((size_t)nj + 1) * 8);
    memcpy(out->cigar, ez.cigar, (size_t)ez.n_cigar * 4);
    std::vector<int32_t> sv_job, sv_pos, sv_len, sv_cpos;
    std::vector<uint8_t> sv_type;
    for (int64_t j = 0; j < nj; ++j) {
      const int c = out->job_cluster[j];
      unsigned rpos = (unsigned)CL->s[c], cpos = 0;
      int nv = 0;
      for (int64_t i = ez.cigar_offs[j]; i < ez.cigar_offs[j + 1]; ++i) {
        const unsigned l = ez.cigar[i] >> 4, op = ez.cigar[i] & 0xf;
        if (op == 0) { rpos += l; cpos += l; }
        else if (op == 1) {
          if (l >= (unsigned)min_sv_length) { sv_job.push_back((int32_t)j); sv_type.push_back(0); sv_pos.push_back((int32_t)rpos); sv_len.push_back((int32_t)l); sv_cpos.push_back((int32_t)cpos); ++nv; }
          cpos += l;
        } else {
          if (l >= (unsigned)min_sv_length) { sv_job.push_back((int32_t)j); sv_type.push_back(1); sv_pos.push_back((int32_t)rpos); sv_len.push_back((int32_t)l); sv_cpos.push_back((int32_t)cpos); ++nv; }
          rpos += l;
        }
      }
      out->job_nv[j] = nv;
    }
    out->n_svs = (int64_t)sv_job.size();
    out->sv_job = halloc<int32_t>(sv_job.size()); out->sv_type = halloc<uint8_t>(sv_job.size()); out->sv_pos = halloc<int32_t>(sv_job.size());
    out->sv_len = halloc<int32_t>(sv_job.size()); out->sv_cpos = halloc<int32_t>(sv_job.size());
    if (!out->sv_job || !out->sv_type || !out->sv_pos || !out->sv_len || !out->sv_cpos) { set_error("out of host memory"); rc = SVB_ENOMEM; goto done; }
    if (!sv_job.empty()) {
      memcpy(out->sv_job, sv_job.data(), sv_job.size() * 4); memcpy(out->sv_type, sv_type.data(), sv_type.size());
      memcpy(out->sv_pos, sv_pos.data(), sv_pos.size() * 4); memcpy(out->sv_len, sv_len.data(), sv_len.size() * 4);
      memcpy(out->sv_cpos, sv_cpos.data(), sv_cpos.size() * 4);
    }
    host_ms += std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t1).count();
    out->host_ms = host_ms;
  }
done:
#undef QCHECK
  pfree(d_sub, 0);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (e2) cudaEventDestroy(e2);
  svb_poa_out_free(&poa);
  svb_ksw_out_free(&ez);
  if (rc != SVB_OK) svb_calls_free(out);
  return rc;
}
