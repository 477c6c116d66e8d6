// Clusterer::run (reference clusterer.cpp:8-52) behind svb_cluster_batch: the step between the SFS table of `search`
// and Caller::pcall.  extend_alignment (per read) and fill_clusters (per cluster) are kernels over the per-item
// functions of cluster_core.cuh; cluster_by_proximity is the sequential host sweep of cluster_host.hpp between them.
// Both kernels are thin: a thread walks a CIGAR of a handful of ops (smoothed reads) or a few hundred (raw reads)
// and touches <= 2 x 107 reference bytes per SFS; the grid is the number of reads with SFSs / of clusters, the
// data per item is tiny, so neither is anywhere near a bandwidth or issue bound -- what matters is that the stage
// no longer runs on (and waits for) host cores between two GPU stages.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <vector>

#include "common.cuh"

#include "cluster_core.cuh"
#include "cluster_host.hpp"

namespace svb {

int check_device(int device);

struct ClRefDev {
  const uint8_t* seq;
  const int64_t* start;
  const int64_t* len;
  int64_t n_contigs;
};

// bam_endpos of every record, and the longest reference span of any (the host's region queries need only that)
__global__ void k_cl_endpos(const int64_t* __restrict__ cigar_offs, const uint32_t* __restrict__ cigar, const int32_t* __restrict__ pos,
                            int64_t n, int32_t* __restrict__ endp, unsigned* __restrict__ max_span) {
  const int64_t a = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int span = 0;
  if (a < n) {
    const int e = cl_endpos(cigar + cigar_offs[a], (int)(cigar_offs[a + 1] - cigar_offs[a]), pos[a]);
    endp[a] = e;
    span = e - pos[a];
  }
  span = __reduce_max_sync(0xffffffffu, span);
  if ((threadIdx.x & 31) == 0 && span > 0) atomicMax(max_span, (unsigned)span);
}

// one thread per accepted read (a read that carries SFSs)
__global__ void k_cl_extend(const int32_t* __restrict__ accepted, int n_acc, const int32_t* __restrict__ tid, const int32_t* __restrict__ pos,
                            const int32_t* __restrict__ endp, const int64_t* __restrict__ cigar_offs, const uint32_t* __restrict__ cigar,
                            const int64_t* __restrict__ sfs_offs, const int32_t* __restrict__ sfs_qs, const int32_t* __restrict__ sfs_len,
                            ClRefDev ref, int flank, int ksize, int clipped, ClExt* __restrict__ ext, int32_t* __restrict__ n_ext,
                            unsigned* __restrict__ cnt, int32_t* __restrict__ clip) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_acc) return;
  const int a = accepted[i];
  const int t = tid[a];
  unsigned c4[4] = {0, 0, 0, 0};
  int cl4[4] = {0, 0, 0, 0};
  int n = 0;
  if (t >= 0 && t < ref.n_contigs && ref.len[t] >= 0) {      // clusterer.cpp:162-163: a chromosome without sequence is skipped
    ClAln A;
    A.cig = cigar + cigar_offs[a]; A.n_cig = (int)(cigar_offs[a + 1] - cigar_offs[a]); A.pos = pos[a];
    A.chrom = ref.seq + ref.start[t]; A.chrom_len = ref.len[t];
    const int64_t s0 = sfs_offs[a];
    n = cl_extend_read(A, sfs_qs + s0, sfs_len + s0, (int)(sfs_offs[a + 1] - s0), flank, ksize, clipped != 0, endp[a], ext + s0, c4, cl4);
  }
  n_ext[i] = n;
  for (int k = 0; k < 4; ++k) if (c4[k]) atomicAdd(cnt + k, c4[k]);
  if (clip) for (int k = 0; k < 4; ++k) clip[(int64_t)a * 4 + k] = cl4[k];
}

struct ClFillDev {
  const int32_t *min_s, *max_e, *lo, *hi;
  const int64_t *mem_offs, *rv_offs;
  const int32_t* members;
};

// one thread per cluster with enough reads
__global__ void k_cl_fill(int n_cl, ClFillDev d, const int32_t* __restrict__ pos, const int32_t* __restrict__ endp, const int32_t* __restrict__ hp,
                          const int64_t* __restrict__ cigar_offs, const uint32_t* __restrict__ cigar, int32_t* __restrict__ sub_aln,
                          int32_t* __restrict__ sub_qs, int32_t* __restrict__ sub_qe, int32_t* __restrict__ sub_hp, int32_t* __restrict__ n_sub,
                          uint8_t* __restrict__ rvec, int32_t* __restrict__ n_rv, int32_t* __restrict__ cov, unsigned* __restrict__ cnt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_cl) return;
  const int64_t m0 = d.mem_offs[c];
  int ns, nr, cv[3];
  unsigned unext = 0;
  cl_fill_cluster(pos, endp, hp, cigar_offs, cigar, d.lo[c], d.hi[c], d.min_s[c], d.max_e[c], d.members + m0, (int)(d.mem_offs[c + 1] - m0),
                  sub_aln + m0, sub_qs + m0, sub_qe + m0, sub_hp + m0, ns, rvec + d.rv_offs[c], nr, cv, unext);
  n_sub[c] = ns; n_rv[c] = nr;
  cov[c * 3] = cv[0]; cov[c * 3 + 1] = cv[1]; cov[c * 3 + 2] = cv[2];
  if (unext) atomicAdd(cnt + 4, unext);
}

namespace {

struct DevArena {   // every device allocation of one call, released together
  std::vector<void*> ptrs;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  ~DevArena() {
    for (void* p : ptrs) pfree(p, stream);
    for (cudaEvent_t e : ev) if (e) cudaEventDestroy(e);
    if (stream) cudaStreamDestroy(stream);
  }
  template <class T>
  cudaError_t alloc(T** p, size_t n) {
    *p = nullptr;
    cudaError_t e = pmalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T), stream);
    if (e == cudaSuccess) ptrs.push_back(*p);
    return e;
  }
  template <class T>
  cudaError_t upload(T** p, const T* h, size_t n) {
    cudaError_t e = alloc(p, n);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(*p, h, n * sizeof(T), cudaMemcpyHostToDevice, stream);
    return e;
  }
};

template <class T>
T* host_alloc(size_t n) { return cl_host_alloc<T>(n); }

}  // namespace
}  // namespace svb

using namespace svb;

extern "C" void svb_clusters_free(svb_clusters_t* o) {
  if (!o) return;
  free(o->tid); free(o->s); free(o->e); free(o->cov0); free(o->cov1); free(o->cov2); free(o->placed); free(o->sub_offs);
  free(o->sub_aln); free(o->sub_qs); free(o->sub_qe); free(o->sub_hp); free(o->rvec_offs); free(o->rvec); free(o->clip);
  memset(o, 0, sizeof(*o));
}

extern "C" int svb_cluster_batch(const svb_alns_t* A, const svb_ref_t* R, int threads, int min_cluster_weight, int flank, int ksize,
                                 int clipped, int device, svb_clusters_t* out) {
  if (!out) { set_error("svb_cluster_batch: null out"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  if (!A || !R || A->n_aln < 0 || A->n_aln > 0x7ffffff0 || R->n_contigs < 0) { set_error("svb_cluster_batch: bad arguments"); return SVB_EINVAL; }
  if (flank < 0 || flank > CL_FLANK_MAX || ksize < 1 || ksize > CL_KSIZE_MAX) { set_error("svb_cluster_batch: flank must be <= %d and ksize <= %d", CL_FLANK_MAX, CL_KSIZE_MAX); return SVB_EINVAL; }
  if (R->fmt != SVB_SEQ_ASCII && R->fmt != SVB_SEQ_NT6) { set_error("svb_cluster_batch: reference format must be SVB_SEQ_ASCII or SVB_SEQ_NT6"); return SVB_EINVAL; }
  SVB_TRY(check_device(device));
  const int64_t n = A->n_aln;
  if (n && (!A->tid || !A->pos || !A->hp || !A->cigar_offs || !A->sfs_offs)) { set_error("svb_cluster_batch: null alignment arrays"); return SVB_EINVAL; }
  std::vector<int32_t> accepted;
  for (int64_t a = 0; a < n; ++a) {
    if (a && (A->tid[a] < A->tid[a - 1] || (A->tid[a] == A->tid[a - 1] && A->pos[a] < A->pos[a - 1]))) {
      set_error("svb_cluster_batch: alignments must be in coordinate order (record %lld is not): fill_clusters needs region queries", (long long)a);
      return SVB_EINVAL;
    }
    if (A->cigar_offs[a + 1] < A->cigar_offs[a] || A->sfs_offs[a + 1] < A->sfs_offs[a]) { set_error("svb_cluster_batch: offsets must be non-decreasing"); return SVB_EINVAL; }
    if (A->sfs_offs[a + 1] > A->sfs_offs[a]) accepted.push_back((int32_t)a);
  }
  const int64_t n_cig = n ? A->cigar_offs[n] : 0, n_sfs = n ? A->sfs_offs[n] : 0;
  if ((n && (A->cigar_offs[0] != 0 || A->sfs_offs[0] != 0)) || n_sfs > 0x7ffffff0) { set_error("svb_cluster_batch: offsets must start at 0"); return SVB_EINVAL; }
  int64_t ref_bytes = 0;
  for (int64_t c = 0; c < R->n_contigs; ++c) {
    if (R->len[c] < 0) continue;            // a chromosome of the BAM header without sequence (clusterer.cpp:162-163)
    if (R->start[c] < 0) { set_error("svb_cluster_batch: bad contig %lld", (long long)c); return SVB_EINVAL; }
    ref_bytes = std::max(ref_bytes, R->start[c] + R->len[c]);
  }
  out->n_clusters = 0;
  if (accepted.empty()) {
    out->sub_offs = host_alloc<int64_t>(1); out->rvec_offs = host_alloc<int64_t>(1);
    if (clipped) out->clip = host_alloc<int32_t>((size_t)n * 4);
    return SVB_OK;
  }
  DevArena D;
#define CCHECK(expr)                                                                        \
  do {                                                                                      \
    cudaError_t _e = (expr);                                                                \
    if (_e != cudaSuccess) {                                                                \
      set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      svb_clusters_free(out);                                                               \
      return SVB_ECUDA;                                                                     \
    }                                                                                       \
  } while (0)
  CCHECK(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
  for (auto& e : D.ev) CCHECK(cudaEventCreate(&e));
  cudaStream_t st = D.stream;
  CCHECK(cudaEventRecord(D.ev[0], st));
  int32_t *d_tid, *d_pos, *d_hp, *d_qs, *d_len, *d_acc, *d_endp, *d_next, *d_clip = nullptr;
  int64_t *d_coff, *d_soff, *d_rstart, *d_rlen;
  uint32_t* d_cig;
  unsigned* d_cnt;
  ClExt* d_ext;
  const uint8_t* d_ref = R->seq;
  CCHECK(D.upload(&d_tid, A->tid, (size_t)n));
  CCHECK(D.upload(&d_pos, A->pos, (size_t)n));
  CCHECK(D.upload(&d_hp, A->hp, (size_t)n));
  CCHECK(D.upload(&d_coff, A->cigar_offs, (size_t)n + 1));
  CCHECK(D.upload(&d_cig, A->cigar, (size_t)n_cig));
  CCHECK(D.upload(&d_soff, A->sfs_offs, (size_t)n + 1));
  CCHECK(D.upload(&d_qs, A->sfs_qs, (size_t)n_sfs));
  CCHECK(D.upload(&d_len, A->sfs_len, (size_t)n_sfs));
  CCHECK(D.upload(&d_acc, accepted.data(), accepted.size()));
  CCHECK(D.upload(&d_rstart, R->start, (size_t)R->n_contigs));
  CCHECK(D.upload(&d_rlen, R->len, (size_t)R->n_contigs));
  out->h2d_bytes = n * 12 + (n + 1) * 16 + n_cig * 4 + n_sfs * 8 + (int64_t)accepted.size() * 4 + R->n_contigs * 16;
  if (R->mem == SVB_MEM_HOST) {
    uint8_t* p;
    CCHECK(D.upload(&p, R->seq, (size_t)ref_bytes));
    d_ref = p;
    out->h2d_bytes += ref_bytes;
  }
  CCHECK(D.alloc(&d_endp, (size_t)n));
  CCHECK(D.alloc(&d_next, accepted.size()));
  CCHECK(D.alloc(&d_ext, (size_t)n_sfs));
  CCHECK(D.alloc(&d_cnt, 8));
  CCHECK(cudaMemsetAsync(d_cnt, 0, 8 * sizeof(unsigned), st));
  if (clipped) { CCHECK(D.alloc(&d_clip, (size_t)n * 4)); CCHECK(cudaMemsetAsync(d_clip, 0, (size_t)n * 16, st)); }
  const int n_acc = (int)accepted.size();
  CCHECK(cudaEventRecord(D.ev[1], st));
  k_cl_endpos<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(d_coff, d_cig, d_pos, n, d_endp, d_cnt + 5);
  ClRefDev rd{d_ref, d_rstart, d_rlen, R->n_contigs};
  k_cl_extend<<<(unsigned)((n_acc + 63) / 64), 64, 0, st>>>(d_acc, n_acc, d_tid, d_pos, d_endp, d_coff, d_cig, d_soff, d_qs, d_len, rd, flank, ksize,
                                                           clipped, d_ext, d_next, d_cnt, d_clip);
  CCHECK(cudaGetLastError());
  CCHECK(cudaEventRecord(D.ev[2], st));
  out->launches = 2;
  std::vector<ClExt> h_ext((size_t)n_sfs);
  std::vector<int32_t> h_next((size_t)n_acc);
  unsigned h_span = 0;
  CCHECK(cudaMemcpyAsync(h_ext.data(), d_ext, (size_t)n_sfs * sizeof(ClExt), cudaMemcpyDeviceToHost, st));
  CCHECK(cudaMemcpyAsync(h_next.data(), d_next, (size_t)n_acc * 4, cudaMemcpyDeviceToHost, st));
  CCHECK(cudaMemcpyAsync(&h_span, d_cnt + 5, 4, cudaMemcpyDeviceToHost, st));
  if (clipped) {
    out->clip = host_alloc<int32_t>((size_t)n * 4);
    if (!out->clip) { set_error("out of host memory"); return SVB_ENOMEM; }
    CCHECK(cudaMemcpyAsync(out->clip, d_clip, (size_t)n * 16, cudaMemcpyDeviceToHost, st));
  }
  CCHECK(cudaStreamSynchronize(st));
  float ms = 0.f, kms = 0.f;
  cudaEventElapsedTime(&ms, D.ev[1], D.ev[2]);
  kms += ms;
  out->d2h_bytes = n_sfs * (int64_t)sizeof(ClExt) + n_acc * 4 + 4 + (clipped ? n * 16 : 0);
  // ---- cluster_by_proximity on the host
  const auto h0 = std::chrono::steady_clock::now();
  ClPlan P;
  cl_plan_fill(accepted.data(), n_acc, h_next.data(), h_ext.data(), A->sfs_offs, A->tid, A->pos, (int)h_span, n, R->name_rank, threads,
               min_cluster_weight, P);
  const std::vector<int32_t>&f_cluster = P.f_cluster, &f_min_s = P.f_min_s, &f_max_e = P.f_max_e, &f_lo = P.f_lo, &f_hi = P.f_hi, &f_members = P.f_members;
  const std::vector<int64_t>&f_moff = P.f_moff, &f_rvoff = P.f_rvoff;
  out->host_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - h0).count();
  // ---- fill_clusters
  const int nf = (int)f_cluster.size();
  std::vector<int32_t> h_sub_aln(f_members.size()), h_sub_qs(f_members.size()), h_sub_qe(f_members.size()), h_sub_hp(f_members.size());
  std::vector<int32_t> h_nsub((size_t)nf), h_nrv((size_t)nf), h_cov((size_t)nf * 3);
  std::vector<uint8_t> h_rvec((size_t)f_rvoff.back());
  unsigned h_cnt[8] = {0};
  if (nf) {
    ClFillDev fd;
    int32_t *p_min, *p_max, *p_lo, *p_hi, *p_mem, *d_sa, *d_sq, *d_se, *d_sh, *d_nsub, *d_nrv, *d_cov;
    int64_t *p_moff, *p_rvoff;
    uint8_t* d_rvec;
    CCHECK(D.upload(&p_min, f_min_s.data(), (size_t)nf)); CCHECK(D.upload(&p_max, f_max_e.data(), (size_t)nf));
    CCHECK(D.upload(&p_lo, f_lo.data(), (size_t)nf)); CCHECK(D.upload(&p_hi, f_hi.data(), (size_t)nf));
    CCHECK(D.upload(&p_mem, f_members.data(), f_members.size()));
    CCHECK(D.upload(&p_moff, f_moff.data(), (size_t)nf + 1)); CCHECK(D.upload(&p_rvoff, f_rvoff.data(), (size_t)nf + 1));
    out->h2d_bytes += (int64_t)nf * 32 + (int64_t)f_members.size() * 4;
    CCHECK(D.alloc(&d_sa, f_members.size())); CCHECK(D.alloc(&d_sq, f_members.size()));
    CCHECK(D.alloc(&d_se, f_members.size())); CCHECK(D.alloc(&d_sh, f_members.size()));
    CCHECK(D.alloc(&d_nsub, (size_t)nf)); CCHECK(D.alloc(&d_nrv, (size_t)nf)); CCHECK(D.alloc(&d_cov, (size_t)nf * 3));
    CCHECK(D.alloc(&d_rvec, (size_t)f_rvoff.back()));
    fd.min_s = p_min; fd.max_e = p_max; fd.lo = p_lo; fd.hi = p_hi; fd.mem_offs = p_moff; fd.rv_offs = p_rvoff; fd.members = p_mem;
    CCHECK(cudaEventRecord(D.ev[1], st));
    k_cl_fill<<<(unsigned)((nf + 63) / 64), 64, 0, st>>>(nf, fd, d_pos, d_endp, d_hp, d_coff, d_cig, d_sa, d_sq, d_se, d_sh, d_nsub, d_rvec,
                                                         d_nrv, d_cov, d_cnt);
    CCHECK(cudaGetLastError());
    CCHECK(cudaEventRecord(D.ev[2], st));
    out->launches += 1;
    CCHECK(cudaMemcpyAsync(h_sub_aln.data(), d_sa, f_members.size() * 4, cudaMemcpyDeviceToHost, st));
    CCHECK(cudaMemcpyAsync(h_sub_qs.data(), d_sq, f_members.size() * 4, cudaMemcpyDeviceToHost, st));
    CCHECK(cudaMemcpyAsync(h_sub_qe.data(), d_se, f_members.size() * 4, cudaMemcpyDeviceToHost, st));
    CCHECK(cudaMemcpyAsync(h_sub_hp.data(), d_sh, f_members.size() * 4, cudaMemcpyDeviceToHost, st));
    CCHECK(cudaMemcpyAsync(h_nsub.data(), d_nsub, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
    CCHECK(cudaMemcpyAsync(h_nrv.data(), d_nrv, (size_t)nf * 4, cudaMemcpyDeviceToHost, st));
    CCHECK(cudaMemcpyAsync(h_cov.data(), d_cov, (size_t)nf * 12, cudaMemcpyDeviceToHost, st));
    if (!h_rvec.empty()) CCHECK(cudaMemcpyAsync(h_rvec.data(), d_rvec, h_rvec.size(), cudaMemcpyDeviceToHost, st));
    out->d2h_bytes += (int64_t)f_members.size() * 16 + (int64_t)nf * 20 + (int64_t)h_rvec.size();
  }
  CCHECK(cudaMemcpyAsync(h_cnt, d_cnt, sizeof(h_cnt), cudaMemcpyDeviceToHost, st));
  CCHECK(cudaEventRecord(D.ev[3], st));
  CCHECK(cudaStreamSynchronize(st));
  if (nf) { cudaEventElapsedTime(&ms, D.ev[1], D.ev[2]); kms += ms; }
  cudaEventElapsedTime(&out->device_ms, D.ev[0], D.ev[3]);
  out->device_ms -= out->host_ms;   // the host sweep sits between the two halves of the stream's work
  out->kernel_ms = kms;
#undef CCHECK
  out->unplaced = h_cnt[0]; out->s_unplaced = h_cnt[1]; out->e_unplaced = h_cnt[2]; out->unknown = h_cnt[3]; out->unextended = h_cnt[4];
  // ---- assemble
  if (!cl_assemble(P, h_nsub.data(), h_sub_aln.data(), h_sub_qs.data(), h_sub_qe.data(), h_sub_hp.data(), h_nrv.data(), h_rvec.data(), h_cov.data(),
                   min_cluster_weight, out)) {
    set_error("out of host memory");
    svb_clusters_free(out);
    return SVB_ENOMEM;
  }
  return SVB_OK;
}

extern "C" int svb_index_ref(const svb_index_t* idx, svb_ref_t* out, int64_t* start_buf, int64_t* len_buf) {
  if (!idx || !out || !start_buf || !len_buf) { set_error("svb_index_ref: null argument"); return SVB_EINVAL; }
  const IndexDev& d = idx->dev;
  if (!d.d_text || !d.d_tstart) { set_error("svb_index_ref: this index carries no text (built from a bare BWT)"); return SVB_EINVAL; }
  SVB_TRY(check_device(d.device));
  std::vector<int64_t> ts((size_t)d.n_contigs + 1);
  SVB_CUDA(cudaMemcpy(ts.data(), d.d_tstart, ts.size() * 8, cudaMemcpyDeviceToHost));
  for (int64_t c = 0; c < d.n_contigs; ++c) { start_buf[c] = ts[(size_t)c]; len_buf[c] = (ts[(size_t)c + 1] - ts[(size_t)c] - 2) / 2; }
  memset(out, 0, sizeof(*out));
  out->n_contigs = d.n_contigs; out->seq = d.d_text; out->start = start_buf; out->len = len_buf; out->name_rank = nullptr;
  out->fmt = SVB_SEQ_NT6; out->mem = SVB_MEM_DEVICE;
  return SVB_OK;
}
