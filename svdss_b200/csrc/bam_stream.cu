// svb_bamstream_*: the BAM loader of `search` (PingPong::load_batch_bam, ping_pong.cpp:53-131, and the filters of
// :66-79,196-203) on the device.  The host only reads the file and finds the BGZF members (host/io.hpp); a window of
// members is inflated into HBM (k_bgzf_inflate_warp), the records are walked where they lie (a record may start in one
// window and end in the next: the unfinished tail is carried over), every record is parsed by a thread (core fields,
// aux walk for XF / HP), and the bases of the reads that will be searched are decoded nt16 -> nt6 straight into the
// device batch svb_sfs_resident runs on.  What crosses PCIe: the compressed file one way, names / flags / tags the other.
#include <algorithm>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "bam_core.cuh"

namespace svb {

int check_device(int device);
void launch_inflate(const uint8_t* d_in, const int64_t* d_io, const int64_t* d_oo, int64_t n_members, uint8_t* d_out, int32_t* d_st, int device,
                    cudaStream_t sq);   // bgzf_inflate.cu

// record offsets of a window: one thread follows the block_size fields (each depends on the one before)
__global__ void k_bam_walk(const uint8_t* __restrict__ win, int64_t start, int64_t total, int64_t* __restrict__ rec_off, int64_t cap,
                           int64_t* __restrict__ res) {
  int64_t end = start;
  int flag = 0;
  const int64_t n = bam_chase(win, start, total, total, rec_off, cap, &end, &flag);
  for (int64_t i = 0; i < n; ++i) rec_off[i] += 4;   // from the block_size field to the record
  res[0] = n; res[1] = end; res[2] = flag == 2;
}

// ---- the same walk in parallel.  The chain of block_size fields is serial (11 ms for the 16 k records of a window:
// every hop is a round trip to L2), so the window is cut into segments and a warp per segment GUESSES where a record
// starts in it (first offset that looks like a record whose successor looks like one too: the lanes test 32 offsets at
// a time) and lane 0 follows the chain to the segment's end.  One CTA then links the segments: entering segment s at
// the true position, thread 0 looks that position up in the segment's guessed chain -- found: the rest of the chain is
// the truth (the walk is deterministic from any true record start) and everybody copies it; not found: it follows the
// fields itself until it meets the chain or leaves the segment.  The guess only decides how fast the walk is, never
// what it returns (bam_core.cuh; tests/test_bam_emul.py plants fake records in the payload).
__global__ void __launch_bounds__(128) k_bam_walk_seg(const uint8_t* __restrict__ win, int64_t start, int64_t total, int64_t seg_len, int n_seg, int n_ref,
                                                      int64_t* __restrict__ seg_pos, int64_t seg_cap, int64_t* __restrict__ seg_cnt,
                                                      int64_t* __restrict__ seg_end_pos, int* __restrict__ seg_end_flag) {
  const int s = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = (int)(threadIdx.x & 31);
  if (s >= n_seg) return;
  const int64_t a = start + (int64_t)s * seg_len, b = min(total, a + seg_len);
  int64_t p = -1;
  if (s == 0) p = a;
  else
    for (int64_t q0 = a; q0 < b; q0 += 32) {
      const int64_t q = q0 + lane;
      const unsigned m = __ballot_sync(0xffffffffu, q < b && bam_guess(win, q, total, n_ref));
      if (m) { p = q0 + (__ffs((int)m) - 1); break; }
    }
  if (lane != 0) return;
  int64_t n = 0, end = p;
  int flag = 0;
  if (p >= 0) n = bam_chase(win, p, b, total, seg_pos + (int64_t)s * seg_cap, seg_cap, &end, &flag);
  seg_cnt[s] = n; seg_end_pos[s] = end; seg_end_flag[s] = flag;
}

__global__ void __launch_bounds__(256) k_bam_walk_link(const uint8_t* __restrict__ win, int64_t start, int64_t total, int64_t seg_len, int n_seg,
                                                       const int64_t* __restrict__ seg_pos, int64_t seg_cap, const int64_t* __restrict__ seg_cnt,
                                                       const int64_t* __restrict__ seg_end_pos, const int* __restrict__ seg_end_flag,
                                                       int64_t* __restrict__ rec_off, int64_t cap, int64_t* __restrict__ res) {
  __shared__ int64_t sh_cur, sh_n, sh_join;
  __shared__ int sh_stop, sh_err, sh_joined;
  if (threadIdx.x == 0) { sh_cur = start; sh_n = 0; sh_stop = 0; sh_err = 0; sh_joined = 0; }
  __syncthreads();
  for (int s = 0; s < n_seg; ++s) {
    if (sh_stop) break;                               // uniform: written before the barrier at the end of the last round
    const int64_t a = start + (int64_t)s * seg_len, b = min(total, a + seg_len);
    const int64_t* chain = seg_pos + (int64_t)s * seg_cap;
    const int64_t c = seg_cnt[s];
    if (threadIdx.x == 0) {
      int64_t cur = sh_cur, n = sh_n, join = -1;
      int err = 0;
      if (!bam_link_segment(win, b, total, chain, c, seg_end_pos[s], seg_end_flag[s], rec_off, cap, &cur, &n, &join, &err)) sh_stop = 1;
      if (err) sh_err = 1;
      if (join >= 0) ++sh_joined;
      sh_cur = cur; sh_n = n; sh_join = join;
    }
    __syncthreads();
    if (sh_join >= 0) {
      const int64_t lo = sh_join, n0 = sh_n;
      for (int64_t k = lo + threadIdx.x; k < c; k += blockDim.x) rec_off[n0 + (k - lo)] = chain[k] + 4;
      __syncthreads();
      if (threadIdx.x == 0) sh_n = n0 + (c - lo);
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) { res[0] = sh_n; res[1] = sh_cur; res[2] = sh_err; res[3] = sh_joined; }
}

// one thread per record (bam_parse_record)
__global__ void k_bam_parse(const uint8_t* __restrict__ win, const int64_t* __restrict__ rec_off, int64_t n, int putative,
                            BamMeta* __restrict__ meta, int64_t* __restrict__ seq_off, BamAln* __restrict__ aln, int* __restrict__ err) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  BamMeta m;
  BamAln al;
  int64_t rel = 0;
  if (!bam_parse_record(win + rec_off[i], putative, &m, &rel, aln ? &al : nullptr)) atomicExch(err, 1);
  if (aln) aln[i] = al;
  seq_off[i] = rec_off[i] + rel;
  meta[i] = m;
}

// svb_bamstream_fetch: CIGAR words and packed bases of the selected records of the window, back to back
__global__ void __launch_bounds__(128) k_bam_fetch(const uint8_t* __restrict__ win, const int64_t* __restrict__ rec_off, const int64_t* __restrict__ seq_off,
                                                   const BamMeta* __restrict__ meta, const BamAln* __restrict__ aln, const int64_t* __restrict__ sel,
                                                   const int64_t* __restrict__ cig_offs, const int64_t* __restrict__ seq_offs, uint32_t* __restrict__ cig,
                                                   uint8_t* __restrict__ seq) {
  const int64_t i = sel[blockIdx.x];
  const uint8_t* c = win + rec_off[i] + aln[i].cigar_rel;
  uint32_t* co = cig + cig_offs[blockIdx.x];
  for (int k = threadIdx.x; k < aln[i].n_cigar; k += blockDim.x) co[k] = bam_ld32(c + 4 * (int64_t)k);
  const uint8_t* q = win + seq_off[i];
  uint8_t* so = seq + seq_offs[blockIdx.x];
  const int nb = (meta[i].l_qseq + 1) / 2;
  for (int k = threadIdx.x; k < nb; k += blockDim.x) so[k] = q[k];
}

// exclusive sums over the records of a window: name bytes of the records the host hears about, bases and count of the searched ones
struct BamSums { long long names, bases, reads; };
__global__ void __launch_bounds__(1024) k_bam_scan(const BamMeta* __restrict__ meta, int64_t n, int64_t* __restrict__ name_off,
                                                   int64_t* __restrict__ base_off, int64_t* __restrict__ rank, BamSums* __restrict__ tot) {
  __shared__ long long sh[3][32];
  __shared__ long long carry[3];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (threadIdx.x < 3) carry[threadIdx.x] = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024) {
    const int64_t i = base + threadIdx.x;
    long long v[3] = {0, 0, 0};
    if (i < n) {
      const BamMeta m = meta[i];
      v[0] = m.name_len;
      if (m.state == 2) { v[1] = m.l_qseq; v[2] = 1; }
    }
    long long inc[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      long long x = v[k];
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
      inc[k] = x;
      if (lane == 31) sh[k][wid] = x;
    }
    __syncthreads();
    if (wid == 0) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        long long x = sh[k][lane];
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) { const long long y = __shfl_up_sync(0xffffffffu, x, d); if (lane >= d) x += y; }
        sh[k][lane] = x;   // inclusive over the warps
      }
    }
    __syncthreads();
    long long pre[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) pre[k] = carry[k] + (wid ? sh[k][wid - 1] : 0) + inc[k] - v[k];
    if (i < n) { name_off[i] = pre[0]; base_off[i] = pre[1]; rank[i] = pre[2]; }
    __syncthreads();
    if (threadIdx.x < 3) carry[threadIdx.x] += sh[threadIdx.x][31];
    __syncthreads();
  }
  if (threadIdx.x == 0) { tot->names = carry[0]; tot->bases = carry[1]; tot->reads = carry[2]; name_off[n] = carry[0]; }
}

// a CTA per record: its name to the names blob; if it is searched, its bases decoded into the batch
// (ping_pong.cpp:88-94: seq_nt16_str, then seq_nt6_table -- A C G T -> 1 2 3 4, everything else 5)
__global__ void __launch_bounds__(128) k_bam_gather(const uint8_t* __restrict__ win, const int64_t* __restrict__ rec_off, const int64_t* __restrict__ seq_off,
                                                    const BamMeta* __restrict__ meta, const int64_t* __restrict__ name_off,
                                                    const int64_t* __restrict__ base_off, const int64_t* __restrict__ rank, char* __restrict__ names,
                                                    uint8_t* __restrict__ batch, int64_t* __restrict__ batch_offs, int64_t batch_reads,
                                                    int64_t batch_bases) {
  const int64_t i = blockIdx.x;
  const BamMeta m = meta[i];
  if (m.state == 0 || m.state == 3) return;
  const uint8_t* name = win + rec_off[i] + 32;
  for (int k = threadIdx.x; k < m.name_len; k += blockDim.x) names[name_off[i] + k] = (char)name[k];
  if (m.state != 2) return;
  const uint8_t* s = win + seq_off[i];
  uint8_t* d = batch + batch_bases + base_off[i];
  for (int k = threadIdx.x; k < m.l_qseq; k += blockDim.x) {
    const uint8_t b = s[k >> 1];
    const unsigned c = (k & 1) ? (b & 0xfu) : (b >> 4);
    d[k] = bam_nt6_of_nt16(c);
  }
  if (threadIdx.x == 0) batch_offs[batch_reads + rank[i] + 1] = batch_bases + base_off[i] + m.l_qseq;
}

}  // namespace svb

using namespace svb;

namespace {

// grow-only device buffer from the stream-ordered pool
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t need(size_t bytes, size_t keep = 0) {
    if (bytes <= cap) return cudaSuccess;
    const size_t want = std::max(bytes, cap + cap / 2);
    void* q = nullptr;
    cudaError_t e = pmalloc(&q, want, 0);
    if (e != cudaSuccess) return e;
    if (keep && p) e = cudaMemcpyAsync(q, p, keep, cudaMemcpyDeviceToDevice, 0);
    pfree(p, 0);
    p = q; cap = want;
    return e;
  }
  void release() { pfree(p, 0); p = nullptr; cap = 0; }
};

}  // namespace

struct svb_bamstream {
  int device = 0;
  int putative = 1;
  int64_t skip_left = 0;       // bytes of the BAM header not yet passed
  bool skip_set = false;
  int64_t carry = 0;           // bytes of an unfinished record at the front of win
  DevBuf comp, io, oo, st, win, rec_off, seq_off, meta, name_off, base_off, rank, names, sums, res, err, seg_pos, seg_aux, aln, sel, f_co, f_so, f_cig, f_seq;
  bool align_mode = false;     // svb_bamstream_open with putative < 0: the Clusterer's scan -- alignments instead of a search batch
  int64_t n_last = 0;          // records of the last window (what svb_bamstream_fetch indexes)
  int64_t tail_from = 0, tail_len = 0;   // alignment mode: the unfinished record still sits at win[tail_from, +tail_len)
  std::vector<BamAln> h_aln;
  std::vector<int32_t> a_pos, a_end, a_ncig;
  std::vector<uint8_t> a_mapq;
  std::vector<uint32_t> f_hcig;
  std::vector<uint8_t> f_hseq;
  std::vector<int64_t> f_hco, f_hso;
  int n_ref = 1 << 30;         // reference sequences of the header (svb_bamstream_set_refs): the walk's plausibility test uses it
  DevBuf batch, batch_offs;
  int64_t batch_reads = 0, batch_bases = 0;
  // host copies handed to the caller, valid until the next call
  std::vector<BamMeta> h_meta;
  std::vector<uint16_t> flag;
  std::vector<int32_t> tid, l_qseq, xf, hp;
  std::vector<uint8_t> state;
  std::vector<int64_t> h_name_off;
  std::vector<char> h_names;
};

extern "C" int svb_bamstream_open(int device, int putative, int n_ref, svb_bamstream_t** out) {
  if (!out) { set_error("svb_bamstream_open: null out"); return SVB_EINVAL; }
  SVB_TRY(check_device(device));
  svb_bamstream* s = new svb_bamstream();
  s->device = device;
  s->putative = putative > 0 ? 1 : 0;
  s->align_mode = putative < 0;
  if (n_ref > 0) s->n_ref = n_ref;
  *out = s;
  return SVB_OK;
}

extern "C" void svb_bamstream_close(svb_bamstream_t* s) {
  if (!s) return;
  cudaSetDevice(s->device);
  for (DevBuf* b : {&s->comp, &s->io, &s->oo, &s->st, &s->win, &s->rec_off, &s->seq_off, &s->meta, &s->name_off, &s->base_off, &s->rank, &s->names,
                    &s->sums, &s->res, &s->err, &s->seg_pos, &s->seg_aux, &s->aln, &s->sel, &s->f_co, &s->f_so, &s->f_cig, &s->f_seq, &s->batch, &s->batch_offs})
    b->release();
  cudaStreamSynchronize(0);
  delete s;
}

extern "C" int64_t svb_bamstream_pending_bytes(const svb_bamstream_t* s) { return s ? s->carry : 0; }

extern "C" int svb_bamstream_window(svb_bamstream_t* s, const uint8_t* comp, const int64_t* in_offs, const int64_t* out_offs, int64_t n_members,
                                    int64_t skip_bytes, svb_bam_recs_t* recs) {
  if (!s || !recs || !in_offs || !out_offs || n_members < 0 || n_members > 0x7fffffff) { set_error("svb_bamstream_window: bad arguments"); return SVB_EINVAL; }
  memset(recs, 0, sizeof(*recs));
  SVB_TRY(check_device(s->device));
  if (!s->skip_set) { s->skip_left = skip_bytes < 0 ? 0 : skip_bytes; s->skip_set = true; }
  const int64_t in_total = n_members ? in_offs[n_members] : 0, out_total = n_members ? out_offs[n_members] : 0;
  if (n_members && (in_offs[0] != 0 || out_offs[0] != 0 || in_total < 0 || out_total < 0 || (in_total > 0 && !comp))) {
    set_error("svb_bamstream_window: offsets must start at 0"); return SVB_EINVAL;
  }
  for (int64_t m = 0; m < n_members; ++m)
    if (in_offs[m + 1] < in_offs[m] || out_offs[m + 1] < out_offs[m] || out_offs[m + 1] - out_offs[m] > 65536) {
      set_error("svb_bamstream_window: member %lld: offsets not ascending or more than 64 KiB of payload", (long long)m); return SVB_EINVAL;
    }
  int rc = SVB_OK;
  StageLog slog("bamstream");
  auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == SVB_OK) { set_error("svb_bamstream_window: %s", cudaGetErrorString(e)); rc = SVB_ECUDA; } };
  if (s->align_mode && s->tail_len > 0 && s->tail_from > 0) {   // the move deferred by the last call (see below)
    if (s->tail_len <= s->tail_from) fail(cudaMemcpyAsync(s->win.p, static_cast<uint8_t*>(s->win.p) + s->tail_from, (size_t)s->tail_len, cudaMemcpyDeviceToDevice, 0));
    else {
      void* tmp = nullptr;
      fail(pmalloc(&tmp, (size_t)s->tail_len, 0));
      if (rc == SVB_OK) {
        fail(cudaMemcpyAsync(tmp, static_cast<uint8_t*>(s->win.p) + s->tail_from, (size_t)s->tail_len, cudaMemcpyDeviceToDevice, 0));
        fail(cudaMemcpyAsync(s->win.p, tmp, (size_t)s->tail_len, cudaMemcpyDeviceToDevice, 0));
      }
      pfree(tmp, 0);
    }
    s->tail_from = 0;
  }
  s->n_last = 0;
  // ---- inflate behind the carried tail
  const int64_t total = s->carry + out_total;
  fail(s->win.need((size_t)total + 64, (size_t)s->carry));
  if (n_members && rc == SVB_OK) {
    fail(s->comp.need((size_t)in_total + 16));
    fail(s->io.need((size_t)(n_members + 1) * 8));
    fail(s->oo.need((size_t)(n_members + 1) * 8));
    fail(s->st.need((size_t)n_members * 4));
    if (rc == SVB_OK) {
      if (in_total) fail(cudaMemcpyAsync(s->comp.p, comp, (size_t)in_total, cudaMemcpyHostToDevice, 0));
      fail(cudaMemcpyAsync(s->io.p, in_offs, (size_t)(n_members + 1) * 8, cudaMemcpyHostToDevice, 0));
      fail(cudaMemcpyAsync(s->oo.p, out_offs, (size_t)(n_members + 1) * 8, cudaMemcpyHostToDevice, 0));
      launch_inflate(static_cast<const uint8_t*>(s->comp.p), static_cast<const int64_t*>(s->io.p), static_cast<const int64_t*>(s->oo.p), n_members,
                     static_cast<uint8_t*>(s->win.p) + s->carry, static_cast<int32_t*>(s->st.p), s->device, 0);
      fail(cudaGetLastError());
      std::vector<int32_t> st((size_t)n_members);
      fail(cudaMemcpy(st.data(), s->st.p, (size_t)n_members * 4, cudaMemcpyDeviceToHost));
      if (rc == SVB_OK)
        for (int64_t m = 0; m < n_members; ++m)
          if (st[(size_t)m] != 0) { set_error("BGZF member %lld does not inflate (code %d): truncated or corrupt file", (long long)m, st[(size_t)m]); return SVB_EIO; }
    }
  }
  if (rc != SVB_OK) return rc;
  slog.lap("H2D + inflate");
  // ---- the BAM header (its length comes from the caller, who parsed it) is passed over
  int64_t start = 0;
  if (s->skip_left > 0) {
    start = std::min<int64_t>(s->skip_left, total);
    s->skip_left -= start;
  }
  // ---- walk the records
  const int64_t cap = std::max<int64_t>(1024, (total - start) / 36 + 1);   // a record is at least 36 bytes
  fail(s->rec_off.need((size_t)cap * 8));
  fail(s->res.need(4 * 8));
  int64_t res[4] = {0, start, 0, 0};
  const char* ew = getenv("SVB_BAM_WALK");
  const bool serial_walk = (ew && strcmp(ew, "serial") == 0) || (total - start < (1 << 20) && !(ew && strcmp(ew, "parallel") == 0));   // tests force either
  if (rc == SVB_OK && serial_walk) {
    k_bam_walk<<<1, 1>>>(static_cast<const uint8_t*>(s->win.p), start, total, static_cast<int64_t*>(s->rec_off.p), cap, static_cast<int64_t*>(s->res.p));
    fail(cudaGetLastError());
    fail(cudaMemcpy(res, s->res.p, 3 * 8, cudaMemcpyDeviceToHost));
  } else if (rc == SVB_OK) {
    const int n_seg = 512;
    const int64_t seg_len = (total - start + n_seg - 1) / n_seg, seg_cap = seg_len / 36 + 2;
    fail(s->seg_pos.need((size_t)n_seg * (size_t)seg_cap * 8));
    fail(s->seg_aux.need((size_t)n_seg * 24));
    if (rc == SVB_OK) {
      int64_t* seg_cnt = static_cast<int64_t*>(s->seg_aux.p);
      int64_t* seg_end = seg_cnt + n_seg;
      int* seg_flag = reinterpret_cast<int*>(seg_end + n_seg);
      k_bam_walk_seg<<<(n_seg + 3) / 4, 128>>>(static_cast<const uint8_t*>(s->win.p), start, total, seg_len, n_seg, s->n_ref, static_cast<int64_t*>(s->seg_pos.p), seg_cap,
                                                seg_cnt, seg_end, seg_flag);
      k_bam_walk_link<<<1, 256>>>(static_cast<const uint8_t*>(s->win.p), start, total, seg_len, n_seg, static_cast<const int64_t*>(s->seg_pos.p), seg_cap, seg_cnt, seg_end,
                                seg_flag, static_cast<int64_t*>(s->rec_off.p), cap, static_cast<int64_t*>(s->res.p));
      fail(cudaGetLastError());
      fail(cudaMemcpy(res, s->res.p, sizeof(res), cudaMemcpyDeviceToHost));
      if (slog.on) fprintf(stderr, "[svb-stage] bamstream: %lld of %d segments joined on their guessed chain\n", (long long)res[3], n_seg);
    }
  }
  if (rc != SVB_OK) return rc;
  if (res[2]) { set_error("BAM record with a block_size below 32: corrupt file"); return SVB_EIO; }
  const int64_t n = res[0], p_end = res[1];
  slog.lap("record walk");
  // ---- parse, lay out, gather
  BamSums sums = {0, 0, 0};
  if (n) {
    fail(s->seq_off.need((size_t)n * 8));
    fail(s->meta.need((size_t)n * sizeof(BamMeta)));
    fail(s->name_off.need((size_t)(n + 1) * 8));
    fail(s->base_off.need((size_t)n * 8));
    fail(s->rank.need((size_t)n * 8));
    fail(s->sums.need(sizeof(BamSums)));
    fail(s->err.need(4));
    if (rc == SVB_OK) {
      fail(cudaMemsetAsync(s->err.p, 0, 4, 0));
      if (s->align_mode) fail(s->aln.need((size_t)n * sizeof(BamAln)));
      k_bam_parse<<<(unsigned)((n + 127) / 128), 128>>>(static_cast<const uint8_t*>(s->win.p), static_cast<const int64_t*>(s->rec_off.p), n, s->putative,
                                                          static_cast<BamMeta*>(s->meta.p), static_cast<int64_t*>(s->seq_off.p),
                                                          s->align_mode ? static_cast<BamAln*>(s->aln.p) : nullptr, static_cast<int*>(s->err.p));
      k_bam_scan<<<1, 1024>>>(static_cast<const BamMeta*>(s->meta.p), n, static_cast<int64_t*>(s->name_off.p), static_cast<int64_t*>(s->base_off.p),
                              static_cast<int64_t*>(s->rank.p), static_cast<BamSums*>(s->sums.p));
      fail(cudaGetLastError());
      int err = 0;
      fail(cudaMemcpy(&err, s->err.p, 4, cudaMemcpyDeviceToHost));
      fail(cudaMemcpy(&sums, s->sums.p, sizeof(sums), cudaMemcpyDeviceToHost));
      if (rc == SVB_OK && err) { set_error("malformed BAM record (field lengths or aux tags run past its block_size)"); return SVB_EIO; }
    }
    if (rc == SVB_OK) {
      fail(s->names.need((size_t)std::max<long long>(sums.names, 1)));
      // the batch keeps what earlier windows put there; 64 readable bytes behind the last base for the search kernels
      fail(s->batch.need((size_t)(s->batch_bases + sums.bases) + 128, (size_t)s->batch_bases));
      fail(s->batch_offs.need((size_t)(s->batch_reads + sums.reads + 1) * 8, (size_t)(s->batch_reads + 1) * 8));
      if (rc == SVB_OK && s->batch_reads == 0) fail(cudaMemsetAsync(s->batch_offs.p, 0, 8, 0));
    }
    if (rc == SVB_OK) {
      k_bam_gather<<<(unsigned)n, 128>>>(static_cast<const uint8_t*>(s->win.p), static_cast<const int64_t*>(s->rec_off.p), static_cast<const int64_t*>(s->seq_off.p),
                                         static_cast<const BamMeta*>(s->meta.p), static_cast<const int64_t*>(s->name_off.p),
                                         static_cast<const int64_t*>(s->base_off.p), static_cast<const int64_t*>(s->rank.p), static_cast<char*>(s->names.p),
                                         static_cast<uint8_t*>(s->batch.p), static_cast<int64_t*>(s->batch_offs.p), s->batch_reads, s->batch_bases);
      fail(cudaGetLastError());
      s->h_meta.resize((size_t)n);
      s->h_name_off.resize((size_t)n + 1);
      s->h_names.resize((size_t)std::max<long long>(sums.names, 1));
      fail(cudaMemcpy(s->h_meta.data(), s->meta.p, (size_t)n * sizeof(BamMeta), cudaMemcpyDeviceToHost));
      fail(cudaMemcpy(s->h_name_off.data(), s->name_off.p, (size_t)(n + 1) * 8, cudaMemcpyDeviceToHost));
      if (sums.names) fail(cudaMemcpy(s->h_names.data(), s->names.p, (size_t)sums.names, cudaMemcpyDeviceToHost));
      if (s->align_mode) { s->h_aln.resize((size_t)n); fail(cudaMemcpy(s->h_aln.data(), s->aln.p, (size_t)n * sizeof(BamAln), cudaMemcpyDeviceToHost)); }
    }
    if (rc != SVB_OK) return rc;
    s->batch_reads += sums.reads;
    s->batch_bases += sums.bases;
  } else {
    s->h_meta.clear(); s->h_name_off.assign(1, 0); s->h_names.assign(1, 0);
  }
  slog.lap("parse + scan + gather + D2H");
  // ---- what is left of the window is the head of a record the next window completes (the move does not touch the
  // records before p_end when left <= p_end; in alignment mode svb_bamstream_fetch reads them after this call, so there
  // the tail is moved at the start of the next call instead)
  const int64_t left = total - p_end;
  s->n_last = n;
  if (s->align_mode) { s->tail_from = p_end; s->tail_len = left; s->carry = left; }
  else if (left > 0 && p_end > 0) {
    if (left <= p_end) fail(cudaMemcpyAsync(s->win.p, static_cast<uint8_t*>(s->win.p) + p_end, (size_t)left, cudaMemcpyDeviceToDevice, 0));
    else {   // overlapping ranges: through a scratch buffer
      void* tmp = nullptr;
      fail(pmalloc(&tmp, (size_t)left, 0));
      if (rc == SVB_OK) {
        fail(cudaMemcpyAsync(tmp, static_cast<uint8_t*>(s->win.p) + p_end, (size_t)left, cudaMemcpyDeviceToDevice, 0));
        fail(cudaMemcpyAsync(s->win.p, tmp, (size_t)left, cudaMemcpyDeviceToDevice, 0));
      }
      pfree(tmp, 0);
    }
  }
  if (!s->align_mode) s->carry = left;
  fail(cudaStreamSynchronize(0));
  if (rc != SVB_OK) return rc;
  // ---- the caller's view
  s->flag.resize((size_t)n); s->tid.resize((size_t)n); s->l_qseq.resize((size_t)n); s->xf.resize((size_t)n); s->hp.resize((size_t)n); s->state.resize((size_t)n);
  for (int64_t i = 0; i < n; ++i) {
    const BamMeta& m = s->h_meta[(size_t)i];
    s->flag[(size_t)i] = m.flag; s->tid[(size_t)i] = m.tid; s->l_qseq[(size_t)i] = m.l_qseq; s->xf[(size_t)i] = m.xf; s->hp[(size_t)i] = m.hp;
    s->state[(size_t)i] = m.state;
  }
  recs->n = n;
  recs->flag = s->flag.data(); recs->tid = s->tid.data(); recs->l_qseq = s->l_qseq.data(); recs->xf = s->xf.data(); recs->hp = s->hp.data();
  recs->state = s->state.data(); recs->name_offs = s->h_name_off.data(); recs->names = s->h_names.data();
  recs->batch_reads = s->batch_reads; recs->batch_bases = s->batch_bases;
  if (s->align_mode) {
    s->a_pos.resize((size_t)n); s->a_end.resize((size_t)n); s->a_ncig.resize((size_t)n); s->a_mapq.resize((size_t)n);
    for (int64_t i = 0; i < n; ++i) {
      const BamAln& al = s->h_aln[(size_t)i];
      s->a_pos[(size_t)i] = al.pos; s->a_end[(size_t)i] = al.endpos; s->a_ncig[(size_t)i] = al.n_cigar; s->a_mapq[(size_t)i] = al.mapq;
    }
    recs->pos = s->a_pos.data(); recs->endpos = s->a_end.data(); recs->n_cigar = s->a_ncig.data(); recs->mapq = s->a_mapq.data();
  }
  recs->h2d_bytes = in_total + (n_members + 1) * 16;
  recs->d2h_bytes = n * (int64_t)sizeof(BamMeta) + (n + 1) * 8 + sums.names + n_members * 4;
  return SVB_OK;
}

extern "C" int svb_bamstream_search(svb_bamstream_t* s, const svb_index_t* idx, int overlap, int assemble, svb_sfs_out_t* out) {
  if (!s || !idx || !out) { set_error("svb_bamstream_search: null argument"); return SVB_EINVAL; }
  memset(out, 0, sizeof(*out));
  SVB_TRY(check_device(s->device));
  if (s->batch_reads == 0) {
    out->offs = (int64_t*)calloc(1, 8);
    return out->offs ? SVB_OK : SVB_ENOMEM;
  }
  SVB_CUDA(cudaMemsetAsync(static_cast<uint8_t*>(s->batch.p) + s->batch_bases, 0, 64, 0));
  SVB_CUDA(cudaStreamSynchronize(0));
  svb_reads_t* R = nullptr;
  SVB_TRY(svb_reads_upload(static_cast<const uint8_t*>(s->batch.p), static_cast<const int64_t*>(s->batch_offs.p), s->batch_reads, SVB_MEM_DEVICE, s->device, &R));
  const int rc = svb_sfs_resident(idx, R, overlap, assemble, out);
  svb_reads_free(R);
  s->batch_reads = 0;
  s->batch_bases = 0;
  return rc;
}

extern "C" int svb_bamstream_fetch(svb_bamstream_t* s, const int64_t* rec, int64_t n_rec, const uint32_t** cigar, const int64_t** cigar_offs,
                                   const uint8_t** seq4, const int64_t** seq4_offs) {
  if (!s || !s->align_mode || n_rec < 0 || (n_rec > 0 && !rec) || !cigar || !cigar_offs || !seq4 || !seq4_offs) {
    set_error("svb_bamstream_fetch: bad arguments (the stream must be opened for alignments)"); return SVB_EINVAL;
  }
  SVB_TRY(check_device(s->device));
  s->f_hco.assign((size_t)n_rec + 1, 0); s->f_hso.assign((size_t)n_rec + 1, 0);
  for (int64_t k = 0; k < n_rec; ++k) {
    if (rec[k] < 0 || rec[k] >= s->n_last) { set_error("svb_bamstream_fetch: record %lld is not of the last window", (long long)rec[k]); return SVB_EINVAL; }
    s->f_hco[(size_t)k + 1] = s->f_hco[(size_t)k] + s->h_aln[(size_t)rec[k]].n_cigar;
    s->f_hso[(size_t)k + 1] = s->f_hso[(size_t)k] + ((int64_t)s->h_meta[(size_t)rec[k]].l_qseq + 1) / 2;
  }
  const int64_t nc = s->f_hco[(size_t)n_rec], ns = s->f_hso[(size_t)n_rec];
  s->f_hcig.resize((size_t)std::max<int64_t>(nc, 1)); s->f_hseq.resize((size_t)std::max<int64_t>(ns, 1));
  if (n_rec) {
    SVB_CUDA(s->sel.need((size_t)n_rec * 8));
    SVB_CUDA(s->f_co.need((size_t)(n_rec + 1) * 8));
    SVB_CUDA(s->f_so.need((size_t)(n_rec + 1) * 8));
    SVB_CUDA(s->f_cig.need((size_t)std::max<int64_t>(nc, 1) * 4));
    SVB_CUDA(s->f_seq.need((size_t)std::max<int64_t>(ns, 1)));
    SVB_CUDA(cudaMemcpyAsync(s->sel.p, rec, (size_t)n_rec * 8, cudaMemcpyHostToDevice, 0));
    SVB_CUDA(cudaMemcpyAsync(s->f_co.p, s->f_hco.data(), (size_t)(n_rec + 1) * 8, cudaMemcpyHostToDevice, 0));
    SVB_CUDA(cudaMemcpyAsync(s->f_so.p, s->f_hso.data(), (size_t)(n_rec + 1) * 8, cudaMemcpyHostToDevice, 0));
    k_bam_fetch<<<(unsigned)n_rec, 128>>>(static_cast<const uint8_t*>(s->win.p), static_cast<const int64_t*>(s->rec_off.p), static_cast<const int64_t*>(s->seq_off.p),
                                          static_cast<const BamMeta*>(s->meta.p), static_cast<const BamAln*>(s->aln.p), static_cast<const int64_t*>(s->sel.p),
                                          static_cast<const int64_t*>(s->f_co.p), static_cast<const int64_t*>(s->f_so.p), static_cast<uint32_t*>(s->f_cig.p),
                                          static_cast<uint8_t*>(s->f_seq.p));
    SVB_CUDA(cudaGetLastError());
    if (nc) SVB_CUDA(cudaMemcpy(s->f_hcig.data(), s->f_cig.p, (size_t)nc * 4, cudaMemcpyDeviceToHost));
    if (ns) SVB_CUDA(cudaMemcpy(s->f_hseq.data(), s->f_seq.p, (size_t)ns, cudaMemcpyDeviceToHost));
    SVB_CUDA(cudaStreamSynchronize(0));
  }
  *cigar = s->f_hcig.data(); *cigar_offs = s->f_hco.data(); *seq4 = s->f_hseq.data(); *seq4_offs = s->f_hso.data();
  return SVB_OK;
}
