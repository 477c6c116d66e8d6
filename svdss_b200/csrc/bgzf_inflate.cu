// svb_bgzf_inflate_device: a window of BGZF members inflated on the GPU (inflate_kernel.cuh, one thread per
// member).  The caller has walked the gzip headers (SAM spec 4.1: BC extra subfield = member size, ISIZE = payload
// size) -- sequential and cheap, what host/io.hpp's BgzfSource::fill does before its parallel zlib loop
// (the reference: htslib bgzf_mt, ping_pong.cpp:249, clusterer.cpp:13) -- and hands over the raw-deflate payloads
// back to back with their offsets and the offsets of their inflated bytes.  Round 1: a measured-later building
// block with its own parity tests (tests/test_inflate_emul.py on the CPU, tests/test_gpu_zz_inflate.py on the GPU);
// Round 2: timed -- the kernel takes ~200 ms whatever the window holds up to ~75 k members (one 64 KiB member per
// thread is a latency, not a throughput), i.e. 1.4 GB/s on a 256 MB window and 10.4 GB/s on a 2 GB one
// (profiles/r02p_inflate.txt): 32 lanes on 32 different streams diverge (one member per warp: 16 ms, profiles/r02q_inflate.txt).
// Hence k_bgzf_inflate_warp, the default now, and BgzfSource's opt-in `--gpu-inflate` (128 MiB compressed windows).
#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "common.cuh"
#include "inflate_kernel.cuh"

namespace svb {

int check_device(int device);

}  // namespace svb

using namespace svb;

namespace svb {

// the inflate kernel over members already in HBM (also used by bam_stream.cu): the warp-per-member kernel by default,
// SVB_INFLATE_KERNEL=thread keeps the first one measurable
void launch_inflate(const uint8_t* d_in, const int64_t* d_io, const int64_t* d_oo, int64_t n_members, uint8_t* d_out, int32_t* d_st, int device,
                    cudaStream_t sq) {
  if (n_members <= 0) return;
  const char* kind = getenv("SVB_INFLATE_KERNEL");
  if (kind && strcmp(kind, "thread") == 0) {
    // one warp per CTA: members differ in length by an order of magnitude, small CTAs retire independently.
    // Members per warp: as few as still fill the warp slots of the device (lanes on different streams diverge)
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    int64_t slots = (int64_t)sms * 32;
    if (const char* e = getenv("SVB_INFLATE_WARPS_PER_SM")) { const int v = atoi(e); if (v > 0) slots = (int64_t)sms * v; }
    int mpw = (int)std::min<int64_t>(32, std::max<int64_t>(1, (n_members + slots - 1) / slots));
    if (const char* e = getenv("SVB_INFLATE_MPW")) { const int v = atoi(e); if (v >= 1 && v <= 32) mpw = v; }
    k_bgzf_inflate<<<(unsigned)((n_members + mpw - 1) / mpw), 32, 0, sq>>>(d_in, d_io, d_oo, n_members, d_out, d_st, mpw);
  } else if (getenv("SVB_INFLATE_OCC12")) {
    k_bgzf_inflate_warp<12><<<(unsigned)((n_members + 1) / 2), 64, 2 * sizeof(InfWarpMem), sq>>>(d_in, d_io, d_oo, n_members, d_out, d_st);
  } else {
    k_bgzf_inflate_warp<16><<<(unsigned)((n_members + 1) / 2), 64, 2 * sizeof(InfWarpMem), sq>>>(d_in, d_io, d_oo, n_members, d_out, d_st);
  }
}

}  // namespace svb

extern "C" int svb_bgzf_inflate_device(const uint8_t* comp, const int64_t* in_offs, const int64_t* out_offs, int64_t n_members,
                                       int device, uint8_t* out_host, int32_t* status_host, float* kernel_ms) {
  if (!in_offs || !out_offs || n_members < 0 || n_members > 0x7fffffff) { set_error("svb_bgzf_inflate_device: bad arguments"); return SVB_EINVAL; }
  if (kernel_ms) *kernel_ms = 0.f;
  SVB_TRY(check_device(device));
  if (n_members == 0) return SVB_OK;
  const int64_t in_total = in_offs[n_members], out_total = out_offs[n_members];
  if (in_offs[0] != 0 || out_offs[0] != 0 || in_total < 0 || out_total < 0 || (in_total > 0 && !comp) || (out_total > 0 && !out_host)) {
    set_error("svb_bgzf_inflate_device: offsets must start at 0 and buffers must be given"); return SVB_EINVAL;
  }
  for (int64_t m = 0; m < n_members; ++m)
    if (in_offs[m + 1] < in_offs[m] || out_offs[m + 1] < out_offs[m] || out_offs[m + 1] - out_offs[m] > 65536) {
      set_error("svb_bgzf_inflate_device: member %lld: offsets not ascending or more than 64 KiB of payload", (long long)m); return SVB_EINVAL;
    }
  uint8_t *d_in = nullptr, *d_out = nullptr;
  int64_t *d_io = nullptr, *d_oo = nullptr;
  int32_t* d_st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  // a stream of its own that does not synchronise with the legacy default stream: a reader thread may call this
  // while another thread of the process has a streamed search in flight on the same device
  cudaStream_t sq = nullptr;
  std::vector<int32_t> st((size_t)n_members, 0);
  int rc = SVB_OK;
  auto fail = [&](cudaError_t e) { if (e != cudaSuccess && rc == SVB_OK) { set_error("svb_bgzf_inflate_device: %s", cudaGetErrorString(e)); rc = SVB_ECUDA; } };
  fail(cudaStreamCreateWithFlags(&sq, cudaStreamNonBlocking));
  // window after window of a file: the buffers come from the stream-ordered pool (no cudaMalloc / cudaFree of GBs per window)
  if (rc == SVB_OK) fail(pmalloc((void**)&d_in, (size_t)in_total + 16, sq));
  if (rc == SVB_OK) fail(pmalloc((void**)&d_out, (size_t)out_total + 16, sq));
  if (rc == SVB_OK) fail(pmalloc((void**)&d_io, (size_t)(n_members + 1) * 8, sq));
  if (rc == SVB_OK) fail(pmalloc((void**)&d_oo, (size_t)(n_members + 1) * 8, sq));
  if (rc == SVB_OK) fail(pmalloc((void**)&d_st, (size_t)n_members * 4, sq));
  fail(cudaEventCreate(&e0));
  fail(cudaEventCreate(&e1));
  if (rc == SVB_OK) {
    if (in_total) fail(cudaMemcpyAsync(d_in, comp, (size_t)in_total, cudaMemcpyHostToDevice, sq));
    fail(cudaMemcpyAsync(d_io, in_offs, (size_t)(n_members + 1) * 8, cudaMemcpyHostToDevice, sq));
    fail(cudaMemcpyAsync(d_oo, out_offs, (size_t)(n_members + 1) * 8, cudaMemcpyHostToDevice, sq));
  }
  if (rc == SVB_OK) {
    fail(cudaEventRecord(e0, sq));
    launch_inflate(d_in, d_io, d_oo, n_members, d_out, d_st, device, sq);
    fail(cudaGetLastError());
    fail(cudaEventRecord(e1, sq));
    if (out_total) fail(cudaMemcpyAsync(out_host, d_out, (size_t)out_total, cudaMemcpyDeviceToHost, sq));
    fail(cudaMemcpyAsync(st.data(), d_st, (size_t)n_members * 4, cudaMemcpyDeviceToHost, sq));
    fail(cudaStreamSynchronize(sq));
    if (rc == SVB_OK && kernel_ms) fail(cudaEventElapsedTime(kernel_ms, e0, e1));
  }
  if (sq) { pfree(d_in, sq); pfree(d_out, sq); pfree(d_io, sq); pfree(d_oo, sq); pfree(d_st, sq); cudaStreamSynchronize(sq); }
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (sq) cudaStreamDestroy(sq);
  if (rc != SVB_OK) return rc;
  if (status_host) memcpy(status_host, st.data(), (size_t)n_members * 4);
  for (int64_t m = 0; m < n_members; ++m)
    if (st[(size_t)m] != 0) { set_error("BGZF member %lld does not inflate (code %d): truncated or corrupt file", (long long)m, st[(size_t)m]); return SVB_EIO; }
  return SVB_OK;
}

// Pinned host memory for the windows a reader hands to / gets back from svb_bgzf_inflate_device (a pageable
// destination would halve the copy rate).  NULL without a device or when the allocation fails.
extern "C" void* svb_host_alloc_pinned(size_t bytes) {
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return p;
}
extern "C" void svb_host_free_pinned(void* p) {
  if (p) cudaFreeHost(p);
}
