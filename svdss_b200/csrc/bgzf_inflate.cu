// svb_bgzf_inflate_device: a window of BGZF members inflated on the GPU (inflate_kernel.cuh, one thread per
// member).  The caller has walked the gzip headers (SAM spec 4.1: BC extra subfield = member size, ISIZE = payload
// size) -- sequential and cheap, what host/io.hpp's BgzfSource::fill does before its parallel zlib loop
// (the reference: htslib bgzf_mt, ping_pong.cpp:249, clusterer.cpp:13) -- and hands over the raw-deflate payloads
// back to back with their offsets and the offsets of their inflated bytes.  Round 1: a measured-later building
// block with its own parity tests (tests/test_inflate_emul.py on the CPU, tests/test_gpu_zz_inflate.py on the GPU);
// BgzfSource still inflates on the host.
#include <algorithm>
#include <cstring>

#include "common.cuh"
#include "inflate_kernel.cuh"

namespace svb {

int check_device(int device);

}  // namespace svb

using namespace svb;

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
  fail(cudaMalloc((void**)&d_in, (size_t)in_total + 16));
  fail(cudaMalloc((void**)&d_out, (size_t)out_total + 16));
  fail(cudaMalloc((void**)&d_io, (size_t)(n_members + 1) * 8));
  fail(cudaMalloc((void**)&d_oo, (size_t)(n_members + 1) * 8));
  fail(cudaMalloc((void**)&d_st, (size_t)n_members * 4));
  fail(cudaEventCreate(&e0));
  fail(cudaEventCreate(&e1));
  if (rc == SVB_OK) {
    if (in_total) fail(cudaMemcpyAsync(d_in, comp, (size_t)in_total, cudaMemcpyHostToDevice, sq));
    fail(cudaMemcpyAsync(d_io, in_offs, (size_t)(n_members + 1) * 8, cudaMemcpyHostToDevice, sq));
    fail(cudaMemcpyAsync(d_oo, out_offs, (size_t)(n_members + 1) * 8, cudaMemcpyHostToDevice, sq));
  }
  if (rc == SVB_OK) {
    // one warp per CTA: members differ in length by an order of magnitude, small CTAs retire independently
    fail(cudaEventRecord(e0, sq));
    k_bgzf_inflate<<<(unsigned)((n_members + 31) / 32), 32, 0, sq>>>(d_in, d_io, d_oo, n_members, d_out, d_st);
    fail(cudaGetLastError());
    fail(cudaEventRecord(e1, sq));
    if (out_total) fail(cudaMemcpyAsync(out_host, d_out, (size_t)out_total, cudaMemcpyDeviceToHost, sq));
    fail(cudaMemcpyAsync(st.data(), d_st, (size_t)n_members * 4, cudaMemcpyDeviceToHost, sq));
    fail(cudaStreamSynchronize(sq));
    if (rc == SVB_OK && kernel_ms) fail(cudaEventElapsedTime(kernel_ms, e0, e1));
  }
  if (sq) cudaStreamSynchronize(sq);
  cudaFree(d_in); cudaFree(d_out); cudaFree(d_io); cudaFree(d_oo); cudaFree(d_st);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (sq) cudaStreamDestroy(sq);
  if (rc != SVB_OK) return rc;
  if (status_host) memcpy(status_host, st.data(), (size_t)n_members * 4);
  for (int64_t m = 0; m < n_members; ++m)
    if (st[(size_t)m] != 0) { set_error("BGZF member %lld does not inflate (code %d): truncated or corrupt file", (long long)m, st[(size_t)m]); return SVB_EIO; }
  return SVB_OK;
}
