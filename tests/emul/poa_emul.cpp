// k_poa (svdss_b200/csrc/poa_kernel.cuh, the source the GPU library is built from) compiled for the
// CPU with the lock-step warp emulator: one warp works through all clusters of a call.  Used by
// tests/test_poa_emul.py to check both kernel variants against the oracle without a GPU.
#include <string.h>

#include "warp_emul.hpp"
#include "../../svdss_b200/csrc/poa_kernel.cuh"

namespace svb { int poa_smem[1 << 18]; }

struct Launch { svb::PoaParams P; int variant; };

static void body(void* a) {
  Launch* l = static_cast<Launch*>(a);
  switch (l->variant) {
#define CASE(V) case V: svb::k_poa<V>(l->P); break;
    CASE(0) CASE(1) CASE(2) CASE(3) CASE(4) CASE(6) CASE(7) CASE(8) CASE(14) CASE(15) CASE(16) CASE(18) CASE(30) CASE(31)
#undef CASE
    default: break;
  }
}

extern "C" int emul_poa(const uint8_t* seqs, const int64_t* seq_offs, const int64_t* cluster_offs, int n_clusters, int smem,
                        int ncap, int ecap, int wcap, int lmax, uint8_t* cons, const int64_t* cons_off, int32_t* cons_len,
                        int32_t* status, unsigned long long* cells) {
  if ((smem & 1) && 6 * (long long)wcap > (long long)(sizeof(svb::poa_smem) / sizeof(int))) return -2;
  const int64_t stride = svb::poa_ws_carve(nullptr, ncap, ecap, wcap, lmax, nullptr);
  std::vector<uint8_t> ws((size_t)stride + 256, 0xA5);   // not zeroed, like a cudaMalloc'ed workspace
  std::vector<uint32_t> order((size_t)n_clusters);
  for (int i = 0; i < n_clusters; ++i) order[(size_t)i] = (uint32_t)i;
  unsigned work = 0;
  Launch l;
  memset(&l, 0, sizeof(l));
  l.variant = smem;
  svb::PoaParams& P = l.P;
  P.seqs = seqs; P.seq_offs = seq_offs; P.cluster_offs = cluster_offs; P.order = order.data(); P.n = n_clusters;
  P.work = &work; P.ws = ws.data(); P.ws_stride = stride; P.ncap = ncap; P.ecap = ecap; P.wcap = wcap; P.lmax = lmax;
  P.cons = cons; P.cons_off = cons_off; P.cons_len = cons_len; P.status = status; P.cells = cells; P.phase = nullptr;
  P.match = 2; P.mismatch = 4; P.o1 = 4; P.e1 = 2; P.o2 = 24; P.e2 = 1; P.wb = 10; P.wf = 0.01f;   // as svb_poa_batch sets them
  blockDim.x = 32; blockIdx.x = 0;
  return emu::run_warp(body, &l) ? 0 : -1;
}
