// k_poa (svdss_b200/csrc/poa_kernel.cuh, the source the GPU library is built from) compiled for the
// CPU with the lock-step warp emulator: one warp works through all clusters of a call.  Used by
// tests/test_poa_emul.py to check both kernel variants against the oracle without a GPU.
#include <string.h>

#include "warp_emul.hpp"
#include "../../svdss_b200/csrc/poa_kernel.cuh"

namespace svb { int poa_smem[1 << 18]; }

struct Launch { svb::PoaParams P; int variant, group; };

static void body(void* a) {
  Launch* l = static_cast<Launch*>(a);
  switch (l->group * 100 + l->variant) {
#define CASE(V) case 3200 + V: svb::k_poa<V, 32>(l->P); break;
    CASE(0) CASE(1) CASE(2) CASE(4) CASE(7) CASE(32) CASE(39) CASE(68) CASE(71) CASE(135) CASE(199) CASE(263) CASE(455) CASE(487) CASE(967) CASE(1479) CASE(512) CASE(1024)
#undef CASE
#define CASE(V) case 1600 + V: svb::k_poa<V, 16>(l->P); break; case 800 + V: svb::k_poa<V, 8>(l->P); break;
    CASE(0) CASE(7) CASE(39) CASE(455) CASE(487)
#undef CASE
    default: fprintf(stderr, "poa_emul: variant %d / group %d not built\n", l->variant, l->group); abort();
  }
}

extern "C" int emul_poa(const uint8_t* seqs, const int64_t* seq_offs, const int64_t* cluster_offs, int n_clusters, int smem, int group,
                        int ncap, int ecap, int wcap, int lmax, int swcap, uint8_t* cons, const int64_t* cons_off, int32_t* cons_len,
                        int32_t* status, unsigned long long* cells) {
  const int groups = 32 / group;   // clusters in flight: one workspace slot and one shared-memory slice each
  if (swcap <= 0) swcap = wcap < 128 ? wcap : 128;   // as svb_poa_batch sizes it
  if ((smem & 1) && 6 * (long long)swcap * groups > (long long)(sizeof(svb::poa_smem) / sizeof(int))) return -2;
  for (size_t i = 0; i < sizeof(svb::poa_smem) / sizeof(int); ++i) svb::poa_smem[i] = 0x7badbad;   // shared memory starts undefined
  const int64_t stride = svb::poa_ws_carve(nullptr, ncap, ecap, wcap, lmax, nullptr);
  std::vector<int64_t> slot_off((size_t)groups + 1);
  for (int i = 0; i <= groups; ++i) slot_off[(size_t)i] = stride * i;
  std::vector<int4> dims((size_t)(n_clusters > 0 ? n_clusters : 1), make_int4(ncap, ecap, wcap, lmax));
  std::vector<uint8_t> ws((size_t)stride * (size_t)groups + 256, 0xA5);   // not zeroed, like a cudaMalloc'ed workspace
  std::vector<uint32_t> order((size_t)n_clusters);
  for (int i = 0; i < n_clusters; ++i) order[(size_t)i] = (uint32_t)i;
  unsigned work = 0;
  Launch l;
  memset(&l, 0, sizeof(l));
  l.variant = smem; l.group = group;
  svb::PoaParams& P = l.P;
  P.seqs = seqs; P.seq_offs = seq_offs; P.cluster_offs = cluster_offs; P.order = order.data(); P.n = n_clusters;
  P.work = &work; P.ws = ws.data(); P.slot_off = slot_off.data(); P.dims = dims.data(); P.n_slots = groups; P.swcap = swcap;
  P.cons = cons; P.cons_off = cons_off; P.cons_len = cons_len; P.status = status; P.cells = cells; P.phase = nullptr;
  P.match = 2; P.mismatch = 4; P.o1 = 4; P.e1 = 2; P.o2 = 24; P.e2 = 1; P.wb = 10; P.wf = 0.01f;   // as svb_poa_batch sets them
  blockDim.x = 32; blockIdx.x = 0;
  return emu::run_warp(body, &l) ? 0 : -1;
}
