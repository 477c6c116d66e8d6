// k_ksw_extd2 (svdss_b200/csrc/ksw_kernel.cuh, the source the GPU library is built from) compiled for the
// CPU with the lock-step warp emulator: one warp works through all pairs of a call.  Used by
// tests/test_ksw_emul.py to check both kernel variants against the oracle without a GPU.
#include <string.h>

#include "warp_emul.hpp"
#include "../../svdss_b200/csrc/ksw_kernel.cuh"

struct Launch { svb::KswParams P; int variant; };

static void body(void* a) {
  Launch* l = static_cast<Launch*>(a);
  switch (l->variant) {
    case 1: svb::k_ksw_extd2<1>(l->P); break;
    case 2: svb::k_ksw_extd2<2>(l->P); break;
    case 3: svb::k_ksw_extd2<3>(l->P); break;
    default: svb::k_ksw_extd2<0>(l->P); break;
  }
}

// cigar_out: forward-order ops of pair p at cigar_off[p] (capacity ql + tl + 2 each), n_cigar[p] ops
extern "C" int emul_ksw(const uint8_t* q, const int64_t* qoff, const uint8_t* t, const int64_t* toff, int n_pairs, int variant,
                        int a, int b, int sc_n, int q1, int e1, int q2, int e2, int32_t* score, uint32_t* cigar_out,
                        const int64_t* cigar_off, int32_t* n_cigar) {
  std::vector<uint32_t> order((size_t)n_pairs);
  std::vector<int64_t> tb_off((size_t)n_pairs + 1, 0), bnd_off((size_t)n_pairs + 1, 0), cg_off((size_t)n_pairs + 1, 0);
  for (int p = 0; p < n_pairs; ++p) {
    order[(size_t)p] = (uint32_t)p;
    const int64_t ql = qoff[p + 1] - qoff[p], tl = toff[p + 1] - toff[p];
    tb_off[(size_t)p + 1] = tb_off[(size_t)p] + svb::ksw_tb_bytes(variant, ql, tl);     // as the host wave planner sizes them
    bnd_off[(size_t)p + 1] = bnd_off[(size_t)p] + svb::ksw_bnd_ints(variant, ql, tl);
    cg_off[(size_t)p + 1] = cg_off[(size_t)p] + ql + tl + 2;
  }
  std::vector<uint8_t> tb((size_t)tb_off[(size_t)n_pairs] + 16, 0xEE);     // not zeroed, like a cudaMalloc'ed buffer
  std::vector<int32_t> bnd((size_t)bnd_off[(size_t)n_pairs] + 4, 0x5A5A5A5A);
  std::vector<uint32_t> cg((size_t)cg_off[(size_t)n_pairs] + 4, 0);
  std::vector<int32_t> cg_n((size_t)n_pairs + 1, 0);
  unsigned work = 0;
  Launch l;
  memset(&l, 0, sizeof(l));
  l.variant = variant;
  svb::KswParams& P = l.P;
  P.q = q; P.qoff = qoff; P.t = t; P.toff = toff; P.order = order.data(); P.n = n_pairs;
  P.tb_off = tb_off.data(); P.bnd_off = bnd_off.data(); P.cg_off = cg_off.data();
  P.tb = tb.data(); P.bnd = bnd.data(); P.cg = cg.data(); P.cg_n = cg_n.data(); P.score = score; P.work = &work;
  P.a = a; P.b = b; P.sc_n = sc_n; P.q1 = q1; P.e1 = e1; P.q2 = q2; P.e2 = e2;
  blockDim.x = 32; blockIdx.x = 0;
  if (!emu::run_warp(body, &l)) return -1;
  for (int p = 0; p < n_pairs; ++p) {   // k_ksw_gather: the kernel leaves the ops in reverse order
    const int n = cg_n[(size_t)p];
    n_cigar[p] = n;
    for (int k = 0; k < n; ++k) cigar_out[cigar_off[p] + k] = cg[(size_t)cg_off[(size_t)p] + (size_t)(n - 1 - k)];
  }
  return 0;
}
