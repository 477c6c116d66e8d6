// inflate_kernel.cuh (BGZF members inflated on the device, one thread per member) compiled for the CPU with the
// warp emulator: warps of 32 members at a time, as the kernel's grid would run them.
#include <string.h>

#include "warp_emul.hpp"
namespace svb { unsigned long long inf_smem[1024]; }   // the kernel's dynamic shared memory (one warp)
#include "../../svdss_b200/csrc/inflate_kernel.cuh"

struct Job { const uint8_t* comp; const int64_t* in_offs; const int64_t* out_offs; int64_t n; uint8_t* out; int32_t* status; int mpw; };

static void body(void* a) {
  Job* j = static_cast<Job*>(a);
  svb::k_bgzf_inflate(j->comp, j->in_offs, j->out_offs, j->n, j->out, j->status, j->mpw);
}

// mpw: members per warp (1..32), the launch parameter svb_bgzf_inflate_device picks from the window size
extern "C" int emul_bgzf_inflate_mpw(const uint8_t* comp, const int64_t* in_offs, const int64_t* out_offs, int64_t n, uint8_t* out, int32_t* status, int mpw) {
  Job j = {comp, in_offs, out_offs, n, out, status, mpw};
  blockDim.x = 32;
  gridDim.x = (unsigned)((n + mpw - 1) / mpw);
  for (unsigned b = 0; b < gridDim.x; ++b) {
    blockIdx.x = b;
    if (!emu::run_warp(body, &j)) return -1;
  }
  return 0;
}

extern "C" int emul_bgzf_inflate(const uint8_t* comp, const int64_t* in_offs, const int64_t* out_offs, int64_t n, uint8_t* out, int32_t* status) {
  return emul_bgzf_inflate_mpw(comp, in_offs, out_offs, n, out, status, 32);
}

// the warp-per-member kernel: one warp per CTA here
static void body_warp(void* a) {
  Job* j = static_cast<Job*>(a);
  svb::k_bgzf_inflate_warp<16>(j->comp, j->in_offs, j->out_offs, j->n, j->out, j->status);
}
extern "C" int emul_bgzf_inflate_warp(const uint8_t* comp, const int64_t* in_offs, const int64_t* out_offs, int64_t n, uint8_t* out, int32_t* status) {
  static_assert(sizeof(svb::InfWarpMem) <= sizeof(svb::inf_smem), "shared memory of one warp");
  Job j = {comp, in_offs, out_offs, n, out, status, 0};
  blockDim.x = 32;
  gridDim.x = (unsigned)n;
  for (unsigned b = 0; b < gridDim.x; ++b) {
    blockIdx.x = b;
    if (!emu::run_warp(body_warp, &j)) return -1;
  }
  return 0;
}
