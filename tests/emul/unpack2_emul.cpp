// unpack2.cuh (device side of the 2-bit read transport) compiled for the CPU with the warp emulator: one warp
// decodes all reads, optionally range by range like the streamed pipeline would.
#include <string.h>

struct uint4 { unsigned x, y, z, w; };
static inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { uint4 v = {x, y, z, w}; return v; }

#include "warp_emul.hpp"
#include "../../svdss_b200/csrc/unpack2.cuh"

struct Job { const uint8_t* pk; const int64_t* pk_offs; const int64_t* offs; int64_t n, total, chunk; uint8_t* out; };

static void body(void* a) {
  Job* j = static_cast<Job*>(a);
  const int lane = threadIdx.x & 31;
  if (j->chunk <= 0) { svb::k_unpack2(j->pk, j->pk_offs, j->offs, j->n, j->total, j->out); return; }
  for (int64_t A = 0; A < j->total; A += j->chunk)                 // chunk by chunk: every read against every range
    for (int64_t r = 0; r < j->n; ++r) svb::unpack2_read(j->pk, j->pk_offs, j->offs, r, A, A + j->chunk < j->total ? A + j->chunk : j->total, j->out, lane);
}

extern "C" int emul_unpack2(const uint8_t* pk, const int64_t* pk_offs, const int64_t* offs, int64_t n, int64_t total, int64_t chunk, uint8_t* out) {
  Job j = {pk, pk_offs, offs, n, total, chunk, out};
  blockDim.x = 32; blockIdx.x = 0; gridDim.x = 1;
  return emu::run_warp(body, &j) ? 0 : -1;
}

// one chunk of the streamed pipeline: reads r_lo..r_hi against output positions [A, B)
struct RangeJob { const uint8_t* pk; const int64_t* pk_offs; const int64_t* offs; int64_t r_lo, r_hi, A, B; uint8_t* out; };
static void range_body(void* a) {
  RangeJob* j = static_cast<RangeJob*>(a);
  const int lane = threadIdx.x & 31;
  for (int64_t r = j->r_lo; r <= j->r_hi; ++r) svb::unpack2_read(j->pk, j->pk_offs, j->offs, r, j->A, j->B, j->out, lane);
}
extern "C" int emul_unpack2_range(const uint8_t* pk, const int64_t* pk_offs, const int64_t* offs, int64_t r_lo, int64_t r_hi, int64_t A, int64_t B,
                                  uint8_t* out) {
  RangeJob j = {pk, pk_offs, offs, r_lo, r_hi, A, B, out};
  blockDim.x = 32; blockIdx.x = 0; gridDim.x = 1;
  return emu::run_warp(range_body, &j) ? 0 : -1;
}
