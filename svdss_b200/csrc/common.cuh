// Shared declarations for libsvdss_b200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/svdss_b200.h"

namespace svb {

// thread-local last error text, readable through svb_last_error()
void set_error(const char* fmt, ...);

#define SVB_CUDA(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) {                                                                   \
      svb::set_error("%s:%d: %s failed: %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return SVB_ECUDA;                                                                        \
    }                                                                                          \
  } while (0)

#define SVB_TRY(expr)            \
  do {                           \
    int _r = (expr);             \
    if (_r != SVB_OK) return _r; \
  } while (0)

// ---------------------------------------------------------------------------------------------
// FM-index block array ("sampled-Occ BWT blocks") in HBM.
//
// A block is G consecutive 16-byte *slices*; slice j of a block is one uint4
//     .x = Occ count slot j at block start, u32, relative to the block's 2^32-symbol superblock
//     .y/.z/.w = bit-planes 0/1/2 of the nt6 codes of BWT symbols [32j, 32j+32) of the block
// so a G-lane group fetches a whole block with ONE coalesced 16-byte load per lane and every lane
// owns the 32 symbols whose planes it loaded.
//   G = 4 : 64-byte blocks, 128 symbols, slots {A,C,G,T};       N from the side array cntN[]
//   G = 8 : 128-byte blocks, 256 symbols, slots {A,C,G,T,N, cum64, cum128, cum192} where cumX packs,
//           one byte per symbol A,C,G,T, the in-block count before symbol X (sub-block sampling)
// Padding symbols past n carry code 7 (matches nothing).  Superblock table:
//     sbase[sb*8 + c] = acc[c] + Occ(c, sb << 32)           (int64, c = 0..5)
// ---------------------------------------------------------------------------------------------
struct IndexDev {
  int device = 0;
  int G = 8;                 // lanes per block: 4 or 8
  int64_t n = 0;             // BWT length (= text length, both strands + sentinels)
  int64_t acc[7] = {0};
  int64_t n_blocks = 0;
  uint4* d_blocks = nullptr; // n_blocks * G slices
  uint32_t* d_cntN = nullptr;// G == 4 only: Occ(N) at block start, relative to superblock
  int64_t* d_sbase = nullptr;
  int n_sb = 0;
  int64_t n_contigs = 0;
};

}  // namespace svb

struct svb_index {
  svb::IndexDev dev;
};
