// Device code of the two-piece affine alignment (see ksw_extd2.cu for the design notes): parameters and
// kernels, free of host / runtime-API code so that tests/emul can compile the same source for the CPU
// with the lock-step warp emulator (tests/emul/warp_emul.hpp).
#pragma once
#include <stdint.h>
#include <stddef.h>

namespace svb {

constexpr int KSW_NEG_INF = -0x40000000;
constexpr int KR = 4;           // rows per lane
constexpr int KBAND = 32 * KR;  // target rows per band

struct KswParams {
  const uint8_t* __restrict__ q;
  const int64_t* __restrict__ qoff;
  const uint8_t* __restrict__ t;
  const int64_t* __restrict__ toff;
  const uint32_t* __restrict__ order;  // pair indices of this wave, biggest first
  int n;                               // pairs in this wave
  const int64_t* __restrict__ tb_off;  // per wave slot: offset into tb
  const int64_t* __restrict__ bnd_off; // per wave slot: offset into bnd (int32 triples per column)
  const int64_t* __restrict__ cg_off;  // per wave slot: offset into cigar scratch (capacity ql+tl+2)
  uint8_t* tb;
  int32_t* bnd;
  uint32_t* cg;      // reverse-order ops per pair
  int32_t* cg_n;     // per wave slot: number of ops
  int32_t* score;    // per pair (global index)
  unsigned int* work;
  int a, b, sc_n, q1, e1, q2, e2;
};

__device__ __forceinline__ int gapcost(int k, int q1, int e1, int q2, int e2) { return min(q1 + k * e1, q2 + k * e2); }

__device__ __forceinline__ void ksw_prefetch(const void* p) {
#ifdef __CUDA_ARCH__
  asm volatile("{ .reg .u64 a; cvta.to.global.u64 a, %0; prefetch.global.L1 [a]; }" ::"l"(p));
#else
  (void)p;
#endif
}

// Variants (template bit mask KV, SVB_KSW_VARIANT on the host; 0 = the kernel measured in round 1):
//  1 TBPF  the backtrack reads one traceback byte per step, each in a different 128-byte line (a diagonal
//          step moves one wavefront step back), written long before: a cache-missing dependent walk by one
//          lane.  With TBPF it runs in windows of 32 steps, and before each window all lanes prefetch the
//          bytes the path would read 32..63 steps ahead if it stayed on its diagonal (0..31 as well for the
//          first window).  Addresses are pure arithmetic on (i, j); a wrong guess is a wasted prefetch.
//  2 CKPT  checkpointed traceback for pairs of at least KSW_CKPT_BANDS bands.  One traceback byte per cell
//          makes a 10 kb x 10 kb pair cost 100 MB, so the memory budget -- not the GPU -- decides how many
//          such pairs are in flight (a few hundred for 2 960 warp slots).  With CKPT the forward pass stores
//          no traceback, only the boundary row (H, E, E2) under every band: 12 bytes x ql per band.  The
//          backtrack then walks the bands from the last to the first, recomputing one band at a time -- only
//          the columns left of the path's current position -- into a single band-sized traceback buffer.
//          About 1.5x the cells for a tenth of the memory, hence ten times the pairs in flight.
// Scores and CIGARs are identical in every variant.
constexpr int KSW_V_TBPF = 1, KSW_V_CKPT = 2;
constexpr int KSW_CKPT_BANDS = 4;

// bytes of traceback / boundary ints a pair needs (host wave planner and kernel agree through these)
__host__ __device__ inline bool ksw_uses_ckpt(int variant, int64_t tl) { return (variant & KSW_V_CKPT) && (tl + KBAND - 1) / KBAND >= KSW_CKPT_BANDS; }
__host__ __device__ inline int64_t ksw_tb_bytes(int variant, int64_t ql, int64_t tl) {
  if (ql <= 0 || tl <= 0) return 0;
  const int64_t nbands = (tl + KBAND - 1) / KBAND, nsteps = ql + 31;
  return (ksw_uses_ckpt(variant, tl) ? 1 : nbands) * nsteps * KBAND;
}
__host__ __device__ inline int64_t ksw_bnd_ints(int variant, int64_t ql, int64_t tl) {
  if (ql <= 0 || tl <= 0) return 0;
  const int64_t nbands = (tl + KBAND - 1) / KBAND;
  return (ksw_uses_ckpt(variant, tl) ? nbands : 1) * 3 * ql;
}

template <int KV>
__global__ void __launch_bounds__(128, 5) k_ksw_extd2(const KswParams P) {   // 5 CTAs per SM (<= 102 registers), as measured in round 1
  constexpr bool TBPF = (KV & KSW_V_TBPF) != 0;
  const int lane = threadIdx.x & 31;
  const int NEG = -0x1fffffff;
  int q1 = P.q1, e1 = P.e1, q2 = P.q2, e2 = P.e2;
  if (q2 + e2 < q1 + e1) { int x = q1; q1 = q2; q2 = x; x = e1; e1 = e2; e2 = x; }  // ksw2 swaps the pieces
  for (;;) {
    unsigned w = 0;
    if (lane == 0) w = atomicAdd(P.work, 1u);
    w = __shfl_sync(0xffffffffu, w, 0);
    if (w >= (unsigned)P.n) break;
    const uint32_t pid = P.order[w];
    const uint8_t* Q = P.q + P.qoff[pid];
    const uint8_t* T = P.t + P.toff[pid];
    const int ql = (int)(P.qoff[pid + 1] - P.qoff[pid]);
    const int tl = (int)(P.toff[pid + 1] - P.toff[pid]);
    if (ql <= 0 || tl <= 0) {  // ksw_reset_extz + early return
      if (lane == 0) { P.score[pid] = KSW_NEG_INF; P.cg_n[w] = 0; }
      continue;
    }
    uint8_t* tb = P.tb + P.tb_off[w];
    int32_t* bnd = P.bnd + P.bnd_off[w];  // [3][ql]: H, E, E2 of the row above the current band (CKPT: one such slab per band)
    const int nbands = (tl + KBAND - 1) / KBAND;
    const int nsteps = ql + 31;
    const bool ckpt = ksw_uses_ckpt(KV, tl);
    int final_score = 0;
    // one band of 128 target rows over the first `nst` wavefront steps: boundary row above from `bin` (band > 0),
    // boundary row below to `bout` (or nullptr), traceback bytes to `tbb` (or nullptr)
    auto run_band = [&](int band, const int32_t* bin, int32_t* bout, uint8_t* tbb, int nst) {
      const int i0 = band * KBAND + lane * KR;  // first row of this lane
      uint8_t tc[KR];
      int hl[KR], f[KR], f2[KR];
#pragma unroll
      for (int r = 0; r < KR; ++r) {
        const int i = i0 + r;
        tc[r] = i < tl ? T[i] : 4;
        hl[r] = -gapcost(i + 1, q1, e1, q2, e2);  // H(i,-1)
        f[r] = hl[r] - q1 - e1;                   // F(i,0)
        f2[r] = hl[r] - q2 - e2;
      }
      int hup_prev = i0 ? -gapcost(i0, q1, e1, q2, e2) : 0;  // H(i0-1,-1)
      // values handed down from the row above at this lane's current column
      int in_h = 0, in_e = NEG, in_e2 = NEG;
      int out_h = 0, out_e = NEG, out_e2 = NEG;  // this lane's last row at its previous column
      int qc = 4, qnext = 4;
      int bh = 0, be = NEG, be2 = NEG;  // lane 0's boundary inputs, prefetched 32 columns at a time
      int wh = 0, we = NEG, we2 = NEG;  // lane 31's boundary outputs, flushed 32 columns at a time
      for (int t = 0; t < nst; ++t) {
        // ---- inputs for this step
        if ((t & 31) == 0) {
          const int jj = t + lane;  // cooperative prefetch of 32 columns of query + boundary
          qnext = jj < ql ? Q[jj] : 4;
          if (band == 0) {
            bh = jj < ql ? -gapcost(jj + 1, q1, e1, q2, e2) : 0;  // H(-1,j)
            be = NEG; be2 = NEG;
          } else if (jj < ql) {
            bh = bin[jj]; be = bin[ql + jj]; be2 = bin[2 * ql + jj];
          }
        }
        // query char: lane 0 takes column t, others inherit from the lane above (one step later)
        const int q_in = __shfl_sync(0xffffffffu, qnext, t & 31);
        const int q_up = __shfl_up_sync(0xffffffffu, qc, 1);
        qc = lane == 0 ? q_in : q_up;
        const int b_h = __shfl_sync(0xffffffffu, bh, t & 31);
        const int b_e = __shfl_sync(0xffffffffu, be, t & 31);
        const int b_e2 = __shfl_sync(0xffffffffu, be2, t & 31);
        const int u_h = __shfl_up_sync(0xffffffffu, out_h, 1);
        const int u_e = __shfl_up_sync(0xffffffffu, out_e, 1);
        const int u_e2 = __shfl_up_sync(0xffffffffu, out_e2, 1);
        in_h = lane == 0 ? b_h : u_h;
        in_e = lane == 0 ? b_e : u_e;
        in_e2 = lane == 0 ? b_e2 : u_e2;
        const int j = t - lane;
        const bool act = j >= 0 && j < ql;
        unsigned tbw = 0;
        if (act) {
          int hup = in_h, eup = in_e, e2up = in_e2;
          int hdiag = hup_prev;
          hup_prev = in_h;
#pragma unroll
          for (int r = 0; r < KR; ++r) {
            const int ee = max(hup - q1, eup) - e1;
            const int ee2 = max(hup - q2, e2up) - e2;
            const int sc = (tc[r] == 4 || qc == 4) ? P.sc_n : (tc[r] == qc ? P.a : P.b);
            int h = hdiag + sc;
            unsigned d = 0;
            if (ee > h) { h = ee; d = 1; }
            if (f[r] > h) { h = f[r]; d = 2; }
            if (ee2 > h) { h = ee2; d = 3; }
            if (f2[r] > h) { h = f2[r]; d = 4; }
            const int ho1 = h - q1, ho2 = h - q2;
            d |= (ee > ho1) ? 0x08u : 0u;
            d |= (f[r] > ho1) ? 0x10u : 0u;
            d |= (ee2 > ho2) ? 0x20u : 0u;
            d |= (f2[r] > ho2) ? 0x40u : 0u;
            tbw |= d << (8 * r);
            f[r] = max(ho1, f[r]) - e1;
            f2[r] = max(ho2, f2[r]) - e2;
            hdiag = hl[r];
            hl[r] = h;
            hup = h; eup = ee; e2up = ee2;
            if (j == ql - 1 && i0 + r == tl - 1) final_score = h;
          }
          out_h = hup; out_e = eup; out_e2 = e2up;
          if (tbb) *reinterpret_cast<unsigned*>(tbb + (size_t)t * KBAND + lane * KR) = tbw;
        }
        // ---- lane 31 hands its last row to the next band: collect 32 columns, flush coalesced
        if (bout) {
          const int j31 = t - 31;  // column lane 31 just finished
          const int s_h = __shfl_sync(0xffffffffu, out_h, 31);
          const int s_e = __shfl_sync(0xffffffffu, out_e, 31);
          const int s_e2 = __shfl_sync(0xffffffffu, out_e2, 31);
          if (j31 >= 0 && j31 < ql) {
            if ((j31 & 31) == lane) { wh = s_h; we = s_e; we2 = s_e2; }
            if ((j31 & 31) == 31 || j31 == ql - 1) {
              const int jj = (j31 & ~31) + lane;
              if (jj <= j31) { bout[jj] = wh; bout[ql + jj] = we; bout[2 * ql + jj] = we2; }
            }
          }
        }
      }
      __syncwarp();
    };
    if (!ckpt) {
      // the boundary buffer is reused in place: a band reads columns ahead of the ones it writes
      for (int band = 0; band < nbands; ++band)
        run_band(band, bnd, band + 1 < nbands ? bnd : nullptr, tb + (size_t)band * nsteps * KBAND, nsteps);
    } else {
      // forward pass without traceback: band b reads slab b - 1, writes slab b
      for (int band = 0; band < nbands; ++band)
        run_band(band, band ? bnd + (size_t)(band - 1) * 3 * ql : nullptr, band + 1 < nbands ? bnd + (size_t)band * 3 * ql : nullptr, nullptr, nsteps);
    }
    // score lives in the lane that owned row tl-1
    {
      const int owner = ((tl - 1) % KBAND) / KR;
      final_score = __shfl_sync(0xffffffffu, final_score, owner);
    }
    __syncwarp();
    // ---- ksw_backtrack (ksw2.h) by lane 0
    {
      uint32_t* cg = P.cg + P.cg_off[w];
      int n = 0, i = tl - 1, j = ql - 1, state = 0;   // lane 0's copies count
      uint32_t last = 0;  // open run: len << 4 | op, 0 = none
      auto push = [&](unsigned op, unsigned len) {
        if (last && (last & 0xfu) == op) last += len << 4;
        else { if (last) cg[n++] = last; last = (len << 4) | op; }
      };
      // traceback byte of cell (ii, jj); `band_base` = first row of the band the buffer `tbp` holds
      auto tb_addr = [&](const uint8_t* tbp, int band_base, int ii, int jj) -> const uint8_t* {
        const int l = ((ii - band_base) % KBAND) / KR, r = ii % KR;
        return tbp + (size_t)((ii - band_base) / KBAND) * nsteps * KBAND + (size_t)(jj + l) * KBAND + l * KR + r;
      };
      // walk while the path stays at or below row `i_min` (0 for the whole matrix)
      auto walk = [&](const uint8_t* tbp, int band_base, int i_min) {
        auto step = [&]() {
          const unsigned tmp = *tb_addr(tbp, band_base, i, j);
          if (state == 0) state = tmp & 7;
          else if (!((tmp >> (state + 2)) & 1)) state = 0;
          if (state == 0) state = tmp & 7;
          if (state == 0) { push(0, 1); --i; --j; }
          else if (state == 1 || state == 3) { push(2, 1); --i; }
          else { push(1, 1); --j; }
        };
        if (!TBPF) {
          if (lane == 0) while (i >= i_min && j >= 0) step();
        } else {
          bool first = true;
          for (;;) {
            const int ci = __shfl_sync(0xffffffffu, i, 0), cj = __shfl_sync(0xffffffffu, j, 0);
            if (ci < i_min || cj < 0) break;
            for (int d = first ? lane : 32 + lane; d < 64; d += 32)
              if (ci - d >= i_min && cj - d >= 0) ksw_prefetch(tb_addr(tbp, band_base, ci - d, cj - d));
            first = false;
            __syncwarp();
            if (lane == 0) for (int st = 0; st < 32 && i >= i_min && j >= 0; ++st) step();
            __syncwarp();
          }
        }
      };
      if (!ckpt) walk(tb, 0, 0);
      else {
        for (int band = nbands - 1; band >= 0; --band) {
          const int ci = __shfl_sync(0xffffffffu, i, 0), cj = __shfl_sync(0xffffffffu, j, 0);
          if (ci < 0 || cj < 0) break;
          if (ci < band * KBAND) continue;            // the path has already left this band (a long deletion)
          // the path never moves right: columns beyond cj are not needed (lane l reaches column cj at step cj + l)
          run_band(band, band ? bnd + (size_t)(band - 1) * 3 * ql : nullptr, nullptr, tb, min(nsteps, cj + 32));
          walk(tb, band * KBAND, band * KBAND);
        }
      }
      if (lane == 0) {
        if (i >= 0) push(2, (unsigned)(i + 1));
        if (j >= 0) push(1, (unsigned)(j + 1));
        if (last) cg[n++] = last;
        P.cg_n[w] = n;
        P.score[pid] = final_score;
      }
    }
    __syncwarp();
  }
}

// reverse the per-pair op lists of one wave into a dense wave-local table
__global__ void k_ksw_gather(const uint32_t* __restrict__ cg, const int64_t* __restrict__ cg_off, int n_wave,
                             const int32_t* __restrict__ cg_n, const int64_t* __restrict__ dense_off,
                             uint32_t* __restrict__ dense) {
  const int w = blockIdx.x;
  if (w >= n_wave) return;
  const int n = cg_n[w];
  const uint32_t* src = cg + cg_off[w];
  uint32_t* dst = dense + dense_off[w];
  for (int k = threadIdx.x; k < n; k += blockDim.x) dst[k] = src[n - 1 - k];
}


}  // namespace svb
