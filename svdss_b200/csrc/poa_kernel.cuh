// Device code of the cluster POA (see poa.cu for the design notes): parameters, workspace layout, graph
// primitives and the k_poa kernel.  Kept free of host / runtime-API code so that tests/emul can compile
// the very same source for the CPU with a lock-step warp emulator (tests/emul/warp_emul.hpp) and check
// kernel variants against the oracle without a GPU.
#pragma once
#include <stdint.h>

namespace svb {

constexpr int PNEG = -(1 << 29);
constexpr int POA_OK = 0, POA_OVERFLOW = 1, POA_CLAMPED = 2;

struct PoaParams {
  const uint8_t* __restrict__ seqs;
  const int64_t* __restrict__ seq_offs;      // n_seqs + 1
  const int64_t* __restrict__ cluster_offs;  // n_clusters + 1 (indexes seq_offs)
  const uint32_t* __restrict__ order;        // clusters of this launch, biggest first
  int n;                                     // clusters in this launch
  int n_slots;                               // workspace slots = warps: slot s takes cluster order[s] first, the rest are handed out by `work`
  unsigned int* work;
  // workspace: slot s (one per resident warp) = ws + slot_off[s]; a warp carves its slot anew for every cluster from that
  // cluster's own capacities dims[cluster] = (ncap, ecap, wcap, lmax) -- slots are sized for the biggest clusters in
  // hand-out order, every later cluster is smaller than any slot (poa.cu)
  uint8_t* ws;
  const int64_t* __restrict__ slot_off;
  const int4* __restrict__ dims;
  int swcap;          // columns per array of the shared-memory row copy (variants with POA_V_SMEM)
  // outputs
  uint8_t* cons;                   // cons_cap bytes per cluster, at cons_off[cluster]
  const int64_t* __restrict__ cons_off;
  int32_t* cons_len;
  int32_t* status;
  unsigned long long* cells;
  unsigned long long* phase;   // SVB_POA_TIMING: clock cycles per phase, summed over warps (lane 0)
  int match, mismatch, o1, e1, o2, e2, wb;
  float wf;
};

// per-slot workspace carving (all int32 unless noted); must match poa_ws_bytes()
struct PoaWs {
  uint8_t* base;
  int *rank, *order, *first_in, *last_in, *first_out, *last_out, *ring, *remain, *mpl, *mpr, *beg, *end, *cnt, *in1;
  int *efrom, *eto, *ew, *enin, *enout;
  int *op_node, *op_q, *new_anchor, *new_id;
  int *H, *E1, *E2;
  unsigned* TB;
  int4* rowinfo;   // per rank: (node, in1, c = ql - remain + 1, base) of the read being added (POA_V_ROWS)
};

__host__ __device__ inline int64_t align16(int64_t x) { return (x + 15) & ~(int64_t)15; }

__host__ __device__ inline int64_t poa_ws_carve(uint8_t* p, int ncap, int ecap, int wcap, int lmax, PoaWs* w) {
  int64_t o = 0;
  auto take = [&](int64_t bytes) { int64_t at = o; o = align16(o + bytes); return at; };
  int64_t a;
#define TAKE_I(f, n) a = take((int64_t)(n) * 4); if (w) w->f = reinterpret_cast<int*>(p + a)
  a = take(ncap); if (w) w->base = p + a;
  TAKE_I(rank, ncap); TAKE_I(order, ncap); TAKE_I(first_in, ncap); TAKE_I(last_in, ncap); TAKE_I(first_out, ncap);
  TAKE_I(last_out, ncap); TAKE_I(ring, ncap); TAKE_I(remain, ncap); TAKE_I(mpl, ncap); TAKE_I(mpr, ncap);
  TAKE_I(beg, ncap); TAKE_I(end, ncap); TAKE_I(cnt, ncap + 2); TAKE_I(in1, ncap);
  TAKE_I(efrom, ecap); TAKE_I(eto, ecap); TAKE_I(ew, ecap); TAKE_I(enin, ecap); TAKE_I(enout, ecap);
  TAKE_I(op_node, ncap + lmax + 4); TAKE_I(op_q, ncap + lmax + 4); TAKE_I(new_anchor, lmax + 2); TAKE_I(new_id, lmax + 2);
  a = take((int64_t)ncap * 16); if (w) w->rowinfo = reinterpret_cast<int4*>(p + a);
  TAKE_I(H, (int64_t)ncap * wcap); TAKE_I(E1, (int64_t)ncap * wcap); TAKE_I(E2, (int64_t)ncap * wcap);
  a = take((int64_t)ncap * wcap * 4); if (w) w->TB = reinterpret_cast<unsigned*>(p + a);
#undef TAKE_I
  return align16(o + 240) & ~(int64_t)255;
}

struct Graph {
  PoaWs w;
  int n, ne, ncap, ecap;
  bool overflow;
  __device__ int node(uint8_t b) {
    if (n >= ncap) { overflow = true; return ncap - 1; }
    const int v = n++;
    w.base[v] = b; w.rank[v] = -1; w.first_in[v] = w.last_in[v] = w.first_out[v] = w.last_out[v] = -1; w.ring[v] = v; w.in1[v] = 0;
    return v;
  }
  __device__ void edge(int u, int v) {
    for (int e = w.first_out[u]; e >= 0; e = w.enout[e])
      if (w.eto[e] == v) { w.ew[e]++; return; }
    if (ne >= ecap) { overflow = true; return; }
    const int e = ne++;
    w.efrom[e] = u; w.eto[e] = v; w.ew[e] = 1; w.enin[e] = -1; w.enout[e] = -1;
    if (w.last_out[u] < 0) w.first_out[u] = e; else w.enout[w.last_out[u]] = e;
    w.last_out[u] = e;
    if (w.last_in[v] < 0) { w.first_in[v] = e; w.in1[v] = u << 1; } else { w.enin[w.last_in[v]] = e; w.in1[v] |= 1; }
    w.last_in[v] = e;
  }
};

template <int G>
__device__ __forceinline__ int warp_incl_max(int v, int lane, unsigned gmask) {
#pragma unroll
  for (int o = 1; o < G; o <<= 1) {
    const int u = __shfl_up_sync(gmask, v, o, G);
    if (lane >= o) v = max(v, u);
  }
  return v;
}

// ---- DP rows of one read by the four warps of a CTA (k_poa<.., NW = 4>): warp w owns column group w of every chunk of
// 128 columns, all warps run the same row loop (same band arithmetic from the same inputs) and meet twice per chunk-row:
// once to hand the scans' carries and the last lane's flags across the groups (six ints per group through shared
// memory), once at the end of the row for the row maximum and the shared-memory copy of the row.  Same values, same
// comparisons as the one-warp row: the serial chain of a big cluster (30 reads x 5 000 rows of 121 columns) gets four
// schedulers instead of one.  xch: 9 * NW * 2 ints of shared memory.
template <int NW>
__device__ __forceinline__ void poa_cta_sync() {
#ifdef __CUDA_ARCH__
  __syncthreads();
#endif
}
template <int NW>
__device__ void poa_dp_rows_cta(const PoaParams& P, const PoaWs& W, const uint8_t* __restrict__ q, const int ql, const int n_ord, const int Wc,
                                const int Ws, int* const sbuf, int* const xch, const int wid, const int lane, const int mm, int& status,
                                unsigned long long& cells) {
  constexpr int G = 32;
  const unsigned gmask = 0xffffffffu;
  const int w = P.wb + (int)(P.wf * (float)ql);
  int4 ri_cur = make_int4(0, 0, 0, 0), ri_nxt = make_int4(0, 0, 0, 0);
  if (lane < n_ord) ri_nxt = W.rowinfo[lane];
  int v_prev = 0, b_prev = 0, en_prev = W.end[0], l_prev = 1, r_prev = 1;   // the source row
  bool prev_sm = en_prev < Ws;
  for (int r = 0; r < n_ord; ++r) {
    const int src = r & (G - 1);
    if (src == 0) {
      ri_cur = ri_nxt;
      if (r + G + lane < n_ord) ri_nxt = W.rowinfo[r + G + lane];
    }
    const int v = __shfl_sync(gmask, ri_cur.x, src, G), in1 = __shfl_sync(gmask, ri_cur.y, src, G);
    const int c = __shfl_sync(gmask, ri_cur.z, src, G), bvw = __shfl_sync(gmask, ri_cur.w, src, G);
    const int bv = bvw & 0xff;
    const bool need_g = (bvw >> 8) != 0;     // some reader of this row's scores is not the next row (poa_kernel.cuh, setup)
    const bool single = !(in1 & 1);
    const int p0 = in1 >> 1;
    int p0b, p0e, pl, pr;
    if (single) {
      if (p0 == v_prev) { p0b = b_prev; p0e = en_prev; pl = l_prev; pr = r_prev; }
      else { p0b = W.beg[p0]; p0e = W.end[p0]; pl = W.mpl[p0]; pr = W.mpr[p0]; }
    } else {
      p0b = 0; p0e = -1; pl = 0x7fffffff; pr = -1;
      for (int e = W.first_in[v]; e >= 0; e = W.enin[e]) {
        const int p = W.efrom[e];
        // the row just finished hands its band over in registers (its stores are ordered by the NEXT barrier only)
        pl = min(pl, p == v_prev ? l_prev : W.mpl[p]); pr = max(pr, p == v_prev ? r_prev : W.mpr[p]);
      }
    }
    int b = max(0, min(pl, c) - w), en = min(ql, max(pr, c) + w);
    if (b > en) b = en;
    if (en - b + 1 > Wc) { en = b + Wc - 1; status |= POA_CLAMPED; }
    int* hrow = W.H + (int64_t)v * Wc; int* e1row = W.E1 + (int64_t)v * Wc; int* e2row = W.E2 + (int64_t)v * Wc;
    unsigned* tbrow = W.TB + (int64_t)v * Wc;
    const int* const sprev = sbuf + (r & 1) * 3 * Ws;
    int* const scur = sbuf + ((r + 1) & 1) * 3 * Ws;
    const bool cur_sm = en - b + 1 <= Ws;
    int carry1 = PNEG, carry2 = PNEG, prevflags = 0;
    int rmax = PNEG - 1, rleft = 0, rright = 0;
    int par = 0;
    for (int jc = b; jc <= en; jc += G * NW, par ^= 1) {
      const int j = jc + wid * G + lane;
      const bool act = j <= en;
      int m = PNEG, x1 = PNEG, x2 = PNEG, pm = 0, p1 = 0, p2 = 0, x1ext = 0, x2ext = 0;
      const int qb = (act && j >= 1) ? q[j - 1] : 4;
      auto consider = [&](int p, int bp, int ep, int ord) {
        const int* ph = W.H + (int64_t)p * Wc;
        const int* pe1 = W.E1 + (int64_t)p * Wc;
        const int* pe2 = W.E2 + (int64_t)p * Wc;
        if (prev_sm && p == v_prev) { ph = sprev; pe1 = sprev + Ws; pe2 = sprev + 2 * Ws; }
        if (act && j >= 1 && j - 1 >= bp && j - 1 <= ep) {
          const int sc_ = (bv >= 4 || qb >= 4) ? 0 : (bv == qb ? P.match : -mm);
          const int cval = ph[j - 1 - bp] + sc_;
          if (cval > m) { m = cval; pm = ord; }
        }
        if (act && j >= bp && j <= ep) {
          const int hj = ph[j - bp];
          int op = hj - P.o1, ex = pe1[j - bp];
          int cval = max(op, ex) - P.e1;
          if (cval > x1) { x1 = cval; p1 = ord; x1ext = ex > op; }
          op = hj - P.o2; ex = pe2[j - bp];
          cval = max(op, ex) - P.e2;
          if (cval > x2) { x2 = cval; p2 = ord; x2ext = ex > op; }
        }
      };
      if (single) consider(p0, p0b, p0e, 0);
      else {
        int ord = 0;
        for (int e = W.first_in[v]; e >= 0; e = W.enin[e], ++ord) {
          const int p = W.efrom[e];
          consider(p, p == v_prev ? b_prev : W.beg[p], p == v_prev ? en_prev : W.end[p], ord);
        }
      }
      m = max(m, PNEG); x1 = max(x1, PNEG); x2 = max(x2, PNEG);
      int hp = m; unsigned hps = 0;
      if (x1 > hp) { hp = x1; hps = 1; }
      if (x2 > hp) { hp = x2; hps = 2; }
      const int B1 = act ? hp + j * P.e1 : PNEG, B2 = act ? hp + j * P.e2 : PNEG;
      const int inc1 = warp_incl_max<G>(B1, lane, gmask), inc2 = warp_incl_max<G>(B2, lane, gmask);
      int* const X = xch + (par * NW + wid) * 6;
      if (lane == G - 1) { X[0] = inc1; X[1] = inc2; X[4] = B1; X[5] = B2; }
      if (lane == G - 2) { X[2] = inc1; X[3] = inc2; }
      poa_cta_sync<NW>();
      // every warp replays the carries of all groups of the chunk: its own carry-in and flag-in, and the chunk's carry-out
      int cin1 = carry1, cin2 = carry2, pfl = prevflags;
#pragma unroll
      for (int g_ = 0; g_ < NW; ++g_) {
        if (g_ == wid) { cin1 = carry1; cin2 = carry2; pfl = prevflags; }
        const int* Y = xch + (par * NW + g_) * 6;
        const int x1_31 = max(carry1, Y[2]), x2_31 = max(carry2, Y[3]);       // lane 31 of group g_: X = max(carry, inclusive scan at lane 30)
        prevflags = (x1_31 > Y[4] ? 1 : 0) | (x2_31 > Y[5] ? 2 : 0);
        carry1 = max(carry1, Y[0]); carry2 = max(carry2, Y[1]);
      }
      int ex1 = __shfl_up_sync(gmask, inc1, 1, G), ex2 = __shfl_up_sync(gmask, inc2, 1, G);
      if (lane == 0) { ex1 = PNEG; ex2 = PNEG; }
      const int X1 = max(cin1, ex1), X2 = max(cin2, ex2);
      int f1 = PNEG, f2 = PNEG;
      if (j > b) { f1 = max(X1 - P.o1 - j * P.e1, PNEG); f2 = max(X2 - P.o2 - j * P.e2, PNEG); }
      const int myflags = (X1 > B1 ? 1 : 0) | (X2 > B2 ? 2 : 0);
      int pf = __shfl_up_sync(gmask, myflags, 1, G);
      if (lane == 0) pf = pfl;
      const int f1ext = (j > b) && (pf & 1), f2ext = (j > b) && (pf & 2);
      int hh = hp; unsigned hs = hps;
      if (f1 > hh) { hh = f1; hs = 3; }
      if (f2 > hh) { hh = f2; hs = 4; }
      if (act) {
        if (need_g || !cur_sm) { hrow[j - b] = hh; e1row[j - b] = x1; e2row[j - b] = x2; }
        if (cur_sm) { scur[j - b] = hh; scur[Ws + j - b] = x1; scur[2 * Ws + j - b] = x2; }
        tbrow[j - b] = hs | (hps << 3) | ((unsigned)x1ext << 5) | ((unsigned)x2ext << 6) | ((unsigned)f1ext << 7) | ((unsigned)f2ext << 8) |
                       ((unsigned)(pm & 0xff) << 12) | ((unsigned)(p1 & 0x3f) << 20) | ((unsigned)(p2 & 0x3f) << 26);
        if (hh > rmax) { rmax = hh; rleft = j; rright = j; }
        else if (hh == rmax) rright = j;
      }
    }
    cells += (unsigned long long)(en - b + 1);
    // row maximum with its first and last column: inside the warp, then across the warps
    int gmax = __reduce_max_sync(gmask, rmax);
    int l_ = __reduce_min_sync(gmask, (rmax == gmax) ? rleft : 0x7fffffff);
    int r_ = __reduce_max_sync(gmask, (rmax == gmax) ? rright : -1);
    int* const XR = xch + 12 * NW + (r & 1) * 3 * NW;
    if (lane == 0) { XR[wid * 3] = gmax; XR[wid * 3 + 1] = l_; XR[wid * 3 + 2] = r_; }
    poa_cta_sync<NW>();
    gmax = XR[0];
#pragma unroll
    for (int g_ = 1; g_ < NW; ++g_) gmax = max(gmax, XR[g_ * 3]);
    l_ = 0x7fffffff; r_ = -1;
#pragma unroll
    for (int g_ = 0; g_ < NW; ++g_) if (XR[g_ * 3] == gmax) { l_ = min(l_, XR[g_ * 3 + 1]); r_ = max(r_, XR[g_ * 3 + 2]); }
    if (wid == 0 && lane == 0) { W.beg[v] = b; W.end[v] = en; W.mpl[v] = l_ + 1; W.mpr[v] = r_ + 1; }
    v_prev = v; b_prev = b; en_prev = en; l_prev = l_ + 1; r_prev = r_ + 1; prev_sm = cur_sm;
  }
  poa_cta_sync<NW>();   // every store of the rows is done (and visible) before the master walks back through them
}

#ifndef SVB_POA_MINB
#define SVB_POA_MINB 4
#endif
// Variants (template bit mask V, SVB_POA_VARIANT on the host; 0 = the kernel measured in round 1).  The
// workspace of all resident warps is far larger than L2 (profiles/r01_poa_full.txt: 21 % L2 hit rate,
// long-scoreboard stalls dominate), so every dependent load of a warp's own graph or score rows is a DRAM
// round trip; each bit removes some of them and none changes a result (same values, same comparisons):
//  1 SMEM   the scores (H, E1, E2) of the row just finished are also kept in shared memory (two buffers per
//           warp, 6 * wcap ints); a row whose predecessor is that row -- the common case, a chain -- reads
//           them from there.  Needs the dynamic shared memory the host sizes from wcap.
//  2 TBIN1  traceback: the first predecessor comes from in1[v], loaded next to beg[v], instead of through
//           first_in -> efrom after the traceback word: two dependent loads per step, not four
//  4 PARN   the two per-read loops over all nodes (remain[], re-rank) are done by the whole warp
//  8, 16   (round 1: prefetching windows for the graph update and the traceback; measured without effect in round 2 --
//           profiles/r02a_variants_sweep.txt -- and replaced by UPDPAR / TBSPEC below; the bits are ignored now)
// 32 LEAN   DP rows with fewer dependent shuffles: the row maximum and its first / last column through REDUX
//           (__reduce_max_sync / __reduce_min_sync, 3 instructions instead of 15 shuffle steps), the two
//           "gap was extended" comparisons handed to the next lane as two bits instead of their four operands
//  64 ROWS    (needs PARN) the per-read setup leaves one 16-byte record per rank -- node, first predecessor, band
//            centre, base -- and the row loop fetches 32 of them with one coalesced load per lane, a block ahead,
//            handing them out by shuffle: no dependent order[] -> in1[] / remain[] / base[] chain per row any more
// 128 TBSPEC traceback by speculation: 32 lanes read the traceback words of the next 32 steps assuming the path
//            keeps running down the diagonal of a chain of consecutive node ids (the common case by far); a ballot
//            finds how many steps the guess holds for, those are emitted at once, lane 0 takes the odd step
// 256 UPDPAR graph update in windows of 32 alignment ops: every lane checks "its node carries the read's base and the
//            edge from the previous op's node exists" and bumps that edge's weight; lane 0 handles the first op of a
//            window that is not of that kind (new node, new edge, aligned-ring lookup) and the window restarts after it
// 512 / 1024 ILP2 / ILP4: a DP row is cut into chunks of 2 (4) x 32 columns and a lane works on its 2 (4) columns of a
//            chunk at once: the predecessor loads, the per-column recurrences and the two max-plus scans of the chunk's
//            column groups are independent instruction streams (the groups meet only in the carries), so a warp that
//            runs alone on its scheduler -- the tail of a batch is one big cluster, 30 reads x 5 000 rows of 121 columns
//            -- no longer pays one dependent-issue latency per instruction
constexpr int POA_V_SMEM = 1, POA_V_TBIN1 = 2, POA_V_PARN = 4, POA_V_LEAN = 32, POA_V_ROWS = 64,
              POA_V_TBSPEC = 128, POA_V_UPDPAR = 256, POA_V_ILP2 = 512, POA_V_ILP4 = 1024;


// G = lanes per cluster (32, 16 or 8; SVB_POA_GROUP on the host).  With G < 32 a warp carries 32/G clusters,
// each on its own group of lanes with its own control flow (group-masked shuffles and __syncwarp): the kernel
// is bound by the latency of dependent loads at a quarter of the issue rate, a band is 35-60 columns wide, and
// the sequential walks use one lane -- so more, narrower instruction streams per warp hide more latency with
// the same registers.  Everything below is written for a "group" of G lanes; `lane` is the lane in the group.
template <int V, int G = 32, int MB = SVB_POA_MINB, int NW = 1>
__global__ void __launch_bounds__(128, MB) k_poa(const PoaParams P) {
  extern __shared__ int poa_smem[];
  static_assert(NW == 1 || (NW == 4 && G == 32), "one warp per cluster, or the four warps of a CTA");
  static_assert(NW == 1 || ((V & (POA_V_SMEM | POA_V_ROWS | POA_V_PARN)) == (POA_V_SMEM | POA_V_ROWS | POA_V_PARN)), "the CTA rows need SMEM, ROWS and PARN");
  constexpr bool SMEM = (V & POA_V_SMEM) != 0, TBIN1 = (V & POA_V_TBIN1) != 0, PARN = (V & POA_V_PARN) != 0,
                 LEAN = (V & POA_V_LEAN) != 0, ROWS = (V & POA_V_ROWS) != 0, TBSPEC = (V & POA_V_TBSPEC) != 0,
                 UPDPAR = (V & POA_V_UPDPAR) != 0;
  constexpr int S = (V & POA_V_ILP4) ? 4 : (V & POA_V_ILP2) ? 2 : 0;   // column groups per chunk of the ILP row (0 = the one-group loop)
  static_assert(!ROWS || PARN, "ROWS builds its records in the warp-parallel setup");
  static_assert(G == 32 || G == 16 || G == 8, "group width");
  const int lane = threadIdx.x & (G - 1);
  const unsigned gmask = G == 32 ? 0xffffffffu : (((1u << (G & 31)) - 1u) << ((threadIdx.x & 31) & ~(G - 1)));
  const int slot = NW > 1 ? (int)blockIdx.x : (int)((blockIdx.x * blockDim.x + threadIdx.x) / G);
  const int wid = NW > 1 ? (int)(threadIdx.x >> 5) : 0;
  uint8_t* wsp = P.ws + P.slot_off[slot];
  Graph g;
  const PoaWs& W = g.w;
  const int mm = P.mismatch < 0 ? -P.mismatch : P.mismatch;
  // NW = 4: warp 0 is the master and runs everything below; warps 1-3 only join the DP rows of every read.  ctl (shared):
  // [0] command (1 = rows of a read, 2 = exit), [1] cluster, [2..3] sequence index, [4] nodes in rank order
  int* const xch = NW > 1 ? poa_smem + 6 * P.swcap : nullptr;
  volatile int* const ctl = NW > 1 ? poa_smem + 6 * P.swcap + 9 * NW * 2 : nullptr;
  if (NW > 1 && wid != 0) {
#ifdef __CUDA_ARCH__
    for (;;) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (ctl[0] == 2) return;
      const uint32_t cid_h = (uint32_t)ctl[1];
      const int64_t si_h = ((int64_t)(uint32_t)ctl[3] << 32) | (uint32_t)ctl[2];
      const int4 dmh = P.dims[cid_h];
      PoaWs Wh;
      poa_ws_carve(wsp, dmh.x, dmh.y, dmh.z, dmh.w, &Wh);
      int st_h = 0;
      unsigned long long cells_h = 0;
      poa_dp_rows_cta<NW>(P, Wh, P.seqs + P.seq_offs[si_h], (int)(P.seq_offs[si_h + 1] - P.seq_offs[si_h]), ctl[4], dmh.z, P.swcap, poa_smem, xch,
                          wid, (int)(threadIdx.x & 31), mm, st_h, cells_h);
    }
#endif
  }
  // shared copy of the row just finished: [2 buffers][H, E1, E2][Ws] per group.  Ws = P.swcap may be smaller than
  // the workspace row (wcap is a worst-case bound, a band is usually 35-60 columns): a row wider than Ws is simply
  // not copied and its successor reads the workspace (prev_sm says which).
  const int Ws = SMEM ? P.swcap : 0;
  int* const sbuf = SMEM ? poa_smem + (NW > 1 ? 0 : (int)(threadIdx.x / G) * 6 * Ws) : nullptr;
  bool prev_sm = false;
  for (bool first_item = true;; first_item = false) {
    unsigned wi = 0;
    if (first_item) wi = (unsigned)slot;             // slot s is sized for cluster order[s] (poa.cu): it takes that one first
    else {
      if (lane == 0) wi = (unsigned)P.n_slots + atomicAdd(P.work, 1u);
      wi = __shfl_sync(gmask, wi, 0, G);
    }
    if (wi >= (unsigned)P.n) { if (first_item) continue; break; }
    const uint32_t cid = P.order[wi];
    const int4 dm = P.dims[cid];
    poa_ws_carve(wsp, dm.x, dm.y, dm.z, dm.w, &g.w);
    g.ncap = dm.x; g.ecap = dm.y;
    const int Wc = dm.z, Lmax = dm.w;
    const int64_t s0 = P.cluster_offs[cid], s1 = P.cluster_offs[cid + 1];
    g.n = 0; g.ne = 0; g.overflow = false;
    int status = POA_OK;
    unsigned long long cells = 0;
    long long t_setup = 0, t_dp = 0, t_tb = 0, t_upd = 0, t_cons = 0, tc = clock64();
#define PHASE(acc) do { if (P.phase) { const long long _n = clock64(); acc += _n - tc; tc = _n; } } while (0)
    if (lane == 0) { g.node(0); g.node(0); }  // source, sink
    g.n = 2;
    __syncwarp(gmask);
    for (int64_t si = s0; si < s1 && !g.overflow; ++si) {
      const uint8_t* q = P.seqs + P.seq_offs[si];
      const int ql = (int)(P.seq_offs[si + 1] - P.seq_offs[si]);
      if (ql <= 0) continue;
      if (ql > Lmax) { g.overflow = true; break; }
      const int N = g.n;
      if (N == 2) {  // first read: a chain (warp-parallel)
        if (ql + 2 > g.ncap || ql + 1 > g.ecap) { g.overflow = true; break; }
        for (int j = lane; j < ql; j += G) {
          const int v = 2 + j;
          W.base[v] = q[j]; W.rank[v] = j; W.ring[v] = v;
          // edge j: (j ? v-1 : source) -> v ; edge ql: last -> sink
          W.efrom[j] = j ? v - 1 : 0; W.eto[j] = v; W.ew[j] = 1; W.enin[j] = -1; W.enout[j] = -1;
          W.first_in[v] = W.last_in[v] = j; W.in1[v] = (j ? v - 1 : 0) << 1;
          W.first_out[v] = W.last_out[v] = j + 1;
        }
        if (lane == 0) {
          W.efrom[ql] = 2 + ql - 1; W.eto[ql] = 1; W.ew[ql] = 1; W.enin[ql] = -1; W.enout[ql] = -1;
          W.first_out[0] = W.last_out[0] = 0;
          W.first_in[1] = W.last_in[1] = ql; W.in1[1] = (2 + ql - 1) << 1;
        }
        g.n = 2 + ql; g.ne = ql + 1;
        __syncwarp(gmask);
        continue;
      }
      const int n_ord = N - 2;
      for (int v = 2 + lane; v < N; v += G) { W.order[W.rank[v]] = v; }
      for (int s = lane; s <= n_ord + 1; s += G) W.cnt[s] = 0;
      __syncwarp(gmask);
      // remain[]: heaviest out-neighbour chain length to the sink, reverse rank order
      if (PARN) {
        // 32 ranks at a time: every lane finds the heaviest out-neighbour bv of its own node (independent
        // walks, their latencies overlap), then the recurrence remain[v] = remain[bv] + 1 is resolved inside
        // the chunk with shuffles -- lanes hold descending ranks, a node's successors have higher ranks, so
        // lane s is final once lanes < s are.  Two shuffles per node instead of four dependent loads.
        if (lane == 0) W.remain[1] = 0;
        __syncwarp(gmask);
        for (int r0 = n_ord - 1; r0 >= -1; r0 -= G) {
          const int r = r0 - lane;
          const bool valid = r >= -1;
          int v = -1, bv = 1;
          int need_g = 0;   // ROWS: some successor is not the next row in rank order (or is the sink): the row's scores must reach the workspace
          if (valid) {
            v = r >= 0 ? W.order[r] : 0;
            int bw = -1;
            for (int e = W.first_out[v]; e >= 0; e = W.enout[e]) {
              const int o_ = W.eto[e];
              if (W.ew[e] > bw) { bw = W.ew[e]; bv = o_; }
              if (ROWS && (o_ == 1 || W.rank[o_] != r + 1)) need_g = 1;
            }
          }
          int rem = valid ? W.remain[bv] : 0;      // final unless bv sits in this chunk (then replaced below)
          for (int s = 0; s < G; ++s) {
            const int vs = __shfl_sync(gmask, v, s, G);
            const int rs = __shfl_sync(gmask, rem, s, G) + 1;   // remain[vs], final at step s
            if (lane > s && valid && bv == vs) rem = rs;
          }
          if (valid) W.remain[v] = rem + 1;
          if (ROWS && valid && r >= 0) W.rowinfo[r] = make_int4(v, W.in1[v], ql - (rem + 1) + 1, (int)W.base[v] | (need_g << 8));
          __syncwarp(gmask);
        }
      } else if (lane == 0) {
        W.remain[1] = 0;
        for (int r = n_ord - 1; r >= -1; --r) {
          const int v = r >= 0 ? W.order[r] : 0;
          int bw = -1, bv = 1;
          for (int e = W.first_out[v]; e >= 0; e = W.enout[e])
            if (W.ew[e] > bw) { bw = W.ew[e]; bv = W.eto[e]; }
          W.remain[v] = W.remain[bv] + 1;
        }
      }
      __syncwarp(gmask);
      const int w = P.wb + (int)(P.wf * (float)ql);
      PHASE(t_setup);
      // ---- source row
      {
        int end0 = min(w, ql);
        if (end0 + 1 > Wc) { end0 = Wc - 1; status |= POA_CLAMPED; }
        for (int j = lane; j <= end0; j += G) {
          const int c1 = P.o1 + j * P.e1, c2 = P.o2 + j * P.e2;
          W.H[j] = j ? -min(c1, c2) : 0; W.E1[j] = PNEG; W.E2[j] = PNEG;
          if (SMEM && end0 < Ws) { sbuf[j] = j ? -min(c1, c2) : 0; sbuf[Ws + j] = PNEG; sbuf[2 * Ws + j] = PNEG; }   // buffer 0 = the row before rank 0
          unsigned t = j ? (c1 <= c2 ? 3u : 4u) : 0u;
          if (j > 1) t |= (c1 <= c2) ? (1u << 7) : (1u << 8);
          W.TB[j] = t;
        }
        // mpl[v] / mpr[v] = (first / last column of row v's maximum) + 1: what v offers to its
        // successors' bands.  A row PULLS the min / max over its in-edges (the same min / max abPOA
        // pushes along out-edges after each row) -- no out-edge walk, no per-read reset of the arrays.
        if (lane == 0) { W.beg[0] = 0; W.end[0] = end0; W.mpl[0] = 1; W.mpr[0] = 1; }
        __syncwarp(gmask);
      }
      // ---- graph rows in rank order
      // Row r needs: its node v (rank order), v's fields, its in-edges, the predecessors' bands, their
      // score rows -- a chain of dependent loads.  The node fields of row r+1 are fetched while row r is
      // computed, the first predecessor sits next to them (in1 = pred << 1 | has-more-in-edges), and a
      // predecessor that is the row just finished hands its band over in registers: the common row
      // (one in-edge, from the previous row) waits for its predecessor's scores only.
      if (NW > 1) {
#ifdef __CUDA_ARCH__
        if (lane == 0) { ctl[1] = (int)cid; ctl[2] = (int)(uint32_t)(si & 0xffffffffll); ctl[3] = (int)(uint32_t)((uint64_t)si >> 32); ctl[4] = n_ord; ctl[0] = 1; }
        __syncwarp(gmask);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        poa_dp_rows_cta<NW>(P, W, q, ql, n_ord, Wc, Ws, sbuf, xch, 0, lane, mm, status, cells);
#endif
      } else {
      int v_next = (!ROWS && n_ord > 0) ? W.order[0] : 0;
      int nx_in1 = ROWS ? 0 : W.in1[v_next], nx_remain = ROWS ? 0 : W.remain[v_next], nx_base = ROWS ? 0 : W.base[v_next];
      int4 ri_cur = make_int4(0, 0, 0, 0), ri_nxt = make_int4(0, 0, 0, 0);   // ROWS: this lane's record of the current / next block of G rows
      if (ROWS && lane < n_ord) ri_nxt = W.rowinfo[lane];
      int v_prev = 0, b_prev = 0, en_prev = W.end[0], l_prev = 1, r_prev = 1;   // the source row
      prev_sm = SMEM && en_prev < Ws;
      for (int r = 0; r < n_ord; ++r) {
        int v, in1, bv, c;
        bool need_g = true;            // ROWS: false when the only reader of this row's scores is the next row, through shared memory
        if (ROWS) {
          const int src = r & (G - 1);
          if (src == 0) {                               // a new block: the records fetched a block ago, and the next fetch goes out
            ri_cur = ri_nxt;
            if (r + G + lane < n_ord) ri_nxt = W.rowinfo[r + G + lane];
          }
          v = __shfl_sync(gmask, ri_cur.x, src, G); in1 = __shfl_sync(gmask, ri_cur.y, src, G);
          c = __shfl_sync(gmask, ri_cur.z, src, G); bv = __shfl_sync(gmask, ri_cur.w, src, G);
          need_g = (bv >> 8) != 0; bv &= 0xff;
        } else {
          v = v_next; in1 = nx_in1; bv = nx_base;
          c = ql - nx_remain + 1;
          if (r + 1 < n_ord) {                          // in flight while this row is computed
            v_next = W.order[r + 1];
            nx_in1 = W.in1[v_next]; nx_remain = W.remain[v_next]; nx_base = W.base[v_next];
          }
        }
        const bool single = !(in1 & 1);
        const int p0 = in1 >> 1;
        int p0b, p0e, pl, pr;
        if (single) {
          if (p0 == v_prev) { p0b = b_prev; p0e = en_prev; pl = l_prev; pr = r_prev; }
          else { p0b = W.beg[p0]; p0e = W.end[p0]; pl = W.mpl[p0]; pr = W.mpr[p0]; }
        } else {
          p0b = 0; p0e = -1; pl = 0x7fffffff; pr = -1;
          for (int e = W.first_in[v]; e >= 0; e = W.enin[e]) {
            const int p = W.efrom[e];
            pl = min(pl, W.mpl[p]); pr = max(pr, W.mpr[p]);
          }
        }
        int b = max(0, min(pl, c) - w), en = min(ql, max(pr, c) + w);
        if (b > en) b = en;
        if (en - b + 1 > Wc) { en = b + Wc - 1; status |= POA_CLAMPED; }
        int* hrow = W.H + (int64_t)v * Wc; int* e1row = W.E1 + (int64_t)v * Wc; int* e2row = W.E2 + (int64_t)v * Wc;
        unsigned* tbrow = W.TB + (int64_t)v * Wc;
        const int* const sprev = SMEM ? sbuf + (r & 1) * 3 * Ws : nullptr;        // scores of row v_prev (if prev_sm)
        int* const scur = SMEM ? sbuf + ((r + 1) & 1) * 3 * Ws : nullptr;
        const bool cur_sm = SMEM && en - b + 1 <= Ws;
        int carry1 = PNEG, carry2 = PNEG;      // running max of B1/B2 over columns before this segment
        int prevX1 = PNEG, prevB1 = PNEG, prevX2 = PNEG, prevB2 = PNEG;  // column j-1 of lane 0
        int prevflags = 0;                                               // LEAN: the same as two bits
        int rmax = PNEG - 1, rleft = 0, rright = 0;
        if (S > 0) {
          constexpr int SS = S > 0 ? S : 1;
          for (int jc = b; jc <= en; jc += G * SS) {
            int hp_[SS], x1_[SS], x2_[SS], inc1_[SS], inc2_[SS], B1_[SS], B2_[SS];
            unsigned tb_[SS];
#pragma unroll
            for (int g_ = 0; g_ < SS; ++g_) {          // independent per column group: loads + recurrences
              const int j = jc + g_ * G + lane;
              const bool act = j <= en;
              int m = PNEG, x1 = PNEG, x2 = PNEG, pm = 0, p1 = 0, p2 = 0, x1ext = 0, x2ext = 0;
              const int qb = (act && j >= 1) ? q[j - 1] : 4;
              auto consider = [&](int p, int bp, int ep, int ord) {
                const int* ph = W.H + (int64_t)p * Wc;
                const int* pe1 = W.E1 + (int64_t)p * Wc;
                const int* pe2 = W.E2 + (int64_t)p * Wc;
                if (SMEM && prev_sm && p == v_prev) { ph = sprev; pe1 = sprev + Ws; pe2 = sprev + 2 * Ws; }
                if (act && j >= 1 && j - 1 >= bp && j - 1 <= ep) {
                  const int sc_ = (bv >= 4 || qb >= 4) ? 0 : (bv == qb ? P.match : -mm);
                  const int cval = ph[j - 1 - bp] + sc_;
                  if (cval > m) { m = cval; pm = ord; }
                }
                if (act && j >= bp && j <= ep) {
                  const int hj = ph[j - bp];
                  int op = hj - P.o1, ex = pe1[j - bp];
                  int cval = max(op, ex) - P.e1;
                  if (cval > x1) { x1 = cval; p1 = ord; x1ext = ex > op; }
                  op = hj - P.o2; ex = pe2[j - bp];
                  cval = max(op, ex) - P.e2;
                  if (cval > x2) { x2 = cval; p2 = ord; x2ext = ex > op; }
                }
              };
              if (single) consider(p0, p0b, p0e, 0);
              else {
                int ord = 0;
                for (int e = W.first_in[v]; e >= 0; e = W.enin[e], ++ord) {
                  const int p = W.efrom[e];
                  consider(p, W.beg[p], W.end[p], ord);
                }
              }
              m = max(m, PNEG); x1 = max(x1, PNEG); x2 = max(x2, PNEG);
              int hp = m; unsigned hps = 0;
              if (x1 > hp) { hp = x1; hps = 1; }
              if (x2 > hp) { hp = x2; hps = 2; }
              hp_[g_] = hp; x1_[g_] = x1; x2_[g_] = x2;
              tb_[g_] = (hps << 3) | ((unsigned)x1ext << 5) | ((unsigned)x2ext << 6) | ((unsigned)(pm & 0xff) << 12) | ((unsigned)(p1 & 0x3f) << 20) |
                        ((unsigned)(p2 & 0x3f) << 26);
              B1_[g_] = act ? hp + j * P.e1 : PNEG; B2_[g_] = act ? hp + j * P.e2 : PNEG;
            }
#pragma unroll
            for (int g_ = 0; g_ < SS; ++g_) { inc1_[g_] = warp_incl_max<G>(B1_[g_], lane, gmask); inc2_[g_] = warp_incl_max<G>(B2_[g_], lane, gmask); }
#pragma unroll
            for (int g_ = 0; g_ < SS; ++g_) {          // the groups meet in the carries only
              const int j = jc + g_ * G + lane;
              const bool act = j <= en;
              int ex1 = __shfl_up_sync(gmask, inc1_[g_], 1, G), ex2 = __shfl_up_sync(gmask, inc2_[g_], 1, G);
              if (lane == 0) { ex1 = PNEG; ex2 = PNEG; }
              const int X1 = max(carry1, ex1), X2 = max(carry2, ex2);
              int f1 = PNEG, f2 = PNEG;
              if (j > b) { f1 = max(X1 - P.o1 - j * P.e1, PNEG); f2 = max(X2 - P.o2 - j * P.e2, PNEG); }
              const int myflags = (X1 > B1_[g_] ? 1 : 0) | (X2 > B2_[g_] ? 2 : 0);
              int pf = __shfl_up_sync(gmask, myflags, 1, G);
              if (lane == 0) pf = prevflags;
              const int f1ext = (j > b) && (pf & 1), f2ext = (j > b) && (pf & 2);
              int hh = hp_[g_]; unsigned hs = (tb_[g_] >> 3) & 3u;
              if (f1 > hh) { hh = f1; hs = 3; }
              if (f2 > hh) { hh = f2; hs = 4; }
              if (act) {
                if (need_g || !cur_sm) { hrow[j - b] = hh; e1row[j - b] = x1_[g_]; e2row[j - b] = x2_[g_]; }
                if (cur_sm) { scur[j - b] = hh; scur[Ws + j - b] = x1_[g_]; scur[2 * Ws + j - b] = x2_[g_]; }
                tbrow[j - b] = hs | tb_[g_] | ((unsigned)f1ext << 7) | ((unsigned)f2ext << 8);
                if (hh > rmax) { rmax = hh; rleft = j; rright = j; }
                else if (hh == rmax) rright = j;
              }
              carry1 = max(carry1, __shfl_sync(gmask, inc1_[g_], G - 1, G));
              carry2 = max(carry2, __shfl_sync(gmask, inc2_[g_], G - 1, G));
              prevflags = __shfl_sync(gmask, myflags, G - 1, G);
            }
          }
        } else
        for (int j0 = b; j0 <= en; j0 += G) {
          const int j = j0 + lane;
          const bool act = j <= en;
          int m = PNEG, x1 = PNEG, x2 = PNEG, pm = 0, p1 = 0, p2 = 0, x1ext = 0, x2ext = 0;
          const int qb = (act && j >= 1) ? q[j - 1] : 4;
          auto consider = [&](int p, int bp, int ep, int ord) {   // predecessor p with band [bp, ep], in-edge ordinal ord
            const int* ph = W.H + (int64_t)p * Wc;
            const int* pe1 = W.E1 + (int64_t)p * Wc;
            const int* pe2 = W.E2 + (int64_t)p * Wc;
            if (SMEM && prev_sm && p == v_prev) { ph = sprev; pe1 = sprev + Ws; pe2 = sprev + 2 * Ws; }
            if (act && j >= 1 && j - 1 >= bp && j - 1 <= ep) {
              const int s = (bv >= 4 || qb >= 4) ? 0 : (bv == qb ? P.match : -mm);
              const int cval = ph[j - 1 - bp] + s;
              if (cval > m) { m = cval; pm = ord; }
            }
            if (act && j >= bp && j <= ep) {
              const int hj = ph[j - bp];
              int op = hj - P.o1, ex = pe1[j - bp];
              int cval = max(op, ex) - P.e1;
              if (cval > x1) { x1 = cval; p1 = ord; x1ext = ex > op; }
              op = hj - P.o2; ex = pe2[j - bp];
              cval = max(op, ex) - P.e2;
              if (cval > x2) { x2 = cval; p2 = ord; x2ext = ex > op; }
            }
          };
          if (single) {
            consider(p0, p0b, p0e, 0);
          } else {
            int ord = 0;
            for (int e = W.first_in[v]; e >= 0; e = W.enin[e], ++ord) {
              const int p = W.efrom[e];
              consider(p, W.beg[p], W.end[p], ord);
            }
          }
          m = max(m, PNEG); x1 = max(x1, PNEG); x2 = max(x2, PNEG);
          int hp = m; unsigned hps = 0;
          if (x1 > hp) { hp = x1; hps = 1; }
          if (x2 > hp) { hp = x2; hps = 2; }
          // F1/F2: exclusive max-plus prefix scan of B(k) = Hp(k) + k*e over the row
          const int B1 = act ? hp + j * P.e1 : PNEG, B2 = act ? hp + j * P.e2 : PNEG;
          const int inc1 = warp_incl_max<G>(B1, lane, gmask), inc2 = warp_incl_max<G>(B2, lane, gmask);
          int ex1 = __shfl_up_sync(gmask, inc1, 1, G), ex2 = __shfl_up_sync(gmask, inc2, 1, G);
          if (lane == 0) { ex1 = PNEG; ex2 = PNEG; }
          const int X1 = max(carry1, ex1), X2 = max(carry2, ex2);
          int f1 = PNEG, f2 = PNEG;
          if (j > b) { f1 = max(X1 - P.o1 - j * P.e1, PNEG); f2 = max(X2 - P.o2 - j * P.e2, PNEG); }
          // ext flag of column j: F(j-1) > Hp(j-1) - o  <=>  X(j-1) > B(j-1)
          int f1ext, f2ext;
          const int myflags = (X1 > B1 ? 1 : 0) | (X2 > B2 ? 2 : 0);   // LEAN: the two comparisons travel, not the four operands
          if (LEAN) {
            int pf = __shfl_up_sync(gmask, myflags, 1, G);
            if (lane == 0) pf = prevflags;
            f1ext = (j > b) && (pf & 1); f2ext = (j > b) && (pf & 2);
          } else {
            int pX1 = __shfl_up_sync(gmask, X1, 1, G), pB1 = __shfl_up_sync(gmask, B1, 1, G);
            int pX2 = __shfl_up_sync(gmask, X2, 1, G), pB2 = __shfl_up_sync(gmask, B2, 1, G);
            if (lane == 0) { pX1 = prevX1; pB1 = prevB1; pX2 = prevX2; pB2 = prevB2; }
            f1ext = (j > b) && (pX1 > pB1); f2ext = (j > b) && (pX2 > pB2);
          }
          int hh = hp; unsigned hs = hps;
          if (f1 > hh) { hh = f1; hs = 3; }
          if (f2 > hh) { hh = f2; hs = 4; }
          if (act) {
            if (need_g || !cur_sm) { hrow[j - b] = hh; e1row[j - b] = x1; e2row[j - b] = x2; }
            if (cur_sm) { scur[j - b] = hh; scur[Ws + j - b] = x1; scur[2 * Ws + j - b] = x2; }
            tbrow[j - b] = hs | (hps << 3) | ((unsigned)x1ext << 5) | ((unsigned)x2ext << 6) | ((unsigned)f1ext << 7) |
                           ((unsigned)f2ext << 8) | ((unsigned)(pm & 0xff) << 12) | ((unsigned)(p1 & 0x3f) << 20) |
                           ((unsigned)(p2 & 0x3f) << 26);
            if (hh > rmax) { rmax = hh; rleft = j; rright = j; }
            else if (hh == rmax) rright = j;
          }
          carry1 = max(carry1, __shfl_sync(gmask, inc1, G - 1, G));
          carry2 = max(carry2, __shfl_sync(gmask, inc2, G - 1, G));
          if (LEAN) prevflags = __shfl_sync(gmask, myflags, G - 1, G);
          else {
            prevX1 = __shfl_sync(gmask, X1, G - 1, G); prevB1 = __shfl_sync(gmask, B1, G - 1, G);
            prevX2 = __shfl_sync(gmask, X2, G - 1, G); prevB2 = __shfl_sync(gmask, B2, G - 1, G);
          }
        }
        cells += (unsigned long long)(en - b + 1);
        // row maximum, its first and last column (lanes hold strided columns: reduce)
        int gmax = rmax, l_, r_;
        if (LEAN || S > 0) {   // REDUX: one instruction per reduction instead of log2(G) shuffle steps
          gmax = __reduce_max_sync(gmask, rmax);
          l_ = __reduce_min_sync(gmask, (rmax == gmax) ? rleft : 0x7fffffff);
          r_ = __reduce_max_sync(gmask, (rmax == gmax) ? rright : -1);
        } else {
#pragma unroll
          for (int o = G / 2; o; o >>= 1) gmax = max(gmax, __shfl_xor_sync(gmask, gmax, o, G));
          l_ = (rmax == gmax) ? rleft : 0x7fffffff; r_ = (rmax == gmax) ? rright : -1;
#pragma unroll
          for (int o = G / 2; o; o >>= 1) {
            l_ = min(l_, __shfl_xor_sync(gmask, l_, o, G));
            r_ = max(r_, __shfl_xor_sync(gmask, r_, o, G));
          }
        }
        if (lane == 0) { W.beg[v] = b; W.end[v] = en; W.mpl[v] = l_ + 1; W.mpr[v] = r_ + 1; }
        v_prev = v; b_prev = b; en_prev = en; l_prev = l_ + 1; r_prev = r_ + 1; prev_sm = cur_sm;
        __syncwarp(gmask);
      }
      }   // NW == 1
      PHASE(t_dp);
      // ---- end point, traceback, graph update, re-rank: lane 0
      int n_new_b = 0, nop_b = 0;
      // graph update (abpoa_add_graph_alignment), one alignment op at a time in forward order; lane 0's copies
      // of the state are the ones that count
      int u_n_new = 0, u_prev = 0, u_anchor = 0;
      const int n_old = N;
      auto add_op = [&](int k) -> bool {   // false: capacity overflow, stop
        if (g.overflow) return false;
        const int v = W.op_node[k], qi = W.op_q[k];
        if (v >= 0) {
          int mr = W.rank[v];
          for (int u = W.ring[v]; u != v; u = W.ring[u]) if (u < n_old && W.rank[u] > mr) mr = W.rank[u];
          u_anchor = mr + 1;
        }
        if (qi < 0) return true;
        int use;
        if (v >= 0) {
          const uint8_t bq = q[qi];
          if (W.base[v] == bq) use = v;
          else {
            use = -1;
            for (int u = W.ring[v]; u != v; u = W.ring[u]) if (W.base[u] == bq) { use = u; break; }
            if (use < 0) {
              use = g.node(bq);
              if (g.overflow) return false;
              W.ring[use] = W.ring[v]; W.ring[v] = use;
              W.new_anchor[u_n_new] = u_anchor; W.new_id[u_n_new] = use; ++u_n_new; W.cnt[u_anchor]++;
            }
          }
        } else {
          use = g.node(q[qi]);
          if (g.overflow) return false;
          W.new_anchor[u_n_new] = u_anchor; W.new_id[u_n_new] = use; ++u_n_new; W.cnt[u_anchor]++;
        }
        g.edge(u_prev, use);
        u_prev = use;
        return true;
      };
      auto finish_update = [&]() {
        if (!g.overflow) g.edge(u_prev, 1);
        n_new_b = u_n_new;
        if (!PARN && !g.overflow) {
          // re-rank: exclusive prefix of cnt over slots (slot 0 = source, r+1 = old rank r)
          int acc = 0;
          for (int s = 0; s <= n_ord; ++s) { const int c_ = W.cnt[s]; W.cnt[s] = acc; acc += c_; }
          for (int r = 0; r < n_ord; ++r) W.rank[W.order[r]] = r + W.cnt[r + 1];
          // new nodes follow their anchor in creation order: reuse mpl[] as the per-slot counter
          for (int k = 0; k < u_n_new; ++k) W.mpl[W.new_anchor[k] < N ? W.new_anchor[k] : 0] = 0;
          for (int k = 0; k < u_n_new; ++k) {
            const int s = W.new_anchor[k];
            // slots range over 0..n_ord <= N-2, so mpl[s] is a valid scratch cell
            W.rank[W.new_id[k]] = (s - 1 + W.cnt[s]) + 1 + W.mpl[s];
            W.mpl[s]++;
          }
        }
      };
      // traceback state (lane 0's copies count): position (t_v, t_j), DP state, ops written so far
      int t_v = 0, t_j = 0, t_state = 0, nop = 0;   // state: 0 H, 5 Hp, 1 E1, 2 E2, 3 F1, 4 F2
      auto tb_step = [&]() {
        int v = t_v, j = t_j, state = t_state;
        if (v == 0) { W.op_node[nop] = -1; W.op_q[nop] = j - 1; ++nop; t_j = j - 1; return; }
        // TBIN1: the first predecessor comes from in1[v] (= efrom[first_in[v]] << 1 | more), loaded next to beg[v]
        const int in1v = TBIN1 ? W.in1[v] : 0;
        const unsigned t = W.TB[(int64_t)v * Wc + j - W.beg[v]];
        if (state == 0) state = (int)(t & 7);
        else if (state == 5) state = (int)((t >> 3) & 3);
        if (state == 0) {
          int ord = (int)((t >> 12) & 0xff);
          W.op_node[nop] = v; W.op_q[nop] = j - 1; ++nop;
          if (TBIN1 && ord == 0) v = in1v >> 1;
          else { int e = W.first_in[v]; while (ord--) e = W.enin[e]; v = W.efrom[e]; }
          --j; state = 0;
        } else if (state == 1 || state == 2) {
          int ord = (int)((t >> (state == 1 ? 20 : 26)) & 0x3f);
          const int ext = (int)((t >> (state == 1 ? 5 : 6)) & 1);
          W.op_node[nop] = v; W.op_q[nop] = -1; ++nop;
          if (TBIN1 && ord == 0) v = in1v >> 1;
          else { int e = W.first_in[v]; while (ord--) e = W.enin[e]; v = W.efrom[e]; }
          if (!ext) state = 0;
          if (v == 0) state = 0;
        } else {
          const int ext = (int)((t >> (state == 3 ? 7 : 8)) & 1);
          W.op_node[nop] = -1; W.op_q[nop] = j - 1; ++nop;
          --j;
          if (!ext) state = 5;
        }
        t_v = v; t_j = j; t_state = state;
      };
      if (lane == 0) {
        int best_p = -1, best = PNEG - 1;
        for (int e = W.first_in[1]; e >= 0; e = W.enin[e]) {
          const int p = W.efrom[e];
          const int val = (ql >= W.beg[p] && ql <= W.end[p]) ? W.H[(int64_t)p * Wc + ql - W.beg[p]] : PNEG;
          if (val > best) { best = val; best_p = p; }
        }
        t_v = best_p; t_j = ql; t_state = 0;
        if (!TBSPEC) while (t_v != 0 || t_j > 0) tb_step();
      }
      if (TBSPEC) {
        for (;;) {
          const int cv = __shfl_sync(gmask, t_v, 0, G), cj = __shfl_sync(gmask, t_j, 0, G), cs = __shfl_sync(gmask, t_state, 0, G);
          const int cnop = __shfl_sync(gmask, nop, 0, G);
          if (cv == 0 && cj <= 0) break;
          int m = 0, nv = 0;
          if (cs == 0 && cv >= 2) {
            // lane k guesses step k: node cv - k, column cj - k, arriving in state H.  The guess holds when the cell's
            // H came from the diagonal through in-edge 0 (then the step emits (node, column - 1) and moves to the first
            // predecessor); the NEXT guess holds only if that predecessor is node - 1.
            const int pv = cv - lane, pj = cj - lane;
            bool ok = pv >= 2 && pj >= 1;
            int pred = 0;
            if (ok) {
              const int in1v = W.in1[pv], bg = W.beg[pv], en_ = W.end[pv];
              ok = pj >= bg && pj <= en_;
              if (ok) {
                const unsigned t = W.TB[(int64_t)pv * Wc + pj - bg];
                ok = (t & 7u) == 0u && ((t >> 12) & 0xffu) == 0u;
                pred = in1v >> 1;
              }
            }
            const unsigned low = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
            const unsigned shift = (threadIdx.x & 31) & ~(G - 1);
            const unsigned okm = (__ballot_sync(gmask, ok) >> shift) & low, chm = (__ballot_sync(gmask, ok && pred == pv - 1) >> shift) & low;
            const int bad_ok = (~okm & low) ? __ffs((int)(~okm & low)) - 1 : G, bad_ch = (~chm & low) ? __ffs((int)(~chm & low)) - 1 : G;
            m = min(bad_ok, bad_ch + 1);
            if (m > G) m = G;
            if (lane < m) { W.op_node[cnop + lane] = pv; W.op_q[cnop + lane] = pj - 1; }
            nv = __shfl_sync(gmask, pred, m > 0 ? m - 1 : 0, G);
          }
          __syncwarp(gmask);
          if (lane == 0) {
            if (m > 0) { nop += m; t_v = nv; t_j = cj - m; t_state = 0; }
            if (m < G && (t_v != 0 || t_j > 0)) tb_step();
          }
          __syncwarp(gmask);
        }
      }
      if (lane == 0) {
        PHASE(t_tb);
        nop_b = nop;
        if (!UPDPAR) {
          for (int k = nop - 1; k >= 0; --k) if (!add_op(k)) break;
          finish_update();
        }
      }
      if (UPDPAR) {
        const int nop_all = __shfl_sync(gmask, nop_b, 0, G);
        const unsigned low = G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u);
        const unsigned shift = (threadIdx.x & 31) & ~(G - 1);
        int lastv = -1;                       // lane 0: node of the last fast op; its anchor is computed when a slow op needs it
        int k0 = nop_all - 1;
        while (k0 >= 0) {
          const int kk = k0 - lane;
          const int v = kk >= 0 ? W.op_node[kk] : -2, qi = kk >= 0 ? W.op_q[kk] : -1;
          const bool reg = kk >= 0 && v >= 0 && qi >= 0 && W.base[v] == q[qi];
          const unsigned rm = (__ballot_sync(gmask, reg) >> shift) & low;
          const int m1 = (~rm & low) ? __ffs((int)(~rm & low)) - 1 : G;
          int u = __shfl_up_sync(gmask, v, 1, G);
          const int up = __shfl_sync(gmask, u_prev, 0, G);
          if (lane == 0) u = up;
          int e = -1;
          if (lane < m1) for (e = W.first_out[u]; e >= 0; e = W.enout[e]) if (W.eto[e] == v) break;
          const unsigned fm = (__ballot_sync(gmask, lane < m1 && e >= 0) >> shift) & low;
          const int m2 = (~fm & low) ? __ffs((int)(~fm & low)) - 1 : G;
          if (lane < m2) W.ew[e]++;
          const int vlast = __shfl_sync(gmask, v, m2 > 0 ? m2 - 1 : 0, G);
          __syncwarp(gmask);
          bool stop = false;
          if (lane == 0) {
            if (m2 > 0) { u_prev = vlast; lastv = vlast; }
            if (m2 < G && k0 - m2 >= 0) {
              if (lastv >= 0) {               // abpoa's anchor of the last aligned op: 1 + the highest old rank in its aligned ring
                int mr = W.rank[lastv];
                for (int x = W.ring[lastv]; x != lastv; x = W.ring[x]) if (x < n_old && W.rank[x] > mr) mr = W.rank[x];
                u_anchor = mr + 1;
                lastv = -1;
              }
              stop = !add_op(k0 - m2);
            }
          }
          stop = __shfl_sync(gmask, (int)stop, 0, G) != 0;
          __syncwarp(gmask);
          if (stop) break;
          k0 -= m2 + (m2 < G ? 1 : 0);
        }
        if (lane == 0) finish_update();
      }
      if (PARN) {
        // the same re-rank by the whole warp: exclusive scan of cnt over the slots, old nodes shifted by the
        // new nodes anchored before them; only the (few) new nodes are placed by lane 0
        const bool ovf = __shfl_sync(gmask, (int)g.overflow, 0, G) != 0;
        const int n_new = __shfl_sync(gmask, n_new_b, 0, G);
        __syncwarp(gmask);
        if (!ovf) {
          int carry = 0;
          for (int s0 = 0; s0 <= n_ord; s0 += G) {
            const int s_ = s0 + lane;
            const int c_ = s_ <= n_ord ? W.cnt[s_] : 0;
            int inc = c_;
#pragma unroll
            for (int o = 1; o < G; o <<= 1) { const int u = __shfl_up_sync(gmask, inc, o, G); if (lane >= o) inc += u; }
            if (s_ <= n_ord) W.cnt[s_] = carry + inc - c_;
            carry += __shfl_sync(gmask, inc, G - 1, G);
          }
          __syncwarp(gmask);
          for (int r = lane; r < n_ord; r += G) W.rank[W.order[r]] = r + W.cnt[r + 1];
          __syncwarp(gmask);
          if (lane == 0) {
            for (int k = 0; k < n_new; ++k) W.mpl[W.new_anchor[k] < N ? W.new_anchor[k] : 0] = 0;
            for (int k = 0; k < n_new; ++k) {
              const int s_ = W.new_anchor[k];
              W.rank[W.new_id[k]] = (s_ - 1 + W.cnt[s_]) + 1 + W.mpl[s_];
              W.mpl[s_]++;
            }
          }
        }
      }
      PHASE(t_upd);
      // lane 0's graph size / overflow flag are the truth
      g.n = __shfl_sync(gmask, g.n, 0, G);
      g.ne = __shfl_sync(gmask, g.ne, 0, G);
      g.overflow = __shfl_sync(gmask, (int)g.overflow, 0, G) != 0;
      __syncwarp(gmask);
    }
    // ---- consensus: heaviest bundling (lane 0)
    if (lane == 0) {
      int len = 0;
      if (!g.overflow && g.n > 2) {
        const int N = g.n, n_ord = N - 2;
        for (int v = 2; v < N; ++v) W.order[W.rank[v]] = v;
        int* score = W.remain; int* nxt = W.mpr;
        score[1] = 0;
        for (int r = n_ord - 1; r >= -1; --r) {
          const int v = r >= 0 ? W.order[r] : 0;
          int mw = -1, mi = -1;
          for (int e = W.first_out[v]; e >= 0; e = W.enout[e]) {
            const int o = W.eto[e], wgt = W.ew[e];
            if (mw < wgt) { mw = wgt; mi = o; }
            else if (mw == wgt && score[mi] <= score[o]) mi = o;
          }
          nxt[v] = mi;
          score[v] = mi >= 0 ? mw + score[mi] : 0;
        }
        uint8_t* out = P.cons + P.cons_off[cid];
        const int cap = (int)(P.cons_off[cid + 1] - P.cons_off[cid]);
        for (int v = nxt[0]; v > 1; v = nxt[v]) { if (len < cap) out[len] = W.base[v]; ++len; }
        if (len > cap) { status |= POA_OVERFLOW; }
      }
      if (g.overflow) status |= POA_OVERFLOW;
      P.cons_len[cid] = len;
      P.status[cid] = status;
      atomicAdd(P.cells, cells);
      PHASE(t_cons);
      if (P.phase) {
        atomicAdd(P.phase + 0, (unsigned long long)t_setup); atomicAdd(P.phase + 1, (unsigned long long)t_dp);
        atomicAdd(P.phase + 2, (unsigned long long)t_tb); atomicAdd(P.phase + 3, (unsigned long long)t_upd);
        atomicAdd(P.phase + 4, (unsigned long long)t_cons);
      }
    }
    __syncwarp(gmask);
  }
  if (NW > 1) {   // release the helpers for good
#ifdef __CUDA_ARCH__
    if (lane == 0) ctl[0] = 2;
    __syncwarp(gmask);
    asm volatile("bar.sync 1, 128;" ::: "memory");
#endif
  }
}


}  // namespace svb
