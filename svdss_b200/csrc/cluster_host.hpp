// Host half of the Clusterer (reference clusterer.cpp:8-52, 405-475): the order in which per-thread results are
// concatenated, the sort of the extended SFSs and cluster_by_proximity's two sequential sweeps.  They run over a few
// hundred thousand 20-byte records per batch and carry the reference's order-dependent quirks (interval cuts that
// look at the FIRST record of the interval only, per-thread std::maps keyed by (low, high) alone), so they stay a
// plain sequential restatement; the per-read and per-cluster work either side of them is cluster_core.cuh on the
// GPU.  Pure functions, no CUDA: tests/emul compiles this header for the CPU suite.
#pragma once
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <map>
#include <utility>
#include <vector>

#include "../../include/svdss_b200.h"
#include "cluster_core.cuh"

namespace svb {

struct ClExtSfs { int aln, rs, re, qs, qe; };

struct ClRaw {               // one raw cluster = the extended SFSs under one (low, high) key of one thread's map
  int tid;                   // Cluster::chrom = chrom of its first SFS
  std::vector<ClExtSfs> sfs;
};

// ext_of[sfs_offs[a] .. + n_ext[k]) = merged extended SFSs of the k-th accepted read (alignment a = accepted[k]; accepted =
// carries SFSs, in BAM order).  The
// reference deals accepted reads round robin to `threads` slots (clusterer.cpp:109-133) and concatenates the slots
// (:21-25); then std::sort by (chrom name, rs) -- stable here, the reference's tie order is unspecified -- and
// cluster_by_proximity (:405-475).  `rank_of_tid` orders the chromosome NAMES (SFS::operator<, sfs.hpp:64-72).
inline void cl_cluster_by_proximity(const int32_t* accepted, int n_acc, const int32_t* n_ext, const ClExt* ext_of, const int64_t* sfs_offs,
                                    const int32_t* tid_of_aln, const int32_t* rank_of_tid, int threads, std::vector<ClRaw>& out,
                                    int64_t& n_extended, int& max_ext_len, int& dist) {
  const size_t T = (size_t)std::max(1, threads);
  std::vector<ClExtSfs> ext;
  {
    size_t tot = 0;
    for (int i = 0; i < n_acc; ++i) tot += (size_t)n_ext[i];
    ext.reserve(tot);
  }
  for (size_t t = 0; t < T; ++t)
    for (size_t k = t; k < (size_t)n_acc; k += T) {
      const int a = accepted[k];
      const ClExt* e = ext_of + sfs_offs[a];
      for (int x = 0; x < n_ext[k]; ++x) ext.push_back(ClExtSfs{a, e[x].rs, e[x].re, e[x].qs, e[x].qe});
    }
  n_extended = (int64_t)ext.size();
  max_ext_len = 0; dist = 0;
  out.clear();
  if (ext.empty()) return;
  auto chrom_rank = [&](const ClExtSfs& s) { const int t = tid_of_aln[s.aln]; return rank_of_tid ? rank_of_tid[t] : t; };
  std::stable_sort(ext.begin(), ext.end(), [&](const ClExtSfs& a, const ClExtSfs& b) {
    const int ra = chrom_rank(a), rb = chrom_rank(b);
    return ra != rb ? ra < rb : a.rs < b.rs;
  });
  for (const ClExtSfs& s : ext) max_ext_len = std::max(max_ext_len, s.re - s.rs);
  dist = (int)((double)max_ext_len * 1.1);
  // :419-441 -- intervals for the parallel loop; prev_e is the `re` of the interval's FIRST record
  std::vector<std::pair<size_t, size_t>> intervals;
  size_t prev_i = 0;
  int prev_e = ext[0].re, prev_chrom = chrom_rank(ext[0]);
  for (size_t i = 1; i < ext.size(); ++i) {
    const int cr = chrom_rank(ext[i]);
    if (cr != prev_chrom) { prev_chrom = cr; intervals.emplace_back(prev_i, i - 1); prev_i = i; prev_e = ext[i].re; }
    else if (ext[i].rs - prev_e > dist) { intervals.emplace_back(prev_i, i - 1); prev_e = ext[i].re; prev_i = i; }
  }
  intervals.emplace_back(prev_i, ext.size() - 1);
  // :443-474 -- schedule(static, 1): interval i belongs to thread i % T; each thread owns a std::map keyed by
  // (low, high) ALONE, so equal keys of one thread merge even across chromosomes
  std::vector<std::map<std::pair<int, int>, std::vector<ClExtSfs>>> maps(T);
  for (size_t i = 0; i < intervals.size(); ++i) {
    auto& mine = maps[i % T];
    size_t j = intervals[i].first;
    int low = ext[j].rs, high = ext[j].re;
    size_t last_j = j;
    for (++j; j <= intervals[i].second; ++j) {
      if (ext[j].rs <= high) { low = std::min(low, ext[j].rs); high = std::max(high, ext[j].re); }
      else {
        auto& v = mine[std::make_pair(low, high)];
        v.insert(v.end(), ext.begin() + (long)last_j, ext.begin() + (long)j);
        low = ext[j].rs; high = ext[j].re; last_j = j;
      }
    }
    auto& v = mine[std::make_pair(low, high)];
    v.insert(v.end(), ext.begin() + (long)last_j, ext.begin() + (long)intervals[i].second + 1);
  }
  for (size_t t = 0; t < T; ++t)
    for (auto& kv : maps[t]) {   // :33-36
      ClRaw c;
      c.tid = tid_of_aln[kv.second[0].aln];
      c.sfs.swap(kv.second);
      out.push_back(std::move(c));
    }
}

// what fill_clusters needs of a raw cluster before it looks at any alignment (clusterer.cpp:497-520)
struct ClDesc {
  int tid, min_s, max_e;
  int n_reads;               // distinct reads among its SFSs (the std::set<string> of qnames)
  std::vector<int> alns;     // those reads, ascending alignment index
};

inline ClDesc cl_describe(const ClRaw& c) {
  ClDesc d;
  d.tid = c.tid; d.min_s = 0x7fffffff; d.max_e = 0;
  for (const ClExtSfs& s : c.sfs) { d.min_s = std::min(d.min_s, s.rs); d.max_e = std::max(d.max_e, s.re); d.alns.push_back(s.aln); }
  std::sort(d.alns.begin(), d.alns.end());
  d.alns.erase(std::unique(d.alns.begin(), d.alns.end()), d.alns.end());
  d.n_reads = (int)d.alns.size();
  return d;
}

// ---- the host work between the extend stage and the fill stage, and the assembly of svb_clusters_t after it.
// Shared by cluster.cu (stages = kernels) and tests/emul/cluster_emul.cpp (stages = loops over the same per-item
// functions).

struct ClPlan {
  std::vector<ClDesc> desc;            // every raw cluster, output order
  std::vector<int32_t> f_cluster;      // the ones that go to the fill stage (enough reads, clusterer.cpp:516-519)
  std::vector<int32_t> f_min_s, f_max_e, f_lo, f_hi, f_members;
  std::vector<int64_t> f_moff, f_rvoff;
  int64_t small_clusters = 0, n_extended = 0;
  int max_ext_len = 0, dist = 0;
};

// ext[sfs_offs[a] .. + n_ext[i]) = merged extended SFSs of accepted read i (alignment a = accepted[i])
inline void cl_plan_fill(const int32_t* accepted, int n_acc, const int32_t* n_ext, const ClExt* ext, const int64_t* sfs_offs,
                         const int32_t* tid, const int32_t* pos, int max_span, int64_t n_aln, const int32_t* rank_of_tid,
                         int threads, int min_cluster_weight, ClPlan& P) {
  std::vector<ClRaw> raw;
  cl_cluster_by_proximity(accepted, n_acc, n_ext, ext, sfs_offs, tid, rank_of_tid, threads, raw, P.n_extended, P.max_ext_len, P.dist);
  // candidate alignments of a cluster: [lo, hi) in BAM order = the records of its chromosome with beg - max_span <= pos < end
  // (the in-memory stand-in for the .bai query of clusterer.cpp:485-492; max_span = the longest reference span of any
  // record, so nothing that overlaps is left out, and the fill stage tests every candidate's own end)
  P.desc.resize(raw.size());
  P.f_moff.assign(1, 0); P.f_rvoff.assign(1, 0);
  for (size_t c = 0; c < raw.size(); ++c) {
    P.desc[c] = cl_describe(raw[c]);
    const ClDesc& d = P.desc[c];
    if (d.n_reads < min_cluster_weight) { ++P.small_clusters; continue; }
    int64_t lo = 0, hi = 0;
    {
      const int32_t* t0p = std::lower_bound(tid, tid + n_aln, d.tid);
      const int32_t* t1p = std::upper_bound(t0p, tid + n_aln, d.tid);
      const int64_t t0 = t0p - tid, t1 = t1p - tid;
      if (t1 > t0) {
        const int beg = std::max(0, d.min_s - 1), end = d.max_e;   // region "chrom:min_s-max_e", 1-based inclusive for htslib
        lo = std::lower_bound(pos + t0, pos + t1, beg - max_span) - pos;
        hi = std::lower_bound(pos + t0, pos + t1, end) - pos;
        if (hi < lo) hi = lo;
      }
    }
    P.f_cluster.push_back((int32_t)c);
    P.f_min_s.push_back(d.min_s); P.f_max_e.push_back(d.max_e); P.f_lo.push_back((int32_t)lo); P.f_hi.push_back((int32_t)hi);
    P.f_members.insert(P.f_members.end(), d.alns.begin(), d.alns.end());
    P.f_moff.push_back((int64_t)P.f_members.size());
    P.f_rvoff.push_back(P.f_rvoff.back() + (hi - lo));
  }
}

template <class T>
inline T* cl_host_alloc(size_t n) { return (T*)calloc(std::max<size_t>(n, 1), sizeof(T)); }

// results of the fill stage (indexed like P.f_*) -> svb_clusters_t; false = out of host memory
inline bool cl_assemble(const ClPlan& P, const int32_t* n_sub, const int32_t* sub_aln, const int32_t* sub_qs, const int32_t* sub_qe,
                        const int32_t* sub_hp, const int32_t* n_rv, const uint8_t* rvec, const int32_t* cov, int min_cluster_weight,
                        svb_clusters_t* out) {
  const int64_t nc = (int64_t)P.desc.size();
  const int nf = (int)P.f_cluster.size();
  out->n_clusters = nc;
  out->small_clusters = P.small_clusters; out->n_extended = P.n_extended; out->max_ext_len = P.max_ext_len; out->dist = P.dist;
  out->tid = cl_host_alloc<int32_t>((size_t)nc); out->s = cl_host_alloc<int32_t>((size_t)nc); out->e = cl_host_alloc<int32_t>((size_t)nc);
  out->cov0 = cl_host_alloc<int32_t>((size_t)nc); out->cov1 = cl_host_alloc<int32_t>((size_t)nc); out->cov2 = cl_host_alloc<int32_t>((size_t)nc);
  out->placed = cl_host_alloc<uint8_t>((size_t)nc);
  out->sub_offs = cl_host_alloc<int64_t>((size_t)nc + 1); out->rvec_offs = cl_host_alloc<int64_t>((size_t)nc + 1);
  int64_t tot_sub = 0, tot_rv = 0;
  for (int f = 0; f < nf; ++f) { tot_sub += n_sub[f]; if (n_sub[f] >= min_cluster_weight) tot_rv += n_rv[f]; }
  out->sub_aln = cl_host_alloc<int32_t>((size_t)tot_sub); out->sub_qs = cl_host_alloc<int32_t>((size_t)tot_sub);
  out->sub_qe = cl_host_alloc<int32_t>((size_t)tot_sub); out->sub_hp = cl_host_alloc<int32_t>((size_t)tot_sub);
  out->rvec = cl_host_alloc<uint8_t>((size_t)tot_rv);
  if (!out->tid || !out->s || !out->e || !out->cov0 || !out->cov1 || !out->cov2 || !out->placed || !out->sub_offs || !out->rvec_offs ||
      !out->sub_aln || !out->sub_qs || !out->sub_qe || !out->sub_hp || !out->rvec)
    return false;
  int f = 0;
  int64_t so = 0, ro = 0;
  for (int64_t c = 0; c < nc; ++c) {
    out->tid[c] = P.desc[(size_t)c].tid;
    out->sub_offs[c] = so; out->rvec_offs[c] = ro;
    if (f < nf && P.f_cluster[(size_t)f] == c) {
      out->placed[c] = 1; out->s[c] = P.f_min_s[(size_t)f]; out->e[c] = P.f_max_e[(size_t)f];
      const int ns = n_sub[f];
      const int64_t m0 = P.f_moff[(size_t)f];
      memcpy(out->sub_aln + so, sub_aln + m0, (size_t)ns * 4); memcpy(out->sub_qs + so, sub_qs + m0, (size_t)ns * 4);
      memcpy(out->sub_qe + so, sub_qe + m0, (size_t)ns * 4); memcpy(out->sub_hp + so, sub_hp + m0, (size_t)ns * 4);
      so += ns;
      if (ns >= min_cluster_weight) {       // clusterer.cpp:592-600: coverage and RVEC only for clusters that keep enough sub-reads
        out->cov0[c] = cov[f * 3]; out->cov1[c] = cov[f * 3 + 1]; out->cov2[c] = cov[f * 3 + 2];
        memcpy(out->rvec + ro, rvec + P.f_rvoff[(size_t)f], (size_t)n_rv[f]);
        ro += n_rv[f];
      } else ++out->small_clusters_2;
      ++f;
    }
  }
  out->sub_offs[nc] = so; out->rvec_offs[nc] = ro;
  return true;
}

}  // namespace svb
