// Clipper of `SVDSS call --clipped` (reference clipper.cpp:3-215, caller.cpp:37-55):
// imprecise INS/DEL records from soft clips whose bases carry an SFS that could not be placed.
// Experimental in the reference and off by default (config.hpp:94).  Host side: a few thousand
// clips per genome, sequential containers.
//
// The pipeline per side (left = clip at the alignment start, right = clip at its end):
//   remove_duplicates (first clip per read name) -> combine (one clip per breakpoint, w = reads,
//   l = longest) -> filter_lowcovered (w >= 2) -> filter_tooclose_clips (not within 1000 bp of a
//   called SV) -> cluster (greedy, 1000 bp) -> sort by position;
// then INS = a left and a right clip less than 1000 bp apart, DEL = a right clip followed by a left
// clip 2000..50000 bp downstream with w >= 5.
//
// Kept from the reference on purpose, because they change the output:
//  * combine() walks a std::unordered_map<unsigned, ...> per chromosome, so the order in which
//    breakpoints reach cluster() is the container's.  The same container with the same insertion
//    order is used here, which gives the reference's order when both are built with libstdc++.
//  * cluster() keys its std::map by position only: clips of different chromosomes within 1000 bp
//    of each other's coordinate are merged, and `it->first - r` is unsigned (a cluster below 1000
//    never absorbs anything).
//  * the partner lookup compares positions only, never chromosomes (clipper.cpp:131-155).
// Not kept: binary_search() recursing with `m - 1` for m == 0 wraps to UINT_MAX and reads
// clips[2^31-1] (clipper.cpp:118-121; a crash when the query lies left of every clip) -- here that
// case returns clip 0, which is what the function's own comment asks for ("smallest right that is
// larger than query").  SV fields the reference leaves uninitialised for these records (cov0..2,
// gtq; sv.cpp:7-27) print as 0.
#pragma once
#include <algorithm>
#include <climits>
#include <cstdlib>
#include <map>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "clusterer.hpp"

namespace svdss {

class Clipper {
 public:
  // `chromosomes`: FASTA order (chromosomes.cpp:5); `seqs`: upper-cased sequences;
  // `sv_regions`: [sv.s - 1000, sv.e + 1000] of every called SV (caller.cpp:40-42), closed intervals
  Clipper(const std::vector<Clip>& clips, const std::vector<std::string>* chromosomes,
          const std::unordered_map<std::string, std::string>* seqs)
      : clips_(clips), chromosomes_(chromosomes), seqs_(seqs) {}

  struct Call { std::string type, chrom, refbase; unsigned s, w, l; };   // SV(type, chrom, s, refbase, "<type>", w, 0, 0, 0, true, l)
  std::vector<std::vector<Call>> p_calls;                                 // _p_svs, clipper.hpp:61
  std::vector<Clip> rclips, lclips;                                       // after preprocessing (tests, --verbose)
  std::vector<Clip> r_combined, l_combined;                               // in the order combine() produced

  void call(int threads, std::vector<std::pair<int, int>> sv_regions) {   // clipper.cpp:124-215
    std::sort(sv_regions.begin(), sv_regions.end());
    reg_lo_.clear(); reg_pmax_hi_.clear();
    int m = INT_MIN;
    for (const auto& r : sv_regions) { reg_lo_.push_back(r.first); m = std::max(m, r.second); reg_pmax_hi_.push_back(m); }
    for (const Clip& c : clips_) (c.starting ? lclips : rclips).push_back(c);
    preprocess(rclips, r_combined);
    preprocess(lclips, l_combined);
    const size_t T = (size_t)std::max(1, threads);
    p_calls.assign(T, std::vector<Call>());
    if (lclips.empty() || rclips.empty()) return;
    for (size_t i = 0; i < lclips.size(); i++) {             // insertions, :164-186; schedule(static, 1)
      const Clip& lc = lclips[i];
      const int r = partner(rclips, lc);
      if (r == -1) continue;
      const Clip& rc = rclips[(size_t)r];
      if (rc.w == 0) continue;
      if (std::abs((int)rc.p - (int)lc.p) < 1000) {
        const unsigned s = lc.w > rc.w ? lc.p : rc.p;
        p_calls[i % T].push_back(Call{"INS", lc.chrom, refbase(lc.chrom, s), s, std::max(lc.w, rc.w), std::max(lc.l, rc.l)});
      }
    }
    for (size_t i = 0; i < rclips.size(); i++) {             // deletions, :188-214
      const Clip& rc = rclips[i];
      const int l = partner(lclips, rc);
      if (l == -1) continue;
      const Clip& lc = lclips[(size_t)l];
      if (lc.w == 0) continue;
      const unsigned d = lc.p - rc.p;                        // unsigned like the reference: lc left of rc wraps and fails `<= 50000`
      if (d >= 2000 && d <= 50000) {
        const unsigned w = std::max(lc.w, rc.w);
        if (w >= 5) p_calls[i % T].push_back(Call{"DEL", rc.chrom, refbase(rc.chrom, rc.p), rc.p, w, d + 1});
      }
    }
  }

 private:
  std::vector<Clip> clips_;
  const std::vector<std::string>* chromosomes_;
  const std::unordered_map<std::string, std::string>* seqs_;
  std::vector<int> reg_lo_, reg_pmax_hi_;

  std::string refbase(const std::string& chrom, unsigned s) const {       // string(chromosome_seqs[chrom] + s, 1)
    auto it = seqs_->find(chrom);
    if (it == seqs_->end() || s >= it->second.size()) return std::string(1, 'N');
    return it->second.substr(s, 1);
  }

  void preprocess(std::vector<Clip>& v, std::vector<Clip>& combined) {    // :141-156
    v = remove_duplicates(v);
    v = combine(v);
    combined = v;
    v = filter_lowcovered(v, 2);
    v = filter_tooclose_clips(v);
    v = cluster(v, 1000);
    std::sort(v.begin(), v.end());                                         // keys of a std::map: positions are distinct
  }

  static std::vector<Clip> remove_duplicates(const std::vector<Clip>& clips) {   // :5-15
    std::vector<Clip> unique_clips;
    std::unordered_map<std::string, int> qnames;
    for (const Clip& clip : clips)
      if (qnames.find(clip.name) == qnames.end()) { qnames[clip.name] = 0; unique_clips.push_back(clip); }
    return unique_clips;
  }

  std::vector<Clip> combine(const std::vector<Clip>& clips) const {        // :17-52
    const int threads = 4;
    std::vector<std::vector<Clip>> p_combined((size_t)threads);
    std::unordered_map<std::string, std::unordered_map<unsigned, std::vector<Clip>>> clips_dict;
    for (const Clip& c : clips) clips_dict[c.chrom][c.p].push_back(c);
    for (size_t i = 0; i < chromosomes_->size(); i++) {                    // schedule(static, 1) over 4 slots
      const std::string& chrom = (*chromosomes_)[i];
      auto cit = clips_dict.find(chrom);
      if (cit == clips_dict.end()) continue;
      for (auto it = cit->second.begin(); it != cit->second.end(); ++it) {
        unsigned max_l = 0;
        for (const Clip& c : it->second) max_l = std::max(max_l, c.l);
        p_combined[i % (size_t)threads].push_back(Clip("", chrom, it->first, max_l, it->second.front().starting, (unsigned)it->second.size()));
      }
    }
    std::vector<Clip> combined;
    for (int i = 0; i < threads; i++) combined.insert(combined.begin(), p_combined[(size_t)i].begin(), p_combined[(size_t)i].end());
    return combined;
  }

  static std::vector<Clip> filter_lowcovered(const std::vector<Clip>& clips, unsigned w) {   // :54-63
    std::vector<Clip> out;
    for (const Clip& c : clips) if (c.w >= w) out.push_back(c);
    return out;
  }

  // :97-106: keep a clip unless [p, p+1] overlaps a closed SV region (lib_interval_tree's default interval kind)
  std::vector<Clip> filter_tooclose_clips(const std::vector<Clip>& clips) const {
    std::vector<Clip> out;
    for (const Clip& c : clips) {
      const int p = (int)c.p;
      // regions sorted by low: those with low <= p + 1 form a prefix; one of them reaches p iff the prefix maximum of high does
      const size_t n = (size_t)(std::upper_bound(reg_lo_.begin(), reg_lo_.end(), p + 1) - reg_lo_.begin());
      if (n == 0 || reg_pmax_hi_[n - 1] < p) out.push_back(c);
    }
    return out;
  }

  // :67-95: every existing cluster within r of the clip absorbs it (all of them, not only the first);
  // a clip absorbed by none starts a cluster at its own position.  The reference scans the whole map
  // per clip; the keys that can satisfy `key - r <= p && p <= key + r` in unsigned arithmetic are
  // exactly those in [max(r, p - r), p + r], visited here through lower_bound.
  static std::vector<Clip> cluster(const std::vector<Clip>& clips, unsigned r) {
    std::map<unsigned, Clip> by_pos;
    for (const Clip& c : clips) {
      bool found = false;
      const unsigned lo = std::max(r, c.p > r ? c.p - r : 0u);
      const unsigned long long hi = (unsigned long long)c.p + r;
      for (auto it = by_pos.lower_bound(lo); it != by_pos.end() && (unsigned long long)it->first <= hi; ++it) {
        found = true;
        it->second.l = std::max(it->second.l, c.l);
        it->second.w += c.w;
      }
      if (!found) by_pos[c.p] = c;
    }
    std::vector<Clip> out;
    for (auto& kv : by_pos) out.push_back(kv.second);
    return out;
  }

  // :107-122 binary_search(clips, 0, size - 1, query): index of the first clip right of the query
  // (the next one when a clip sits exactly at the query position, or that clip itself if it is
  // the last), -1 if every clip lies left of the query
  static int partner(const std::vector<Clip>& clips, const Clip& query) {
    long long begin = 0, end = (long long)clips.size() - 1;
    while (true) {
      if (begin > end || begin >= (long long)clips.size()) return -1;
      const long long m = (begin + end) / 2;
      if (clips[(size_t)m].p == query.p) return (int)((size_t)m + 1 < clips.size() ? m + 1 : m);
      if (clips[(size_t)m].p > query.p) {
        if (m > 0 && clips[(size_t)m - 1].p < query.p) return (int)m;
        if (m == 0) return 0;                      // the reference wraps `m - 1` to UINT_MAX here and reads out of bounds
        end = m - 1;
      } else begin = m + 1;
    }
  }
};

}  // namespace svdss
