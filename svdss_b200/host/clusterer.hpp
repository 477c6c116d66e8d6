// Clusterer of `SVDSS call` (reference clusterer.cpp:8-54): the step between the `.sfs` file and
// Caller::pcall.  It places every SFS on the reference through its read's alignment, extends it to
// the nearest unique 7-mers of the 100-bp flanks, clusters the extended SFSs by proximity, and
// cuts the same reference window out of every read covering a cluster (one sub-read per read).
//
// Host side by design: it is BAM parsing, per-read CIGAR walks and a sort + sweep, a few seconds
// of CPU next to the GPU stages on either side.  htslib is not available offline, so the indexed
// region fetch of fill_clusters (clusterer.cpp:485-492) is served from an in-memory table built
// during the single sequential scan (alignment payloads are kept only for reads that carry SFSs).
//
// Output order follows the reference for a given --threads: accepted reads are dealt round-robin
// to thread slots (clusterer.cpp:109-133), per-thread results are concatenated in slot order
// (:21-25), intervals are dealt round-robin to per-thread std::maps keyed by (low, high) and
// concatenated (:33-36, :376-413).  Two deliberate differences: std::sort on equal (chrom, rs)
// keys is replaced by a stable sort (the reference's tie order is unspecified), and clusters that
// never received coordinates (fewer than min_cluster_weight reads, clusterer.cpp:449-452) are
// dropped instead of being written with uninitialised s/e.
#pragma once
#include <algorithm>
#include <climits>
#include <fstream>
#include <map>
#include <set>
#include <sstream>
#include <string>
#include <unordered_map>
#include <utility>
#include <vector>

#include "io.hpp"

namespace svdss {

struct SFS {  // sfs.hpp:31-79
  std::string chrom, qname;
  int rs = 0, re = 0, qs = 0, qe = 0, l = 0, htag = 0;
  SFS(const std::string& qname_, int qs_, int l_, int htag_) : qname(qname_), qs(qs_), qe(qs_ + l_), l(l_), htag(htag_) {}
  SFS(const std::string& chrom_, const std::string& qname_, int rs_, int re_, int qs_, int qe_, int htag_)
      : chrom(chrom_), qname(qname_), rs(rs_), re(re_), qs(qs_), qe(qe_), l(qe_ - qs_ + 1), htag(htag_) {}
  bool operator<(const SFS& c) const {  // sfs.hpp:64-72
    if (chrom.empty() || c.chrom.empty()) return qs < c.qs;
    return chrom == c.chrom ? rs < c.rs : chrom < c.chrom;
  }
};

// sfs.cpp:5-30: `<qname or *>\t<qs>\t<l>\t<htag>\t`, the name only on a read's first line
inline bool parse_sfsfile(const std::string& path, std::unordered_map<std::string, std::vector<SFS>>& out, size_t& total) {
  std::ifstream inf(path);
  if (!inf.is_open()) return false;
  std::string line, info[4], read_name;
  total = 0;
  while (std::getline(inf, line)) {
    std::stringstream ss(line);
    int i = 0;
    while (ss.good() && i < 4) ss >> info[i++];
    if (i < 4) continue;
    if (info[0] != "*") { read_name = info[0]; out[read_name] = std::vector<SFS>(); }
    out[read_name].push_back(SFS(read_name, atoi(info[1].c_str()), atoi(info[2].c_str()), atoi(info[3].c_str())));
    ++total;
  }
  return true;
}

struct SubRead { std::string name, seq; int htag; size_t size() const { return seq.size(); } };  // clusterer.hpp:24-36

struct Cluster {  // clusterer.hpp:38-139 (fields kept)
  std::string chrom;
  int s = 0, e = 0, cov = 0, cov0 = 0, cov1 = 0, cov2 = 0;
  std::vector<SFS> SFSs;
  std::vector<std::pair<int, int>> reads;  // (has an SFS here, haplotype 1/2 or 3 = untagged): RVEC
  std::vector<SubRead> subreads;
  bool placed = false;                     // set_coordinates() was called
  size_t size() const { return subreads.size(); }
  int get_len() const {  // integer mean, clusterer.hpp:103-111
    unsigned l = 0, n = 0;
    for (const auto& sr : subreads) { ++n; l += (unsigned)sr.size(); }
    return (int)(l / n);
  }
};

typedef std::vector<std::pair<int, int>> AlPairs;

// bam.cpp:92-134 get_aligned_pairs: (read_pos or -1, ref_pos or -1) per alignment column
inline AlPairs get_aligned_pairs(int32_t pos, const std::vector<uint32_t>& cigar) {
  AlPairs r;
  size_t n = 0;
  for (uint32_t c : cigar) if ((c & 0xf) != 5 && (c & 0xf) != 6) n += c >> 4;
  r.reserve(n);
  int ref = pos, rd = 0;
  for (uint32_t c : cigar) {
    const uint32_t op = c & 0xf, len = c >> 4;
    if (op == 0 || op == 7 || op == 8) { for (uint32_t i = 0; i < len; ++i) r.emplace_back(rd++, ref++); }
    else if (op == 1 || op == 4) { for (uint32_t i = 0; i < len; ++i) r.emplace_back(rd++, -1); }
    else if (op == 2 || op == 3) { for (uint32_t i = 0; i < len; ++i) r.emplace_back(-1, ref++); }
  }
  return r;
}

struct ClusterConfig {
  std::string bam, clusters_out, clips_out;
  int threads = 4, batch_size = 10000;
  unsigned flank = 100, ksize = 7, min_mapq = 20, min_cluster_weight = 2;  // config.hpp:84-89
  bool clipped = false;                                                    // config.hpp:94 (--clipped)
};

// clipper.hpp:21-43: a soft clip whose bases carry an SFS that could not be placed on the reference.
// `p` = alignment start (left clip, starting) or bam_endpos (right clip); `l` = clipped bases.
struct Clip {
  std::string name, chrom;
  unsigned p = 0, l = 0;
  bool starting = false;
  unsigned w = 0;
  Clip() {}
  Clip(const std::string& name_, const std::string& chrom_, unsigned p_, unsigned l_, bool starting_, unsigned w_ = 0)
      : name(name_), chrom(chrom_), p(p_), l(l_), starting(starting_), w(w_) {}
  bool operator<(const Clip& c) const { return p < c.p; }
};

class Clusterer {
 public:
  Clusterer(const ClusterConfig& c, const std::unordered_map<std::string, std::vector<SFS>>* sfss,
            const std::unordered_map<std::string, std::string>* chromosome_seqs)
      : cfg_(c), SFSs_(sfss), chroms_(chromosome_seqs) {}

  std::vector<Cluster> clusters;
  std::vector<Clip> clips;   // clusterer.hpp:183; filled only with --clipped
  // book keeping (clusterer.hpp:150-160)
  unsigned unplaced = 0, s_unplaced = 0, e_unplaced = 0, unknown = 0, unextended = 0, small_clusters = 0, small_clusters_2 = 0;
  size_t n_extended = 0;
  int max_ext_len = 0, dist = 0;
  std::string error;

  bool run() {
    if (!scan()) return false;
    if (!cfg_.clips_out.empty() && !store_clips()) return false;
    if (extended_.empty()) return true;
    cluster_by_proximity();
    fill_clusters();
    if (!cfg_.clusters_out.empty() && !store_clusters()) return false;
    return true;
  }

 private:
  struct Aln {      // one primary, mapq-passing record (what both passes look at)
    int32_t tid, pos, end, l_qseq;
    int hp;
    std::string qname;
    int64_t payload;  // index into payload_ or -1 (read carries no SFS)
  };
  struct Payload { std::vector<uint32_t> cigar; std::vector<uint8_t> seq4; };

  ClusterConfig cfg_;
  const std::unordered_map<std::string, std::vector<SFS>>* SFSs_;
  const std::unordered_map<std::string, std::string>* chroms_;
  std::vector<std::string> ref_names_;
  std::vector<Aln> alns_;                    // file order
  std::vector<Payload> payload_;
  std::vector<std::vector<size_t>> by_tid_;  // indices into alns_, file order
  std::vector<std::vector<int32_t>> pmax_end_;
  std::vector<char> sorted_tid_;
  std::vector<SFS> extended_;

  // ---- pass 1 (clusterer.cpp:58-153): one sequential scan, reads dealt to thread slots
  bool scan() {
    BamReader bam(cfg_.bam);
    if (!bam.ok()) { error = "cannot read BAM " + cfg_.bam; return false; }
    bam.want_alignment(true);
    ref_names_ = bam.ref_names();
    by_tid_.resize(ref_names_.size());
    const int T = std::max(1, cfg_.threads);
    std::vector<std::vector<SFS>> p_ext((size_t)T);
    std::vector<size_t> accepted;   // indices into alns_
    BamRecord r;
    int st;
    while ((st = bam.next(r)) == 1) {
      if (r.flag & 0x4 || r.flag & 0x800 || r.flag & 0x100) continue;   // :116-120
      if (r.mapq < cfg_.min_mapq) continue;                             // :121-122
      if (r.tid < 0 || (size_t)r.tid >= ref_names_.size()) continue;
      const bool has = SFSs_->find(r.qname) != SFSs_->end();            // :123-125
      Aln a{r.tid, r.pos, r.endpos(), r.l_qseq, r.has_hp ? (int)r.hp : 0, r.qname, -1};
      if (has) {
        a.payload = (int64_t)payload_.size();
        payload_.push_back(Payload{r.cigar, r.seq4});
        accepted.push_back(alns_.size());
      }
      by_tid_[(size_t)r.tid].push_back(alns_.size());
      alns_.push_back(std::move(a));
    }
    if (st < 0) { error = "truncated or corrupt BAM " + cfg_.bam; return false; }
    // extend_alignment per accepted read; thread slot = n % threads (batch_size is a multiple of threads)
    std::vector<std::vector<SFS>> per_read(accepted.size());
    std::vector<unsigned> cnt(accepted.size() * 4, 0);
    std::vector<std::pair<unsigned, unsigned>> lr_clip(cfg_.clipped ? accepted.size() * 2 : 0, std::make_pair(0u, 0u));
#pragma omp parallel for schedule(dynamic, 64)
    for (long long n = 0; n < (long long)accepted.size(); ++n)
      extend_alignment(alns_[accepted[(size_t)n]], per_read[(size_t)n], &cnt[(size_t)n * 4],
                       cfg_.clipped ? &lr_clip[(size_t)n * 2] : nullptr);
    std::vector<std::vector<Clip>> p_clips((size_t)T);
    for (size_t n = 0; n < accepted.size(); ++n) {
      for (auto& s : per_read[n]) p_ext[n % (size_t)T].push_back(std::move(s));
      unplaced += cnt[n * 4]; s_unplaced += cnt[n * 4 + 1]; e_unplaced += cnt[n * 4 + 2]; unknown += cnt[n * 4 + 3];
      if (cfg_.clipped) {                                                                             // :339-345
        const Aln& a = alns_[accepted[n]];
        const std::string& chrom = ref_names_[(size_t)a.tid];
        if (lr_clip[n * 2].second > 0) p_clips[n % (size_t)T].push_back(Clip(a.qname, chrom, lr_clip[n * 2].first, lr_clip[n * 2].second, true));
        if (lr_clip[n * 2 + 1].second > 0) p_clips[n % (size_t)T].push_back(Clip(a.qname, chrom, lr_clip[n * 2 + 1].first, lr_clip[n * 2 + 1].second, false));
      }
    }
    for (int t = 0; t < T; ++t) {
      for (auto& s : p_ext[(size_t)t]) extended_.push_back(std::move(s));                            // :21-25
      clips.insert(clips.begin(), p_clips[(size_t)t].begin(), p_clips[(size_t)t].end());             // :24 (front insertion)
    }
    n_extended = extended_.size();
    // region-fetch tables
    pmax_end_.resize(by_tid_.size());
    sorted_tid_.assign(by_tid_.size(), 1);
    for (size_t t = 0; t < by_tid_.size(); ++t) {
      int32_t m = INT32_MIN, last = INT32_MIN;
      pmax_end_[t].reserve(by_tid_[t].size());
      for (size_t i : by_tid_[t]) {
        if (alns_[i].pos < last) sorted_tid_[t] = 0;
        last = alns_[i].pos;
        m = std::max(m, alns_[i].end);
        pmax_end_[t].push_back(m);
      }
    }
    return true;
  }

  // clusterer.cpp:156-345
  // `lr` (only with --clipped): [0] = left clip (alignment start, clipped bases), [1] = right clip
  // (bam_endpos, clipped bases) of a read whose SFS lies in a soft clip (:211-226)
  void extend_alignment(const Aln& aln, std::vector<SFS>& out, unsigned* cnt, std::pair<unsigned, unsigned>* lr) const {
    const std::string& chrom = ref_names_[(size_t)aln.tid];
    auto cit = chroms_->find(chrom);
    if (cit == chroms_->end()) return;                                   // :162-163
    const std::string& cseq = cit->second;
    const std::vector<uint32_t>& cig = payload_[(size_t)aln.payload].cigar;
    const AlPairs alpairs = get_aligned_pairs(aln.pos, cig);
    int last_pos = 0;
    std::vector<SFS> local;
    for (const SFS& sfs : SFSs_->at(aln.qname)) {
      const int s = sfs.qs, e = sfs.qs + sfs.l - 1;
      int aln_start = -1, aln_end = -1, refs = -1, refe = -1;
      for (size_t i = (size_t)last_pos; i < alpairs.size(); i++) {       // :183-201
        const int q = alpairs[i].first, r = alpairs[i].second;
        if (q == -1 || r == -1) continue;
        else if (q < s) { last_pos = (int)i; refs = r; aln_start = (int)i; }
        else if (q > e) { refe = r; aln_end = (int)i; break; }
      }
      if (refs == -1 && refe == -1) { ++cnt[0]; continue; }              // :206-211
      else if (refs == -1) {                                             // :211-218
        const uint32_t c0 = cig.empty() ? 0 : cig.front();
        if (lr && (c0 & 0xf) == 4) lr[0] = std::make_pair((unsigned)aln.pos, (unsigned)(c0 >> 4));
        else ++cnt[1];
        continue;
      } else if (refe == -1) {                                           // :219-226
        const uint32_t c1 = cig.empty() ? 0 : cig.back();
        if (lr && (c1 & 0xf) == 4) lr[1] = std::make_pair((unsigned)aln.end, (unsigned)(c1 >> 4));
        else ++cnt[2];
        continue;
      }
      AlPairs local_alpairs;
      {
        int last_r = refs - 1;
        for (int i = aln_start; i <= aln_end; i++) {                     // :229-244
          const int q = alpairs[(size_t)i].first, r = alpairs[(size_t)i].second;
          if (r == -1) { if (refs <= last_r && last_r <= refe) local_alpairs.emplace_back(q, r); }
          else { last_r = r; if (refs <= r && r <= refe) local_alpairs.emplace_back(q, r); }
          if (q != -1 && r != -1 && r >= refe) break;
        }
      }
      AlPairs pre, post;
      {
        unsigned n = 0;
        for (int i = aln_start - 1; i >= 0; --i) { pre.push_back(alpairs[(size_t)i]); if (++n == cfg_.flank) break; }   // :251-260
        std::reverse(pre.begin(), pre.end());
        n = 0;
        for (size_t i = (size_t)aln_end + 1; i < alpairs.size(); i++) { post.push_back(alpairs[i]); if (++n == cfg_.flank) break; }  // :263-271
      }
      std::pair<int, int> prekmer = get_unique_kmers(pre, cfg_.ksize, true, cseq);
      std::pair<int, int> postkmer = get_unique_kmers(post, cfg_.ksize, false, cseq);
      if (prekmer.first == -1 || prekmer.second == -1) prekmer = local_alpairs.front();     // :284-287
      if (postkmer.first == -1 || postkmer.second == -1) postkmer = local_alpairs.back();   // :288-291
      if (prekmer.first == -1 || prekmer.second == -1 || postkmer.first == -1 || postkmer.second == -1) { ++cnt[3]; continue; }  // :294-299
      if ((unsigned)prekmer.second > (unsigned)postkmer.second + cfg_.ksize) continue;     // :301-303 (warning only)
      local.push_back(SFS(chrom, aln.qname, prekmer.second, postkmer.second + (int)cfg_.ksize, prekmer.first,
                          postkmer.first + (int)cfg_.ksize, sfs.htag));
    }
    // merge overlapping extended SFSs of this read (:314-337)
    for (size_t i = 0; i < local.size(); ++i) {
      size_t j;
      for (j = 0; j < out.size(); ++j)
        if ((local[i].rs <= out[j].rs && out[j].rs <= local[i].re) || (out[j].rs <= local[i].rs && local[i].rs <= out[j].re)) break;
      if (j < out.size()) {
        out[j].rs = std::min(out[j].rs, local[i].rs); out[j].re = std::max(out[j].re, local[i].re);
        out[j].qs = std::min(out[j].qs, local[i].qs); out[j].qe = std::max(out[j].qe, local[i].qe);
      } else out.push_back(local[i]);
    }
  }

  // clusterer.cpp:350-403: first (from the inner end) clean k-mer of the flank that is unique in it
  static std::pair<int, int> get_unique_kmers(const AlPairs& alpairs, const unsigned k, const bool from_end, const std::string& cseq) {
    if (alpairs.size() < k) return std::make_pair(-1, -1);
    std::map<std::string, int> kmers;
    auto kmer_at = [&](int r) {
      // the reference reads k chars of a C string; never past the terminator
      return (size_t)r >= cseq.size() ? std::string() : cseq.substr((size_t)r, k);
    };
    size_t i = 0;
    while (i < alpairs.size() - k + 1) {
      bool skip = false;
      for (size_t j = i; j < i + k; j++)
        if (alpairs[j].first == -1 || alpairs[j].second == -1) { skip = true; i = j + 1; break; }
      if (skip) continue;
      ++kmers[kmer_at(alpairs[i].second)];
      ++i;
    }
    std::pair<int, int> last_kmer = std::make_pair(-1, -1);
    i = 0;
    while (i < alpairs.size() - k + 1) {
      size_t offset = i;
      if (from_end) offset = alpairs.size() - k - i;
      bool skip = false;
      for (size_t j = offset; j < offset + k; j++)
        if (alpairs[j].first == -1 || alpairs[j].second == -1) { skip = true; i += (j - offset); break; }
      if (skip) { ++i; continue; }
      last_kmer = alpairs[offset];
      if (kmers[kmer_at(alpairs[offset].second)] == 1) break;
      ++i;
    }
    return last_kmer;
  }

  std::vector<std::map<std::pair<int, int>, std::vector<SFS>>> p_sfs_clusters_;

  // clusterer.cpp:405-475
  void cluster_by_proximity() {
    std::stable_sort(extended_.begin(), extended_.end());
    for (const SFS& s : extended_) max_ext_len = std::max(max_ext_len, s.re - s.rs);
    dist = (int)((double)max_ext_len * 1.1);
    size_t prev_i = 0;
    int prev_e = extended_[0].re;
    std::string prev_chrom = extended_[0].chrom;
    std::vector<std::pair<size_t, size_t>> intervals;
    for (size_t i = 1; i < extended_.size(); i++) {
      const SFS& sfs = extended_[i];
      if (sfs.chrom != prev_chrom) {
        prev_chrom = sfs.chrom;
        intervals.emplace_back(prev_i, i - 1);
        prev_i = i; prev_e = sfs.re;
      } else if (sfs.rs - prev_e > dist) {
        intervals.emplace_back(prev_i, i - 1);
        prev_e = sfs.re; prev_i = i;
      }
    }
    intervals.emplace_back(prev_i, extended_.size() - 1);
    const size_t T = (size_t)std::max(1, cfg_.threads);
    p_sfs_clusters_.assign(T, std::map<std::pair<int, int>, std::vector<SFS>>());
    for (size_t i = 0; i < intervals.size(); i++) {
      auto& mine = p_sfs_clusters_[i % T];   // schedule(static, 1)
      size_t j = intervals[i].first;
      int low = extended_[j].rs, high = extended_[j].re;
      size_t last_j = j;
      j++;
      for (; j <= intervals[i].second; j++) {
        const SFS& sfs = extended_[j];
        if (sfs.rs <= high) { low = std::min(low, sfs.rs); high = std::max(high, sfs.re); }
        else {
          for (size_t k = last_j; k < j; k++) mine[std::make_pair(low, high)].push_back(extended_[k]);
          low = sfs.rs; high = sfs.re; last_j = j;
        }
      }
      for (size_t k = last_j; k <= intervals[i].second; k++) mine[std::make_pair(low, high)].push_back(extended_[k]);
    }
    for (size_t t = 0; t < T; ++t)
      for (auto& kv : p_sfs_clusters_[t]) {   // clusterer.cpp:33-36, Cluster(const vector<SFS>&)
        Cluster c;
        c.SFSs = kv.second;
        c.chrom = kv.second[0].chrom;
        clusters.push_back(std::move(c));
      }
    p_sfs_clusters_.clear();
  }

  // query base aligned to the last matched column with ref <= target (scan from the end, :558-568),
  // and to the first matched column with ref >= target (:569-579); -1 if none.  CIGAR walk instead
  // of materialising the pairs: matched columns are exactly the M/=/X runs.
  static void window_on_read(int32_t pos, const std::vector<uint32_t>& cigar, int min_s, int max_e, int& qs, int& qe) {
    qs = -1; qe = -1;
    int ref = pos, rd = 0;
    for (uint32_t c : cigar) {
      const uint32_t op = c & 0xf;
      const int len = (int)(c >> 4);
      if (op == 0 || op == 7 || op == 8) {
        if (len > 0) {
          if (ref <= min_s) { const int o = std::min(len - 1, min_s - ref); qs = rd + o; }   // later runs overwrite: last one wins
          if (qe == -1 && ref + len - 1 >= max_e) { const int o = std::max(0, max_e - ref); qe = rd + o; }
        }
        ref += len; rd += len;
      } else if (op == 1 || op == 4) rd += len;
      else if (op == 2 || op == 3) ref += len;
    }
  }

  // clusterer.cpp:478-610
  void fill_clusters() {
    std::unordered_map<std::string, int> tid_of;
    for (size_t t = 0; t < ref_names_.size(); ++t) tid_of.emplace(ref_names_[t], (int)t);   // first name wins, like bam_name2id
    std::vector<unsigned> cnt(clusters.size() * 3, 0);
#pragma omp parallel for schedule(dynamic, 16)
    for (long long ci = 0; ci < (long long)clusters.size(); ci++) {
      Cluster& cluster = clusters[(size_t)ci];
      std::set<std::string> reads;
      int min_s = INT_MAX, max_e = 0;
      for (const SFS& sfs : cluster.SFSs) { min_s = std::min(min_s, sfs.rs); max_e = std::max(max_e, sfs.re); reads.insert(sfs.qname); }
      if (reads.size() < cfg_.min_cluster_weight) { ++cnt[(size_t)ci * 3]; continue; }
      cluster.s = min_s; cluster.e = max_e; cluster.placed = true;
      int coverages[3] = {0, 0, 0};
      // region "chrom:min_s-max_e" is 1-based inclusive for htslib: records with pos < max_e and endpos > min_s-1
      const int beg = std::max(0, min_s - 1), end = max_e;
      auto tit = tid_of.find(cluster.chrom);
      if (tit != tid_of.end()) {
        const size_t t = (size_t)tit->second;
        const std::vector<size_t>& ids = by_tid_[t];
        size_t lo = 0, hi = ids.size();
        if (sorted_tid_[t]) {
          lo = (size_t)(std::upper_bound(pmax_end_[t].begin(), pmax_end_[t].end(), beg) - pmax_end_[t].begin());
          size_t a = lo, b = ids.size();
          while (a < b) { const size_t m = (a + b) / 2; if (alns_[ids[m]].pos < end) a = m + 1; else b = m; }
          hi = a;
        }
        for (size_t k = lo; k < hi; ++k) {
          const Aln& aln = alns_[ids[k]];
          if (!(aln.pos < end && aln.end > beg)) continue;
          const int hp_t = (aln.hp == 1 || aln.hp == 2) ? aln.hp : 0;   // other values index out of range in the reference
          ++coverages[hp_t];
          cluster.reads.emplace_back(0, hp_t == 0 ? 3 : hp_t);
          if (reads.find(aln.qname) == reads.end()) continue;
          cluster.reads.back().first = 1;
          // aln.payload is >= 0 here: the read carries SFSs, so pass 1 kept its alignment
          const Payload& pl = payload_[(size_t)aln.payload];
          int qs, qe;
          window_on_read(aln.pos, pl.cigar, min_s, max_e, qs, qe);
          if (qs == -1 || qe == -1) { ++cnt[(size_t)ci * 3 + 1]; continue; }
          std::string seq;
          if (qe >= qs) {
            seq.resize((size_t)(qe - qs + 1));
            for (int i = qs; i <= qe; ++i) { const uint8_t b = pl.seq4[(size_t)i >> 1]; seq[(size_t)(i - qs)] = nt16_char((i & 1) ? (b & 0xf) : (b >> 4)); }
          }
          cluster.subreads.push_back(SubRead{aln.qname, seq, hp_t});
        }
      }
      if (cluster.size() >= cfg_.min_cluster_weight) {
        cluster.cov0 = coverages[0]; cluster.cov1 = coverages[1]; cluster.cov2 = coverages[2];
        cluster.cov = cluster.cov0 + cluster.cov1 + cluster.cov2;
      } else {
        cluster.reads.clear();
        ++cnt[(size_t)ci * 3 + 2];
      }
    }
    for (size_t ci = 0; ci < clusters.size(); ++ci) { small_clusters += cnt[ci * 3]; unextended += cnt[ci * 3 + 1]; small_clusters_2 += cnt[ci * 3 + 2]; }
  }

  // test/debug output of `--clips FILE` (no reference counterpart): name, chrom, p, l, L|R in `clips` order
  bool store_clips() const {
    std::ofstream f(cfg_.clips_out);
    if (!f.is_open()) return false;
    for (const Clip& c : clips) f << c.name << "\t" << c.chrom << "\t" << c.p << "\t" << c.l << "\t" << (c.starting ? "L" : "R") << "\n";
    return true;
  }

  // clusterer.cpp:613-626
  bool store_clusters() const {
    std::ofstream f(cfg_.clusters_out);
    if (!f.is_open()) return false;
    for (const Cluster& c : clusters) {
      if (!c.placed) continue;
      f << c.chrom << ":" << c.s + 1 << "-" << c.e + 1 << "\t" << c.size();
      for (const SubRead& sr : c.subreads) f << "\t" << sr.name << ":" << sr.seq;
      f << "\n";
    }
    return true;
  }
};

}  // namespace svdss
