// Clusterer of `SVDSS call` (reference clusterer.cpp:8-54): the step between the `.sfs` file and
// Caller::pcall.  It places every SFS on the reference through its read's alignment, extends it to
// the nearest unique 7-mers of the 100-bp flanks, clusters the extended SFSs by proximity, and
// cuts the same reference window out of every read covering a cluster (one sub-read per read).
//
// The scan of the BAM stays on the host (own BGZF/BAM reader, io.hpp); everything after it is svb_cluster_batch
// (csrc/cluster.cu: extend_alignment and fill_clusters as kernels, cluster_by_proximity as the host sweep between
// them).  htslib is not available offline, so the indexed region fetch of fill_clusters (clusterer.cpp:485-492) is
// served from the coordinate-sorted record arrays themselves; an unsorted BAM is refused (the reference needs the .bai).
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

#include "../../include/svdss_b200.h"
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

struct ClusterConfig {
  std::string bam, clusters_out, clips_out;
  int threads = 4, batch_size = 10000;
  unsigned flank = 100, ksize = 7, min_mapq = 20, min_cluster_weight = 2;  // config.hpp:84-89
  bool clipped = false;                                                    // config.hpp:94 (--clipped)
  int device = 0;
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
  bool scanned_on_device = false;   // the BAM scan ran through svb_bamstream_* (`--gpu-inflate`)

  // Clusterer::run (clusterer.cpp:8-52): one sequential scan of the BAM on the host (the records that pass the
  // filters become the arrays of svb_alns_t), then svb_cluster_batch does extend_alignment / cluster_by_proximity /
  // fill_clusters (the per-read and per-cluster work on the GPU), and the arrays that come back are turned into the
  // reference's Cluster objects for the rest of `call`.
  bool run() {
    if (!scan()) return false;
    if (!place_and_fill()) return false;
    if (!cfg_.clips_out.empty() && !store_clips()) return false;
    if (!cfg_.clusters_out.empty() && !store_clusters()) return false;
    return true;
  }

 private:
  ClusterConfig cfg_;
  const std::unordered_map<std::string, std::vector<SFS>>* SFSs_;
  const std::unordered_map<std::string, std::string>* chroms_;
  std::vector<std::string> ref_names_;
  // what Clusterer keeps of the BAM: primary, mapq-passing records in file order
  std::vector<int32_t> tid_, pos_, hp_;
  std::vector<std::string> qname_;
  std::vector<int64_t> cigar_offs_{0}, sfs_offs_{0};
  std::vector<uint32_t> cigar_;
  std::vector<int32_t> sfs_qs_, sfs_len_;
  std::vector<int64_t> payload_of_;             // per record: index into seq4_ or -1 (read carries no SFS)
  std::vector<std::vector<uint8_t>> seq4_;      // 4-bit sequences of the reads that carry SFSs

  // pass 1 on the device (`--gpu-inflate`, files of 1 GiB or more): the records are inflated, walked and parsed in HBM
  // (svb_bamstream_*, alignment mode); per record 28 bytes and the name come back, and the CIGAR + packed bases of
  // the reads that carry SFSs are fetched from the window where they still lie.  1 = done, 0 = not applicable, -1 = failed.
  int scan_device() {
    if (bgzf_gpu_device() < 0 || getenv("SVB_BAM_HOST_PARSE")) return 0;
    BgzfSource src(cfg_.bam);
    if (!src.ok() || !src.device_inflate()) return 0;
    int64_t header_bytes = 0;
    {
      const int dev = bgzf_gpu_device();
      bgzf_gpu_device() = -1;                          // the header is read by the host reader
      BamReader hdr(cfg_.bam, (size_t)1 << 20);
      bgzf_gpu_device() = dev;
      if (!hdr.ok()) { error = "cannot read BAM " + cfg_.bam; return -1; }
      ref_names_ = hdr.ref_names();
      header_bytes = hdr.header_bytes();
    }
    svb_bamstream_t* bs = nullptr;
    if (svb_bamstream_open(bgzf_gpu_device(), -1, (int)ref_names_.size(), &bs) != SVB_OK) { error = std::string("svb_bamstream_open: ") + svb_last_error(); return -1; }
    const uint8_t* base = nullptr;
    std::vector<int64_t> io, oo, want;
    std::vector<int8_t> keep;
    std::vector<const std::vector<SFS>*> sfs_of;
    int rc = 1;
    while (rc == 1 && src.next_members(base, io, oo)) {
      svb_bam_recs_t recs;
      if (svb_bamstream_window(bs, base, io.data(), oo.data(), (int64_t)io.size() - 1, header_bytes, &recs) != SVB_OK) {
        error = "truncated or corrupt BAM " + cfg_.bam + ": " + svb_last_error(); rc = -1; break;
      }
      // which records count (clusterer.cpp:116-125), and which of them carry SFSs
      keep.assign((size_t)recs.n, 0); sfs_of.assign((size_t)recs.n, nullptr); want.clear();
      for (int64_t i = 0; i < recs.n; ++i) {
        if (recs.state[i] == 0) continue;                                        // :116-120
        if (recs.mapq[i] < cfg_.min_mapq) continue;                               // :121-122
        if (recs.tid[i] < 0 || (size_t)recs.tid[i] >= ref_names_.size()) continue;
        keep[(size_t)i] = 1;
        auto it = SFSs_->find(std::string(recs.names + recs.name_offs[i], (size_t)(recs.name_offs[i + 1] - recs.name_offs[i])));   // :123-125
        if (it != SFSs_->end()) { sfs_of[(size_t)i] = &it->second; want.push_back(i); }
      }
      const uint32_t* f_cig = nullptr; const int64_t* f_co = nullptr; const uint8_t* f_seq = nullptr; const int64_t* f_so = nullptr;
      if (svb_bamstream_fetch(bs, want.data(), (int64_t)want.size(), &f_cig, &f_co, &f_seq, &f_so) != SVB_OK) { error = std::string("svb_bamstream_fetch: ") + svb_last_error(); rc = -1; break; }
      size_t w = 0;
      for (int64_t i = 0; i < recs.n; ++i) {
        if (!keep[(size_t)i]) continue;
        tid_.push_back(recs.tid[i]); pos_.push_back(recs.pos[i]); hp_.push_back(recs.hp[i]);
        qname_.emplace_back(recs.names + recs.name_offs[i], (size_t)(recs.name_offs[i + 1] - recs.name_offs[i]));
        if (sfs_of[(size_t)i]) {
          cigar_.insert(cigar_.end(), f_cig + f_co[w], f_cig + f_co[w + 1]);
          for (const SFS& s : *sfs_of[(size_t)i]) { sfs_qs_.push_back(s.qs); sfs_len_.push_back(s.l); }
          payload_of_.push_back((int64_t)seq4_.size());
          seq4_.emplace_back(f_seq + f_so[w], f_seq + f_so[w + 1]);
          ++w;
        } else {
          cigar_.push_back((uint32_t)(recs.endpos[i] - recs.pos[i]) << 4);
          payload_of_.push_back(-1);
        }
        cigar_offs_.push_back((int64_t)cigar_.size());
        sfs_offs_.push_back((int64_t)sfs_qs_.size());
      }
    }
    if (rc == 1 && (src.failed() || svb_bamstream_pending_bytes(bs) != 0)) { error = "truncated or corrupt BAM " + cfg_.bam; rc = -1; }
    svb_bamstream_close(bs);
    if (rc == 1) scanned_on_device = true;
    return rc;
  }

  // pass 1 (clusterer.cpp:58-153): one sequential scan
  bool scan() {
    const int on_device = scan_device();
    if (on_device != 0) return on_device > 0;
    BamReader bam(cfg_.bam);
    if (!bam.ok()) { error = "cannot read BAM " + cfg_.bam; return false; }
    bam.want_alignment(true);
    ref_names_ = bam.ref_names();
    BamRecord r;
    int st;
    while ((st = bam.next(r)) == 1) {
      if (r.flag & 0x4 || r.flag & 0x800 || r.flag & 0x100) continue;   // :116-120
      if (r.mapq < cfg_.min_mapq) continue;                             // :121-122
      if (r.tid < 0 || (size_t)r.tid >= ref_names_.size()) continue;
      auto it = SFSs_->find(r.qname);                                   // :123-125
      tid_.push_back(r.tid); pos_.push_back(r.pos); hp_.push_back(r.has_hp ? (int32_t)r.hp : 0);
      qname_.push_back(r.qname);
      if (it != SFSs_->end()) {
        cigar_.insert(cigar_.end(), r.cigar.begin(), r.cigar.end());
        for (const SFS& s : it->second) { sfs_qs_.push_back(s.qs); sfs_len_.push_back(s.l); }
        payload_of_.push_back((int64_t)seq4_.size());
        seq4_.push_back(r.seq4);
      } else {
        // a read without SFSs only counts into coverage: its reference span is all that is needed of its alignment
        cigar_.push_back((uint32_t)(r.endpos() - r.pos) << 4);
        payload_of_.push_back(-1);
      }
      cigar_offs_.push_back((int64_t)cigar_.size());
      sfs_offs_.push_back((int64_t)sfs_qs_.size());
    }
    if (st < 0) { error = "truncated or corrupt BAM " + cfg_.bam; return false; }
    return true;
  }

  bool place_and_fill() {
    // chromosome_seqs in tid order, one buffer; a BAM chromosome the FASTA does not hold has length -1 (clusterer.cpp:162-163)
    std::string refcat;
    std::vector<int64_t> rstart(ref_names_.size(), 0), rlen(ref_names_.size(), -1);
    for (size_t t = 0; t < ref_names_.size(); ++t) {
      auto it = chroms_->find(ref_names_[t]);
      if (it == chroms_->end()) continue;
      bool seen = false;
      for (size_t u = 0; u < t && !seen; ++u) if (ref_names_[u] == ref_names_[t]) { rstart[t] = rstart[u]; rlen[t] = rlen[u]; seen = true; }
      if (seen) continue;
      rstart[t] = (int64_t)refcat.size(); rlen[t] = (int64_t)it->second.size();
      refcat += it->second;
    }
    // SFS::operator< compares chromosome NAMES (sfs.hpp:64-72): rank of every tid's name
    std::vector<int32_t> order(ref_names_.size()), rank(ref_names_.size());
    for (size_t t = 0; t < order.size(); ++t) order[t] = (int32_t)t;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return ref_names_[(size_t)a] < ref_names_[(size_t)b]; });
    for (size_t k = 0, r = 0; k < order.size(); ++k) {
      if (k && ref_names_[(size_t)order[k]] != ref_names_[(size_t)order[k - 1]]) r = k;
      rank[(size_t)order[k]] = (int32_t)r;
    }
    svb_alns_t A;
    A.n_aln = (int64_t)tid_.size(); A.tid = tid_.data(); A.pos = pos_.data(); A.hp = hp_.data();
    A.cigar_offs = cigar_offs_.data(); A.cigar = cigar_.data(); A.sfs_offs = sfs_offs_.data(); A.sfs_qs = sfs_qs_.data(); A.sfs_len = sfs_len_.data();
    svb_ref_t R;
    R.n_contigs = (int64_t)ref_names_.size(); R.seq = reinterpret_cast<const uint8_t*>(refcat.data()); R.start = rstart.data(); R.len = rlen.data();
    R.name_rank = rank.data(); R.fmt = SVB_SEQ_ASCII; R.mem = SVB_MEM_HOST;
    svb_clusters_t C;
    if (svb_cluster_batch(&A, &R, cfg_.threads, (int)cfg_.min_cluster_weight, (int)cfg_.flank, (int)cfg_.ksize, cfg_.clipped ? 1 : 0, cfg_.device, &C) != SVB_OK) {
      error = std::string("svb_cluster_batch: ") + svb_last_error();
      return false;
    }
    unplaced = (unsigned)C.unplaced; s_unplaced = (unsigned)C.s_unplaced; e_unplaced = (unsigned)C.e_unplaced; unknown = (unsigned)C.unknown;
    unextended = (unsigned)C.unextended; small_clusters = (unsigned)C.small_clusters; small_clusters_2 = (unsigned)C.small_clusters_2;
    n_extended = (size_t)C.n_extended; max_ext_len = C.max_ext_len; dist = C.dist;
    clusters.resize((size_t)C.n_clusters);
    for (int64_t c = 0; c < C.n_clusters; ++c) {
      Cluster& cl = clusters[(size_t)c];
      cl.chrom = ref_names_[(size_t)C.tid[c]];
      cl.placed = C.placed[c] != 0;
      cl.s = C.s[c]; cl.e = C.e[c];
      cl.cov0 = C.cov0[c]; cl.cov1 = C.cov1[c]; cl.cov2 = C.cov2[c]; cl.cov = cl.cov0 + cl.cov1 + cl.cov2;
      for (int64_t k = C.rvec_offs[c]; k < C.rvec_offs[c + 1]; ++k) cl.reads.emplace_back(C.rvec[k] & 1, C.rvec[k] >> 1);
      for (int64_t k = C.sub_offs[c]; k < C.sub_offs[c + 1]; ++k) {
        const int a = C.sub_aln[k], qs = C.sub_qs[k], qe = C.sub_qe[k];
        const std::vector<uint8_t>& s4 = seq4_[(size_t)payload_of_[(size_t)a]];
        std::string seq;
        if (qe >= qs) {
          seq.resize((size_t)(qe - qs + 1));
          for (int i = qs; i <= qe; ++i) { const uint8_t b = s4[(size_t)i >> 1]; seq[(size_t)(i - qs)] = nt16_char((i & 1) ? (b & 0xf) : (b >> 4)); }
        }
        cl.subreads.push_back(SubRead{qname_[(size_t)a], seq, C.sub_hp[k]});
      }
    }
    if (cfg_.clipped && C.clip) {
      // the reference's order: accepted reads dealt round robin to thread slots, per-slot vectors inserted at the front (:24)
      const size_t T = (size_t)std::max(1, cfg_.threads);
      std::vector<std::vector<Clip>> p_clips(T);
      size_t n = 0;
      for (size_t a = 0; a < tid_.size(); ++a) {
        if (sfs_offs_[a + 1] == sfs_offs_[a]) continue;
        const int32_t* q = C.clip + a * 4;
        const std::string& chrom = ref_names_[(size_t)tid_[a]];
        if (q[1] > 0) p_clips[n % T].push_back(Clip(qname_[a], chrom, (unsigned)q[0], (unsigned)q[1], true));
        if (q[3] > 0) p_clips[n % T].push_back(Clip(qname_[a], chrom, (unsigned)q[2], (unsigned)q[3], false));
        ++n;
      }
      for (size_t t = 0; t < T; ++t) clips.insert(clips.begin(), p_clips[t].begin(), p_clips[t].end());
    }
    svb_clusters_free(&C);
    return true;
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
