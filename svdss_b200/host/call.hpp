// `SVDSS call` on the GPU: Caller::run (reference caller.cpp:3-60) -- parse the .sfs file, Clusterer
// (clusterer.hpp), then Caller::pcall (caller.cpp:311-406): split_cluster -> run_poa ->
// ksw_extd2_sse -> CIGAR walk -> SV records, with run_poa and ksw2 batched over all sub-clusters
// through svb_poa_batch / svb_ksw_extd2_batch; then clean_dups, filter_sv_chains, VCF (+ SAM of
// the consensus alignments with --poa).
//
// Clusters come either from --bam + --sfs (the reference's own inputs) or from a file in the
// format the reference writes with `--clusters` (clusterer.cpp:613-626):
//     chrom:s+1-e+1 <TAB> n { <TAB> name:seq }
// That file carries no haplotype tags or coverage vectors, so with --clusters-in every sub-read
// has htag 0, cov = cov0 = n and RVEC is empty.
// --clipped (caller.cpp:37-55) appends the Clipper's imprecise records (clipper.hpp) after the VCF body.
// Not carried over: the reference's reuse of the OpenMP loop variable `i` inside pcall's encode
// loops (caller.cpp:339-342), which is undefined behaviour.
#pragma once
#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/svdss_b200.h"
#include "clipper.hpp"
#include "clusterer.hpp"
#include "io.hpp"

namespace svdss {

struct SV {  // sv.hpp:12-62, sv.cpp:7-27
  std::string type, chrom, idx, refall, altall, gt = "./.", cigar, reads, rvec;
  int s = 0, e = 0, cov = 0, cov0 = 0, cov1 = 0, cov2 = 0, l = 0, ngaps = 0, score = 0, gtq = 0;
  unsigned w = 0;
  bool imprecise = false;
  SV(const std::string& type_, const std::string& chrom_, unsigned s_, const std::string& refall_,
     const std::string& altall_, unsigned w_, unsigned cov_, int ngaps_, int score_, bool imprecise_, unsigned l_,
     const std::string& cigar_)
      : type(type_), chrom(chrom_), refall(refall_), altall(altall_), cigar(cigar_), s((int)s_), cov((int)cov_),
        l((int)l_), ngaps(ngaps_), score(score_), w(w_), imprecise(imprecise_) {
    e = s + (int)refall.size() - 1;
    idx = type + "_" + chrom + ":" + std::to_string(s) + "-" + std::to_string(e) + "_" + std::to_string(l < 0 ? -l : l);
  }
  bool operator<(const SV& c) const { return chrom != c.chrom ? chrom < c.chrom : s < c.s; }  // sv.hpp:45-53
  std::string vcf_line() const {  // sv.cpp:53-80
    std::string o = chrom + "\t" + std::to_string(s) + "\t" + idx + "\t" + refall + "\t" + altall + "\t.\tPASS\t";
    o += "VARTYPE=SV;SVTYPE=" + type + ";SVLEN=" + std::to_string(type == "DEL" ? -l : l) + ";END=" + std::to_string(e) +
         ";WEIGHT=" + std::to_string(w) + ";COV=" + std::to_string(cov) + ";COV0=" + std::to_string(cov0) +
         ";COV1=" + std::to_string(cov1) + ";COV2=" + std::to_string(cov2) + ";AS=" + std::to_string(score) +
         ";NV=" + std::to_string(ngaps) + ";CIGAR=" + cigar + ";RVEC=" + rvec + ";READS=" + reads +
         (imprecise ? ";IMPRECISE\t" : "\t") + "GT:GQ\t" + gt + ":" + std::to_string(gtq);
    return o;
  }
};

inline const uint8_t* char26_table() {  // caller.hpp:25-37
  static uint8_t t[256];
  static bool init = false;
  if (!init) {
    memset(t, 4, sizeof(t));
    t[0] = 0; t[1] = 1; t[2] = 2; t[3] = 3;
    const char* s = "AaCcGgTtUu";
    const uint8_t v[] = {0, 0, 1, 1, 2, 2, 3, 3, 3, 3};
    for (int i = 0; s[i]; ++i) t[(uint8_t)s[i]] = v[i];
    init = true;
  }
  return t;
}

// caller.cpp:78-97
inline std::vector<Cluster> split_cluster_by_len(const Cluster& cluster, float min_ratio) {
  std::vector<Cluster> sub;
  for (const SubRead& sr : cluster.subreads) {
    size_t i;
    for (i = 0; i < sub.size(); i++) {
      float cl = (float)sub[i].get_len(), sl = (float)sr.size();
      if (std::min(cl, sl) / std::max(cl, sl) >= min_ratio) break;
    }
    if (i == sub.size()) {
      Cluster c;
      c.chrom = cluster.chrom; c.s = cluster.s; c.e = cluster.e;
      c.cov = cluster.cov; c.cov0 = cluster.cov0; c.cov1 = cluster.cov1; c.cov2 = cluster.cov2;
      sub.push_back(c);
    }
    sub[i].subreads.push_back(sr);
  }
  return sub;
}

static inline Cluster shell_of(const Cluster& c, int cov, int cov0, int cov1, int cov2) {  // Cluster(chrom,s,e,cov..)
  Cluster o;
  o.chrom = c.chrom; o.s = c.s; o.e = c.e; o.cov = cov; o.cov0 = cov0; o.cov1 = cov1; o.cov2 = cov2;
  return o;
}

static inline int largest(const std::vector<Cluster>& v) {
  unsigned v_max = 0; int i_max = -1;
  for (unsigned i = 0; i < v.size(); ++i) if (v[i].size() > v_max) { v_max = (unsigned)v[i].size(); i_max = (int)i; }
  return i_max;
}

// caller.cpp:100-255: by haplotype tag first (unless --noht), by length inside each haplotype;
// at most two sub-clusters come back
inline std::vector<Cluster> split_cluster(const Cluster& cluster, float min_ratio, bool useht) {
  Cluster c0 = shell_of(cluster, cluster.cov, cluster.cov0, cluster.cov1, cluster.cov2), c1 = c0, c2 = c0;
  for (const SubRead& sr : cluster.subreads) {
    if (useht && sr.htag == 1) c1.subreads.push_back(sr);
    else if (useht && sr.htag == 2) c2.subreads.push_back(sr);
    else c0.subreads.push_back(sr);
  }
  c0.cov1 = -1; c0.cov2 = -1; c1.cov0 = -1; c1.cov2 = -1; c2.cov0 = -1; c2.cov1 = -1;
  std::vector<Cluster> out;
  if (c1.size() == 0 && c2.size() == 0) {   // no alignment is tagged, use length: the two largest groups
    std::vector<Cluster> sub = split_cluster_by_len(c0, min_ratio);
    int i1 = -1, i2 = -1;
    unsigned v1 = 0, v2 = 0;
    for (unsigned i = 0; i < sub.size(); ++i) {
      if (sub[i].size() > v1) { v2 = v1; i2 = i1; v1 = (unsigned)sub[i].size(); i1 = (int)i; }
      else if (sub[i].size() > v2) { v2 = (unsigned)sub[i].size(); i2 = (int)i; }
    }
    if (i1 != -1) out.push_back(sub[i1]);
    if (i2 != -1) out.push_back(sub[i2]);
    return out;
  }
  const int both = (c1.size() > 0 ? 1 : 0) + (c2.size() > 0 ? 2 : 0);
  std::vector<Cluster> sub1 = split_cluster_by_len(c1, min_ratio), sub2 = split_cluster_by_len(c2, min_ratio);
  Cluster fresh = shell_of(cluster, cluster.cov, cluster.cov0, -1, -1);
  for (const SubRead& sr : c0.subreads) {
    const float sl = (float)sr.size();
    // best_ratio_* are declared int in the reference (caller.cpp:162,172): the ratio truncates
    // to 0 (or 1 for equal lengths) when stored, so `r > best_ratio` compares against that
    int best_1 = -1, best_ratio_1 = -1, best_2 = -1, best_ratio_2 = -1;
    for (unsigned i = 0; i < sub1.size(); i++) {
      const float cl = (float)sub1[i].get_len(), r = std::min(cl, sl) / std::max(cl, sl);
      if (r >= min_ratio && r > (float)best_ratio_1) { best_1 = (int)i; best_ratio_1 = (int)r; }
    }
    for (unsigned i = 0; i < sub2.size(); i++) {
      const float cl = (float)sub2[i].get_len(), r = std::min(cl, sl) / std::max(cl, sl);
      if (r >= min_ratio && r > (float)best_ratio_2) { best_2 = (int)i; best_ratio_2 = (int)r; }
    }
    if (both == 1) {
      if (best_1 == -1) fresh.subreads.push_back(sr);
      else { sub1[best_1].subreads.push_back(sr); ++sub1[best_1].cov1; --fresh.cov0; }
    } else if (both == 2) {
      if (best_2 == -1) fresh.subreads.push_back(sr);
      else { sub2[best_2].subreads.push_back(sr); ++sub2[best_2].cov2; --fresh.cov0; }
    } else {
      if (best_1 != -1 && best_ratio_1 > best_ratio_2) { sub1[best_1].subreads.push_back(sr); ++sub1[best_1].cov1; --fresh.cov0; }
      else if (best_2 != -1 && best_ratio_2 > best_ratio_1) { sub2[best_2].subreads.push_back(sr); ++sub2[best_2].cov2; --fresh.cov0; }
    }
  }
  int i_max = largest(sub1);
  if (i_max != -1) out.push_back(sub1[i_max]);
  i_max = largest(sub2);
  if (i_max != -1) out.push_back(sub2[i_max]);
  if (both != 3) {
    std::vector<Cluster> subn = split_cluster_by_len(fresh, min_ratio);
    i_max = largest(subn);
    if (i_max != -1) {
      if (both == 1) subn[i_max].cov1 = -1; else subn[i_max].cov2 = -1;
      out.push_back(subn[i_max]);
    }
  }
  return out;
}

// rapidfuzz::fuzz::ratio (caller.cpp:455-458): normalised Indel similarity in percent,
// 100 * 2*LCS / (|a|+|b|); LCS by the bit-parallel recurrence S' = (S + (S & M)) | (S & ~M)
inline double fuzz_ratio(const std::string& a, const std::string& b) {
  const size_t n = a.size(), m = b.size();
  if (n + m == 0) return 100.0;
  if (n == 0 || m == 0) return 0.0;
  const size_t W = (n + 63) / 64;
  std::vector<uint64_t> pm(256 * W, 0), S(W, ~0ULL);
  for (size_t i = 0; i < n; ++i) pm[(size_t)(uint8_t)a[i] * W + (i >> 6)] |= 1ULL << (i & 63);
  for (size_t j = 0; j < m; ++j) {
    const uint64_t* M = &pm[(size_t)(uint8_t)b[j] * W];
    unsigned carry = 0;
    for (size_t w = 0; w < W; ++w) {
      const uint64_t u = S[w] & M[w];
      const uint64_t t = S[w] + u;
      const uint64_t sum = t + carry;
      carry = (t < S[w]) || (sum < t);
      S[w] = sum | (S[w] & ~M[w]);
    }
  }
  size_t lcs = 0;
  for (size_t i = 0; i < n; ++i) lcs += !((S[i >> 6] >> (i & 63)) & 1);
  return 100.0 * 2.0 * (double)lcs / (double)(n + m);
}

// caller.hpp:39-71: a sub-cluster's POA consensus placed on the reference; sam_line() is its operator<<
struct Consensus {
  std::string seq, chrom, cigar;
  int s = 0, e = 0;
  Consensus(std::string seq_, std::string cigar_, std::string chrom_, int s_, int e_)
      : seq(std::move(seq_)), chrom(std::move(chrom_)), cigar(std::move(cigar_)), s(s_), e(e_) {}
  std::string sam_line() const {
    return chrom + ":" + std::to_string(s + 1) + "-" + std::to_string(e + 1) + "\t0\t" + chrom + "\t" + std::to_string(s + 1) + "\t60\t" + cigar +
           "\t*\t0\t0\t" + seq + "\t*";
  }
};

struct CallConfig {
  std::string reference, clusters_in, poa_out, bam, sfs, clusters_out, clips_out;
  unsigned min_cluster_weight = 2, min_sv_length = 25, min_mapq = 20;
  float min_ratio = 0.97f;
  int device = 0, threads = 4, batch_size = 10000;
  bool useht = true;
  bool clipped = false;        // --clipped: imprecise SVs from clipped SFSs (experimental in the reference)
  bool cluster_only = false;   // stop after the Clusterer (writes --clusters); no GPU needed
};

inline void print_vcf_header(const std::vector<std::string>& chroms, const std::unordered_map<std::string, std::string>& seqs) {
  // caller.cpp:477-550
  std::string h = "##fileformat=VCFv4.2\n##reference=ftp://ftp.1000genomes.ebi.ac.uk/vol1/ftp/data_collections/HGSVC2/technical/reference/20200513_hg38_NoALT/hg38.no_alt.fa.gz\n";
  for (const auto& c : chroms) h += "##contig=<ID=" + c + ",length=" + std::to_string(seqs.at(c).size()) + ">\n";
  h += "##FILTER=<ID=PASS,Description=\"All filters passed\">\n"
       "##INFO=<ID=VARTYPE,Number=A,Type=String,Description=\"Variant class\">\n"
       "##INFO=<ID=SVTYPE,Number=1,Type=String,Description=\"Variant type\">\n"
       "##INFO=<ID=SVLEN,Number=1,Type=Integer,Description=\"Difference in length between REF and ALT alleles\">\n"
       "##INFO=<ID=END,Number=1,Type=Integer,Description=\"End position of the variant described in this record\">\n"
       "##INFO=<ID=WEIGHT,Number=1,Type=Integer,Description=\"Number of alignments supporting this record\">\n"
       "##INFO=<ID=COV,Number=1,Type=Integer,Description=\"Total number of alignments covering this locus\">\n"
       "##INFO=<ID=COV0,Number=1,Type=Integer,Description=\"Total number of alignments covering this locus (no HP)\">\n"
       "##INFO=<ID=COV1,Number=1,Type=Integer,Description=\"Total number of alignments covering this locus (HP=1)\">\n"
       "##INFO=<ID=COV2,Number=1,Type=Integer,Description=\"Total number of alignments covering this locus (HP=2)\">\n"
       "##INFO=<ID=AS,Number=1,Type=Integer,Description=\"Alignment score\">\n"
       "##INFO=<ID=NV,Number=1,Type=Integer,Description=\"Number of variations on same consensus\">\n"
       "##INFO=<ID=IMPRECISE,Number=0,Type=Flag,Description=\"Imprecise structural variation\">\n"
       "##INFO=<ID=CIGAR,Number=A,Type=String,Description=\"CIGAR of consensus\">\n"
       "##INFO=<ID=READS,Number=.,Type=String,Description=\"Reads identifiers supporting the call\">\n"
       "##INFO=<ID=RVEC,Number=.,Type=String,Description=\"Reads vector used by genotyper\">\n"
       "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">\n"
       "##FORMAT=<ID=GQ,Number=1,Type=Integer,Description=\"Genotype quality\">\n"
       "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\tDEFAULT\n";
  fwrite(h.data(), 1, h.size(), stdout);
}

// Clipper::call + the collection loop of caller.cpp:45-52: per-thread vectors are inserted at the front
inline std::vector<SV> clipped_calls(const std::vector<Clip>& clips, const std::vector<std::string>& chroms,
                                     const std::unordered_map<std::string, std::string>& seqs, int threads,
                                     const std::vector<std::pair<int, int>>& sv_regions, Clipper* keep = nullptr) {
  Clipper clipper(clips, &chroms, &seqs);
  clipper.call(threads, sv_regions);
  std::vector<SV> out;
  for (const auto& slot : clipper.p_calls) {
    std::vector<SV> v;
    for (const Clipper::Call& k : slot) v.push_back(SV(k.type, k.chrom, k.s, k.refbase, "<" + k.type + ">", k.w, 0, 0, 0, true, k.l, "."));
    out.insert(out.begin(), v.begin(), v.end());
  }
  if (keep) *keep = clipper;
  return out;
}

// returns process exit code; `log` prints to stderr
inline int run_call(const CallConfig& c, void (*log)(const char*, const std::string&)) {
  // load_chromosomes (chromosomes.cpp:10-27): upper-cased
  std::vector<std::string> chroms;
  std::unordered_map<std::string, std::string> seqs;
  {
    FastxReader fx(c.reference);
    if (!fx.ok()) { log("critical", "cannot open reference " + c.reference); return 1; }
    FastxRecord r;
    while (fx.next(r)) {
      for (auto& ch : r.seq) ch = (char)toupper((unsigned char)ch);
      chroms.push_back(r.name);
      seqs[r.name] = r.seq;
    }
  }
  std::vector<Cluster> clusters;
  std::vector<Clip> clips;
  if (c.clusters_in.empty()) {
    // Caller::run, caller.cpp:9-13
    std::unordered_map<std::string, std::vector<SFS>> sfss;
    size_t total = 0;
    log("info", "Loading SFSs from " + c.sfs + "..");
    if (!parse_sfsfile(c.sfs, sfss, total)) { log("critical", "cannot open " + c.sfs); return 1; }
    log("info", "Loaded " + std::to_string(total) + " SFSs from " + std::to_string(sfss.size()) + " reads.");
    ClusterConfig cc;
    cc.bam = c.bam; cc.clusters_out = c.clusters_out; cc.threads = c.threads; cc.batch_size = c.batch_size;
    cc.min_mapq = c.min_mapq; cc.min_cluster_weight = c.min_cluster_weight;
    cc.clipped = c.clipped; cc.clips_out = c.clips_out; cc.device = c.device;
    Clusterer C(cc, &sfss, &seqs);
    log("info", "Placing SFSs on reference genome");
    if (!C.run()) { log("critical", C.error.empty() ? "cannot write " + c.clusters_out : C.error); return 1; }
    if (C.scanned_on_device) log("info", "BAM records decoded on GPU " + std::to_string(c.device));
    log("info", std::to_string(C.unplaced) + "/" + std::to_string(C.s_unplaced) + "/" + std::to_string(C.e_unplaced) +
                    " unplaced SFSs. " + std::to_string(C.unknown) + " erroneus SFSs. " + std::to_string(C.clips.size()) + " clipped SFSs.");
    log("info", "Clustered " + std::to_string(C.n_extended) + " SFSs. Maximum extended SFS length: " + std::to_string(C.max_ext_len) +
                    "bp. Using separation distance: " + std::to_string(C.dist) + "bp.");
    log("info", "Filtered " + std::to_string(C.unextended) + " SFSs. Filtered " + std::to_string(C.small_clusters) +
                    " clusters. Filtered " + std::to_string(C.small_clusters_2) + " global clusters.");
    clusters.swap(C.clusters);
    clips.swap(C.clips);
    if (c.cluster_only) return 0;
  } else {
    // clusters (clusterer.cpp:613-626 format)
    {
      GzSource src(c.clusters_in);
      if (!src.ok()) { log("critical", "cannot open clusters file " + c.clusters_in); return 1; }
      std::string line;
      while (src.getline(line)) {
        if (line.empty()) continue;
        std::vector<std::string> tok;
        size_t b = 0;
        while (true) { size_t t = line.find('\t', b); tok.push_back(line.substr(b, t == std::string::npos ? t : t - b)); if (t == std::string::npos) break; b = t + 1; }
        size_t colon = tok[0].rfind(':'), dash = tok[0].rfind('-');
        if (tok.size() < 2 || colon == std::string::npos || dash == std::string::npos || dash < colon) { log("critical", "malformed cluster line"); return 1; }
        Cluster cl;
        cl.chrom = tok[0].substr(0, colon);
        cl.s = atoi(tok[0].substr(colon + 1, dash - colon - 1).c_str()) - 1;
        cl.e = atoi(tok[0].substr(dash + 1).c_str()) - 1;
        for (size_t i = 2; i < tok.size(); ++i) {
          size_t k = tok[i].rfind(':');
          if (k == std::string::npos) { log("critical", "malformed sub-read in cluster line"); return 1; }
          cl.subreads.push_back(SubRead{tok[i].substr(0, k), tok[i].substr(k + 1), 0});
        }
        cl.cov = cl.cov0 = (int)cl.size(); cl.cov1 = cl.cov2 = 0;
        if (!seqs.count(cl.chrom) || cl.s < 1 || cl.e < cl.s || (size_t)cl.e >= seqs[cl.chrom].size()) { log("critical", "cluster " + tok[0] + " is outside the reference"); return 1; }
        clusters.push_back(cl);
      }
    }
  }
  log("info", "Calling SVs from " + std::to_string(clusters.size()) + " clusters..");
  // Caller::pcall (caller.cpp:311-406) through svb_call_batch: the clusters as arrays, the sub-reads and the chromosomes as
  // ASCII in one buffer each; split_cluster, POA, ksw2 and the CIGAR walk happen behind the call
  std::unordered_map<std::string, int32_t> tid_of;
  std::string refcat;
  std::vector<int64_t> rstart, rlen;
  for (size_t t = 0; t < chroms.size(); ++t) {
    if (!tid_of.emplace(chroms[t], (int32_t)t).second) { rstart.push_back(0); rlen.push_back(-1); continue; }   // a repeated name: the first one wins (std::map semantics)
    rstart.push_back((int64_t)refcat.size()); rlen.push_back((int64_t)seqs[chroms[t]].size());
    refcat += seqs[chroms[t]];
  }
  std::vector<int32_t> c_tid, c_s, c_e, c_cov0, c_cov1, c_cov2, sub_aln, sub_qs, sub_qe, sub_hp;
  std::vector<uint8_t> c_placed;
  std::vector<int64_t> sub_offs(1, 0), seq_offs;
  std::string seqcat;
  for (size_t ci = 0; ci < clusters.size(); ++ci) {
    const Cluster& cl = clusters[ci];
    auto it = tid_of.find(cl.chrom);
    const bool inside = it != tid_of.end() && !(cl.s < 1 || cl.e < cl.s || (size_t)cl.e >= seqs[cl.chrom].size());
    if (cl.size() >= c.min_cluster_weight && !inside)   // the reference would read outside the chromosome
      log("warning", "cluster " + cl.chrom + ":" + std::to_string(cl.s + 1) + "-" + std::to_string(cl.e + 1) + " is outside the reference, skipped");
    c_tid.push_back(it == tid_of.end() ? -1 : it->second); c_s.push_back(cl.s); c_e.push_back(cl.e);
    c_cov0.push_back(cl.cov0); c_cov1.push_back(cl.cov1); c_cov2.push_back(cl.cov2);
    c_placed.push_back(1);
    for (const SubRead& sr : cl.subreads) {
      sub_aln.push_back((int32_t)seq_offs.size()); sub_qs.push_back(0); sub_qe.push_back((int32_t)sr.seq.size() - 1); sub_hp.push_back(sr.htag);
      seq_offs.push_back((int64_t)seqcat.size());
      seqcat += sr.seq;
    }
    sub_offs.push_back((int64_t)sub_aln.size());
  }
  if (seq_offs.empty()) seq_offs.push_back(0);
  if (sub_aln.empty()) { sub_aln.push_back(0); sub_qs.push_back(0); sub_qe.push_back(-1); sub_hp.push_back(0); }
  svb_clusters_t CL;
  memset(&CL, 0, sizeof(CL));
  CL.n_clusters = (int64_t)clusters.size();
  CL.tid = c_tid.data(); CL.s = c_s.data(); CL.e = c_e.data(); CL.cov0 = c_cov0.data(); CL.cov1 = c_cov1.data(); CL.cov2 = c_cov2.data();
  CL.placed = c_placed.data(); CL.sub_offs = sub_offs.data(); CL.sub_aln = sub_aln.data(); CL.sub_qs = sub_qs.data(); CL.sub_qe = sub_qe.data();
  CL.sub_hp = sub_hp.data();
  svb_seqs_t RD;
  RD.n = (int64_t)seq_offs.size(); RD.seq = reinterpret_cast<const uint8_t*>(seqcat.data()); RD.offs = seq_offs.data(); RD.fmt = SVB_SEQ_ASCII; RD.mem = SVB_MEM_HOST;
  svb_ref_t RF;
  RF.n_contigs = (int64_t)chroms.size(); RF.seq = reinterpret_cast<const uint8_t*>(refcat.data()); RF.start = rstart.data(); RF.len = rlen.data();
  RF.name_rank = nullptr; RF.fmt = SVB_SEQ_ASCII; RF.mem = SVB_MEM_HOST;
  svb_calls_t K;
  if (svb_call_batch(&CL, &RD, &RF, (int)c.min_cluster_weight, (int)c.min_sv_length, c.min_ratio, c.useht ? 1 : 0, c.device, &K) != SVB_OK) {
    log("critical", std::string("svb_call_batch: ") + svb_last_error());
    return 1;
  }
  const size_t T = (size_t)std::max(1, c.threads);
  std::vector<std::vector<SV>> p_svs(T);
  std::vector<std::vector<Consensus>> p_sam(T);          // _p_alignments, caller.cpp:312-314
  int64_t sv_k = 0;
  for (int64_t k = 0; k < K.n_jobs; ++k) {
    const size_t parent = (size_t)K.job_cluster[k];
    std::vector<SV>& svs = p_svs[parent % T];            // schedule(static, 1), caller.cpp:312-314
    std::vector<Consensus>& sam = p_sam[parent % T];
    std::string rvec;                                        // SV::set_rvec, sv.cpp:42-46
    for (const auto& r : clusters[parent].reads) rvec += std::to_string(r.first) + ":" + std::to_string(r.second) + "-";
    if (!rvec.empty()) rvec.pop_back();
    const Cluster& cl = clusters[parent];
    const std::string& chromseq = seqs[cl.chrom];
    const int score = K.score[k];
    std::string cons, cigar_str, reads;
    for (int64_t i = K.cons_offs[k]; i < K.cons_offs[k + 1]; ++i) cons += "ACGTN"[K.cons[i]];   // :295-297
    for (int64_t i = K.cigar_offs[k]; i < K.cigar_offs[k + 1]; ++i) cigar_str += std::to_string(K.cigar[i] >> 4) + "MID"[K.cigar[i] & 0xf];  // :352-355
    sam.emplace_back(cons, cigar_str, cl.chrom, cl.s, cl.e);                                              // caller.cpp:357
    const int64_t n_sub = K.job_sub_offs[k + 1] - K.job_sub_offs[k];
    for (int64_t i = K.job_sub_offs[k]; i < K.job_sub_offs[k + 1]; ++i) reads += cl.subreads[(size_t)(K.job_sub[i] - sub_offs[parent])].name + ",";
    if (!reads.empty()) reads.pop_back();
    const int cov = K.job_cov[k * 4], cov0 = K.job_cov[k * 4 + 1], cov1 = K.job_cov[k * 4 + 2], cov2 = K.job_cov[k * 4 + 3];
    for (; sv_k < K.n_svs && K.sv_job[sv_k] == k; ++sv_k) {  // :359-401
      const unsigned rpos = (unsigned)K.sv_pos[sv_k], l = (unsigned)K.sv_len[sv_k], cpos = (unsigned)K.sv_cpos[sv_k];
      SV sv = K.sv_type[sv_k] == 0
                  ? SV("INS", cl.chrom, rpos, chromseq.substr(rpos - 1, 1), chromseq.substr(rpos - 1, 1) + cons.substr(cpos, l), (unsigned)n_sub, (unsigned)cov, 0, score, false, l, cigar_str)
                  : SV("DEL", cl.chrom, rpos, chromseq.substr(rpos - 1, l + 1), chromseq.substr(rpos - 1, 1), (unsigned)n_sub, (unsigned)cov, 0, score, false, l, cigar_str);
      sv.reads = reads;
      sv.ngaps = K.job_nv[k]; sv.gt = "0/1"; sv.gtq = 100;
      sv.cov = cov; sv.cov0 = cov0; sv.cov1 = cov1; sv.cov2 = cov2;
      sv.rvec = rvec;
      svs.push_back(sv);
    }
  }
  svb_calls_free(&K);
  // caller.cpp:17-29: per-thread vectors are inserted at the front, then sort / clean_dups /
  // filter_sv_chains / sort (stable here; the reference's std::sort leaves ties unspecified)
  std::vector<SV> svs;
  std::vector<Consensus> sam;
  for (size_t slot = 0; slot < T; ++slot) {
    svs.insert(svs.begin(), p_svs[slot].begin(), p_svs[slot].end());
    sam.insert(sam.begin(), p_sam[slot].begin(), p_sam[slot].end());
  }
  std::stable_sort(svs.begin(), svs.end());
  {  // clean_dups, caller.cpp:409-427
    std::vector<SV> keep;
    std::string last_chrom, last_refall, last_altall;
    int last_pos = -1;
    for (const SV& sv : svs) {
      if (last_chrom != sv.chrom || last_pos != sv.s || last_refall != sv.refall || last_altall != sv.altall) keep.push_back(sv);
      last_chrom = sv.chrom; last_pos = sv.s; last_refall = sv.refall; last_altall = sv.altall;
    }
    svs.swap(keep);
  }
  log("info", std::to_string(svs.size()) + " SVs before chain filtering.");
  if (svs.size() >= 2) {  // filter_sv_chains, caller.cpp:430-475 (`prev` there aliases svs[0] and is assigned through)
    std::vector<SV> keep;
    SV prev = svs[0];
    bool reset = false;
    for (size_t i = 1; i < svs.size(); i++) {
      if (reset) { reset = false; prev = svs[i]; continue; }
      const SV& sv = svs[i];
      if (sv.chrom == prev.chrom && sv.s - prev.e < 2 * sv.l && prev.type == sv.type) {
        const double w_r = std::min((double)sv.w, (double)prev.w) / std::max((double)sv.w, (double)prev.w);
        const double l_r = std::min((double)sv.l, (double)prev.l) / std::max((double)sv.l, (double)prev.l);
        const int d = sv.s - prev.s;
        if (d < 100 && w_r >= 0.9 && l_r >= c.min_ratio) {
          const double sim = sv.type == "DEL" ? fuzz_ratio(sv.refall, prev.refall) : fuzz_ratio(sv.altall, prev.altall);
          if (sim > 70) {
            keep.push_back(sv.w > prev.w ? sv : prev);
            reset = true;
            continue;
          }
        }
      }
      keep.push_back(prev);
      prev = sv;
    }
    keep.push_back(prev);
    svs.swap(keep);
    std::stable_sort(svs.begin(), svs.end());
  }
  log("info", "Writing " + std::to_string(svs.size()) + " SVs.");
  print_vcf_header(chroms, seqs);
  for (const SV& sv : svs) { std::string l = sv.vcf_line() + "\n"; fwrite(l.data(), 1, l.size(), stdout); }
  if (!c.poa_out.empty()) {  // write_sam, caller.cpp:65-75
    FILE* f = fopen(c.poa_out.c_str(), "w");
    if (!f) { log("critical", "cannot write " + c.poa_out); return 1; }
    fprintf(f, "@HD\tVN:1.4\n");
    for (const auto& ch : chroms) fprintf(f, "@SQ\tSN:%s\tLN:%zu\n", ch.c_str(), seqs[ch].size());
    for (const Consensus& a : sam) fprintf(f, "%s\n", a.sam_line().c_str());
    fclose(f);
  }
  if (c.clipped) {  // caller.cpp:37-55
    log("warning", "Calling imprecise SVs from clipped alignments is experimental");
    if (!c.clusters_in.empty()) log("warning", "--clusters-in carries no alignments, hence no clips");
    std::vector<std::pair<int, int>> regions;
    for (const SV& sv : svs) regions.emplace_back(sv.s - 1000, sv.e + 1000);
    const std::vector<SV> clipped_svs = clipped_calls(clips, chroms, seqs, c.threads, regions);
    log("info", "Predicted " + std::to_string(clipped_svs.size()) + " SVs from clipped alignments");
    for (const SV& sv : clipped_svs) { std::string l = sv.vcf_line() + "\n"; fwrite(l.data(), 1, l.size(), stdout); }
  }
  return 0;
}

}  // namespace svdss
