// `SVDSS call` compute core on the GPU: Caller::pcall (reference caller.cpp:311-406) --
// split_cluster -> run_poa -> ksw_extd2_sse -> CIGAR walk -> SV records -> VCF -- with run_poa and
// ksw2 batched over all sub-clusters through svb_poa_batch / svb_ksw_extd2_batch.
//
// The reference builds its clusters with Clusterer (clusterer.cpp: BAM + .sfs -> clusters), which
// is host glue outside this round's scope; this shell therefore takes the clusters
// from the file the reference itself writes with `--clusters` (clusterer.cpp:613-626):
//     chrom:s+1-e+1 <TAB> n { <TAB> name:seq }
// Sub-read haplotype tags are not part of that file, so every sub-read has htag 0 and
// split_cluster takes its "no alignment is tagged, use length" branch (caller.cpp:130-150).
// clean_dups / filter_sv_chains (rapidfuzz, caller.cpp:409-475) are not applied.
#pragma once
#include <algorithm>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/svdss_b200.h"
#include "io.hpp"

namespace svdss {

struct SubRead { std::string name, seq; int htag; size_t size() const { return seq.size(); } };  // clusterer.hpp:24-36

struct Cluster {  // clusterer.hpp:38-139 (fields kept)
  std::string chrom;
  int s = 0, e = 0, cov = 0, cov0 = 0, cov1 = 0, cov2 = 0;
  std::vector<SubRead> subreads;
  size_t size() const { return subreads.size(); }
  int get_len() const {  // integer mean, clusterer.hpp:103-111
    unsigned l = 0, n = 0;
    for (const auto& sr : subreads) { ++n; l += (unsigned)sr.size(); }
    return (int)(l / n);
  }
};

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

// caller.cpp:100-150, untagged branch: the two largest length groups
inline std::vector<Cluster> split_cluster(const Cluster& cluster, float min_ratio) {
  Cluster c0 = cluster;
  c0.cov1 = -1; c0.cov2 = -1;
  std::vector<Cluster> sub = split_cluster_by_len(c0, min_ratio), out;
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

struct CallConfig {
  std::string reference, clusters_in, poa_out;
  unsigned min_cluster_weight = 2, min_sv_length = 25;
  float min_ratio = 0.97f;
  int device = 0;
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
  // clusters (clusterer.cpp:613-626 format)
  std::vector<Cluster> clusters;
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
  log("info", "Calling SVs from " + std::to_string(clusters.size()) + " clusters..");
  // pcall: collect sub-clusters (caller.cpp:311-330)
  std::vector<Cluster> jobs;
  for (const Cluster& cl : clusters) {
    if (cl.size() < c.min_cluster_weight) continue;   // :316
    for (const Cluster& sc : split_cluster(cl, c.min_ratio)) jobs.push_back(sc);
  }
  const uint8_t* t26 = char26_table();
  // run_poa for all jobs (caller.cpp:257-308)
  std::vector<uint8_t> seqcat; std::vector<int64_t> soff(1, 0), coff(1, 0);
  for (const Cluster& j : jobs) {
    for (const SubRead& sr : j.subreads) { for (char ch : sr.seq) seqcat.push_back(t26[(uint8_t)ch]); soff.push_back((int64_t)seqcat.size()); }
    coff.push_back((int64_t)soff.size() - 1);
  }
  svb_poa_out_t poa;
  if (seqcat.empty()) seqcat.push_back(0);
  if (svb_poa_batch(seqcat.data(), soff.data(), coff.data(), (int64_t)jobs.size(), c.device, &poa) != SVB_OK) { log("critical", std::string("svb_poa_batch: ") + svb_last_error()); return 1; }
  std::vector<std::string> cons(jobs.size()), refs(jobs.size());
  std::vector<uint8_t> q, t; std::vector<int64_t> qo(1, 0), to(1, 0);
  for (size_t k = 0; k < jobs.size(); ++k) {
    for (int64_t i = poa.cons_offs[k]; i < poa.cons_offs[k + 1]; ++i) cons[k] += "ACGTN"[poa.cons[i]];   // :295-297
    refs[k] = seqs[jobs[k].chrom].substr((size_t)jobs[k].s, (size_t)(jobs[k].e - jobs[k].s + 1));         // :329
    for (char ch : cons[k]) q.push_back(t26[(uint8_t)ch]);
    for (char ch : refs[k]) t.push_back(t26[(uint8_t)ch]);
    qo.push_back((int64_t)q.size()); to.push_back((int64_t)t.size());
  }
  svb_poa_out_free(&poa);
  svb_ksw_out_t ez;
  if (q.empty()) q.push_back(0);
  if (t.empty()) t.push_back(0);
  if (svb_ksw_extd2_batch(q.data(), qo.data(), t.data(), to.data(), (int64_t)jobs.size(), 1, -9, -1, 16, 2, 41, 1, c.device, &ez) != SVB_OK) {  // :333-349
    log("critical", std::string("svb_ksw_extd2_batch: ") + svb_last_error());
    return 1;
  }
  std::vector<SV> svs;
  std::vector<std::string> sam;
  for (size_t k = 0; k < jobs.size(); ++k) {
    const Cluster& cl = jobs[k];
    const std::string& chromseq = seqs[cl.chrom];
    const int score = ez.score[k];
    std::string cigar_str;
    for (int64_t i = ez.cigar_offs[k]; i < ez.cigar_offs[k + 1]; ++i) cigar_str += std::to_string(ez.cigar[i] >> 4) + "MID"[ez.cigar[i] & 0xf];  // :352-355
    sam.push_back(cl.chrom + ":" + std::to_string(cl.s + 1) + "-" + std::to_string(cl.e + 1) + "\t0\t" + cl.chrom + "\t" +
                  std::to_string(cl.s + 1) + "\t60\t" + cigar_str + "\t*\t0\t0\t" + cons[k] + "\t*");        // caller.hpp:56-70
    std::vector<SV> _svs;
    unsigned rpos = (unsigned)cl.s, cpos = 0;
    int nv = 0;
    std::string reads;
    for (const SubRead& sr : cl.subreads) reads += sr.name + ",";
    if (!reads.empty()) reads.pop_back();
    for (int64_t i = ez.cigar_offs[k]; i < ez.cigar_offs[k + 1]; ++i) {  // :359-395
      const unsigned l = ez.cigar[i] >> 4;
      const char op = "MID"[ez.cigar[i] & 0xf];
      if (op == 'M') { rpos += l; cpos += l; }
      else if (op == 'I') {
        if (l >= c.min_sv_length) {
          SV sv("INS", cl.chrom, rpos, chromseq.substr(rpos - 1, 1), chromseq.substr(rpos - 1, 1) + cons[k].substr(cpos, l),
                (unsigned)cl.size(), (unsigned)cl.cov, nv, score, false, l, cigar_str);
          sv.reads = reads; _svs.push_back(sv); nv++;
        }
        cpos += l;
      } else {
        if (l >= c.min_sv_length) {
          SV sv("DEL", cl.chrom, rpos, chromseq.substr(rpos - 1, l + 1), chromseq.substr(rpos - 1, 1), (unsigned)cl.size(),
                (unsigned)cl.cov, nv, score, false, l, cigar_str);
          sv.reads = reads; _svs.push_back(sv); nv++;
        }
        rpos += l;
      }
    }
    for (SV& sv : _svs) {  // :396-401
      sv.ngaps = nv; sv.gt = "0/1"; sv.gtq = 100;
      sv.cov = cl.cov; sv.cov0 = cl.cov0; sv.cov1 = cl.cov1; sv.cov2 = cl.cov2;
      svs.push_back(sv);
    }
  }
  svb_ksw_out_free(&ez);
  std::stable_sort(svs.begin(), svs.end());   // caller.cpp:23,28 (clean_dups / filter_sv_chains not applied)
  log("info", "Writing " + std::to_string(svs.size()) + " SVs.");
  print_vcf_header(chroms, seqs);
  for (const SV& sv : svs) { std::string l = sv.vcf_line() + "\n"; fwrite(l.data(), 1, l.size(), stdout); }
  if (!c.poa_out.empty()) {  // write_sam, caller.cpp:65-75
    FILE* f = fopen(c.poa_out.c_str(), "w");
    if (!f) { log("critical", "cannot write " + c.poa_out); return 1; }
    fprintf(f, "@HD\tVN:1.4\n");
    for (const auto& ch : chroms) fprintf(f, "@SQ\tSN:%s\tLN:%zu\n", ch.c_str(), seqs[ch].size());
    for (const auto& a : sam) fprintf(f, "%s\n", a.c_str());
    fclose(f);
  }
  return 0;
}

}  // namespace svdss
