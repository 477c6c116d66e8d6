// SVDSS shell for the B200 hot path: `SVDSS index | search` with the reference's flag surface
// (reference main.cpp:34-81, config.cpp:26-107, config.hpp:68-103), C++14, calling the CUDA
// library only through the C ABI of include/svdss_b200.h.
//
//   SVDSS index  [-t N] [-d] [-o OUT] FASTA          (main_build flags the pipeline uses, run_svdss:142)
//   SVDSS search --index IDX (--bam BAM | --fastx FQ) [--threads N] [--bsize N] [--noputative]
//                [--noassemble] [--omax N] [--verbose]        > specifics.sfs
//
// stdout = payload, stderr = log, EXIT_FAILURE on bad arguments like the reference.
// Output order is the reference's (ping_pong.cpp:213-236,329-361): logical batches of --bsize
// accepted reads, reads dealt round-robin to --threads slots, each slot printed in qname order.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <iostream>
#include <iterator>
#include <fstream>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/svdss_b200.h"
#include "io.hpp"
#include "call.hpp"
#include "smoother.hpp"
#include "rld.hpp"
#include <unistd.h>
static long getpid_portable() { return (long)getpid(); }

using namespace std;
using namespace svdss;

static const char* VERSION = "v2.1.1-b200";
static const char* MAIN_USAGE =
    "Usage: SVDSS <index|smooth|search|call> --help\n"
    "  index   build the FMD index of a reference on the GPU\n"
    "  smooth  remove everything but putative SVs from the alignments of a BAM\n"
    "  search  extract sample-specific strings (SFS) from a BAM/FASTX\n"
    "  call    POA consensus + ksw2 realignment + SV extraction from SFS clusters";
static const char* INDEX_USAGE =
    "Usage: SVDSS index [-t threads] [-d] [-o index] [--fmd <out.fmd>] <reference.fa[.gz]> [...]\n"
    "       SVDSS index --from-fmd <ropebwt3.fmd> [-o index]          (same as -i <ropebwt3.fmd>)\n"
    "  -o FILE     output [stdout]\n"
    "  -L          one sequence per line\n"
    "  --fmd       also dump the BWT in ropebwt3's FMD format (what the reference's `index -d` writes)\n"
    "  --from-fmd  convert an index built by the reference (ropebwt3 FMD) instead of reading a FASTA\n"
    "  ropebwt3 build's tuning flags (-m -l -n -p NUM, -2 -s -r -d) are accepted and have no effect on\n"
    "  search results; -F / -R (one strand only), -b and -T (other output formats) are refused.";
static const char* CALL_USAGE =
    "Usage: SVDSS call --reference <fa> (--bam <bam> --sfs <sfs> | --clusters-in <clusters.txt>) [--threads 4]\n"
    "                  [--min-cluster-weight 2] [--min-sv-length 25] [--min-mapq 20] [-l 0.97] [--noht]\n"
    "                  [--poa <out.sam>] [--clusters <out.txt>] [--cluster-only] [--clipped [--clips <out.tsv>]]\n"
    "  Clusters are built on the host from --bam/--sfs (Clusterer) or read back from a `--clusters` file;\n"
    "  POA consensus + realignment (Caller::pcall) run on the GPU. --cluster-only stops after --clusters.";
static const char* SMOOTH_USAGE =
    "Usage: SVDSS smooth --reference <fa> --bam <bam> [--threads 4] [--min-mapq 20] [--accp 0.98] > smoothed.bam";
static const char* SEARCH_USAGE =
    "Usage: SVDSS search --index <index> (--bam <bam> | --fastx <fastx>) [--threads 4] [--bsize 10000]\n"
    "                    [--noputative] [--noassemble] [--verbose]";

struct Config {
  string index, bam, fastx, out, reference, sfs, clusters_in, clusters_out, poa, clips_out, clips_in, regions_in, fmd_out, fmd_in;
  int min_cluster_weight = 2, min_sv_length = 25, min_mapq = 20;
  float min_ratio = 0.97f, accp = 0.98f;
  bool noht = false, clipped = false, cluster_only = false, line_input = false;
  int threads = 4, bsize = 10000, omax = 100000, device = 0;
  bool gpu_inflate = false;
  bool assemble = true, putative = true, verbose = false, help = false, version = false;
  int overlap = -1;  // config.hpp:82: never settable from the command line
};

static double now_s() { return chrono::duration<double>(chrono::steady_clock::now().time_since_epoch()).count(); }
static void logmsg(const char* lvl, const string& m) { fprintf(stderr, "[svdss-b200] [%s] %s\n", lvl, m.c_str()); }

static bool parse_common(int argc, char** argv, Config& c, vector<string>& positional) {
  // `SVDSS --version` / `SVDSS --help`: no subcommand, options start at argv[1]
  for (int i = (argv[1][0] == '-') ? 1 : 2; i < argc; ++i) {
    string a = argv[i], attached;
    bool has_attached = false;   // cxxopts also takes `--option=value`
    if (a.size() > 2 && a[0] == '-' && a[1] == '-') {
      const size_t eq = a.find('=');
      if (eq != string::npos) { attached = a.substr(eq + 1); a.resize(eq); has_attached = true; }
    }
    auto val = [&](string& dst) {
      if (has_attached) { dst = attached; return true; }
      if (i + 1 >= argc) return false;
      dst = argv[++i];
      return true;
    };
    auto ival = [&](int& dst) { string s; if (!val(s)) return false; dst = atoi(s.c_str()); return true; };
    bool ok = true;
    if (a == "--index") ok = val(c.index);
    else if (a == "--bam") ok = val(c.bam);
    else if (a == "--fastx") ok = val(c.fastx);
    else if (a == "--threads" || a == "-t") ok = ival(c.threads);
    else if (a.size() > 2 && a.compare(0, 2, "-t") == 0 && isdigit((unsigned char)a[2])) c.threads = atoi(a.c_str() + 2);
    else if (a == "--bsize") ok = ival(c.bsize);
    else if (a == "--reference") ok = val(c.reference);
    else if (a == "--sfs") ok = val(c.sfs);
    else if (a == "--clusters-in") ok = val(c.clusters_in);
    else if (a == "--clusters") ok = val(c.clusters_out);
    else if (a == "--cluster-only") c.cluster_only = true;
    else if (a == "--fmd") ok = val(c.fmd_out);
    else if (a == "--from-fmd") ok = val(c.fmd_in);
    else if (a == "--clips") ok = val(c.clips_out);
    else if (a == "--clips-in") ok = val(c.clips_in);
    else if (a == "--regions-in") ok = val(c.regions_in);
    else if (a == "--poa") ok = val(c.poa);
    else if (a == "--min-cluster-weight") ok = ival(c.min_cluster_weight);
    else if (a == "--min-sv-length") { ok = ival(c.min_sv_length); c.min_sv_length = max(25, c.min_sv_length); }  // config.cpp:87
    else if (a == "--min-mapq") ok = ival(c.min_mapq);
    else if (a == "-l") { string v; ok = val(v); if (ok) c.min_ratio = (float)atof(v.c_str()); }
    else if (a == "--accp") { string v; ok = val(v); if (ok) c.accp = (float)atof(v.c_str()); }
    else if (a == "--noht") c.noht = true;
    else if (a == "--clipped") c.clipped = true;
    else if (a == "--omax") ok = ival(c.omax);
    else if (a == "--append") { string unused; ok = val(unused); }   // registered by the reference (config.cpp:35), read nowhere
    else if (a == "--binary") {}                                       // config.cpp:51,96: parsed, never used
    else if (a == "--device") ok = ival(c.device);
    else if (a == "--gpu-inflate") c.gpu_inflate = true;               // BGZF windows inflated on the device (io.hpp)
    else if (a == "-o") ok = val(c.out);
    else if (a == "-d") {}
    else if (a == "--noassemble") c.assemble = false;
    else if (a == "--noputative") c.putative = false;
    else if (a == "--verbose") c.verbose = true;
    else if (a == "--help" || a == "-h") c.help = true;
    else if (a == "--version") c.version = true;
    else if (!a.empty() && a[0] == '-') { logmsg("critical", "unknown option " + a); return false; }
    else positional.push_back(a);
    if (!ok) { logmsg("critical", "option " + a + " needs a value"); return false; }
  }
  if (c.threads < 1) c.threads = 1;
  if (const char* e = getenv("SVB_BGZF_GPU")) c.gpu_inflate = c.gpu_inflate || atoi(e) != 0;
  if (c.gpu_inflate) bgzf_gpu_device() = c.device;
  c.bsize = (c.bsize / c.threads) * c.threads;  // config.cpp:106
  if (c.bsize < c.threads) c.bsize = c.threads;
  return true;
}

// ropebwt3 FMD -> the forward strands it indexes (rld.hpp): decode, invert the BWT, drop one strand of every pair
static bool fmd_to_contigs(const string& path, vector<uint8_t>& cat, vector<int64_t>& offs) {
  const double t0 = now_s();
  RldFile rf;
  string err;
  vector<uint8_t> bwt;
  if (!Rld::read(path, rf, err) || !Rld::decode_bwt(rf, bwt, err)) { logmsg("critical", err); return false; }
  if (rf.asize != 6) { logmsg("critical", path + ": alphabet of " + to_string(rf.asize) + " symbols, expected 6 ($ACGTN)"); return false; }
  rf.words.clear(); rf.words.shrink_to_fit();
  const double t1 = now_s();
  vector<string> seqs;
  {
    BwtInverter inv(bwt.data(), bwt.size());
    if (!inv.sequences(seqs)) { logmsg("critical", path + ": the BWT does not invert to '$'-terminated sequences"); return false; }
  }
  bwt.clear(); bwt.shrink_to_fit();
  vector<size_t> keep;
  if (!forward_strands(seqs, keep)) {
    logmsg("critical", path + ": some sequences have no reverse complement in the index (built with ropebwt3 -R?); SVDSS searches both strands");
    return false;
  }
  cat.clear(); offs.assign(1, 0);
  for (size_t k : keep) {
    cat.insert(cat.end(), seqs[k].begin(), seqs[k].end());
    offs.push_back((int64_t)cat.size());
    string().swap(seqs[k]);
  }
  char tb[200];
  snprintf(tb, sizeof(tb), "ropebwt3 FMD index: %zu sequences (%zu after dropping reverse strands), %zu bp; decode %.2f s, BWT inversion %.2f s",
           seqs.size(), keep.size(), cat.size(), t1 - t0, now_s() - t1);
  logmsg("info", tb);
  return true;
}

// `SVDSS index` = ropebwt3's `build` command line (main.cpp:34-37 hands argv to main_build): getopt-style
// short options, values attached or separate, several input files.  Returns false on a usage error.
static bool parse_index(int argc, char** argv, Config& c, vector<string>& positional) {
  for (int i = 2; i < argc; ++i) {
    const string a = argv[i];
    if (a == "--fmd" || a == "--from-fmd" || a == "--device" || a == "--threads") {
      if (i + 1 >= argc) { logmsg("critical", "option " + a + " needs a value"); return false; }
      const string v = argv[++i];
      if (a == "--fmd") c.fmd_out = v; else if (a == "--from-fmd") c.fmd_in = v; else if (a == "--device") c.device = atoi(v.c_str()); else c.threads = atoi(v.c_str());
      continue;
    }
    if (a == "--help") { c.help = true; continue; }
    if (a == "--version") { c.version = true; continue; }
    if (a.size() < 2 || a[0] != '-' || a == "-") { positional.push_back(a); continue; }
    if (a[1] == '-') { logmsg("critical", "unknown option " + a); return false; }
    for (size_t k = 1; k < a.size(); ++k) {            // a cluster of short options, like getopt
      const char o = a[k];
      if (strchr("tomlnpiS", o)) {                       // options with a value: rest of the word, or the next word
        string v = a.substr(k + 1);
        if (v.empty()) { if (i + 1 >= argc) { logmsg("critical", string("option -") + o + " needs a value"); return false; } v = argv[++i]; }
        if (o == 't') c.threads = atoi(v.c_str());
        else if (o == 'o') c.out = v;
        else if (o == 'i') c.fmd_in = v;                 // ropebwt3: read an existing FMD/FMR index
        else if (o == 'S') { logmsg("critical", "-S (save after each input file) is not supported"); return false; }
        break;                                           // -m -l -n -p: tuning of ropebwt3's builder, nothing to tune here
      }
      if (o == 'd' || o == '2' || o == 's' || o == 'r') continue;   // FMD output is implied; BCR / RLO / RCLO only permute the BWT
      if (o == 'L') { c.line_input = true; continue; }
      if (o == 'h') { c.help = true; continue; }
      if (o == 'F' || o == 'R') { logmsg("critical", string("-") + o + ": SVDSS search needs both strands in the index"); return false; }
      if (o == 'b' || o == 'T') { logmsg("critical", string("-") + o + ": only this library's index format (and --fmd) can be written"); return false; }
      logmsg("critical", string("unknown option -") + o);
      return false;
    }
  }
  if (c.threads < 1) c.threads = 1;
  return true;
}

static int run_index(const Config& c, const vector<string>& pos) {
  vector<uint8_t> cat;
  vector<int64_t> offs(1, 0);
  if (!c.fmd_in.empty()) {
    if (!pos.empty()) { logmsg("critical", "adding sequences to an existing index is not supported"); return EXIT_FAILURE; }
    if (!Rld::is_rld(c.fmd_in)) { logmsg("critical", c.fmd_in + " is not a ropebwt3 FMD (RLD\\3) file"); return EXIT_FAILURE; }
    if (!fmd_to_contigs(c.fmd_in, cat, offs)) return EXIT_FAILURE;
  } else {
    if (pos.empty()) { cerr << INDEX_USAGE << endl; return EXIT_FAILURE; }
    const uint8_t* t6 = nt6_table();
    for (const string& path : pos) {                     // ropebwt3 build takes any number of input files
      if (c.line_input) {                                // -L: one sequence per line
        GzSource src(path);
        if (!src.ok()) { logmsg("critical", "cannot open " + path); return EXIT_FAILURE; }
        string line;
        while (src.getline(line)) {
          if (line.empty()) continue;
          for (char ch : line) cat.push_back(t6[(uint8_t)ch]);
          offs.push_back((int64_t)cat.size());
        }
        continue;
      }
      FastxReader fx(path);
      if (!fx.ok()) { logmsg("critical", "cannot open " + path); return EXIT_FAILURE; }
      FastxRecord r;
      while (fx.next(r)) {
        for (char ch : r.seq) cat.push_back(t6[(uint8_t)ch]);
        offs.push_back((int64_t)cat.size());
      }
    }
    if (offs.size() < 2) { logmsg("critical", "no sequences in " + pos[0]); return EXIT_FAILURE; }
  }
  logmsg("info", "indexing " + to_string(offs.size() - 1) + " sequences, " + to_string(cat.size()) + " bp (both strands) on GPU " + to_string(c.device));
  svb_index_t* idx = nullptr;
  if (svb_index_build(cat.data(), offs.data(), (int64_t)offs.size() - 1, SVB_MEM_HOST, c.device, 0, &idx) != SVB_OK) {
    logmsg("critical", string("svb_index_build: ") + svb_last_error());
    return EXIT_FAILURE;
  }
  if (!c.fmd_out.empty()) {   // the reference's own index format, for the reference's own `search`
    svb_index_info_t info;
    vector<uint8_t> bwt;
    string err;
    bool ok = svb_index_info(idx, &info) == SVB_OK;
    if (ok) { bwt.resize((size_t)info.n); ok = svb_index_get_bwt(idx, bwt.data()) == SVB_OK; }
    if (!ok) { logmsg("critical", string("svb_index_get_bwt: ") + svb_last_error()); svb_index_free(idx); return EXIT_FAILURE; }
    if (!Rld::write(c.fmd_out, bwt.data(), bwt.size(), err)) { logmsg("critical", err); svb_index_free(idx); return EXIT_FAILURE; }
    logmsg("info", "BWT written in ropebwt3 FMD format to " + c.fmd_out);
  }
  string out = c.out;
  const bool to_stdout = out.empty();  // ropebwt3 build writes to stdout without -o (README.md:113)
  if (to_stdout) out = "/tmp/svdss_b200_index." + to_string((long)getpid_portable());
  int rc = svb_index_save(idx, out.c_str());
  svb_index_free(idx);
  if (rc != SVB_OK) { logmsg("critical", string("svb_index_save: ") + svb_last_error()); return EXIT_FAILURE; }
  if (to_stdout) {
    FILE* f = fopen(out.c_str(), "rb");
    vector<char> buf(1 << 20);
    size_t n;
    while (f && (n = fread(buf.data(), 1, buf.size(), f)) > 0) fwrite(buf.data(), 1, n, stdout);
    if (f) fclose(f);
    remove(out.c_str());
  }
  return EXIT_SUCCESS;
}

static double g_gpu_s = 0;   // wall time inside svb_sfs_batch* (stage report of `search`)

struct PendingRead { string qname; int hp; int64_t lo, hi; bool search; int32_t l_qseq; };

// one GPU submission covering many logical batches; prints in the reference's order.
// packed == false: `cat` holds one nt6 byte per base (FASTX);  packed == true: `cat` holds the reads
// as the BAM stores them (4 bits per base) and the decode of ping_pong.cpp:90-94 runs on the GPU.
static bool flush(svb_index_t* idx, const Config& c, vector<PendingRead>& reads, vector<uint8_t>& cat,
                  uint64_t& total_sfs, bool packed) {
  if (reads.empty()) return true;
  // `cat` holds the sequences of the searched reads only, in read order (reads filtered by the XF rule
  // keep their slot in the output order but carry no bases, ping_pong.cpp:202-203)
  vector<int64_t> offs(1, 0);
  vector<int32_t> lq;
  vector<int64_t> slot_of(reads.size(), -1);
  int64_t n = 0;
  for (size_t i = 0; i < reads.size(); ++i)
    if (reads[i].search) { slot_of[i] = n++; offs.push_back(reads[i].hi); lq.push_back(reads[i].l_qseq); }
  const vector<uint8_t>& sub = cat;
  svb_sfs_out_t out;
  const double t_gpu = now_s();
  const int rc = packed ? svb_sfs_batch_bam4(idx, sub.data(), offs.data(), lq.data(), n, c.overlap, c.assemble ? 1 : 0, &out)
                        : svb_sfs_batch(idx, sub.data(), offs.data(), n, c.overlap, c.assemble ? 1 : 0, &out);
  g_gpu_s += now_s() - t_gpu;
  if (rc != SVB_OK) {
    logmsg("critical", string(packed ? "svb_sfs_batch_bam4: " : "svb_sfs_batch: ") + svb_last_error());
    return false;
  }
  string line;
  for (size_t b0 = 0; b0 < reads.size(); b0 += (size_t)c.bsize) {          // logical batch
    const size_t b1 = min(reads.size(), b0 + (size_t)c.bsize);
    for (int t = 0; t < c.threads; ++t) {                                   // thread slot
      map<string, vector<size_t>> slot;                                     // batch_type_t is a std::map
      for (size_t i = b0 + (size_t)t; i < b1; i += (size_t)c.threads)
        if (reads[i].search) slot[reads[i].qname].push_back(i);
      for (auto& kv : slot) {
        bool first = true;
        for (size_t i : kv.second) {
          const int64_t r = slot_of[i];
          for (int64_t k = out.offs[r]; k < out.offs[r + 1]; ++k) {         // ping_pong.cpp:227-228
            line = (first ? kv.first : string("*")) + "\t" + to_string(out.qs[k]) + "\t" + to_string(out.len[k]) +
                   "\t" + to_string(reads[i].hp) + "\t\n";
            fwrite(line.data(), 1, line.size(), stdout);
            first = false;
            ++total_sfs;
          }
        }
      }
    }
  }
  svb_sfs_out_free(&out);
  reads.clear();
  cat.clear();
  return true;
}

// ---- `search --bam --gpu-inflate`: the BAM loader on the device (svb_bamstream_*, csrc/bam_stream.cu).  The host reads
// the file and finds the BGZF members; records are inflated, walked, parsed and filtered in HBM, the bases of the reads to
// search never leave it.  Same output, byte for byte, as the host loader.
struct DevRead { string qname; int hp; bool search; vector<pair<int32_t, int32_t>> sfs; };

// the reference's output order (ping_pong.cpp:213-236) for reads [0, n): logical batches of --bsize, thread slots, qname order
static void print_reads(const Config& c, const vector<DevRead>& reads, size_t n, uint64_t& total_sfs) {
  string line;
  for (size_t b0 = 0; b0 < n; b0 += (size_t)c.bsize) {
    const size_t b1 = min(n, b0 + (size_t)c.bsize);
    for (int t = 0; t < c.threads; ++t) {
      map<string, vector<size_t>> slot;
      for (size_t i = b0 + (size_t)t; i < b1; i += (size_t)c.threads)
        if (reads[i].search) slot[reads[i].qname].push_back(i);
      for (auto& kv : slot) {
        bool first = true;
        for (size_t i : kv.second)
          for (const auto& f : reads[i].sfs) {
            line = (first ? kv.first : string("*")) + "\t" + to_string(f.first) + "\t" + to_string(f.second) + "\t" + to_string(reads[i].hp) + "\t\n";
            fwrite(line.data(), 1, line.size(), stdout);
            first = false;
            ++total_sfs;
          }
      }
    }
  }
}

// 1 = done, 0 = not applicable (the caller takes the host loader), -1 = failed
static int search_bam_on_device(svb_index_t* idx, const Config& c, size_t gpu_bases, uint64_t& processed, uint64_t& total_sfs) {
  BgzfSource src(c.bam);
  if (!src.ok() || !src.device_inflate()) return 0;   // not BGZF, or a file too small for the device path to pay
  int64_t header_bytes = 0;
  int n_ref = 0;
  {
    const int dev = bgzf_gpu_device();
    bgzf_gpu_device() = -1;                            // the header is read by the host reader
    BamReader hdr(c.bam, (size_t)1 << 20);
    bgzf_gpu_device() = dev;
    if (!hdr.ok()) { logmsg("critical", "cannot read BAM " + c.bam); return -1; }
    header_bytes = hdr.header_bytes();
    n_ref = (int)hdr.ref_names().size();
  }
  svb_bamstream_t* bs = nullptr;
  if (svb_bamstream_open(c.device, c.putative ? 1 : 0, n_ref, &bs) != SVB_OK) { logmsg("critical", string("svb_bamstream_open: ") + svb_last_error()); return -1; }
  vector<DevRead> reads;
  size_t attached = 0;   // reads [0, attached) carry their results already
  auto search_and_print = [&](bool final) -> bool {
    svb_sfs_out_t out;
    const double t_gpu = now_s();
    const int rc = svb_bamstream_search(bs, idx, c.overlap, c.assemble ? 1 : 0, &out);
    g_gpu_s += now_s() - t_gpu;
    if (rc != SVB_OK) { logmsg("critical", string("svb_bamstream_search: ") + svb_last_error()); return false; }
    int64_t r = 0;
    for (size_t i = attached; i < reads.size(); ++i)
      if (reads[i].search) {
        for (int64_t k = out.offs[r]; k < out.offs[r + 1]; ++k) reads[i].sfs.emplace_back(out.qs[k], out.len[k]);
        ++r;
      }
    svb_sfs_out_free(&out);
    attached = reads.size();
    // whole logical batches only: the rest waits for the reads that complete its batch
    const size_t n_print = final ? reads.size() : (reads.size() / (size_t)c.bsize) * (size_t)c.bsize;
    print_reads(c, reads, n_print, total_sfs);
    reads.erase(reads.begin(), reads.begin() + (long)n_print);
    attached -= n_print;
    return true;
  };
  const uint8_t* base = nullptr;
  vector<int64_t> io, oo;
  bool ok = true;
  while (ok && src.next_members(base, io, oo)) {
    svb_bam_recs_t recs;
    const double t_gpu = now_s();
    const int rc = svb_bamstream_window(bs, base, io.data(), oo.data(), (int64_t)io.size() - 1, header_bytes, &recs);
    g_gpu_s += now_s() - t_gpu;
    if (rc != SVB_OK) { logmsg("critical", string("truncated or corrupt BAM: ") + svb_last_error()); ok = false; break; }
    for (int64_t i = 0; i < recs.n; ++i) {
      ++processed;
      if (recs.state[i] == 0) continue;                                      // ping_pong.cpp:66-69
      if (recs.state[i] == 3) {                                              // :70-75
        logmsg("warning", "Alignment filtered due to l_qseq. Why are we here? Please check");
        continue;
      }
      if (recs.tid[i] < 0) { logmsg("critical", "core.tid < 0. Why are we here? Please check"); svb_bamstream_close(bs); svb_index_free(idx); exit(1); }  // :76-79
      reads.push_back(DevRead{string(recs.names + recs.name_offs[i], (size_t)(recs.name_offs[i + 1] - recs.name_offs[i])), (int)recs.hp[i], recs.state[i] == 2, {}});
    }
    if ((size_t)recs.batch_bases >= gpu_bases) ok = search_and_print(false);
  }
  if (ok && (src.failed() || svb_bamstream_pending_bytes(bs) != 0)) { logmsg("critical", "truncated or corrupt BAM"); ok = false; }
  if (ok) ok = search_and_print(true);
  svb_bamstream_close(bs);
  if (ok) logmsg("info", "BAM records decoded on GPU " + to_string(c.device) + ": " + to_string(processed));
  return ok ? 1 : -1;
}

static int run_search(const Config& c) {
  if (c.index.empty() || (c.fastx.empty() && c.bam.empty())) { cerr << SEARCH_USAGE << endl; return EXIT_FAILURE; }
  const double t_start = now_s();
  logmsg("info", "Restoring index..");
  svb_index_t* idx = nullptr;
  if (Rld::is_rld(c.index)) {   // an index written by the reference's `index -d` (ropebwt3 FMD)
    logmsg("warning", "ropebwt3 FMD index: re-indexing on the GPU for this run; convert it once with `SVDSS index --from-fmd`");
    vector<uint8_t> cat;
    vector<int64_t> offs;
    if (!fmd_to_contigs(c.index, cat, offs)) return EXIT_FAILURE;
    if (svb_index_build(cat.data(), offs.data(), (int64_t)offs.size() - 1, SVB_MEM_HOST, c.device, 0, &idx) != SVB_OK) {
      logmsg("critical", string("svb_index_build: ") + svb_last_error());
      return EXIT_FAILURE;
    }
  } else if (svb_index_load(c.index.c_str(), c.device, &idx) != SVB_OK) {
    logmsg("critical", string("svb_index_load: ") + svb_last_error());
    return EXIT_FAILURE;
  }
  const double t_loaded = now_s();
  // A GPU submission covers as many logical batches (--bsize) as fit ~1 Gbases of reads (SVB_SEARCH_SUBMIT_BASES): output
  // flows after every submission, like the reference's per-batch output (--omax); round 1 waited for 4 Gbases.
  // One deviation from the reference stays: when the same qname lands in one thread slot twice, the reference's
  // std::map<qname, vector<SFS>> lets Assembler::assemble merge the SFSs of both records; here every record is assembled
  // on its own (primary alignments carry unique names, so this needs a malformed BAM).
  size_t gpu_bases = (size_t)1 << 30;
  if (const char* e = getenv("SVB_SEARCH_SUBMIT_BASES")) { const long long v = atoll(e); if (v > 0) gpu_bases = (size_t)v; }
  vector<PendingRead> reads;
  vector<uint8_t> cat;
  uint64_t total_sfs = 0, processed = 0;
  const bool packed = !c.bam.empty();   // BAM records are handed to the GPU as stored: 4 bits per base
  auto maybe_flush = [&]() {
    if (cat.size() >= (packed ? gpu_bases / 2 : gpu_bases) && reads.size() % (size_t)c.bsize == 0) return flush(idx, c, reads, cat, total_sfs, packed);
    return true;
  };
  logmsg("info", "Extracting SFS strings on GPU " + to_string(c.device) + " (ordering as with " + to_string(c.threads) + " threads)..");
  bool ok = true;
  int on_device = 0;
  if (!c.bam.empty() && c.gpu_inflate && !getenv("SVB_BAM_HOST_PARSE")) {
    on_device = search_bam_on_device(idx, c, gpu_bases, processed, total_sfs);
    if (on_device < 0) ok = false;
  }
  if (on_device != 0) {
    // done (or failed) above
  } else if (!c.bam.empty()) {
    BamReader bam(c.bam);
    if (!bam.ok()) { logmsg("critical", "cannot read BAM " + c.bam); svb_index_free(idx); return EXIT_FAILURE; }
    bam.want_view(true);        // the packed sequence stays where it was inflated; no host-side decode, no copy of the reads that are not searched
    BamRecord r;
    int st;
    while (ok && (st = bam.next(r)) == 1) {
      ++processed;
      if (r.flag & 0x4 || r.flag & 0x800 || r.flag & 0x100) continue;       // ping_pong.cpp:66-69
      if (r.l_qseq < 100) {                                                  // :70-75
        logmsg("warning", "Alignment filtered due to l_qseq. Why are we here? Please check");
        continue;
      }
      if (r.tid < 0) { logmsg("critical", "core.tid < 0. Why are we here? Please check"); svb_index_free(idx); exit(1); }  // :76-79
      const int xf = r.has_xf ? (int)r.xf : 0, hp = r.has_hp ? (int)r.hp : 0; // :196-201
      PendingRead pr{r.qname, hp, (int64_t)cat.size(), 0, !(c.putative && xf != 0), r.l_qseq};
      if (pr.search) cat.insert(cat.end(), r.seq4_view, r.seq4_view + ((size_t)r.l_qseq + 1) / 2);
      pr.hi = (int64_t)cat.size();
      reads.push_back(pr);
      ok = maybe_flush();
    }
    if (ok && st < 0) { logmsg("critical", "truncated or corrupt BAM"); ok = false; }
  } else {
    logmsg("warning", "FASTX mode is not optimized (higher running times and larger SFSs set).");  // ping_pong.cpp:253-254
    FastxReader fx(c.fastx);
    if (!fx.ok()) { logmsg("critical", "cannot open " + c.fastx); svb_index_free(idx); return EXIT_FAILURE; }
    FastxRecord r;
    const uint8_t* t6 = nt6_table();
    while (ok && fx.next(r)) {
      ++processed;
      PendingRead pr{r.name, 0, (int64_t)cat.size(), 0, true, (int32_t)r.seq.size()};
      for (char ch : r.seq) cat.push_back(t6[(uint8_t)ch]);                  // rb3_char2nt6, ping_pong.cpp:158
      pr.hi = (int64_t)cat.size();
      reads.push_back(pr);
      ok = maybe_flush();
    }
  }
  if (ok) ok = flush(idx, c, reads, cat, total_sfs, packed);
  fflush(stdout);
  svb_index_free(idx);
  if (!ok) return EXIT_FAILURE;
  char tbuf[160];
  snprintf(tbuf, sizeof(tbuf), " (index load %.2f s, input + output %.2f s, GPU calls %.2f s)", t_loaded - t_start,
           now_s() - t_loaded - g_gpu_s, g_gpu_s);
  logmsg("info", "records read: " + to_string(processed) + ", SFS lines written: " + to_string(total_sfs) + tbuf);
  return EXIT_SUCCESS;
}

// `SVDSS _clipper --reference FA --clips-in CLIPS.tsv [--regions-in REGIONS.tsv] [--threads N] [--verbose]`:
// clips in the format of `call --clipped --clips` (name chrom p l L|R), regions as `low high` lines;
// prints the imprecise records like the tail of `call --clipped`; --verbose lists the combined
// breakpoints in the order they reach Clipper::cluster (std::unordered_map iteration order).
static int run_clipper_hook(const Config& c) {
  if (c.reference.empty() || c.clips_in.empty()) return EXIT_FAILURE;
  vector<string> chroms;
  unordered_map<string, string> seqs;
  FastxReader fx(c.reference);
  if (!fx.ok()) return EXIT_FAILURE;
  FastxRecord r;
  while (fx.next(r)) { for (auto& ch : r.seq) ch = (char)toupper((unsigned char)ch); chroms.push_back(r.name); seqs[r.name] = r.seq; }
  vector<Clip> clips;
  {
    ifstream f(c.clips_in);
    if (!f.is_open()) return EXIT_FAILURE;
    string name, chrom, side;
    unsigned p, l;
    while (f >> name >> chrom >> p >> l >> side) clips.push_back(Clip(name, chrom, p, l, side == "L"));
  }
  vector<pair<int, int>> regions;
  if (!c.regions_in.empty()) {
    ifstream f(c.regions_in);
    if (!f.is_open()) return EXIT_FAILURE;
    int lo, hi;
    while (f >> lo >> hi) regions.emplace_back(lo, hi);
  }
  Clipper kept(clips, &chroms, &seqs);
  const vector<SV> svs = clipped_calls(clips, chroms, seqs, c.threads, regions, &kept);
  if (c.verbose) {
    for (const Clip& k : kept.r_combined) fprintf(stderr, "combined\tR\t%s\t%u\t%u\t%u\n", k.chrom.c_str(), k.p, k.l, k.w);
    for (const Clip& k : kept.l_combined) fprintf(stderr, "combined\tL\t%s\t%u\t%u\t%u\n", k.chrom.c_str(), k.p, k.l, k.w);
    for (const Clip& k : kept.rclips) fprintf(stderr, "clustered\tR\t%s\t%u\t%u\t%u\n", k.chrom.c_str(), k.p, k.l, k.w);
    for (const Clip& k : kept.lclips) fprintf(stderr, "clustered\tL\t%s\t%u\t%u\t%u\n", k.chrom.c_str(), k.p, k.l, k.w);
  }
  for (const SV& sv : svs) cout << sv.vcf_line() << "\n";
  return EXIT_SUCCESS;
}

// `SVDSS _fmd encode BWT.bin OUT.fmd | decode IN.fmd BWT.bin | contigs IN.fmd OUT.fa`: rld.hpp without the
// GPU (tests/test_fmd_cpu.py).  BWT.bin = one nt6 code per byte; OUT.fa = the forward strands, named seq<k>.
static int run_fmd_hook(const vector<string>& pos) {
  if (pos.size() != 3) return EXIT_FAILURE;
  string err;
  if (pos[0] == "encode") {
    ifstream f(pos[1], ios::binary);
    if (!f.is_open()) return EXIT_FAILURE;
    vector<uint8_t> bwt((istreambuf_iterator<char>(f)), istreambuf_iterator<char>());
    if (!Rld::write(pos[2], bwt.data(), bwt.size(), err)) { logmsg("critical", err); return EXIT_FAILURE; }
    return EXIT_SUCCESS;
  }
  if (pos[0] == "decode") {
    RldFile rf;
    vector<uint8_t> bwt;
    if (!Rld::read(pos[1], rf, err) || !Rld::decode_bwt(rf, bwt, err)) { logmsg("critical", err); return EXIT_FAILURE; }
    ofstream o(pos[2], ios::binary);
    o.write((const char*)bwt.data(), (streamsize)bwt.size());
    return o.good() ? EXIT_SUCCESS : EXIT_FAILURE;
  }
  if (pos[0] == "contigs") {
    vector<uint8_t> cat;
    vector<int64_t> offs;
    if (!fmd_to_contigs(pos[1], cat, offs)) return EXIT_FAILURE;
    ofstream o(pos[2]);
    for (size_t k = 0; k + 1 < offs.size(); ++k) {
      o << ">seq" << k << "\n";
      for (int64_t i = offs[k]; i < offs[k + 1]; ++i) o << "$ACGTN"[cat[(size_t)i]];
      o << "\n";
    }
    return o.good() ? EXIT_SUCCESS : EXIT_FAILURE;
  }
  return EXIT_FAILURE;
}

int main(int argc, char** argv) {
  time_t t0;
  time(&t0);
  if (argc == 1) { cerr << MAIN_USAGE << endl; exit(EXIT_FAILURE); }   // main.cpp:27-31
  Config c;
  vector<string> pos;
  const string mode = argv[1];
  if (mode == "index" ? !parse_index(argc, argv, c, pos) : !parse_common(argc, argv, c, pos)) exit(EXIT_FAILURE);
  if (c.version) { cout << "SVDSS, " << VERSION << endl; exit(EXIT_SUCCESS); }
  if (c.help) { cerr << (mode == "index" ? INDEX_USAGE : mode == "smooth" ? SMOOTH_USAGE : mode == "search" ? SEARCH_USAGE : mode == "call" ? CALL_USAGE : MAIN_USAGE) << endl; exit(EXIT_SUCCESS); }
  int rc;
  if (mode == "_ratio") {   // test hook: fuzz_ratio of filter_sv_chains (tests/test_cluster_cpu.py)
    if (pos.size() != 2) exit(EXIT_FAILURE);
    printf("%.6f\n", fuzz_ratio(pos[0], pos[1]));
    return 0;
  }
  if (mode == "_bamread") {   // measurement hook: what the BAM reader hands `search` per second (tools/bench_bamread.py), no GPU
    if (pos.size() != 1) return EXIT_FAILURE;
    double cuda_init_s = 0;
    if (c.gpu_inflate) {   // the CUDA context is created before the clock starts and reported apart (0.3 s as a rule, seconds on a busy box)
      const double ti = now_s();
      svb_host_free_pinned(svb_host_alloc_pinned(1));
      cuda_init_s = now_s() - ti;
    }
    if (c.gpu_inflate && getenv("SVB_BAMREAD_DEVICE")) {   // the device loader of `search` (svb_bamstream_*), without the search
      const double t0 = now_s();
      BgzfSource src(pos[0]);
      if (!src.ok() || !src.device_inflate()) return EXIT_FAILURE;
      int64_t header_bytes = 0;
      int n_ref = 0;
      { const int dev = bgzf_gpu_device(); bgzf_gpu_device() = -1; BamReader hdr(pos[0], (size_t)1 << 20); bgzf_gpu_device() = dev; if (!hdr.ok()) return EXIT_FAILURE; header_bytes = hdr.header_bytes(); n_ref = (int)hdr.ref_names().size(); }
      svb_bamstream_t* bs = nullptr;
      const bool align = string(getenv("SVB_BAMREAD_DEVICE")) == "align";   // the Clusterer's scan: alignments + a fetch of the reads with XF = 0
      if (svb_bamstream_open(c.device, align ? -1 : 1, n_ref, &bs) != SVB_OK) return EXIT_FAILURE;
      const uint8_t* base = nullptr;
      vector<int64_t> io, oo, want;
      uint64_t n = 0, bases = 0, kept = 0, name_bytes = 0;
      double t_dev = 0;
      bool ok = true;
      while (src.next_members(base, io, oo)) {
        svb_bam_recs_t recs;
        const double t1 = now_s();
        if (svb_bamstream_window(bs, base, io.data(), oo.data(), (int64_t)io.size() - 1, header_bytes, &recs) != SVB_OK) { cerr << svb_last_error() << endl; ok = false; break; }
        t_dev += now_s() - t1;
        for (int64_t i = 0; i < recs.n; ++i) { ++n; bases += (uint64_t)recs.l_qseq[i]; if (recs.state[i] == 2) ++kept; }
        if (align) {
          want.clear();
          for (int64_t i = 0; i < recs.n; ++i) if (recs.state[i] == 1 && recs.xf[i] == 0) want.push_back(i);
          const uint32_t* fc = nullptr; const int64_t* fco = nullptr; const uint8_t* fs = nullptr; const int64_t* fso = nullptr;
          const double t2 = now_s();
          if (svb_bamstream_fetch(bs, want.data(), (int64_t)want.size(), &fc, &fco, &fs, &fso) != SVB_OK) { cerr << svb_last_error() << endl; ok = false; break; }
          t_dev += now_s() - t2;
          kept += want.size();
        }
        name_bytes += (uint64_t)recs.name_offs[recs.n];
      }
      ok = ok && !src.failed() && svb_bamstream_pending_bytes(bs) == 0;
      const double dt = now_s() - t0;
      svb_bamstream_close(bs);
      printf("{\"records\": %llu, \"kept\": %llu, \"bases\": %llu, \"seq_sum\": 0, \"name_bytes\": %llu, \"seconds\": %.3f, \"cuda_init_seconds\": %.3f, \"device_call_seconds\": %.3f, \"records_per_s\": %.0f, \"Gbases_per_s\": %.3f}\n",
             (unsigned long long)n, (unsigned long long)kept, (unsigned long long)bases, (unsigned long long)name_bytes, dt, cuda_init_s, t_dev, n / dt, bases / dt / 1e9);
      return ok ? EXIT_SUCCESS : EXIT_FAILURE;
    }
    const double t0 = now_s();                     // the first window is inflated by the constructor
    BamReader bam(pos[0]);
    if (!bam.ok()) return EXIT_FAILURE;
    bam.want_view(true);        // as run_search reads it
    BamRecord r;
    int st;
    uint64_t n = 0, bases = 0, kept = 0, seq_sum = 0;
    vector<uint8_t> cat;
    while ((st = bam.next(r)) == 1) {
      ++n; bases += (uint64_t)r.l_qseq;
      const size_t seq_bytes = ((size_t)r.l_qseq + 1) / 2;
      // a checksum the host and the device inflate must agree on: three bytes of every record (more would time the checksum's cache misses)
      if (seq_bytes) seq_sum = (seq_sum * 31 + r.seq4_view[0]) * 31 + r.seq4_view[seq_bytes / 2] + 7 * r.seq4_view[seq_bytes - 1];
      if (!(r.has_xf && r.xf != 0)) { cat.insert(cat.end(), r.seq4_view, r.seq4_view + seq_bytes); ++kept; }   // what run_search keeps of a record
      if (cat.size() > ((size_t)1 << 30)) cat.clear();
    }
    const double dt = now_s() - t0;
    printf("{\"records\": %llu, \"kept\": %llu, \"bases\": %llu, \"seq_sum\": %llu, \"seconds\": %.3f, \"cuda_init_seconds\": %.3f, \"records_per_s\": %.0f, \"Gbases_per_s\": %.3f}\n",
           (unsigned long long)n, (unsigned long long)kept, (unsigned long long)bases, (unsigned long long)seq_sum, dt, cuda_init_s, n / dt, bases / dt / 1e9);
    return st < 0 ? EXIT_FAILURE : EXIT_SUCCESS;
  }
  if (mode == "_fmd") return run_fmd_hook(pos);
  if (mode == "_clipper") return run_clipper_hook(c);   // test hook: Clipper::call without the GPU stages (tests/test_clipper_cpu.py)
  if (mode == "index") rc = run_index(c, pos);
  else if (mode == "smooth") {
    if (c.reference.empty() || c.bam.empty()) { cerr << SMOOTH_USAGE << endl; exit(EXIT_FAILURE); }   // main.cpp:72-75
    SmoothConfig sc;
    sc.bam = c.bam; sc.reference = c.reference; sc.min_mapq = (unsigned)c.min_mapq; sc.accp = c.accp; sc.threads = c.threads;
    rc = run_smooth(sc, [](const char* l, const string& m) { logmsg(l, m); });
  }
  else if (mode == "search") rc = run_search(c);
  else if (mode == "call") {
    if (c.reference.empty() || (c.clusters_in.empty() && (c.bam.empty() || c.sfs.empty()))) { cerr << CALL_USAGE << endl; exit(EXIT_FAILURE); }  // main.cpp:56-59
    CallConfig cc;
    cc.reference = c.reference; cc.clusters_in = c.clusters_in; cc.poa_out = c.poa;
    cc.min_cluster_weight = (unsigned)c.min_cluster_weight; cc.min_sv_length = (unsigned)c.min_sv_length;
    cc.min_ratio = c.min_ratio; cc.device = c.device;
    cc.bam = c.bam; cc.sfs = c.sfs; cc.clusters_out = c.clusters_out; cc.min_mapq = (unsigned)c.min_mapq;
    cc.threads = c.threads; cc.batch_size = c.bsize; cc.useht = !c.noht; cc.cluster_only = c.cluster_only;
    cc.clipped = c.clipped; cc.clips_out = c.clips_out;
    rc = run_call(cc, [](const char* l, const string& m) { logmsg(l, m); });
  }
  else { cerr << MAIN_USAGE << endl; exit(EXIT_FAILURE); }
  if (rc != EXIT_SUCCESS) return rc;
  time_t t1;
  time(&t1);
  logmsg("info", "All done! Runtime: " + to_string((long)(t1 - t0)) + " seconds");   // main.cpp:85
  return 0;
}
