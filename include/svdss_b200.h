/*
 * svdss_b200.h -- C ABI of libsvdss_b200: the B200 (sm_100a) implementation of SVDSS's
 * data-parallel hot path (SFS extraction by ping-pong FMD search; cluster POA; ksw2 realignment).
 *
 * The reference (Parsoa/SVDSS v2.1.1) has no plugin/FFI layer: its hot path is reached by direct
 * C++ calls into three statically linked C libraries.  Each entry point below names the reference
 * call it stands in for (file:line relative to the reference tree).  INTEGRATION.md shows the
 * host-side binding a maintainer would add.
 *
 * Conventions: every function returns SVB_OK (0) or a negative errno-style code; no exceptions
 * cross the ABI; svb_last_error() returns a thread-local message for the last failure; the caller
 * owns every buffer it passes in; outputs are released with the matching *_free.  `mem` arguments
 * say where caller buffers live: SVB_MEM_HOST (0) or SVB_MEM_DEVICE (1, a CUDA device pointer on
 * `device`).  All sequences are nt6-coded bytes ($=0 A=1 C=2 G=3 T=4 N=5; ping_pong.hpp:46-52).
 * There is NO CPU fallback: without a CUDA device every compute entry fails with SVB_ECUDA.
 */
#ifndef SVDSS_B200_H
#define SVDSS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SVB_OK 0
#define SVB_EINVAL (-22)
#define SVB_ENOMEM (-12)
#define SVB_ECUDA (-5)
#define SVB_EIO (-74)
#define SVB_ERANGE (-34)

#define SVB_MEM_HOST 0
#define SVB_MEM_DEVICE 1

#define SVB_API __attribute__((visibility("default")))

typedef struct svb_index svb_index_t; /* immutable once built; shareable between host threads */
typedef struct svb_reads svb_reads_t; /* a batch of reads resident in HBM */

SVB_API const char* svb_last_error(void);
SVB_API int svb_device_count(void);
SVB_API const char* svb_version(void);

/* ------------------------------------------------------------------ index (a2, a3, `index`) -- */

typedef struct {
  int64_t n;           /* BWT length = 2 * sum(len_i + 1)                                        */
  int64_t acc[7];      /* acc[c] = #symbols < c  (rb3_fmi_t::acc)                    */
  int64_t n_blocks;
  int32_t block_bytes; /* 64 or 128                                                              */
  int32_t block_syms;  /* 128 or 256                                                             */
  int64_t n_contigs;
  int64_t device_bytes;
  int32_t device;
} svb_index_info_t;

/* Build the FMD index of {S_i, rc(S_i)} on the GPU: suffix sort, BWT, sampled-Occ block array.
 * Replaces `SVDSS index` = ropebwt3 main_build (main.cpp:15-17,34-37; CMakeLists.txt:154-169).
 * block_bytes: 64, 128, or 0 for the library default. */
SVB_API int svb_index_build(const uint8_t* contigs_nt6, const int64_t* offs /* n_contigs+1 */,
                            int64_t n_contigs, int mem, int device, int block_bytes,
                            svb_index_t** out);
/* Build the block array from a ready BWT (nt6 codes). Test hook for the rank / search kernels. */
SVB_API int svb_index_from_bwt(const uint8_t* bwt, int64_t n, int mem, int device, int block_bytes,
                               svb_index_t** out);
/* rb3_fmi_restore(&index, path, 0) (ping_pong.cpp:244-245): load an index file into HBM of `device`.
 * Two formats, told apart by their magic: the file svb_index_save writes (layout private to this
 * library; the reference treats the index as opaque between `index` and `search`, run_svdss:137-164),
 * and the file the reference's own `index -d` writes -- ropebwt3's FMD ("RLD\3", rld0.c) -- which is
 * decoded, inverted to the sequences it indexes and re-indexed on the GPU (1-2 minutes of host work
 * for a human genome; the FMD layout is restated from memory, ropebwt3 not being vendored by the
 * reference: see svdss_b200/host/rld.hpp). */
SVB_API int svb_index_load(const char* path, int device, svb_index_t** out);
SVB_API int svb_index_save(const svb_index_t* idx, const char* path);
SVB_API void svb_index_free(svb_index_t* idx);
SVB_API int svb_index_info(const svb_index_t* idx, svb_index_info_t* info);
/* Decode the block array back to BWT symbols (host buffer of n bytes). */
SVB_API int svb_index_get_bwt(const svb_index_t* idx, uint8_t* bwt_host);
/* Debug/test: suffix array of a '$'-terminated nt6 text with sentinels ordered by position. */
SVB_API int svb_suffix_array(const uint8_t* text, int64_t n, int mem, int device, int64_t* sa_host);

/* ------------------------------------------------------------------------------ rank (a2) -- */

/* rb3_fmi_rank2a(f, k, l, ok, ol) for n query pairs: Occ of all six symbols at k[i] and l[i]
 * (ropebwt3 fm-index.h, reached from ping_pong.cpp:20,35 via rb3_fmd_extend). Host buffers.
 * ok6/ol6 are n*6 int64. */
SVB_API int svb_rank2a(const svb_index_t* idx, const int64_t* k, const int64_t* l, int64_t n,
                       int64_t* ok6, int64_t* ol6);

/* "FMD rank GB/s" microbenchmark: n_queries uniformly random (k, k+delta) pairs, one
 * backward extension each (Occ of one symbol at both ends), generated on the device; timed with
 * CUDA events over `iters` launches.  bytes = block_bytes * distinct blocks touched. */
SVB_API int svb_rank_bench(const svb_index_t* idx, int64_t n_queries, int64_t delta, uint64_t seed,
                           int iters, float* ms_per_iter, int64_t* blocks_touched_per_iter);

/* ------------------------------------------------------------------ SFS search (a1, a4, a10) -- */

typedef struct {
  /* results: one record per (assembled) SFS, grouped by read in input order; within a read
   * ascending qs when assemble != 0 (assembler.cpp:36), descending qs otherwise (emit order of
   * ping_pong.cpp:39-41).  offs[r]..offs[r+1] indexes the records of read r. */
  int64_t n_reads;
  int64_t n_sfs;
  int64_t* offs; /* n_reads + 1 */
  int32_t* qs;   /* SFS::qs  (sfs.hpp:36) */
  int32_t* len;  /* SFS::l   (sfs.hpp:38) */
  /* measurement (filled by every search call) */
  int64_t n_ext;            /* extensions performed (= rb3_fmd_extend calls of the reference)   */
  int64_t n_blocks_touched; /* sum over extensions of distinct index blocks read (1 or 2)       */
  float kernel_ms;          /* search kernel only, CUDA events on the launch stream             */
  float device_ms;          /* whole device pipeline of this call (H2D + kernels + D2H)         */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int32_t launches;         /* kernels launched by this call                                    */
  int32_t block_bytes;
  int64_t n_text_ext;       /* of n_ext: extensions answered by comparing the read with the text
                               (located-match mode, 2 bytes each) instead of an index block fetch */
} svb_sfs_out_t;

/* PingPong::process_batch (ping_pong.cpp:176-209) for a whole batch: runs
 * PingPong::ping_pong_search (ping_pong.cpp:4-49) on every read and, if assemble != 0,
 * Assembler::assemble (assembler.cpp:34-56) per read.  HOST buffers; copies are pipelined with the
 * kernel internally, the call is synchronous at return.  overlap must be <= 0 (config.hpp:82 fixes
 * it at -1; 0 is the "relaxed" branch of ping_pong.cpp:44-45).  Reads of length 0 yield nothing.
 * The BAM-level filters of ping_pong.cpp:66-79,196-203 stay with the caller. */
SVB_API int svb_sfs_batch(const svb_index_t* idx, const uint8_t* nt6_concat,
                          const int64_t* offs /* n_reads+1 */, int64_t n_reads, int overlap,
                          int assemble, svb_sfs_out_t* out);

/* The same call for reads as BAM stores them -- what load_batch_bam gets from bam_get_seq before its
 * decode loop (ping_pong.cpp:88-94): 4 bits per base, two bases per byte, first base in the high
 * nibble, htslib nt16 codes (=ACMGRSVTWYHKDBN), every read starting on a byte boundary.
 * seq4_offs[r] is the byte offset of read r in seq4 (n_reads + 1 entries), l_qseq[r] its length in
 * bases (bam1_core_t::l_qseq).  The nt16 -> ASCII -> nt6 decode of ping_pong.cpp:90-94 runs on the
 * GPU, so half as many bytes cross PCIe and the host does not touch the bases at all.  Results are
 * identical to svb_sfs_batch on the decoded reads.  128-byte-block indexes only.  HOST buffers. */
SVB_API int svb_sfs_batch_bam4(const svb_index_t* idx, const uint8_t* seq4, const int64_t* seq4_offs /* n_reads+1 */,
                               const int32_t* l_qseq /* n_reads */, int64_t n_reads, int overlap, int assemble,
                               svb_sfs_out_t* out);
/* Test / bench utility: pack nt6 reads resident on `device` into that BAM layout (device buffers). */
SVB_API int svb_pack4_device(const uint8_t* d_nt6_concat, const int64_t* d_offs /* n_reads+1 */,
                             const int64_t* d_seq4_offs /* n_reads+1 */, int64_t n_reads, int device, uint8_t* d_out);

/* Host utility (used by the streamed search with SVB_STREAM_PACK2=1, off by default: slower on a 16-core host): re-pack reads from BAM's 4 bits per base to
 * 2 bits per base with all host threads (threads <= 0) -- half the bytes again for the H2D copy that bounds
 * the end-to-end search.  Read r goes to (l_qseq[r] + 3) / 4 bytes at out_offs[r], first base in the two high
 * bits, A C G T -> 0 1 2 3; exception[r] = 1 if the read holds any other code (it then has to travel as it is). */
SVB_API int svb_pack2_host(const uint8_t* seq4, const int64_t* seq4_offs /* n_reads+1 */, const int32_t* l_qseq,
                           int64_t n_reads, uint8_t* out, const int64_t* out_offs /* n_reads */,
                           uint8_t* exception /* n_reads */, int threads);

/* The streamed form of the same packing: the part of reads r_lo..r_hi inside base positions [o, o + nb) of the
 * concatenated batch, widened to whole packed bytes, into `stage` = bytes [*pa2, *pe2) of the packed layout;
 * positions of bases without a 2-bit form (they all decode to N) go to exc_pos (*n_exc of them). */
SVB_API int svb_pack2_chunk(const uint8_t* seq4, const int64_t* seq4_offs, const int64_t* offs, const int64_t* packed_offs,
                            int64_t r_lo, int64_t r_hi, int64_t o, int64_t nb, uint8_t* stage, int64_t stage_cap,
                            int64_t* pa2, int64_t* pe2, int64_t* exc_pos, int64_t exc_cap, int64_t* n_exc, int threads);

/* Test utility for the other half: decode such a batch on `device` (k_unpack2) and return one nt6 byte per base. */
SVB_API int svb_unpack2_device(const uint8_t* packed, const int64_t* packed_offs /* n_reads+1 */,
                               const int64_t* offs /* n_reads+1, bases */, int64_t n_reads, int device, uint8_t* out_host);

/* BGZF members inflated on the GPU (the reference inflates through htslib's bgzf_mt threads,
 * ping_pong.cpp:249, clusterer.cpp:13).  The caller walks the gzip headers and passes the raw-deflate payloads of
 * n_members members back to back (member m = comp[in_offs[m], in_offs[m+1])) and where their ISIZE bytes go
 * (out_host[out_offs[m], out_offs[m+1]), at most 64 KiB each); both offset arrays start at 0.  status_host (optional, one per member): 0 or
 * the reason a member did not inflate; any non-zero status makes the call return SVB_EIO after the copies.
 * kernel_ms (optional): device time of the inflate kernel.  host/io.hpp uses it with `--gpu-inflate`. */
SVB_API int svb_bgzf_inflate_device(const uint8_t* comp, const int64_t* in_offs /* n_members+1 */,
                                    const int64_t* out_offs /* n_members+1 */, int64_t n_members, int device,
                                    uint8_t* out_host, int32_t* status_host, float* kernel_ms);

/* Pinned host memory for a reader's windows (svb_bgzf_inflate_device copies at PCIe speed only from / to pinned
 * buffers).  NULL when there is no device or the allocation fails. */
SVB_API void* svb_host_alloc_pinned(size_t bytes);
SVB_API void svb_host_free_pinned(void* p);

/* ---- the BAM loader of `search` on the device (PingPong::load_batch_bam, ping_pong.cpp:53-131, with the filters of
 * :66-79 and :196-203).  The caller reads the file and finds the BGZF members, as for svb_bgzf_inflate_device; a
 * stream inflates window after window into HBM, walks the records where they lie (the unfinished record at the end of
 * a window is completed by the next), parses core fields and the XF / HP tags, and decodes the bases of the reads
 * that will be searched (nt16 -> nt6, :88-94) into a device batch.  Names, flags and tags come back; bases never do. */
typedef struct svb_bamstream svb_bamstream_t;
typedef struct {
  int64_t n;                 /* records completed by this window, in file order                                    */
  const uint16_t* flag;      /* bam1_core_t::flag                                                                   */
  const int32_t* tid;        /* ::tid (the reference exits on a kept record with tid < 0, ping_pong.cpp:76-79)      */
  const int32_t* l_qseq;     /* ::l_qseq                                                                            */
  const int32_t* xf;         /* XF:i or 0 (bam_aux_get / bam_aux2i, :196-198)                                       */
  const int32_t* hp;         /* HP:i or 0 (:199-201)                                                                */
  const uint8_t* state;      /* 0 dropped (unmapped / secondary / supplementary, :66-69); 3 dropped for l_qseq < 100
                              * (:70-75); 1 kept, not searched (XF != 0 with the putative filter, :202); 2 searched:
                              * its bases were appended to the device batch                                         */
  const int64_t* name_offs;  /* n + 1: qname of record i = names[name_offs[i], name_offs[i+1]) (empty for dropped)  */
  const char* names;
  int64_t batch_reads, batch_bases;   /* the device batch after this window                                         */
  int64_t h2d_bytes, d2h_bytes;
  /* alignment mode only (svb_bamstream_open with putative < 0), else NULL: what the Clusterer's scan reads of a record
   * (clusterer.cpp:58-153); state is then 0, 3 or 1 and no bases are batched */
  const int32_t* pos;        /* bam1_core_t::pos                                                                    */
  const int32_t* endpos;     /* bam_endpos                                                                          */
  const int32_t* n_cigar;    /* ops of the CIGAR in force (the CG:B,I tag's when the field holds the placeholder)   */
  const uint8_t* mapq;
} svb_bam_recs_t;
/* putative: 1 = the XF rule of ping_pong.cpp:202, 0 = --noputative, -1 = alignment mode (no search batch: the Clusterer's
 * scan, see svb_bam_recs_t and svb_bamstream_fetch); n_ref: reference sequences in the BAM header (the record walk
 * uses it to tell a record start from payload; <= 0 if unknown) */
SVB_API int svb_bamstream_open(int device, int putative, int n_ref, svb_bamstream_t** out);
/* One window: n_members BGZF members as svb_bgzf_inflate_device takes them (host buffers).  skip_bytes (first call
 * only): inflated bytes to pass over before the first record = the BAM header, which the caller has parsed.
 * recs points into memory owned by the stream, valid until the next call on it. */
SVB_API int svb_bamstream_window(svb_bamstream_t* s, const uint8_t* comp, const int64_t* in_offs /* n_members+1 */,
                                 const int64_t* out_offs /* n_members+1 */, int64_t n_members, int64_t skip_bytes,
                                 svb_bam_recs_t* recs);
/* Alignment mode: CIGAR ops and packed bases (BAM's 4 bits per base) of n_rec records of the LAST window (indices into
 * its svb_bam_recs_t, ascending or not), back to back; the arrays belong to the stream until its next call. */
SVB_API int svb_bamstream_fetch(svb_bamstream_t* s, const int64_t* rec, int64_t n_rec, const uint32_t** cigar,
                                const int64_t** cigar_offs /* n_rec+1 */, const uint8_t** seq4, const int64_t** seq4_offs /* n_rec+1, bytes */);
/* bytes of an unfinished record left behind the last window: not 0 at the end of the file = truncated BAM */
SVB_API int64_t svb_bamstream_pending_bytes(const svb_bamstream_t* s);
/* svb_sfs_resident over the reads batched so far (read r of the result = the r-th record of state 2 since the last
 * search); empties the batch. */
SVB_API int svb_bamstream_search(svb_bamstream_t* s, const svb_index_t* idx, int overlap, int assemble, svb_sfs_out_t* out);
SVB_API void svb_bamstream_close(svb_bamstream_t* s);

/* Same search with the batch already resident in HBM (kernel-only measurement; multi-batch reuse). */
SVB_API int svb_reads_upload(const uint8_t* nt6_concat, const int64_t* offs, int64_t n_reads,
                             int mem, int device, svb_reads_t** out);
SVB_API void svb_reads_free(svb_reads_t* reads);
SVB_API int svb_sfs_resident(const svb_index_t* idx, const svb_reads_t* reads, int overlap,
                             int assemble, svb_sfs_out_t* out);
SVB_API void svb_sfs_out_free(svb_sfs_out_t* out);

/* ------------------------------------------------------------------ ksw2 realignment (a6) -- */

typedef struct {
  int64_t n_pairs;
  int32_t* score;       /* ez.score per pair (KSW_NEG_INF = -0x40000000 for an empty query/target) */
  int64_t* cigar_offs;  /* n_pairs + 1, indexes cigar[]                                          */
  uint32_t* cigar;      /* ez.cigar: len << 4 | op, op 0=M 1=I 2=D (caller.cpp:352-354)          */
  int64_t n_cigar;
  int64_t cells;        /* sum of ql*tl                                                          */
  float kernel_ms;      /* DP + traceback kernels, CUDA events                                   */
  float device_ms;      /* whole call on the device incl. H2D/D2H                                */
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int32_t launches;
  int32_t waves;        /* batches the pairs were split into to bound traceback memory           */
} svb_ksw_out_t;

/* ksw_extd2_sse(0, ql, qs, tl, ts, 5, mat, gapo, gape, gapo2, gape2, -1, -1, -1, 0, &ez) for every
 * (query = consensus, target = reference window) pair of a batch (caller.cpp:332-355): global
 * alignment, two-piece affine gaps, full matrix, score + left-aligned CIGAR.  Sequences are
 * _char26_table codes 0..4 (caller.hpp:25-37).  `sc_n` is the score of any pair involving code 4:
 * ksw2 without KSW_EZ_GENERIC_SC uses -gape2 when mat[24] == 0, which is what the reference passes.
 * HOST buffers. */
SVB_API int svb_ksw_extd2_batch(const uint8_t* q_concat, const int64_t* q_offs /* n_pairs+1 */,
                                const uint8_t* t_concat, const int64_t* t_offs /* n_pairs+1 */,
                                int64_t n_pairs, int match, int mismatch, int sc_n, int gapo, int gape,
                                int gapo2, int gape2, int device, svb_ksw_out_t* out);
SVB_API void svb_ksw_out_free(svb_ksw_out_t* out);

/* ------------------------------------------------------------------ cluster POA (a7) -- */

typedef struct {
  int64_t n_clusters;
  int64_t* cons_offs;   /* n_clusters + 1, indexes cons[]                                         */
  uint8_t* cons;        /* consensus bases, codes 0..4 (abc->cons_base[0], caller.cpp:292-297)    */
  int32_t* status;      /* per cluster: 0 ok; bit 1 = workspace overflow, bit 2 = band clamped    */
  int64_t cells;        /* DP cells computed (sum over reads and rows of the band width)          */
  float kernel_ms;
  float device_ms;
  int64_t h2d_bytes;
  int64_t d2h_bytes;
  int32_t launches;
  int32_t reruns;       /* clusters redone with worst-case workspace                              */
} svb_poa_out_t;

/* Caller::run_poa (caller.cpp:257-308) for every cluster of a batch: abpoa_msa with the reference's
 * parameters (global, convex gap 4/2 + 24/1, match 2, mismatch 4, adaptive band 10 + 0.01*qlen,
 * no seeding, not progressive => reads added in input order) and the heaviest-bundling consensus
 * (max_n_cons = 1).  Sequences are _char26_table codes 0..4; cluster c owns sequences
 * cluster_offs[c] .. cluster_offs[c+1] of seq_offs.  A cluster without sequences yields an empty
 * consensus (n_cons == 0 => "" at caller.cpp:294).  HOST buffers. */
SVB_API int svb_poa_batch(const uint8_t* seqs_concat, const int64_t* seq_offs /* n_seqs+1 */,
                          const int64_t* cluster_offs /* n_clusters+1 */, int64_t n_clusters, int device,
                          svb_poa_out_t* out);
SVB_API void svb_poa_out_free(svb_poa_out_t* out);

/* ------------------------------------------------------------------ Clusterer (8f #1) -- */

/* What Clusterer::run (clusterer.cpp:8-52) keeps of the BAM: the records that pass its filters (not unmapped /
 * secondary / supplementary, mapq >= --min-mapq, clusterer.cpp:116-122) in file order, which must be coordinate
 * order (the reference needs the .bai, i.e. a sorted BAM, for fill_clusters' region queries).  All HOST arrays. */
typedef struct {
  int64_t n_aln;
  const int32_t* tid;          /* bam1_core_t::tid, non-decreasing                                          */
  const int32_t* pos;          /* bam1_core_t::pos, non-decreasing inside a tid                              */
  const int32_t* hp;           /* HP:i aux value, 0 when absent                                              */
  const int64_t* cigar_offs;   /* n_aln + 1                                                                  */
  const uint32_t* cigar;       /* bam_get_cigar: len << 4 | op                                               */
  const int64_t* sfs_offs;     /* n_aln + 1: SFSs->at(qname) of the record's read, in .sfs order (empty if none) */
  const int32_t* sfs_qs;       /* SFS::qs                                                                    */
  const int32_t* sfs_len;      /* SFS::l                                                                     */
} svb_alns_t;

#define SVB_SEQ_ASCII 0        /* upper-case characters, as load_chromosomes keeps them (chromosomes.cpp:10-27) */
#define SVB_SEQ_NT6 1          /* nt6 codes, one byte per base                                                */
#define SVB_SEQ_BAM4 2         /* reads only: 4 bits per base as BAM stores them, every read on a byte boundary */

/* chromosome_seqs (chromosomes.hpp): contig c = seq[start[c], start[c] + len[c]).  `mem` says where seq lives;
 * start / len / name_rank are HOST arrays.  name_rank[tid] orders the chromosome NAMES (SFS::operator< compares
 * strings, sfs.hpp:64-72); NULL = tid order.  len[c] < 0: chromosome c of the BAM header has no sequence
 * (reads on it are skipped, clusterer.cpp:162-163). */
typedef struct {
  int64_t n_contigs;
  const uint8_t* seq;
  const int64_t* start;
  const int64_t* len;
  const int32_t* name_rank;
  int fmt;                     /* SVB_SEQ_ASCII or SVB_SEQ_NT6 */
  int mem;
} svb_ref_t;

typedef struct {
  /* Clusterer::clusters in the reference's order for `threads` (clusterer.cpp:33-36); a cluster with fewer than
   * min_cluster_weight reads has placed = 0 and no coordinates (the reference leaves them uninitialised) */
  int64_t n_clusters;
  int32_t* tid;                /* Cluster::chrom                                                             */
  int32_t* s;                  /* Cluster::s, ::e (clusterer.cpp:520)                                        */
  int32_t* e;
  int32_t* cov0;               /* Cluster::cov0..2; cov = their sum (clusterer.cpp:592-596)                  */
  int32_t* cov1;
  int32_t* cov2;
  uint8_t* placed;
  int64_t* sub_offs;           /* n_clusters + 1: Cluster::subreads, in BAM order                            */
  int32_t* sub_aln;            /* index into the alignments                                                  */
  int32_t* sub_qs;             /* SubRead::seq = read bases [qs, qe] (empty when qe < qs)                    */
  int32_t* sub_qe;
  int32_t* sub_hp;             /* SubRead::htag: 1, 2 or 0                                                   */
  int64_t* rvec_offs;          /* n_clusters + 1: Cluster::reads, one byte each: has-SFS | hap code << 1 (1, 2, 3 = untagged) */
  uint8_t* rvec;
  int32_t* clip;               /* with clipped: 4 per alignment (left pos, left bases, right pos, right bases; bases 0 = none), else NULL */
  /* bookkeeping of clusterer.hpp:150-160, as logged by Caller::run */
  int64_t unplaced, s_unplaced, e_unplaced, unknown, unextended, small_clusters, small_clusters_2, n_extended;
  int32_t max_ext_len, dist;
  float kernel_ms;             /* the two kernels, CUDA events                                               */
  float device_ms;             /* uploads + kernels + downloads                                              */
  float host_ms;               /* the sort + sweep between them (cluster_by_proximity)                       */
  int32_t launches;
  int64_t h2d_bytes, d2h_bytes;
} svb_clusters_t;

/* Clusterer::run (clusterer.cpp:8-52) on the GPU: extend_alignment (:156-345, one thread per read that carries
 * SFSs: placement through the CIGAR, unique 7-mers of the 100-bp flanks, per-read merge) and fill_clusters
 * (:478-610, one thread per cluster: coverage, RVEC, sub-read windows) are kernels; cluster_by_proximity (:405-475)
 * is a sort + two sequential sweeps over the extended SFSs on the host between them.  flank / ksize are
 * config.hpp:85-86 (100 / 7; ksize <= 8, flank <= 128).  `threads` only fixes the output order. */
SVB_API int svb_cluster_batch(const svb_alns_t* alns, const svb_ref_t* ref, int threads, int min_cluster_weight,
                              int flank, int ksize, int clipped, int device, svb_clusters_t* out);
SVB_API void svb_clusters_free(svb_clusters_t* out);

/* ------------------------------------------------------------------ Caller::pcall (a8, a9) -- */

/* The sequences sub-reads are cut from: sequence i = seq[offs[i] ...) -- offs in BYTES for SVB_SEQ_BAM4 (every read
 * starts on a byte boundary, l_qseq bases each, bam_get_seq), in bases for SVB_SEQ_NT6 / SVB_SEQ_ASCII; offs[i] < 0 =
 * sequence not available (a read that was never searched cannot carry an SFS, hence never yields a sub-read).
 * `mem` says where seq lives (SVB_MEM_DEVICE: SVB_SEQ_NT6 only); offs is a HOST array of n entries. */
typedef struct {
  int64_t n;
  const uint8_t* seq;
  const int64_t* offs;
  int fmt;
  int mem;
} svb_seqs_t;

typedef struct {
  /* one job = one sub-cluster that split_cluster kept (caller.cpp:100-255, at most two per cluster), cluster order */
  int64_t n_jobs;
  int32_t* job_cluster;        /* index into the clusters                                                      */
  int32_t* job_cov;            /* Cluster::cov, cov0, cov1, cov2 of the sub-cluster (-1 = not applicable), 4 per job */
  int64_t* job_sub_offs;       /* n_jobs + 1: the sub-reads of the job = indices into the clusters' sub_* arrays */
  int32_t* job_sub;
  int64_t* cons_offs;          /* n_jobs + 1: run_poa's consensus, codes 0..4 = ACGTN (caller.cpp:292-297)     */
  uint8_t* cons;
  int32_t* score;              /* ez.score (caller.cpp:351)                                                    */
  int64_t* cigar_offs;         /* n_jobs + 1: ez.cigar, len << 4 | op with op 0 1 2 = M I D                    */
  uint32_t* cigar;
  /* SV records of the CIGAR walk (caller.cpp:359-401), job order then CIGAR order */
  int64_t n_svs;
  int32_t* sv_job;
  uint8_t* sv_type;            /* 0 INS, 1 DEL                                                                 */
  int32_t* sv_pos;             /* SV::s = rpos (1-based position of the anchor base)                           */
  int32_t* sv_len;             /* SV::l                                                                        */
  int32_t* sv_cpos;            /* INS: the inserted bases are cons[cpos, cpos + l) of the job                  */
  int32_t* job_nv;             /* SV::ngaps of every record of the job                                         */
  int64_t skipped_outside;     /* clusters whose window leaves the chromosome (the reference would read out of bounds) */
  /* measurement */
  int64_t poa_cells, ksw_cells;
  float poa_kernel_ms, ksw_kernel_ms, gather_ms;
  float device_ms;             /* all device work of the call                                                  */
  float host_ms;               /* split_cluster + CIGAR walk on the host                                       */
  int32_t launches, poa_reruns, ksw_waves;
  int64_t h2d_bytes, d2h_bytes;
  float poa_ms, ksw_ms;        /* wall clock of the whole POA / ksw2 stage of the call (uploads, workspace, kernels, downloads) */
} svb_calls_t;

/* Caller::pcall (caller.cpp:311-406) for the clusters of svb_cluster_batch (or any svb_clusters_t a caller fills:
 * tid, s, e, cov0..2, placed, sub_offs, sub_aln, sub_qs, sub_qe, sub_hp): clusters with fewer than
 * min_cluster_weight sub-reads are skipped (:316), split_cluster (host: it looks at lengths and haplotype tags
 * only), then for every sub-cluster run_poa over its sub-reads (gathered on the GPU when the reads live there,
 * k_poa), ksw_extd2 of the consensus against chromosome[s, e] (k_ksw_extd2) and the CIGAR walk to INS / DEL
 * records of at least min_sv_length bases.  min_ratio is config->min_ratio (0.97), useht = !--noht. */
SVB_API int svb_call_batch(const svb_clusters_t* clusters, const svb_seqs_t* reads, const svb_ref_t* ref, int min_cluster_weight,
                           int min_sv_length, float min_ratio, int useht, int device, svb_calls_t* out);
SVB_API void svb_calls_free(svb_calls_t* out);

/* The indexed sequences as an svb_ref_t that lives on the index's device (the text the located-match mode keeps
 * next to the BWT): contig c = nt6 codes of the forward strand.  start_buf / len_buf: caller-owned HOST arrays of
 * n_contigs entries that `out` points into.  Fails for an index without text (svb_index_from_bwt). */
SVB_API int svb_index_ref(const svb_index_t* idx, svb_ref_t* out, int64_t* start_buf, int64_t* len_buf);

#ifdef __cplusplus
}
#endif
#endif /* SVDSS_B200_H */
