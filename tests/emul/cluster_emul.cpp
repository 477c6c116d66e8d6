// CPU harness for the Clusterer kernels: the per-item functions of svdss_b200/csrc/cluster_core.cuh and the host
// half of cluster_host.hpp, compiled as they are, with the two kernels of cluster.cu replaced by plain loops over the
// same functions.  Same signature and output as svb_cluster_batch (reference in HOST memory), so one Python wrapper
// serves both and the CPU suite can hold the GPU path's logic against tests/cluster_model.py without a GPU.
#include <chrono>
#include <cstdio>
#include <vector>

#include "../../svdss_b200/csrc/cluster_host.hpp"

using namespace svb;

extern "C" int emul_cluster_batch(const svb_alns_t* A, const svb_ref_t* R, int threads, int min_cluster_weight, int flank, int ksize,
                                  int clipped, svb_clusters_t* out) {
  memset(out, 0, sizeof(*out));
  const int64_t n = A->n_aln;
  std::vector<int32_t> accepted;
  for (int64_t a = 0; a < n; ++a) if (A->sfs_offs[a + 1] > A->sfs_offs[a]) accepted.push_back((int32_t)a);
  const int64_t n_sfs = n ? A->sfs_offs[n] : 0;
  std::vector<int32_t> endp((size_t)n), n_ext(accepted.size());
  std::vector<ClExt> ext((size_t)n_sfs + 1);
  unsigned cnt[8] = {0};
  if (clipped) out->clip = cl_host_alloc<int32_t>((size_t)n * 4);
  int max_span = 0;
  for (int64_t a = 0; a < n; ++a) {
    endp[(size_t)a] = cl_endpos(A->cigar + A->cigar_offs[a], (int)(A->cigar_offs[a + 1] - A->cigar_offs[a]), A->pos[a]);
    if (endp[(size_t)a] - A->pos[a] > max_span) max_span = endp[(size_t)a] - A->pos[a];
  }
  for (size_t i = 0; i < accepted.size(); ++i) {      // k_cl_extend
    const int a = accepted[i], t = A->tid[a];
    int cl4[4] = {0, 0, 0, 0};
    int m = 0;
    if (t >= 0 && t < R->n_contigs && R->len[t] >= 0) {
      ClAln al;
      al.cig = A->cigar + A->cigar_offs[a]; al.n_cig = (int)(A->cigar_offs[a + 1] - A->cigar_offs[a]); al.pos = A->pos[a];
      al.chrom = R->seq + R->start[t]; al.chrom_len = R->len[t];
      const int64_t s0 = A->sfs_offs[a];
      m = cl_extend_read(al, A->sfs_qs + s0, A->sfs_len + s0, (int)(A->sfs_offs[a + 1] - s0), flank, ksize, clipped != 0, endp[(size_t)a],
                         ext.data() + s0, cnt, cl4);
    }
    n_ext[i] = m;
    if (clipped) for (int k = 0; k < 4; ++k) out->clip[(int64_t)a * 4 + k] = cl4[k];
  }
  const auto t0_ = std::chrono::steady_clock::now();
  ClPlan P;
  cl_plan_fill(accepted.data(), (int)accepted.size(), n_ext.data(), ext.data(), A->sfs_offs, A->tid, A->pos, max_span, n, R->name_rank,
               threads, min_cluster_weight, P);
  if (getenv("SVB_EMUL_TIMING")) fprintf(stderr, "[cluster_emul] cl_plan_fill %.2f ms\n", std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0_).count());
  const int nf = (int)P.f_cluster.size();
  std::vector<int32_t> sa(P.f_members.size() + 1), sq(P.f_members.size() + 1), se(P.f_members.size() + 1), sh(P.f_members.size() + 1);
  std::vector<int32_t> nsub((size_t)nf + 1), nrv((size_t)nf + 1), cov((size_t)nf * 3 + 3);
  std::vector<uint8_t> rvec((size_t)P.f_rvoff.back() + 1);
  for (int c = 0; c < nf; ++c) {                      // k_cl_fill
    const int64_t m0 = P.f_moff[(size_t)c];
    int ns, nr, cv[3];
    unsigned unext = 0;
    cl_fill_cluster(A->pos, endp.data(), A->hp, A->cigar_offs, A->cigar, P.f_lo[(size_t)c], P.f_hi[(size_t)c], P.f_min_s[(size_t)c], P.f_max_e[(size_t)c],
                    P.f_members.data() + m0, (int)(P.f_moff[(size_t)c + 1] - m0), sa.data() + m0, sq.data() + m0, se.data() + m0, sh.data() + m0, ns,
                    rvec.data() + P.f_rvoff[(size_t)c], nr, cv, unext);
    nsub[(size_t)c] = ns; nrv[(size_t)c] = nr; cov[(size_t)c * 3] = cv[0]; cov[(size_t)c * 3 + 1] = cv[1]; cov[(size_t)c * 3 + 2] = cv[2];
    cnt[4] += unext;
  }
  out->unplaced = cnt[0]; out->s_unplaced = cnt[1]; out->e_unplaced = cnt[2]; out->unknown = cnt[3]; out->unextended = cnt[4];
  return cl_assemble(P, nsub.data(), sa.data(), sq.data(), se.data(), sh.data(), nrv.data(), rvec.data(), cov.data(), min_cluster_weight, out) ? 0 : -12;
}

extern "C" void emul_clusters_free(svb_clusters_t* o) {
  free(o->tid); free(o->s); free(o->e); free(o->cov0); free(o->cov1); free(o->cov2); free(o->placed); free(o->sub_offs);
  free(o->sub_aln); free(o->sub_qs); free(o->sub_qe); free(o->sub_hp); free(o->rvec_offs); free(o->rvec); free(o->clip);
  memset(o, 0, sizeof(*o));
}
