// gtsb_mle.cu -- batched distance MLE between contig pairs (SURVEY.md §8(f) rank 4;
// gt_scaffolder_bamparser.c:385-598: window, calculate_fragment_dist, compute_likelihood,
// maximum_likelihood_estimate, estimate_dist_using_mle).  One thread per (contig pair, theta)
// evaluates the likelihood sum, the pair count and the normalising constant exactly
// (gtsb_mle_core.h); the host takes the two logarithms with its own libm and decides.
#include <math.h>

#include <algorithm>
#include <vector>

#include "gtsb_context.h"
#include "gtsb_mle_core.h"

using namespace gtsb;
using namespace gtsbi;
using namespace gtsbm;

namespace gtsbmle {

__global__ void __launch_bounds__(128) km_eval(uint32_t npairs, uint64_t nslots, const MlePair *__restrict__ pairs,
                                               const uint64_t *__restrict__ size, const uint64_t *__restrict__ count,
                                               const double *__restrict__ pmf, const double *__restrict__ logp,
                                               uint64_t pmf_nof, double minp, double *__restrict__ L,
                                               double *__restrict__ c, uint64_t *__restrict__ n,
                                               double *__restrict__ g, uint32_t *__restrict__ slot_pair) {
  const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots) return;
  uint32_t lo = 0, hi = npairs;                    // the pair of slot t: last p with out_off[p] <= t
  while (hi - lo > 1u) {
    const uint32_t mid = lo + (hi - lo) / 2u;
    if (pairs[mid].out_off <= t) lo = mid; else hi = mid;
  }
  const MlePair p = pairs[lo];
  const int64_t theta = p.lo + (int64_t) (t - p.out_off);
  double Lt, ct;
  uint64_t nt;
  mle_eval(p, theta, size, count, pmf, logp, pmf_nof, minp, &Lt, &nt, &ct);
  L[t] = Lt;
  c[t] = ct;
  n[t] = nt;
  g[t] = Lt - (double) p.nfp * log(ct);            // the device's own logarithm: ranking only
  slot_pair[t] = lo;
}

// warp per pair: the best ranking value among the thetas with pairs, and the size of the terms
__global__ void __launch_bounds__(256) km_best(uint32_t npairs, const MlePair *__restrict__ pairs,
                                               const double *__restrict__ L, const double *__restrict__ c,
                                               const uint64_t *__restrict__ n, const double *__restrict__ g,
                                               double *__restrict__ gmax, double *__restrict__ mag) {
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31u;
  if (w >= npairs) return;
  const MlePair p = pairs[w];
  double best = -INFINITY, m = 0.0;
  if (p.lo <= p.hi) {
    const uint64_t len = (uint64_t) (p.hi - p.lo) + 1u;
    for (uint64_t k = lane; k < len; k += 32u) {
      const uint64_t t = p.out_off + k;
      if (n[t] == 0) continue;
      const double v = g[t], a = fabs(L[t]) + fabs(g[t] - L[t]);
      if (v == v && v > best) best = v;
      if (a == a && a > m && !isinf(a)) m = a;
    }
  }
  for (int d = 16; d >= 1; d >>= 1) {
    best = fmax(best, __shfl_xor_sync(0xffffffffu, best, d));
    m = fmax(m, __shfl_xor_sync(0xffffffffu, m, d));
  }
  if (lane == 0) {
    gmax[w] = best;
    mag[w] = m;
  }
}

__global__ void __launch_bounds__(256) km_candidates(uint64_t nslots, const MlePair *__restrict__ pairs,
                                                     const uint32_t *__restrict__ slot_pair,
                                                     const double *__restrict__ L, const double *__restrict__ c,
                                                     const uint64_t *__restrict__ n, const double *__restrict__ g,
                                                     const double *__restrict__ gmax, const double *__restrict__ mag,
                                                     MleCandidate *__restrict__ out, uint32_t cap,
                                                     uint32_t *__restrict__ n_out) {
  const uint64_t t = (uint64_t) blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nslots || n[t] == 0) return;
  const uint32_t p = slot_pair[t];
  const double tol = 1e-9 * (1.0 + mag[p]);
  if (g[t] < gmax[p] - tol) return;                // NaN and infinities stay in
  const uint32_t at = atomicAdd(n_out, 1u);
  if (at >= cap) return;
  MleCandidate x;
  x.pair = p;
  x.pad = 0;
  x.theta = pairs[p].lo + (int64_t) (t - pairs[p].out_off);
  x.L = L[t];
  x.c = c[t];
  x.n = n[t];
  out[at] = x;
}

}  // namespace gtsbmle

using namespace gtsbmle;

extern "C" {

int gtsb_mle_host(gtsb_context *c, uint64_t nof_pairs, const uint64_t *frag_off, const int64_t *frag_start,
                  const int64_t *frag_end, const uint64_t *ma, const uint64_t *len_ref, const uint64_t *len_mref,
                  const double *pmf, uint64_t pmf_nof, double minp, int rf, int64_t min_dist, int64_t max_dist,
                  int64_t *dist, uint64_t *pairs_used) {
  if (c == nullptr) return -1;
  CK(cudaSetDevice(c->device));
  if (nof_pairs == 0) return 0;
  if (frag_off == nullptr || frag_start == nullptr || frag_end == nullptr || ma == nullptr || len_ref == nullptr ||
      len_mref == nullptr || pmf == nullptr || dist == nullptr || pairs_used == nullptr)
    return fail(c, "gtsb_mle_host: null argument");
  if (nof_pairs >= 0xFFFFFFF0ull) return fail(c, "gtsb_mle_host: too many contig pairs in one call");
  if (pmf_nof == 0) return fail(c, "gtsb_mle_host: empty distribution");
  ProfScope prof(c);
  cudaStream_t s = c->stream;
  // ---- host: fragment-size tables, theta ranges, log(p) table (gtsb_mle_core.h)
  std::vector<MlePair> pairs;
  std::vector<uint64_t> size, count;
  std::vector<double> logp;
  uint64_t nslots = 0;
  const char *msg = mle_prepare(nof_pairs, frag_off, frag_start, frag_end, ma, len_ref, len_mref, pmf, pmf_nof, minp, rf,
                                min_dist, max_dist, pairs, size, count, logp, &nslots);
  if (msg != nullptr) return fail(c, "%s", msg);

  // ---- device: every (pair, theta)
  std::vector<MleCandidate> cand;
  uint32_t ncand = 0;
  if (nslots) {
    if (nslots >= (1ull << 32)) return fail(c, "gtsb_mle_host: too many (pair, theta) slots for one call, pass fewer pairs");
    ENSURE(c->m_pairs, nof_pairs * sizeof(MlePair));
    ENSURE(c->m_size, size.size() * 8 + 8);
    ENSURE(c->m_count, count.size() * 8 + 8);
    ENSURE(c->m_pmf, pmf_nof * 8);
    ENSURE(c->m_logp, (pmf_nof + 1) * 8);
    ENSURE(c->m_L, nslots * 8);
    ENSURE(c->m_c, nslots * 8);
    ENSURE(c->m_n, nslots * 8);
    ENSURE(c->m_g, nslots * 8);
    ENSURE(c->m_slot_pair, nslots * 4);
    ENSURE(c->m_gmax, nof_pairs * 8);
    ENSURE(c->m_mag, nof_pairs * 8);
    ENSURE(c->p_flags, 16);
    uint32_t cap = (uint32_t) std::min<uint64_t>(nslots, std::max<uint64_t>(4 * nof_pairs + 1024, 65536));
    CK(cudaMemcpyAsync(c->m_pairs.p, pairs.data(), nof_pairs * sizeof(MlePair), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->m_size.p, size.data(), size.size() * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->m_count.p, count.data(), count.size() * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->m_pmf.p, pmf, pmf_nof * 8, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(c->m_logp.p, logp.data(), (pmf_nof + 1) * 8, cudaMemcpyHostToDevice, s));
    {
      GTSB_TIMED("km_eval", s);
      km_eval<<<(uint32_t) ((nslots + 127) / 128), 128, 0, s>>>(
          (uint32_t) nof_pairs, nslots, c->m_pairs.as<MlePair>(), c->m_size.as<uint64_t>(), c->m_count.as<uint64_t>(),
          c->m_pmf.as<double>(), c->m_logp.as<double>(), pmf_nof, minp, c->m_L.as<double>(), c->m_c.as<double>(),
          c->m_n.as<uint64_t>(), c->m_g.as<double>(), c->m_slot_pair.as<uint32_t>());
    }
    {
      GTSB_TIMED("km_best", s);
      km_best<<<(uint32_t) ((nof_pairs * 32 + 255) / 256), 256, 0, s>>>(
          (uint32_t) nof_pairs, c->m_pairs.as<MlePair>(), c->m_L.as<double>(), c->m_c.as<double>(),
          c->m_n.as<uint64_t>(), c->m_g.as<double>(), c->m_gmax.as<double>(), c->m_mag.as<double>());
    }
    c->stats.kernel_launches += 2;
    for (;;) {                                     // the candidate list grows when a plateau overflows it
      ENSURE(c->m_cand, (size_t) cap * sizeof(MleCandidate));
      CK(cudaMemsetAsync(c->p_flags.p, 0, 16, s));
      {
        GTSB_TIMED("km_candidates", s);
        km_candidates<<<(uint32_t) ((nslots + 255) / 256), 256, 0, s>>>(
            nslots, c->m_pairs.as<MlePair>(), c->m_slot_pair.as<uint32_t>(), c->m_L.as<double>(),
            c->m_c.as<double>(), c->m_n.as<uint64_t>(), c->m_g.as<double>(), c->m_gmax.as<double>(),
            c->m_mag.as<double>(), c->m_cand.as<MleCandidate>(), cap, c->p_flags.as<uint32_t>());
      }
      c->stats.kernel_launches++;
      CK(cudaMemcpyAsync(&ncand, c->p_flags.p, 4, cudaMemcpyDeviceToHost, s));
      CK(cudaStreamSynchronize(s));
      CK(cudaGetLastError());
      if (ncand <= cap) break;
      cap = ncand;
    }
    cand.resize(ncand);
    if (ncand) CK(cudaMemcpyAsync(cand.data(), c->m_cand.p, (size_t) ncand * sizeof(MleCandidate), cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }

  // ---- host: the reference's decision on the candidates
  mle_decide(cand, pairs, ma, rf, min_dist, dist, pairs_used);
  return 0;
}

}  // extern "C"
