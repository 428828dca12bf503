// mle_emul.cpp -- TEST INFRASTRUCTURE.  gtsb_mle_host with the device scan replaced by plain loops
// over the same stage function (gt-scaffold_b200/csrc/gtsb_mle_core.h): preparation, every
// (pair, theta) slot, the ranking and tolerance rule of the kernels, the host decision.  Compared
// with the reference's own estimate_dist_using_mle (oracle/ref_bam_driver.c).
#include <math.h>
#include <stdint.h>

#include <vector>

#include "../../gt-scaffold_b200/csrc/gtsb_mle_core.h"

using namespace gtsbm;

extern "C" {

// keep_all != 0: every theta with pairs goes to the decision (no ranking filter)
int emul_mle(uint64_t nof_pairs, const uint64_t *frag_off, const int64_t *frag_start, const int64_t *frag_end,
             const uint64_t *ma, const uint64_t *len_ref, const uint64_t *len_mref, const double *pmf,
             uint64_t pmf_nof, double minp, int rf, int64_t min_dist, int64_t max_dist, int keep_all, int64_t *dist,
             uint64_t *pairs_used, uint64_t *slots_out, uint64_t *cand_out) {
  std::vector<MlePair> pairs;
  std::vector<uint64_t> size, count;
  std::vector<double> logp;
  uint64_t nslots = 0;
  if (mle_prepare(nof_pairs, frag_off, frag_start, frag_end, ma, len_ref, len_mref, pmf, pmf_nof, minp, rf, min_dist,
                  max_dist, pairs, size, count, logp, &nslots) != nullptr)
    return -1;
  std::vector<MleCandidate> cand;
  for (uint64_t p = 0; p < nof_pairs; p++) {
    const MlePair &q = pairs[p];
    if (q.lo > q.hi) continue;
    const uint64_t len = (uint64_t) (q.hi - q.lo) + 1u;
    std::vector<double> L(len), c(len), g(len);
    std::vector<uint64_t> n(len);
    double gmax = -INFINITY, mag = 0.0;
    for (uint64_t k = 0; k < len; k++) {
      mle_eval(q, q.lo + (int64_t) k, size.data(), count.data(), pmf, logp.data(), pmf_nof, minp, &L[k], &n[k], &c[k]);
      // a different logarithm than the decision's, as on the device: log2 * ln 2
      g[k] = L[k] - (double) q.nfp * (log2(c[k]) * 0.6931471805599453);
      if (n[k] == 0) continue;
      const double a = fabs(L[k]) + fabs(g[k] - L[k]);
      if (g[k] == g[k] && g[k] > gmax) gmax = g[k];
      if (a == a && a > mag && !isinf(a)) mag = a;
    }
    const double tol = 1e-9 * (1.0 + mag);
    for (uint64_t k = 0; k < len; k++) {
      if (n[k] == 0) continue;
      if (!keep_all && g[k] < gmax - tol) continue;
      MleCandidate x;
      x.pair = (uint32_t) p;
      x.pad = 0;
      x.theta = q.lo + (int64_t) k;
      x.L = L[k];
      x.c = c[k];
      x.n = n[k];
      cand.push_back(x);
    }
  }
  if (slots_out) *slots_out = nslots;
  if (cand_out) *cand_out = cand.size();
  mle_decide(cand, pairs, ma, rf, min_dist, dist, pairs_used);
  return 0;
}

}  // extern "C"
