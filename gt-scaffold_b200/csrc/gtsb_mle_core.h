// gtsb_mle_core.h -- the distance estimator of gt_scaffolder_bamparser.c (:385-598) as stage
// functions shared by the kernels of gtsb_mle.cu and their host build (tests/emul/mle_emul.cpp).
//
// For one contig pair the reference scans theta = min_dist .. max_dist and keeps the first theta
// with the largest
//     likelihood(theta) = sum_i count_i * log(p(size_i + theta)) - nof_frag_pos * log(c(theta))
//     c(theta)          = sum_{x < pmf.nof} pmf[x] * window(len_ref, len_mref, x - theta)
// among the thetas with n(theta) > 0 fragments inside the distribution (:496-551).  All of it is
// double arithmetic summed in index order, which a thread reproduces bit for bit (IEEE add / mul /
// div, no FMA contraction) -- except the two logarithms, which come from the C library.
// log(p) is a table the host fills once per distribution with its own libm.  log(c) is taken on
// the host too: the device evaluates L(theta) = sum_i count_i * logp[...], c(theta) and n(theta)
// exactly, ranks the thetas with its own logarithm, and hands back only those within a tolerance
// (>= 1e5 times the error of that ranking) of the best; the host finishes these few with libm and
// applies the reference's "first strictly larger" rule.  The scan -- O(range * (pmf.nof + sizes))
// per pair -- stays on the device, the decision is the reference's own arithmetic.
#pragma once
#include <stdint.h>

#ifndef GTSB_HD
#if defined(__CUDACC__)
#define GTSB_HD __host__ __device__ __forceinline__
#else
#define GTSB_HD inline
#endif
#endif

namespace gtsbm {

struct MlePair {            // one contig pair, as maximum_likelihood_estimate sees it
  uint64_t tab_off;         // its (size, count) table in the size / count arrays
  uint32_t tab_n;
  uint32_t pad;
  int64_t lo, hi;           // theta range after the clamps of :511-512 (lo > hi: nothing to scan)
  uint64_t out_off;         // first slot of its thetas in the per-theta arrays
  int64_t x1, x2;           // len_ref, len_mref after estimate_dist_using_mle's adjustments (:567-575)
  uint64_t nfp;             // FragmentData.nof_frag_pos
};

// window(), bamparser.c:385-403
GTSB_HD double mle_window(int64_t x1, int64_t x2, int64_t x) {
  int64_t r;
  const int64_t x3 = x1 + x2;
  if (x <= 0) r = 1;
  else if (x < x1) r = x;
  else if (x < x2) r = x1;
  else if (x < x3) r = x3 - x;
  else r = 1;
  return (double) r / (double) x1;
}

// one theta of one pair: L = the sum of compute_likelihood (:459-494) with log(p) from the table
// (logp[pmf_nof] = log(minp)), n = its pair count, c = the normalising constant (:518-528)
GTSB_HD void mle_eval(const MlePair &p, int64_t theta, const uint64_t *size, const uint64_t *count,
                      const double *pmf, const double *logp, uint64_t pmf_nof, double minp, double *L_out,
                      uint64_t *n_out, double *c_out) {
  double c = 0.0;
  for (uint64_t x = 0; x < pmf_nof; x++) {
    const double w = mle_window(p.x1, p.x2, (int64_t) (x - (uint64_t) theta));
#if defined(__CUDA_ARCH__)
    c = __dadd_rn(c, __dmul_rn(pmf[x], w));
#else
    c += pmf[x] * w;
#endif
  }
  double L = 0.0;
  uint64_t n = 0;
  for (uint32_t i = 0; i < p.tab_n; i++) {
    const int64_t fs = (int64_t) (size[p.tab_off + i] + (uint64_t) theta);
    const bool inside = fs >= 0 && (uint64_t) fs < pmf_nof;
    const double prob = inside ? pmf[fs] : minp;
    const double lp = inside ? logp[fs] : logp[pmf_nof];
    const double cnt = (double) count[p.tab_off + i];
#if defined(__CUDA_ARCH__)
    L = __dadd_rn(L, __dmul_rn(cnt, lp));
#else
    L += cnt * lp;
#endif
    if (prob > minp) n += count[p.tab_off + i];
  }
  *L_out = L;
  *n_out = n;
  *c_out = c;
}

struct MleCandidate {
  uint32_t pair;
  uint32_t pad;
  int64_t theta;
  double L, c;
  uint64_t n;
};

}  // namespace gtsbm
#include <math.h>

#include <algorithm>
#include <vector>
namespace gtsbm {

// Host half, before the scan: the (size, count) tables of calculate_fragment_dist (:423-456), the
// adjusted contig lengths (:567-575), the theta ranges (:511-512) and the log(p) table.
// Returns nullptr, or a message.
inline const char *mle_prepare(uint64_t nof_pairs, const uint64_t *frag_off, const int64_t *frag_start,
                               const int64_t *frag_end, const uint64_t *ma, const uint64_t *len_ref,
                               const uint64_t *len_mref, const double *pmf, uint64_t pmf_nof, double minp, int rf,
                               int64_t min_dist, int64_t max_dist, std::vector<MlePair> &pairs,
                               std::vector<uint64_t> &size, std::vector<uint64_t> &count, std::vector<double> &logp,
                               uint64_t *nslots_out) {
  // compute_likelihood stops with "negative probability" when it meets one (:484-488); the whole
  // distribution is checked here instead
  if (minp < 0) return "negative probability";
  for (uint64_t x = 0; x < pmf_nof; x++)
    if (pmf[x] < 0) return "negative probability";
  pairs.resize(nof_pairs);
  size.clear();
  count.clear();
  std::vector<uint64_t> tmp;
  uint64_t nslots = 0;
  for (uint64_t p = 0; p < nof_pairs; p++) {
    const uint64_t f0 = frag_off[p], f1 = frag_off[p + 1];
    if (f1 <= f0) return "a contig pair without fragments";
    MlePair &q = pairs[p];
    uint64_t lr = len_ref[p] - (ma[p] - 1), lm = len_mref[p] - (ma[p] - 1);      // :567-568 (GtUword arithmetic)
    if (lr > lm) std::swap(lr, lm);                                              // :571-575
    const uint64_t factor = rf ? 0 : 2 * (ma[p] - 1);                            // :579, :586
    tmp.resize(f1 - f0);
    for (uint64_t k = f0; k < f1; k++) tmp[k - f0] = (uint64_t) frag_end[k] - (uint64_t) frag_start[k];
    std::sort(tmp.begin(), tmp.end());                                           // compare_fragments: unsigned sizes
    q.tab_off = size.size();
    for (uint64_t k = 0; k < tmp.size(); k++) {
      if (k == 0 || tmp[k] != tmp[k - 1]) {
        size.push_back(tmp[k] - factor);
        count.push_back(1);
      } else {
        count.back()++;
      }
    }
    q.tab_n = (uint32_t) (size.size() - q.tab_off);
    q.pad = 0;
    const uint64_t min_frag = size[q.tab_off], max_frag = size.back();
    // :511-512: MAX / MIN of a GtWord and a GtUword compare as unsigned
    const uint64_t a_lo = (uint64_t) min_dist, b_lo = 0 - min_frag;
    q.lo = (int64_t) (a_lo > b_lo ? a_lo : b_lo);
    const uint64_t a_hi = (uint64_t) max_dist, b_hi = pmf_nof - max_frag - 1;
    q.hi = (int64_t) (a_hi < b_hi ? a_hi : b_hi);
    q.out_off = nslots;
    q.x1 = (int64_t) lr;
    q.x2 = (int64_t) lm;
    q.nfp = f1 - f0;
    if (q.lo <= q.hi) {
      const uint64_t len = (uint64_t) (q.hi - q.lo) + 1u;
      if (len > (1ull << 31)) return "distance range of a contig pair too long";
      nslots += len;
    }
  }
  logp.resize(pmf_nof + 1);
  for (uint64_t x = 0; x < pmf_nof; x++) logp[x] = log(pmf[x]);
  logp[pmf_nof] = log(minp);
  *nslots_out = nslots;
  return nullptr;
}

// Host half, after the scan: the reference's decision (:530-546, :578-590) on the thetas the scan
// kept, with the C library's log(c)
inline void mle_decide(std::vector<MleCandidate> &cand, const std::vector<MlePair> &pairs, const uint64_t *ma, int rf,
                       int64_t min_dist, int64_t *dist, uint64_t *pairs_used) {
  const uint64_t nof_pairs = pairs.size();
  std::sort(cand.begin(), cand.end(), [](const MleCandidate &a, const MleCandidate &b) {
    return a.pair != b.pair ? a.pair < b.pair : a.theta < b.theta;
  });
  std::vector<int64_t> best_theta(nof_pairs, min_dist);
  std::vector<uint64_t> best_n(nof_pairs, 0);
  std::vector<double> best_lik(nof_pairs, (double) INT64_MIN);
  for (const MleCandidate &x : cand) {
    const double lik = x.L - (double) pairs[x.pair].nfp * log(x.c);
    if (x.n > 0 && lik > best_lik[x.pair]) {
      best_lik[x.pair] = lik;
      best_theta[x.pair] = x.theta;
      best_n[x.pair] = x.n;
    }
  }
  for (uint64_t p = 0; p < nof_pairs; p++) {
    pairs_used[p] = best_n[p];
    if (rf) {
      dist[p] = best_theta[p];
    } else {                                       // :589: forward-reverse libraries
      const int64_t d = best_theta[p] - 2 * (int64_t) (ma[p] - 1);
      dist[p] = min_dist > d ? min_dist : d;
    }
  }
}

}  // namespace gtsbm
