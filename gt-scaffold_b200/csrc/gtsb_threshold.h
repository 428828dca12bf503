/* gtsb_threshold.h -- see gtsb_threshold.c */
#ifndef GTSB_THRESHOLD_H
#define GTSB_THRESHOLD_H
#ifdef __cplusplus
extern "C" {
#endif
/* tail of gt_scaffolder_graph_ambiguousorder (algorithms.c:187-192) */
int gtsb_ambig_tail(float interval, float cutoff);
/* reduce it to thresholds; returns -1 if it is not a step function of
   |interval| on either side for this cutoff and libm (then the caller must
   refuse rather than guess) */
int gtsb_ambig_thresholds(float cutoff, float *t_pos, float *t_neg, int *inf_true);
#ifdef __cplusplus
}
#endif
#endif
