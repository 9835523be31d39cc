/* Plain-C oracle for the integer tail of the evaluation path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  PARITY UNPINNED: the
 * histogram restates [TF-1.12] metrics_impl._streaming_confusion_matrix reached
 * from code/estimator/define_estimator_hierarchical.py:185-194; the composition
 * restates code/models/resnet50_extended_model_hierarchical.py:113-117.
 */
#include <stdint.h>

/* cm[label * C + decision] += 1; returns the number of out-of-range pairs (skipped). */
int oracle_confusion_matrix(const int32_t* labels, const int32_t* decisions, long long n,
                            int num_classes, int64_t* cm) {
  int bad = 0;
  for (long long i = 0; i < n; ++i) {
    int32_t l = labels[i], d = decisions[i];
    if (l < 0 || l >= num_classes || d < 0 || d >= num_classes) { ++bad; continue; }
    cm[(int64_t)l * num_classes + d] += 1;
  }
  return bad;
}

/* decs = where(l1 == veh, veh_map[l2v], where(l1 == hum, hum_map[l2h], l1_map[l1])) */
void oracle_compose_decisions(const int32_t* l1, const int32_t* l2v, const int32_t* l2h, long long n,
                              int cid_vehicle, int cid_human, const int32_t* l1_map,
                              const int32_t* veh_map, const int32_t* hum_map, int32_t* out) {
  for (long long i = 0; i < n; ++i) {
    if (l1[i] == cid_vehicle) out[i] = veh_map[l2v[i]];
    else if (l1[i] == cid_human) out[i] = hum_map[l2h[i]];
    else out[i] = l1_map[l1[i]];
  }
}
