/* CPU ORACLE — TEST INFRASTRUCTURE ONLY (never linked into the product library).
 * C restatement of utils/nms_utils.py:8-48 (gpu_nms) on top of TensorFlow's
 * NonMaxSuppression CPU kernel semantics (TensorFlow itself is an
 * un-vendored dependency: parity unpinned for the kernel, pinned for the Python layer by
 * tests/golden/nms.npz).  Same arithmetic as oracle/yolov3_oracle.py:tf_nms_cpu, which it is
 * checked against in tests/test_oracle_golden.py; exists because the numpy version is too slow
 * for the 100k x 80 stress configuration and for an honest single-thread CPU baseline
 * (TF's kernel is single-threaded C++ as well).
 * Build: make -C oracle   ->  oracle/libnms_oracle.so */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float score; int32_t idx; } cand_t;

static int cmp_cand(const void* a, const void* b) {
  const cand_t* x = (const cand_t*)a; const cand_t* y = (const cand_t*)b;
  if (x->score > y->score) return -1;
  if (x->score < y->score) return 1;
  return (x->idx > y->idx) - (x->idx < y->idx);          /* ties -> lower index first */
}

static inline float fminf_(float a, float b) { return a < b ? a : b; }
static inline float fmaxf_(float a, float b) { return a > b ? a : b; }

/* boxes [n,4] (any corner order), scores [n]; returns #selected, indices in selection order */
int yo_tf_nms(const float* boxes, const float* scores, int n, int max_out, float iou_thr, int32_t* out_idx) {
  if (n <= 0 || max_out <= 0) return 0;
  cand_t* c = (cand_t*)malloc(sizeof(cand_t) * (size_t)n);
  for (int i = 0; i < n; ++i) { c[i].score = scores[i] + 0.0f; c[i].idx = i; }
  qsort(c, (size_t)n, sizeof(cand_t), cmp_cand);
  float* sel = (float*)malloc(sizeof(float) * 5 * (size_t)max_out);   /* x0,y0,x1,y1,area */
  int k = 0;
  for (int t = 0; t < n && k < max_out; ++t) {
    const float* b = boxes + 4 * (size_t)c[t].idx;
    const float x0 = fminf_(b[0], b[2]), x1 = fmaxf_(b[0], b[2]);
    const float y0 = fminf_(b[1], b[3]), y1 = fmaxf_(b[1], b[3]);
    volatile float w = x1 - x0, h = y1 - y0;              /* volatile: no FMA contraction / excess precision */
    volatile float area = w * h;
    int keep = 1;
    for (int j = k - 1; j >= 0; --j) {                    /* most recently selected first, like TF */
      const float* s = sel + 5 * j;
      if (s[4] <= 0.0f || area <= 0.0f) continue;
      volatile float iw = fminf_(s[2], x1) - fmaxf_(s[0], x0);
      volatile float ih = fminf_(s[3], y1) - fmaxf_(s[1], y0);
      if (iw < 0.0f) iw = 0.0f;
      if (ih < 0.0f) ih = 0.0f;
      volatile float inter = iw * ih;
      volatile float sum = s[4] + area;
      volatile float den = sum - inter;
      volatile float iou = inter / den;
      if (iou > iou_thr) { keep = 0; break; }
    }
    if (keep) {
      float* s = sel + 5 * k;
      s[0] = x0; s[1] = y0; s[2] = x1; s[3] = y1; s[4] = area;
      out_idx[k++] = c[t].idx;
    }
  }
  free(sel); free(c);
  return k;
}

/* utils/nms_utils.py:8-48 for one image.  boxes [B,4], scores [B,C].  Outputs sized C*max_boxes.
 * Returns K.  out_index = index of the kept box in the original [B] axis. */
int yo_gpu_nms(const float* boxes, const float* scores, int B, int C, int max_boxes, float score_thr, float iou_thr,
               float* out_boxes, float* out_scores, int32_t* out_labels, int32_t* out_index) {
  float* fb = (float*)malloc(sizeof(float) * 4 * (size_t)(B > 0 ? B : 1));
  float* fs = (float*)malloc(sizeof(float) * (size_t)(B > 0 ? B : 1));
  int32_t* pos = (int32_t*)malloc(sizeof(int32_t) * (size_t)(B > 0 ? B : 1));
  int32_t* idx = (int32_t*)malloc(sizeof(int32_t) * (size_t)(max_boxes > 0 ? max_boxes : 1));
  int K = 0;
  for (int c = 0; c < C; ++c) {
    int m = 0;
    for (int i = 0; i < B; ++i) {
      const float s = scores[(size_t)i * C + c];
      if (s >= score_thr) { memcpy(fb + 4 * (size_t)m, boxes + 4 * (size_t)i, 16); fs[m] = s; pos[m] = i; ++m; }
    }
    const int k = yo_tf_nms(fb, fs, m, max_boxes, iou_thr, idx);
    for (int j = 0; j < k; ++j) {
      memcpy(out_boxes + 4 * (size_t)K, fb + 4 * (size_t)idx[j], 16);
      out_scores[K] = fs[idx[j]]; out_labels[K] = c; out_index[K] = pos[idx[j]];
      ++K;
    }
  }
  free(fb); free(fs); free(pos); free(idx);
  return K;
}
