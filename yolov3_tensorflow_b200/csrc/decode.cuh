// Arithmetic of the anchor-box decode (model.py:82-137, 140-190), shared by predict_kernel (csrc/decode.cu) and the
// detection-head epilogue of the conv kernels (csrc/conv_igemm.cu) so that the fused and the stand-alone path produce
// the same bits.  __f*_rn intrinsics keep the reference's op order free of FMA contraction.
#pragma once
#include "common.cuh"

namespace yb {

__device__ __forceinline__ float sigmoid_ref(float x) { return 1.0f / (1.0f + expf(-x)); }

// one axis of a box: centre and size in input-image pixels -> (min, max) corner coordinates
//   centre = (sigmoid(t_c) + grid offset) * ratio            model.py:118-120
//   size   = exp(t_s) * (anchor / ratio) * ratio             model.py:94,123-126
//   corner = centre -/+ size / 2                             model.py:182-188
__device__ __forceinline__ void decode_axis(float t_c, float t_s, float off, float ratio, float anchor, float& lo,
                                            float& hi) {
  const float center = __fmul_rn(__fadd_rn(sigmoid_ref(t_c), off), ratio);
  const float size = __fmul_rn(__fmul_rn(expf(t_s), __fdiv_rn(anchor, ratio)), ratio);
  const float half = __fmul_rn(size, 0.5f);
  lo = __fsub_rn(center, half);
  hi = __fadd_rn(center, half);
}

// Decode fused into the detection-head conv (yb_net_detect): what the epilogue needs to turn one accumulator row
// (one grid cell: 3 anchors x (5 + C) logits) into 3 boxes and the (score >= thresh) candidates of the NMS.
struct DetParams {
  float* boxes;            // [n, B, 4] xmin, ymin, xmax, ymax
  int* cand_count;         // [n, C]       NMS workspace (csrc/nms.cu), zeroed by the caller
  float* cand_score;       // [n, C, B]
  int* cand_idx;           // [n, C, B]
  int B, C, E;             // boxes per image over the three scales, classes, 5 + C
  int box_off;             // first box of this scale inside an image
  float ratio_w, ratio_h;  // input pixels per grid cell
  float anchor_w[3], anchor_h[3];
  float thr;               // score threshold (score >= thr is a candidate, utils/nms_utils.py:30)
  float logit_lo;          // logits below this cannot reach thr (sigmoid(logit_lo) < thr with margin); -inf: no pre-filter
  int on;                  // 0: plain conv
};

}  // namespace yb
