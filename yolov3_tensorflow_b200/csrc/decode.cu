// Anchor-box decode.  One launch covers the three scales; one thread per output element
// so that both the logit reads ([n, g, g, 3*(5+C)] fp32 is a dense [boxes, 5+C] matrix)
// and the prob/score writes are fully coalesced.  HBM-bound: 904,995 logits in,
// 10,647 x (4+1+C[+C]) floats out per 416x416 image.
// Replaces model.py:82-137 (reorg_layer), :140-190 (predict) and the caller's
// pred_scores = pred_confs * pred_probs (test_single_image.py:55).
#include "common.cuh"
#include "decode.cuh"

namespace yb {

struct DecodeParams {
  const float* fm[3];
  int gh[3], gw[3];
  int box_off[4];        // prefix of boxes per image over the scales: 0, B1, B1+B2, B
  float ratio_h[3], ratio_w[3];
  float anchor_w[9], anchor_h[9];   // pixels, reference order (small -> large)
  int n, C, E, B;
  float* boxes;          // [n,B,4]
  float* confs;          // [n,B,1]
  float* probs;          // [n,B,C]
  float* scores;         // [n,B,C] or null
};

__device__ __forceinline__ float sigmoidf_(float x) { return sigmoid_ref(x); }

// one WARP per box: the 5+C logits of a box are contiguous (coalesced 128-byte reads), the box logits are
// broadcast by shuffle, and probs / scores rows are written coalesced.  confs / probs / scores are optional
// (the detection pipeline needs only boxes + scores).
__global__ void __launch_bounds__(256) predict_kernel(const DecodeParams p) {
  const int lane = threadIdx.x & 31;
  const long nbox = (long)p.n * p.B;
  for (long gb = (long)blockIdx.x * 8 + (threadIdx.x >> 5); gb < nbox; gb += (long)gridDim.x * 8) {
    const int img = (int)(gb / p.B);
    const int b = (int)(gb - (long)img * p.B);
    const int s = b < p.box_off[1] ? 0 : (b < p.box_off[2] ? 1 : 2);
    const int lb = b - p.box_off[s];
    const int per_img = p.box_off[s + 1] - p.box_off[s];
    const float* row = p.fm[s] + ((long)img * per_img + lb) * p.E;
    const float head = lane < 5 ? row[lane] : 0.f;
    const float conf = sigmoidf_(__shfl_sync(0xffffffffu, head, 4));
    if (lane < 4) {
      const int a = lb % 3;
      const int cell = lb / 3;
      const int axis = lane & 1;  // 0: x / width, 1: y / height
      const float t_c = __shfl_sync(0xfu, head, axis);
      const float t_s = __shfl_sync(0xfu, head, 2 + axis);
      const float off = axis == 0 ? (float)(cell % p.gw[s]) : (float)(cell / p.gw[s]);
      const float ratio = axis == 0 ? p.ratio_w[s] : p.ratio_h[s];
      const int ai = (2 - s) * 3 + a;                                     // anchor groups 6:9, 3:6, 0:3
      const float anc = axis == 0 ? p.anchor_w[ai] : p.anchor_h[ai];
      float lo, hi;
      decode_axis(t_c, t_s, off, ratio, anc, lo, hi);                      // model.py:118-126, 182-188
      p.boxes[gb * 4 + lane] = lane < 2 ? lo : hi;
    }
    if (lane == 0 && p.confs) p.confs[gb] = conf;                          // model.py:167
    for (int k = lane; k < p.C; k += 32) {
      const float pr = sigmoidf_(row[5 + k]);
      if (p.probs) p.probs[gb * p.C + k] = pr;                             // model.py:168
      if (p.scores) p.scores[gb * p.C + k] = __fmul_rn(conf, pr);          // test_single_image.py:55
    }
  }
}

struct ReorgParams {
  const float* fm;
  int n, gh, gw, C, E;
  float ratio_h, ratio_w;
  float anchor_w[3], anchor_h[3];
  float* xy_offset;   // [gh,gw,1,2]
  float* boxes;       // [n,gh,gw,3,4] cx,cy,w,h
  float* conf_logits; // [n,gh,gw,3,1]
  float* prob_logits; // [n,gh,gw,3,C]
};

__global__ void __launch_bounds__(256) reorg_kernel(const ReorgParams p) {
  const long nb = (long)p.n * p.gh * p.gw * 3;
  const long total = nb * p.E;
  for (long e = blockIdx.x * (long)blockDim.x + threadIdx.x; e < total; e += (long)gridDim.x * blockDim.x) {
    const int j = (int)(e % p.E);
    const long gb = e / p.E;
    const int a = (int)(gb % 3);
    const long cell = (gb / 3) % ((long)p.gh * p.gw);
    const float v = p.fm[e];
    if (j >= 5) {
      if (p.prob_logits) p.prob_logits[gb * p.C + (j - 5)] = v;
    } else if (j == 4) {
      if (p.conf_logits) p.conf_logits[gb] = v;
    } else if (p.boxes) {
      const int axis = j & 1;
      const float off = axis == 0 ? (float)(cell % p.gw) : (float)(cell / p.gw);
      const float ratio = axis == 0 ? p.ratio_w : p.ratio_h;
      float o;
      if (j < 2) o = __fmul_rn(__fadd_rn(sigmoidf_(v), off), ratio);
      else o = __fmul_rn(__fmul_rn(expf(v), __fdiv_rn(axis == 0 ? p.anchor_w[a] : p.anchor_h[a], ratio)), ratio);
      p.boxes[gb * 4 + j] = o;
    }
    if (p.xy_offset && gb < (long)p.gh * p.gw * 3 && a == 0 && j < 2) {
      p.xy_offset[cell * 2 + j] = j == 0 ? (float)(cell % p.gw) : (float)(cell / p.gw);
    }
  }
}

static int grid_for(long total) {
  long g = (total + 255) / 256;
  const long cap = (long)num_sms() * 32;
  return (int)(g < cap ? g : cap);
}

}  // namespace yb

using namespace yb;

extern "C" int yb_predict(const float* fm1, const float* fm2, const float* fm3, int n, int img_h, int img_w,
                          int class_num, const float* anchors9x2, float* boxes, float* confs, float* probs,
                          float* scores, void* stream) {
  YB_REQUIRE(fm1 && fm2 && fm3 && anchors9x2 && boxes, "predict: null pointer");
  YB_REQUIRE(n > 0 && class_num > 0, "predict: bad n/class_num");
  YB_REQUIRE(img_h % 32 == 0 && img_w % 32 == 0 && img_h > 0 && img_w > 0,
             "predict: image size must be a multiple of 32 (got %dx%d)", img_h, img_w);
  DecodeParams p;
  p.fm[0] = fm1; p.fm[1] = fm2; p.fm[2] = fm3;
  const int div[3] = {32, 16, 8};
  p.box_off[0] = 0;
  for (int s = 0; s < 3; ++s) {
    p.gh[s] = img_h / div[s];
    p.gw[s] = img_w / div[s];
    p.box_off[s + 1] = p.box_off[s] + 3 * p.gh[s] * p.gw[s];
    p.ratio_h[s] = (float)((double)img_h / (double)p.gh[s]);   // model.py:91 (float64 divide, cast to f32)
    p.ratio_w[s] = (float)((double)img_w / (double)p.gw[s]);
  }
  for (int i = 0; i < 9; ++i) { p.anchor_w[i] = anchors9x2[2 * i]; p.anchor_h[i] = anchors9x2[2 * i + 1]; }
  p.n = n; p.C = class_num; p.E = 5 + class_num; p.B = p.box_off[3];
  p.boxes = boxes; p.confs = confs; p.probs = probs; p.scores = scores;
  long blocks = ((long)n * p.B + 7) / 8;
  const long capb = (long)num_sms() * 32;
  if (blocks > capb) blocks = capb;
  predict_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

extern "C" int yb_reorg_layer(const float* feature_map, int n, int gh, int gw, int img_h, int img_w, int class_num,
                              const float* anchors3x2, float* xy_offset, float* boxes, float* conf_logits,
                              float* prob_logits, void* stream) {
  YB_REQUIRE(feature_map && anchors3x2, "reorg_layer: null pointer");
  YB_REQUIRE(n > 0 && gh > 0 && gw > 0 && class_num > 0, "reorg_layer: bad shape");
  ReorgParams p;
  p.fm = feature_map; p.n = n; p.gh = gh; p.gw = gw; p.C = class_num; p.E = 5 + class_num;
  p.ratio_h = (float)((double)img_h / (double)gh);
  p.ratio_w = (float)((double)img_w / (double)gw);
  for (int i = 0; i < 3; ++i) { p.anchor_w[i] = anchors3x2[2 * i]; p.anchor_h[i] = anchors3x2[2 * i + 1]; }
  p.xy_offset = xy_offset; p.boxes = boxes; p.conf_logits = conf_logits; p.prob_logits = prob_logits;
  const long total = (long)n * gh * gw * 3 * p.E;
  reorg_kernel<<<grid_for(total), 256, 0, static_cast<cudaStream_t>(stream)>>>(p);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}
