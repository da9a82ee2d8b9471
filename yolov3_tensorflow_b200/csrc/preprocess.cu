// The two host-side steps either side of the hot path, moved onto the device:
//   * process_box (utils/data_utils.py:51-115): ground-truth box lists -> the three y_true tensors consumed by
//     loss_layer.  The reference builds 3.66 MB per 416x416 image in numpy and ships it host -> device; here only the
//     box lists (<= 50 x 24 B per image) cross PCIe and the tensors are produced at HBM speed.
//   * letterbox_resize + BGR->RGB + /255 (utils/data_aug.py:274-293, test_single_image.py:39-46): uint8 BGR image ->
//     float32 RGB network input, nearest-neighbour (the reference's interp=0) with the 128-grey border.
// Both are bit-exact restatements: float32 operations in the reference's order (__f*_rn: no FMA contraction), the
// resize index in double like OpenCV's resizeNN.
#include "common.cuh"

namespace yb {

// y_true[..., :] = 0 and y_true[..., -1] = 1 (utils/data_utils.py:72-79) for the three scales in one launch:
// float4 stores over each tensor's multiple-of-4 prefix, the <= 3 trailing floats by one thread.
__global__ void __launch_bounds__(256) ytrue_fill_kernel(float* __restrict__ y1, long n1, float* __restrict__ y2, long n2,
                                                          float* __restrict__ y3, long n3, int E1 /* 6 + C */) {
  const long q1 = n1 / 4, q2 = n2 / 4, q3 = n3 / 4;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < q1 + q2 + q3; i += (long)gridDim.x * blockDim.x) {
    long e;
    float* dst;
    if (i < q1) { e = i * 4; dst = y1 + e; }
    else if (i < q1 + q2) { e = (i - q1) * 4; dst = y2 + e; }
    else { e = (i - q1 - q2) * 4; dst = y3 + e; }
    const int m = (int)(e % E1);                 // position of the first of the 4 elements inside its box record
    float4 v;                                    // the record's last element (mix-up weight) is the only 1
    v.x = ((m + 0) % E1 == E1 - 1) ? 1.f : 0.f;
    v.y = ((m + 1) % E1 == E1 - 1) ? 1.f : 0.f;
    v.z = ((m + 2) % E1 == E1 - 1) ? 1.f : 0.f;
    v.w = ((m + 3) % E1 == E1 - 1) ? 1.f : 0.f;
    *reinterpret_cast<float4*>(dst) = v;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (long e = q1 * 4; e < n1; ++e) y1[e] = (e % E1 == E1 - 1) ? 1.f : 0.f;
    for (long e = q2 * 4; e < n2; ++e) y2[e] = (e % E1 == E1 - 1) ? 1.f : 0.f;
    for (long e = q3 * 4; e < n3; ++e) y3[e] = (e % E1 == E1 - 1) ? 1.f : 0.f;
  }
}

struct PBoxParams {
  const float* boxes;      // [n, vmax, 5] x_min, y_min, x_max, y_max, mixup weight
  const int* labels;       // [n, vmax]
  const int* counts;       // [n] valid boxes per image (<= vmax)
  int n, vmax, C;
  int gw[3], gh[3];        // grid sizes of y_true_13 / _26 / _52 (named after the 416 case)
  float* y[3];
  float aw[9], ah[9];      // anchors, reference order (small -> large)
};

static constexpr int PBOX_MAX = 256;   // boxes per image handled by one CTA

// One CTA per image.  Phase 1: every box picks its best anchor (IoU of the centred sizes, first maximum wins like
// np.argmax) and its grid cell.  Phase 2: the reference writes the boxes in list order, so for a (scale, cell, anchor)
// slot hit by several boxes the LAST one's coordinates / mix weight survive while every box's class bit stays set.
__global__ void __launch_bounds__(PBOX_MAX) process_box_kernel(const PBoxParams p) {
  __shared__ int s_slot[PBOX_MAX];     // linear slot id (scale, y, x, k) or -1
  const int img = blockIdx.x;
  const int v = min(p.counts[img], p.vmax);
  const int i = threadIdx.x;
  float cx = 0.f, cy = 0.f, sw = 0.f, sh = 0.f, mixw = 0.f;
  int g = 0, k = 0, gx = 0, gy = 0, cls = 0, slot = -1;
  if (i < v) {
    const float* b = p.boxes + ((long)img * p.vmax + i) * 5;
    const float x0 = b[0], y0 = b[1], x1 = b[2], y1 = b[3];
    mixw = b[4];
    cx = __fdiv_rn(__fadd_rn(x0, x1), 2.f);          // (boxes[:, 0:2] + boxes[:, 2:4]) / 2
    cy = __fdiv_rn(__fadd_rn(y0, y1), 2.f);
    sw = __fsub_rn(x1, x0);                          // boxes[:, 2:4] - boxes[:, 0:2]
    sh = __fsub_rn(y1, y0);
    const float hw = __fdiv_rn(sw, 2.f), hh = __fdiv_rn(sh, 2.f);
    float best = 0.f;
    int bi = 0;
#pragma unroll
    for (int a = 0; a < 9; ++a) {
      const float aw2 = __fdiv_rn(p.aw[a], 2.f), ah2 = __fdiv_rn(p.ah[a], 2.f);
      const float w = __fsub_rn(fminf(hw, aw2), fmaxf(-hw, -aw2));    // maxs - mins
      const float h = __fsub_rn(fminf(hh, ah2), fmaxf(-hh, -ah2));
      const float inter = __fmul_rn(w, h);
      const float den = __fadd_rn(__fsub_rn(__fadd_rn(__fmul_rn(sw, sh), __fmul_rn(p.aw[a], p.ah[a])), inter), 1e-10f);
      const float iou = __fdiv_rn(inter, den);
      if (a == 0 || iou > best) { best = iou; bi = a; }               // np.argmax: first maximum
    }
    g = 2 - bi / 3;                                  // 0,1,2 -> y_true_52 ; 6,7,8 -> y_true_13
    k = bi % 3;
    const float ratio = g == 0 ? 32.f : (g == 1 ? 16.f : 8.f);
    gx = (int)floorf(__fdiv_rn(cx, ratio));
    gy = (int)floorf(__fdiv_rn(cy, ratio));
    cls = p.labels[(long)img * p.vmax + i];
    if (gx >= 0 && gx < p.gw[g] && gy >= 0 && gy < p.gh[g] && cls >= 0 && cls < p.C)
      slot = ((g * 4096 + gy) * 4096 + gx) * 3 + k;  // (the reference raises IndexError for a box outside the image)
  }
  s_slot[i] = slot;
  __syncthreads();
  if (slot < 0) return;
  bool last = true;
  for (int j = i + 1; j < v; ++j)
    if (s_slot[j] == slot) { last = false; break; }
  const int E1 = 6 + p.C;
  float* rec = p.y[g] + ((((long)img * p.gh[g] + gy) * p.gw[g] + gx) * 3 + k) * E1;
  rec[5 + cls] = 1.f;                                // every box leaves its class bit
  if (last) {
    rec[0] = cx; rec[1] = cy; rec[2] = sw; rec[3] = sh;
    rec[4] = 1.f;
    rec[E1 - 1] = mixw;
  }
}

// letterbox_resize(img, new_w, new_h, interp=0) -> cvtColor(BGR2RGB) -> float32 / 255
__global__ void __launch_bounds__(256)
letterbox_kernel(const uint8_t* __restrict__ src, int sh, int sw, long src_pitch, int rh, int rw, int dh, int dw, int nh,
                 int nw, double ify, double ifx, float* __restrict__ dst) {
  const long total = (long)nh * nw;
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < total; i += (long)gridDim.x * blockDim.x) {
    const int y = (int)(i / nw), x = (int)(i - (long)y * nw);
    float r = 128.f, g = 128.f, b = 128.f;          // np.full(..., 128, np.uint8)
    const int ry = y - dh, rx = x - dw;
    if (ry >= 0 && ry < rh && rx >= 0 && rx < rw) {
      // OpenCV resizeNN: src index = min(floor(dst index * (1 / (dst size / src size))), src size - 1), in double
      const int sy = min((int)floor(ry * ify), sh - 1);
      const int sx = min((int)floor(rx * ifx), sw - 1);
      const uint8_t* px = src + (long)sy * src_pitch + (long)sx * 3;
      b = (float)px[0]; g = (float)px[1]; r = (float)px[2];
    }
    float* o = dst + i * 3;
    o[0] = __fdiv_rn(r, 255.f); o[1] = __fdiv_rn(g, 255.f); o[2] = __fdiv_rn(b, 255.f);
  }
}

}  // namespace yb

using namespace yb;

extern "C" int yb_process_box(const float* boxes, const int32_t* labels, const int32_t* counts, int n, int vmax,
                              int img_w, int img_h, int class_num, const float* anchors9x2, float* y_true_1,
                              float* y_true_2, float* y_true_3, void* stream) {
  YB_REQUIRE(boxes && labels && counts && anchors9x2 && y_true_1 && y_true_2 && y_true_3, "process_box: null pointer");
  YB_REQUIRE(n > 0 && vmax > 0 && vmax <= PBOX_MAX, "process_box: vmax must be in 1..%d (got %d)", PBOX_MAX, vmax);
  YB_REQUIRE(class_num > 0 && img_w > 0 && img_h > 0 && img_w % 32 == 0 && img_h % 32 == 0 && img_w / 8 < 4096 && img_h / 8 < 4096,
             "process_box: image size must be a multiple of 32 (got %dx%d)", img_w, img_h);
  YB_REQUIRE(((uintptr_t)y_true_1 & 15) == 0 && ((uintptr_t)y_true_2 & 15) == 0 && ((uintptr_t)y_true_3 & 15) == 0,
             "process_box: y_true tensors must be 16-byte aligned");
  PBoxParams p;
  p.boxes = boxes; p.labels = labels; p.counts = counts; p.n = n; p.vmax = vmax; p.C = class_num;
  const int div[3] = {32, 16, 8};
  long cnt[3];
  float* ys[3] = {y_true_1, y_true_2, y_true_3};
  for (int s = 0; s < 3; ++s) {
    p.gw[s] = img_w / div[s]; p.gh[s] = img_h / div[s]; p.y[s] = ys[s];
    cnt[s] = (long)n * p.gh[s] * p.gw[s] * 3 * (6 + class_num);
  }
  for (int a = 0; a < 9; ++a) { p.aw[a] = anchors9x2[2 * a]; p.ah[a] = anchors9x2[2 * a + 1]; }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long total4 = (cnt[0] + cnt[1] + cnt[2]) / 4;
  long blocks = (total4 + 255) / 256;
  const long cap = (long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  ytrue_fill_kernel<<<(int)blocks, 256, 0, st>>>(y_true_1, cnt[0], y_true_2, cnt[1], y_true_3, cnt[2], 6 + class_num);
  YB_CUDA(cudaGetLastError());
  process_box_kernel<<<n, PBOX_MAX, 0, st>>>(p);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

extern "C" int yb_letterbox_params(int src_h, int src_w, int new_h, int new_w, double* resize_ratio, int* resize_h,
                                   int* resize_w, int* dh, int* dw) {
  YB_REQUIRE(src_h > 0 && src_w > 0 && new_h > 0 && new_w > 0 && resize_ratio && resize_h && resize_w && dh && dw,
             "letterbox_params: bad argument");
  const double a = (double)new_w / (double)src_w, b = (double)new_h / (double)src_h;    // utils/data_aug.py:280
  const double ratio = a < b ? a : b;
  *resize_ratio = ratio;
  *resize_w = (int)(ratio * src_w);                    // int(): truncation towards zero
  *resize_h = (int)(ratio * src_h);
  *dw = (int)((new_w - *resize_w) / 2.0);
  *dh = (int)((new_h - *resize_h) / 2.0);
  return YB_OK;
}

extern "C" int yb_letterbox_normalize(const uint8_t* bgr, int src_h, int src_w, long src_pitch_bytes, int new_h,
                                      int new_w, float* out_rgb, void* stream) {
  YB_REQUIRE(bgr && out_rgb && src_pitch_bytes >= 3L * src_w, "letterbox: bad argument");
  double ratio;
  int rh, rw, dh, dw;
  int rc = yb_letterbox_params(src_h, src_w, new_h, new_w, &ratio, &rh, &rw, &dh, &dw);
  if (rc) return rc;
  YB_REQUIRE(rh > 0 && rw > 0, "letterbox: image too small for the target size");
  // OpenCV: inv_scale = dsize / ssize; ifx = 1. / inv_scale  (modules/imgproc/src/resize.cpp, resizeNN)
  const double ifx = 1.0 / ((double)rw / (double)src_w), ify = 1.0 / ((double)rh / (double)src_h);
  const long total = (long)new_h * new_w;
  long blocks = (total + 255) / 256;
  const long cap = (long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  letterbox_kernel<<<(int)blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(bgr, src_h, src_w, src_pitch_bytes, rh, rw, dh,
                                                                             dw, new_h, new_w, ify, ifx, out_rgb);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}
