// Network plan: Darknet-53 + 3-scale YOLOv3 head (model.py:30-80, utils/layer_utils.py:24-79)
// for a fixed (batch, H, W, dtype): layer schedule, activation/parameter arena layout,
// TMA tensor maps.  concat (model.py:62,72) and NN-upsample (utils/layer_utils.py:82-87)
// never run as ops: producers store straight into channel slices of the concat buffers.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"
#include "conv.cuh"

#include "net.cuh"

namespace yb {
void train_layout(yb_net* net);
int train_bind(yb_net* net, cudaStream_t st);
int train_refresh_dgrad_weights(yb_net* net, int layer, void* stream);
int nms_candidate_buffers(void* workspace, size_t workspace_bytes, int n_images, int num_boxes, int num_classes,
                          int max_boxes, int** cand_count, float** cand_score, int** cand_idx);
int nms_select_gather(const float* boxes, int n_images, int num_boxes, int num_classes, int max_boxes, float iou_thresh,
                      void* workspace, size_t workspace_bytes, float* out_boxes, float* out_scores, int32_t* out_labels,
                      int32_t* out_indices, int32_t* out_counts, cudaStream_t st);

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

struct Builder {
  yb_net* net;
  int body_k = 0, head_k = 0;

  int new_buf(int h, int w, int ld, int fp32) {
    Buf b; b.h = h; b.w = w; b.ld = ld; b.fp32 = fp32;
    net->bufs.push_back(b);
    return (int)net->bufs.size() - 1;
  }
  // conv2d of utils/layer_utils.py:9-22.  dst: optional pre-allocated destination view.
  Ten conv(const Ten& x, int cout, int k, int s, bool head, bool bn = true, const Ten* shortcut = nullptr,
           const Ten* dst = nullptr, bool upsample = false) {
    Layer L;
    memset(&L.info, 0, sizeof(L.info));
    L.info.index = (int)net->layers.size();
    L.info.cin = x.c; L.info.cout = cout; L.info.ksize = k; L.info.stride = s; L.info.has_bn = bn ? 1 : 0;
    L.info.in_h = x.h; L.info.in_w = x.w; L.info.out_h = x.h / s; L.info.out_w = x.w / s;
    L.info.is_head = head ? 1 : 0;
    L.info.scope_index = head ? head_k++ : body_k++;
    L.in = x;
    L.upsample = upsample;
    L.info.upsample2x = upsample ? 1 : 0;
    L.out_fp32 = !bn;
    L.cout_pad = yb_conv_cout_pad(cout);
    if (shortcut) L.res = *shortcut; else L.res.buf = -2;
    Ten y;
    if (dst) {
      y = *dst;
    } else {
      y.buf = new_buf(x.h / s, x.w / s, cout, !bn);
      y.off = 0;
    }
    y.c = cout;
    y.h = (x.h / s) * (upsample ? 2 : 1);
    y.w = (x.w / s) * (upsample ? 2 : 1);
    L.out = y;
    net->layers.push_back(L);
    return y;
  }
  Ten res_block(const Ten& x, int f, const Ten* dst = nullptr) {   // utils/layer_utils.py:25-32
    Ten a = conv(x, f, 1, 1, false);
    return conv(a, 2 * f, 3, 1, false, true, &x, dst);
  }
  void yolo_block(const Ten& x, int f, Ten* route, Ten* net_out) {  // utils/layer_utils.py:71-79
    Ten t = conv(x, f, 1, 1, true);
    t = conv(t, 2 * f, 3, 1, true);
    t = conv(t, f, 1, 1, true);
    t = conv(t, 2 * f, 3, 1, true);
    t = conv(t, f, 1, 1, true);
    *route = t;
    *net_out = conv(t, 2 * f, 3, 1, true);
  }

  void build() {
    const int H = net->h, W = net->w, D = 3 * (5 + net->class_num);
    // concat buffers (model.py:62,72): [upsampled | route]
    const int cat1 = new_buf(H / 16, W / 16, 256 + 512, 0);
    const int cat2 = new_buf(H / 8, W / 8, 128 + 256, 0);
    Ten x; x.buf = -1; x.c = 3; x.h = H; x.w = W;
    // ---- darknet53_body (utils/layer_utils.py:35-66) ----
    Ten t = conv(x, 32, 3, 1, false);
    t = conv(t, 64, 3, 2, false);
    t = res_block(t, 32);
    t = conv(t, 128, 3, 2, false);
    for (int i = 0; i < 2; ++i) t = res_block(t, 64);
    t = conv(t, 256, 3, 2, false);
    Ten r1dst; r1dst.buf = cat2; r1dst.off = 128;
    for (int i = 0; i < 8; ++i) t = res_block(t, 128, i == 7 ? &r1dst : nullptr);
    const Ten route1 = t;
    t = conv(t, 512, 3, 2, false);
    Ten r2dst; r2dst.buf = cat1; r2dst.off = 256;
    for (int i = 0; i < 8; ++i) t = res_block(t, 256, i == 7 ? &r2dst : nullptr);
    const Ten route2 = t;
    t = conv(t, 1024, 3, 2, false);
    for (int i = 0; i < 4; ++i) t = res_block(t, 512);
    const Ten route3 = t;
    (void)route1; (void)route2;
    // ---- yolov3_head (model.py:53-78) ----
    Ten inter, nt;
    yolo_block(route3, 512, &inter, &nt);
    Ten fm1 = conv(nt, D, 1, 1, true, false);
    Ten up1; up1.buf = cat1; up1.off = 0;
    conv(inter, 256, 1, 1, true, true, nullptr, &up1, true);
    Ten c1; c1.buf = cat1; c1.off = 0; c1.c = 768; c1.h = H / 16; c1.w = W / 16;
    yolo_block(c1, 256, &inter, &nt);
    Ten fm2 = conv(nt, D, 1, 1, true, false);
    Ten up2; up2.buf = cat2; up2.off = 0;
    conv(inter, 128, 1, 1, true, true, nullptr, &up2, true);
    Ten c2; c2.buf = cat2; c2.off = 0; c2.c = 384; c2.h = H / 8; c2.w = W / 8;
    yolo_block(c2, 128, &inter, &nt);
    Ten fm3 = conv(nt, D, 1, 1, true, false);
    net->fm_buf[0] = fm1.buf; net->fm_buf[1] = fm2.buf; net->fm_buf[2] = fm3.buf;

    // ---- arenas ----
    const size_t esz = 2;
    size_t o = 0;
    for (auto& b : net->bufs) {
      b.bytes = (size_t)net->n * b.h * b.w * b.ld * (b.fp32 ? 4 : esz);
      b.offset = o;
      o = align256(o + b.bytes);
    }
    net->act_bytes = o;
    o = 0;
    for (auto& L : net->layers) {
      const size_t kk = (size_t)L.info.ksize * L.info.ksize * L.info.cin;
      L.w_master = o; o = align256(o + (size_t)L.info.cout * kk * 4);
      L.w_packed = o; o = align256(o + (size_t)L.cout_pad * kk * esz);
      const size_t cb = (size_t)L.cout_pad * 4;
      if (L.info.has_bn) {
        L.gamma = o; o = align256(o + cb);
        L.beta = o;  o = align256(o + cb);
        L.mean = o;  o = align256(o + cb);
        L.var = o;   o = align256(o + cb);
      } else {
        L.bias = o;  o = align256(o + cb);
      }
      L.scale = o; o = align256(o + cb);
      if (L.info.has_bn) { L.shift = o; o = align256(o + cb); }
      else L.shift = L.bias;   // detection convs: the epilogue's shift IS the (trainable) bias
    }
    net->param_bytes = o;
  }
};

static void* ten_ptr(const yb_net* net, const Ten& t) {
  const Buf& b = net->bufs[t.buf];
  return net->act + b.offset + (size_t)t.off * (b.fp32 ? 4 : 2);
}

__global__ void fill_kernel(float* p, int n, float v) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

}  // namespace yb

using namespace yb;

extern "C" int yb_net_create(yb_net** out, int class_num, int n, int h, int w, int dtype, int training) {
  YB_REQUIRE(out, "net_create: null out pointer");
  YB_REQUIRE(class_num > 0 && n > 0, "net_create: bad class_num/batch");
  YB_REQUIRE(h > 0 && w > 0 && h % 32 == 0 && w % 32 == 0, "net_create: H,W must be multiples of 32 (got %dx%d)", h, w);
  YB_REQUIRE(dtype == YB_F16 || dtype == YB_BF16, "net_create: dtype must be f16 or bf16");
  yb_net* net = new yb_net();
  net->class_num = class_num; net->n = n; net->h = h; net->w = w; net->dtype = dtype; net->training = training;
  Builder b{net};
  b.build();
  if (training) train_layout(net);
  *out = net;
  return YB_OK;
}

extern "C" int yb_net_destroy(yb_net* net) { delete net; return YB_OK; }

extern "C" int yb_net_num_layers(const yb_net* net) { return net ? (int)net->layers.size() : YB_ERR_INVALID_ARGUMENT; }

extern "C" int yb_net_layer_info(const yb_net* net, int layer, yb_layer_info* info) {
  YB_REQUIRE(net && info && layer >= 0 && layer < (int)net->layers.size(), "layer_info: bad argument");
  *info = net->layers[layer].info;
  return YB_OK;
}

extern "C" int yb_net_arena_bytes(const yb_net* net, size_t* activation_bytes, size_t* param_bytes) {
  YB_REQUIRE(net && activation_bytes && param_bytes, "arena_bytes: bad argument");
  *activation_bytes = net->act_bytes;
  *param_bytes = net->param_bytes;
  return YB_OK;
}

extern "C" int yb_net_bind(yb_net* net, void* activation_arena, size_t activation_bytes, void* param_arena,
                           size_t param_bytes, void* stream) {
  YB_REQUIRE(net && activation_arena && param_arena, "bind: null pointer");
  YB_REQUIRE(activation_bytes >= net->act_bytes && param_bytes >= net->param_bytes, "bind: arena too small");
  YB_REQUIRE(((uintptr_t)activation_arena & 255) == 0 && ((uintptr_t)param_arena & 255) == 0,
             "bind: arenas must be 256-byte aligned");
  net->act = static_cast<uint8_t*>(activation_arena);
  net->par = static_cast<uint8_t*>(param_arena);
  // prepare every tensor-core conv (layer 0 is the CUDA-core stem)
  for (size_t i = 1; i < net->layers.size(); ++i) {
    Layer& L = net->layers[i];
    yb_conv_desc d;
    memset(&d, 0, sizeof(d));
    d.n = net->n; d.h = L.info.in_h; d.w = L.info.in_w; d.cin = L.info.cin; d.cout = L.info.cout;
    d.ksize = L.info.ksize; d.stride = L.info.stride;
    d.in_ld = net->bufs[L.in.buf].ld; d.out_ld = net->bufs[L.out.buf].ld;
    d.res_ld = L.res.buf >= 0 ? net->bufs[L.res.buf].ld : 0;
    d.dtype = net->dtype; d.out_fp32 = L.out_fp32; d.leaky = L.info.has_bn; d.upsample2x = L.upsample;
    int cp = 0;
    int rc = conv_prepare(&d, ten_ptr(net, L.in), net->par + L.w_packed,
                          reinterpret_cast<const float*>(net->par + L.scale),
                          reinterpret_cast<const float*>(net->par + L.shift),
                          L.res.buf >= 0 ? ten_ptr(net, L.res) : nullptr, ten_ptr(net, L.out), nullptr, nullptr, &L.tmA,
                          &L.tmB, &L.params, &cp);
    if (rc) return rc;
    L.prepared = true;
    L.halo_ok = false;
    if (L.info.has_bn && conv_halo_supported(&d)) {
      L.halo_desc = d;
      rc = conv_halo_prepare(&d, ten_ptr(net, L.in), net->par + L.w_packed, reinterpret_cast<const float*>(net->par + L.scale),
                             reinterpret_cast<const float*>(net->par + L.shift),
                             L.res.buf >= 0 ? ten_ptr(net, L.res) : nullptr, ten_ptr(net, L.out), &L.halo_maps, &L.halo_params);
      if (rc) return rc;
      L.halo_ok = true;
    }
    if (!L.info.has_bn) {
      // fused-decode variant of the head (yb_net_detect); class counts without a kernel keep the unfused pipeline
      L.det_ok = conv_prepare_det(&d, net->class_num, ten_ptr(net, L.in), net->par + L.w_packed,
                                  reinterpret_cast<const float*>(net->par + L.scale),
                                  reinterpret_cast<const float*>(net->par + L.shift), &L.det_tmA, &L.det_tmB, &L.det_params,
                                  &L.det_cout_pad) == YB_OK;
    }
  }
  if (net->training) return train_bind(net, static_cast<cudaStream_t>(stream));
  return YB_OK;
}

extern "C" int yb_net_set_conv_params(yb_net* net, int layer, const float* w, int layout, const float* gamma,
                                      const float* beta, const float* mean, const float* var, const float* bias,
                                      void* stream) {
  YB_REQUIRE(net && net->par, "set_conv_params: net not bound");
  YB_REQUIRE(layer >= 0 && layer < (int)net->layers.size() && w, "set_conv_params: bad argument");
  Layer& L = net->layers[layer];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int c = L.info.cout;
  int rc = yb_pack_conv_weights(w, layout, c, L.info.cin, L.info.ksize, c, YB_F32, net->par + L.w_master, stream);
  if (rc) return rc;
  rc = yb_pack_conv_weights(w, layout, c, L.info.cin, L.info.ksize, L.cout_pad, net->dtype, net->par + L.w_packed, stream);
  if (rc) return rc;
  float* scale = reinterpret_cast<float*>(net->par + L.scale);
  float* shift = reinterpret_cast<float*>(net->par + L.shift);
  if (L.info.has_bn) {
    YB_REQUIRE(gamma && beta && mean && var, "set_conv_params: layer %d needs gamma/beta/mean/var", layer);
    YB_CUDA(cudaMemcpyAsync(net->par + L.gamma, gamma, c * 4, cudaMemcpyDeviceToDevice, st));
    YB_CUDA(cudaMemcpyAsync(net->par + L.beta, beta, c * 4, cudaMemcpyDeviceToDevice, st));
    YB_CUDA(cudaMemcpyAsync(net->par + L.mean, mean, c * 4, cudaMemcpyDeviceToDevice, st));
    YB_CUDA(cudaMemcpyAsync(net->par + L.var, var, c * 4, cudaMemcpyDeviceToDevice, st));
    rc = yb_bn_fold(gamma, beta, mean, var, c, 1e-5f, scale, shift, stream);   // model.py:37 epsilon
    if (rc) return rc;
  } else {
    YB_REQUIRE(bias, "set_conv_params: layer %d needs a bias", layer);
    YB_CUDA(cudaMemsetAsync(shift, 0, L.cout_pad * 4, st));          // shift aliases the bias buffer
    YB_CUDA(cudaMemcpyAsync(net->par + L.bias, bias, c * 4, cudaMemcpyDeviceToDevice, st));
    fill_kernel<<<ceil_div(L.cout_pad, 128), 128, 0, st>>>(scale, L.cout_pad, 1.0f);
    YB_CUDA(cudaGetLastError());
  }
  if (net->training) return train_refresh_dgrad_weights(net, layer, stream);
  return YB_OK;
}

extern "C" int yb_net_refold_bn(yb_net* net, void* stream) {
  YB_REQUIRE(net && net->par, "refold_bn: net not bound");
  for (auto& L : net->layers) {
    if (!L.info.has_bn) continue;
    int rc = yb_bn_fold(reinterpret_cast<const float*>(net->par + L.gamma), reinterpret_cast<const float*>(net->par + L.beta),
                        reinterpret_cast<const float*>(net->par + L.mean), reinterpret_cast<const float*>(net->par + L.var),
                        L.info.cout, net->bn_eps, reinterpret_cast<float*>(net->par + L.scale),
                        reinterpret_cast<float*>(net->par + L.shift), stream);
    if (rc) return rc;
  }
  net->fold_dirty = false;
  return YB_OK;
}

extern "C" int yb_net_forward(yb_net* net, const float* images, float* fm1, float* fm2, float* fm3, void* stream) {
  return yb_net_forward_layers(net, images, fm1, fm2, fm3, 0, 1 << 30, stream);
}

static int forward_layers_impl(yb_net* net, const float* images, float* fm1, float* fm2, float* fm3, int first, int last,
                               const DetParams* det /* [3] or NULL */, void* stream);

extern "C" int yb_net_forward_layers(yb_net* net, const float* images, float* fm1, float* fm2, float* fm3, int first,
                                     int last, void* stream) {
  return forward_layers_impl(net, images, fm1, fm2, fm3, first, last, nullptr, stream);
}

static int forward_layers_impl(yb_net* net, const float* images, float* fm1, float* fm2, float* fm3, int first, int last,
                               const DetParams* det, void* stream) {
  YB_REQUIRE(net && net->act && net->par, "forward: net not bound");
  YB_REQUIRE(images, "forward: null images");
  YB_REQUIRE(first >= 0 && first <= last, "forward: bad layer range");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  float* user_fm[3] = {fm1, fm2, fm3};
  if (net->fold_dirty) {   // BN parameters / moving statistics changed by a training step: refold for inference
    int rc = yb_net_refold_bn(net, stream);
    if (rc) return rc;
  }
  const bool thin = opt("YB_THIN")[0] != '0';   // A/B switch for the thin-layer kernels
  // detect path: heads 1 and 2 on a side stream (YB_HEAD_STREAM=0: everything on the caller's stream)
  const bool head_overlap = det != nullptr && opt("YB_HEAD_STREAM")[0] != '0';
  bool forked = false;
  if (head_overlap && net->side_stream == nullptr) {
    YB_CUDA(cudaStreamCreateWithFlags(&net->side_stream, cudaStreamNonBlocking));
    YB_CUDA(cudaEventCreateWithFlags(&net->side_fork, cudaEventDisableTiming));
    YB_CUDA(cudaEventCreateWithFlags(&net->side_join, cudaEventDisableTiming));
  }
  // stem fused into Conv_1 (csrc/conv_halo.cu): layer 0's output is never written.  YB_STEM_FUSE=0: two launches.
  // (a layer range that stops at layer 0 then launches nothing: the stem no longer exists as a launch of its own)
  const bool fuse_stem = opt("YB_STEM_FUSE")[0] != '0' && net->layers.size() > 1 && net->layers[1].halo_ok &&
                         net->layers[1].info.cin == 32 && net->layers[1].info.stride == 2 && opt("YB_HALO")[0] != '0';
  if (first == 0 && !fuse_stem) {
    Layer& L = net->layers[0];
    int rc;
    if (thin)
      rc = yb_stem_conv_fwd_tc(images, reinterpret_cast<const float*>(net->par + L.w_master),
                               reinterpret_cast<const float*>(net->par + L.scale),
                               reinterpret_cast<const float*>(net->par + L.shift), net->n, net->h, net->w, net->dtype, 1,
                               ten_ptr(net, L.out), stream);
    else
      rc = yb_stem_conv_fwd(images, reinterpret_cast<const float*>(net->par + L.w_master),
                            reinterpret_cast<const float*>(net->par + L.scale),
                            reinterpret_cast<const float*>(net->par + L.shift), net->n, net->h, net->w, L.info.cout,
                            net->dtype, 1, ten_ptr(net, L.out), stream);
    if (rc) return rc;
  }
  for (size_t i = first > 1 ? first : 1; i < net->layers.size() && (int)i <= last; ++i) {
    Layer& L = net->layers[i];
    // (after the r01_j pipeline-loop fixes the tcgen05 kernel runs these two layers in ~455 us against ~500-550 us for
    //  the mma.sync halo kernel, so the latter is now opt-in: YB_THIN=2)
    const bool thin_cin32 = opt("YB_THIN")[0] == '2';
    if (thin_cin32 && L.info.ksize == 3 && L.info.cin == 32 && L.info.has_bn && !L.upsample) {
      // Cin = 32: 64-byte im2col rows halve the TMA line rate -> direct halo-tile kernel (csrc/conv_thin.cu)
      yb_conv_desc d;
      memset(&d, 0, sizeof(d));
      d.n = net->n; d.h = L.info.in_h; d.w = L.info.in_w; d.cin = 32; d.cout = L.info.cout; d.ksize = 3;
      d.stride = L.info.stride; d.in_ld = net->bufs[L.in.buf].ld; d.out_ld = net->bufs[L.out.buf].ld;
      d.res_ld = L.res.buf >= 0 ? net->bufs[L.res.buf].ld : 0; d.dtype = net->dtype; d.leaky = 1;
      int rc = yb_conv3x3_thin_fwd(&d, ten_ptr(net, L.in), net->par + L.w_packed,
                                   reinterpret_cast<const float*>(net->par + L.scale),
                                   reinterpret_cast<const float*>(net->par + L.shift),
                                   L.res.buf >= 0 ? ten_ptr(net, L.res) : nullptr, ten_ptr(net, L.out), stream);
      if (rc) return rc;
      continue;
    }
    if (i == 1 && fuse_stem && first <= 1) {
      Layer& L0 = net->layers[0];
      HaloMaps hm;
      HaloParams hp;
      int rc = conv_stem_halo_prepare(&L.halo_desc, images, reinterpret_cast<const float*>(net->par + L0.w_master),
                                      reinterpret_cast<const float*>(net->par + L0.scale),
                                      reinterpret_cast<const float*>(net->par + L0.shift), net->par + L.w_packed,
                                      reinterpret_cast<const float*>(net->par + L.scale),
                                      reinterpret_cast<const float*>(net->par + L.shift), ten_ptr(net, L.out), &hm, &hp);
      if (rc) return rc;
      rc = conv_stem_halo_launch(&L.halo_desc, hm, hp, st);
      if (rc) return rc;
      continue;
    }
    // halo-tile kernel: measured (profiles/r02_c, batch 64 @416) 302 vs 381 us on Conv_3 (32->64 @208^2) and 244 vs 382 us
    // on Conv_1 (32->64 /2), but 227 vs 188 us on the 64->128 layers (their weights leave room for two halo stages only):
    // default on for Cin = 32.  YB_HALO=0: never, YB_HALO=1: wherever supported.
    const char* hopt = opt("YB_HALO");
    if (L.halo_ok && hopt[0] != '0' && (hopt[0] == '1' || L.info.cin == 32)) {
      int rc = conv_halo_launch(&L.halo_desc, L.halo_maps, L.halo_params, st);
      if (rc) return rc;
      continue;
    }
    ConvParams* p = &L.params;
    if (!L.info.has_bn) {
      int which = L.out.buf == net->fm_buf[0] ? 0 : (L.out.buf == net->fm_buf[1] ? 1 : 2);
      if (det) {                                      // decode + candidate filter in the epilogue, no feature map
        ConvParams dp = L.det_params;
        dp.det = det[which];
        // The first two heads are leaves of the graph (nothing but the NMS reads what they produce) and small (43 / 170
        // tiles): they run on the plan's side stream beside the upsampling branch that continues on the caller's stream;
        // the last head is on the critical path.  Every layer output has its own buffer, so the head's input stays intact.
        cudaStream_t hs = st;
        if (head_overlap && which < 2) {
          YB_CUDA(cudaEventRecord(net->side_fork, st));
          YB_CUDA(cudaStreamWaitEvent(net->side_stream, net->side_fork, 0));
          hs = net->side_stream;
          forked = true;
        }
        int rc = conv_launch(net->dtype, L.det_cout_pad, L.det_tmA, L.det_tmB, dp, hs);
        if (rc) return rc;
        continue;
      }
      p->out = user_fm[which] ? (void*)user_fm[which] : ten_ptr(net, L.out);
    }
    int rc = conv_launch(net->dtype, L.cout_pad, L.tmA, L.tmB, *p, st);
    if (rc) return rc;
  }
  if (forked) {
    YB_CUDA(cudaEventRecord(net->side_join, net->side_stream));
    YB_CUDA(cudaStreamWaitEvent(st, net->side_join, 0));
  }
  return YB_OK;
}

// ---------------------------------------------------------------------------------------------------------------
// forward -> predict -> score = conf * prob -> per-image gpu_nms in one call (test_single_image.py:50-57): the decode and
// the score filter run inside the three detection-head epilogues, then the greedy selection reads the candidate lists.
// ---------------------------------------------------------------------------------------------------------------
static int det_num_boxes(const yb_net* net) {
  int b = 0;
  for (int s = 0; s < 3; ++s) b += 3 * (net->h / (32 >> s)) * (net->w / (32 >> s));
  return b;
}

extern "C" int yb_net_detect_supported(const yb_net* net) {
  if (!net || !net->act) return 0;
  for (auto& L : net->layers) if (!L.info.has_bn && !L.det_ok) return 0;
  return 1;
}

extern "C" int yb_net_detect_workspace_bytes(const yb_net* net, int max_boxes, size_t* bytes) {
  YB_REQUIRE(net && bytes && max_boxes >= 0, "detect_workspace_bytes: bad argument");
  return yb_nms_workspace_bytes(net->n, det_num_boxes(net), net->class_num, max_boxes, bytes);
}

extern "C" int yb_net_detect(yb_net* net, const float* images, const float* anchors9x2, int max_boxes, float score_thresh,
                             float iou_thresh, void* workspace, size_t workspace_bytes, float* boxes, float* out_boxes,
                             float* out_scores, int32_t* out_labels, int32_t* out_indices, int32_t* out_counts,
                             void* stream) {
  return yb_net_detect_phases(net, images, anchors9x2, max_boxes, score_thresh, iou_thresh, workspace, workspace_bytes,
                              boxes, out_boxes, out_scores, out_labels, out_indices, out_counts, 7, stream);
}

extern "C" int yb_net_detect_phases(yb_net* net, const float* images, const float* anchors9x2, int max_boxes,
                                    float score_thresh, float iou_thresh, void* workspace, size_t workspace_bytes,
                                    float* boxes, float* out_boxes, float* out_scores, int32_t* out_labels,
                                    int32_t* out_indices, int32_t* out_counts, int phases, void* stream) {
  YB_REQUIRE(net && net->act && net->par, "detect: net not bound");
  YB_REQUIRE(images && anchors9x2 && workspace && boxes && out_counts, "detect: null pointer");
  YB_REQUIRE(max_boxes >= 0, "detect: bad max_boxes");
  if (!yb_net_detect_supported(net)) {
    set_error("detect: no fused-decode kernel for %d classes (use forward + predict + nms)", net->class_num);
    return YB_ERR_UNSUPPORTED;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = net->n, B = det_num_boxes(net), C = net->class_num;
  if (max_boxes == 0) {
    YB_CUDA(cudaMemsetAsync(out_counts, 0, sizeof(int32_t) * n, st));
    return YB_OK;
  }
  YB_REQUIRE(out_boxes && out_scores && out_labels && out_indices, "detect: null pointer");
  YB_REQUIRE(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)out_boxes & 15) == 0, "detect: boxes must be 16-byte aligned");
  int* cand_count; float* cand_score; int* cand_idx;
  int rc = nms_candidate_buffers(workspace, workspace_bytes, n, B, C, max_boxes, &cand_count, &cand_score, &cand_idx);
  if (rc) return rc;
  if (phases & 1) YB_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)n * C, st));
  DetParams det[3];
  int off = 0;
  for (int s = 0; s < 3; ++s) {
    DetParams& d = det[s];
    memset(&d, 0, sizeof(d));
    const int gh = net->h / (32 >> s), gw = net->w / (32 >> s);
    d.boxes = boxes; d.cand_count = cand_count; d.cand_score = cand_score; d.cand_idx = cand_idx;
    d.B = B; d.C = C; d.E = 5 + C; d.box_off = off;
    off += 3 * gh * gw;
    d.ratio_h = (float)((double)net->h / (double)gh);        // model.py:91 (float64 divide, cast to f32)
    d.ratio_w = (float)((double)net->w / (double)gw);
    for (int a = 0; a < 3; ++a) {                            // anchor groups 6:9, 3:6, 0:3 (model.py:148-150)
      d.anchor_w[a] = anchors9x2[2 * ((2 - s) * 3 + a)];
      d.anchor_h[a] = anchors9x2[2 * ((2 - s) * 3 + a) + 1];
    }
    d.thr = score_thresh;
    // logits below logit(thr) - 0.01 give sigmoid < thr by > 1e-3 (float error ~1e-7): exact pre-filter
    d.logit_lo = (score_thresh > 0.f && score_thresh < 1.f)
                     ? (float)(log((double)score_thresh / (1.0 - (double)score_thresh)) - 0.01) : -INFINITY;
    d.on = 1;
  }
  if (phases & 1) {
    rc = forward_layers_impl(net, images, nullptr, nullptr, nullptr, 0, 0, det, stream);
    if (rc) return rc;
  }
  if (phases & 2) {
    rc = forward_layers_impl(net, images, nullptr, nullptr, nullptr, 1, 1 << 30, det, stream);
    if (rc) return rc;
  }
  if (!(phases & 4)) return YB_OK;
  return nms_select_gather(boxes, n, B, C, max_boxes, iou_thresh, workspace, workspace_bytes, out_boxes, out_scores,
                           out_labels, out_indices, out_counts, st);
}

extern "C" int yb_net_layer_output(const yb_net* net, int layer, void** ptr, int* ld, int* dtype) {
  YB_REQUIRE(net && net->act && layer >= 0 && layer < (int)net->layers.size() && ptr && ld && dtype,
             "layer_output: bad argument");
  const Layer& L = net->layers[layer];
  *ptr = ten_ptr(net, L.out);
  *ld = net->bufs[L.out.buf].ld;
  *dtype = L.out_fp32 ? YB_F32 : net->dtype;
  return YB_OK;
}

extern "C" int yb_net_forward_launches(const yb_net* net) { return net ? (int)net->layers.size() : 0; }
