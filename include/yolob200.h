/* yolob200.h — C ABI of libyolob200.so: the B200-native (sm_100a) replacement for the
 * TensorFlow ops that wizyoung/YOLOv3_TensorFlow's hot path builds its graph from.
 *
 * The reference has no FFI layer of its own (it is pure Python over TensorFlow), so
 * each entry point below names the reference *call site* (file:line, relative to the
 * reference repo) whose TF ops it replaces.  The Python package
 * `yolov3_tensorflow_b200` binds these with ctypes and re-exposes the reference's
 * Python API (model.yolov3, utils.nms_utils.gpu_nms, utils.misc_utils.load_weights).
 *
 * Conventions
 *   - every function returns YB_OK (0) or a negative yb_status; text via
 *     yb_last_error_string() (thread-local).
 *   - the CALLER owns every buffer (inputs, outputs, workspaces, arenas); nothing is
 *     allocated or freed on the device behind the caller's back.
 *   - all work is enqueued on the cudaStream_t passed as `stream` (void*); no entry
 *     point synchronises the device unless its comment says so.
 *   - device pointers unless marked "host".  Activations are NHWC.
 */
#ifndef YOLOB200_H_
#define YOLOB200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum yb_status {
  YB_OK = 0,
  YB_ERR_INVALID_ARGUMENT = -1,
  YB_ERR_CUDA = -2,
  YB_ERR_UNSUPPORTED = -3,
  YB_ERR_WORKSPACE = -4
} yb_status;

typedef enum yb_dtype { YB_F16 = 0, YB_BF16 = 1, YB_F32 = 2 } yb_dtype;

/* layout of a conv weight tensor handed to the packer */
typedef enum yb_wlayout {
  YB_W_HWIO = 0, /* TensorFlow variable layout [kh,kw,Cin,Cout] (utils/misc_utils.py:120)   */
  YB_W_OIHW = 1, /* darknet .weights stream layout (Cout,Cin,kh,kw) (utils/misc_utils.py:117) */
  YB_W_OHWI = 2  /* engine layout [Cout,kh,kw,Cin] (K-major for the implicit GEMM)           */
} yb_wlayout;

int yb_version(void);
const char* yb_last_error_string(void);
/* Runtime switches for A/B experiments and tests ("YB_CONV_MODE", "YB_CONV_EPI", ...).  The table is
 * seeded once from the equally named environment variables when the library is first used; afterwards only
 * yb_set_option() changes it (value NULL or "" = default).  No entry point calls getenv() on its own. */
int yb_set_option(const char* key, const char* value);
const char* yb_get_option(const char* key);
/* device 0..: SM count and compute capability of the current device. */
int yb_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ---------------------------------------------------------------------------------
 * Convolution  (replaces slim.conv2d + slim.batch_norm + leaky_relu + tf.add +
 * tf.pad + resize_nearest_neighbor + tf.concat: utils/layer_utils.py:9-22,30,82-87,
 * model.py:43-49,55-57,62,72)
 * --------------------------------------------------------------------------------- */
typedef struct yb_conv_desc {
  int n, h, w;      /* input batch / height / width                                         */
  int cin, cout;    /* real channel counts (cout = 255 for the detection heads)             */
  int ksize;        /* 1 or 3                                                               */
  int stride;       /* 1 (SAME) or 2 (pad 1 each side then VALID — darknet rule)            */
  int in_ld;        /* elements between consecutive input pixels  (>= cin; concat slices)   */
  int out_ld;       /* elements between consecutive output pixels (>= cout)                 */
  int res_ld;       /* same for the residual input (ignored when res == NULL)               */
  int dtype;        /* yb_dtype of x / w / res (and of out unless out_fp32)                 */
  int out_fp32;     /* 1: write float32 (detection heads), 0: write `dtype`                 */
  int leaky;        /* 1: leaky_relu(0.1) after scale/shift                                 */
  int upsample2x;   /* 1: nearest-neighbour 2x: every output pixel is stored to its 4 places
                          in an [n, 2*ho, 2*wo, out_ld] buffer (model.py:61,71)             */
} yb_conv_desc;

/* Epilogue note (round 2): 16-bit outputs leave each CTA through swizzled shared memory and cp.async.bulk.tensor
 * (TMA) stores, the residual arrives by TMA into the same staging tile; scale = shift = NULL means identity.
 *
 * out[p, co] = act( scale[co] * sum_{r,s,ci} x[p*stride + (r,s) - pad, ci] * w[co,r,s,ci] + shift[co] ) (+ res[p, co])
 *   x        [n,h,w,in_ld]   dtype
 *   w_packed [cout_pad, ksize, ksize, cin] dtype, cout_pad = yb_conv_cout_pad(cout) (zero rows beyond cout)
 *   scale, shift  float32 [cout_pad]   (BN folded: gamma/sqrt(var+eps), beta-mean*scale; heads: 1, bias)
 *   res      nullable, [n,ho,wo,res_ld] dtype — added AFTER the activation (utils/layer_utils.py:30)
 *   out      [n,ho,wo,out_ld] (or the 2x-upsampled buffer)
 *   stat_sum/stat_sqsum nullable float32 [cout_pad]: when given, the per-channel sum and sum of squares of
 *            the raw convolution result (before scale/shift) are atomically accumulated (BN batch statistics).
 * Requires cin % 32 == 0 (the 3-channel stem has its own entry point).  tcgen05 implicit GEMM. */
int yb_conv2d_fwd(const yb_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                  const float* shift, const void* res, void* out, float* stat_sum, float* stat_sqsum,
                  void* stream);
int yb_conv_cout_pad(int cout);
/* Profiling aid (tools/conv_trace.py): convs prepared after this call make CTA 0 stamp clock64 at its pipeline events
 * into buf ([10 warps][64 tile iterations][32 slots] + 2 int64, device memory, caller-zeroed); NULL switches it off. */
int yb_debug_set_conv_trace(long long* buf);

/* First layer (darknet53_body/Conv, 3->32, 3x3 s1; utils/layer_utils.py:35): float32 NHWC image in,
 * `dtype` NHWC out.  w is OHWI float32 [32,3,3,3]. */
int yb_stem_conv_fwd(const float* x, const float* w_ohwi, const float* scale, const float* shift, int n, int h,
                     int w, int cout, int dtype, int leaky, void* out, void* stream);

/* Thin-layer fast paths (HBM-bound layers at the top of Darknet-53; see csrc/conv_thin.cu): same contract as
 * yb_conv2d_fwd / yb_stem_conv_fwd, restricted to 3x3, cin = 32, cout in {32,64} (no statistics, 16-bit output),
 * resp. the 3->32 stem.  Halo tile + resident weights in shared memory, warp-level tensor path. */
int yb_conv3x3_thin_fwd(const yb_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                        const float* shift, const void* res, void* out, void* stream);
int yb_stem_conv_fwd_tc(const float* x, const float* w_ohwi, const float* scale, const float* shift, int n, int h,
                        int w, int dtype, int leaky, void* out, void* stream);
/* ... and, for the training forward, also ACCUMULATING the per-channel sum / sum of squares of the stored outputs
 * (fp32 [32] each, zeroed by the caller; both NULL = yb_stem_conv_fwd_tc): the batch statistics of
 * slim.batch_norm(is_training=True) (reference model.py:35-41) without a second pass over the tensor. */
int yb_stem_conv_fwd_tc_stats(const float* x, const float* w_ohwi, const float* scale, const float* shift, int n, int h,
                              int w, int dtype, int leaky, void* out, float* stat_sum, float* stat_sqsum, void* stream);

/* 3x3 convs with cin in {32, 64} and cout in {64, 128} (darknet53_body Conv_1/3/6/8, utils/layer_utils.py:36-44) from
 * a shared-memory HALO tile: one tiled TMA load per 16x8-pixel output tile (four parity planes for stride 2), the nine
 * taps are nine UMMA descriptors into that tile, weights resident in shared memory (csrc/conv_halo.cu).  Same contract
 * as yb_conv2d_fwd without statistics; 16-bit output; (w / stride) % 8 == 0.  yb_conv3x3_halo_supported: 1 if d fits. */
int yb_conv3x3_halo_supported(const yb_conv_desc* d);
int yb_conv3x3_halo_fwd(const yb_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                        const float* shift, const void* res, void* out, void* stream);
/* darknet53_body/Conv (3->32, 3x3/1) fused into darknet53_body/Conv_1 (32->64, 3x3/2) (utils/layer_utils.py:35-36):
 * the stem is computed on the fly as the producer of Conv_1's shared-memory halo planes, its 64 B/pixel output is never
 * written.  image float32 [n, h, w, 3]; d describes Conv_1 (d->h, d->w = image size, cin 32, cout 64, stride 2);
 * stem_w_ohwi float32 [32][27], stem_scale / stem_shift float32 [32] (folded BN); out [n, h/2, w/2, out_ld] 16-bit. */
int yb_stem_conv1_fused_fwd(const yb_conv_desc* d, const float* image, const float* stem_w_ohwi, const float* stem_scale,
                            const float* stem_shift, const void* w_packed, const float* scale, const float* shift, void* out,
                            void* stream);

/* ---------------------------------------------------------------------------------
 * Pre-processing either side of the hot path, on the device
 * --------------------------------------------------------------------------------- */
/* process_box (utils/data_utils.py:51-115) for a batch: ground-truth lists -> y_true_13/26/52.
 *   boxes  float32 [n, vmax, 5] (x_min, y_min, x_max, y_max, mixup weight), labels int32 [n, vmax], counts int32 [n]
 *   (boxes beyond counts[i] are ignored); anchors9x2 host float[18]; y_true_s float32 [n, h/s, w/s, 3, 6 + class_num],
 *   fully overwritten (zeros, mix weight 1, then the boxes in list order: the last box of a slot wins, class bits
 *   accumulate — exactly the reference's loop).  Bit-exact vs the reference.  vmax <= 256. */
int yb_process_box(const float* boxes, const int32_t* labels, const int32_t* counts, int n, int vmax, int img_w,
                   int img_h, int class_num, const float* anchors9x2, float* y_true_1, float* y_true_2,
                   float* y_true_3, void* stream);
/* letterbox_resize(img, new_w, new_h, interp=0) (utils/data_aug.py:274-293) + BGR->RGB + float32 / 255
 * (test_single_image.py:39-46): uint8 BGR [src_h, src_w, 3] (row pitch in bytes) -> float32 RGB [new_h, new_w, 3].
 * yb_letterbox_params returns the host-side scalars of the same call (resize_ratio, resized size, dh, dw). */
int yb_letterbox_params(int src_h, int src_w, int new_h, int new_w, double* resize_ratio, int* resize_h, int* resize_w,
                        int* dh, int* dw);
int yb_letterbox_normalize(const uint8_t* bgr, int src_h, int src_w, long src_pitch_bytes, int new_h, int new_w,
                           float* out_rgb, void* stream);

/* Weight repack (utils/misc_utils.py:114-123 does (Cout,Cin,kh,kw) -> HWIO on the host):
 * src float32 in `layout` -> dst `dtype` (or float32) OHWI [cout_pad,k,k,cin], rows >= cout zeroed. */
int yb_pack_conv_weights(const float* src, int layout, int cout, int cin, int ksize, int cout_pad, int dtype,
                         void* dst, void* stream);
/* BN inference fold (model.py:35-41, eps=1e-5): scale = gamma/sqrt(var+eps), shift = beta - mean*scale. */
int yb_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, int c, float eps,
               float* scale, float* shift, void* stream);

/* ---------------------------------------------------------------------------------
 * Training-mode pieces around the conv (replace slim.batch_norm(is_training=True) and TF autodiff of
 * slim.conv2d / batch_norm / leaky_relu: model.py:35-49, train.py:108-112)
 * --------------------------------------------------------------------------------- */
/* dw[cout, k*k*cin] (fp32, OHWI, ACCUMULATED) += sum over output pixels of dz[p,co] * im2col(x)[p,(r,s,ci)].
 * d describes the FORWARD conv (n,h,w,cin,cout,ksize,stride,in_ld,dtype); x is its input activation,
 * dz [n*ho*wo, dz_ld] the gradient w.r.t. its raw output; dz_dilated=1: dz is stored zero-inserted in an
 * [n, 2ho, 2wo, dz_ld] buffer (what the stride-2 dgrad consumes).  tcgen05, MN-major operands, split over pixels. */
int yb_conv2d_wgrad(const yb_conv_desc* d, const void* x, const void* dz, int dz_ld, int dz_dilated, float* dw,
                    void* stream);
/* wgrad of the 3-channel stem: x float32 [n,h,w,3], dz [n*h*w, 32] 16-bit -> dw [32,3,3,3] accumulated. */
int yb_stem_conv_wgrad(const float* x, const void* dz, int dtype, int n, int h, int w, float* dw, void* stream);
/* Same contract on the warp-level tensor path (image split into 16-bit head + remainder: float32-exact to ~1e-5);
 * yb_stem_conv_wgrad dispatches to it by default. */
int yb_stem_conv_wgrad_tc(const float* x, const void* dz, int dtype, int n, int h, int w, float* dw, void* stream);
/* dgrad weights: dst[ci][r][s][co] = w_ohwi[co][k-1-r][k-1-s][ci] (k_cout >= cout, cin_pad >= cin zero-padded):
 * the data gradient of a stride-1 conv is yb_conv2d_fwd(dz, dst) (stride-2: on the zero-inserted dz). */
int yb_pack_dgrad_weights(const float* w_ohwi, int cout, int cin, int ksize, int k_cout, int cin_pad, int dtype,
                          void* dst, void* stream);
/* Data gradient of a 3x3 STRIDE-2 conv (pad 1 + VALID, utils/layer_utils.py:17-27) without zero insertion: the input
 * pixels are split by parity (a, b) = (row & 1, col & 1); class (a, b) is a (1+a) x (1+b)-tap conv over the plain
 * dz [n, h/2, w/2, dz_ld] whose result is stored at (2i + a, 2j + b) of dx [n, h, w, dx_ld] (+ res at the same pixel,
 * res nullable, may alias dx).  yb_pack_dgrad_weights_s2 lays the four weight matrices [cin_pad][(1+a)(1+b)*k_cout]
 * out back to back (9 * cin_pad * k_cout elements).  `fwd` is the FORWARD conv's descriptor (n, h, w, cin, cout, dtype). */
int yb_pack_dgrad_weights_s2(const float* w_ohwi, int cout, int cin, int k_cout, int cin_pad, int dtype, void* dst,
                             void* stream);
int yb_conv2d_dgrad_s2(const yb_conv_desc* fwd, const void* dz, int dz_ld, int k_cout, const void* w_dgrad_s2,
                       const void* res, int res_ld, void* dx, int dx_ld, void* stream);
/* Host-only: the split-K plan yb_conv2d_wgrad picks for a layer with `tiles` output tiles (tap groups x ci chunks x
 * co tiles) over `num_pixel_blocks` 64-pixel blocks on `sms` SMs — the count that minimises
 * waves x (blocks per CTA + epi_blocks); every block is covered exactly once. */
int yb_wgrad_split_plan(long num_pixel_blocks, long tiles, int sms, int epi_blocks, long* splits,
                        long* blocks_per_split);
/* BN batch statistics -> scale/shift for bn_act_apply, saved mean/invstd for the backward, moving-stat update
 * (biased variance normalises, unbiased variance feeds the moving average; moving_* nullable). */
int yb_bn_finalize(const float* sum, const float* sqsum, long count, int c, const float* gamma, const float* beta,
                   float eps, float decay, float* moving_mean, float* moving_var, float* scale, float* shift,
                   float* save_mean, float* save_invstd, void* stream);
/* out = leaky(z*scale+shift) (+res); z/res/out 16-bit [n*h*w, ld]; upsample2x stores every row to its 4 places. */
int yb_bn_act_apply(const void* z, long z_ld, const float* scale, const float* shift, const void* res, long res_ld,
                    void* out, long out_ld, int n, int h, int w, int c, int dtype, int leaky, int upsample2x,
                    void* stream);
/* yb_bn_finalize + yb_bn_act_apply in ONE launch (count = n*h*w; same arguments, bit-identical results): the training
 * forward of a BN conv (slim.batch_norm(is_training=True) + leaky_relu, reference model.py:35-41). */
int yb_bn_stats_act_apply(const void* z, long z_ld, const float* sum, const float* sqsum, const float* gamma,
                          const float* beta, float eps, float decay, float* moving_mean, float* moving_var,
                          float* scale, float* shift, float* save_mean, float* save_invstd, const void* res,
                          long res_ld, void* out, long out_ld, int n, int h, int w, int c, int dtype, int leaky,
                          int upsample2x, void* stream);
/* dgamma/dbeta (fp32 [c], overwritten) from dA (gradient w.r.t. the layer output; upsample2x: summed over the
 * 4 copies) and the saved z.  workspace: NULL (atomic accumulation) or yb_bn_bwd_reduce_workspace_bytes() bytes,
 * zero-initialised once (two-stage deterministic reduction, no same-address atomics). */
int yb_bn_bwd_reduce_workspace_bytes(size_t* bytes);
int yb_bn_bwd_reduce(const void* dA, long dA_ld, const void* z, long z_ld, const float* scale, const float* shift,
                     const float* save_mean, const float* save_invstd, int n, int h, int w, int c, int dtype,
                     int leaky, int upsample2x, float* dgamma, float* dbeta, void* workspace, void* stream);
/* dz = gamma*invstd*(dact - dbeta/M - zhat*dgamma/M); dilate2x stores row (p,q) at (2p,2q) of [n,2h,2w,dz_ld]. */
int yb_bn_bwd_apply(const void* dA, long dA_ld, const void* z, long z_ld, const float* gamma, const float* scale,
                    const float* shift, const float* save_mean, const float* save_invstd, const float* dgamma,
                    const float* dbeta, int n, int h, int w, int c, int dtype, int leaky, int upsample2x,
                    int dilate2x, void* dz, long dz_ld, void* stream);
/* out[c] (fp32, overwritten) = column sums of x [rows, ld] (bias gradient of the detection convs). */
int yb_col_sum(const void* x, long ld, long rows, int c, int dtype, float* out, void* stream);
/* column sums and sums of squares (BN batch statistics of the stem's raw output). */
int yb_col_stats(const void* x, long ld, long rows, int c, int dtype, float* sum, float* sqsum, void* stream);

/* ---------------------------------------------------------------------------------
 * Decode  (replaces the ~40 elementwise TF ops of model.py:82-190 and the caller's
 * pred_scores = pred_confs * pred_probs, test_single_image.py:55)
 * --------------------------------------------------------------------------------- */
/* reorg_layer (model.py:82-137) for one scale.  anchors3x2: host float[6] (w,h) pixels.
 * xy_offset [gh,gw,1,2], boxes [n,gh,gw,3,4] (cx,cy,w,h px), conf_logits [n,gh,gw,3,1],
 * prob_logits [n,gh,gw,3,C]; any output may be NULL. */
int yb_reorg_layer(const float* feature_map, int n, int gh, int gw, int img_h, int img_w, int class_num,
                   const float* anchors3x2, float* xy_offset, float* boxes, float* conf_logits,
                   float* prob_logits, void* stream);
/* predict (model.py:140-190) over the three scales (/32,/16,/8).  anchors9x2: host float[18].
 * boxes [n,B,4] xyxy, confs [n,B,1] (nullable), probs [n,B,C] (nullable), scores [n,B,C] = conf*prob (nullable),
 * B = 3*(h/32*w/32 + h/16*w/16 + h/8*w/8). */
int yb_predict(const float* fm1, const float* fm2, const float* fm3, int n, int img_h, int img_w, int class_num,
               const float* anchors9x2, float* boxes, float* confs, float* probs, float* scores, void* stream);

/* ---------------------------------------------------------------------------------
 * NMS  (replaces utils/nms_utils.py:8-48: greater_equal + boolean_mask x2 +
 * tf.image.non_max_suppression + gather x3 per class + concat x3)
 * --------------------------------------------------------------------------------- */
int yb_nms_workspace_bytes(int n_images, int num_boxes, int num_classes, int max_boxes, size_t* bytes);
/* Per image and class: keep score >= score_thresh, greedy NMS (IoU > iou_thresh suppresses; TF CPU-kernel
 * arithmetic), at most max_boxes per class; classes concatenated ascending, descending score inside a class.
 *   boxes [n,B,4] xyxy, scores [n,B,C]
 *   out_boxes [n, C*max_boxes, 4], out_scores / out_labels / out_indices [n, C*max_boxes]
 *   out_indices = index of the kept box in the ORIGINAL [B] axis (not exposed by the reference)
 *   out_counts [n] (device int32) = K of each image. */
int yb_nms(const float* boxes, const float* scores, int n_images, int num_boxes, int num_classes, int max_boxes,
           float score_thresh, float iou_thresh, void* workspace, size_t workspace_bytes, float* out_boxes,
           float* out_scores, int32_t* out_labels, int32_t* out_indices, int32_t* out_counts, void* stream);

/* ---------------------------------------------------------------------------------
 * Loss  (replaces model.py:192-304 loss_layer, :307-345 box_iou, :348-365 compute_loss and
 * the part of TF autodiff (train.py:112) that differentiates them)
 * --------------------------------------------------------------------------------- */
int yb_loss_workspace_bytes(int n, int gh, int gw, size_t* bytes);
/* One scale.  feature_map [n,gh,gw,3*(5+C)] f32 logits, y_true [n,gh,gw,3,5+C+1] f32 (utils/data_utils.py:51-115
 * format: cx,cy,w,h px | obj | one-hot | mix-up weight), anchors3x2 host float[6].
 * loss4 (device double[4]: xy, wh, conf, class) is ACCUMULATED into (zero it first; every term already
 * carries the 1/N of model.py:276-302, N = 1/inv_batch).
 * dfm (nullable) receives d(total)/d(feature_map): dfm_dtype YB_F32 -> same layout as feature_map;
 * YB_F16/YB_BF16 -> [n*gh*gw, dfm_ld] rows (dfm_ld >= 3*(5+C), padding columns zeroed) for the backward GEMMs. */
/* loss_scale (> 0; 1 = off) multiplies the stored gradient only (not the loss values): the fp16 backward chain
 * needs it to keep small gradients above the subnormal range; the optimizer divides it out again
 * (yb_optimizer.grad_scale) and skips a step whose gradient is non-finite. */
int yb_loss_layer(const float* feature_map, const float* y_true, int n, int gh, int gw, int img_h, int img_w,
                  int class_num, const float* anchors3x2, int use_label_smooth, int use_focal_loss,
                  float inv_batch, float loss_scale, void* workspace, size_t workspace_bytes, double* loss4,
                  void* dfm, int dfm_dtype, int dfm_ld, void* stream);
/* out5 (device float[5]) = [total, xy, wh, conf, class] (model.py:364-365). */
int yb_loss_finalize(const double* loss4, float* out5, void* stream);
/* box_iou (model.py:307-345): pred_boxes [P,4], true_boxes [V,4] (cx,cy,w,h) -> iou [P,V]. */
int yb_box_iou(const float* pred_boxes, const float* true_boxes, long num_pred, int num_true, float* iou,
               void* stream);

/* ---------------------------------------------------------------------------------
 * Network plan: the 75-conv Darknet-53 + YOLOv3 head of model.py:30-80 for a fixed
 * (batch, H, W, dtype).  Holds tensor maps and the layer schedule; buffers are the
 * caller's: one activation arena and one parameter arena.
 * --------------------------------------------------------------------------------- */
typedef struct yb_net yb_net;

typedef struct yb_layer_info {
  int index;          /* creation order == TF variable order == darknet .weights order  */
  int cin, cout, ksize, stride;
  int has_bn;         /* 0 for the three detection convs (bias, linear)                  */
  int in_h, in_w, out_h, out_w;
  int is_head;        /* 1: under yolov3_head, 0: darknet53_body                        */
  int scope_index;    /* k of "Conv_k" inside its variable scope                        */
  int upsample2x;     /* 1: output is stored 2x nearest-neighbour upsampled (model.py:61,71) */
} yb_layer_info;

int yb_net_create(yb_net** net, int class_num, int n, int h, int w, int dtype, int training);
int yb_net_destroy(yb_net* net);
int yb_net_num_layers(const yb_net* net);
int yb_net_layer_info(const yb_net* net, int layer, yb_layer_info* info);
int yb_net_arena_bytes(const yb_net* net, size_t* activation_bytes, size_t* param_bytes);
/* Bind caller-owned arenas (256-byte aligned).  Must be called before set_params/forward.  The PARAMETER arena layout
 * depends only on (class_num, dtype, training): plans of different batch / image sizes may share one parameter arena
 * (master weights, 16-bit copies, folded BN, optimizer slots) — this is how one model serves multi-scale training
 * (train.py:47-49 multi_scale_train) with a single set of weights and a single optimizer state.  Binding fills
 * constants and this plan's activation-arena scratch on `stream`; it never touches weights or optimizer slots. */
int yb_net_bind(yb_net* net, void* activation_arena, size_t activation_bytes, void* param_arena,
                size_t param_bytes, void* stream);
/* Re-fold every BN layer's (gamma, beta, moving mean, moving variance) into the inference scale/shift (needed after
 * training steps made through ANOTHER plan that shares the parameter arena; a plan refolds by itself after its own). */
int yb_net_refold_bn(yb_net* net, void* stream);
/* Upload one conv's parameters (device float32 pointers).  BN layers: gamma,beta,mean,var (bias NULL);
 * detection convs: bias (others NULL).  Repacks/folds on `stream`. */
int yb_net_set_conv_params(yb_net* net, int layer, const float* w, int layout, const float* gamma,
                           const float* beta, const float* mean, const float* var, const float* bias,
                           void* stream);
/* forward (model.py:30-80), inference mode: images float32 [n,h,w,3] -> fm1 [n,h/32,w/32,D],
 * fm2 [n,h/16,w/16,D], fm3 [n,h/8,w/8,D] float32, D = 3*(5+class_num). */
int yb_net_forward(yb_net* net, const float* images, float* fm1, float* fm2, float* fm3, void* stream);
/* The whole detection pipeline of test_single_image.py:50-57 in one call:
 *   forward (model.py:30-80) -> predict (model.py:140-190) -> pred_scores = confs * probs (test_single_image.py:55)
 *   -> gpu_nms per image (utils/nms_utils.py:8-48; max_boxes per class, score >= score_thresh, IoU > iou_thresh suppresses).
 * The decode and the score filter run INSIDE the epilogues of the three detection-head convs (fp32 accumulators ->
 * boxes + per-(image, class) candidate lists); the feature maps and the [n, B, C] score tensor are never written.
 * Results are bit-identical to yb_net_forward + yb_predict + yb_nms.
 *   anchors9x2 host float[18] (w,h pixels, small -> large);  boxes [n, B, 4] float32 out: every decoded box
 *   (xmin,ymin,xmax,ymax), B = 3*(h/32*w/32 + h/16*w/16 + h/8*w/8);  workspace >= yb_net_detect_workspace_bytes;
 *   out_* as yb_nms (fixed shape [n, class_num*max_boxes(,4)], out_counts [n]).
 * yb_net_detect_supported: 1 when a fused kernel exists for the plan's class count (80 and 20), else 0 — callers
 * then use the three separate calls. */
int yb_net_detect_supported(const yb_net* net);
int yb_net_detect_workspace_bytes(const yb_net* net, int max_boxes, size_t* bytes);
int yb_net_detect(yb_net* net, const float* images, const float* anchors9x2, int max_boxes, float score_thresh,
                  float iou_thresh, void* workspace, size_t workspace_bytes, float* boxes, float* out_boxes,
                  float* out_scores, int32_t* out_labels, int32_t* out_indices, int32_t* out_counts, void* stream);
/* yb_net_detect split for benchmarks that bracket its parts with their own events: phases is a bit mask,
 * 1 = candidate-list reset + stem (layer 0), 2 = the 74 tensor-core convs (decode fused into the heads), 4 = NMS
 * selection + gather.  yb_net_detect == phases 7. */
int yb_net_detect_phases(yb_net* net, const float* images, const float* anchors9x2, int max_boxes, float score_thresh,
                         float iou_thresh, void* workspace, size_t workspace_bytes, float* boxes, float* out_boxes,
                         float* out_scores, int32_t* out_labels, int32_t* out_indices, int32_t* out_counts, int phases,
                         void* stream);
/* Same, restricted to layers [first, last] (creation order) — lets a benchmark bracket the CUDA-core stem
 * (layer 0) and the tensor-core convs (1..74) with its own events. */
int yb_net_forward_layers(yb_net* net, const float* images, float* fm1, float* fm2, float* fm3, int first,
                          int last, void* stream);
/* ---- training plan (yb_net_create(..., training=1)); one reference training step (train.py:105-115) is
 *      yb_net_train_fwd_bwd -> [all-reduce of yb_net_grad_buffer across ranks] -> yb_net_train_update ---- */
/* forward with BN batch statistics (updating the moving statistics with `bn_decay`, train.py:108-109) ->
 * compute_loss (model.py:348-365; loss4 = device double[4] xy,wh,conf,class, overwritten) -> backward into the
 * flat gradient buffer (data term only).  YB_TRAIN_FORWARD_ONLY stops after the forward (y_true*, loss4 may be NULL).
 * y_true_k: [n, g_k, g_k, 3, 5+C+1] float32 for the /32, /16, /8 maps; anchors9x2 host float[18].
 * fm1..3 nullable (then the arena-owned float32 outputs are used). */
/* flags: YB_TRAIN_FORWARD_ONLY stops after the forward; YB_TRAIN_BN_FROZEN normalises with the moving statistics
 * (forward(is_training=False) inside the training graph: fine-tuning with frozen BN; statistics are constants of the
 * backward pass and are not updated) — per-image results then do not depend on the rest of the batch, which is what
 * makes an N-rank data-parallel step bit-comparable with a 1-rank step on the concatenated batch. */
enum { YB_TRAIN_FORWARD_ONLY = 1, YB_TRAIN_BN_FROZEN = 2, YB_TRAIN_NO_BACKWARD = 4 };
int yb_net_train_fwd_bwd(yb_net* net, const float* images, const float* y_true_1, const float* y_true_2,
                         const float* y_true_3, const float* anchors9x2, int use_label_smooth, int use_focal_loss,
                         float bn_decay, float loss_scale, float* fm1, float* fm2, float* fm3, double* loss4,
                         int flags, void* stream);
/* Bucketed data parallelism: yb_net_train_fwd_bwd(..., flags | YB_TRAIN_NO_BACKWARD) stops after the
 * loss; yb_net_train_backward then runs the backward of layers last_layer .. first_layer (descending, same flags), and
 * yb_net_grad_range returns the contiguous slice of the flat gradient those layers own — the caller issues the
 * all-reduce of a finished bucket (detection heads first) while the next bucket's backward is running. */
int yb_net_train_backward(yb_net* net, const float* images, int first_layer, int last_layer, int flags, void* stream);
int yb_net_grad_range(yb_net* net, int first_layer, int last_layer, float** ptr, size_t* count);
/* the flat float32 gradient of all 222 trainable tensors (creation order: per conv w [OHWI], then gamma, beta
 * | bias; each padded to 4 floats) — the buffer a data-parallel wrapper all-reduces. */
int yb_net_grad_buffer(yb_net* net, float** ptr, size_t* count);
/* The optimizers of utils/misc_utils.py:151-161 (config_optimizer) with TensorFlow 1.x update rules. */
typedef enum yb_opt_kind {
  YB_OPT_SGD = 0,      /* GradientDescentOptimizer: w -= lr*g                                                    */
  YB_OPT_MOMENTUM = 1, /* MomentumOptimizer(momentum): v = m*v + g; w -= lr*v  (the reference's default)         */
  YB_OPT_RMSPROP = 2,  /* RMSPropOptimizer(decay, momentum, epsilon=1e-10): ms = d*ms + (1-d)g^2 (ms starts at 1);
                          mom = m*mom + lr*g/sqrt(ms+eps); w -= mom                                              */
  YB_OPT_ADAM = 3      /* AdamOptimizer(beta1=.9, beta2=.999, epsilon=1e-8): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
                          m = b1*m + (1-b1)g; v = b2*v + (1-b2)g^2; w -= lr_t*m/(sqrt(v)+eps)                    */
} yb_opt_kind;
typedef struct yb_optimizer {
  int kind;            /* yb_opt_kind                                                                            */
  float lr;            /* THIS step's learning rate: the host evaluates the schedule (utils/misc_utils.py:129-148,
                          warm-up train.py:93-99; Python: utils.misc_utils.config_learning_rate)                 */
  float grad_scale;    /* multiplies the raw gradient buffer: 1/world_size (data-parallel mean) x 1/loss_scale   */
  float momentum, decay, beta1, beta2, epsilon;
  float weight_decay;  /* slim.l2_regularizer on conv weights only (model.py:49, train.py:78)                    */
  float clip_norm;     /* per-tensor tf.clip_by_norm (train.py:113-114); <= 0: off                               */
} yb_optimizer;
/* g = grad_scale*grad + weight_decay*w (conv weights only); per-tensor clip_by_norm; the optimizer's rule on the fp32
 * master weights and its slots; refresh of the 16-bit compute copies and dgrad weights.  A step whose gradient holds
 * a non-finite value is skipped entirely (yb_net_opt_state: ctrl[2] counts skipped steps, ctrl[1] applied ones). */
int yb_net_train_update(yb_net* net, const yb_optimizer* opt, void* stream);
/* Zero the optimizer slots (rmsprop: mean-square slot = 1 like TF), the gradient buffer and the step counters.
 * Call once when a parameter arena starts training (or the optimizer changes); binding a plan never does this. */
int yb_net_train_reset_state(yb_net* net, int optimizer_kind, void* stream);
/* The optimizer slots: num_slots x count_per_slot floats laid out like the gradient buffer, and int ctrl[3] =
 * {non-finite flag of the running step, updates applied, steps skipped} (checkpoint save/restore of save_optimizer,
 * train.py:101-104,118-121). */
int yb_net_opt_state(yb_net* net, float** slots, size_t* count_per_slot, int* num_slots, int** ctrl);
/* train.py:81 update_part: exclude a conv (weights + gamma/beta | bias) from / include it in the update. */
int yb_net_set_trainable(yb_net* net, int layer, int trainable, void* stream);
/* Re-derive the dgrad weight layouts from the fp32 master weights (after the arena's weights were replaced). */
int yb_net_train_refresh_dgrad(yb_net* net, void* stream);
/* device pointers of one conv's float32 master parameters (w is OHWI [cout,k,k,cin]) / of its gradients. */
int yb_net_get_conv_params(yb_net* net, int layer, float** w_ohwi, float** gamma, float** beta, float** mean,
                           float** var, float** bias);
int yb_net_layer_grad(yb_net* net, int layer, float** dw, float** dgamma, float** dbeta, float** dbias);
/* training scratch of one layer (tests): which = 0 raw conv output z, 1 its gradient dz (zero-inserted at the input
 * resolution for stride-2 layers), 2 gradient w.r.t. the layer output, 3 the layer's input activation.
 * All 16-bit [n, h, w, ld]. */
int yb_net_train_buffer(yb_net* net, int layer, int which, void** ptr, int* ld, int* h, int* w);
/* device pointer + geometry of one layer's output activation (tests / debugging). */
int yb_net_layer_output(const yb_net* net, int layer, void** ptr, int* ld, int* dtype);
/* number of kernels one yb_net_forward enqueues (for bench.py's gpu_launches). */
int yb_net_forward_launches(const yb_net* net);

#ifdef __cplusplus
}
#endif
#endif /* YOLOB200_H_ */
