// Shared between conv_igemm.cu (kernel + launcher) and net.cu (network plan).
#pragma once
#include "common.cuh"
#include "decode.cuh"

namespace yb {

struct ConvParams {
  int M, P, Q;            // output pixels (n*P*Q), output height, width
  int cout;               // valid output channels
  int cin;                // input channels (K per filter tap)
  int ksize, stride, pad;
  int kh, kw;             // filter window (ksize x ksize for the forward convs; 1x1 .. 2x2 for the stride-2 dgrad classes)
  int scatter;            // 0, or 1 + 2a + b: output pixel (p, q) is stored at (2p + a, 2q + b) of a [n, 2P, 2Q] grid
  int im2col;             // 1: A via im2col TMA, 0: A via 2D tiled TMA
  int two_cta;            // 1: cta_group::2 kernel (256-row tiles per CTA pair)
  int b_resident;         // 1-CTA kernel: the whole [BN, K] weight tile stays in shared memory (single n-tile, fits)
  int epi_groups;         // 1 | 2 sets of four epilogue warps (2: two tiles' epilogues run concurrently; short-K layers)
  int kps;                // 1-CTA kernel: k-blocks per barrier phase (one empty/full handshake per kps k-blocks)
  int dbg;                // timing experiments only (YB_CONV_DBG bitmask: 1 skip A loads, 2 skip B loads, 4 skip MMAs)
  int mc_m, mc_n;         // cluster of mc_m x mc_n pairs with TMA multicast (1,1: plain pair kernel)
  int num_m_tiles, num_n_tiles;
  const float* scale;     // [cout_pad]
  const float* shift;     // [cout_pad]
  void* out;
  long out_ld;
  const void* res;
  long res_ld;
  int out_fp32, leaky, upsample;
  float* stat_sum;        // nullable: BN batch statistics of the raw conv result
  float* stat_sqsum;
  int epi_tma;            // 1: 16-bit output tiles leave through shared memory + TMA stores (tmO), the residual comes in by TMA (tmR)
  CUtensorMap tmO;        // [M, cout] view of the output slice, box 32 rows x 32 channels, SWIZZLE_64B
  CUtensorMap tmR;        // same view of the residual
  DetParams det;          // det.on: decode + NMS candidate filter instead of the fp32 feature-map store (detection heads)
  long long* trace;       // debugging only (yb_debug_set_conv_trace): CTA 0 stamps clock64 at its pipeline events, else NULL
};

int conv_prepare(const yb_conv_desc* d, const void* x, const void* w_packed, const float* scale, const float* shift,
                 const void* res, void* out, float* stat_sum, float* stat_sqsum, CUtensorMap* tmA, CUtensorMap* tmB,
                 ConvParams* p, int* cout_pad_out);
// Window variant used by the stride-2 dgrad: a kh x kw window whose taps sit at offsets (0..kh-1, 0..kw-1) from the
// output pixel (zero-filled past the border), stride 1, output scattered to parity class `scatter`.
// w_packed is [cout_pad][kh*kw*cin].
int conv_prepare_win(const yb_conv_desc* d, int kh, int kw, int scatter, const void* x, const void* w_packed,
                     const float* scale, const float* shift, const void* res, void* out, CUtensorMap* tmA,
                     CUtensorMap* tmB, ConvParams* p, int* cout_pad_out);
int conv_prepare_det(const yb_conv_desc* d, int class_num, const void* x, const void* w_packed, const float* scale,
                     const float* shift, CUtensorMap* tmA, CUtensorMap* tmB, ConvParams* p, int* cout_pad_out);
// halo-tile conv for the Cin <= 64 3x3 layers (csrc/conv_halo.cu)
struct HaloMaps { CUtensorMap plane[4]; CUtensorMap w; CUtensorMap in3d; };
struct HaloParams {
  int n, ho, wo;               // output geometry
  int tiles_x, tiles_y, num_tiles;
  int cout;
  int leaky;
  const float* scale;
  const float* shift;
  const void* res;             // nullable, [n, ho, wo, res_ld]
  long res_ld;
  void* out;                   // [n, ho, wo, out_ld]
  long out_ld;
  // fused stem (darknet53_body/Conv 3->32 computed on the fly as the producer of Conv_1's halo planes)
  const float* stem_w;         // [32][27] float32 OHWI, NULL: plain halo conv
  const float* stem_scale;     // [32]
  const float* stem_shift;     // [32]
  int in_h, in_w;              // image size (= the stem's output size)
  int direct;                  // 1: direct 256-bit register stores (no residual, 32-byte aligned dense rows)
};
bool conv_halo_supported(const yb_conv_desc* d);
int conv_halo_prepare(const yb_conv_desc* d, const void* x, const void* w_packed, const float* scale, const float* shift,
                      const void* res, void* out, HaloMaps* maps, HaloParams* p);
int conv_halo_launch(const yb_conv_desc* d, const HaloMaps& maps, const HaloParams& p, cudaStream_t st);
// stem (3->32, 3x3/1, BN + leaky) fused into Conv_1 (32->64, 3x3/2): d describes Conv_1 (its input = the stem's output,
// which is never written); image float32 [n, d->h, d->w, 3].
int conv_stem_halo_prepare(const yb_conv_desc* d, const float* image, const float* stem_w, const float* stem_scale,
                           const float* stem_shift, const void* w_packed, const float* scale, const float* shift, void* out,
                           HaloMaps* maps, HaloParams* p);
int conv_stem_halo_launch(const yb_conv_desc* d, const HaloMaps& maps, const HaloParams& p, cudaStream_t st);
int conv_launch(int dtype, int cout_pad, const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvParams& p,
                cudaStream_t st);

}  // namespace yb
