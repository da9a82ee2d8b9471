// Plan data structures shared by net.cu (schedule + inference) and net_train.cu (training).
#pragma once
#include <vector>

#include "common.cuh"
#include "conv.cuh"
#include "optim.cuh"

namespace yb {

struct Ten {          // a view of an activation: buffer id + channel slice
  int buf = -1;       // -1: network input image
  int off = 0;        // first channel inside the buffer
  int c = 0, h = 0, w = 0;
};

struct Buf {
  int h, w, ld;       // [n, h, w, ld]
  int fp32;           // detection outputs are float32
  size_t offset = 0, bytes = 0;
};

struct Layer {
  yb_layer_info info;
  Ten in, out, res;   // res.buf == -2: none
  bool upsample = false, out_fp32 = false;
  int cout_pad = 0;
  // parameter arena offsets (bytes)
  size_t w_master = 0, w_packed = 0, gamma = 0, beta = 0, mean = 0, var = 0, bias = 0, scale = 0, shift = 0;
  // prepared launch state
  CUtensorMap tmA, tmB;
  ConvParams params;
  bool prepared = false;
  // Cin <= 64 3x3 layers: halo-tile kernel (csrc/conv_halo.cu), inference forward only
  HaloMaps halo_maps;
  HaloParams halo_params;
  yb_conv_desc halo_desc;
  bool halo_ok = false;
  // detection heads: the same conv with the decode + NMS candidate filter fused into its epilogue (yb_net_detect)
  CUtensorMap det_tmA, det_tmB;
  ConvParams det_params;
  int det_cout_pad = 0;
  bool det_ok = false;
  // ---- training plan (net_train.cu) ----
  size_t z_off = 0, dz_off = 0;      // raw conv output z / its gradient (activation arena); dz is zero-inserted for stride 2
  int dz_ld = 0, dz_dilated = 0, k_cout = 0;
  int dgrad_parity = 0;              // stride-2 layer whose dgrad runs as 4 parity-class convs on the plain dz
  CUtensorMap d4_tmA[4], d4_tmB[4];
  ConvParams d4_params[4];
  int d4_cout_pad[4] = {0, 0, 0, 0};
  size_t st_sum = 0, st_sqsum = 0, st_mean = 0, st_invstd = 0, st_scale = 0, st_shift = 0;   // fp32 [cout_pad] each
  size_t w_dgrad = 0;                // [cin_pad, k, k, k_cout] 16-bit (param arena)
  long g_w = -1, g_gamma = -1, g_beta = -1, g_bias = -1;   // float offsets into the flat gradient / velocity buffers
  ConvParams tparams;                // training-mode forward conv (raw z + statistics)
  CUtensorMap d_tmA, d_tmB;          // dgrad (forward kernel on dz with flipped/transposed weights)
  ConvParams dparams;
  int d_cout_pad = 0;
};

}  // namespace yb

struct yb_net {
  int class_num, n, h, w, dtype, training;
  std::vector<yb::Layer> layers;
  std::vector<yb::Buf> bufs;
  int fm_buf[3];
  size_t act_bytes = 0, param_bytes = 0;
  uint8_t* act = nullptr;
  uint8_t* par = nullptr;
  // ---- training plan ----
  std::vector<size_t> gbuf_offset;   // gradient mirror of every 16-bit activation buffer
  size_t dfm_off[3] = {0, 0, 0};     // 16-bit [rows, 256] loss gradients of the three detection maps
  size_t stats_off = 0, stats_bytes = 0;   // per-step zeroed BN sums
  size_t lossws_off = 0, lossws_bytes = 0;
  size_t bnws_off = 0, bnws_bytes = 0;     // two-stage BN-backward reduction scratch (zeroed at bind)
  size_t ones_off = 0, zeros_off = 0;      // fp32 [1024] constants (param arena)
  size_t grad_off = 0, vel_off = 0; long grad_count = 0;   // flat fp32 gradient / optimizer slots (param arena)
  int opt_state_slots = 2;            // slot 1: momentum / adam m; slot 2: rmsprop mean square / adam v
  size_t opt_step_off = 0;            // int ctrl[64]: non-finite flag, updates applied, steps skipped
  size_t opt_tensors_off = 0, opt_chunks_off = 0, opt_norm_off = 0; int num_opt_tensors = 0, num_opt_chunks = 0;
  std::vector<yb::OptTensor> opt_tensors;
  std::vector<yb::OptChunk> opt_chunks;
  size_t pack_jobs_off = 0; int pack_tiles = 0;   // multi-tensor dgrad-weight repack (optim.cuh: PackJob)
  std::vector<yb::PackJob> pack_jobs;
  bool fold_dirty = false;
  float bn_eps = 1e-5f;
  // side stream of the backward pass: layer L's wgrad runs beside its dgrad (net_train.cu); created on first use
  cudaStream_t side_stream = nullptr;
  cudaEvent_t side_fork = nullptr, side_join = nullptr;
  ~yb_net() {
    if (side_fork) cudaEventDestroy(side_fork);
    if (side_join) cudaEventDestroy(side_join);
    if (side_stream) cudaStreamDestroy(side_stream);
  }
};

