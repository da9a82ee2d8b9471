// Training plan: one step of train.py:105-115 on the B200 engine.
//   forward  : conv (raw z + BN batch sums in the epilogue) -> bn_finalize -> bn_act_apply
//              (model.py:30-80 with is_training=True; UPDATE_OPS moving stats, train.py:108-109)
//   loss     : yb_loss_layer x3, gradients written 16-bit into the backward GEMM operands
//   backward : per layer, last to first: bn_bwd_reduce/apply -> wgrad (tcgen05) -> dgrad (forward
//              kernel on dz with flipped/transposed weights; the residual / second-consumer
//              contributions are folded into the dgrad epilogue's residual add)
//   update   : L2 + per-tensor clip_by_norm + momentum over one flat gradient buffer (which the
//              data-parallel wrapper all-reduces between backward and update).
#include <string.h>

#include <algorithm>
#include <map>

#include "net.cuh"

namespace yb {

int conv_prepare(const yb_conv_desc* d, const void* x, const void* w_packed, const float* scale, const float* shift,
                 const void* res, void* out, float* stat_sum, float* stat_sqsum, CUtensorMap* tmA, CUtensorMap* tmB,
                 ConvParams* p, int* cout_pad_out);

static size_t al256(size_t v) { return (v + 255) & ~size_t(255); }

// Extend the arenas computed by Builder::build() with everything training needs.
void train_layout(yb_net* net) {
  const size_t esz = 2;
  size_t o = net->act_bytes;
  // gradient mirrors of the 16-bit activation buffers
  net->gbuf_offset.assign(net->bufs.size(), 0);
  for (size_t b = 0; b < net->bufs.size(); ++b) {
    if (net->bufs[b].fp32) continue;
    net->gbuf_offset[b] = o;
    o = al256(o + net->bufs[b].bytes);
  }
  int head_i = 0;
  for (auto& L : net->layers) {
    const size_t out_rows = (size_t)net->n * L.info.out_h * L.info.out_w;
    L.k_cout = (L.info.cout + 31) / 32 * 32;
    if (L.info.has_bn) {
      L.z_off = o; o = al256(o + out_rows * L.info.cout * esz);
      L.dz_ld = L.info.cout;
      // stride-2 layers: the input gradient is computed per parity class on the plain dz (4 small convs, no zeros);
      // YB_DGRAD_S2=dilated selects the first version (one 3x3 conv over a zero-inserted dz: 4x the MMA work)
      const char* s2 = opt("YB_DGRAD_S2");
      L.dgrad_parity = (L.info.stride == 2 && !(s2 && s2[0] == 'd')) ? 1 : 0;
      L.dz_dilated = L.info.stride == 2 && !L.dgrad_parity;
      const size_t dz_rows = L.dz_dilated ? (size_t)net->n * L.info.in_h * L.info.in_w : out_rows;
      L.dz_off = o; o = al256(o + dz_rows * L.dz_ld * esz);
    } else {
      L.dz_ld = L.k_cout;                      // 255 -> 256, zero padded by the loss kernel
      L.dz_off = o; o = al256(o + out_rows * L.dz_ld * esz);
      net->dfm_off[head_i++] = L.dz_off;
    }
  }
  // per-step-zeroed BN sums, then the other per-channel scratch
  net->stats_off = o;
  for (auto& L : net->layers) {
    if (!L.info.has_bn) continue;
    L.st_sum = o; o += (size_t)L.cout_pad * 4;
    L.st_sqsum = o; o += (size_t)L.cout_pad * 4;
  }
  o = al256(o);
  net->stats_bytes = o - net->stats_off;
  for (auto& L : net->layers) {
    if (!L.info.has_bn) continue;
    L.st_mean = o; o = al256(o + (size_t)L.cout_pad * 4);
    L.st_invstd = o; o = al256(o + (size_t)L.cout_pad * 4);
    L.st_scale = o; o = al256(o + (size_t)L.cout_pad * 4);
    L.st_shift = o; o = al256(o + (size_t)L.cout_pad * 4);
  }
  size_t lw = 0;
  for (int s = 0; s < 3; ++s) {
    size_t b = 0;
    const int div = 32 >> s;
    yb_loss_workspace_bytes(net->n, net->h / div, net->w / div, &b);
    lw = std::max(lw, b);
  }
  net->lossws_off = o; net->lossws_bytes = lw; o = al256(o + lw);
  yb_bn_bwd_reduce_workspace_bytes(&net->bnws_bytes);
  net->bnws_off = o; o = al256(o + net->bnws_bytes);
  net->act_bytes = o;

  // ---- parameter arena ----
  o = net->param_bytes;
  net->ones_off = o; o = al256(o + 1024 * 4);
  net->zeros_off = o; o = al256(o + 1024 * 4);
  for (auto& L : net->layers) {
    if (L.info.index == 0) continue;
    const size_t cin_pad = yb_conv_cout_pad(L.info.cin);
    L.w_dgrad = o;
    o = al256(o + cin_pad * L.info.ksize * L.info.ksize * L.k_cout * esz);
  }
  long g = 0;
  auto take = [&](long n) { long at = g; g += (n + 3) / 4 * 4; return at; };
  for (auto& L : net->layers) {
    L.g_w = take((long)L.info.cout * L.info.ksize * L.info.ksize * L.info.cin);
    if (L.info.has_bn) { L.g_gamma = take(L.info.cout); L.g_beta = take(L.info.cout); }
    else L.g_bias = take(L.info.cout);
  }
  net->grad_count = g;
  net->grad_off = o; o = al256(o + (size_t)g * 4);
  net->vel_off = o; o = al256(o + (size_t)g * 4 * net->opt_state_slots);
  // optimizer tables
  net->opt_tensors.clear(); net->opt_chunks.clear();
  const long CH = 1 << 16;
  auto add = [&](long n, int l2) {
    OptTensor t; memset(&t, 0, sizeof(t));
    t.n = n; t.l2 = l2; t.trainable = 1;
    const int id = (int)net->opt_tensors.size();
    net->opt_tensors.push_back(t);
    for (long b = 0; b < n; b += CH) { OptChunk c; c.tensor = id; c.begin = b; c.end = std::min(n, b + CH); net->opt_chunks.push_back(c); }
  };
  for (auto& L : net->layers) {
    add((long)L.info.cout * L.info.ksize * L.info.ksize * L.info.cin, 1);
    if (L.info.has_bn) { add(L.info.cout, 0); add(L.info.cout, 0); } else add(L.info.cout, 0);
  }
  net->num_opt_tensors = (int)net->opt_tensors.size();
  net->num_opt_chunks = (int)net->opt_chunks.size();
  net->opt_tensors_off = o; o = al256(o + net->opt_tensors.size() * sizeof(OptTensor));
  net->opt_chunks_off = o; o = al256(o + net->opt_chunks.size() * sizeof(OptChunk));
  net->opt_norm_off = o; o = al256(o + net->opt_tensors.size() * 4);
  net->opt_step_off = o; o = al256(o + 256);
  net->pack_jobs_off = o; o = al256(o + net->layers.size() * sizeof(PackJob));
  net->param_bytes = o;
}

static void* ten_ptr2(const yb_net* net, const Ten& t) {
  const Buf& b = net->bufs[t.buf];
  return net->act + b.offset + (size_t)t.off * 2;
}
static void* gten_ptr(const yb_net* net, const Ten& t) {
  return net->act + net->gbuf_offset[t.buf] + (size_t)t.off * 2;
}
static float* fpar(const yb_net* net, size_t off) { return reinterpret_cast<float*>(net->par + off); }
static float* fact(const yb_net* net, size_t off) { return reinterpret_cast<float*>(net->act + off); }
static float* gradp(const yb_net* net, long idx) { return reinterpret_cast<float*>(net->par + net->grad_off) + idx; }

__global__ void fill_f32_kernel(float* p, long n, float v) {
  for (long i = blockIdx.x * (long)blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) p[i] = v;
}

// Called from yb_net_bind for training plans.  Everything is enqueued on the caller's stream.  The optimizer state
// (velocity) lives in the parameter arena, which several plans of different shapes may share: binding never touches it —
// yb_net_train_reset_state() zeroes it once, when the arena is created.
int train_bind(yb_net* net, cudaStream_t st) {
  // constants (idempotent) + this plan's activation-arena scratch
  fill_f32_kernel<<<4, 256, 0, st>>>(fpar(net, net->ones_off), 1024, 1.0f);
  YB_CUDA(cudaGetLastError());
  YB_CUDA(cudaMemsetAsync(net->par + net->zeros_off, 0, 1024 * 4, st));
  YB_CUDA(cudaMemsetAsync(net->act + net->bnws_off, 0, net->bnws_bytes, st));
  const float* ones = fpar(net, net->ones_off);
  const float* zeros = fpar(net, net->zeros_off);
  for (auto& L : net->layers) {
    if (L.dz_dilated)
      YB_CUDA(cudaMemsetAsync(net->act + L.dz_off, 0, (size_t)net->n * L.info.in_h * L.info.in_w * L.dz_ld * 2, st));
  }
  // ---- training-mode forward convs: raw z + statistics ----
  for (size_t i = 1; i < net->layers.size(); ++i) {
    Layer& L = net->layers[i];
    if (!L.info.has_bn) { L.tparams = L.params; continue; }   // detection convs run as in inference
    yb_conv_desc d; memset(&d, 0, sizeof(d));
    d.n = net->n; d.h = L.info.in_h; d.w = L.info.in_w; d.cin = L.info.cin; d.cout = L.info.cout;
    d.ksize = L.info.ksize; d.stride = L.info.stride;
    d.in_ld = net->bufs[L.in.buf].ld; d.out_ld = L.info.cout; d.res_ld = 0;
    d.dtype = net->dtype; d.out_fp32 = 0; d.leaky = 0; d.upsample2x = 0;
    CUtensorMap a, b; int cp = 0;
    int rc = conv_prepare(&d, ten_ptr2(net, L.in), net->par + L.w_packed, ones, zeros, nullptr, net->act + L.z_off,
                          fact(net, L.st_sum), fact(net, L.st_sqsum), &a, &b, &L.tparams, &cp);
    if (rc) return rc;
  }
  // ---- dgrad convs + residual bookkeeping (reverse order) ----
  std::map<int, std::vector<std::pair<int, int>>> written;   // gradient buffer -> channel intervals already produced
  struct Pending { const void* ptr; long ld; };
  std::map<std::pair<int, int>, Pending> pending;            // (buf, off) -> residual pass-through source
  auto covered = [&](const Ten& t) {
    auto it = written.find(t.buf);
    if (it == written.end()) return false;
    for (auto& iv : it->second) if (iv.first <= t.off && t.off + t.c <= iv.first + iv.second) return true;
    return false;
  };
  for (int i = (int)net->layers.size() - 1; i >= 1; --i) {
    Layer& L = net->layers[i];
    // this layer's residual input receives dA(out) unchanged
    if (L.res.buf >= 0) pending[{L.res.buf, L.res.off}] = Pending{gten_ptr(net, L.out), (long)net->bufs[L.out.buf].ld};
    // dgrad: dA(in) (+)= conv_s1(dz [zero-inserted if stride 2], Wd)
    yb_conv_desc d; memset(&d, 0, sizeof(d));
    d.n = net->n; d.h = L.info.in_h; d.w = L.info.in_w;      // dz (dilated for stride 2) lives at the INPUT resolution
    d.cin = L.k_cout; d.cout = L.info.cin; d.ksize = L.info.ksize; d.stride = 1;
    d.in_ld = L.dz_ld; d.out_ld = net->bufs[L.in.buf].ld;
    d.dtype = net->dtype; d.out_fp32 = 0; d.leaky = 0; d.upsample2x = 0;
    const void* res = nullptr;
    auto pit = pending.find({L.in.buf, L.in.off});
    const bool cov = covered(L.in);
    if (cov && pit != pending.end()) { set_error("train plan: tensor with both a written gradient and a pending residual"); return YB_ERR_UNSUPPORTED; }
    if (cov) { res = gten_ptr(net, L.in); d.res_ld = d.out_ld; }
    else if (pit != pending.end()) { res = pit->second.ptr; d.res_ld = (int)pit->second.ld; pending.erase(pit); }
    int rc = YB_OK;
    if (L.dgrad_parity) {
      d.h = L.info.out_h; d.w = L.info.out_w;                // plain dz at the OUTPUT resolution
      const size_t esz = 2, per = (size_t)yb_conv_cout_pad(L.info.cin) * L.k_cout;
      const size_t woff[4] = {0, per, 3 * per, 5 * per};
      for (int c = 0; c < 4 && rc == YB_OK; ++c)
        rc = conv_prepare_win(&d, 1 + (c >> 1), 1 + (c & 1), 1 + c, net->act + L.dz_off, net->par + L.w_dgrad + woff[c] * esz,
                              ones, zeros, res, gten_ptr(net, L.in), &L.d4_tmA[c], &L.d4_tmB[c], &L.d4_params[c],
                              &L.d4_cout_pad[c]);
    } else {
      rc = conv_prepare(&d, net->act + L.dz_off, net->par + L.w_dgrad, ones, zeros, res, gten_ptr(net, L.in), nullptr,
                        nullptr, &L.d_tmA, &L.d_tmB, &L.dparams, &L.d_cout_pad);
    }
    if (rc) return rc;
    written[L.in.buf].push_back({L.in.off, L.in.c});
  }
  if (!pending.empty()) { set_error("train plan: unconsumed residual gradient"); return YB_ERR_UNSUPPORTED; }
  // ---- optimizer tables ----
  {
    size_t ti = 0;
    float* gbase = reinterpret_cast<float*>(net->par + net->grad_off);
    float* vbase = reinterpret_cast<float*>(net->par + net->vel_off);
    float* v2base = vbase + net->grad_count;
    for (auto& L : net->layers) {
      OptTensor& tw = net->opt_tensors[ti++];
      tw.w = fpar(net, L.w_master); tw.g = gbase + L.g_w; tw.v = vbase + L.g_w; tw.v2 = v2base + L.g_w; tw.w16 = net->par + L.w_packed;
      if (L.info.index == 0) tw.w16 = nullptr;     // the stem reads its fp32 master weights
      if (L.info.has_bn) {
        OptTensor& tg = net->opt_tensors[ti++];
        tg.w = fpar(net, L.gamma); tg.g = gbase + L.g_gamma; tg.v = vbase + L.g_gamma; tg.v2 = v2base + L.g_gamma; tg.w16 = nullptr;
        OptTensor& tb = net->opt_tensors[ti++];
        tb.w = fpar(net, L.beta); tb.g = gbase + L.g_beta; tb.v = vbase + L.g_beta; tb.v2 = v2base + L.g_beta; tb.w16 = nullptr;
      } else {
        OptTensor& tb = net->opt_tensors[ti++];
        tb.w = fpar(net, L.bias); tb.g = gbase + L.g_bias; tb.v = vbase + L.g_bias; tb.v2 = v2base + L.g_bias; tb.w16 = nullptr;
      }
    }
    // (pageable host source: the copies are staged before the calls return; the vectors live as long as the plan)
    YB_CUDA(cudaMemcpyAsync(net->par + net->opt_tensors_off, net->opt_tensors.data(), net->opt_tensors.size() * sizeof(OptTensor),
                            cudaMemcpyHostToDevice, st));
    YB_CUDA(cudaMemcpyAsync(net->par + net->opt_chunks_off, net->opt_chunks.data(), net->opt_chunks.size() * sizeof(OptChunk),
                            cudaMemcpyHostToDevice, st));
    // dgrad-weight repack table: every layer but the stem, one launch (pack_dgrad_all)
    net->pack_jobs.clear();
    int tile0 = 0;
    for (auto& L : net->layers) {
      if (L.info.index == 0) continue;
      PackJob j; memset(&j, 0, sizeof(j));
      j.w = fpar(net, L.w_master); j.dst = net->par + L.w_dgrad;
      j.cout = L.info.cout; j.cin = L.info.cin; j.ks = L.info.ksize; j.kco = L.k_cout;
      j.cin_pad = yb_conv_cout_pad(L.info.cin); j.s2 = L.dgrad_parity ? 1 : 0;
      j.tiles_ci = (j.cin_pad + 31) / 32; j.tiles_co = (j.kco + 127) / 128;
      j.tile0 = tile0;
      tile0 += j.ks * j.ks * j.tiles_ci * j.tiles_co;
      net->pack_jobs.push_back(j);
    }
    net->pack_tiles = tile0;
    YB_CUDA(cudaMemcpyAsync(net->par + net->pack_jobs_off, net->pack_jobs.data(), net->pack_jobs.size() * sizeof(PackJob),
                            cudaMemcpyHostToDevice, st));
  }
  return YB_OK;
}

// dgrad weights follow the master weights (after set_conv_params and after every update)
int train_refresh_dgrad_weights(yb_net* net, int layer, void* stream) {
  Layer& L = net->layers[layer];
  if (layer == 0) return YB_OK;
  if (L.dgrad_parity)
    return yb_pack_dgrad_weights_s2(fpar(net, L.w_master), L.info.cout, L.info.cin, L.k_cout, yb_conv_cout_pad(L.info.cin),
                                    net->dtype, net->par + L.w_dgrad, stream);
  return yb_pack_dgrad_weights(fpar(net, L.w_master), L.info.cout, L.info.cin, L.info.ksize, L.k_cout,
                               yb_conv_cout_pad(L.info.cin), net->dtype, net->par + L.w_dgrad, stream);
}

// ... all layers: one multi-tensor launch (YB_PACK_MT=0: the per-layer kernels)
static int refresh_all_dgrad_weights(yb_net* net, void* stream) {
  if (opt("YB_PACK_MT")[0] != '0' && !net->pack_jobs.empty())
    return pack_dgrad_all(reinterpret_cast<const PackJob*>(net->par + net->pack_jobs_off), (int)net->pack_jobs.size(),
                          net->pack_tiles, net->dtype, static_cast<cudaStream_t>(stream));
  for (size_t i = 1; i < net->layers.size(); ++i) {
    int rc = train_refresh_dgrad_weights(net, (int)i, stream);
    if (rc) return rc;
  }
  return YB_OK;
}

}  // namespace yb

using namespace yb;

extern "C" int yb_net_train_fwd_bwd(yb_net* net, const float* images, const float* y_true_1, const float* y_true_2,
                                    const float* y_true_3, const float* anchors9x2, int use_label_smooth,
                                    int use_focal_loss, float bn_decay, float loss_scale, float* fm1, float* fm2,
                                    float* fm3, double* loss4, int flags, void* stream) {
  const int forward_only = flags & YB_TRAIN_FORWARD_ONLY;
  const bool bn_frozen = (flags & YB_TRAIN_BN_FROZEN) != 0;
  YB_REQUIRE(net && net->training && net->act && net->par, "train_fwd_bwd: not a bound training plan");
  YB_REQUIRE(images, "train_fwd_bwd: null images");
  YB_REQUIRE(forward_only || (y_true_1 && y_true_2 && y_true_3 && anchors9x2 && loss4), "train_fwd_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = net->n, dt = net->dtype;
  const float* ones = fpar(net, net->ones_off);
  const float* zeros = fpar(net, net->zeros_off);
  float* user_fm[3] = {fm1, fm2, fm3};
  const float* y_true[3] = {y_true_1, y_true_2, y_true_3};
  int rc;
  YB_CUDA(cudaMemsetAsync(net->act + net->stats_off, 0, net->stats_bytes, st));
  if (!forward_only) {
    YB_CUDA(cudaMemsetAsync(net->par + net->grad_off, 0, (size_t)net->grad_count * 4, st));
    YB_CUDA(cudaMemsetAsync(loss4, 0, 4 * sizeof(double), st));
  }
  // ------------------------------------------------ forward (is_training=True)
  float* fm_ptr[3] = {nullptr, nullptr, nullptr};
  const bool fuse_fin = opt("YB_BN_FIN")[0] != '0';
  const bool stem_tc = opt("YB_STEM_TRAIN")[0] != 'c';
  for (size_t i = 0; i < net->layers.size(); ++i) {
    Layer& L = net->layers[i];
    const long rows = (long)n * L.info.out_h * L.info.out_w;
    if (i == 0) {
      if (stem_tc) {   // warp-level tensor path, batch statistics accumulated by the same kernel (the sums were zeroed above)
        rc = yb_stem_conv_fwd_tc_stats(images, fpar(net, L.w_master), ones, zeros, n, net->h, net->w, dt, 0,
                                       net->act + L.z_off, fact(net, L.st_sum), fact(net, L.st_sqsum), stream);
        if (rc) return rc;
      } else {         // YB_STEM_TRAIN=cuda: the fp32 CUDA-core stem + a column-statistics pass (the first version)
        rc = yb_stem_conv_fwd(images, fpar(net, L.w_master), ones, zeros, n, net->h, net->w, L.info.cout, dt, 0,
                              net->act + L.z_off, stream);
        if (rc) return rc;
        rc = yb_col_stats(net->act + L.z_off, L.info.cout, rows, L.info.cout, dt, fact(net, L.st_sum), fact(net, L.st_sqsum), stream);
        if (rc) return rc;
      }
    } else {
      ConvParams p = L.tparams;
      if (!L.info.has_bn) {
        const int which = L.out.buf == net->fm_buf[0] ? 0 : (L.out.buf == net->fm_buf[1] ? 1 : 2);
        p.out = user_fm[which] ? (void*)user_fm[which] : (void*)(net->act + net->bufs[L.out.buf].offset);
        fm_ptr[which] = static_cast<float*>(p.out);
      }
      rc = conv_launch(dt, L.cout_pad, L.tmA, L.tmB, p, st);
      if (rc) return rc;
    }
    if (L.info.has_bn) {
      const void* resp = L.res.buf >= 0 ? ten_ptr2(net, L.res) : nullptr;
      const long res_ld = L.res.buf >= 0 ? net->bufs[L.res.buf].ld : 0;
      if (fuse_fin) {     // statistics -> scale/shift inside the apply kernel (one launch per BN layer instead of two)
        rc = yb_bn_stats_act_apply(net->act + L.z_off, L.info.cout, bn_frozen ? nullptr : fact(net, L.st_sum),
                                   bn_frozen ? nullptr : fact(net, L.st_sqsum), fpar(net, L.gamma), fpar(net, L.beta),
                                   net->bn_eps, bn_decay, fpar(net, L.mean), fpar(net, L.var), fact(net, L.st_scale),
                                   fact(net, L.st_shift), fact(net, L.st_mean), fact(net, L.st_invstd), resp, res_ld,
                                   ten_ptr2(net, L.out), net->bufs[L.out.buf].ld, n, L.info.out_h, L.info.out_w,
                                   L.info.cout, dt, 1, L.upsample ? 1 : 0, stream);
        if (rc) return rc;
        continue;
      }
      rc = yb_bn_finalize(bn_frozen ? nullptr : fact(net, L.st_sum), bn_frozen ? nullptr : fact(net, L.st_sqsum), rows,
                          L.info.cout, fpar(net, L.gamma), fpar(net, L.beta),
                          net->bn_eps, bn_decay, fpar(net, L.mean), fpar(net, L.var), fact(net, L.st_scale),
                          fact(net, L.st_shift), fact(net, L.st_mean), fact(net, L.st_invstd), stream);
      if (rc) return rc;
      rc = yb_bn_act_apply(net->act + L.z_off, L.info.cout, fact(net, L.st_scale), fact(net, L.st_shift), resp, res_ld,
                           ten_ptr2(net, L.out), net->bufs[L.out.buf].ld, n, L.info.out_h, L.info.out_w, L.info.cout, dt, 1,
                           L.upsample ? 1 : 0, stream);
      if (rc) return rc;
    }
  }
  if (!bn_frozen) net->fold_dirty = true;
  if (forward_only) return YB_OK;
  // ------------------------------------------------ loss + d(loss)/d(feature maps)
  for (int s = 0; s < 3; ++s) {
    const int div = 32 >> s;
    rc = yb_loss_layer(fm_ptr[s], y_true[s], n, net->h / div, net->w / div, net->h, net->w, net->class_num,
                       anchors9x2 + 2 * 3 * (2 - s), use_label_smooth, use_focal_loss, 1.0f / (float)n, loss_scale,
                       net->act + net->lossws_off, net->lossws_bytes, loss4, net->act + net->dfm_off[s], dt,
                       (3 * (5 + net->class_num) + 31) / 32 * 32, stream);
    if (rc) return rc;
  }
  if (flags & YB_TRAIN_NO_BACKWARD) return YB_OK;
  return yb_net_train_backward(net, images, 0, (int)net->layers.size() - 1, flags, stream);
}

// Backward of layers last_layer .. first_layer (descending), after yb_net_train_fwd_bwd(..., YB_TRAIN_NO_BACKWARD):
// lets a data-parallel caller all-reduce the finished head-side gradient buckets while the backbone is still running.
extern "C" int yb_net_train_backward(yb_net* net, const float* images, int first_layer, int last_layer, int flags,
                                     void* stream) {
  YB_REQUIRE(net && net->training && net->act && net->par && images, "train_backward: not a bound training plan");
  YB_REQUIRE(first_layer >= 0 && first_layer <= last_layer && last_layer < (int)net->layers.size(), "train_backward: bad layer range");
  const bool bn_frozen = (flags & YB_TRAIN_BN_FROZEN) != 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int n = net->n, dt = net->dtype;
  const float* zeros = fpar(net, net->zeros_off);
  int rc;
  // Layer L's weight gradient and input gradient both only read dz_L: the wgrad goes to a side stream and runs beside
  // the dgrad (and the next layer's BN backward).  Both kernels own a whole SM per CTA, so the gain is in the tails:
  // the SMs a finishing kernel frees — and the ~15 us every wgrad CTA spends flushing fp32 atomics at its end — are
  // picked up by the other kernel's CTAs instead of idling until the launch boundary.  dz_L and the forward activations
  // are never rewritten during the backward (one buffer per layer), the join below orders the side stream before
  // whatever the caller enqueues next (all-reduce of this layer range, optimizer, next step's gradient memset).
  const bool overlap = opt("YB_WGRAD_STREAM")[0] != '0';
  bool forked = false;
  if (overlap && net->side_stream == nullptr) {
    YB_CUDA(cudaStreamCreateWithFlags(&net->side_stream, cudaStreamNonBlocking));
    YB_CUDA(cudaEventCreateWithFlags(&net->side_fork, cudaEventDisableTiming));
    YB_CUDA(cudaEventCreateWithFlags(&net->side_join, cudaEventDisableTiming));
  }
  for (int i = last_layer; i >= first_layer; --i) {
    Layer& L = net->layers[i];
    const long rows = (long)n * L.info.out_h * L.info.out_w;
    if (L.info.has_bn) {
      float* dgamma = gradp(net, L.g_gamma);
      float* dbeta = gradp(net, L.g_beta);
      const void* dA = gten_ptr(net, L.out);
      const long dA_ld = net->bufs[L.out.buf].ld;
      rc = yb_bn_bwd_reduce(dA, dA_ld, net->act + L.z_off, L.info.cout, fact(net, L.st_scale), fact(net, L.st_shift),
                            fact(net, L.st_mean), fact(net, L.st_invstd), n, L.info.out_h, L.info.out_w, L.info.cout, dt, 1,
                            L.upsample ? 1 : 0, dgamma, dbeta, net->act + net->bnws_off, stream);
      if (rc) return rc;
      // frozen BN: mean / variance are constants -> dz = gamma * invstd * dact (the batch-statistic terms vanish)
      rc = yb_bn_bwd_apply(dA, dA_ld, net->act + L.z_off, L.info.cout, fpar(net, L.gamma), fact(net, L.st_scale),
                           fact(net, L.st_shift), fact(net, L.st_mean), fact(net, L.st_invstd), bn_frozen ? zeros : dgamma,
                           bn_frozen ? zeros : dbeta, n,
                           L.info.out_h, L.info.out_w, L.info.cout, dt, 1, L.upsample ? 1 : 0, L.dz_dilated,
                           net->act + L.dz_off, L.dz_ld, stream);
      if (rc) return rc;
    } else {
      rc = yb_col_sum(net->act + L.dz_off, L.dz_ld, rows, L.info.cout, dt, gradp(net, L.g_bias), stream);
      if (rc) return rc;
    }
    if (i == 0) {
      rc = yb_stem_conv_wgrad(images, net->act + L.dz_off, dt, n, net->h, net->w, gradp(net, L.g_w), stream);
      if (rc) return rc;
      break;
    }
    yb_conv_desc d; memset(&d, 0, sizeof(d));
    d.n = n; d.h = L.info.in_h; d.w = L.info.in_w; d.cin = L.info.cin; d.cout = L.info.cout;
    d.ksize = L.info.ksize; d.stride = L.info.stride; d.in_ld = net->bufs[L.in.buf].ld; d.dtype = dt;
    void* wstream = stream;
    if (overlap) {
      YB_CUDA(cudaEventRecord(net->side_fork, st));
      YB_CUDA(cudaStreamWaitEvent(net->side_stream, net->side_fork, 0));
      wstream = net->side_stream;
      forked = true;
    }
    rc = yb_conv2d_wgrad(&d, ten_ptr2(net, L.in), net->act + L.dz_off, L.dz_ld, L.dz_dilated, gradp(net, L.g_w), wstream);
    if (rc) return rc;
    if (L.dgrad_parity) {
      for (int c = 0; c < 4; ++c) {
        rc = conv_launch(dt, L.d4_cout_pad[c], L.d4_tmA[c], L.d4_tmB[c], L.d4_params[c], st);
        if (rc) return rc;
      }
    } else {
      rc = conv_launch(dt, L.d_cout_pad, L.d_tmA, L.d_tmB, L.dparams, st);
      if (rc) return rc;
    }
  }
  if (forked) {
    YB_CUDA(cudaEventRecord(net->side_join, net->side_stream));
    YB_CUDA(cudaStreamWaitEvent(st, net->side_join, 0));
  }
  return YB_OK;
}

extern "C" int yb_net_train_reset_state(yb_net* net, int optimizer_kind, void* stream) {
  YB_REQUIRE(net && net->training && net->par, "train_reset_state: not a bound training plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  YB_CUDA(cudaMemsetAsync(net->par + net->vel_off, 0, (size_t)net->grad_count * 4 * net->opt_state_slots, st));
  if (optimizer_kind == YB_OPT_RMSPROP) {       // [TF] RMSPropOptimizer creates its `rms` slot with ones
    fill_f32_kernel<<<num_sms() * 4, 256, 0, st>>>(reinterpret_cast<float*>(net->par + net->vel_off) + net->grad_count,
                                                   net->grad_count, 1.0f);
    YB_CUDA(cudaGetLastError());
  }
  YB_CUDA(cudaMemsetAsync(net->par + net->grad_off, 0, (size_t)net->grad_count * 4, st));
  YB_CUDA(cudaMemsetAsync(net->par + net->opt_step_off, 0, 256, st));
  return YB_OK;
}

extern "C" int yb_net_opt_state(yb_net* net, float** slots, size_t* count_per_slot, int* num_slots, int** ctrl) {
  YB_REQUIRE(net && net->training && net->par, "opt_state: not a bound training plan");
  if (slots) *slots = reinterpret_cast<float*>(net->par + net->vel_off);
  if (count_per_slot) *count_per_slot = (size_t)net->grad_count;
  if (num_slots) *num_slots = net->opt_state_slots;
  if (ctrl) *ctrl = reinterpret_cast<int*>(net->par + net->opt_step_off);
  return YB_OK;
}

// train.py:81 `update_part`: restrict the update to some convs (their weights, gamma/beta or bias)
extern "C" int yb_net_set_trainable(yb_net* net, int layer, int trainable, void* stream) {
  YB_REQUIRE(net && net->training && net->par && layer >= 0 && layer < (int)net->layers.size(), "set_trainable: bad argument");
  size_t ti = 0;
  for (int i = 0; i < layer; ++i) ti += net->layers[i].info.has_bn ? 3 : 2;
  const int cnt = net->layers[layer].info.has_bn ? 3 : 2;
  for (int j = 0; j < cnt; ++j) net->opt_tensors[ti + j].trainable = trainable ? 1 : 0;
  YB_CUDA(cudaMemcpyAsync(net->par + net->opt_tensors_off + ti * sizeof(OptTensor), &net->opt_tensors[ti], cnt * sizeof(OptTensor),
                          cudaMemcpyHostToDevice, static_cast<cudaStream_t>(stream)));
  return YB_OK;
}

extern "C" int yb_net_train_refresh_dgrad(yb_net* net, void* stream) {
  YB_REQUIRE(net && net->training && net->par, "train_refresh_dgrad: not a bound training plan");
  return refresh_all_dgrad_weights(net, stream);
}

extern "C" int yb_net_grad_buffer(yb_net* net, float** ptr, size_t* count) {
  YB_REQUIRE(net && net->training && net->par && ptr && count, "grad_buffer: not a bound training plan");
  *ptr = reinterpret_cast<float*>(net->par + net->grad_off);
  *count = (size_t)net->grad_count;
  return YB_OK;
}

// flat-gradient slice of layers [first_layer, last_layer] (contiguous: the buffer is laid out in creation order)
extern "C" int yb_net_grad_range(yb_net* net, int first_layer, int last_layer, float** ptr, size_t* count) {
  YB_REQUIRE(net && net->training && net->par && ptr && count, "grad_range: not a bound training plan");
  YB_REQUIRE(first_layer >= 0 && first_layer <= last_layer && last_layer < (int)net->layers.size(), "grad_range: bad layer range");
  const long lo = net->layers[first_layer].g_w;
  const long hi = last_layer + 1 < (int)net->layers.size() ? net->layers[last_layer + 1].g_w : net->grad_count;
  *ptr = reinterpret_cast<float*>(net->par + net->grad_off) + lo;
  *count = (size_t)(hi - lo);
  return YB_OK;
}

extern "C" int yb_net_train_update(yb_net* net, const yb_optimizer* opt, void* stream) {
  YB_REQUIRE(net && net->training && net->par && opt, "train_update: not a bound training plan");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = opt_step(reinterpret_cast<const OptTensor*>(net->par + net->opt_tensors_off),
                    reinterpret_cast<const OptChunk*>(net->par + net->opt_chunks_off), net->num_opt_tensors,
                    net->num_opt_chunks, fpar(net, net->opt_norm_off), reinterpret_cast<int*>(net->par + net->opt_step_off),
                    net->dtype, *opt, st);
  if (rc) return rc;
  rc = refresh_all_dgrad_weights(net, stream);
  if (rc) return rc;
  net->fold_dirty = true;
  return YB_OK;
}

extern "C" int yb_net_get_conv_params(yb_net* net, int layer, float** w_ohwi, float** gamma, float** beta,
                                      float** mean, float** var, float** bias) {
  YB_REQUIRE(net && net->par && layer >= 0 && layer < (int)net->layers.size(), "get_conv_params: bad argument");
  Layer& L = net->layers[layer];
  if (w_ohwi) *w_ohwi = fpar(net, L.w_master);
  if (gamma) *gamma = L.info.has_bn ? fpar(net, L.gamma) : nullptr;
  if (beta) *beta = L.info.has_bn ? fpar(net, L.beta) : nullptr;
  if (mean) *mean = L.info.has_bn ? fpar(net, L.mean) : nullptr;
  if (var) *var = L.info.has_bn ? fpar(net, L.var) : nullptr;
  if (bias) *bias = L.info.has_bn ? nullptr : fpar(net, L.bias);
  return YB_OK;
}

extern "C" int yb_net_layer_grad(yb_net* net, int layer, float** dw, float** dgamma, float** dbeta, float** dbias) {
  YB_REQUIRE(net && net->training && net->par && layer >= 0 && layer < (int)net->layers.size(), "layer_grad: bad argument");
  Layer& L = net->layers[layer];
  if (dw) *dw = gradp(net, L.g_w);
  if (dgamma) *dgamma = L.info.has_bn ? gradp(net, L.g_gamma) : nullptr;
  if (dbeta) *dbeta = L.info.has_bn ? gradp(net, L.g_beta) : nullptr;
  if (dbias) *dbias = L.info.has_bn ? nullptr : gradp(net, L.g_bias);
  return YB_OK;
}

extern "C" int yb_net_train_buffer(yb_net* net, int layer, int which, void** ptr, int* ld, int* rows_h, int* rows_w) {
  YB_REQUIRE(net && net->training && net->act && layer >= 0 && layer < (int)net->layers.size() && ptr && ld,
             "train_buffer: bad argument");
  Layer& L = net->layers[layer];
  int h = L.info.out_h, w = L.info.out_w;
  switch (which) {
    case 0: YB_REQUIRE(L.info.has_bn, "train_buffer: no z for detection convs"); *ptr = net->act + L.z_off; *ld = L.info.cout; break;
    case 1: *ptr = net->act + L.dz_off; *ld = L.dz_ld; if (L.dz_dilated) { h = L.info.in_h; w = L.info.in_w; } break;
    case 2: YB_REQUIRE(L.info.has_bn, "train_buffer: no dA for detection convs"); *ptr = gten_ptr(net, L.out); *ld = net->bufs[L.out.buf].ld;
            if (L.upsample) { h *= 2; w *= 2; } break;
    case 3: YB_REQUIRE(layer > 0, "train_buffer: layer 0 reads the image"); *ptr = ten_ptr2(net, L.in); *ld = net->bufs[L.in.buf].ld;
            h = L.info.in_h; w = L.info.in_w; break;
    case 4: YB_REQUIRE(layer > 0 && net->par, "train_buffer: the stem has no dgrad weights");   // [cin_pad * k * k][k_cout], 16-bit
            *ptr = net->par + L.w_dgrad; *ld = L.k_cout; h = yb_conv_cout_pad(L.info.cin); w = L.info.ksize * L.info.ksize; break;
    default: set_error("train_buffer: which must be 0..4"); return YB_ERR_INVALID_ARGUMENT;
  }
  if (rows_h) *rows_h = h;
  if (rows_w) *rows_w = w;
  return YB_OK;
}
