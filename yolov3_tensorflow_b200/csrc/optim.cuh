// Device-side tables of the multi-tensor optimizer (optim.cu), built by the training plan (net_train.cu).
#pragma once
#include "common.cuh"

namespace yb {

struct OptTensor {
  float* w;      // fp32 master values
  float* g;      // fp32 gradient (data term; L2 is added on the fly)
  float* v;      // optimizer slot 1: momentum accumulator (momentum, rmsprop) / first moment (adam)
  float* v2;     // optimizer slot 2: mean square (rmsprop, initialised to 1 like TF) / second moment (adam)
  void* w16;     // 16-bit compute copy laid out like w (conv weights), or null
  long n;
  int l2;        // 1: conv weight (slim.l2_regularizer applies), 0: gamma/beta/bias
  int trainable; // 0: not in update_vars (train.py:81 `update_part`): never touched
};
struct OptChunk {
  int tensor;
  long begin, end;
};
// One conv layer of the multi-tensor dgrad-weight repack (pack_dgrad_all): the flip + transpose W[co][r][s][ci] ->
// Wd[ci][r'][s'][co] of every layer in ONE launch (a block = one 32(ci) x 128(co) tile of one tap, transposed through
// shared memory so that both the fp32 reads and the 16-bit writes are coalesced).
struct PackJob {
  const float* w;   // fp32 master weights [cout][ks][ks][cin]
  void* dst;        // 16-bit dgrad weights (layouts of yb_pack_dgrad_weights / yb_pack_dgrad_weights_s2)
  int cout, cin, ks, kco, cin_pad, s2;
  int tiles_ci, tiles_co;
  int tile0;        // first block of this layer
};
int pack_dgrad_all(const PackJob* jobs, int num_jobs, int total_tiles, int dtype, cudaStream_t st);

// ctrl[0]: a non-finite gradient was seen this step (the whole update is skipped), ctrl[1]: updates applied so far
// (Adam's t - 1), ctrl[2]: steps skipped
int opt_step(const OptTensor* tensors, const OptChunk* chunks, int num_tensors, int num_chunks, float* sqnorm, int* ctrl,
             int dtype, const yb_optimizer& o, cudaStream_t st);

}  // namespace yb
