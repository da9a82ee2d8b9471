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
// ctrl[0]: a non-finite gradient was seen this step (the whole update is skipped), ctrl[1]: updates applied so far
// (Adam's t - 1), ctrl[2]: steps skipped
int opt_step(const OptTensor* tensors, const OptChunk* chunks, int num_tensors, int num_chunks, float* sqnorm, int* ctrl,
             int dtype, const yb_optimizer& o, cudaStream_t st);

}  // namespace yb
