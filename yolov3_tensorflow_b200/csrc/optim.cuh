// Device-side tables of the multi-tensor optimizer (optim.cu), built by the training plan (net.cu).
#pragma once
#include "common.cuh"

namespace yb {

struct OptTensor {
  float* w;      // fp32 master values
  float* g;      // fp32 gradient (data term; L2 is added on the fly)
  float* v;      // momentum accumulator
  void* w16;     // 16-bit compute copy laid out like w (conv weights), or null
  long n;
  int l2;        // 1: conv weight (slim.l2_regularizer applies), 0: gamma/beta/bias
};
struct OptChunk {
  int tensor;
  long begin, end;
};

int opt_step(const OptTensor* tensors, const OptChunk* chunks, int num_tensors, int num_chunks, float* sqnorm,
             int dtype, float lr, float grad_scale, float momentum, float weight_decay, float clip, cudaStream_t st);

}  // namespace yb
