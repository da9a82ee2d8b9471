// Batch-norm (training mode) + leaky-ReLU forward/backward around the tensor-core convs.
// TF semantics (slim.batch_norm fused, model.py:35-41): batch mean and
// *biased* variance normalise; moving stats are updated with the *unbiased* variance
// (train.py:108-109 UPDATE_OPS).  The conv epilogue already produced the per-channel
// sum / sum of squares of the raw conv output z (yb_conv2d_fwd stat_sum/stat_sqsum).
//
//   bn_finalize      : sums -> mean/var -> scale/shift (+ saved mean/invstd, moving update)
//   bn_act_apply     : a = leaky(z*scale+shift) (+ residual), optional 2x-upsample store
//   bn_bwd_reduce    : dbeta = sum(dact), dgamma = sum(dact * zhat), dact = dA * leaky'(y)
//   bn_bwd_apply     : dz = gamma*invstd*(dact - dbeta/M - zhat*dgamma/M), optional
//                      zero-insertion (dilated) store for the dgrad of stride-2 convs
//   col_sum          : bias gradient of the detection convs
// All are HBM-bound streaming kernels over [rows, C] NHWC 16-bit tensors (C % 8 == 0):
// a thread owns 8 consecutive channels (one 16-byte load), a block a slab of rows.
#include "common.cuh"

namespace yb {

__global__ void bn_finalize_kernel(const float* __restrict__ sum, const float* __restrict__ sqsum, float count, int c,
                                   const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                   float decay, float* moving_mean, float* moving_var, float* scale, float* shift,
                                   float* save_mean, float* save_invstd) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  if (sum == nullptr) {
    // frozen BN (forward(is_training=False) under a gradient tape): normalise with the MOVING statistics, which are
    // constants of the backward pass and are not updated
    const float invstd = rsqrtf(moving_var[i] + eps);
    const float sc = gamma[i] * invstd;
    scale[i] = sc;
    shift[i] = beta[i] - moving_mean[i] * sc;
    save_mean[i] = moving_mean[i];
    save_invstd[i] = invstd;
    return;
  }
  const float mean = sum[i] / count;
  float var = sqsum[i] / count - mean * mean;   // biased
  var = fmaxf(var, 0.f);
  const float invstd = rsqrtf(var + eps);
  const float sc = gamma[i] * invstd;
  scale[i] = sc;
  shift[i] = beta[i] - mean * sc;
  save_mean[i] = mean;
  save_invstd[i] = invstd;
  if (moving_mean) {
    const float unb = count > 1.f ? var * count / (count - 1.f) : var;
    moving_mean[i] = moving_mean[i] * decay + (1.f - decay) * mean;
    moving_var[i] = moving_var[i] * decay + (1.f - decay) * unb;
  }
}

struct RowGeom {
  long rows;        // n*h*w
  int h, w;         // spatial (for the upsample / dilate address maps)
  int c;
};

template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  float2 f;
  f = Pack2<T>::unpack(u.x); v[0] = f.x; v[1] = f.y;
  f = Pack2<T>::unpack(u.y); v[2] = f.x; v[3] = f.y;
  f = Pack2<T>::unpack(u.z); v[4] = f.x; v[5] = f.y;
  f = Pack2<T>::unpack(u.w); v[6] = f.x; v[7] = f.y;
}
template <typename T>
__device__ __forceinline__ void store8(T* p, const float (&v)[8]) {
  uint4 u;
  u.x = Pack2<T>::pack(v[0], v[1]); u.y = Pack2<T>::pack(v[2], v[3]);
  u.z = Pack2<T>::pack(v[4], v[5]); u.w = Pack2<T>::pack(v[6], v[7]);
  *reinterpret_cast<uint4*>(p) = u;
}
// row index in the 2x-upsampled [n, 2h, 2w] grid of the top-left copy of row r of [n, h, w]
__device__ __forceinline__ long up_row(long r, int h, int w) {
  const long q = r % w, pp = (r / w) % h, img = r / ((long)w * h);
  return ((img * 2 * h + 2 * pp) * (2L * w)) + 2 * q;
}

// ---- streaming kernels: thread = (8-channel vector cv, row lane); per-channel coefficients live in registers,
// ---- rows are walked 4 at a time so several 16-byte loads are in flight per thread.
template <typename T>
__device__ __forceinline__ void load_dA(const T* dA, long dA_ld, long r, int c0, const RowGeom& g, int upsample,
                                        float (&v)[8]) {
  if (!upsample) {
    load8(dA + r * dA_ld + c0, v);
  } else {   // the forward stored this row 2x-upsampled: its gradient is the sum of the 4 copies
    const long b = up_row(r, g.h, g.w);
    const long W2 = 2L * g.w;
    float t0[8], t1[8], t2[8];
    load8(dA + b * dA_ld + c0, v);
    load8(dA + (b + 1) * dA_ld + c0, t0);
    load8(dA + (b + W2) * dA_ld + c0, t1);
    load8(dA + (b + W2 + 1) * dA_ld + c0, t2);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += t0[j] + t1[j] + t2[j];
  }
}

struct StreamGeom {
  int cv;             // channel vectors per row (c / 8)
  int lanes;          // row lanes per block (256 / cv)
  long rows_per_block;
};

template <typename T>
__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  float2 f;
  f = Pack2<T>::unpack(u.x); v[0] = f.x; v[1] = f.y;
  f = Pack2<T>::unpack(u.y); v[2] = f.x; v[3] = f.y;
  f = Pack2<T>::unpack(u.z); v[4] = f.x; v[5] = f.y;
  f = Pack2<T>::unpack(u.w); v[6] = f.x; v[7] = f.y;
}
template <typename T>
__global__ void __launch_bounds__(256, 3)
bn_act_apply_kernel(const T* __restrict__ z, long z_ld, const float* __restrict__ scale, const float* __restrict__ shift,
                    const T* __restrict__ res, long res_ld, T* __restrict__ out, long out_ld, RowGeom g, StreamGeom sg,
                    int leaky, int upsample) {
  const int cvi = threadIdx.x % sg.cv, lane_r = threadIdx.x / sg.cv;
  const int c0 = cvi * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; }
  const long r0 = blockIdx.x * sg.rows_per_block, r1 = min(r0 + sg.rows_per_block, g.rows);
  for (long rb = r0 + lane_r; rb < r1; rb += 4L * sg.lanes) {
    uint4 zq[4], rq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long r = rb + (long)u * sg.lanes;
      if (r < r1) {
        zq[u] = *reinterpret_cast<const uint4*>(z + r * z_ld + c0);
        if (res) rq[u] = *reinterpret_cast<const uint4*>(res + r * res_ld + c0);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long r = rb + (long)u * sg.lanes;
      if (r >= r1) break;
      float v[8], rv[8];
      unpack8<T>(zq[u], v);
      if (res) unpack8<T>(rq[u], rv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float y = fmaf(v[j], sc[j], sh[j]);
        if (leaky) y = leaky01(y);
        if (res) y += rv[j];
        v[j] = y;
      }
      if (!upsample) {
        store8(out + r * out_ld + c0, v);
      } else {
        const long bb = up_row(r, g.h, g.w);
        const long W2 = 2L * g.w;
        store8(out + bb * out_ld + c0, v); store8(out + (bb + 1) * out_ld + c0, v);
        store8(out + (bb + W2) * out_ld + c0, v); store8(out + (bb + W2 + 1) * out_ld + c0, v);
      }
    }
  }
}

// Stage 1: every block reduces its row slab to per-channel partial sums.  With a (zero-initialised) workspace the
// partials are added into one of BN_SLOTS rows [slot][2][c] and the LAST block to finish (ticket counter) sums the
// rows, writes dgamma/dbeta and zeroes the rows again; without one they are added atomically to dgamma/dbeta
// (~600-deep same-address atomics: slow, kept for callers without a workspace).
static constexpr int BN_SLOTS = 16;

template <typename T>
__global__ void __launch_bounds__(256, 3)
bn_bwd_reduce_kernel(const T* __restrict__ dA, long dA_ld, const T* __restrict__ z, long z_ld,
                     const float* __restrict__ scale, const float* __restrict__ shift,
                     const float* __restrict__ save_mean, const float* __restrict__ save_invstd, RowGeom g,
                     StreamGeom sg, int leaky, int upsample, float* __restrict__ dgamma, float* __restrict__ dbeta,
                     float* __restrict__ partial, unsigned int* __restrict__ ticket) {
  __shared__ float s_g[256][9], s_b[256][9];
  __shared__ unsigned int s_last;
  const int cvi = threadIdx.x % sg.cv, lane_r = threadIdx.x / sg.cv;
  const int c0 = cvi * 8;
  // per-channel coefficients live in registers (the first version re-loaded four of them per ELEMENT through the LSU,
  // which made this kernel 2x slower than bn_bwd_apply on the same data: profiles/r01_k).  The invstd factor of
  // zhat = (z - mean) * invstd is folded in once per block at the end.
  float sc[8], sh[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { sc[j] = scale[c0 + j]; sh[j] = shift[c0 + j]; }
  float ag[8], ab[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) ag[j] = ab[j] = 0.f;
  const long r0 = blockIdx.x * sg.rows_per_block, r1 = min(r0 + sg.rows_per_block, g.rows);
  for (long rb = r0 + lane_r; rb < r1; rb += 4L * sg.lanes) {
    uint4 zq[4], dq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long r = rb + (long)u * sg.lanes;
      if (r < r1) { zq[u] = *reinterpret_cast<const uint4*>(z + r * z_ld + c0); if (!upsample) dq[u] = *reinterpret_cast<const uint4*>(dA + r * dA_ld + c0); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if (rb + (long)u * sg.lanes >= r1) break;
      float zv[8], dv[8];
      unpack8<T>(zq[u], zv);
      if (!upsample) unpack8<T>(dq[u], dv);
      else load_dA(dA, dA_ld, rb + (long)u * sg.lanes, c0, g, 1, dv);   // fp32 sum of the 4 upsampled copies
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = fmaf(zv[j], sc[j], sh[j]);
        const float da = (leaky && y <= 0.f) ? 0.1f * dv[j] : dv[j];
        ab[j] += da;
        ag[j] = fmaf(da, zv[j], ag[j]);        // sum(da * z); centred and scaled below
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)   // sum(da * zhat) = invstd * (sum(da * z) - mean * sum(da))   (linear: exact per block)
    ag[j] = (ag[j] - __ldg(save_mean + c0 + j) * ab[j]) * __ldg(save_invstd + c0 + j);
#pragma unroll
  for (int j = 0; j < 8; ++j) { s_g[threadIdx.x][j] = ag[j]; s_b[threadIdx.x][j] = ab[j]; }
  __syncthreads();
  if (lane_r == 0) {
    for (int y = 1; y < sg.lanes; ++y) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { ag[j] += s_g[y * sg.cv + cvi][j]; ab[j] += s_b[y * sg.cv + cvi][j]; }
    }
    if (partial == nullptr) {
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(dgamma + c0 + j, ag[j]); atomicAdd(dbeta + c0 + j, ab[j]); }
    } else {
      // BN_SLOTS partial rows (zero on entry): ~grid/BN_SLOTS-deep atomics per address instead of grid-deep, and the
      // final pass reads BN_SLOTS x 2c floats instead of grid x 2c (which made one SM stream megabytes: r01_i)
      float* pg = partial + (long)(blockIdx.x % BN_SLOTS) * 2 * g.c;
#pragma unroll
      for (int j = 0; j < 8; ++j) { atomicAdd(pg + c0 + j, ag[j]); atomicAdd(pg + g.c + c0 + j, ab[j]); }
    }
  }
  if (partial == nullptr) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = (atomicAdd(ticket, 1u) == gridDim.x - 1) ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = threadIdx.x; c < 2 * g.c; c += 256) {
    float v[BN_SLOTS];
#pragma unroll
    for (int sl = 0; sl < BN_SLOTS; ++sl) v[sl] = __ldcg(partial + (long)sl * 2 * g.c + c);
    float acc = 0.f;
#pragma unroll
    for (int sl = 0; sl < BN_SLOTS; ++sl) { acc += v[sl]; partial[(long)sl * 2 * g.c + c] = 0.f; }   // re-armed
    if (c < g.c) dgamma[c] = acc; else dbeta[c - g.c] = acc;
  }
  if (threadIdx.x == 0) *ticket = 0u;   // ready for the next launch
}

template <typename T>
__global__ void __launch_bounds__(256, 3)
bn_bwd_apply_kernel(const T* __restrict__ dA, long dA_ld, const T* __restrict__ z, long z_ld,
                    const float* __restrict__ gamma, const float* __restrict__ scale, const float* __restrict__ shift,
                    const float* __restrict__ save_mean, const float* __restrict__ save_invstd,
                    const float* __restrict__ dgamma, const float* __restrict__ dbeta, RowGeom g, StreamGeom sg,
                    int leaky, int upsample, int dilate, T* __restrict__ dz, long dz_ld) {
  const int cvi = threadIdx.x % sg.cv, lane_r = threadIdx.x / sg.cv;
  const int c0 = cvi * 8;
  const float inv_m = 1.f / (float)g.rows;
  // dz = k1*dact + k2*z + k3   with   k1 = gamma*invstd, k2 = -k1*invstd*dgamma/M, k3 = -k1*dbeta/M - k2*mean
  float sc[8], sh[8], k1[8], k2[8], k3[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    sc[j] = scale[c]; sh[j] = shift[c];
    const float is = save_invstd[c];
    k1[j] = gamma[c] * is;
    k2[j] = -k1[j] * is * dgamma[c] * inv_m;
    k3[j] = -k1[j] * dbeta[c] * inv_m - k2[j] * save_mean[c];
  }
  const long r0 = blockIdx.x * sg.rows_per_block, r1 = min(r0 + sg.rows_per_block, g.rows);
  for (long rb = r0 + lane_r; rb < r1; rb += 4L * sg.lanes) {
    uint4 zq[4], dq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long r = rb + (long)u * sg.lanes;
      if (r < r1) { zq[u] = *reinterpret_cast<const uint4*>(z + r * z_ld + c0); if (!upsample) dq[u] = *reinterpret_cast<const uint4*>(dA + r * dA_ld + c0); }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long r = rb + (long)u * sg.lanes;
      if (r >= r1) break;
      float zv[8], dv[8], o[8];
      unpack8<T>(zq[u], zv);
      if (!upsample) unpack8<T>(dq[u], dv);
      else load_dA(dA, dA_ld, r, c0, g, 1, dv);   // fp32 sum of the 4 upsampled copies
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float y = fmaf(zv[j], sc[j], sh[j]);
        const float da = (leaky && y <= 0.f) ? 0.1f * dv[j] : dv[j];
        o[j] = fmaf(k1[j], da, fmaf(k2[j], zv[j], k3[j]));
      }
      const long orow = dilate ? up_row(r, g.h, g.w) : r;   // (2p, 2q) of a zero-initialised [n,2h,2w] buffer
      store8(dz + orow * dz_ld + c0, o);
    }
  }
}

static StreamGeom stream_geom(long rows, int c, int* grid) {
  StreamGeom sg;
  sg.cv = c / 8;
  sg.lanes = 256 / sg.cv;
  if (sg.lanes < 1) sg.lanes = 1;
  long blocks = (long)num_sms() * 4;
  long rpb = (rows + blocks - 1) / blocks;
  const long unit = 4L * sg.lanes;
  rpb = (rpb + unit - 1) / unit * unit;
  sg.rows_per_block = rpb;
  *grid = (int)((rows + rpb - 1) / rpb);
  return sg;
}

template <typename T>
__global__ void __launch_bounds__(256)
col_sum_kernel(const T* __restrict__ x, long ld, long rows, int c, long rows_per_block, float* __restrict__ out,
               float* __restrict__ out_sq) {
  // block (32, 8): 32 channels x 8 row lanes
  __shared__ float s[8][33], s2[8][33];
  const int ch = blockIdx.y * 32 + threadIdx.x;
  float a = 0.f, a2 = 0.f;
  const long r0 = blockIdx.x * rows_per_block, r1 = min(r0 + rows_per_block, rows);
  if (ch < c)
    for (long r = r0 + threadIdx.y; r < r1; r += 8) {
      const float v = static_cast<float>(x[r * ld + ch]);
      a += v; a2 += v * v;
    }
  s[threadIdx.y][threadIdx.x] = a; s2[threadIdx.y][threadIdx.x] = a2;
  __syncthreads();
  if (threadIdx.y == 0 && ch < c) {
    for (int y = 1; y < 8; ++y) { a += s[y][threadIdx.x]; a2 += s2[y][threadIdx.x]; }
    atomicAdd(out + ch, a);
    if (out_sq) atomicAdd(out_sq + ch, a2);
  }
}

static int grid1d(long total) {
  long g = (total + 255) / 256;
  const long cap = (long)num_sms() * 16;
  return (int)(g < cap ? g : cap);
}

}  // namespace yb

using namespace yb;

#define YB_BN_COMMON_CHECK(name)                                                                         \
  YB_REQUIRE(n > 0 && h > 0 && w > 0 && c >= 8 && c <= 2048 && (c & (c - 1)) == 0, name ": channels must be a power of two in [8, 2048]"); \
  YB_REQUIRE(dtype == YB_F16 || dtype == YB_BF16, name ": dtype must be f16 or bf16");

extern "C" int yb_bn_finalize(const float* sum, const float* sqsum, long count, int c, const float* gamma,
                              const float* beta, float eps, float decay, float* moving_mean, float* moving_var,
                              float* scale, float* shift, float* save_mean, float* save_invstd, void* stream) {
  YB_REQUIRE(gamma && beta && scale && shift && save_mean && save_invstd && c > 0 && count > 0, "bn_finalize: bad argument");
  YB_REQUIRE((sum == nullptr) == (sqsum == nullptr), "bn_finalize: sum/sqsum must both be given (both NULL: frozen BN)");
  YB_REQUIRE((moving_mean == nullptr) == (moving_var == nullptr), "bn_finalize: moving_mean/var must both be given");
  YB_REQUIRE(sum || moving_mean, "bn_finalize: frozen BN needs the moving statistics");
  bn_finalize_kernel<<<ceil_div(c, 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(
      sum, sqsum, (float)count, c, gamma, beta, eps, decay, moving_mean, moving_var, scale, shift, save_mean, save_invstd);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

extern "C" int yb_bn_act_apply(const void* z, long z_ld, const float* scale, const float* shift, const void* res,
                               long res_ld, void* out, long out_ld, int n, int h, int w, int c, int dtype, int leaky,
                               int upsample2x, void* stream) {
  YB_BN_COMMON_CHECK("bn_act_apply");
  YB_REQUIRE(z && scale && shift && out, "bn_act_apply: null pointer");
  RowGeom g{(long)n * h * w, h, w, c};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid;
  const StreamGeom sg = stream_geom(g.rows, c, &grid);
  if (dtype == YB_F16)
    bn_act_apply_kernel<__half><<<grid, 256, 0, st>>>((const __half*)z, z_ld, scale, shift, (const __half*)res, res_ld,
                                                     (__half*)out, out_ld, g, sg, leaky, upsample2x);
  else
    bn_act_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)z, z_ld, scale, shift,
                                                            (const __nv_bfloat16*)res, res_ld, (__nv_bfloat16*)out,
                                                            out_ld, g, sg, leaky, upsample2x);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

extern "C" int yb_bn_bwd_reduce_workspace_bytes(size_t* bytes) {
  YB_REQUIRE(bytes, "bn_bwd_reduce_workspace_bytes: null pointer");
  *bytes = 256 + (size_t)16 * 2 * 2048 * sizeof(float);   // ticket + BN_SLOTS x [2][c <= 2048]
  return YB_OK;
}

extern "C" int yb_bn_bwd_reduce(const void* dA, long dA_ld, const void* z, long z_ld, const float* scale,
                                const float* shift, const float* save_mean, const float* save_invstd, int n, int h,
                                int w, int c, int dtype, int leaky, int upsample2x, float* dgamma, float* dbeta,
                                void* workspace, void* stream) {
  YB_BN_COMMON_CHECK("bn_bwd_reduce");
  YB_REQUIRE(dA && z && scale && shift && save_mean && save_invstd && dgamma && dbeta, "bn_bwd_reduce: null pointer");
  RowGeom g{(long)n * h * w, h, w, c};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid;
  const StreamGeom sg = stream_geom(g.rows, c, &grid);
  // workspace (zero-initialised once by the caller): [0,256) ticket counter, then per-block partials
  unsigned int* ticket = static_cast<unsigned int*>(workspace);
  float* partial = workspace ? reinterpret_cast<float*>(static_cast<uint8_t*>(workspace) + 256) : nullptr;
  if (!workspace) {
    YB_CUDA(cudaMemsetAsync(dgamma, 0, c * 4, st));
    YB_CUDA(cudaMemsetAsync(dbeta, 0, c * 4, st));
  }
  if (dtype == YB_F16)
    bn_bwd_reduce_kernel<__half><<<grid, 256, 0, st>>>((const __half*)dA, dA_ld, (const __half*)z, z_ld, scale, shift,
                                                      save_mean, save_invstd, g, sg, leaky, upsample2x, dgamma, dbeta,
                                                      partial, ticket);
  else
    bn_bwd_reduce_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dA, dA_ld, (const __nv_bfloat16*)z,
                                                             z_ld, scale, shift, save_mean, save_invstd, g, sg, leaky,
                                                             upsample2x, dgamma, dbeta, partial, ticket);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

extern "C" int yb_bn_bwd_apply(const void* dA, long dA_ld, const void* z, long z_ld, const float* gamma,
                               const float* scale, const float* shift, const float* save_mean,
                               const float* save_invstd, const float* dgamma, const float* dbeta, int n, int h, int w,
                               int c, int dtype, int leaky, int upsample2x, int dilate2x, void* dz, long dz_ld,
                               void* stream) {
  YB_BN_COMMON_CHECK("bn_bwd_apply");
  YB_REQUIRE(dA && z && gamma && scale && shift && save_mean && save_invstd && dgamma && dbeta && dz,
             "bn_bwd_apply: null pointer");
  RowGeom g{(long)n * h * w, h, w, c};
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int grid;
  const StreamGeom sg = stream_geom(g.rows, c, &grid);
  if (dtype == YB_F16)
    bn_bwd_apply_kernel<__half><<<grid, 256, 0, st>>>((const __half*)dA, dA_ld, (const __half*)z, z_ld, gamma, scale,
                                                     shift, save_mean, save_invstd, dgamma, dbeta, g, sg, leaky,
                                                     upsample2x, dilate2x, (__half*)dz, dz_ld);
  else
    bn_bwd_apply_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>((const __nv_bfloat16*)dA, dA_ld, (const __nv_bfloat16*)z,
                                                            z_ld, gamma, scale, shift, save_mean, save_invstd, dgamma,
                                                            dbeta, g, sg, leaky, upsample2x, dilate2x,
                                                            (__nv_bfloat16*)dz, dz_ld);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

extern "C" int yb_col_stats(const void* x, long ld, long rows, int c, int dtype, float* sum, float* sqsum, void* stream);
extern "C" int yb_col_sum(const void* x, long ld, long rows, int c, int dtype, float* out, void* stream) {
  return yb_col_stats(x, ld, rows, c, dtype, out, nullptr, stream);
}
extern "C" int yb_col_stats(const void* x, long ld, long rows, int c, int dtype, float* out, float* out_sq, void* stream) {
  YB_REQUIRE(x && out && rows > 0 && c > 0, "col_sum: bad argument");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  YB_CUDA(cudaMemsetAsync(out, 0, c * 4, st));
  if (out_sq) YB_CUDA(cudaMemsetAsync(out_sq, 0, c * 4, st));
  const int gy = ceil_div(c, 32);
  long slabs = (long)num_sms() * 8 / gy;
  if (slabs < 1) slabs = 1;
  long rpb = (rows + slabs - 1) / slabs;
  if (rpb < 32) rpb = 32;
  dim3 grid(ceil_div(rows, rpb), gy), block(32, 8);
  if (dtype == YB_F16) col_sum_kernel<__half><<<grid, block, 0, st>>>((const __half*)x, ld, rows, c, rpb, out, out_sq);
  else if (dtype == YB_BF16) col_sum_kernel<__nv_bfloat16><<<grid, block, 0, st>>>((const __nv_bfloat16*)x, ld, rows, c, rpb, out, out_sq);
  else if (dtype == YB_F32) col_sum_kernel<float><<<grid, block, 0, st>>>((const float*)x, ld, rows, c, rpb, out, out_sq);
  else { set_error("col_sum: bad dtype"); return YB_ERR_UNSUPPORTED; }
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}
