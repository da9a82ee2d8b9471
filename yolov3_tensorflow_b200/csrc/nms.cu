// Per-class score filter + greedy NMS, all classes and images in one grid.
// Replaces utils/nms_utils.py:8-48 (80 serialized TF sub-graphs per image) with
//   1. nms_compact_kernel : score >= thr  ->  per-(image,class) candidate lists
//      (block-aggregated slot reservation; list order is irrelevant, see below)
//   2. nms_select_kernel  : one CTA per (image,class).  Round r picks the best live
//      candidate (score desc, ties -> lower original index = TF's stable order) and
//      in the same sweep kills every live candidate whose IoU with the previous pick
//      is > thr.  That is exactly the TF CPU kernel's greedy rule: a candidate is
//      selected iff no previously selected box overlaps it by more than thr — but it
//      needs no sort, and stops after max_boxes rounds.
//   3. nms_gather_kernel  : concat classes ascending (utils/nms_utils.py:44-46),
//      also emitting the original box index of every kept box.
// IoU arithmetic is the TF CPU kernel's: float32, min/max-normalised
// corners, true divide, strict '>'; written with __f*_rn intrinsics so that no FMA
// contraction can change a rounding (indices must be bit-exact vs the oracle).
#include "common.cuh"

namespace yb {

static constexpr int COMPACT_THREADS = 256;
static constexpr int COMPACT_EPT = 16;          // elements per thread
static constexpr int SELECT_THREADS = 256;

__global__ void __launch_bounds__(COMPACT_THREADS)
nms_compact_kernel(const float* __restrict__ scores, int B, int C, int boxes_per_block, float thr,
                   int* __restrict__ cand_count, float* __restrict__ cand_score, int* __restrict__ cand_idx) {
  extern __shared__ int s_mem[];
  int* s_cnt = s_mem;        // [C]
  int* s_base = s_mem + C;   // [C]
  const int img = blockIdx.y;
  const int b0 = blockIdx.x * boxes_per_block;
  const int nb = min(boxes_per_block, B - b0);
  const int nelem = nb * C;
  for (int c = threadIdx.x; c < C; c += COMPACT_THREADS) s_cnt[c] = 0;
  __syncthreads();
  const float* src = scores + ((long)img * B + b0) * C;
  int lrank[COMPACT_EPT];
  float val[COMPACT_EPT];
#pragma unroll
  for (int k = 0; k < COMPACT_EPT; ++k) {
    const int e = threadIdx.x + k * COMPACT_THREADS;
    lrank[k] = -1;
    if (e < nelem) {
      const float s = src[e];
      val[k] = s;
      if (s >= thr) lrank[k] = atomicAdd(&s_cnt[e % C], 1);   // utils/nms_utils.py:30  (>=)
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < C; c += COMPACT_THREADS)
    s_base[c] = s_cnt[c] ? atomicAdd(&cand_count[img * C + c], s_cnt[c]) : 0;
  __syncthreads();
#pragma unroll
  for (int k = 0; k < COMPACT_EPT; ++k) {
    if (lrank[k] >= 0) {
      const int e = threadIdx.x + k * COMPACT_THREADS;
      const int c = e % C;
      const long seg = ((long)img * C + c) * B;
      const int slot = s_base[c] + lrank[k];
      cand_score[seg + slot] = val[k];
      cand_idx[seg + slot] = b0 + e / C;
    }
  }
}

struct BoxN { float x0, y0, x1, y1, area; };

__device__ __forceinline__ BoxN load_box(const float4* boxes, long i) {
  const float4 b = __ldg(boxes + i);
  BoxN r;
  r.x0 = fminf(b.x, b.z); r.x1 = fmaxf(b.x, b.z);
  r.y0 = fminf(b.y, b.w); r.y1 = fmaxf(b.y, b.w);
  r.area = __fmul_rn(__fsub_rn(r.x1, r.x0), __fsub_rn(r.y1, r.y0));
  return r;
}
// [TF] NonMaxSuppression CPU kernel IOU()
__device__ __forceinline__ bool iou_gt(const BoxN& a, const BoxN& b, float thr) {
  if (a.area <= 0.f || b.area <= 0.f) return false;
  const float iw = fmaxf(__fsub_rn(fminf(a.x1, b.x1), fmaxf(a.x0, b.x0)), 0.f);
  const float ih = fmaxf(__fsub_rn(fminf(a.y1, b.y1), fmaxf(a.y0, b.y0)), 0.f);
  const float inter = __fmul_rn(iw, ih);
  const float iou = __fdiv_rn(inter, __fsub_rn(__fadd_rn(a.area, b.area), inter));
  return iou > thr;
}
// monotone float -> uint (so that uint compare == float compare), -0 canonicalised to +0
__device__ __forceinline__ uint32_t f2ord(float f) {
  f = f + 0.0f;
  const uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(uint32_t u) {
  return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// block-wide max of a 64-bit key (all threads get the result)
__device__ __forceinline__ unsigned long long block_max_u64(unsigned long long v, unsigned long long* s_red,
                                                            unsigned long long* s_out) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, v, o);
    v = other > v ? other : v;
  }
  if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    unsigned long long t = threadIdx.x < SELECT_THREADS / 32 ? s_red[threadIdx.x] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, t, o);
      t = other > t ? other : t;
    }
    if (threadIdx.x == 0) *s_out = t;
  }
  __syncthreads();
  const unsigned long long r = *s_out;
  __syncthreads();
  return r;
}

static constexpr int BAND_CAP = 1024;      // candidates staged in shared memory per band
static constexpr int BAND_TARGET = 512;    // stop widening a band once it holds this many
static constexpr int HIST_BINS = 4096;     // top 12 bits of the order-preserving score key

// One CTA per (image, class).  Exact greedy NMS in descending-score BANDS:
//   1. a 12-bit histogram of the candidates' score keys below the current upper bound picks a lower bound so that the
//      band [lower, upper) holds ~1-2 k candidates (bisection inside one histogram bin if that bin alone overflows);
//   2. the band is staged in shared memory (key + normalised box), candidates overlapping an already selected box are
//      dropped, then rounds of (block-wide argmax || suppress-by-previous-pick) select from it;
//   3. if fewer than max_boxes were selected the next (strictly lower-scored) band is visited.
// Every candidate of a lower band scores below every candidate of a higher band, so this is exactly the TF kernel's
// visit order; typical inputs finish inside the first band.  More than BAND_CAP candidates with IDENTICAL score fall
// back to the unbanded global-memory sweep (use_global).
__global__ void __launch_bounds__(SELECT_THREADS)
nms_select_kernel(const float* __restrict__ boxes, int B, int C, int max_boxes, float iou_thr,
                  const int* __restrict__ cand_count, const float* __restrict__ cand_score,
                  int* __restrict__ cand_idx, int* __restrict__ sel_count, int* __restrict__ sel_idx,
                  float* __restrict__ sel_score) {
  extern __shared__ __align__(16) uint8_t nms_smem[];
  unsigned long long* bkey = reinterpret_cast<unsigned long long*>(nms_smem);        // [BAND_CAP]
  BoxN* bbox = reinterpret_cast<BoxN*>(bkey + BAND_CAP);                              // [BAND_CAP]
  BoxN* sbox = bbox + BAND_CAP;                                                       // [max_boxes]
  int* hist = reinterpret_cast<int*>(nms_smem);   // [HIST_BINS] aliases the band staging (used only before staging)
  __shared__ unsigned long long s_red[SELECT_THREADS / 32];
  __shared__ unsigned long long s_best;
  __shared__ int s_cnt, s_flag;
  const int c = blockIdx.x, img = blockIdx.y;
  const long segi = (long)img * C + c;
  const int cnt = min(cand_count[segi], B);
  const float* sc = cand_score + segi * B;
  int* ix = cand_idx + segi * B;
  const float4* bx = reinterpret_cast<const float4*>(boxes) + (long)img * B;
  int nsel = 0;
  bool use_global = false;
  unsigned long long upper = 1ull << 32;          // exclusive upper bound on the 32-bit score key

  while (nsel < max_boxes && !use_global) {
    // ---- 1. choose the band [lower, top) just below `upper` ----
    unsigned long long lower = 0ull;
    if (cnt <= BAND_CAP && upper == (1ull << 32)) {
      // common case (a few dozen candidates per class after the score filter): everything fits in one band
      goto stage_band;
    }
    {
    // highest remaining key
    unsigned long long mk = 0ull;
    for (int i = threadIdx.x; i < cnt; i += SELECT_THREADS) {
      const unsigned long long o = f2ord(sc[i]);
      if (o < upper && o + 1 > mk) mk = o + 1;
    }
    const unsigned long long top = block_max_u64(mk, s_red, &s_best);      // exclusive; 0: nothing left
    if (top == 0ull) break;
    // descending histogram below `top`: coarse bins of 2^20 keys, refined to 2^8 keys if the first bin overflows
    bool found = false;
    for (int level = 0; level < 2 && !found; ++level) {
      const int shift = level == 0 ? 20 : 8;
      for (int i = threadIdx.x; i < HIST_BINS; i += SELECT_THREADS) hist[i] = 0;
      __syncthreads();
      for (int i = threadIdx.x; i < cnt; i += SELECT_THREADS) {
        const unsigned long long o = f2ord(sc[i]);
        if (o < top) {
          const unsigned long long d = (top - 1 - o) >> shift;
          if (d < HIST_BINS) atomicAdd(&hist[(int)d], 1);
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        int acc = 0, nb = 0;
        for (int bsel = 0; bsel < HIST_BINS; ++bsel) {
          const int hb = hist[bsel];
          if (acc > 0 && acc + hb > BAND_CAP) break;
          acc += hb; nb = bsel + 1;
          if (acc > BAND_CAP || acc >= BAND_TARGET) break;
        }
        s_cnt = acc; s_flag = nb;
      }
      __syncthreads();
      const int acc = s_cnt, nb = s_flag;
      __syncthreads();
      if (acc <= BAND_CAP) {
        const unsigned long long span = (unsigned long long)nb << shift;
        lower = top > span ? top - span : 0ull;
        found = true;
      }
    }
    if (!found) {
      // > BAND_CAP candidates inside 256 consecutive key values: bisect; identical keys beyond capacity -> global sweep
      unsigned long long lo = top > 256 ? top - 256 : 0ull, hi = top;       // count([hi,top)) <= CAP < count([lo,top))
      while (hi - lo > 1) {
        const unsigned long long mid = lo + (hi - lo) / 2;
        if (threadIdx.x == 0) s_cnt = 0;
        __syncthreads();
        int local = 0;
        for (int i = threadIdx.x; i < cnt; i += SELECT_THREADS) {
          const unsigned long long o = f2ord(sc[i]);
          local += (o >= mid && o < top) ? 1 : 0;
        }
        if (local) atomicAdd(&s_cnt, local);
        __syncthreads();
        const int cm = s_cnt;
        __syncthreads();
        if (cm > BAND_CAP) lo = mid; else hi = mid;
      }
      if (hi >= top) { use_global = true; break; }
      lower = hi;
    }
    upper = top;
    }
  stage_band:
    // ---- 2. stage the band ----
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int i = threadIdx.x; i < cnt; i += SELECT_THREADS) {
      const uint32_t o = f2ord(sc[i]);
      if (o >= lower && o < upper) {
        const int id = ix[i];
        const int slot = atomicAdd(&s_cnt, 1);
        bkey[slot] = ((unsigned long long)o << 32) | (0xFFFFFFFFu - (uint32_t)id);
        bbox[slot] = load_box(bx, id);
      }
    }
    __syncthreads();
    const int n = s_cnt;
    // drop what the boxes selected in higher bands already suppress
    for (int j = threadIdx.x; j < n; j += SELECT_THREADS) {
      const BoxN bj = bbox[j];
      for (int k = 0; k < nsel; ++k)
        if (iou_gt(bj, sbox[k], iou_thr)) { bkey[j] = 0ull; break; }
    }
    __syncthreads();
    // ---- 3. greedy rounds inside the band ----
    unsigned long long lastkey = 0ull;
    BoxN lastb = {0, 0, 0, 0, 0};
    while (nsel < max_boxes) {
      unsigned long long best = 0ull;
      for (int j = threadIdx.x; j < n; j += SELECT_THREADS) {
        const unsigned long long kj = bkey[j];
        if (kj == 0ull) continue;
        if (lastkey != 0ull && (kj == lastkey || iou_gt(bbox[j], lastb, iou_thr))) { bkey[j] = 0ull; continue; }
        best = kj > best ? kj : best;
      }
      const unsigned long long k = block_max_u64(best, s_red, &s_best);
      if (k == 0ull) break;
      const int id = (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
      lastkey = k;
      lastb = load_box(bx, id);
      if (threadIdx.x == 0) {
        sel_idx[segi * max_boxes + nsel] = id;
        sel_score[segi * max_boxes + nsel] = ord2f((uint32_t)(k >> 32));
        sbox[nsel] = lastb;
      }
      ++nsel;
    }
    __syncthreads();
    upper = lower;
    if (upper == 0ull) break;
  }

  if (use_global) {
    // unbanded sweep over the global candidate list (marks dead entries in cand_idx)
    nsel = 0;
    int last = -1;
    BoxN lastb = {0, 0, 0, 0, 0};
    while (nsel < max_boxes) {
      unsigned long long best = 0ull;
      for (int i = threadIdx.x; i < cnt; i += SELECT_THREADS) {
        const int id = ix[i];
        if (id < 0) continue;
        if (last >= 0) {
          if (id == last || iou_gt(load_box(bx, id), lastb, iou_thr)) { ix[i] = -1; continue; }
        }
        const unsigned long long key = ((unsigned long long)f2ord(sc[i]) << 32) | (0xFFFFFFFFu - (uint32_t)id);
        best = key > best ? key : best;
      }
      const unsigned long long k = block_max_u64(best, s_red, &s_best);
      if (k == 0ull) break;
      last = (int)(0xFFFFFFFFu - (uint32_t)(k & 0xFFFFFFFFull));
      lastb = load_box(bx, last);
      if (threadIdx.x == 0) {
        sel_idx[segi * max_boxes + nsel] = last;
        sel_score[segi * max_boxes + nsel] = ord2f((uint32_t)(k >> 32));
      }
      ++nsel;
    }
  }
  if (threadIdx.x == 0) sel_count[segi] = nsel;
}

__global__ void __launch_bounds__(256)
nms_gather_kernel(const float* __restrict__ boxes, int B, int C, int max_boxes, const int* __restrict__ sel_count,
                  const int* __restrict__ sel_idx, const float* __restrict__ sel_score,
                  float* __restrict__ out_boxes, float* __restrict__ out_scores, int* __restrict__ out_labels,
                  int* __restrict__ out_indices, int* __restrict__ out_counts) {
  extern __shared__ int s_off[];   // [C+1]
  const int img = blockIdx.x;
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int c = 0; c < C; ++c) { s_off[c] = acc; acc += sel_count[(long)img * C + c]; }
    s_off[C] = acc;
    out_counts[img] = acc;
  }
  __syncthreads();
  const long cap = (long)C * max_boxes;
  for (long e = threadIdx.x; e < cap; e += blockDim.x) {
    const int c = (int)(e / max_boxes), k = (int)(e % max_boxes);
    const long segi = (long)img * C + c;
    if (k < sel_count[segi]) {
      const long o = (long)img * cap + s_off[c] + k;
      const int id = sel_idx[segi * max_boxes + k];
      reinterpret_cast<float4*>(out_boxes)[o] = __ldg(reinterpret_cast<const float4*>(boxes) + (long)img * B + id);
      out_scores[o] = sel_score[segi * max_boxes + k];
      out_labels[o] = c;
      out_indices[o] = id;
    }
  }
}

struct NmsWs {
  size_t cand_count, cand_score, cand_idx, sel_count, sel_idx, sel_score, total;
};
static NmsWs nms_layout(long n, long B, long C, long mb) {
  auto al = [](size_t v) { return (v + 255) & ~size_t(255); };
  NmsWs w;
  size_t o = 0;
  w.cand_count = o; o = al(o + n * C * 4);
  w.cand_score = o; o = al(o + n * C * B * 4);
  w.cand_idx = o;   o = al(o + n * C * B * 4);
  w.sel_count = o;  o = al(o + n * C * 4);
  w.sel_idx = o;    o = al(o + n * C * mb * 4);
  w.sel_score = o;  o = al(o + n * C * mb * 4);
  w.total = o;
  return w;
}

}  // namespace yb

using namespace yb;

extern "C" int yb_nms_workspace_bytes(int n_images, int num_boxes, int num_classes, int max_boxes, size_t* bytes) {
  YB_REQUIRE(bytes && n_images > 0 && num_boxes >= 0 && num_classes > 0 && max_boxes >= 0, "nms: bad argument");
  *bytes = nms_layout(n_images, num_boxes > 0 ? num_boxes : 1, num_classes, max_boxes > 0 ? max_boxes : 1).total;
  return YB_OK;
}

namespace yb {

// Workspace carve-up for callers that fill the candidate lists themselves (the fused detection-head epilogue).
int nms_candidate_buffers(void* workspace, size_t workspace_bytes, int n_images, int num_boxes, int num_classes,
                          int max_boxes, int** cand_count, float** cand_score, int** cand_idx) {
  const NmsWs w = nms_layout(n_images, num_boxes, num_classes, max_boxes > 0 ? max_boxes : 1);
  if (workspace_bytes < w.total) {
    set_error("nms: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return YB_ERR_WORKSPACE;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  *cand_count = reinterpret_cast<int*>(ws + w.cand_count);
  *cand_score = reinterpret_cast<float*>(ws + w.cand_score);
  *cand_idx = reinterpret_cast<int*>(ws + w.cand_idx);
  return YB_OK;
}

// Greedy selection + class-ascending gather over candidate lists that are already in the workspace.
int nms_select_gather(const float* boxes, int n_images, int num_boxes, int num_classes, int max_boxes, float iou_thresh,
                      void* workspace, size_t workspace_bytes, float* out_boxes, float* out_scores, int32_t* out_labels,
                      int32_t* out_indices, int32_t* out_counts, cudaStream_t st) {
  const NmsWs w = nms_layout(n_images, num_boxes, num_classes, max_boxes);
  if (workspace_bytes < w.total) {
    set_error("nms: workspace too small (%zu < %zu)", workspace_bytes, w.total);
    return YB_ERR_WORKSPACE;
  }
  uint8_t* ws = static_cast<uint8_t*>(workspace);
  int* cand_count = reinterpret_cast<int*>(ws + w.cand_count);
  float* cand_score = reinterpret_cast<float*>(ws + w.cand_score);
  int* cand_idx = reinterpret_cast<int*>(ws + w.cand_idx);
  int* sel_count = reinterpret_cast<int*>(ws + w.sel_count);
  int* sel_idx = reinterpret_cast<int*>(ws + w.sel_idx);
  float* sel_score = reinterpret_cast<float*>(ws + w.sel_score);
  dim3 g2(num_classes, n_images);
  static_assert(HIST_BINS * sizeof(int) <= BAND_CAP * (sizeof(unsigned long long) + sizeof(BoxN)), "histogram must fit the staging area");
  const size_t sel_smem = (size_t)BAND_CAP * (sizeof(unsigned long long) + sizeof(BoxN)) + (size_t)max_boxes * sizeof(BoxN);
  constexpr int SEL_SMEM_MAX = 200 * 1024;
  YB_REQUIRE(sel_smem <= (size_t)SEL_SMEM_MAX, "nms: max_boxes %d too large for the shared-memory staging", max_boxes);
  static DeviceOnce once;       // per device, not per process
  { const int rc = ensure_smem_attr(once, reinterpret_cast<const void*>(nms_select_kernel), SEL_SMEM_MAX); if (rc) return rc; }
  nms_select_kernel<<<g2, SELECT_THREADS, sel_smem, st>>>(boxes, num_boxes, num_classes, max_boxes, iou_thresh, cand_count,
                                                          cand_score, cand_idx, sel_count, sel_idx, sel_score);
  YB_CUDA(cudaGetLastError());
  nms_gather_kernel<<<n_images, 256, (num_classes + 1) * sizeof(int), st>>>(
      boxes, num_boxes, num_classes, max_boxes, sel_count, sel_idx, sel_score, out_boxes, out_scores, out_labels,
      out_indices, out_counts);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

}  // namespace yb

extern "C" int yb_nms(const float* boxes, const float* scores, int n_images, int num_boxes, int num_classes,
                      int max_boxes, float score_thresh, float iou_thresh, void* workspace, size_t workspace_bytes,
                      float* out_boxes, float* out_scores, int32_t* out_labels, int32_t* out_indices,
                      int32_t* out_counts, void* stream) {
  YB_REQUIRE(n_images > 0 && n_images <= 65535 && num_boxes >= 0 && num_classes > 0 && max_boxes >= 0,
             "nms: bad shape");
  YB_REQUIRE(num_classes <= COMPACT_THREADS * COMPACT_EPT, "nms: num_classes %d too large", num_classes);
  YB_REQUIRE(out_counts && workspace, "nms: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (num_boxes == 0 || max_boxes == 0) {
    YB_CUDA(cudaMemsetAsync(out_counts, 0, sizeof(int32_t) * n_images, st));
    return YB_OK;
  }
  YB_REQUIRE(boxes && scores && out_boxes && out_scores && out_labels && out_indices, "nms: null pointer");
  YB_REQUIRE(((uintptr_t)boxes & 15) == 0 && ((uintptr_t)out_boxes & 15) == 0, "nms: boxes must be 16-byte aligned");
  int* cand_count; float* cand_score; int* cand_idx;
  int rc = nms_candidate_buffers(workspace, workspace_bytes, n_images, num_boxes, num_classes, max_boxes, &cand_count,
                                 &cand_score, &cand_idx);
  if (rc) return rc;
  YB_CUDA(cudaMemsetAsync(cand_count, 0, sizeof(int) * (size_t)n_images * num_classes, st));
  int bpb = (COMPACT_THREADS * COMPACT_EPT) / num_classes;
  if (bpb < 1) bpb = 1;
  dim3 g1(ceil_div(num_boxes, bpb), n_images);
  nms_compact_kernel<<<g1, COMPACT_THREADS, 2 * num_classes * sizeof(int), st>>>(
      scores, num_boxes, num_classes, bpb, score_thresh, cand_count, cand_score, cand_idx);
  YB_CUDA(cudaGetLastError());
  return nms_select_gather(boxes, n_images, num_boxes, num_classes, max_boxes, iou_thresh, workspace, workspace_bytes,
                           out_boxes, out_scores, out_labels, out_indices, out_counts, st);
}
