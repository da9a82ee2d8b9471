// 3x3 convolution for the Cin <= 64 layers at the top of Darknet-53 (utils/layer_utils.py:35-44: Conv_1 32->64 /2,
// Conv_3 32->64, Conv_4 64->128 /2, Conv_6 / Conv_8 64->128) on the tensor cores, with the A operand taken from a
// shared-memory HALO TILE instead of nine im2col gathers.
//
// Why: with 32 / 64 input channels an im2col row is 64 / 128 bytes and one k-block is one filter tap, so the generic
// implicit-GEMM kernel (conv_igemm.cu) issues 9 TMA requests of 128 rows per 128-pixel tile — every input pixel is
// fetched 9x from L2, the TMA row rate (not bandwidth, not the MMA) paces the layer, and these five layers cost 1.4 ms
// of a 6.2 ms step (profiles/r02_b) against a ~0.5 ms HBM bound.  Here:
//   * a tile is 16 x 8 output pixels (UMMA M = 128: sixteen 8-row groups, group g = output row g);
//   * stride 1: ONE tiled TMA load brings the (16+2) x (8+2) input halo [18][10][Cin] into shared memory (borders and
//     the bottom tail zero-filled by the TMA), 1.4x instead of 9x the tile's pixels;
//     stride 2: four loads with traversal stride 2 bring the four (row, col)-parity planes of the 33 x 17 halo, so
//     that every tap again reads a dense window of one plane;
//   * tap (r, s) needs NO data movement: its A operand is the same shared-memory tile, addressed by a UMMA descriptor
//     that starts (dr * plane_width + ds) rows further down and uses the plane width as the 8-row-group stride (SBO).
//     The 128B / 64B swizzle is a function of the shared-memory address bits, so a descriptor may start at any row of
//     a TMA-written tile (tools/probes/umma_shift_probe.cu, profiles/r01_j_umma_shift_probe.txt);
//   * all 9 taps' weights [Cout][9 * Cin] stay resident in shared memory for the whole persistent CTA;
//   * the epilogue is the staging-tile + coalesced-store one of conv_igemm.cu, with pixel (not row) addressing.
// Warp roles: warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue; accumulators double-buffered in TMEM.
// Inference only (folded BN): scale/shift + leaky + optional residual, 16-bit NHWC in and out.
#include <cudaTypedefs.h>
#include <string.h>

#include <type_traits>

#include "common.cuh"
#include "conv.cuh"

namespace yb {

int make_tmap_2d(CUtensorMap* tm, const void* base, int dtype, long rows, long cols, long ld, int box_rows, int box_cols,
                 int weights);
int make_tmap_tiled4d(CUtensorMap* tm, const void* base, int dtype, int n, int h, int w, int c, long ld, int box_c,
                      int box_w, int box_h, int estride);

static constexpr int HT_H = 16, HT_W = 8;       // output tile
static constexpr int HALO_THREADS = 192;

template <int CIN, int COUT, int STRIDE>
struct HaloCfg {
  static constexpr int ROWB = CIN * 2;                                   // bytes per pixel row of a plane (one swizzle span)
  static constexpr int NPLANE = STRIDE == 1 ? 1 : 4;
  // plane geometry: stride 1: one [18][10] plane; stride 2: (odd|even rows) x (odd|even cols): 17|16 x 9|8
  static constexpr int ph(int p) { return STRIDE == 1 ? HT_H + 2 : ((p >> 1) == 0 ? HT_H + 1 : HT_H); }
  static constexpr int pw(int p) { return STRIDE == 1 ? HT_W + 2 : ((p & 1) == 0 ? HT_W + 1 : HT_W); }
  static constexpr int pbytes(int p) { return (ph(p) * pw(p) * ROWB + 1023) / 1024 * 1024; }
  static constexpr int poff(int p) { return p == 0 ? 0 : poff(p - 1) + pbytes(p - 1); }
  static constexpr int STAGE_BYTES = poff(NPLANE - 1) + pbytes(NPLANE - 1);
  static constexpr int STAGE_TX = STRIDE == 1 ? ph(0) * pw(0) * ROWB
                                              : (ph(0) * pw(0) + ph(1) * pw(1) + ph(2) * pw(2) + ph(3) * pw(3)) * ROWB;
  static constexpr int B_TAP_BYTES = COUT * ROWB;                        // one tap's [COUT][CIN] weight tile
  static constexpr int B_BYTES = 9 * B_TAP_BYTES;
  static constexpr int EPI_BYTES = 4 * 2 * 2048;                         // 4 warps x 2 staging tiles of [32][32] 16-bit
  static constexpr int MISC_BYTES = 1024;                                // barriers, TMEM slot, scale / shift
  static constexpr int BUDGET = 227 * 1024 - 1024 /*alignment slack*/;
  static constexpr int NST_RAW = (BUDGET - B_BYTES - EPI_BYTES - MISC_BYTES - 2 * COUT * 4) / STAGE_BYTES;
  static constexpr int NST = NST_RAW > 6 ? 6 : NST_RAW;
  static_assert(NST >= 1, "halo conv: configuration does not fit shared memory");
  static constexpr int SMEM_BYTES = 1024 + B_BYTES + NST * STAGE_BYTES + EPI_BYTES + MISC_BYTES + 2 * COUT * 4;
  static constexpr int TMEM_COLS = 2 * COUT;                             // 128 or 256
  static constexpr uint32_t SWIZZLE = CIN == 64 ? 2u : 4u;               // UMMA layout_type: 128B / 64B
  // tap (r, s) -> plane and row/col offset inside it.  stride 2 (pad 1 + VALID): input row 2i + r - 1:
  //   r = 0 -> odd-row plane, offset 0; r = 1 -> even-row plane, offset 0; r = 2 -> odd-row plane, offset 1
  static constexpr int tap_plane(int r, int s) { return STRIDE == 1 ? 0 : (((r == 1) ? 2 : 0) | ((s == 1) ? 1 : 0)); }
  static constexpr int tap_dr(int r) { return STRIDE == 1 ? r : (r == 2 ? 1 : 0); }
  static constexpr int tap_ds(int s) { return STRIDE == 1 ? s : (s == 2 ? 1 : 0); }
};

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

template <typename T, int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(HALO_THREADS, 1)
conv_halo_kernel(const __grid_constant__ HaloMaps maps, const __grid_constant__ HaloParams p) {
  using C = HaloCfg<CIN, COUT, STRIDE>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // pointer arithmetic: stays in the shared space
  uint8_t* sB = smem;                                        // [9][COUT][CIN]   swizzled, resident
  uint8_t* sA = smem + C::B_BYTES;                           // [NST][planes]    swizzled halo tiles
  uint8_t* sE = sA + C::NST * C::STAGE_BYTES;                // [4 warps][2][2 KB] epilogue staging
  uint64_t* bars = reinterpret_cast<uint64_t*>(sE + C::EPI_BYTES);
  uint64_t* full_bar = bars;            // [NST]
  uint64_t* empty_bar = bars + 8;       // [NST]
  uint64_t* tfull_bar = bars + 16;      // [2]
  uint64_t* tempty_bar = bars + 18;     // [2]
  uint64_t* b_bar = bars + 20;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);
  float* s_ss = reinterpret_cast<float*>(sE + C::EPI_BYTES + C::MISC_BYTES);   // [2][COUT] scale / shift

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
#pragma unroll
    for (int i = 0; i < C::NPLANE; ++i) tma_prefetch_desc(&maps.plane[i]);
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < C::NST; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    mbar_init(b_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  for (int c = threadIdx.x; c < COUT; c += HALO_THREADS) {
    s_ss[c] = c < p.cout ? __ldg(p.scale + c) : 0.f;
    s_ss[COUT + c] = c < p.cout ? __ldg(p.shift + c) : 0.f;
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      mbar_arrive_expect_tx(b_bar, (uint32_t)C::B_BYTES);
#pragma unroll
      for (int t = 0; t < 9; ++t) tma_load_2d(sB + t * C::B_TAP_BYTES, &maps.w, b_bar, t * CIN, 0);
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
        const int tx = tile % p.tiles_x;
        const int ty = (tile / p.tiles_x) % p.tiles_y;
        const int img = tile / (p.tiles_x * p.tiles_y);
        const int h0 = ty * HT_H, w0 = tx * HT_W;
        mbar_wait(&empty_bar[stage], phase ^ 1);
        mbar_arrive_expect_tx(&full_bar[stage], (uint32_t)C::STAGE_TX);
        uint8_t* dst = sA + stage * C::STAGE_BYTES;
        if (STRIDE == 1) {
          tma_load_4d(dst, &maps.plane[0], &full_bar[stage], 0, w0 - 1, h0 - 1, img);
        } else {
#pragma unroll
          for (int pl = 0; pl < 4; ++pl) {
            // plane (odd|even rows, odd|even cols): first input row 2 h0 - 1 (odd plane) or 2 h0 (even plane)
            const int hs = 2 * h0 - ((pl >> 1) == 0 ? 1 : 0), ws = 2 * w0 - ((pl & 1) == 0 ? 1 : 0);
            tma_load_4d(dst + C::poff(pl), &maps.plane[pl], &full_bar[stage], 0, ws, hs, img);
          }
        }
        if (++stage == C::NST) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_f16(128, COUT, std::is_same<T, __nv_bfloat16>::value);
      const uint32_t a_base = smem_u32(sA), b_base = smem_u32(sB);
      mbar_wait(b_bar, 0);
      int stage = 0, it = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
        mbar_wait(&full_bar[stage], phase);
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * COUT;
        const uint32_t st_base = a_base + stage * C::STAGE_BYTES;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
#pragma unroll
          for (int s = 0; s < 3; ++s) {
            const int pl = C::tap_plane(r, s);
            const uint32_t a_tap = st_base + C::poff(pl) + (C::tap_dr(r) * C::pw(pl) + C::tap_ds(s)) * C::ROWB;
            const uint32_t b_tap = b_base + (r * 3 + s) * C::B_TAP_BYTES;
#pragma unroll
            for (int k = 0; k < CIN / 16; ++k) {
              const uint64_t adesc = make_kmajor_desc(a_tap + k * 32, C::pw(pl) * C::ROWB, C::SWIZZLE);
              const uint64_t bdesc = make_kmajor_desc(b_tap + k * 32, 8 * C::ROWB, C::SWIZZLE);
              umma_f16(d_tmem, adesc, bdesc, idesc, (r | s | k) != 0);
            }
          }
        }
        umma_commit(&empty_bar[stage]);
        umma_commit(&tfull_bar[acc]);
        if (++stage == C::NST) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                  // TMEM lanes [32 q, 32 q + 32): output rows 4q .. 4q+3 of the tile
    uint8_t* stage2 = sE + (warp - 2) * 4096;
    const int sw = (lane >> 1) & 3;
    const int cq = lane & 3, cr0 = lane >> 2;      // coalesced layout: staging row 8k + cr0 (tile row 4q + k, col cr0), piece cq
    const bool has_res = p.res != nullptr;
    constexpr int NCH = COUT / 32;
    uint32_t cnt = 0;
    int it = 0;
    uint4 rnext[4];
    bool prefetched = false;
    auto pix_base = [&](int tile, long (&off)[4], bool (&ok)[4]) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int img = tile / (p.tiles_x * p.tiles_y);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int oh = ty * HT_H + 4 * quarter + k, ow = tx * HT_W + cr0;
        ok[k] = oh < p.ho && ow < p.wo;
        off[k] = ((long)img * p.ho + oh) * p.wo + ow;
      }
    };
    auto fetch_res = [&](const long (&off)[4], const bool (&ok)[4], int col0) {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        rnext[k] = ok[k] ? __ldg(reinterpret_cast<const uint4*>(static_cast<const T*>(p.res) + off[k] * p.res_ld + col0 + cq * 8))
                         : make_uint4(0u, 0u, 0u, 0u);
    };
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      long off[4], noff[4];
      bool ok[4], nok[4];
      pix_base(tile, off, ok);
      const int ntile = tile + gridDim.x;
      const bool has_next = ntile < p.num_tiles;
      if (has_res && has_next) pix_base(ntile, noff, nok);
      if (has_res && !prefetched) fetch_res(off, ok, 0);
      mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
      tcgen05_fence_after();
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * COUT;
      // all of the tile's TMEM loads in flight at once: tcgen05.ld is latency-bound (~1000 cycles, profiles/r02_b)
      uint32_t r0[32], r1[32], r2[32], r3[32];
      tmem_ld_32x32(t_row, r0);
      tmem_ld_32x32(t_row + 32, r1);
      if (NCH > 2) { tmem_ld_32x32(t_row + 64, r2); tmem_ld_32x32(t_row + 96, r3); }
      auto chunk = [&](const uint32_t (&r)[32], const int ch) {
        uint8_t* buf = stage2 + (cnt & 1u) * 2048;
        uint4 rcur[4];
        if (has_res) {
#pragma unroll
          for (int k = 0; k < 4; ++k) rcur[k] = rnext[k];
          const bool last = ch + 1 >= NCH;
          if (!last) fetch_res(off, ok, (ch + 1) * 32);
          else if (has_next) fetch_res(noff, nok, 0);
          if (last) prefetched = has_next;
        }
        float v[32];
        const float4* sc4 = reinterpret_cast<const float4*>(s_ss + ch * 32);
        const float4* sh4 = reinterpret_cast<const float4*>(s_ss + COUT + ch * 32);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 sc = sc4[j];
          const float4 sh = sh4[j];
          v[4 * j + 0] = fmaf(__uint_as_float(r[4 * j + 0]), sc.x, sh.x);
          v[4 * j + 1] = fmaf(__uint_as_float(r[4 * j + 1]), sc.y, sh.y);
          v[4 * j + 2] = fmaf(__uint_as_float(r[4 * j + 2]), sc.z, sh.z);
          v[4 * j + 3] = fmaf(__uint_as_float(r[4 * j + 3]), sc.w, sh.w);
        }
        if (p.leaky) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.1f * v[j]);
        }
        uint4* rowp = reinterpret_cast<uint4*>(buf + lane * 64);
        if (has_res) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const int R = 8 * k + cr0;
            *reinterpret_cast<uint4*>(buf + R * 64 + ((cq ^ ((R >> 1) & 3)) << 4)) = rcur[k];
          }
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const uint4 u = rowp[j ^ sw];
            float2 f;
            f = Pack2<T>::unpack(u.x); v[8 * j + 0] += f.x; v[8 * j + 1] += f.y;
            f = Pack2<T>::unpack(u.y); v[8 * j + 2] += f.x; v[8 * j + 3] += f.y;
            f = Pack2<T>::unpack(u.z); v[8 * j + 4] += f.x; v[8 * j + 5] += f.y;
            f = Pack2<T>::unpack(u.w); v[8 * j + 6] += f.x; v[8 * j + 7] += f.y;
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          uint4 pk;
          pk.x = Pack2<T>::pack(v[8 * j + 0], v[8 * j + 1]);
          pk.y = Pack2<T>::pack(v[8 * j + 2], v[8 * j + 3]);
          pk.z = Pack2<T>::pack(v[8 * j + 4], v[8 * j + 5]);
          pk.w = Pack2<T>::pack(v[8 * j + 6], v[8 * j + 7]);
          rowp[j ^ sw] = pk;
        }
        __syncwarp();
        T* outp = static_cast<T*>(p.out) + ch * 32 + cq * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int R = 8 * k + cr0;
          const uint4 u = *reinterpret_cast<const uint4*>(buf + R * 64 + ((cq ^ ((R >> 1) & 3)) << 4));
          if (ok[k]) *reinterpret_cast<uint4*>(outp + off[k] * p.out_ld) = u;
        }
        ++cnt;
      };
      tmem_ld_wait();
      chunk(r0, 0);
      chunk(r1, 1);
      if (NCH > 2) { chunk(r2, 2); chunk(r3, 3); }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
    }
  }

  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<C::TMEM_COLS>(tmem_base);
  }
}

template <typename T, int CIN, int COUT, int STRIDE>
static int launch_halo(const HaloMaps& maps, const HaloParams& p, cudaStream_t st) {
  using C = HaloCfg<CIN, COUT, STRIDE>;
  static DeviceOnce once;
  auto kern = conv_halo_kernel<T, CIN, COUT, STRIDE>;
  { const int rc = ensure_smem_attr(once, reinterpret_cast<const void*>(kern), C::SMEM_BYTES); if (rc) return rc; }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  kern<<<grid, HALO_THREADS, C::SMEM_BYTES, st>>>(maps, p);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

bool conv_halo_supported(const yb_conv_desc* d) {
  if (d->ksize != 3 || (d->stride != 1 && d->stride != 2)) return false;
  if (!(d->cin == 32 || d->cin == 64) || !(d->cout == 64 || d->cout == 128)) return false;
  if (d->cin == 64 && d->cout == 128 && d->stride == 2) return false;   // weights + one parity-plane stage exceed shared memory
  if (d->out_fp32 || d->upsample2x) return false;
  if (d->h % d->stride || d->w % d->stride || (d->w / d->stride) % HT_W) return false;
  if (d->in_ld % 8 || d->out_ld % 8 || d->in_ld < d->cin || d->out_ld < d->cout) return false;
  return d->dtype == YB_F16 || d->dtype == YB_BF16;
}

int conv_halo_prepare(const yb_conv_desc* d, const void* x, const void* w_packed, const float* scale, const float* shift,
                      const void* res, void* out, HaloMaps* maps, HaloParams* p) {
  YB_REQUIRE(conv_halo_supported(d), "conv_halo: unsupported configuration (k=%d s=%d cin=%d cout=%d w=%d)", d->ksize,
             d->stride, d->cin, d->cout, d->w);
  YB_REQUIRE(x && w_packed && scale && shift && out, "conv_halo: null pointer");
  YB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)out & 15) == 0 && ((uintptr_t)res & 15) == 0,
             "conv_halo: pointers must be 16-byte aligned");
  if (res) YB_REQUIRE(d->res_ld >= d->cout && d->res_ld % 8 == 0, "conv_halo: res_ld %d invalid", d->res_ld);
  memset(maps, 0, sizeof(*maps));
  p->n = d->n; p->ho = d->h / d->stride; p->wo = d->w / d->stride;
  p->tiles_x = p->wo / HT_W; p->tiles_y = ceil_div(p->ho, HT_H);
  p->num_tiles = p->tiles_x * p->tiles_y * d->n;
  p->cout = d->cout; p->leaky = d->leaky; p->scale = scale; p->shift = shift;
  p->res = res; p->res_ld = d->res_ld; p->out = out; p->out_ld = d->out_ld;
  int rc;
  if (d->stride == 1) {
    rc = make_tmap_tiled4d(&maps->plane[0], x, d->dtype, d->n, d->h, d->w, d->cin, d->in_ld, d->cin, HT_W + 2, HT_H + 2, 1);
    if (rc) return rc;
  } else {
    for (int pl = 0; pl < 4; ++pl) {
      // traversal stride 2: a box spanning 2 * count - 1 elements loads `count` of them
      const int rows = (pl >> 1) == 0 ? HT_H + 1 : HT_H, cols = (pl & 1) == 0 ? HT_W + 1 : HT_W;
      rc = make_tmap_tiled4d(&maps->plane[pl], x, d->dtype, d->n, d->h, d->w, d->cin, d->in_ld, d->cin, 2 * cols - 1,
                             2 * rows - 1, 2);
      if (rc) return rc;
    }
  }
  const long K = 9L * d->cin;
  return make_tmap_2d(&maps->w, w_packed, d->dtype, yb_conv_cout_pad(d->cout), K, K, d->cout, d->cin, 1);
}

int conv_halo_launch(const yb_conv_desc* d, const HaloMaps& maps, const HaloParams& p, cudaStream_t st) {
#define YB_HALO(T)                                                                              \
  if (d->cin == 32 && d->cout == 64 && d->stride == 1) return launch_halo<T, 32, 64, 1>(maps, p, st);   \
  if (d->cin == 32 && d->cout == 64 && d->stride == 2) return launch_halo<T, 32, 64, 2>(maps, p, st);   \
  if (d->cin == 32 && d->cout == 128 && d->stride == 1) return launch_halo<T, 32, 128, 1>(maps, p, st); \
  if (d->cin == 32 && d->cout == 128 && d->stride == 2) return launch_halo<T, 32, 128, 2>(maps, p, st); \
  if (d->cin == 64 && d->cout == 64 && d->stride == 1) return launch_halo<T, 64, 64, 1>(maps, p, st);   \
  if (d->cin == 64 && d->cout == 64 && d->stride == 2) return launch_halo<T, 64, 64, 2>(maps, p, st);   \
  if (d->cin == 64 && d->cout == 128 && d->stride == 1) return launch_halo<T, 64, 128, 1>(maps, p, st);
  if (d->dtype == YB_F16) { YB_HALO(__half) }
  else { YB_HALO(__nv_bfloat16) }
#undef YB_HALO
  set_error("conv_halo: no kernel for cin=%d cout=%d stride=%d", d->cin, d->cout, d->stride);
  return YB_ERR_UNSUPPORTED;
}

}  // namespace yb

using namespace yb;

extern "C" int yb_conv3x3_halo_supported(const yb_conv_desc* d) { return d && conv_halo_supported(d) ? 1 : 0; }

extern "C" int yb_conv3x3_halo_fwd(const yb_conv_desc* d, const void* x, const void* w_packed, const float* scale,
                                   const float* shift, const void* res, void* out, void* stream) {
  if (!d) { set_error("conv_halo: null descriptor"); return YB_ERR_INVALID_ARGUMENT; }
  HaloMaps maps;
  HaloParams p;
  int rc = conv_halo_prepare(d, x, w_packed, scale, shift, res, out, &maps, &p);
  if (rc) return rc;
  return conv_halo_launch(d, maps, p, static_cast<cudaStream_t>(stream));
}
