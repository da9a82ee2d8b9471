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
int make_tmap_image3d(CUtensorMap* tm, const float* base, int n, int h, int w, int box_f, int box_h);
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

// Fused stem (darknet53_body/Conv, 3 -> 32, 3x3/1, utils/layer_utils.py:35) as the PRODUCER of Conv_1's parity planes:
// the 709 MB stem output of a batch-64 step is never written nor re-read.  Per 16x8-pixel tile of Conv_1 the stem is
// needed on 33 x 17 pixels; their 35 x 19-pixel float32 input halo arrives by one 3D TMA load, STEMW producer warps
// compute the stem on mma.sync (m16n8k16: A = the 27 -> 32 patch values gathered straight from the halo, B = the stem
// weights held in registers), apply BN + leaky and store the 16-bit results into the swizzled plane tiles the
// tcgen05 descriptors of Conv_1 read.  Stem pixels outside the image are Conv_1's zero padding (utils/layer_utils.py:15-16).
struct StemCfg {
  static constexpr int SH = 2 * HT_H + 1, SW = 2 * HT_W + 1;             // stem pixels per tile: 33 x 17
  static constexpr int NPX = SH * SW;                                     // 561
  static constexpr int NT16 = (NPX + 15) / 16;                            // 36 m16 tiles
  static constexpr int IN_ROWS = SH + 2;                                  // 35 image rows
  static constexpr int IN_ROWF = 60;                                      // floats per halo row: 19 px x 3 = 57, padded to 16 bytes
  static constexpr int IN_BYTES = (IN_ROWS * IN_ROWF * 4 + 127) / 128 * 128;
  static constexpr int NIN = 2;                                           // input halo stages
};

__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// 256-bit global store (sm_100: STG.E.ENL2.256): one full 32-byte sector per lane
__device__ __forceinline__ void st_global_v8(void* ptr, const uint32_t (&v)[8]) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(ptr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]),
               "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void mma16816_f16(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1, bool bf16) {
  if (bf16)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}

// EPIG = 2: a second group of four epilogue warps drains the other accumulator stage (tiles alternate between the groups);
// it uses the DIRECT epilogue only (no residual, dense 32-byte aligned rows: every lane stores its pixel's chunk as two
// 256-bit st.global — with out_ld = 64 a lane owns a whole 128-byte line — so the groups need no staging tiles).
template <typename T, int CIN, int COUT, int STRIDE, int STEMW = 0, int EPIG = 1>
__global__ void __launch_bounds__(64 + 128 * EPIG + 32 * STEMW, 1)
conv_halo_kernel(const __grid_constant__ HaloMaps maps, const __grid_constant__ HaloParams p) {
  using C = HaloCfg<CIN, COUT, STRIDE>;
  constexpr int NTHREADS = 64 + 128 * EPIG + 32 * STEMW;
  constexpr int PROD_WARP0 = 2 + 4 * EPIG;       // first stem-producer warp
  static_assert(STEMW == 0 || (CIN == 32 && STRIDE == 2), "the fused stem feeds Conv_1 (32 -> 64, stride 2)");
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
  uint64_t* in_full = bars + 24;        // [NIN] float32 input halo landed (TMA -> stem producers)
  uint64_t* in_empty = bars + 26;       // [NIN] stem producers -> TMA
  uint8_t* sIn = reinterpret_cast<uint8_t*>(s_ss + 2 * COUT);                  // [NIN][35][60] float32 (STEMW > 0 only; 128-byte aligned)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    if (STEMW == 0) {                            // (with the fused stem the plane maps are unused and zeroed)
#pragma unroll
      for (int i = 0; i < C::NPLANE; ++i) tma_prefetch_desc(&maps.plane[i]);
    }
    tma_prefetch_desc(&maps.w);
    for (int i = 0; i < C::NST; ++i) { mbar_init(&full_bar[i], STEMW > 0 ? STEMW : 1); mbar_init(&empty_bar[i], 1); }
    if (STEMW > 0) {
      tma_prefetch_desc(&maps.in3d);
      for (int i = 0; i < StemCfg::NIN; ++i) { mbar_init(&in_full[i], 1); mbar_init(&in_empty[i], STEMW); }
    }
    for (int i = 0; i < 2; ++i) { mbar_init(&tfull_bar[i], 1); mbar_init(&tempty_bar[i], 4); }
    mbar_init(b_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<C::TMEM_COLS>(tmem_slot);
  for (int c = threadIdx.x; c < COUT; c += NTHREADS) {
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
      if (STEMW > 0) {
        // float32 image halo of the tile: rows 2 h0 - 2 .. 2 h0 + 32, pixels 2 w0 - 2 .. 2 w0 + 16 (x 3 channels),
        // zero-filled outside the image (= the stem's SAME padding)
        for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
          const int tx = tile % p.tiles_x;
          const int ty = (tile / p.tiles_x) % p.tiles_y;
          const int img = tile / (p.tiles_x * p.tiles_y);
          mbar_wait(&in_empty[stage], phase ^ 1);
          mbar_arrive_expect_tx(&in_full[stage], (uint32_t)(StemCfg::IN_ROWS * StemCfg::IN_ROWF * 4));
          // (the innermost TMA coordinate must be 16-byte aligned — an unaligned start is an illegal instruction; the halo's
          //  first float (2 tx 8 - 2) * 3 is always 2 mod 4, so the box starts 2 floats earlier: 2 + 57 <= 60 floats per row)
          tma_load_3d(sIn + stage * StemCfg::IN_BYTES, &maps.in3d, &in_full[stage], (2 * tx * HT_W - 2) * 3 - 2, 2 * ty * HT_H - 2, img);
          if (++stage == StemCfg::NIN) { stage = 0; phase ^= 1; }
        }
      }
      for (int tile = blockIdx.x; STEMW == 0 && tile < p.num_tiles; tile += gridDim.x) {
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
  } else if (STEMW > 0 && warp >= PROD_WARP0) {
    // ===================== stem producers (warps PROD_WARP0 .. PROD_WARP0 + STEMW - 1) =====================
    const int pw_id = warp - PROD_WARP0;
    const bool bf16 = std::is_same<T, __nv_bfloat16>::value;
    const int g = lane >> 2, q = lane & 3;         // mma fragment coordinates: row group, column quad
    // B fragments: stem weights W[n][k] (k = (r*3+s)*3+c < 27), n8 tile nt, k-step ks: b0 = k 2q..2q+1, b1 = k + 8
    uint32_t bfrag[4][2][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int ks = 0; ks < 2; ++ks)
#pragma unroll
        for (int hb = 0; hb < 2; ++hb) {
          const int n = nt * 8 + g, k = ks * 16 + hb * 8 + 2 * q;
          const float w0 = k < 27 ? __ldg(p.stem_w + n * 27 + k) : 0.f;
          const float w1 = k + 1 < 27 ? __ldg(p.stem_w + n * 27 + k + 1) : 0.f;
          bfrag[nt][ks][hb] = Pack2<T>::pack(w0, w1);
        }
    float sc[4][2], sh[4][2];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        sc[nt][j] = __ldg(p.stem_scale + nt * 8 + 2 * q + j);
        sh[nt][j] = __ldg(p.stem_shift + nt * 8 + 2 * q + j);
      }
    // halo offsets of this thread's 8 patch elements: k -> row k / 9, float k % 9 of the 3 x 9 patch
    int koff[2][2][2];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks)
#pragma unroll
      for (int hb = 0; hb < 2; ++hb)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int k = ks * 16 + hb * 8 + 2 * q + j;
          koff[ks][hb][j] = k < 27 ? (k / 9) * StemCfg::IN_ROWF + (k % 9) : 0;
        }
    int in_stage = 0, stage = 0;
    uint32_t in_phase = 0, phase = 0;
    for (int tile = blockIdx.x; tile < p.num_tiles; tile += gridDim.x) {
      const int tx = tile % p.tiles_x;
      const int ty = (tile / p.tiles_x) % p.tiles_y;
      const int gy0 = 2 * ty * HT_H - 1, gx0 = 2 * tx * HT_W - 1;      // image coordinates of stem pixel (0, 0) of the tile
      mbar_wait(&in_full[in_stage], in_phase);
      mbar_wait(&empty_bar[stage], phase ^ 1);
      const float* halo = reinterpret_cast<const float*>(sIn + in_stage * StemCfg::IN_BYTES) + 2;   // see the TMA coordinate
      uint8_t* planes = sA + stage * C::STAGE_BYTES;
      // The 561 stem pixels are walked PLANE BY PLANE in m16 tiles (10 + 9 + 9 + 8 = 36): a tile lies inside one plane, so
      // the plane's constants are warp-uniform, consecutive fragment rows are consecutive 64-byte rows of the plane tile,
      // and the pixel -> (row, col) split is one constant division (the first version split every pixel index by 17 and
      // selected the plane per pixel: ~250 instructions per m16 tile, issue-bound at 6.5 k cycles per output tile).
      struct M16 {                                 // one m16 tile of stem pixels in flight
        uint32_t a[2][4];
        float acc[4][4];
        int rho[2], poff;
        bool live[2], inside[2];
      };
      auto gather = [&](const int t, M16& m) {
        const int pl = t < 10 ? 0 : (t < 19 ? 1 : (t < 28 ? 2 : 3));              // warp-uniform
        const int t0 = pl == 0 ? 0 : (pl == 1 ? 10 : (pl == 2 ? 19 : 28));
        const int pwid = (pl & 1) == 0 ? HT_W + 1 : HT_W;
        const int npl = ((pl >> 1) == 0 ? HT_H + 1 : HT_H) * pwid;                // pixels of this plane
        m.poff = pl == 0 ? C::poff(0) : (pl == 1 ? C::poff(1) : (pl == 2 ? C::poff(2) : C::poff(3)));
        const int ya = pl >> 1, xb = pl & 1;                                       // stem row = 2 * plane row + ya, col likewise
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {                       // fragment rows g and g + 8
          m.rho[hr] = (t - t0) * 16 + g + 8 * hr;
          m.live[hr] = m.rho[hr] < npl;
          const int r_ = m.live[hr] ? m.rho[hr] : 0;
          const int pr = (pl & 1) == 0 ? r_ / (HT_W + 1) : r_ >> 3;
          const int pc = r_ - pr * pwid;
          const int sy = 2 * pr + ya, sx = 2 * pc + xb;
          m.inside[hr] = m.live[hr] && gy0 + sy >= 0 && gy0 + sy < p.in_h && gx0 + sx >= 0 && gx0 + sx < p.in_w;
          const float* base = halo + sy * StemCfg::IN_ROWF + sx * 3;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks)
#pragma unroll
            for (int hb = 0; hb < 2; ++hb) {
              const float v0 = base[koff[ks][hb][0]];            // (k >= 27: offset 0, a finite value times a zero weight)
              const float v1 = base[koff[ks][hb][1]];
              m.a[ks][hb * 2 + hr] = Pack2<T>::pack(v0, v1);     // a0: row g, k lo; a1: row g+8, k lo; a2: row g, k hi; a3: row g+8, k hi
            }
        }
      };
      auto multiply = [&](M16& m) {
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
          m.acc[nt][0] = m.acc[nt][1] = m.acc[nt][2] = m.acc[nt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < 2; ++ks) mma16816_f16(m.acc[nt], m.a[ks], bfrag[nt][ks][0], bfrag[nt][ks][1], bf16);
        }
      };
      auto store = [&](const M16& m) {
#pragma unroll
        for (int hr = 0; hr < 2; ++hr) {
          if (!m.live[hr]) continue;
          uint8_t* row = planes + m.poff + m.rho[hr] * 64 + q * 4;
          const int swz = (m.rho[hr] >> 1) & 3;
          if (m.inside[hr]) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
              float v0 = fmaf(m.acc[nt][2 * hr], sc[nt][0], sh[nt][0]);
              float v1 = fmaf(m.acc[nt][2 * hr + 1], sc[nt][1], sh[nt][1]);
              v0 = fmaxf(v0, 0.1f * v0); v1 = fmaxf(v1, 0.1f * v1);      // leaky_relu(0.1), model.py:47
              *reinterpret_cast<uint32_t*>(row + ((nt ^ swz) << 4)) = Pack2<T>::pack(v0, v1);
            }
          } else {                                                        // outside the image: Conv_1's zero padding
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) *reinterpret_cast<uint32_t*>(row + ((nt ^ swz) << 4)) = 0u;
          }
        }
      };
      // The 561 stem pixels are walked PLANE BY PLANE in m16 tiles (10 + 9 + 9 + 8 = 36): a tile lies inside one plane, so
      // the plane's constants are warp-uniform and consecutive fragment rows are consecutive 64-byte rows of the plane tile.
      // The producers are latency-bound, not issue-bound: 12 warps beat 8 (408 vs 465 us); keeping TWO tiles in flight per
      // warp lost to the register pressure (498 / 538 us) and was dropped.
#pragma unroll 1
      for (int t = pw_id; t < StemCfg::NT16; t += STEMW) {
        M16 m0;
        gather(t, m0);
        multiply(m0);
        store(m0);
      }
      fence_proxy_async();                       // generic-proxy plane writes -> visible to tcgen05.mma (async proxy)
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(&full_bar[stage]);
        mbar_arrive(&in_empty[in_stage]);
      }
      if (++in_stage == StemCfg::NIN) { in_stage = 0; in_phase ^= 1; }
      if (++stage == C::NST) { stage = 0; phase ^= 1; }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int quarter = warp & 3;                  // TMEM lanes [32 q, 32 q + 32): output rows 4q .. 4q+3 of the tile
    const int grp = (warp - 2) >> 2;               // epilogue group: accumulator stage grp, tiles with iteration % EPIG == grp
    uint8_t* stage2 = sE + ((warp - 2) & 3) * 4096;          // (staging is used by the non-direct path, EPIG = 1 only)
    const bool direct = EPIG == 2 || p.direct != 0;
    const int sw = (lane >> 1) & 3;
    const int cq = lane & 3, cr0 = lane >> 2;      // coalesced layout: staging row 8k + cr0 (tile row 4q + k, col cr0), piece cq
    const bool has_res = STEMW == 0 && p.res != nullptr;      // (Conv_1 has no shortcut: the residual code is compiled out of the fused kernel)
    constexpr int NCH = COUT / 32;
    uint32_t cnt = 0;
    int it = grp;
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
    for (int tile = blockIdx.x + grp * gridDim.x; tile < p.num_tiles; tile += EPIG * gridDim.x, it += EPIG) {
      const int acc = it & 1;
      long off[4], noff[4];
      bool ok[4], nok[4];
      pix_base(tile, off, ok);
      const int ntile = tile + EPIG * gridDim.x;
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
      // direct path: this lane's pixel (tile row 4q + lane / 8, column lane % 8) and its 32 channels of a chunk = 64 bytes
      long my_off = 0;
      bool my_ok = false;
      if (direct) {
        const int tx = tile % p.tiles_x;
        const int ty = (tile / p.tiles_x) % p.tiles_y;
        const int img = tile / (p.tiles_x * p.tiles_y);
        const int oh = ty * HT_H + 4 * quarter + (lane >> 3), ow = tx * HT_W + (lane & 7);
        my_ok = oh < p.ho && ow < p.wo;
        my_off = (((long)img * p.ho + oh) * p.wo + ow) * p.out_ld;
      }
      auto chunk_direct = [&](const uint32_t (&r)[32], const int ch) {
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
        if (my_ok) {
          uint32_t lo[8], hi[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            lo[j] = Pack2<T>::pack(v[2 * j], v[2 * j + 1]);
            hi[j] = Pack2<T>::pack(v[16 + 2 * j], v[16 + 2 * j + 1]);
          }
          T* dst = static_cast<T*>(p.out) + my_off + ch * 32;
          st_global_v8(dst, lo);
          st_global_v8(dst + 16, hi);
        }
      };
      tmem_ld_wait();
      if (direct) {
        chunk_direct(r0, 0);
        chunk_direct(r1, 1);
        if (NCH > 2) { chunk_direct(r2, 2); chunk_direct(r3, 3); }
      } else if (EPIG == 1) {
        chunk(r0, 0);
        chunk(r1, 1);
        if (NCH > 2) { chunk(r2, 2); chunk(r3, 3); }
      }
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

template <typename T, int STEM_WARPS, int EPIG>
static int launch_stem_halo(const HaloMaps& maps, const HaloParams& p, cudaStream_t st) {
  using C = HaloCfg<32, 64, 2>;
  constexpr int SMEM = C::SMEM_BYTES + StemCfg::NIN * StemCfg::IN_BYTES;
  static_assert(SMEM <= 227 * 1024, "fused stem + Conv_1 does not fit shared memory");
  static DeviceOnce once;
  auto kern = conv_halo_kernel<T, 32, 64, 2, STEM_WARPS, EPIG>;
  { const int rc = ensure_smem_attr(once, reinterpret_cast<const void*>(kern), SMEM); if (rc) return rc; }
  const int grid = p.num_tiles < num_sms() ? p.num_tiles : num_sms();
  kern<<<grid, 64 + 128 * EPIG + 32 * STEM_WARPS, SMEM, st>>>(maps, p);
  YB_CUDA(cudaGetLastError());
  return YB_OK;
}

int conv_stem_halo_prepare(const yb_conv_desc* d, const float* image, const float* stem_w, const float* stem_scale,
                           const float* stem_shift, const void* w_packed, const float* scale, const float* shift, void* out,
                           HaloMaps* maps, HaloParams* p) {
  YB_REQUIRE(d->ksize == 3 && d->stride == 2 && d->cin == 32 && d->cout == 64 && !d->out_fp32 && !d->upsample2x,
             "conv_stem_halo: Conv_1 is 3x3/2 32->64 (got k=%d s=%d %d->%d)", d->ksize, d->stride, d->cin, d->cout);
  YB_REQUIRE(d->h % 2 == 0 && d->w % 16 == 0 && d->out_ld % 8 == 0 && d->out_ld >= 64, "conv_stem_halo: bad geometry (h=%d w=%d)", d->h, d->w);
  YB_REQUIRE(image && stem_w && stem_scale && stem_shift && w_packed && scale && shift && out, "conv_stem_halo: null pointer");
  YB_REQUIRE(((uintptr_t)image & 15) == 0 && ((uintptr_t)w_packed & 15) == 0 && ((uintptr_t)out & 15) == 0, "conv_stem_halo: pointers must be 16-byte aligned");
  memset(maps, 0, sizeof(*maps));
  memset(p, 0, sizeof(*p));
  p->n = d->n; p->ho = d->h / 2; p->wo = d->w / 2;
  p->tiles_x = p->wo / HT_W; p->tiles_y = ceil_div(p->ho, HT_H);
  p->num_tiles = p->tiles_x * p->tiles_y * d->n;
  p->cout = d->cout; p->leaky = d->leaky; p->scale = scale; p->shift = shift;
  p->res = nullptr; p->res_ld = 0; p->out = out; p->out_ld = d->out_ld;
  p->stem_w = stem_w; p->stem_scale = stem_scale; p->stem_shift = stem_shift; p->in_h = d->h; p->in_w = d->w;
  p->direct = (d->out_ld % 16 == 0 && ((uintptr_t)out & 31) == 0 && opt("YB_HALO_DIRECT")[0] != '0') ? 1 : 0;
  int rc = make_tmap_image3d(&maps->in3d, image, d->n, d->h, d->w, StemCfg::IN_ROWF, StemCfg::IN_ROWS);
  if (rc) return rc;
  const long K = 9L * d->cin;
  return make_tmap_2d(&maps->w, w_packed, d->dtype, yb_conv_cout_pad(d->cout), K, K, d->cout, d->cin, 1);
}

int conv_stem_halo_launch(const yb_conv_desc* d, const HaloMaps& maps, const HaloParams& p, cudaStream_t st) {
  // A/B switches: YB_STEM_WARPS = 8 | 12 producer warps, YB_STEM_EPIG = 1 | 2 epilogue groups
  const bool w12 = opt("YB_STEM_WARPS")[0] != '8';      // default: twelve producer warps
  const bool direct_ok = d->out_ld % 16 == 0 && ((uintptr_t)p.out & 31) == 0;
  const bool eg2 = direct_ok && opt("YB_STEM_EPIG")[0] == '2';
#define YB_STEM_LAUNCH(T)                                                        \
  if (eg2) return w12 ? launch_stem_halo<T, 12, 2>(maps, p, st) : launch_stem_halo<T, 8, 2>(maps, p, st); \
  return w12 ? launch_stem_halo<T, 12, 1>(maps, p, st) : launch_stem_halo<T, 8, 1>(maps, p, st);
  if (d->dtype == YB_F16) { YB_STEM_LAUNCH(__half) }
  if (d->dtype == YB_BF16) { YB_STEM_LAUNCH(__nv_bfloat16) }
#undef YB_STEM_LAUNCH
  set_error("conv_stem_halo: dtype must be f16 or bf16");
  return YB_ERR_UNSUPPORTED;
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
  memset(p, 0, sizeof(*p));
  p->n = d->n; p->ho = d->h / d->stride; p->wo = d->w / d->stride;
  p->tiles_x = p->wo / HT_W; p->tiles_y = ceil_div(p->ho, HT_H);
  p->num_tiles = p->tiles_x * p->tiles_y * d->n;
  p->cout = d->cout; p->leaky = d->leaky; p->scale = scale; p->shift = shift;
  p->res = res; p->res_ld = d->res_ld; p->out = out; p->out_ld = d->out_ld;
  // direct 256-bit stores when there is no residual and the rows are 32-byte aligned
  p->direct = (!res && d->out_ld % 16 == 0 && ((uintptr_t)out & 31) == 0 && opt("YB_HALO_DIRECT")[0] != '0') ? 1 : 0;
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

// Stem + Conv_1 in one kernel (utils/layer_utils.py:35-36): image float32 [n, h, w, 3] -> Conv_1's output [n, h/2, w/2, out_ld].
// d describes Conv_1 (h, w = image size = the stem's output size).  The stem's BN is folded into stem_scale / stem_shift.
extern "C" int yb_stem_conv1_fused_fwd(const yb_conv_desc* d, const float* image, const float* stem_w_ohwi,
                                       const float* stem_scale, const float* stem_shift, const void* w_packed,
                                       const float* scale, const float* shift, void* out, void* stream) {
  if (!d) { set_error("stem_conv1_fused: null descriptor"); return YB_ERR_INVALID_ARGUMENT; }
  HaloMaps maps;
  HaloParams p;
  int rc = conv_stem_halo_prepare(d, image, stem_w_ohwi, stem_scale, stem_shift, w_packed, scale, shift, out, &maps, &p);
  if (rc) return rc;
  return conv_stem_halo_launch(d, maps, p, static_cast<cudaStream_t>(stream));
}
