// Hardware question behind the "flat-shift" implicit GEMM: can a K-major SWIZZLE_128B
// UMMA operand start at an arbitrary 128-byte row of a TMA-written tile (so that the 9 taps of a 3x3 conv are 9
// descriptors into ONE shared-memory tile), and what must the descriptor's base_offset field (bits 49-51) hold?
// A[512][64] bf16 is loaded by one 2D TMA box of 256 rows; B = identity; D[m][n] should equal A[m + shift][n].
// Also tests a non-natural group stride (SBO = 1280 B: 8-pixel rows of a 10-pixel-wide halo tile).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I yolov3_tensorflow_b200/csrc -I include -o build/umma_shift_probe tools/probes/umma_shift_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_bf16.h>
#include "yolob200.h"
#include "common.cuh"

using namespace yb;

__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int shift_rows,
             int sbo_bytes, int base_off, float* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                  // 256 rows x 128 B
  uint8_t* sB = smem + 256 * 128;      // 64 rows x 128 B
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 256 * 128 + 64 * 128);
  uint64_t* done = bar + 1;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(bar, 1); mbar_init(done, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc<64>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(bar, 256 * 128 + 64 * 128);
    tma_load_2d(sA, &tmA, bar, 0, 0);
    tma_load_2d(sB, &tmB, bar, 0, 0);
    mbar_wait(bar, 0);
    tcgen05_fence_after();
    constexpr uint32_t idesc = make_idesc_f16(128, 64, 1);
    for (int k = 0; k < 4; ++k) {
      uint64_t adesc = make_kmajor_desc(smem_u32(sA) + shift_rows * 128 + k * 32, sbo_bytes, 2u);
      adesc |= (uint64_t)(base_off & 7) << 49;
      const uint64_t bdesc = make_kmajor_desc(smem_u32(sB) + k * 32, 1024, 2u);
      umma_f16(tmem_base, adesc, bdesc, idesc, k != 0);
    }
    umma_commit(done);
  }
  __syncthreads();
  mbar_wait(done, 0);
  tcgen05_fence_after();
  for (int ch = 0; ch < 2; ++ch) {
    uint32_t r[32];
    tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32, r);
    tmem_ld_wait();
    for (int j = 0; j < 32; ++j) out[(warp * 32 + lane) * 64 + ch * 32 + j] = __uint_as_float(r[j]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<64>(tmem_base); }
}

static PFN_cuTensorMapEncodeTiled_v12000 enc;
static void make_map(CUtensorMap* tm, void* base, int rows, int box_rows) {
  cuuint64_t dims[2] = {64, (cuuint64_t)rows};
  cuuint64_t strides[1] = {128};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(1); }
}

int main() {
  cudaDriverEntryPointQueryResult qr; void* fn = nullptr;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  enc = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  std::vector<__nv_bfloat16> hA(512 * 64), hB(64 * 64);
  for (int i = 0; i < 512; ++i) for (int k = 0; k < 64; ++k) hA[i * 64 + k] = __float2bfloat16((float)((i * 7 + k * 3) % 13 - 6) + (float)(i % 5) * 16.f);
  for (int n = 0; n < 64; ++n) for (int k = 0; k < 64; ++k) hB[n * 64 + k] = __float2bfloat16(n == k ? 1.f : 0.f);
  __nv_bfloat16 *dA, *dB; float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, 128 * 64 * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  CUtensorMap tmA, tmB;
  make_map(&tmA, dA, 512, 256); make_map(&tmB, dB, 64, 64);
  const int smem = 256 * 128 + 64 * 128 + 1024 + 256;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> hO(128 * 64);
  const int shifts[] = {0, 1, 3, 8, 9, 29, 57};
  const int sbos[] = {1024, 1280};
  for (int sbo : sbos)
    for (int sh : shifts)
      for (int mode = 0; mode < 2; ++mode) {
        const int bo = mode ? (sh & 7) : 0;
        cudaMemset(dO, 0xff, 128 * 64 * 4);
        probe_kernel<<<1, 128, smem>>>(tmA, tmB, sh, sbo, bo, dO);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("sbo %d shift %d base_off %d: CUDA error %s\n", sbo, sh, bo, cudaGetErrorString(e)); return 1; }
        cudaMemcpy(hO.data(), dO, hO.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0, first = -1;
        for (int m = 0; m < 128; ++m) {
          const int src = sh + (m / 8) * (sbo / 128) + (m % 8);     // smem row the descriptor should read
          for (int n = 0; n < 64; ++n) {
            const float want = __bfloat162float(hA[src * 64 + n]);
            if (hO[m * 64 + n] != want) { if (first < 0) first = m * 64 + n; ++bad; }
          }
        }
        printf("sbo %4d shift %2d base_offset %d : %s (%d mismatches, first at row %d col %d)\n", sbo, sh, bo,
               bad ? "WRONG" : "ok", bad, first < 0 ? -1 : first / 64, first < 0 ? -1 : first % 64);
      }
  return 0;
}
