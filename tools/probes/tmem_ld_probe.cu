// Hardware question behind the conv epilogues: how fast does tcgen05.ld move accumulators TMEM -> registers, per shape
// and with 1 or 4 warps (one per 32-lane quarter) reading concurrently?  Every variant reads the same 128 fp32 columns
// of the warp's 32 lanes (16 KB per warp), REPS times, and reports cycles per 128-column read.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -I yolov3_tensorflow_b200/csrc -I include -o build/tmem_ld_probe tools/probes/tmem_ld_probe.cu
#include <cstdio>
#include <cstdlib>
#include "yolob200.h"
#include "common.cuh"

using namespace yb;

#define LD_ASM_16(shape, num, regs)                                                                              \
  asm volatile("tcgen05.ld.sync.aligned." shape "." num ".b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(regs[0]), "=r"(regs[1]), "=r"(regs[2]), "=r"(regs[3]), "=r"(regs[4]), "=r"(regs[5]), "=r"(regs[6]),     \
                 "=r"(regs[7]), "=r"(regs[8]), "=r"(regs[9]), "=r"(regs[10]), "=r"(regs[11]), "=r"(regs[12]),                 \
                 "=r"(regs[13]), "=r"(regs[14]), "=r"(regs[15])                                                              \
               : "r"(addr) : "memory")

__device__ __forceinline__ void ld_32x32_x16(uint32_t addr, uint32_t (&r)[16]) { LD_ASM_16("32x32b", "x16", r); }
__device__ __forceinline__ void ld_16x64_x16(uint32_t addr, uint32_t (&r)[16]) { LD_ASM_16("16x64b", "x16", r); }     // 16 lanes x 32 cols
__device__ __forceinline__ void ld_16x128_x8(uint32_t addr, uint32_t (&r)[16]) { LD_ASM_16("16x128b", "x8", r); }     // 16 lanes x 32 cols
__device__ __forceinline__ void ld_16x256_x4(uint32_t addr, uint32_t (&r)[16]) { LD_ASM_16("16x256b", "x4", r); }     // 16 lanes x 32 cols

// mode 0: 32x32b.x32 x 4 loads, wait after each          (the round-1 epilogue)
// mode 1: 32x32b.x32 x 4 loads issued back to back, one wait
// mode 2: 32x32b.x16 x 8 loads back to back, one wait
// mode 3: 16x64b.x16  : 16 lanes x 32 cols per load -> 8 loads (2 lane halves x 4 column groups), one wait
// mode 4: 16x128b.x8  : same footprint per load
// mode 5: 16x256b.x4  : same footprint per load
// mode 6: no loads at all (loop overhead)
__global__ void __launch_bounds__(128, 1) probe_kernel(int mode, int active_warps, int reps, long long* out, uint32_t* sink) {
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t base = tmem_slot + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  long long t0 = 0, t1 = 0;
  if (warp < active_warps) {
    t0 = clock64();
    for (int it = 0; it < reps; ++it) {
      if (mode == 0) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld_32x32(base + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) acc ^= r[j];
        }
      } else if (mode == 1) {
        uint32_t r0[32], r1[32], r2[32], r3[32];
        tmem_ld_32x32(base, r0); tmem_ld_32x32(base + 32, r1); tmem_ld_32x32(base + 64, r2); tmem_ld_32x32(base + 96, r3);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) acc ^= r0[j] ^ r1[j] ^ r2[j] ^ r3[j];
      } else if (mode >= 2 && mode <= 5) {
        uint32_t r[8][16];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (mode == 2) ld_32x32_x16(base + q * 16, r[q]);
          else {
            const uint32_t a = base + ((uint32_t)((q & 1) * 16) << 16) + (q >> 1) * 32;     // lane half, 32-column group
            if (mode == 3) ld_16x64_x16(a, r[q]);
            else if (mode == 4) ld_16x128_x8(a, r[q]);
            else ld_16x256_x4(a, r[q]);
          }
        }
        tmem_ld_wait();
#pragma unroll
        for (int q = 0; q < 8; ++q)
#pragma unroll
          for (int j = 0; j < 16; ++j) acc ^= r[q][j];
      } else {
        acc ^= (uint32_t)it * 2654435761u;
      }
    }
    t1 = clock64();
  }
  if ((threadIdx.x & 31) == 0) out[warp] = t1 - t0;
  sink[threadIdx.x] = acc;
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 0) { tcgen05_fence_after(); tmem_dealloc<512>(tmem_slot); }
}

int main() {
  long long* d_out; uint32_t* d_sink;
  cudaMalloc(&d_out, 4 * sizeof(long long)); cudaMalloc(&d_sink, 128 * 4);
  const char* names[] = {"32x32b.x32, wait per load", "32x32b.x32 x4 in flight", "32x32b.x16 x8 in flight", "16x64b.x16 x8 in flight",
                         "16x128b.x8 x8 in flight", "16x256b.x4 x8 in flight", "no loads"};
  const int reps = 256;
  for (int aw = 1; aw <= 4; aw += 3) {
    for (int mode = 0; mode < 7; ++mode) {
      long long h[4] = {0, 0, 0, 0};
      probe_kernel<<<1, 128>>>(mode, aw, reps, d_out, d_sink);    // warm-up
      probe_kernel<<<1, 128>>>(mode, aw, reps, d_out, d_sink);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
      long long mx = 0;
      for (int w = 0; w < aw; ++w) mx = h[w] > mx ? h[w] : mx;
      const double cyc = (double)mx / reps;
      printf("%d warp(s) | %-28s | %8.1f cycles per 128-column read of 32 lanes (16 KB) | %6.1f B/clk/warp | %6.1f B/clk/SM\n", aw,
             names[mode], cyc, mode == 6 ? 0.0 : 16384.0 / cyc, mode == 6 ? 0.0 : aw * 16384.0 / cyc);
    }
  }
  return 0;
}
