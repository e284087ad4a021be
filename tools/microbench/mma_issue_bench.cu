// mma_issue_bench.cu -- how long does ONE tcgen05.mma (M = 128, K = 16, kind::f16) take on B200 as a function of N, of
// the shared-memory operand layout (K-major no-swizzle, LBO / SBO as the production kernels use them), of the number of
// independent accumulators it alternates between, and of a 16-byte-row shifted A start address (conv_halo_tc's taps)?
// One CTA, operands static in shared memory, R back-to-back MMAs issued by one elected lane, clock64 around
// issue + commit + wait.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I oar_ocr_b200/csrc -o mma_issue_bench ...
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <cuda_fp16.h>

#include "tc_ptx.cuh"

using namespace oar;

struct Cfg {
  int N, chains, shift_rows, reps, lbo_a_pad;
};

__global__ void __launch_bounds__(128, 1) bench_kernel(Cfg c, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5;
  // A: 4 k-chunk pairs (K = 64) hi only, rows up to 300: chunk stride lbo_a
  const uint32_t lbo_a = 300 * 16 + c.lbo_a_pad;
  const uint32_t a_bytes = 8 * lbo_a;
  const uint32_t b_lbo = c.N * 16;
  const uint32_t off_b = (a_bytes + 127) & ~127u;
  const uint32_t b_bytes = 8 * b_lbo;
  const uint32_t off_ctrl = (off_b + b_bytes + 127) & ~127u;
  for (uint32_t i = tid; i < off_ctrl / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;  // 1.0h pairs
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + off_ctrl + 64);
  const uint32_t bar = sbase + off_ctrl;
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  if (warp == 1) {
    const uint32_t idesc = (1u << 4) | ((uint32_t)(c.N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t adesc0 = make_desc(0, lbo_a, 128), bdesc0 = make_desc(0, b_lbo, 128);
    const uint32_t a0 = sbase >> 4, b0 = (sbase + off_b) >> 4;
    long long t0 = 0, t1 = 0;
    for (int round = 0; round < 2; ++round) {  // round 0 warms up
      __syncwarp();
      t0 = clock64();
      // no divisions, no per-MMA branches: 8 MMAs per iteration, operands from adds on uniform values
      const uint32_t cmask = (uint32_t)c.chains - 1u;  // chains is a power of two
      const uint32_t sstep = (uint32_t)c.shift_rows;
      for (int r = 0; r < c.reps; r += 8) {
        const uint32_t acc = r ? 1u : 0u;
        if (elect_one_sync()) {
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            const uint32_t j = u & 3;
            const uint64_t ad = adesc0 + (uint64_t)(a0 + j * (lbo_a >> 3) + (u * sstep));
            const uint64_t bd = bdesc0 + (uint64_t)(b0 + j * (b_lbo >> 3));
            const uint32_t d = tmem_base + ((uint32_t)u & cmask) * (uint32_t)c.N;
            umma_f16(d, ad, bd, idesc, (u < c.chains) ? acc : 1u);
          }
        }
        __syncwarp();
      }
      if (elect_one_sync()) umma_commit(bar);
      __syncwarp();
      mbar_wait_warp(bar, (uint32_t)round & 1u);
      t1 = clock64();
    }
    if ((tid & 31) == 0) out[0] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 8);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const int reps = 512;
  printf("%5s %6s %10s %8s %12s %10s\n", "N", "chains", "shift_rows", "lbo_pad", "cycles/MMA", "math_cyc");
  for (int pad : {16, 32 + 112}) {
    for (int N : {32, 64, 128, 256}) {
      for (int chains : {1, 2, 4, 8}) {
        if (chains * N > 512) continue;
        for (int shift : {0, 1, 27}) {
          Cfg c{N, chains, shift, reps, pad};
          bench_kernel<<<1, 128, 190 * 1024>>>(c, d_out);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) {
            printf("error %s\n", cudaGetErrorString(e));
            return 1;
          }
          long long cyc;
          cudaMemcpy(&cyc, d_out, 8, cudaMemcpyDeviceToHost);
          printf("%5d %6d %10d %8d %12.1f %10.1f\n", N, chains, shift, pad, (double)cyc / reps, N / 2.0);
        }
      }
    }
  }
  return 0;
}
