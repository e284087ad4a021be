// tma_row_bench.cu -- what bounds a TMA box load of an fp32 NHWC tensor on B200: bytes, or the number of inner rows?
// The production kernels (fused_tc.cu, conv_halo_tc.cu) load [32 channels] x [cols] x [rows] boxes: every box row is one
// 128-byte segment of global memory.  This bench streams the same tensor through a shared-memory ring with boxes whose
// inner dimension is 32 / 64 channels (128 / 256-byte rows) and with 1-4 stages in flight, one producer lane per CTA,
// consumers that only hand the stage back.  148 CTAs x 1 (one per SM), persistent.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I oar_ocr_b200/csrc -o tma_row_bench tools/microbench/tma_row_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>

#include "tc_ptx.cuh"

using namespace oar;

struct Cfg {
  int cin_box;   // channels per box (32 or 64)
  int cols, rows;  // box pixels
  int stages;
  int n_boxes;   // per CTA
  int tiles_w, tiles_h, cblocks;
  uint32_t box_bytes;
};

__global__ void __launch_bounds__(64, 1) tma_bench(const Cfg c, const __grid_constant__ CUtensorMap tm, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar0 = sbase + c.stages * c.box_bytes;  // full[stages], empty[stages]
  if (threadIdx.x == 0) {
    for (int i = 0; i < c.stages; ++i) {
      mbar_init(bar0 + 8 * i, 1);
      mbar_init(bar0 + 8 * (c.stages + i), 1);
    }
    fence_mbar_init();
  }
  __syncthreads();
  const long long t0 = clock64();
  if (threadIdx.x == 0) {  // producer
    int item = blockIdx.x;
    for (int i = 0; i < c.n_boxes; ++i, item += gridDim.x) {
      const int s = i % c.stages, ph = (i / c.stages) & 1;
      mbar_wait(bar0 + 8 * (c.stages + s), ph ^ 1);
      const int cb = item % c.cblocks, r = item / c.cblocks;
      const int tw = r % c.tiles_w, r2 = r / c.tiles_w;
      const int th = r2 % c.tiles_h, b = r2 / c.tiles_h;
      mbar_expect_tx(bar0 + 8 * s, c.box_bytes);
      tma_load_4d(sbase + s * c.box_bytes, &tm, bar0 + 8 * s, cb * c.cin_box, tw * c.cols, th * c.rows, b);
    }
  } else if (threadIdx.x == 32) {  // consumer: wait, hand back
    for (int i = 0; i < c.n_boxes; ++i) {
      const int s = i % c.stages, ph = (i / c.stages) & 1;
      mbar_wait(bar0 + 8 * s, ph);
      mbar_arrive(bar0 + 8 * (c.stages + s));
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) cycles[blockIdx.x] = clock64() - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  cudaSetDevice(0);
  cudaFree(0);
  EncodeFn encode = nullptr;
  cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &q);
  if (!encode) {
    printf("no encoder\n");
    return 1;
  }
  // tensor: B x H x W x C fp32 = 32 x 240 x 240 x 96 (the detector neck's 708 MB) and 256 x 12 x 80 x 240 (rec block 5)
  struct Shape {
    int B, H, W, C;
    const char* name;
  } shapes[] = {{32, 240, 240, 96, "det neck 240x240x96"}, {256, 12, 80, 256, "rec 12x80x256"}};
  long long* d_cycles;
  cudaMalloc(&d_cycles, 148 * sizeof(long long));
  int clock_khz = 0;
  cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
  for (auto& sh : shapes) {
    size_t bytes = (size_t)sh.B * sh.H * sh.W * sh.C * 4;
    float* d;
    cudaMalloc(&d, bytes);
    cudaMemset(d, 0, bytes);
    struct Box {
      int cin, cols, rows;
    } boxes[] = {{32, 20, 12}, {64, 20, 12}, {32, 40, 12}, {64, 20, 6}, {32, 20, 6}, {64, 40, 6}, {32, 80, 3}, {64, 80, 3}, {32, 16, 4}, {64, 16, 4}};
    for (auto& bx : boxes) {
      if (sh.C % bx.cin) continue;
      for (int stages : {2, 3, 4, 6}) {
        Cfg c{};
        c.cin_box = bx.cin, c.cols = bx.cols, c.rows = bx.rows, c.stages = stages;
        c.box_bytes = (uint32_t)bx.cin * 4 * bx.cols * bx.rows;
        if ((size_t)stages * c.box_bytes + 256 > 220 * 1024) continue;
        c.tiles_w = sh.W / bx.cols, c.tiles_h = sh.H / bx.rows, c.cblocks = sh.C / bx.cin;
        if (c.tiles_w < 1 || c.tiles_h < 1) continue;
        const long long total = (long long)sh.B * c.tiles_h * c.tiles_w * c.cblocks;
        c.n_boxes = (int)(total / 148);
        if (c.n_boxes > 4000) c.n_boxes = 4000;
        if (c.n_boxes < 8) continue;
        CUtensorMap tm;
        cuuint64_t dims[4] = {(cuuint64_t)sh.C, (cuuint64_t)sh.W, (cuuint64_t)sh.H, (cuuint64_t)sh.B};
        cuuint64_t strides[3] = {(cuuint64_t)sh.C * 4, (cuuint64_t)sh.C * 4 * sh.W, (cuuint64_t)sh.C * 4 * sh.W * sh.H};
        cuuint32_t box[4] = {(cuuint32_t)bx.cin, (cuuint32_t)bx.cols, (cuuint32_t)bx.rows, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, d, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
          printf("%s box %dx%dx%d: encode failed %d\n", sh.name, bx.cin, bx.cols, bx.rows, (int)r);
          continue;
        }
        const size_t smem = (size_t)stages * c.box_bytes + 256 + 1024;
        cudaFuncSetAttribute(tma_bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0);
        cudaEventCreate(&e1);
        tma_bench<<<148, 64, smem>>>(c, tm, d_cycles);  // warm-up
        cudaEventRecord(e0);
        tma_bench<<<148, 64, smem>>>(c, tm, d_cycles);
        cudaEventRecord(e1);
        cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) {
          printf("launch failed: %s\n", cudaGetErrorString(err));
          return 1;
        }
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gb = 148.0 * c.n_boxes * c.box_bytes / 1e9;
        const double rows_per_box = (double)bx.cols * bx.rows;
        const double cyc_per_row = ms * 1e-3 * clock_khz * 1e3 / (c.n_boxes * rows_per_box);
        printf("%-22s box %2dch x %2d x %2d (%6u B, %3.0f rows of %3d B) stages %d: %7.3f ms %7.1f GB/s  %5.1f cycles/row\n", sh.name,
               bx.cin, bx.cols, bx.rows, c.box_bytes, rows_per_box, bx.cin * 4, stages, ms, gb / (ms * 1e-3), cyc_per_row);
      }
    }
    cudaFree(d);
  }
  return 0;
}
