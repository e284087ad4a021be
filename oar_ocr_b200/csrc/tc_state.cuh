// tc_state.cuh -- host-side state of the tensor-core engine shared by gemm_tc.cu and fused_tc.cu
#pragma once
#include <cuda.h>

#include <map>

#include "engine.cuh"

namespace oar {

struct TcWeights {
  uint4* packed = nullptr;  // [n_tile][k_block][hi|lo][k-chunk][BN rows][8 halfs]
  int N = 0, K = 0, BN = 0, n_tiles = 0, nkb = 0, KC = 4;
  bool rowtaps = false;  // packed for conv_rowtaps_tc: [n_tile][ky][cin block][kx][hi|lo][k-chunk][BN rows][8 halfs]
  int kh = 1, kw = 1;
};

struct TcState {
  std::map<int, TcWeights> w;
  std::map<int, TcWeights> wf;  // KC = 4 copies for the fused kernel where `w` holds a KC = 8 packing
  std::map<int, float*> dwp;    // depthwise taps for the fused kernel: [k-block][K*K taps | bias][32 ch], zero padded
  // narrow 3 x 3 convs with the kernel COLUMNS folded into N (conv_halo_tc.cu: fold mode): packed as a 3 x 1 convolution
  // with 3 * Cout output channels, row kx * Cout + co = the weights of tap column kx
  std::map<int, TcWeights> wfold;
};

// fp16 hi/lo split + packing into the shared-memory layout of the tcgen05 kernels (gemm_tc.cu)
TcWeights pack_weights(const float* w, int N, int K, bool force_kc4 = false);

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda)
EncodeTiledFn tmap_encoder();

}  // namespace oar
