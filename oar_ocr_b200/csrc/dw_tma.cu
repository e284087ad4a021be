// dw_tma.cu -- stand-alone depthwise k x k convolution (k = 3 / 5) on TMA-staged shared-memory tiles.
//
// In the reference these are depthwise Conv nodes inside ONNX Runtime (ort_infer_execution.rs:178,281).  Most of them
// are fused with their 1x1 neighbour (fused_tc.cu); the ones that are not -- a squeeze-excite gate needs the whole
// image's pool before the 1x1 conv can start, and a 5x5 block wider than one N tile is cheaper split -- ran on a
// register-tiled LDG kernel (engine.cu: dwconv_tiled_kernel) at ~3.0 TB/s: 128 registers, 4 CTAs per SM, 48 LDG.128
// in flight per thread and an L1 that re-serves every input value six times.  This kernel is the depthwise stage of
// lcblock_tc on its own: a TMA producer streams [32 ch] x [tile + halo] boxes of the fp32 NHWC input (zero fill outside
// the image = the padding) with the k-block's taps + bias behind them into a 4-6 deep ring, two teams of 8 warps take
// alternate (pixel tile, k-block) items, thread = 2 channels x (2 x 4) output pixels, every input value read once from
// shared memory (LDS.64, a pixel's 128 bytes per half-warp) and accumulated bias-first in (ky, kx) order with FFMA2 --
// the same arithmetic, bit for bit, as the fused kernel's and engine 1's.  Without weight stages, A buffers and
// staging tiles the ring holds ~200 KB: enough bytes in flight for the HBM rate (tools/microbench/tma_row_bench.cu).
// Outputs leave straight from registers (a half-warp writes one pixel's 128 contiguous bytes per store).
// Optionally the per-(image, tile) channel sums the squeeze-excite pool reads instead of the tensor (se_gap_kernel).
#include <cuda.h>

#include <cstdlib>

#include "engine.cuh"
#include "tc_ptx.cuh"
#include "tc_state.cuh"

namespace oar {

constexpr int DT_THREADS = 32 + 512;  // producer warp + 2 teams x 8 depthwise warps
constexpr int DT_MAX_IN = 6;
constexpr size_t DT_SMEM_MAX = 227 * 1024;
enum { DT_IN_FULL = 0, DT_IN_EMPTY = 6, DT_NBAR = 12 };

struct DtParams {
  const float* dw_pk;  // [k-block][K*K taps | bias][32 channels]
  int act;
  float ps, pb;
  float* out;
  float* tile_sums;  // [image][tile][C] or null
  int C, nkb, H, W, Ho, Wo;
  int TH, TW, tiles_h, tiles_w, n_work;  // n_work = images * tiles (pixel tiles); a CTA walks the k-blocks of its tiles
  int cols_in;
  uint32_t in_bytes, tap_bytes;
  int ns_in;
};

__device__ __forceinline__ float dt_act(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_HSWISH: return v * fminf(fmaxf(v + 3.0f, 0.0f), 6.0f) * 0.16666667f;
    case ACT_SWISH: return __fdividef(v, 1.0f + __expf(-v));
    case ACT_SIGMOID: return __fdividef(1.0f, 1.0f + __expf(-v));
    case ACT_HSIGMOID: return fminf(fmaxf(v * 0.16666667f + 0.5f, 0.0f), 1.0f);
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    default: return v;
  }
}

template <int K, int SH, int SW>
__global__ void __launch_bounds__(DT_THREADS, 1) dw_tma_kernel(const DtParams P, const __grid_constant__ CUtensorMap tm_in) {
  extern __shared__ __align__(1024) uint8_t dt_smem_raw[];
  uint8_t* smem = dt_smem_raw + ((1024u - (smem_u32(dt_smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t stage_bytes = P.in_bytes + P.tap_bytes;
  const uint32_t off_ctrl = (uint32_t)P.ns_in * stage_bytes;
  const uint32_t bar0 = sbase + off_ctrl;
#define DT_BAR(i) (bar0 + 8u * (uint32_t)(i))
  float2* red = reinterpret_cast<float2*>(smem + off_ctrl + 128);  // [team][8 warps][16 pairs] tile-sum partials
  if (tid == 0) {
    for (int i = 0; i < DT_MAX_IN; ++i) {
      mbar_init(DT_BAR(DT_IN_FULL + i), 1);
      mbar_init(DT_BAR(DT_IN_EMPTY + i), 8);
    }
    fence_mbar_init();
  }
  __syncthreads();
  // item i of this CTA = (pixel tile blockIdx.x + (i / nkb) * gridDim.x, k-block i % nkb): k-blocks fastest, so the tile
  // coordinates (three integer divisions) change once per nkb items and the loops below carry counters instead
  const uint32_t n_tiles_cta = P.n_work > (int)blockIdx.x ? (uint32_t)((P.n_work - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) : 0u;
  const uint32_t n_items = n_tiles_cta * (uint32_t)P.nkb;

  if (warp == 0) {
    if (lane == 0) {
      int sp = blockIdx.x, kb = 0;
      int tw = sp % P.tiles_w, r = sp / P.tiles_w;
      uint32_t s = 0, ph = 0;
      for (uint32_t it = 0; it < n_items; ++it) {
        mbar_wait(DT_BAR(DT_IN_EMPTY + s), ph ^ 1u);
        const uint32_t dst = sbase + s * stage_bytes;
        mbar_expect_tx(DT_BAR(DT_IN_FULL + s), stage_bytes);
        tma_load_4d(dst, &tm_in, DT_BAR(DT_IN_FULL + s), kb * 32, tw * P.TW * SW - K / 2, (r % P.tiles_h) * P.TH * SH - K / 2,
                    r / P.tiles_h);
        bulk_load(dst + P.in_bytes, reinterpret_cast<const uint8_t*>(P.dw_pk) + (size_t)kb * P.tap_bytes, P.tap_bytes,
                  DT_BAR(DT_IN_FULL + s));
        if (++s == (uint32_t)P.ns_in) s = 0, ph ^= 1u;
        if (++kb == P.nkb) {
          kb = 0, sp += gridDim.x;
          tw = sp % P.tiles_w, r = sp / P.tiles_w;
        }
      }
    }
  } else {
    constexpr int RT = SH + K;      // input rows feeding a thread's 2 output rows
    constexpr int CT = 3 * SW + K;  // input columns feeding its 4 output columns
    const int team = (tid - 32) >> 8, ct = (tid - 32) & 255;
    const int pair = ct & 15, pg = ct >> 4;
    const int gx = P.TW >> 2;
    const int pgy = pg / gx, pgx = pg - pgy * gx;
    const bool active = pgy < (P.TH >> 1);
    const uint32_t row_stride = (uint32_t)P.cols_in * 128u;
    const uint32_t in_off = (uint32_t)(2 * pgy * SH) * row_stride + (uint32_t)(4 * pgx * SW) * 128u + (uint32_t)pair * 8u;
    const bool hsw = P.act == ACT_HSWISH;
    const bool affine = P.ps != 1.0f || P.pb != 0.0f;
    const int tiles_per_img = P.tiles_h * P.tiles_w;
    // counters of this team's items: it = team, team + 2, ...
    int kb = team % P.nkb, sp = (int)blockIdx.x + (team / P.nkb) * (int)gridDim.x;
    int tw = sp % P.tiles_w, th = (sp / P.tiles_w) % P.tiles_h, img = sp / P.tiles_w / P.tiles_h;
    uint32_t s = (uint32_t)team % (uint32_t)P.ns_in, ph = ((uint32_t)team / (uint32_t)P.ns_in) & 1u;
    for (uint32_t it = (uint32_t)team; it < n_items; it += 2) {
      // the ring depth is even (host), so a stage belongs to one team for good and a plain parity wait is sound
      mbar_wait(DT_BAR(DT_IN_FULL + s), ph);
      const uint8_t* stage = smem + s * stage_bytes;
      const float2* taps = reinterpret_cast<const float2*>(stage + P.in_bytes) + pair;
      const float2 bv = taps[K * K * 16];
      float2 acc[2][4];
#pragma unroll
      for (int ty = 0; ty < 2; ++ty)
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) acc[ty][tx] = bv;
      if (active) {
        const uint8_t* base = stage + in_off;
        float2 w[K][K];
#pragma unroll
        for (int iy = 0; iy < RT; ++iy) {
          float2 x[CT];
#pragma unroll
          for (int cx = 0; cx < CT; ++cx) x[cx] = *reinterpret_cast<const float2*>(base + iy * row_stride + cx * 128);
          if (iy < K) {
#pragma unroll
            for (int kx = 0; kx < K; ++kx) w[iy][kx] = taps[(iy * K + kx) * 16];
          }
#pragma unroll
          for (int ty = 0; ty < 2; ++ty) {
            const int ky = iy - ty * SH;
            if (ky < 0 || ky >= K) continue;
#pragma unroll
            for (int kx = 0; kx < K; ++kx)
#pragma unroll
              for (int tx = 0; tx < 4; ++tx) acc[ty][tx] = __ffma2_rn(x[tx * SW + kx], w[ky][kx], acc[ty][tx]);
          }
        }
      }
      // activation, affine, store, tile sums -- all from registers; the stage goes back once its values were consumed
      // (the arrive is made to depend on the accumulators: fused_tc.cu, dep_zero)
      const int oy0 = th * P.TH + 2 * pgy, ox0 = tw * P.TW + 4 * pgx;
      const int ch = kb * 32 + 2 * pair;
      float2 ssum = make_float2(0.f, 0.f);
#pragma unroll
      for (int ty = 0; ty < 2; ++ty)
#pragma unroll
        for (int tx = 0; tx < 4; ++tx) {
          float2 v = acc[ty][tx];
          if (hsw) {  // same operation order as the stand-alone kernel (engine.cu: dw_tile)
            float2 tq = __fadd2_rn(v, make_float2(3.0f, 3.0f));
            tq.x = fminf(fmaxf(tq.x, 0.0f), 6.0f), tq.y = fminf(fmaxf(tq.y, 0.0f), 6.0f);
            v = __fmul2_rn(__fmul2_rn(v, tq), make_float2(0.16666667f, 0.16666667f));
          } else if (P.act != ACT_NONE) {
            v.x = dt_act(v.x, P.act), v.y = dt_act(v.y, P.act);
          }
          if (affine) v.x = v.x * P.ps + P.pb, v.y = v.y * P.ps + P.pb;
          const int oy = oy0 + ty, ox = ox0 + tx;
          if (active && oy < P.Ho && ox < P.Wo && ch < P.C) {
            *reinterpret_cast<float2*>(P.out + (((size_t)img * P.Ho + oy) * P.Wo + ox) * P.C + ch) = v;
            ssum.x += v.x, ssum.y += v.y;
          }
        }
      {
        uint32_t bits = 0;
#pragma unroll
        for (int ty = 0; ty < 2; ++ty)
#pragma unroll
          for (int tx = 0; tx < 4; ++tx) bits |= __float_as_uint(acc[ty][tx].x);
        const uint32_t dep = bits & (uint32_t)(P.n_work >> 31);  // always 0 (n_work >= 0), but not to the compiler
        __syncwarp();
        if (lane == 0) mbar_arrive(DT_BAR(DT_IN_EMPTY + s) + dep);
      }
      if (P.tile_sums) {
        // sum over the tile's pixel groups: lanes l and l ^ 16 hold the same channel pair, then the team's 8 warps
        ssum.x += __shfl_xor_sync(0xffffffffu, ssum.x, 16);
        ssum.y += __shfl_xor_sync(0xffffffffu, ssum.y, 16);
        const int tw8 = (ct >> 5);
        named_bar_sync(1 + team, 256);  // the previous item's partials have been read
        if (lane < 16) red[(team * 8 + tw8) * 16 + lane] = ssum;
        named_bar_sync(1 + team, 256);
        if (tw8 == 0 && lane < 16) {
          float2 t = red[(team * 8) * 16 + lane];
#pragma unroll
          for (int w8 = 1; w8 < 8; ++w8) {
            const float2 u = red[(team * 8 + w8) * 16 + lane];
            t.x += u.x, t.y += u.y;
          }
          const int c2 = kb * 32 + 2 * lane;
          if (c2 < P.C)
            *reinterpret_cast<float2*>(P.tile_sums + ((size_t)img * tiles_per_img + (size_t)th * P.tiles_w + tw) * P.C + c2) = t;
        }
      }
      // next item of this team: two k-blocks on (the ring depth is even: the stage advances by two with one wrap test)
      s += 2;
      if (s >= (uint32_t)P.ns_in) s -= (uint32_t)P.ns_in, ph ^= 1u;
      kb += 2;
      if (kb >= P.nkb) {
        do {
          kb -= P.nkb, sp += gridDim.x;
        } while (kb >= P.nkb);
        tw = sp % P.tiles_w, th = (sp / P.tiles_w) % P.tiles_h, img = sp / P.tiles_w / P.tiles_h;
      }
    }
  }
#undef DT_BAR
}

using DtKern = void (*)(const DtParams, const CUtensorMap);

static DtKern dt_pick(int k, int sh, int sw) {
#define DT_CASE(KV, SHV, SWV) \
  if (k == KV && sh == SHV && sw == SWV) return dw_tma_kernel<KV, SHV, SWV>;
  DT_CASE(3, 1, 1) DT_CASE(5, 1, 1)
  DT_CASE(3, 2, 2) DT_CASE(5, 2, 2)
  DT_CASE(3, 2, 1) DT_CASE(5, 2, 1)
  DT_CASE(3, 1, 2) DT_CASE(5, 1, 2)
#undef DT_CASE
  return nullptr;
}

// false = shape not covered: the caller runs the register-tiled kernel.  tile_sums (optional): [B][*n_tiles][C]
bool tc_dw_tma(oar_model* m, int dw_key, const float* in, float* out, int B, int H, int W, int C, int Ho, int Wo, int k, int sh,
               int sw, int act, float ps, float pb, float* tile_sums, int* n_tiles, const char* name) {
  // Measured on B200 against the register-tiled kernel (rec blocks 6a-6d, 256 crops): strided layers 0.109 vs 0.118 ms,
  // stride-1 layers 0.169 vs 0.153 (6 x 80 maps) and 0.122 vs 0.084 (3 x 80 maps: a 4 x 28 tile is a quarter empty) --
  // the shared-memory form is bound by its LDS wavefronts (~1000 per item and team), not by HBM.  So it takes the
  // strided layers only; OAR_DWTMA=0 / 2 = never / always.
  static const int mode = getenv("OAR_DWTMA") ? atoi(getenv("OAR_DWTMA")) : 1;
  if (mode == 0 || (mode == 1 && sh * sw == 1)) return false;
  TcState* st = static_cast<TcState*>(m->tc_state);
  if (!st) return false;
  auto itd = st->dwp.find(dw_key);
  if (itd == st->dwp.end()) return false;
  DtKern kern = dt_pick(k, sh, sw);
  if (!kern || (C & 3) || (C & 1) || (((uintptr_t)in) & 15) || (((uintptr_t)out) & 7)) return false;
  if (tile_sums && (((uintptr_t)tile_sums) & 7)) return false;
  if (B <= 0 || Ho <= 0 || Wo <= 0) return false;
  DtParams P{};
  P.dw_pk = itd->second, P.act = act, P.ps = ps, P.pb = pb, P.out = out, P.tile_sums = tile_sums;
  P.C = C, P.nkb = (C + 31) / 32, P.H = H, P.W = W, P.Ho = Ho, P.Wo = Wo;
  P.tap_bytes = (uint32_t)(k * k + 1) * 128u;
  // tile: TH even, TW % 4 == 0, <= 128 pixels; an EVEN ring of at least 4 stages (each stage then belongs to one team
  // for good: no cross-team hand-shake); least tiles first, then least halo
  double best = 1e300;
  int bTH = 0, bTW = 0, bns = 0;
  for (int TH = 2; TH <= 32; TH += 2)
    for (int TW = 4; TW <= 64; TW += 4) {
      if (TH * TW > 128) continue;
      const int rows_in = (TH - 1) * sh + k, cols_in = (TW - 1) * sw + k;
      if (rows_in > 256 || cols_in > 256) continue;
      const size_t stage = (size_t)rows_in * cols_in * 128 + P.tap_bytes;
      int ns = (int)std::min<size_t>(DT_MAX_IN, (DT_SMEM_MAX - 4096) / stage) & ~1;
      if (ns < 2) continue;
      const double tiles = (double)cdiv(Ho, TH) * cdiv(Wo, TW);
      const double cost = tiles * (1000.0 + rows_in * cols_in) * (ns == 2 ? 1.3 : 1.0);
      if (cost < best) best = cost, bTH = TH, bTW = TW, bns = ns;
    }
  if (!bTH) return false;
  P.TH = bTH, P.TW = bTW, P.ns_in = bns;
  P.tiles_h = cdiv(Ho, bTH), P.tiles_w = cdiv(Wo, bTW);
  const int rows_in = (bTH - 1) * sh + k;
  P.cols_in = (bTW - 1) * sw + k;
  P.in_bytes = (uint32_t)rows_in * P.cols_in * 128u;
  const long long n_work = (long long)B * P.tiles_h * P.tiles_w;
  if (n_work * P.nkb > 0x7fffffffLL) return false;
  P.n_work = (int)n_work;
  if (n_tiles) *n_tiles = P.tiles_h * P.tiles_w;
  CUtensorMap tm_in;
  memset(&tm_in, 0, sizeof(tm_in));
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 4, (cuuint64_t)C * 4 * W, (cuuint64_t)C * 4 * W * H};
  cuuint32_t box[4] = {32, (cuuint32_t)P.cols_in, (cuuint32_t)rows_in, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  if (tmap_encoder()(&tm_in, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(in), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
    return false;
  const size_t smem = (size_t)P.ns_in * (P.in_bytes + P.tap_bytes) + 128 + 2 * 8 * 16 * sizeof(float2) + 1024;
  ensure_max_dynamic_smem((const void*)kern, m->ctx->device, (int)DT_SMEM_MAX);
  static const bool dbg = getenv("OAR_DBG_TILES") != nullptr;
  if (dbg)
    fprintf(stderr, "[dwtma] k=%d s=%dx%d B=%d %dx%d C=%d -> tile %dx%d stages %d items %d sums %d smem %zu\n", k, sh, sw, B, Ho, Wo,
            C, P.TH, P.TW, P.ns_in, P.n_work, tile_sums != nullptr, smem);
  const int grid = std::min(P.n_work, m->ctx->sm_count);
  Launch l(m->ctx, name, 2.0 * (double)B * Ho * Wo * C * k * k, 4.0 * ((double)B * H * W * C + (double)B * Ho * Wo * C));
  kern<<<grid, DT_THREADS, smem, m->ctx->stream>>>(P, tm_in);
  return true;
}

}  // namespace oar
