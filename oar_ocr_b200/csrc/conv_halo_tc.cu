// conv_halo_tc.cu -- dense stride-1 k x k convolution (k <= 3) as ONE persistent, warp-specialised tcgen05 kernel with a
// TMA-loaded halo tile ("TMA im2col").
//
// In the reference these layers are Conv nodes inside ONNX Runtime (ort_infer_execution.rs:178,281): the 3x3 convs of
// the detector's neck and head, the (1,3) convs of the recogniser's SVTR neck, the 3x3 stacks of HGNetV2 and the
// CSPRep blocks of the layout detector's encoder.  The per-layer kernel they ran on (gemm_tc.cu: conv_rowtaps_tc)
// gathers one image row through registers per CTA, re-reads every input row once per kernel row, synchronises the
// whole CTA per k-block and wastes the rows of a 128-row MMA tile that a narrow image does not fill (a 20-pixel-wide
// map uses 16 % of the tile): 0.17-0.2 of its roofline.  Here:
//
//   tile      TH x TW output pixels of one image, laid out at PITCH P = TW + kw - 1: accumulator row r = y * P + x.
//             The input box [TH + kh - 1] x [P] x [32 ch] arrives by ONE cp.async.bulk.tensor per 32-channel block
//             (zero-filled outside the image = the convolution's padding) and is split to fp16 hi/lo ONCE into the
//             K-major no-swizzle UMMA layout with row = box pixel.  Because consecutive rows of that layout are 16
//             bytes apart, tap (ky, kx) of output row r is box row r + ky * P + kx: the SAME shared-memory tile read
//             through a descriptor that starts (ky * P + kx) * 16 bytes further.  kh * kw taps, one conversion.
//             Rows with x >= TW are computed and dropped (kw - 1 of every P).
//   warps     0-3 epilogue (TMEM -> bias/activation -> dense swizzled staging tile -> TMA store clipped by the map),
//             4 TMA producer (boxes and per-tap weight blocks on two cursors), 5 MMA issuer, 6-13 split/convert.
//   chains    tcgen05.mma into ONE accumulator is a dependent chain: measured ~155 cycles from issue to the next MMA on
//             the same TMEM tile, whatever N is (ncu: tensor pipe 10 % at N = 32, 40 % at N = 128 with one chain).  A
//             work item is therefore a pixel tile with ALL its N tiles, and the taps are dealt round-robin to KS
//             partial accumulators per N tile (KS = 256 columns / N): n_tiles x KS independent chains keep the pipe
//             busy; the epilogue adds the KS partials.  The A tile is converted once for all N tiles.
//   pipeline  2 accumulator sets of <= 256 columns (epilogue of one item under the main loop of the next), 2 A
//             buffers, 2-3 input stages, 3-6 weight stages; mbarriers only.
// Weights: the row-taps packing of gemm_tc.cu ([n tile][ky][cin block][kx][hi|lo][k-chunk][BN rows][8 halfs]); one
// weight stage = TPS taps (a whole kernel row when it fits: small stages are latency-bound, 6 x 4 KB in flight cannot
// cover an L2 round trip) of one 32-channel block for every N tile = TPS * n_tiles * 128 * BN bytes.
#include <cuda.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <cstdlib>

#include "engine.cuh"
#include "tc_ptx.cuh"
#include "tc_state.cuh"

namespace oar {

constexpr int CH_THREADS = 448;  // 4 epilogue + TMA + MMA + 8 convert warps
constexpr int CH_EPI_THREADS = 128;
constexpr int CH_WARP_TMA = 4, CH_WARP_MMA = 5, CH_WARP_C0 = 6;
constexpr int CH_CWARPS = 8, CH_CTHREADS = CH_CWARPS * 32;
constexpr int CH_MAX_IN = 4, CH_MAX_B = 12;
constexpr uint32_t CH_EP_TILE = 128 * 128;
constexpr size_t CH_SMEM_MAX = 227 * 1024;
constexpr uint32_t CH_BIAS_OFF = 384;  // mbarriers + TMEM slot below, the bias behind
constexpr size_t CH_CTRL_BYTES = CH_BIAS_OFF + (512 + 32) * 4;

enum { CB_IN_FULL = 0, CB_IN_EMPTY = 4, CB_B_FULL = 8, CB_B_EMPTY = 20, CB_A_FULL = 32, CB_A_EMPTY = 34, CB_ACC_FULL = 36,
       CB_ACC_EMPTY = 38, CB_NBAR = 40 };

struct ChParams {
  const uint4* wpk;
  const float* bias;
  int act;
  float ps, pb;
  int N, BN, n_tiles, ncb, kh, kw;
  int KS;   // partial accumulators (independent MMA chains) per N tile
  int dbg;  // timing bisect (OAR_DBG_HALO): 1 = no MMAs, 2 = no conversion, 4 = no epilogue body; results are garbage
  int TPS;  // taps per weight stage: kw (one kernel row per bulk copy) when that fits, else 1
  int TH, TW, P, rows_box;  // rows_box = (TH + kh - 1) * P box pixels per channel block
  int tiles_h, tiles_w, n_work;
  int ph, pw;
  uint32_t in_bytes, b_bytes, lbo_a;
  int ns_in, nb, ep_tiles;
  uint32_t off_in, off_a, off_b, off_ctrl;
  // b_resident = 1: the weight ring holds every stage of an item (nb = ncb * taps / TPS) and is filled ONCE per CTA.  A
  // narrow layer (the detector's 96 -> 24 convs: 108 KB of hi/lo weights) otherwise re-streams all of its weights from
  // L2 for every 128-row tile -- more bytes than the tile's activations -- behind nine stage handshakes.
  int b_resident;
  // fold = 1: a narrow 3 x 3 convolution run as a 3 x 1 one with 3 * fold_n output columns (column block kx = the
  // weights of kernel column kx) on a tile whose pitch keeps two halo columns (P = TW + 2, a multiple of a warp or
  // half of one).  D[r][kx * n + co] is then the partial sum of kernel column kx at box pixel r, and the output is
  // out[r][co] = D[r][co] + D[r + 1][n + co] + D[r + 2][2n + co]: rows r + 1, r + 2 live in the neighbouring TMEM lanes =
  // the neighbouring threads of the same warp, so the epilogue folds with two shuffles per channel and stores straight
  // from registers.  A third of the MMAs and of the operand bytes of the tap-by-tap form.
  int fold, fold_n;
  float* out;
  int out_ld, H, W;
};

__device__ __forceinline__ void ch_warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

__device__ __forceinline__ void ch_split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
  const __half2 h = __halves2half2(ha, hb);
  const __half2 l = __halves2half2(__float2half_rn(a - __half2float(ha)), __float2half_rn(b - __half2float(hb)));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__global__ void __launch_bounds__(CH_THREADS, 1) conv_halo_tc(const ChParams P, const __grid_constant__ CUtensorMap tm_in,
                                                              const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ __align__(1024) uint8_t ch_smem_raw[];
  uint8_t* smem = ch_smem_raw + ((1024u - (smem_u32(ch_smem_raw) & 1023u)) & 1023u);
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = sbase + P.off_ctrl;
#define CH_BAR(i) (bar0 + 8u * (uint32_t)(i))
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P.off_ctrl + 8 * CB_NBAR);
  const uint32_t a_part = 4u * P.lbo_a, a_buf = 2u * a_part;
  const int taps = P.kh * P.kw;

  if (tid == 0) {
    for (int i = 0; i < CH_MAX_IN; ++i) {
      mbar_init(CH_BAR(CB_IN_FULL + i), 1);
      mbar_init(CH_BAR(CB_IN_EMPTY + i), CH_CWARPS);
    }
    for (int i = 0; i < CH_MAX_B; ++i) {
      mbar_init(CH_BAR(CB_B_FULL + i), 1);
      mbar_init(CH_BAR(CB_B_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(CH_BAR(CB_A_FULL + i), CH_CWARPS);
      mbar_init(CH_BAR(CB_A_EMPTY + i), 1);
      mbar_init(CH_BAR(CB_ACC_FULL + i), 1);
      mbar_init(CH_BAR(CB_ACC_EMPTY + i), CH_EPI_THREADS / 32);
    }
    fence_mbar_init();
  }
  if (warp == CH_WARP_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t n_items = P.n_work > (int)blockIdx.x ? (uint32_t)((P.n_work - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) : 0u;

  if (warp == CH_WARP_TMA) {
    // ------------------------------------------------------------------ producer: boxes and weight taps, two cursors;
    // the whole warp runs the loop, one elected lane issues the copies
    {
      const uint32_t n_in = n_items * (uint32_t)P.ncb, n_b = n_in * (uint32_t)taps;
      const uint32_t one = P.b_bytes / (uint32_t)(P.n_tiles * P.TPS);  // one N tile's block of one tap
      const uint32_t n_st = (uint32_t)(taps / P.TPS);                  // weight stages per channel block
      const uint32_t n_b2 = P.b_resident ? (n_items ? (uint32_t)P.ncb * n_st : 0u) : n_in * n_st;
      const size_t nt_stride = (size_t)P.kh * P.ncb * P.kw * one;  // row-taps packing: [n tile][ky][cin block][kx] blocks
      uint32_t it_in = 0, it_b = 0;
      uint32_t s_in = 0, ph_in = 0, s_b = 0, ph_b = 0;
      int t_in = blockIdx.x, cb_in = 0;
      int tw_in = t_in % P.tiles_w, r_in = t_in / P.tiles_w;
      int cb_b = 0, tap_b = 0;  // tap_b = first tap of the stage (= ky * kw + kx)
      (void)n_b;
      while (it_in < n_in || it_b < n_b2) {
        if (it_in < n_in && mbar_test_warp(CH_BAR(CB_IN_EMPTY + s_in), ph_in ^ 1u)) {
          if (elect_one_sync()) {
            mbar_expect_tx(CH_BAR(CB_IN_FULL + s_in), P.in_bytes);
            tma_load_4d(sbase + P.off_in + s_in * P.in_bytes, &tm_in, CH_BAR(CB_IN_FULL + s_in), cb_in * 32,
                        tw_in * P.TW - P.pw, (r_in % P.tiles_h) * P.TH - P.ph, r_in / P.tiles_h);
          }
          __syncwarp();
          ++it_in;
          if (++s_in == (uint32_t)P.ns_in) s_in = 0, ph_in ^= 1u;
          if (++cb_in == P.ncb) {
            cb_in = 0, t_in += gridDim.x;
            tw_in = t_in % P.tiles_w, r_in = t_in / P.tiles_w;
          }
        }
        if (it_b < n_b2 && mbar_test_warp(CH_BAR(CB_B_EMPTY + s_b), ph_b ^ 1u)) {
          // taps tap_b .. tap_b + TPS - 1 of block cb_b are contiguous in the packing (same ky when TPS = kw)
          const int ky = tap_b / P.kw, kx = tap_b - ky * P.kw;
          const size_t blk0 = (((size_t)ky * P.ncb + cb_b) * P.kw + kx) * one;
          if (elect_one_sync()) {
            mbar_expect_tx(CH_BAR(CB_B_FULL + s_b), P.b_bytes);
            for (int nt = 0; nt < P.n_tiles; ++nt)
              bulk_load(sbase + P.off_b + s_b * P.b_bytes + nt * (P.TPS * one),
                        reinterpret_cast<const uint8_t*>(P.wpk) + nt * nt_stride + blk0, P.TPS * one, CH_BAR(CB_B_FULL + s_b));
          }
          __syncwarp();
          ++it_b;
          if (++s_b == (uint32_t)P.nb) s_b = 0, ph_b ^= 1u;
          tap_b += P.TPS;
          if (tap_b == taps) {
            tap_b = 0;
            if (++cb_b == P.ncb) cb_b = 0;
          }
        }
      }
    }
  } else if (warp == CH_WARP_MMA) {
    // ------------------------------------------------------------------ MMA issuer: the whole warp runs the loop (its
    // values are warp-uniform and stay in uniform registers), one elected lane issues the tcgen05 instructions
    {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(P.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)P.BN * 16u, b_part = 4u * b_lbo;
      uint32_t it_a = 0, ti = 0;
      uint32_t sb = 0, phb = 0;  // weight ring cursor (no divisions anywhere in this loop: the issue rate of this one
                                 // warp bounds the kernel -- a tcgen05.mma retires every ~45-65 cycles at N <= 128,
                                 // tools/microbench/mma_issue_bench.cu, and two integer modulos per MMA cost 270)
      const uint32_t one16 = (P.b_bytes / (uint32_t)(P.n_tiles * P.TPS)) >> 4;  // one tap of one N tile, in 16-byte units
      const uint32_t nt16 = one16 * (uint32_t)P.TPS;                              // one N tile's part of a stage
      const uint32_t ks_cols = (uint32_t)(P.KS * P.BN), nt_cols = ks_cols;  // TMEM columns of one N tile's partials
      // descriptor = constant upper half | (address >> 4) in the low 14 bits: one add per operand and MMA
      const uint64_t adesc0 = make_desc(0, P.lbo_a, 128), bdesc0 = make_desc(0, b_lbo, 128);
      const uint32_t a_j = P.lbo_a >> 3, b_j = b_lbo >> 3, b_lo_off = b_part >> 4;
      if (P.fold) {
        // Fold mode: everything about a tile's 54 MMAs is static -- three kernel rows x two k16 steps x (hi*hi, hi*lo,
        // lo*hi) per channel block, weights resident -- so the loop below is three waits and three unrolled issue
        // blocks per tile.  (The general loop spends ~150 instructions of uniform-register bookkeeping per tap: ncu put
        // the MMA warp at ~1100 cycles per tap on this layer, the convert warps waiting on it 63 % of the time.)
        for (int s2 = 0; s2 < P.nb; ++s2) mbar_wait_warp(CH_BAR(CB_B_FULL + s2), 0u);  // resident weights: once
        tc_fence_after();
        const uint32_t b16 = (sbase + P.off_b) >> 4, stage16 = P.b_bytes >> 4, pitch = (uint32_t)P.P;
        const uint32_t ks = (uint32_t)P.KS, bn = (uint32_t)P.BN;
        for (int t = blockIdx.x; t < P.n_work; t += gridDim.x, ++ti) {
          const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
          mbar_wait_warp(CH_BAR(CB_ACC_EMPTY + acc), aph ^ 1u);
          const uint32_t d0 = tmem_base + acc * 256u;
          for (int cb = 0; cb < P.ncb; ++cb, ++it_a) {
            const uint32_t sa = it_a & 1u;
            mbar_wait_warp(CH_BAR(CB_A_FULL + sa), (it_a >> 1) & 1u);
            tc_fence_after();
            const uint32_t a_hi = (sbase + P.off_a + sa * a_buf) >> 4, a_lo = a_hi + (a_part >> 4);
            const uint32_t bcb = b16 + (uint32_t)(cb * 3) * stage16;
            const bool last_cb = cb == P.ncb - 1;
            if (elect_one_sync()) {
#pragma unroll
              for (int ky = 0; ky < 3; ++ky) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  const uint32_t g = (uint32_t)(ky * 2 + j);  // group index within the channel block
                  const uint32_t chain = g % ks;
                  const uint32_t accum = (cb == 0 && g < ks) ? 0u : 1u;  // a chain's first group starts it
                  const uint32_t shift = (uint32_t)ky * pitch;          // kernel row ky: box rows r + ky * P
                  const uint64_t ah = adesc0 + (uint64_t)(a_hi + j * a_j + shift);
                  const uint64_t al = adesc0 + (uint64_t)(a_lo + j * a_j + shift);
                  const uint32_t bb = bcb + (uint32_t)ky * stage16 + j * b_j;
                  const uint64_t bh = bdesc0 + (uint64_t)bb, bl = bdesc0 + (uint64_t)(bb + b_lo_off);
                  const uint32_t d = d0 + chain * bn;
                  umma_f16(d, ah, bh, idesc, accum);
                  umma_f16(d, ah, bl, idesc, 1u);
                  umma_f16(d, al, bh, idesc, 1u);
                }
              }
              umma_commit(CH_BAR(CB_A_EMPTY + sa));
              if (last_cb) umma_commit(CH_BAR(CB_ACC_FULL + acc));
            }
            __syncwarp();
          }
        }
      } else
      for (int t = blockIdx.x; t < P.n_work; t += gridDim.x, ++ti) {
        const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
        mbar_wait_warp(CH_BAR(CB_ACC_EMPTY + acc), aph ^ 1u);
        tc_fence_after();
        const uint32_t d0 = tmem_base + acc * 256u;
        uint32_t chain_col = 0;             // column offset of the chain the next group accumulates into
        uint32_t fresh_left = (uint32_t)P.KS;  // groups that still start their chain (accumulate = 0)
        for (int cb = 0; cb < P.ncb; ++cb, ++it_a) {
          const uint32_t sa = it_a & 1u;
          mbar_wait_warp(CH_BAR(CB_A_FULL + sa), (it_a >> 1) & 1u);
          const uint32_t a_hi = (sbase + P.off_a + sa * a_buf) >> 4, a_lo = a_hi + (a_part >> 4);
          uint32_t row_shift = 0, kx = 0;  // the tap's rows start this many 16-byte rows into the tile
          for (int tap0 = 0; tap0 < taps; tap0 += P.TPS) {
            mbar_wait_warp(CH_BAR(CB_B_FULL + sb), P.b_resident ? 0u : phb);  // resident: filled once, phase 0 stays complete
            tc_fence_after();
            const uint32_t bst = (sbase + P.off_b + sb * P.b_bytes) >> 4;
            // One elected issue block per weight stage (TPS <= 3 taps): every lane does the cheap bookkeeping -- chains,
            // accumulate flags, row shifts -- into registers first, then one lane issues all of the stage's MMAs and
            // commits.  (An elect + warp barrier per tap left the MMA warp at ~1100 cycles per tap in ncu.)
            uint32_t col[3][2], accum[3][2], shift[3];
            const bool last_cb = cb == P.ncb - 1;
#pragma unroll
            for (int tt = 0; tt < 3; ++tt) {
              if (tt < P.TPS) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                  col[tt][j] = chain_col, accum[tt][j] = fresh_left ? 0u : 1u;
                  if (fresh_left) --fresh_left;
                  chain_col += (uint32_t)P.BN;
                  if (chain_col == ks_cols) chain_col = 0;
                }
                shift[tt] = row_shift;
                ++row_shift;
                if (++kx == (uint32_t)P.kw) kx = 0, row_shift += (uint32_t)(P.P - P.kw);
              }
            }
            const bool last_stage = tap0 + P.TPS == taps;
            if (elect_one_sync()) {
#pragma unroll
              for (int tt = 0; tt < 3; ++tt) {
                if (tt < P.TPS) {
                  const uint32_t b0 = bst + (uint32_t)tt * one16;
#pragma unroll
                  for (int j = 0; j < 2; ++j) {
                    const uint64_t ah = adesc0 + (uint64_t)(a_hi + j * a_j + shift[tt]);
                    const uint64_t al = adesc0 + (uint64_t)(a_lo + j * a_j + shift[tt]);
                    uint32_t bb = b0 + j * b_j, d = d0 + col[tt][j];
                    for (int nt = 0; nt < P.n_tiles && !(P.dbg & 1); ++nt, bb += nt16, d += nt_cols) {
                      const uint64_t bh = bdesc0 + (uint64_t)bb, bl = bdesc0 + (uint64_t)(bb + b_lo_off);
                      umma_f16(d, ah, bh, idesc, accum[tt][j]);
                      umma_f16(d, ah, bl, idesc, 1u);
                      umma_f16(d, al, bh, idesc, 1u);
                    }
                  }
                }
              }
              if (!P.b_resident) umma_commit(CH_BAR(CB_B_EMPTY + sb));
              if (last_stage) {
                umma_commit(CH_BAR(CB_A_EMPTY + sa));
                if (last_cb) umma_commit(CH_BAR(CB_ACC_FULL + acc));
              }
            }
            __syncwarp();
            if (++sb == (uint32_t)P.nb) sb = 0, phb ^= 1u;
          }
        }
      }
    }
  } else if (warp >= CH_WARP_C0) {
    // ------------------------------------------------------------------ split / convert: fp32 box -> fp16 hi/lo A tile
    const int ct = tid - CH_WARP_C0 * 32;
    const int q = ct & 7;        // 16-byte piece (4 channels) of a pixel's 128 bytes
    const int r0 = ct >> 3;      // box pixel, stride 32
    const uint32_t a_off = (uint32_t)(q >> 1) * P.lbo_a + (uint32_t)(q & 1) * 8u;
    const uint32_t n_in = n_items * (uint32_t)P.ncb;
    for (uint32_t it = 0; it < n_in; ++it) {
      const uint32_t s = it % (uint32_t)P.ns_in, ph = (it / (uint32_t)P.ns_in) & 1u;
      const uint32_t sa = it & 1u;
      mbar_wait(CH_BAR(CB_IN_FULL + s), ph);
      mbar_wait(CH_BAR(CB_A_EMPTY + sa), ((it >> 1) & 1u) ^ 1u);
      const uint8_t* src = smem + P.off_in + s * P.in_bytes + q * 16;
      uint8_t* ab = smem + P.off_a + sa * a_buf + a_off;
      for (int r = r0; r < P.rows_box && !(P.dbg & 2); r += 32) {
        const float4 x = *reinterpret_cast<const float4*>(src + (size_t)r * 128);
        uint32_t h0, l0, h1, l1;
        ch_split2(x.x, x.y, h0, l0);
        ch_split2(x.z, x.w, h1, l1);
        *reinterpret_cast<uint2*>(ab + (size_t)r * 16) = make_uint2(h0, h1);
        *reinterpret_cast<uint2*>(ab + a_part + (size_t)r * 16) = make_uint2(l0, l1);
      }
      ch_warp_arrive(CH_BAR(CB_IN_EMPTY + s), lane);
      fence_proxy_async_smem();
      ch_warp_arrive(CH_BAR(CB_A_FULL + sa), lane);
    }
  } else {
    // ------------------------------------------------------------------ epilogue (tid = accumulator row = TMEM lane)
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool affine = P.ps != 1.0f || P.pb != 0.0f;
    float* bias_s = reinterpret_cast<float*>(smem + P.off_ctrl + CH_BIAS_OFF);
    const int n_bias = P.fold ? P.fold_n : P.N;  // fold: N counts the folded columns, the bias has fold_n entries
    for (int i = tid; i < P.n_tiles * P.BN + 32; i += CH_EPI_THREADS) bias_s[i] = i < n_bias ? __ldg(P.bias + i) : 0.0f;
    named_bar_sync(1, CH_EPI_THREADS);
    const int y = tid / P.P, x = tid - y * P.P;
    const bool valid = y < P.TH && x < P.TW;
    const int dense = y * P.TW + x;  // row of the dense [TH][TW] staging tile
    uint32_t ti = 0, nstore = 0;
    for (int t = blockIdx.x; t < P.n_work; t += gridDim.x, ++ti) {
      const int sp = t;
      const int tw = sp % P.tiles_w, r = sp / P.tiles_w;
      const int c1 = tw * P.TW, c2 = (r % P.tiles_h) * P.TH, c3 = r / P.tiles_h;
      const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
      mbar_wait(CH_BAR(CB_ACC_FULL + acc), aph);
      tc_fence_after();
      if (P.fold) {
        constexpr int FN = 24;  // fold_n (checked by the host)
        float v[80];
        __syncwarp();
#pragma unroll
        for (int c16 = 0; c16 < 5; ++c16) tmem_ld16(lane_base + acc * 256u + (uint32_t)(c16 * 16), v + 16 * c16);
        for (int ch = 1; ch < P.KS; ++ch) {  // the other partial accumulators (independent MMA chains)
#pragma unroll
          for (int c16 = 0; c16 < 5; ++c16) {
            float u[16];
            tmem_ld16(lane_base + acc * 256u + (uint32_t)(ch * P.BN + c16 * 16), u);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[16 * c16 + i] += u[i];
          }
        }
        float o[FN];
#pragma unroll
        for (int co = 0; co < FN; ++co)
          o[co] = v[co] + __shfl_down_sync(0xffffffffu, v[FN + co], 1) + __shfl_down_sync(0xffffffffu, v[2 * FN + co], 2);
        const int oy = c2 + y, ox = c1 + x;
        if (valid && oy < P.H && ox < P.W) {
#pragma unroll
          for (int co = 0; co < FN; ++co) o[co] += bias_s[co];
          switch (P.act) {
#define CH_ACT_CASE_F(A)                                     \
  case A:                                                    \
    _Pragma("unroll") for (int i = 0; i < FN; ++i) o[i] = act_t<A>(o[i]); \
    break;
            CH_ACT_CASE_F(ACT_RELU)
            CH_ACT_CASE_F(ACT_HSWISH)
            CH_ACT_CASE_F(ACT_SWISH)
            CH_ACT_CASE_F(ACT_SIGMOID)
            CH_ACT_CASE_F(ACT_HSIGMOID)
            CH_ACT_CASE_F(ACT_GELU)
#undef CH_ACT_CASE_F
            default: break;
          }
          if (affine) {
#pragma unroll
            for (int i = 0; i < FN; ++i) o[i] = o[i] * P.ps + P.pb;
          }
          float4* dst = reinterpret_cast<float4*>(P.out + (((size_t)c3 * P.H + oy) * P.W + ox) * P.out_ld);
#pragma unroll
          for (int i = 0; i < FN / 4; ++i) dst[i] = make_float4(o[4 * i], o[4 * i + 1], o[4 * i + 2], o[4 * i + 3]);
        }
        tc_fence_before();
        ch_warp_arrive(CH_BAR(CB_ACC_EMPTY + acc), lane);
        continue;
      }
      for (int nt = 0; nt < P.n_tiles && !(P.dbg & 4); ++nt) {
        const int n_base = nt * P.BN;
        const uint32_t col0 = lane_base + acc * 256u + (uint32_t)(nt * P.KS * P.BN);
        for (int c0 = 0; c0 < P.BN && n_base + c0 < P.N; c0 += 32, ++nstore) {
          float v[32];
          const bool two = c0 + 16 < P.BN;
          __syncwarp();
          tmem_ld16(col0 + (uint32_t)c0, v);
          if (two) {
            tmem_ld16(col0 + (uint32_t)c0 + 16u, v + 16);
          } else {
#pragma unroll
            for (int i = 16; i < 32; ++i) v[i] = 0.0f;
          }
          for (int ch = 1; ch < P.KS; ++ch) {  // the other partial accumulators of this N tile
            float u[16];
            tmem_ld16(col0 + (uint32_t)(ch * P.BN + c0), u);
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] += u[i];
            if (two) {
              tmem_ld16(col0 + (uint32_t)(ch * P.BN + c0) + 16u, u);
#pragma unroll
              for (int i = 0; i < 16; ++i) v[16 + i] += u[i];
            }
          }
          {
            const float4* bs = reinterpret_cast<const float4*>(bias_s + n_base + c0);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 b = bs[i];
              v[4 * i] += b.x, v[4 * i + 1] += b.y, v[4 * i + 2] += b.z, v[4 * i + 3] += b.w;
            }
            switch (P.act) {
#define CH_ACT_CASE(A)                                      \
  case A:                                                   \
    _Pragma("unroll") for (int i = 0; i < 32; ++i) v[i] = act_t<A>(v[i]); \
    break;
              CH_ACT_CASE(ACT_RELU)
              CH_ACT_CASE(ACT_HSWISH)
              CH_ACT_CASE(ACT_SWISH)
              CH_ACT_CASE(ACT_SIGMOID)
              CH_ACT_CASE(ACT_HSIGMOID)
              CH_ACT_CASE(ACT_GELU)
#undef CH_ACT_CASE
              default: break;
            }
          }
          if (affine) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = v[i] * P.ps + P.pb;
          }
          const uint32_t slot = P.ep_tiles == 2 ? (nstore & 1u) : 0u;
          if (P.ep_tiles == 1) {
            if (tid == 0) bulk_wait_read_all();
            named_bar_sync(1, CH_EPI_THREADS);
          }
          if (valid) {
            uint8_t* ep = smem + slot * CH_EP_TILE + dense * 128;
#pragma unroll
            for (int i = 0; i < 8; ++i)
              *reinterpret_cast<float4*>(ep + ((i ^ (dense & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          }
          fence_proxy_async_smem();
          if (P.ep_tiles == 2 && tid == 0) bulk_wait_read_all();
          named_bar_sync(1, CH_EPI_THREADS);
          if (tid == 0) {
            tma_store_4d(&tm_out, sbase + slot * CH_EP_TILE, n_base + c0, c1, c2, c3);
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      ch_warp_arrive(CH_BAR(CB_ACC_EMPTY + acc), lane);
    }
    if (tid == 0) bulk_wait_all();
  }
#undef CH_BAR
  tc_fence_before();
  __syncthreads();
  if (warp == CH_WARP_MMA) tmem_dealloc(tmem_base, 512);
}

static bool ch_encode_map(CUtensorMap* tm, const float* base, const cuuint64_t* dims, const cuuint64_t* strides,
                          const cuuint32_t* box, CUtensorMapSwizzle swz) {
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = tmap_encoder()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

// false = shape not covered: the caller runs the per-layer kernel
bool tc_conv_halo(oar_model* m, int key, const ConvParams& p, const char* name) {
  static const bool off = getenv("OAR_DBG_NOHALO") != nullptr;  // A/B switch
  if (off) return false;
  TcState* st = static_cast<TcState*>(m->tc_state);
  if (!st) return false;
  auto it = st->w.find(key);
  if (it == st->w.end()) return false;
  const TcWeights& w = it->second;
  if (!w.rowtaps || w.KC != 4 || w.kh != p.kh || w.kw != p.kw || w.N != p.N) return false;
  if (p.mode != 0 || p.sh != 1 || p.sw != 1 || p.kh > 3 || p.kw > 3 || p.ph != p.kh / 2 || p.pw != p.kw / 2) return false;
  if ((p.Cin & 31) || p.Ho != p.H || p.Wo != p.W || (p.out_ld & 3) || (p.out_c_off & 3)) return false;
  if ((((uintptr_t)p.in) & 15) || (((uintptr_t)p.out) & 15) || (w.BN & 15) || w.BN > 256) return false;
  if (w.n_tiles * w.BN > 256) return false;  // one accumulator set per item, two sets in TMEM
  // Narrow, shallow convolutions (the detector's 96 -> 24 neck / head convs: 162 small MMAs per tile) are bound by this
  // kernel's per-stage handshakes, not by its MMAs (timing bisect, OAR_DBG_HALO: 0.53 ms of 0.94 ms per launch remain
  // with MMAs, conversion and epilogue all switched off); three co-resident row-taps CTAs per SM do better there.
  static const bool force = getenv("OAR_DBG_HALO_ALL") != nullptr;
  // Resident weights (b_resident, round 2) remove the nine weight-stage handshakes and the L2 re-stream, and it is
  // still slower there: 0.98 ms per launch against 0.63 for conv_rowtaps_tc on the 240 x 240 maps (18.5 k cycles per
  // 120-pixel tile against an MMA issue floor of 7.3 k).  So narrow layers stay on the row-taps kernel unless
  // OAR_HALO_NARROW=1.  (tools/microbench/tma_row_bench.cu rules the TMA unit out: 128-byte box rows stream at
  // 6.1-6.8 TB/s once 60 KB are in flight per SM.)
  static const bool no_narrow = !(getenv("OAR_HALO_NARROW") && atoi(getenv("OAR_HALO_NARROW")) == 1);
  const bool narrow = w.n_tiles * w.BN < 64;
  if (narrow && no_narrow && !force) return false;
  if ((size_t)w.n_tiles * w.BN + 32 > (CH_CTRL_BYTES - 256) / sizeof(float)) return false;
  if (p.M <= 0) return true;
  // a kernel of height 1 never mixes rows: all images stack into one tall image and tiles span several of them
  const int H = p.kh == 1 ? p.B * p.H : p.H, B = p.kh == 1 ? 1 : p.B, W = p.W;
  // tile: TW columns at pitch P = TW + kw - 1, TH rows with TH * P - (kw - 1) <= 128: least (MMA rows + staged box rows)
  int bTH = 0, bTW = 0;
  double best_cost = -1;
  for (int TW = 1; TW <= std::min(W, 128 - (p.kw - 1)); ++TW) {
    const int Pp = TW + p.kw - 1;
    int TH = std::min((128 + p.kw - 1) / Pp, H);
    if (TH < 1) continue;
    if ((TH + p.kh - 1) > 256 || Pp > 256) continue;
    const double tiles = (double)cdiv(H, TH) * cdiv(W, TW);
    const double cost = tiles * (128.0 * p.kh * p.kw + 0.5 * (TH + p.kh - 1) * Pp + 48.0);
    if (best_cost < 0 || cost < best_cost) best_cost = cost, bTH = TH, bTW = TW;
  }
  if (!bTH) return false;
  ChParams P{};
  P.wpk = w.packed, P.bias = p.bias, P.act = p.act, P.ps = p.post_scale, P.pb = p.post_bias;
  P.N = p.N, P.BN = w.BN, P.n_tiles = w.n_tiles, P.ncb = p.Cin / 32, P.kh = p.kh, P.kw = p.kw;
  P.TH = bTH, P.TW = bTW, P.P = bTW + p.kw - 1, P.rows_box = (bTH + p.kh - 1) * P.P;
  P.tiles_h = cdiv(H, bTH), P.tiles_w = cdiv(W, bTW);
  P.ph = p.ph, P.pw = p.pw;
  P.in_bytes = (uint32_t)P.rows_box * 128u;
  static const int dbg_mode = getenv("OAR_DBG_HALO") ? atoi(getenv("OAR_DBG_HALO")) : 0;
  P.dbg = dbg_mode;
  P.TPS = (size_t)p.kw * 128 * w.BN * w.n_tiles <= 40 * 1024 ? p.kw : 1;
  P.b_bytes = 128u * (uint32_t)w.BN * (uint32_t)w.n_tiles * (uint32_t)P.TPS;
  // K-split chains: fill the 256-column accumulator set (BN % 32 keeps every partial on a 32-column boundary)
  P.KS = (w.BN & 31) ? 1 : std::max(1, std::min(8, 256 / (w.n_tiles * w.BN)));
  P.KS = std::min(P.KS, P.ncb * p.kh * p.kw * 2);
  // chunk stride of the A tile: every box row plus the last tap's overhang, + 16 B so the 8-byte split stores of a
  // quarter warp spread over the banks
  // (and the 128 rows an MMA always reads from the last tap's start, so that no read leaves the tile's own region)
  // rounded to 128 B + 32: the four k-chunks of a pixel then land 8 banks apart and the 8-byte split stores of a
  // half-warp (2 pixels x 4 chunks x hi pair) cover all 32 banks once
  P.lbo_a = (((uint32_t)std::max(P.rows_box + p.kw, 128 + (p.kh - 1) * P.P + p.kw) * 16u + 127u) & ~127u) + 32u;
  const long long n_work = (long long)B * P.tiles_h * P.tiles_w;
  if (n_work > 0x7fffffffLL) return false;
  P.n_work = (int)n_work;
  const size_t a_bytes = 2 * 2 * 4 * (size_t)P.lbo_a;
  // resident weights when every stage of an item fits beside two input stages
  static const bool no_resident = getenv("OAR_DBG_HALO_STREAM") != nullptr;  // A/B switch
  const int nb_res = P.ncb * (p.kh * p.kw / P.TPS);
  P.b_resident = 0;
  if (!no_resident && nb_res <= CH_MAX_B) {
    for (P.ep_tiles = 2; P.ep_tiles >= 1 && !P.b_resident; --P.ep_tiles)
      for (P.ns_in = 3; P.ns_in >= 2; --P.ns_in) {
        const size_t need = (size_t)P.ep_tiles * CH_EP_TILE + a_bytes + CH_CTRL_BYTES + 1024 + 512 +
                            (size_t)P.ns_in * P.in_bytes + (size_t)nb_res * P.b_bytes;
        if (need <= CH_SMEM_MAX) {
          P.b_resident = 1, P.nb = nb_res;
          break;
        }
      }
    if (P.b_resident) ++P.ep_tiles;  // the loop header stepped past the accepted value
  }
  if (!P.b_resident)
  for (P.ep_tiles = 2; P.ep_tiles >= 1; --P.ep_tiles) {
    const size_t fixed = (size_t)P.ep_tiles * CH_EP_TILE + a_bytes + CH_CTRL_BYTES + 1024;
    for (P.ns_in = 3; P.ns_in >= 2; --P.ns_in) {
      const size_t left = CH_SMEM_MAX > fixed + (size_t)P.ns_in * P.in_bytes ? CH_SMEM_MAX - fixed - (size_t)P.ns_in * P.in_bytes : 0;
      P.nb = (int)std::min<size_t>(CH_MAX_B, left / P.b_bytes);
      if (P.nb >= 3) break;
    }
    if (P.nb >= 3) break;
  }
  if (!P.b_resident && (P.nb < 3 || P.ns_in < 2)) return false;
  if (narrow && !P.b_resident && !force) return false;
  P.off_in = (uint32_t)P.ep_tiles * CH_EP_TILE;
  P.off_a = P.off_in + (uint32_t)P.ns_in * P.in_bytes;
  P.off_a = (P.off_a + 127u) & ~127u;
  P.off_b = P.off_a + (uint32_t)a_bytes;
  P.off_b = (P.off_b + 127u) & ~127u;
  P.off_ctrl = P.off_b + (uint32_t)P.nb * P.b_bytes;
  P.off_ctrl = (P.off_ctrl + 127u) & ~127u;
  if ((size_t)P.off_ctrl + CH_CTRL_BYTES + 1024 > CH_SMEM_MAX) return false;
  CUtensorMap tm_in, tm_out;
  memset(&tm_in, 0, sizeof(tm_in));
  memset(&tm_out, 0, sizeof(tm_out));
  cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Cin * 4 * W, (cuuint64_t)p.Cin * 4 * W * H};
  cuuint32_t box[4] = {32, (cuuint32_t)P.P, (cuuint32_t)(bTH + p.kh - 1), 1};
  if (!ch_encode_map(&tm_in, p.in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return false;
  cuuint64_t odims[4] = {(cuuint64_t)p.N, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t ostrides[3] = {(cuuint64_t)p.out_ld * 4, (cuuint64_t)p.out_ld * 4 * W, (cuuint64_t)p.out_ld * 4 * W * H};
  cuuint32_t obox[4] = {32, (cuuint32_t)bTW, (cuuint32_t)bTH, 1};
  if (!ch_encode_map(&tm_out, p.out + p.out_c_off, odims, ostrides, obox, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
  const size_t smem = std::max<size_t>((size_t)P.off_ctrl + CH_CTRL_BYTES + 1024, 120 * 1024);  // one CTA per SM: 512 TMEM columns
  ensure_max_dynamic_smem((const void*)conv_halo_tc, m->ctx->device, (int)CH_SMEM_MAX);
  static const bool dbg = getenv("OAR_DBG_TILES") != nullptr;
  if (dbg)
    fprintf(stderr, "[halo] k=%dx%d B=%d %dx%d C=%d N=%d (%d x %d) chains %d -> tile %dx%d pitch %d stages in %d w %d x %d taps resident %d staging %d items %d smem %zu\n",
            p.kh, p.kw, p.B, p.H, p.W, p.Cin, p.N, w.n_tiles, w.BN, P.KS, P.TH, P.TW, P.P, P.ns_in, P.nb, P.TPS, P.b_resident, P.ep_tiles, P.n_work, smem);
  const int grid = std::min(P.n_work, m->ctx->sm_count);
  Launch l(m->ctx, name, 2.0 * p.M * p.N * p.K, 4.0 * ((double)p.M * p.Cin + (double)p.M * p.N));
  conv_halo_tc<<<grid, CH_THREADS, smem, m->ctx->stream>>>(P, tm_in, tm_out);
  return true;
}

// Narrow 3 x 3 convolutions (Cout <= 32: the detector's 96 -> 24 neck / head convs) with the kernel columns folded
// into N: see ChParams::fold.  false = shape not covered.
bool tc_conv_fold(oar_model* m, int key, const ConvParams& p, const char* name) {
  static const bool off = getenv("OAR_DBG_NOFOLD") != nullptr;  // A/B switch
  if (off) return false;
  TcState* st = static_cast<TcState*>(m->tc_state);
  if (!st) return false;
  auto it = st->wfold.find(key);
  if (it == st->wfold.end()) return false;
  const TcWeights& w = it->second;
  if (p.mode != 0 || p.kh != 3 || p.kw != 3 || p.sh != 1 || p.sw != 1 || p.ph != 1 || p.pw != 1) return false;
  if (p.N != 24 || w.N != 3 * p.N || w.n_tiles != 1 || w.BN != 80 || !w.rowtaps || w.kh != 3 || w.kw != 1) return false;
  if ((p.Cin & 31) || p.Ho != p.H || p.Wo != p.W || (p.out_ld & 3) || (p.out_c_off & 3)) return false;
  if ((((uintptr_t)p.in) & 15) || (((uintptr_t)(p.out + p.out_c_off)) & 15)) return false;
  if (p.M <= 0) return true;
  ChParams P{};
  P.wpk = w.packed, P.bias = p.bias, P.act = p.act, P.ps = p.post_scale, P.pb = p.post_bias;
  P.N = w.N, P.BN = w.BN, P.n_tiles = 1, P.ncb = p.Cin / 32, P.kh = 3, P.kw = 1;  // the MMA loop sees a 3 x 1 conv
  P.fold = 1, P.fold_n = p.N, P.out = p.out + p.out_c_off, P.out_ld = p.out_ld, P.H = p.H, P.W = p.W;
  // pitch 32 (one tile row per warp) or 16 (two): the two dropped columns of a row sit at a warp's / half-warp's end, so
  // the shuffles never cross a row; the fewer tiles win
  const long long t32 = (long long)cdiv(p.H, 4) * cdiv(p.W, 30), t16 = (long long)cdiv(p.H, 8) * cdiv(p.W, 14);
  P.P = t32 <= t16 ? 32 : 16;
  P.TW = P.P - 2, P.TH = 128 / P.P;
  P.rows_box = (P.TH + 2) * P.P;
  P.tiles_h = cdiv(p.H, P.TH), P.tiles_w = cdiv(p.W, P.TW);
  P.ph = 1, P.pw = 1;
  P.in_bytes = (uint32_t)P.rows_box * 128u;
  // one accumulator per set: two partial accumulators (independent MMA chains, BN padded to 96) measured the same --
  // tcgen05.mma into one accumulator is not a dependent chain (tools/microbench/mma_issue_bench.cu)
  P.TPS = 1, P.KS = 1, P.dbg = 0;
  P.b_bytes = 128u * (uint32_t)w.BN;
  P.lbo_a = (((uint32_t)std::max(P.rows_box + 1, 128 + 2 * P.P + 1) * 16u + 127u) & ~127u) + 32u;
  const long long n_work = (long long)p.B * P.tiles_h * P.tiles_w;
  if (n_work > 0x7fffffffLL) return false;
  P.n_work = (int)n_work;
  const size_t a_bytes = 2 * 2 * 4 * (size_t)P.lbo_a;
  P.b_resident = 1, P.nb = P.ncb * 3, P.ep_tiles = 0;
  if (P.nb > CH_MAX_B) return false;
  for (P.ns_in = 3; P.ns_in >= 2; --P.ns_in)
    if (a_bytes + CH_CTRL_BYTES + 1024 + 512 + (size_t)P.ns_in * P.in_bytes + (size_t)P.nb * P.b_bytes <= CH_SMEM_MAX) break;
  if (P.ns_in < 2) return false;
  P.off_in = 0;
  P.off_a = (P.off_in + (uint32_t)P.ns_in * P.in_bytes + 127u) & ~127u;
  P.off_b = (P.off_a + (uint32_t)a_bytes + 127u) & ~127u;
  P.off_ctrl = (P.off_b + (uint32_t)P.nb * P.b_bytes + 127u) & ~127u;
  if ((size_t)P.off_ctrl + CH_CTRL_BYTES + 1024 > CH_SMEM_MAX) return false;
  CUtensorMap tm_in, tm_out;
  memset(&tm_in, 0, sizeof(tm_in));
  memset(&tm_out, 0, sizeof(tm_out));
  cuuint64_t dims[4] = {(cuuint64_t)p.Cin, (cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B};
  cuuint64_t strides[3] = {(cuuint64_t)p.Cin * 4, (cuuint64_t)p.Cin * 4 * p.W, (cuuint64_t)p.Cin * 4 * p.W * p.H};
  cuuint32_t box[4] = {32, (cuuint32_t)P.P, (cuuint32_t)(P.TH + 2), 1};
  if (!ch_encode_map(&tm_in, p.in, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return false;
  const size_t smem = std::max<size_t>((size_t)P.off_ctrl + CH_CTRL_BYTES + 1024, 120 * 1024);
  ensure_max_dynamic_smem((const void*)conv_halo_tc, m->ctx->device, (int)CH_SMEM_MAX);
  static const bool dbg = getenv("OAR_DBG_TILES") != nullptr;
  if (dbg)
    fprintf(stderr, "[fold] 3x3 B=%d %dx%d C=%d N=%d -> tile %dx%d pitch %d stages in %d resident w %d items %d smem %zu\n", p.B, p.H,
            p.W, p.Cin, p.N, P.TH, P.TW, P.P, P.ns_in, P.nb, P.n_work, smem);
  const int grid = std::min(P.n_work, m->ctx->sm_count);
  Launch l(m->ctx, name, 2.0 * p.M * p.N * p.K, 4.0 * ((double)p.M * p.Cin + (double)p.M * p.N));
  conv_halo_tc<<<grid, CH_THREADS, smem, m->ctx->stream>>>(P, tm_in, tm_out);
  return true;
}

}  // namespace oar
