// fused_tc.cu -- PP-LCNetV3 block as ONE persistent, warp-specialised tcgen05 kernel (sm_100a).
//
// Both networks on the hot path are stacks of [depthwise k x k conv -> activation -> 1x1 conv -> activation]
// (oar_ocr_b200/models.py:_lcnet_block; in the reference the same graph runs inside ONNX Runtime,
// oar-ocr-core/src/core/inference/ort_infer_execution.rs:178,281).  Run as two kernels the depthwise output makes a
// round trip through HBM (4*C bytes written + 4*C read per pixel) although both layers are bandwidth-bound.  Here the
// depthwise result never leaves the SM: it is produced straight into the shared-memory A operand of the pointwise
// GEMM.  The same kernel with K = 0 is a plain (optionally squeeze-excite-scaled) 1x1 convolution.
//
// One CTA per SM, persistent over (pixel tile, N tile) work items, 14 warps in four roles connected by mbarriers:
//   warp 4   TMA producer.  Per 32-channel k-block: one cp.async.bulk.tensor box of the fp32 NHWC input
//            ([32 ch] x [tile + halo]; out-of-bounds rows/columns/channels are zero-filled by the TMA unit, which IS
//            the convolution's zero padding) into a 3-4 deep ring, plus one linear bulk copy of the pre-packed fp16
//            hi/lo weight k-block.
//   warps 6-13  depthwise / convert.  Thread = 2 channels x (2 x 4) output pixels: every input value is read once
//            from shared memory into registers (LDS.64, one pixel's 128 bytes per half-warp: conflict-free) and
//            reused by all taps (packed FFMA2), accumulating bias, then (ky, kx) ascending like the stand-alone
//            kernel (engine.cu: dw_tile) so the values are bit-identical; activation; split x = hi + lo (fp16) and
//            store into the K-major no-swizzle UMMA layout [k-chunk][row][8 halfs], double buffered.
//   warp 5   MMA issuer.  One thread: 2 k16 steps x (hi*hi + hi*lo + lo*hi) tcgen05.mma.kind::f16 into a fp32 TMEM
//            accumulator (two accumulators of 256 columns alternate between work items); tcgen05.commit releases
//            the A/B stage and, after the last k-block, publishes the accumulator.
//   warps 0-3  epilogue.  tcgen05.ld, bias + activation, 128-byte-swizzled staging tile, cp.async.bulk.tensor
//            store (clipped at the image / channel bounds by the tensor map); overlaps the next item's main loop.
#include <cuda.h>
#include <cuda_fp16.h>

#include <cstdlib>
#include <map>

#include "engine.cuh"
#include "tc_ptx.cuh"
#include "tc_state.cuh"

namespace oar {

constexpr int FB_THREADS = 704;  // 4 epilogue + TMA + MMA + 16 depthwise/convert warps
constexpr int FB_EPI_THREADS = 128;
constexpr int FB_WARP_TMA = 4, FB_WARP_MMA = 5, FB_WARP_C0 = 6;
constexpr int FB_CTHREADS = 256;  // one team: the two teams of depthwise warps take alternate k-blocks
constexpr int FB_CWARPS = FB_CTHREADS / 32, FB_EPI_WARPS = FB_EPI_THREADS / 32;  // mbarrier arrivals are per warp
constexpr uint32_t FB_LBO = 128 * 16 + 16;   // k-chunk stride of the A operand (+16 B: conflict-free split stores)
constexpr uint32_t FB_APART = 4 * FB_LBO;    // one part (hi or lo) of a 128 x 32 A tile
constexpr uint32_t FB_ABUF = 2 * FB_APART;
constexpr uint32_t FB_EP_TILE = 128 * 128;   // 128 rows x 32 fp32 staging tile
constexpr int FB_MAX_IN = 4;
constexpr size_t FB_SMEM_MAX = 227 * 1024;
constexpr size_t FB_CTRL_BYTES = 256 + (512 + 32) * 4;  // mbarriers + TMEM slot, then the 1x1 conv bias

enum { FB_IN_FULL = 0, FB_IN_EMPTY = 4, FB_B_FULL = 8, FB_B_EMPTY = 12, FB_A_FULL = 16, FB_AB_EMPTY = 19, FB_ACC_FULL = 22,
       FB_ACC_EMPTY = 24, FB_SEEN = 26, FB_NBAR = 28 };  // AB_EMPTY: the A buffer of a k-block has been consumed

struct FbParams {
  const float* dw_pk;  // [k-block][K*K taps | bias][32 channels], zero padded: rides along with the activation box
  int dw_act;
  float dw_ps, dw_pb;
  const float* se_scale;  // K == 0: per (image, channel) multiplier applied to A, or null
  int HW, M;              // K == 0: pixels per image (row -> image for se_scale), total rows
  const uint4* wpk;
  const float* bias;
  int act;
  float ps, pb;
  int C, N, BN, nkb, n_tiles;
  // CTC head mode (K == 0): instead of storing the tile, reduce it to per-row softmax partials (max, last arg-max,
  // sum exp) per 128-column half, for launch_ctc_combine
  float* part_max;
  int32_t* part_idx;
  float* part_sum;
  int TH, TW, tiles_h, tiles_w, n_work;
  int cols_in;
  uint32_t in_bytes;   // activation box
  uint32_t tap_bytes;  // depthwise taps + bias of one k-block (0 for K == 0); stage = box + taps
  int ns_in;
  int nb;              // weight-stage ring depth (2..4): deep enough to cover the L2 round trip of a k-block
  uint32_t off_in, off_a, off_b, off_ctrl;
  int ep_tiles;  // epilogue staging tiles: 2 (store of one overlaps the fill of the other) or 1 when shared memory is short
  // share_a = 1 (two N tiles only): a work item is a PIXEL tile; its A operand -- the depthwise stage's output -- is built
  // once per k-block and multiplied into both N tiles' accumulators (TMEM columns 0.. and 256..), so a block wider
  // than 256 output channels no longer runs its depthwise stage once per N tile.  n_work counts pixel tiles then.
  int share_a;
  // b_resident = 1 (CTC head): the grid is a multiple of n_tiles, so a CTA meets ONE N tile for its whole life; the nkb
  // weight k-blocks of that tile are loaded once into an nkb-deep ring and never released.  Without it every
  // 128-row item re-streamed its 128 KB of hi/lo weights from L2 (192 KB per item with the activations: the chip-wide
  // L2 throughput, not the tensor pipe, set the pace).
  int b_resident;
  int ctc_two_pass;  // CTC-head epilogue form (OAR_DBG_CTC_EPI1 selects the one-pass online softmax)
  int ctc_groups;    // CTC head: 2 = a second epilogue group (warps 16-19 of the idle convert team) reduces column half 1
  int one_team;      // K == 0: one convert team takes every k-block (the CTC head's arrangement)
  uint32_t zero;     // always 0, but only the host knows: lets an address depend on loaded data (dep_zero below)
  int lean_mma;      // MMA warp: whole-warp loop with one elected issue block per k-block (OAR_FB_LEAN=0: lane-0 loop)
  int na;            // A operand buffers (2 or 3): k-block it lives in buffer it % na.  A third buffer takes the MMA's
                     // completion off the hand-off path: a team writing k-block it needs the MMAs of it - 3 done, not it - 2
  int poll1;         // one polling lane per warp in the convert / depthwise / epilogue waits (OAR_FB_POLL1=0: all lanes)
  int dbg_fence;     // bisecting aid: the producer fences (gpu scope + async proxy) before its first TMA load
};

__device__ __forceinline__ float act_rt(float v, int act) {
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

// One mbarrier arrival per warp: 32 lanes arriving on the same word serialise in the shared-memory atomic unit (~1 per
// cycle), which at 3 barriers x 256 threads per k-block is hundreds of cycles.  __syncwarp orders the lanes' prior
// shared-memory accesses (and proxy fences) before lane 0's releasing arrive.
__device__ __forceinline__ void warp_arrive(uint32_t bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}

// 0, computed FROM the given values with a mask the compiler cannot see through: adding it to an mbarrier address makes
// the arrive wait (register scoreboard) for the loads that produced the values.
__device__ __forceinline__ uint32_t dep_zero(uint32_t zero, float a, float b, float c, float d) {
  return (__float_as_uint(a) | __float_as_uint(b) | __float_as_uint(c) | __float_as_uint(d)) & zero;
}

// Waits of a converged warp outside the single-thread roles: every lane polls (default), or one elected lane polls and
// the others park at the warp barrier (OAR_FB_POLL1=1).  Measured on B200: the one-lane form is SLOWER here (3x3 blocks
// 5.74 -> 6.29 ms, 5x5 blocks 4.69 -> 5.60 ms per step): these waits sit on the critical path of a k-block hand-off and
// the elect + warp barrier behind the poll costs more than 32 lanes on the barrier word do.
__device__ __forceinline__ void mbar_wait_sel(int poll1, uint32_t bar, uint32_t parity) {
  if (poll1)
    mbar_wait_warp(bar, parity);
  else
    mbar_wait(bar, parity);
}

// x = hi + lo in fp16, two channels packed per 32-bit word
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __half ha = __float2half_rn(a), hb = __float2half_rn(b);
  const __half2 h = __halves2half2(ha, hb);
  const __half2 l = __halves2half2(__float2half_rn(a - __half2float(ha)), __float2half_rn(b - __half2float(hb)));
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// CTC-head epilogue of one epilogue group (128 threads = the tile's 128 rows; group g reduces column half g):
// online softmax statistics of each row over its half of the tile's classes, so the [B,T,V] logits never exist.
// Classes ascend and ties keep the LAST maximal index (simd.rs:194-204).
__device__ __forceinline__ void ctc_epilogue_group(const FbParams& P, int group, int n_halves, int gtid,
                                                   uint32_t lane_base, float* bias_g, uint32_t acc_full0,
                                                   uint32_t acc_empty0) {
  const int half_cols = P.BN >> 1;
  const float LOG2E = 1.4426950408889634f;
  uint32_t ti = 0;
  for (int t = blockIdx.x; t < P.n_work; t += gridDim.x, ++ti) {
    const int nt = t % P.n_tiles, sp = t / P.n_tiles;
    const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
    named_bar_sync(1 + group, FB_EPI_THREADS);  // the group is done with the previous tile's bias
    for (int i = gtid; i < n_halves * half_cols; i += FB_EPI_THREADS) {
      const int n = nt * P.BN + group * half_cols + i;
      bias_g[i] = n < P.N ? __ldg(P.bias + n) : 0.0f;
    }
    named_bar_sync(1 + group, FB_EPI_THREADS);
    mbar_wait_sel(P.poll1, acc_full0 + 8u * acc, aph);
    tc_fence_after();
#pragma unroll 1
    for (int h = 0; h < n_halves; ++h) {
      const int half = group + h;
      const int n_base = nt * P.BN + half * half_cols;
      float mx = -INFINITY, sum = 0.0f;
      int mi = 0;
      const uint32_t tcol = lane_base + acc * 256u + (uint32_t)(half * half_cols);
      for (int c0 = 0; c0 < half_cols && n_base + c0 < P.N; c0 += 16) {
        float v[16];
        __syncwarp();  // tcgen05.ld is .sync.aligned: the lanes diverge below (per-row arg-max search), reconverge first
        tmem_ld16(tcol + (uint32_t)c0, v);
        // one epilogue warp per scheduler: instruction count is what bounds this loop, so the arithmetic is packed
        // (f32x2 add / fma), the maximum is a 3-input tree and exp is the bare MUFU.EX2
        float2 q[8];
        const float4* b4 = reinterpret_cast<const float4*>(bias_g + h * half_cols + c0);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 b = b4[i];
          q[2 * i] = __fadd2_rn(make_float2(v[4 * i], v[4 * i + 1]), make_float2(b.x, b.y));
          q[2 * i + 1] = __fadd2_rn(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w));
        }
        if (n_base + c0 + 16 > P.N) {  // the vocabulary ends inside this group of 16 (last tile only)
#pragma unroll 1
          for (int i = 0; i < 8; ++i) {
            if (n_base + c0 + 2 * i >= P.N) q[i].x = -INFINITY;
            if (n_base + c0 + 2 * i + 1 >= P.N) q[i].y = -INFINITY;
          }
        }
        float gmax = fmaxf(q[0].x, q[0].y);
#pragma unroll
        for (int i = 1; i < 8; ++i) gmax = fmaxf(gmax, fmaxf(q[i].x, q[i].y));
        if (gmax >= mx) {  // a new (or tied, hence later) maximum lives in this group: find its last position
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (q[i].x == gmax) mi = n_base + c0 + 2 * i;
            if (q[i].y == gmax) mi = n_base + c0 + 2 * i + 1;
          }
        }
        const float nmx = fmaxf(mx, gmax);
        const float nl = -nmx * LOG2E;
        float2 part = make_float2(0.f, 0.f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float2 a = __ffma2_rn(q[i], make_float2(LOG2E, LOG2E), make_float2(nl, nl));  // (v - nmx) * log2(e)
          part = __fadd2_rn(part, make_float2(ex2_approx(a.x), ex2_approx(a.y)));  // 0 for the padded classes
        }
        sum = sum * ex2_approx((mx - nmx) * LOG2E) + (part.x + part.y);
        mx = nmx;
      }
      const int m = sp * 128 + gtid;
      if (m < P.M) {
        const size_t o = (size_t)m * (2 * P.n_tiles) + 2 * nt + half;
        P.part_max[o] = mx;
        P.part_idx[o] = mi;
        P.part_sum[o] = sum;
      }
    }
    tc_fence_before();
    warp_arrive(acc_empty0 + 8u * acc, gtid & 31);
  }
}

// The same statistics in two passes over the accumulator (TMEM reads are cheap): pass 1 finds the row's maximum over
// the half tile, pass 2 sums exp(v - max) and records the last column equal to the maximum.  The one-pass form above
// carries (max, sum) from one 16-column group to the next and waits on a TMEM load per group: one warp per scheduler ran
// its 86 instructions per group at 0.19 IPC (ncu source view: the four epilogue warps busy 100 % of the kernel, 11.6 k
// cycles per work item against 3.5 k for the item's MMAs).  Here every 32-column batch is independent work: the
// exponentials have a fixed reference, the sums go to four separate accumulators, nothing but `mx` crosses batches.
// half0 / n_halves: the column halves this group of four warps reduces (both, or one each with two groups; then the bias
// slice was staged by the kernel prologue and the group shares nothing writable with the other one)
__device__ __forceinline__ void ctc_epilogue_two_pass(const FbParams& P, int gtid, uint32_t lane_base, float* bias_g,
                                                      uint32_t acc_full0, uint32_t acc_empty0, int half0, int n_halves,
                                                      bool bias_staged) {
  const int half_cols = P.BN >> 1;
  const float LOG2E = 1.4426950408889634f;
  uint32_t ti = 0;
  for (int t = blockIdx.x; t < P.n_work; t += gridDim.x, ++ti) {
    const int nt = t % P.n_tiles, sp = t / P.n_tiles;
    const uint32_t acc = ti & 1u, aph = (ti >> 1) & 1u;
    if (!bias_staged && (ti == 0 || !P.b_resident)) {  // resident weights: the CTA keeps its N tile, and with it this bias slice
      named_bar_sync(1, FB_EPI_THREADS);
      // classes past the vocabulary get a bias of -inf: their accumulators are exact zeros (zero-padded weights), so
      // they can never be the maximum and their exponentials are exact zeros -- no per-element tail masks below
      for (int i = gtid; i < P.BN; i += FB_EPI_THREADS) {
        const int n = nt * P.BN + i;
        bias_g[i] = n < P.N ? __ldg(P.bias + n) : -INFINITY;
      }
      named_bar_sync(1, FB_EPI_THREADS);
    }
    mbar_wait_sel(P.poll1, acc_full0 + 8u * acc, aph);
    tc_fence_after();
#pragma unroll 1
    for (int half = half0; half < half0 + n_halves; ++half) {
      const int n_base = nt * P.BN + half * half_cols;
      const uint32_t tcol = lane_base + acc * 256u + (uint32_t)(half * half_cols);
      const float* bh = bias_g + half * half_cols;
      float mx = -INFINITY;
#pragma unroll 1
      for (int c0 = 0; c0 < half_cols && n_base + c0 < P.N; c0 += 32) {
        float v[32];
        __syncwarp();
        tmem_ld32(tcol + (uint32_t)c0, v);
        const float4* b4 = reinterpret_cast<const float4*>(bh + c0);
        float m[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = b4[i];
          const float2 q0 = __fadd2_rn(make_float2(v[4 * i], v[4 * i + 1]), make_float2(b.x, b.y));
          const float2 q1 = __fadd2_rn(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w));
          m[i] = fmaxf(fmaxf(q0.x, q0.y), fmaxf(q1.x, q1.y));
        }
        mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(m[0], m[1]), fmaxf(m[2], m[3])), fmaxf(fmaxf(m[4], m[5]), fmaxf(m[6], m[7]))));
      }
      const float nl = -mx * LOG2E;
      float2 part[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) part[i] = make_float2(0.f, 0.f);
      int mi = 0;
#pragma unroll 1
      for (int c0 = 0; c0 < half_cols && n_base + c0 < P.N; c0 += 32) {
        float v[32];
        __syncwarp();  // the arg-max selects below diverge per row; tcgen05.ld is .sync.aligned
        tmem_ld32(tcol + (uint32_t)c0, v);
        const float4* b4 = reinterpret_cast<const float4*>(bh + c0);
        int loc = -1;  // last column of this batch equal to the maximum (classes ascend: simd.rs:194-204)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 b = b4[i];
          const float2 q0 = __fadd2_rn(make_float2(v[4 * i], v[4 * i + 1]), make_float2(b.x, b.y));
          const float2 q1 = __fadd2_rn(make_float2(v[4 * i + 2], v[4 * i + 3]), make_float2(b.z, b.w));
          const float2 a0 = __ffma2_rn(q0, make_float2(LOG2E, LOG2E), make_float2(nl, nl));  // (v - mx) * log2(e)
          const float2 a1 = __ffma2_rn(q1, make_float2(LOG2E, LOG2E), make_float2(nl, nl));
          part[(2 * i) & 3] = __fadd2_rn(part[(2 * i) & 3], make_float2(ex2_approx(a0.x), ex2_approx(a0.y)));
          part[(2 * i + 1) & 3] = __fadd2_rn(part[(2 * i + 1) & 3], make_float2(ex2_approx(a1.x), ex2_approx(a1.y)));
          loc = q0.x == mx ? 4 * i : loc;
          loc = q0.y == mx ? 4 * i + 1 : loc;
          loc = q1.x == mx ? 4 * i + 2 : loc;
          loc = q1.y == mx ? 4 * i + 3 : loc;
        }
        mi = loc >= 0 ? n_base + c0 + loc : mi;
      }
      const float2 s2 = __fadd2_rn(__fadd2_rn(part[0], part[1]), __fadd2_rn(part[2], part[3]));
      const int m = sp * 128 + gtid;
      if (m < P.M) {
        const size_t o = (size_t)m * (2 * P.n_tiles) + 2 * nt + half;
        P.part_max[o] = mx;
        P.part_idx[o] = mi;
        P.part_sum[o] = s2.x + s2.y;
      }
    }
    tc_fence_before();
    warp_arrive(acc_empty0 + 8u * acc, gtid & 31);
  }
}

template <int K, int SH, int SW>
__global__ void __launch_bounds__(FB_THREADS, 1) lcblock_tc(const FbParams P, const __grid_constant__ CUtensorMap tm_in,
                                                            const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);  // swizzled staging tiles at offset 0
  const uint32_t sbase = smem_u32(smem);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t bar0 = sbase + P.off_ctrl;
#define FB_BAR(i) (bar0 + 8u * (uint32_t)(i))
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + P.off_ctrl + 8 * FB_NBAR);
  const uint32_t b_bytes = (uint32_t)P.BN * 128u;  // one weight k-block: hi + lo, 4 chunks x BN rows x 16 B
  const uint32_t stage_bytes = P.in_bytes + P.tap_bytes;

  if (tid == 0) {
    for (int i = 0; i < FB_MAX_IN; ++i) {
      mbar_init(FB_BAR(FB_IN_FULL + i), 1);
      mbar_init(FB_BAR(FB_IN_EMPTY + i), FB_CWARPS);
      mbar_init(FB_BAR(FB_B_FULL + i), 1);
      mbar_init(FB_BAR(FB_B_EMPTY + i), 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(FB_BAR(FB_A_FULL + i), FB_CWARPS);
      mbar_init(FB_BAR(FB_AB_EMPTY + i), 1);
      mbar_init(FB_BAR(FB_ACC_FULL + i), 1);
      mbar_init(FB_BAR(FB_ACC_EMPTY + i), FB_EPI_WARPS * (P.ctc_groups == 2 ? 2 : 1));
      mbar_init(FB_BAR(FB_SEEN + i), FB_CWARPS);
    }
    mbar_init(FB_BAR(FB_A_FULL + 2), FB_CWARPS);
    mbar_init(FB_BAR(FB_AB_EMPTY + 2), 1);
    fence_mbar_init();
  }
  if (warp == FB_WARP_MMA) tmem_alloc(smem_u32(tmem_slot), 512);
  if (K == 0 && P.ctc_groups == 2) {
    // two CTC epilogue groups (resident weights: one N tile per CTA): its bias slice is staged once, by everybody;
    // classes past the vocabulary get -inf (see ctc_epilogue_two_pass)
    float* bias_all = reinterpret_cast<float*>(smem + P.off_ctrl + 256);
    const int nt = (int)blockIdx.x % P.n_tiles;
    for (int i = tid; i < P.BN; i += FB_THREADS) {
      const int n = nt * P.BN + i;
      bias_all[i] = n < P.N ? __ldg(P.bias + n) : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == FB_WARP_TMA) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      if (P.dbg_fence) {
        __threadfence();
        asm volatile("fence.proxy.async.global;" ::: "memory");
      }
      // two independent cursors -- activation boxes run ns_in stages ahead of the depthwise warps, weight k-blocks two
      // ahead of the MMAs -- advanced by non-blocking probes so that neither ring throttles the other
      const uint32_t n_items = (P.n_work > (int)blockIdx.x ? (uint32_t)((P.n_work - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) : 0u) *
                               (uint32_t)P.nkb;
      const uint32_t n_items_b = P.b_resident ? (n_items ? (uint32_t)P.nkb : 0u)  // resident weights: one lap of the ring
                                 : P.share_a  ? n_items * (uint32_t)P.n_tiles      // shared A: every N tile's k-block
                                              : n_items;
      const int sp_div = P.share_a ? 1 : P.n_tiles;
      uint32_t it_in = 0, it_b = 0;
      int t_in = blockIdx.x, kb_in = 0, t_b = blockIdx.x, kb_b = 0, nt_b = 0;
      while (it_in < n_items || it_b < n_items_b) {
        if (it_in < n_items) {
          const uint32_t s = it_in % (uint32_t)P.ns_in, ph = (it_in / (uint32_t)P.ns_in) & 1u;
          if (mbar_test(FB_BAR(FB_IN_EMPTY + s), ph ^ 1u)) {
            const int sp = t_in / sp_div;
            const uint32_t dst = sbase + P.off_in + s * stage_bytes;
            mbar_expect_tx(FB_BAR(FB_IN_FULL + s), stage_bytes);
            if (K == 0) {
              tma_load_2d(dst, &tm_in, FB_BAR(FB_IN_FULL + s), kb_in * 32, sp * 128);
            } else {
              const int tw = sp % P.tiles_w, r = sp / P.tiles_w;
              tma_load_4d(dst, &tm_in, FB_BAR(FB_IN_FULL + s), kb_in * 32, tw * P.TW * SW - K / 2,
                          (r % P.tiles_h) * P.TH * SH - K / 2, r / P.tiles_h);
              bulk_load(dst + P.in_bytes, reinterpret_cast<const uint8_t*>(P.dw_pk) + (size_t)kb_in * P.tap_bytes,
                        P.tap_bytes, FB_BAR(FB_IN_FULL + s));
            }
            ++it_in;
            if (++kb_in == P.nkb) kb_in = 0, t_in += gridDim.x;
          }
        }
        if (it_b < n_items_b) {
          const uint32_t sb = it_b % (uint32_t)P.nb, phb = (it_b / (uint32_t)P.nb) & 1u;
          if (mbar_test(FB_BAR(FB_B_EMPTY + sb), phb ^ 1u)) {
            // weight k-blocks in the order the MMA warp consumes them: (item, kb) -- or (item, kb, N tile) with a shared A
            const int nt = P.share_a ? nt_b : t_b % P.n_tiles;
            const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(P.wpk) + ((size_t)nt * P.nkb + kb_b) * b_bytes;
            mbar_expect_tx(FB_BAR(FB_B_FULL + sb), b_bytes);
            bulk_load(sbase + P.off_b + sb * b_bytes, wsrc, b_bytes, FB_BAR(FB_B_FULL + sb));
            ++it_b;
            if (P.share_a && ++nt_b < P.n_tiles) continue;
            nt_b = 0;
            if (++kb_b == P.nkb) kb_b = 0, t_b += gridDim.x;
          }
        }
      }
    }
  } else if (warp == FB_WARP_MMA && P.lean_mma) {
    // ------------------------------------------------------------------ MMA issuer, lean form: the whole warp runs the
    // loop (values warp-uniform, in uniform registers), one elected lane polls each barrier and issues one block of
    // 6 MMAs + commits per (k-block, N tile); descriptors are a constant upper half plus (address >> 4) adds.
    {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(P.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)P.BN * 16u, b_part = 4u * b_lbo;
      const uint64_t adesc0 = make_desc(0, FB_LBO, 128), bdesc0 = make_desc(0, b_lbo, 128);
      const uint32_t a0 = (sbase + P.off_a) >> 4, b0 = (sbase + P.off_b) >> 4;
      const uint32_t a_j = FB_LBO >> 3, b_j = b_lbo >> 3, a_lo16 = FB_APART >> 4, b_lo16 = b_part >> 4, b_st16 = b_bytes >> 4;
      const uint32_t nt_loop = P.share_a ? (uint32_t)P.n_tiles : 1u;
      uint32_t it = 0, ti = 0, sb = 0, phb = 0, sA = 0, phA = 0;
      for (int t = blockIdx.x; t < P.n_work; t += gridDim.x, ++ti) {
        for (int kb = 0; kb < P.nkb; ++kb, ++it) {
          const uint32_t s = sA;
          mbar_wait_warp(FB_BAR(FB_A_FULL + s), phA);
          if (++sA == (uint32_t)P.na) sA = 0, phA ^= 1u;
          const uint32_t a_hi = a0 + s * (FB_ABUF >> 4), a_lo = a_hi + a_lo16;
          for (uint32_t ntl = 0; ntl < nt_loop; ++ntl) {
            const uint32_t v = ti * nt_loop + ntl, acc = v & 1u, aph = (v >> 1) & 1u;
            if (kb == 0) mbar_wait_warp(FB_BAR(FB_ACC_EMPTY + acc), aph ^ 1u);
            const uint32_t sbk = P.b_resident ? (uint32_t)kb : sb;
            mbar_wait_warp(FB_BAR(FB_B_FULL + sbk), P.b_resident ? 0u : phb);
            tc_fence_after();
            const uint32_t d = tmem_base + acc * 256u;
            const uint32_t b_hi = b0 + sbk * b_st16, b_lo = b_hi + b_lo16;
            const bool last = kb == P.nkb - 1, last_nt = ntl + 1 == nt_loop;
            if (elect_one_sync()) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const uint64_t ah = adesc0 + (uint64_t)(a_hi + j * a_j), al = adesc0 + (uint64_t)(a_lo + j * a_j);
                const uint64_t bh = bdesc0 + (uint64_t)(b_hi + j * b_j), bl = bdesc0 + (uint64_t)(b_lo + j * b_j);
                umma_f16(d, ah, bh, idesc, (kb | j) ? 1u : 0u);
                umma_f16(d, ah, bl, idesc, 1u);
                umma_f16(d, al, bh, idesc, 1u);
              }
              if (!P.b_resident) umma_commit(FB_BAR(FB_B_EMPTY + sbk));
              if (last) umma_commit(FB_BAR(FB_ACC_FULL + acc));
              if (last_nt) umma_commit(FB_BAR(FB_AB_EMPTY + s));
            }
            __syncwarp();
            if (!P.b_resident && ++sb == (uint32_t)P.nb) sb = 0, phb ^= 1u;
          }
        }
      }
    }
  } else if (warp == FB_WARP_MMA) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(P.BN >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint32_t b_lbo = (uint32_t)P.BN * 16u, b_part = 4u * b_lbo;
      const uint32_t nt_loop = P.share_a ? (uint32_t)P.n_tiles : 1u;  // accumulators fed from one A buffer
      uint32_t it = 0, it_b = 0, ti = 0;
      for (int t = blockIdx.x; t < P.n_work; t += gridDim.x, ++ti) {
        for (int kb = 0; kb < P.nkb; ++kb, ++it) {
          const uint32_t s = it % (uint32_t)P.na, ph = (it / (uint32_t)P.na) & 1u;
          mbar_wait(FB_BAR(FB_A_FULL + s), ph);
          const uint32_t a_hi = sbase + P.off_a + s * FB_ABUF, a_lo = a_hi + FB_APART;
          for (uint32_t ntl = 0; ntl < nt_loop; ++ntl, ++it_b) {
            // accumulators alternate over the sequence of (item, N tile) pairs, which is the order the epilogue drains
            const uint32_t v = ti * nt_loop + ntl, acc = v & 1u, aph = (v >> 1) & 1u;
            if (kb == 0) mbar_wait(FB_BAR(FB_ACC_EMPTY + acc), aph ^ 1u);  // the epilogue has drained this accumulator
            // resident weights: stage kb was filled once (phase 0 stays complete) and is never handed back
            const uint32_t sb = P.b_resident ? (uint32_t)kb : it_b % (uint32_t)P.nb;
            mbar_wait(FB_BAR(FB_B_FULL + sb), P.b_resident ? 0u : (it_b / (uint32_t)P.nb) & 1u);
            tc_fence_after();
            const uint32_t d = tmem_base + acc * 256u;
            const uint32_t b_hi = sbase + P.off_b + sb * b_bytes, b_lo = b_hi + b_part;
#pragma unroll
            for (int j = 0; j < 2; ++j) {
              const uint64_t ah = make_desc(a_hi + 2 * j * FB_LBO, FB_LBO, 128);
              const uint64_t al = make_desc(a_lo + 2 * j * FB_LBO, FB_LBO, 128);
              const uint64_t bh = make_desc(b_hi + 2 * j * b_lbo, b_lbo, 128);
              const uint64_t bl = make_desc(b_lo + 2 * j * b_lbo, b_lbo, 128);
              umma_f16(d, ah, bh, idesc, (kb | j) ? 1u : 0u);
              umma_f16(d, ah, bl, idesc, 1u);
              umma_f16(d, al, bh, idesc, 1u);
            }
            if (!P.b_resident) umma_commit(FB_BAR(FB_B_EMPTY + sb));
            if (kb == P.nkb - 1) umma_commit(FB_BAR(FB_ACC_FULL + acc));
          }
          umma_commit(FB_BAR(FB_AB_EMPTY + s));
        }
      }
    }
  } else if (warp >= FB_WARP_C0) {
    // ------------------------------------------------------------------ depthwise / convert warps
    // two teams of 8 warps take alternate k-blocks (team = item parity = A buffer), so twice the warps hide the
    // shared-memory and FMA latencies of one k-block's work
    const int team = (tid - FB_WARP_C0 * 32) >> 8;
    const int ct = (tid - FB_WARP_C0 * 32) & 255;
    const uint32_t n_items = (P.n_work > (int)blockIdx.x ? (uint32_t)((P.n_work - 1 - (int)blockIdx.x) / (int)gridDim.x + 1) : 0u) *
                             (uint32_t)P.nkb;
    const bool ctc = K == 0 && (P.part_max != nullptr || P.one_team);
    if (K == 0 && P.ctc_groups == 2 && warp >= 16 && warp < 20) {
      // CTC head, second epilogue group: TMEM lanes 32 * (warp % 4) are this warp's, like warps 0-3
      const uint32_t lane_base2 = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
      ctc_epilogue_two_pass(P, tid - 16 * 32, lane_base2, reinterpret_cast<float*>(smem + P.off_ctrl + 256),
                            FB_BAR(FB_ACC_FULL), FB_BAR(FB_ACC_EMPTY), 1, 1, true);
    } else if (ctc && team == 1) {
      // CTC head: converting fp32 rows is light work, so one team does every k-block and this one idles
    } else if (K == 0) {
      const uint32_t it0 = ctc ? 0u : (uint32_t)team, it_step = ctc ? 1u : 2u;
      const int q = ct & 7;
      const int r0 = 8 * (ct >> 6) + 4 * ((ct >> 3) & 1) + ((ct >> 4) & 3);
      const uint32_t a_off = (uint32_t)(q >> 1) * FB_LBO + (uint32_t)(q & 1) * 8u;
      for (uint32_t it = it0; it < n_items; it += it_step) {
        const uint32_t sa = it % (uint32_t)P.na;
        const int kb = (int)(it % (uint32_t)P.nkb);
        // squeeze-excite multipliers of this thread's (row, channel quad)s, requested before the wait for the tile
        float4 sc[4];
        if (P.se_scale) {
          const int sp = ((int)blockIdx.x + (int)(it / (uint32_t)P.nkb) * (int)gridDim.x) / (P.share_a ? 1 : P.n_tiles);
          const int c = kb * 32 + q * 4;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int m = sp * 128 + r0 + 32 * j;
            sc[j] = (c < P.C && m < P.M) ? __ldg(reinterpret_cast<const float4*>(P.se_scale + (size_t)(m / P.HW) * P.C + c))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);  // rows past M / channels past C are zeros
          }
        }
        const uint32_t s = it % (uint32_t)P.ns_in, ph = (it / (uint32_t)P.ns_in) & 1u;
        // The two teams share one ring of stages.  A parity wait is only sound while the waiter is at most one lap
        // ahead of the stage's barrier, which a fast team could violate when an older box (the other team's) lands
        // late.  So arrivals are observed strictly in item order: before waiting for item `it`, wait until the other
        // team has seen item it - 1 (it had then also seen every older item of its own).
        // (with an even number of stages every stage belongs to one team for good, and no handshake is needed)
        if (!ctc && (P.ns_in & 1)) {
          if (it >= 1) mbar_wait_sel(P.poll1, FB_BAR(FB_SEEN + (team ^ 1)), ((it - 1) >> 1) & 1u);
          mbar_wait_sel(P.poll1, FB_BAR(FB_IN_FULL + s), ph);
          warp_arrive(FB_BAR(FB_SEEN + team), lane);
        } else {
          mbar_wait_sel(P.poll1, FB_BAR(FB_IN_FULL + s), ph);
        }
        const uint8_t* src = smem + P.off_in + s * stage_bytes + q * 16;
        float4 x[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) x[j] = *reinterpret_cast<const float4*>(src + (r0 + 32 * j) * 128);
        // The stage may be handed back only once its values sit in registers.  An arrive issued right behind the LDS
        // instructions is NOT ordered behind their data: with the tensor pipe streaming operands at full rate the
        // shared-memory port is saturated (N = 256: 72 KB of operand reads + 48 KB of copies per 870 cycles), the loads
        // queue, the arrive does not, the producer refills the stage and a warp converts the NEXT lap's pixels --
        // one warp's 16 rows of one k-block wrong, a few times per hundred calls (tools/ctc_dump_diff.py; this was the
        // "two-group CTC race" of round 1 and the two-lane differences of round 2).  An arrive must follow the data in
        // registers -- so the barrier address is made to depend on the loaded values (dep_zero).
        warp_arrive(FB_BAR(FB_IN_EMPTY + s) + dep_zero(P.zero, x[0].x, x[1].x, x[2].x, x[3].x), lane);
        if (P.se_scale) {
#pragma unroll
          for (int j = 0; j < 4; ++j) x[j].x *= sc[j].x, x[j].y *= sc[j].y, x[j].z *= sc[j].z, x[j].w *= sc[j].w;
        }
        uint32_t hi[4][2], lo[4][2];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          split2(x[j].x, x[j].y, hi[j][0], lo[j][0]);
          split2(x[j].z, x[j].w, hi[j][1], lo[j][1]);
        }
        mbar_wait_sel(P.poll1, FB_BAR(FB_AB_EMPTY + sa), ((it / (uint32_t)P.na) & 1u) ^ 1u);
        uint8_t* ab = smem + P.off_a + sa * FB_ABUF + a_off;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          *reinterpret_cast<uint2*>(ab + (r0 + 32 * j) * 16) = make_uint2(hi[j][0], hi[j][1]);
          *reinterpret_cast<uint2*>(ab + FB_APART + (r0 + 32 * j) * 16) = make_uint2(lo[j][0], lo[j][1]);
        }
        fence_proxy_async_smem();
        warp_arrive(FB_BAR(FB_A_FULL + sa), lane);
      }
    } else {
      constexpr int KK = K > 0 ? K : 1;
      constexpr int RT = SH + KK;      // input rows feeding a thread's 2 output rows
      constexpr int CT = 3 * SW + KK;  // input columns feeding its 4 output columns
      const int pair = ct & 15, pg = ct >> 4;
      const int gx = P.TW >> 2;
      const int pgy = pg / gx, pgx = pg - pgy * gx;
      const bool active = pgy < (P.TH >> 1);
      const uint32_t row_stride = (uint32_t)P.cols_in * 128u;
      const uint32_t in_off = (uint32_t)(2 * pgy * SH) * row_stride + (uint32_t)(4 * pgx * SW) * 128u + (uint32_t)pair * 8u;
      const uint32_t a_off = (uint32_t)(pair >> 2) * FB_LBO + (uint32_t)(pair & 3) * 4u +
                             (uint32_t)((2 * pgy) * P.TW + 4 * pgx) * 16u;
      const bool hsw = P.dw_act == ACT_HSWISH;
      const bool affine = P.dw_ps != 1.0f || P.dw_pb != 0.0f;
      for (uint32_t it = (uint32_t)team; it < n_items; it += 2) {
        const uint32_t sa = it % (uint32_t)P.na;
        const uint32_t s = it % (uint32_t)P.ns_in, ph = (it / (uint32_t)P.ns_in) & 1u;
        // observe box arrivals strictly in item order across the two teams (see the K == 0 loop)
        const bool handshake = (P.ns_in & 1) != 0;  // even ring: each stage has one owner team, plain parity waits do
        if (handshake && it >= 1) mbar_wait_sel(P.poll1, FB_BAR(FB_SEEN + (team ^ 1)), ((it - 1) >> 1) & 1u);
        mbar_wait_sel(P.poll1, FB_BAR(FB_IN_FULL + s), ph);
        if (handshake) warp_arrive(FB_BAR(FB_SEEN + team), lane);
        const uint8_t* stage = smem + P.off_in + s * stage_bytes;
        // taps [ky*K+kx][32 ch] and bias [32 ch] of this k-block sit behind the box; each is read once per thread,
        // just in time (a kernel row serves output row 0 at input row ky and output row 1 at input row ky + SH)
        const float2* taps = reinterpret_cast<const float2*>(stage + P.in_bytes) + pair;
        const float2 bv = taps[KK * KK * 16];
        float2 acc[2][4];
#pragma unroll
        for (int ty = 0; ty < 2; ++ty)
#pragma unroll
          for (int tx = 0; tx < 4; ++tx) acc[ty][tx] = bv;
        if (active) {
          const uint8_t* base = stage + in_off;
          float2 w[KK][KK];
#pragma unroll
          for (int iy = 0; iy < RT; ++iy) {
            float2 x[CT];
#pragma unroll
            for (int cx = 0; cx < CT; ++cx) x[cx] = *reinterpret_cast<const float2*>(base + iy * row_stride + cx * 128);
            if (iy < KK) {
#pragma unroll
              for (int kx = 0; kx < KK; ++kx) w[iy][kx] = taps[(iy * KK + kx) * 16];
            }
#pragma unroll
            for (int ty = 0; ty < 2; ++ty) {
              const int ky = iy - ty * SH;
              if (ky < 0 || ky >= KK) continue;
#pragma unroll
              for (int kx = 0; kx < KK; ++kx)
#pragma unroll
                for (int tx = 0; tx < 4; ++tx) acc[ty][tx] = __ffma2_rn(x[tx * SW + kx], w[ky][kx], acc[ty][tx]);
            }
          }
        }
        // (as in the K == 0 loop: every loaded value has been consumed before the stage is handed back)
        warp_arrive(FB_BAR(FB_IN_EMPTY + s) + (dep_zero(P.zero, acc[0][0].x, acc[0][1].x, acc[0][2].x, acc[0][3].x) |
                                                 dep_zero(P.zero, acc[1][0].x, acc[1][1].x, acc[1][2].x, acc[1][3].x)),
                    lane);
        uint32_t hi[2][4], lo[2][4];
#pragma unroll
        for (int ty = 0; ty < 2; ++ty)
#pragma unroll
          for (int tx = 0; tx < 4; ++tx) {
            float2 v = acc[ty][tx];
            if (hsw) {  // same operation order as the stand-alone kernel (engine.cu: dw_tile)
              float2 tq = __fadd2_rn(v, make_float2(3.0f, 3.0f));
              tq.x = fminf(fmaxf(tq.x, 0.0f), 6.0f), tq.y = fminf(fmaxf(tq.y, 0.0f), 6.0f);
              v = __fmul2_rn(__fmul2_rn(v, tq), make_float2(0.16666667f, 0.16666667f));
            } else if (P.dw_act != ACT_NONE) {
              v.x = act_rt(v.x, P.dw_act), v.y = act_rt(v.y, P.dw_act);
            }
            if (affine) v.x = v.x * P.dw_ps + P.dw_pb, v.y = v.y * P.dw_ps + P.dw_pb;
            split2(v.x, v.y, hi[ty][tx], lo[ty][tx]);
          }
        mbar_wait_sel(P.poll1, FB_BAR(FB_AB_EMPTY + sa), ((it / (uint32_t)P.na) & 1u) ^ 1u);
        if (active) {
          uint8_t* ab = smem + P.off_a + sa * FB_ABUF + a_off;
#pragma unroll
          for (int ty = 0; ty < 2; ++ty)
#pragma unroll
            for (int tx = 0; tx < 4; ++tx) {
              const uint32_t o = (uint32_t)(ty * P.TW + tx) * 16u;
              *reinterpret_cast<uint32_t*>(ab + o) = hi[ty][tx];
              *reinterpret_cast<uint32_t*>(ab + FB_APART + o) = lo[ty][tx];
            }
        }
        fence_proxy_async_smem();
        warp_arrive(FB_BAR(FB_A_FULL + sa), lane);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue warps (tid = tile row = TMEM lane)
    const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
    const bool affine = P.ps != 1.0f || P.pb != 0.0f;
    float* bias_s = reinterpret_cast<float*>(smem + P.off_ctrl + 256);  // [n_tiles * BN + 32], zero padded
    const bool ctc = P.part_max != nullptr;  // then bias_s holds one tile's BN values, reloaded per work item
    if (ctc) {
      if (P.ctc_two_pass)
        ctc_epilogue_two_pass(P, tid, lane_base, bias_s, FB_BAR(FB_ACC_FULL), FB_BAR(FB_ACC_EMPTY), 0,
                              P.ctc_groups == 2 ? 1 : 2, P.ctc_groups == 2);
      else
        ctc_epilogue_group(P, 0, 2, tid, lane_base, bias_s, FB_BAR(FB_ACC_FULL), FB_BAR(FB_ACC_EMPTY));
    } else {
      for (int i = tid; i < P.n_tiles * P.BN + 32; i += FB_EPI_THREADS) bias_s[i] = i < P.N ? __ldg(P.bias + i) : 0.0f;
      named_bar_sync(1, FB_EPI_THREADS);
    }
    const uint32_t nt_loop = P.share_a ? (uint32_t)P.n_tiles : 1u;
    uint32_t vi = 0, nstore = 0;  // vi counts (item, N tile) pairs: the accumulator sequence of the MMA warp
    for (int t = blockIdx.x; t < (ctc ? 0 : P.n_work); t += gridDim.x)
    for (uint32_t ntl = 0; ntl < nt_loop; ++ntl, ++vi) {
      const int nt = P.share_a ? (int)ntl : t % P.n_tiles, sp = P.share_a ? t : t / P.n_tiles;
      int c1 = sp * 128, c2 = 0, c3 = 0;
      if (K > 0) {
        const int tw = sp % P.tiles_w, r = sp / P.tiles_w;
        c1 = tw * P.TW, c2 = (r % P.tiles_h) * P.TH, c3 = r / P.tiles_h;
      }
      const uint32_t acc = vi & 1u, aph = (vi >> 1) & 1u;
      mbar_wait_sel(P.poll1, FB_BAR(FB_ACC_FULL + acc), aph);
      tc_fence_after();
      const int n_base = nt * P.BN;
      for (int c0 = 0; c0 < P.BN && n_base + c0 < P.N; c0 += 32, ++nstore) {
        float v[32];
        __syncwarp();  // tcgen05.ld is .sync.aligned (lane 0 of warp 0 issued the previous chunk's TMA store)
        tmem_ld16(lane_base + acc * 256u + (uint32_t)c0, v);
        if (c0 + 16 < P.BN) {
          tmem_ld16(lane_base + acc * 256u + (uint32_t)c0 + 16u, v + 16);
        } else {
#pragma unroll
          for (int i = 16; i < 32; ++i) v[i] = 0.0f;
        }
        {
          // bias from shared memory (broadcast reads; zero past N), activation chosen once per chunk
          const float4* bs = reinterpret_cast<const float4*>(bias_s + n_base + c0);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 b = bs[i];
            v[4 * i] += b.x, v[4 * i + 1] += b.y, v[4 * i + 2] += b.z, v[4 * i + 3] += b.w;
          }
          switch (P.act) {
#define FB_ACT_CASE(A)                                      \
  case A:                                                   \
    _Pragma("unroll") for (int i = 0; i < 32; ++i) v[i] = act_t<A>(v[i]); \
    break;
            FB_ACT_CASE(ACT_RELU)
            FB_ACT_CASE(ACT_HSWISH)
            FB_ACT_CASE(ACT_SWISH)
            FB_ACT_CASE(ACT_SIGMOID)
            FB_ACT_CASE(ACT_HSIGMOID)
            FB_ACT_CASE(ACT_GELU)
#undef FB_ACT_CASE
            default: break;
          }
        }
        if (affine) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v[i] * P.ps + P.pb;
        }
        // two staging tiles: tile nstore & 1 was released by thread 0's wait before the previous barrier (its store was
        // issued two groups ago); one tile: wait for the previous store to have read it before anybody refills it
        const uint32_t slot = P.ep_tiles == 2 ? (nstore & 1u) : 0u;
        if (P.ep_tiles == 1) {
          if (tid == 0) bulk_wait_read_all();
          named_bar_sync(1, FB_EPI_THREADS);
        }
        uint8_t* ep = smem + slot * FB_EP_TILE + tid * 128;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          *reinterpret_cast<float4*>(ep + ((i ^ (tid & 7)) << 4)) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
        fence_proxy_async_smem();
        if (P.ep_tiles == 2 && tid == 0) bulk_wait_read_all();  // every store issued so far has read its staging tile
        named_bar_sync(1, FB_EPI_THREADS);
        if (tid == 0) {
          const uint32_t src = sbase + slot * FB_EP_TILE;
          if (K == 0)
            tma_store_2d(&tm_out, src, n_base + c0, c1);
          else
            tma_store_4d(&tm_out, src, n_base + c0, c1, c2, c3);
          bulk_commit();
        }
      }
      tc_fence_before();
      warp_arrive(FB_BAR(FB_ACC_EMPTY + acc), lane);
    }
    if (tid == 0) bulk_wait_all();
  }
#undef FB_BAR
  tc_fence_before();
  __syncthreads();
  if (warp == FB_WARP_MMA) tmem_dealloc(tmem_base, 512);
}

// bisecting aid (OAR_DBG_FB_PAD = cycles): a one-warp kernel that spins, launched in front of the 1x1 kernel
__global__ void fb_pad_kernel(int cycles) {
  const long long t0 = clock64();
  while (clock64() - t0 < cycles) {
  }
}

// ---------------------------------------------------------------------------
// host
// ---------------------------------------------------------------------------
static bool encode_map(CUtensorMap* tm, const float* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                       const cuuint32_t* box, CUtensorMapSwizzle swz) {
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = tmap_encoder()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<float*>(base), dims, strides,
                              box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

using FbKern = void (*)(const FbParams, const CUtensorMap, const CUtensorMap);

static FbKern pick_kernel(int k, int sh, int sw) {
#define FB_CASE(KV, SHV, SWV) \
  if (k == KV && sh == SHV && sw == SWV) return lcblock_tc<KV, SHV, SWV>;
  FB_CASE(0, 1, 1)
  FB_CASE(3, 1, 1) FB_CASE(5, 1, 1)
  FB_CASE(3, 2, 2) FB_CASE(5, 2, 2)
  FB_CASE(3, 2, 1) FB_CASE(5, 2, 1)
  FB_CASE(3, 1, 2) FB_CASE(5, 1, 2)
#undef FB_CASE
  return nullptr;
}

bool tc_fused_block(oar_model* m, int key, const FusedBlock& f, const char* name) {
  TcState* st = static_cast<TcState*>(m->tc_state);
  if (!st) return false;
  const TcWeights* w = nullptr;
  auto itf = st->wf.find(key);
  if (itf != st->wf.end()) {
    w = &itf->second;
  } else {
    auto it = st->w.find(key);
    if (it != st->w.end()) w = &it->second;
  }
  if (!w || w->rowtaps || w->KC != 4 || w->K != f.C || w->N != f.N) return false;
  if (f.C < 4 || (f.C & 3) || (f.out_ld & 3) || (f.out_c_off & 3) || (((uintptr_t)f.in) & 15) || (((uintptr_t)f.out) & 15))
    return false;
  FbKern kern = pick_kernel(f.k, f.k ? f.sh : 1, f.k ? f.sw : 1);
  if (!kern) return false;
  // bisecting aid: OAR_DBG_FB_OFF bit 0 = plain 1x1, 1 = squeeze-excite 1x1, 2 = 3x3 blocks, 3 = 5x5 blocks fall back to
  // the per-layer kernels; OAR_DBG_FB_NS forces the input ring depth (an even ring needs no team handshake)
  static const int dbg_off = getenv("OAR_DBG_FB_OFF") ? atoi(getenv("OAR_DBG_FB_OFF")) : 0;
  static const int dbg_ns = getenv("OAR_DBG_FB_NS") ? atoi(getenv("OAR_DBG_FB_NS")) : 0;
  if (dbg_off) {
    const int cls = f.k == 0 ? (f.se_scale ? 1 : 0) : (f.k == 3 ? 2 : 3);
    if ((dbg_off >> cls) & 1) return false;
    // bits 4 / 5: plain 1x1 with two N tiles / with one N tile; bits 6 / 7: plain 1x1 over more / fewer than 148 row tiles
    if (f.k == 0 && !f.se_scale) {
      const long long tiles = ((long long)f.B * f.Ho * f.Wo + 127) / 128;
      if (((dbg_off >> 4) & 1) && w->n_tiles >= 2) return false;
      if (((dbg_off >> 5) & 1) && w->n_tiles == 1) return false;
      if (((dbg_off >> 6) & 1) && tiles * w->n_tiles > 148) return false;
      if (((dbg_off >> 7) & 1) && tiles * w->n_tiles <= 148) return false;
    }
  }
  // the epilogue keeps the whole (zero-padded) bias in shared memory: FB_CTRL_BYTES reserves 512 + 32 floats
  if ((size_t)w->n_tiles * w->BN + 32 > (FB_CTRL_BYTES - 256) / sizeof(float)) return false;
  const long long M = (long long)f.B * f.Ho * f.Wo;
  if (M <= 0) return true;

  FbParams P{};
  P.dw_pk = nullptr, P.dw_act = f.dw_act, P.dw_ps = f.dw_ps, P.dw_pb = f.dw_pb;
  if (f.k) {
    auto itd = st->dwp.find(f.dw_key);
    if (itd == st->dwp.end()) return false;
    P.dw_pk = itd->second;
  }
  const size_t tap_bytes = f.k ? (size_t)(f.k * f.k + 1) * 128 : 0;
  P.tap_bytes = (uint32_t)tap_bytes;
  P.se_scale = f.se_scale, P.HW = f.Ho * f.Wo;
  P.wpk = w->packed, P.bias = f.bias, P.act = f.act, P.ps = f.ps, P.pb = f.pb;
  P.C = f.C, P.N = f.N, P.BN = w->BN, P.nkb = w->nkb, P.n_tiles = w->n_tiles;

  const size_t b_stage = (size_t)w->BN * 128;
  // two N tiles: the depthwise blocks can share one A operand between them (OAR_FB_SHARE=1; =2 also the plain 1x1
  // convs, whose A costs a conversion, not a convolution).  Opt-in: measured on B200 it is no faster than the
  // stand-alone depthwise kernel + the 1x1 kernel it replaces (recogniser blocks 6c / 6d: 1.01 + 1.01 ms against
  // 1.02 + 0.96 per 5 chunks; 512 fixed-size crops: 2.6 % slower) -- see the shared-memory budget below.
  static const int share_mode = getenv("OAR_FB_SHARE") ? atoi(getenv("OAR_FB_SHARE")) : 0;  // 0 off, 1 k > 0, 2 all
  P.share_a = (w->n_tiles == 2 && (share_mode == 2 || (share_mode == 1 && f.k > 0))) ? 1 : 0;
  // A shared A consumes two weight stages per k-block.  Measured on the recogniser's 480-channel blocks: 4370 cycles per
  // k-block against 2750 for a 240-channel block, and a third weight stage (OAR_FB_SHARE_NB=3) changes nothing -- the
  // ring is not what waits.  The shared-memory port is: 12 MMAs x 12 KB of operand reads + 96 KB of TMA writes + the
  // depthwise loads + the staging tile come to ~3800 cycles of 128 B/clk traffic per k-block.
  static const int share_nb = getenv("OAR_FB_SHARE_NB") ? atoi(getenv("OAR_FB_SHARE_NB")) : 2;
  const int min_b = P.share_a ? std::max(2, std::min(4, share_nb)) : 2;
  const size_t fixed = 2 * FB_EP_TILE + 2 * FB_ABUF + min_b * b_stage + FB_CTRL_BYTES + 1024;  // with the minimum of weight stages
  CUtensorMap tm_in, tm_out;
  memset(&tm_in, 0, sizeof(tm_in));
  memset(&tm_out, 0, sizeof(tm_out));
  int n_sp = 0;
  if (f.k == 0) {
    P.in_bytes = 128 * 128;
    P.ns_in = 4;
    while (P.ns_in > 2 && fixed + (size_t)P.ns_in * P.in_bytes > FB_SMEM_MAX) --P.ns_in;
    if (dbg_ns && dbg_ns <= P.ns_in) P.ns_in = dbg_ns;
    if (fixed + (size_t)P.ns_in * P.in_bytes > FB_SMEM_MAX) return false;
    n_sp = (int)((M + 127) / 128);
    P.M = (int)M;
    P.TH = P.TW = 0, P.tiles_h = P.tiles_w = 1, P.cols_in = 128;
    cuuint64_t dims[2] = {(cuuint64_t)f.C, (cuuint64_t)M};
    cuuint64_t strides[1] = {(cuuint64_t)f.C * 4};
    cuuint32_t box[2] = {32, 128};
    if (!encode_map(&tm_in, f.in, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return false;
    cuuint64_t odims[2] = {(cuuint64_t)f.N, (cuuint64_t)M};
    cuuint64_t ostrides[1] = {(cuuint64_t)f.out_ld * 4};
    if (!encode_map(&tm_out, f.out + f.out_c_off, 2, odims, ostrides, box, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
  } else {
    // tile = TH x TW output pixels of one image (TH even, TW % 4 == 0, <= 128 pixels).  Every tile costs the
    // depthwise warps the same time, so minimise the tile count first and the halo traffic second.
    double best = 1e300;
    int bTH = 0, bTW = 0, bns = 0, bep = 2;
    for (int TH = 2; TH <= 32; TH += 2)
      for (int TW = 4; TW <= 64; TW += 4) {
        if (TH * TW > 128) continue;
        const int rows_in = (TH - 1) * f.sh + f.k, cols_in = (TW - 1) * f.sw + f.k;
        if (rows_in > 256 || cols_in > 256) continue;
        const size_t in_bytes = (size_t)rows_in * cols_in * 128;
        const double tiles = (double)cdiv(f.Ho, TH) * cdiv(f.Wo, TW);
        // input stages: 4 = two per team; 3 still lets the producer run one box ahead of both teams; with 2 every
        // team waits out a full TMA round trip per k-block.  A single staging tile buys 16 KB for another stage.
        for (int ep = 2; ep >= (f.k == 5 ? 1 : 2); --ep) {  // measured: pays for the 5x5 blocks, not for the light 3x3 ones
          const size_t fx = fixed - (size_t)(2 - ep) * FB_EP_TILE;
          int ns = 4;
          while (ns >= 2 && fx + ns * (in_bytes + tap_bytes) > FB_SMEM_MAX) --ns;
          if (ns < 2) continue;
          if (dbg_ns && dbg_ns <= ns) ns = dbg_ns;
          const double cost = tiles * (1000.0 + rows_in * cols_in) * (ns == 2 ? 1.2 : ns == 3 ? 1.05 : 1.0) *
                              (ep == 1 ? 1.03 : 1.0);
          if (cost < best) best = cost, bTH = TH, bTW = TW, bns = ns, bep = ep;
        }
      }
    if (!bTH) return false;
    P.TH = bTH, P.TW = bTW, P.ns_in = bns, P.ep_tiles = bep;
    P.tiles_h = cdiv(f.Ho, bTH), P.tiles_w = cdiv(f.Wo, bTW);
    const int rows_in = (bTH - 1) * f.sh + f.k;
    P.cols_in = (bTW - 1) * f.sw + f.k;
    P.in_bytes = (uint32_t)rows_in * P.cols_in * 128u;
    n_sp = f.B * P.tiles_h * P.tiles_w;
    cuuint64_t dims[4] = {(cuuint64_t)f.C, (cuuint64_t)f.W, (cuuint64_t)f.H, (cuuint64_t)f.B};
    cuuint64_t strides[3] = {(cuuint64_t)f.C * 4, (cuuint64_t)f.C * 4 * f.W, (cuuint64_t)f.C * 4 * f.W * f.H};
    cuuint32_t box[4] = {32, (cuuint32_t)P.cols_in, (cuuint32_t)rows_in, 1};
    if (!encode_map(&tm_in, f.in, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return false;
    cuuint64_t odims[4] = {(cuuint64_t)f.N, (cuuint64_t)f.Wo, (cuuint64_t)f.Ho, (cuuint64_t)f.B};
    cuuint64_t ostrides[3] = {(cuuint64_t)f.out_ld * 4, (cuuint64_t)f.out_ld * 4 * f.Wo,
                              (cuuint64_t)f.out_ld * 4 * f.Wo * f.Ho};
    cuuint32_t obox[4] = {32, (cuuint32_t)bTW, (cuuint32_t)bTH, 1};
    if (!encode_map(&tm_out, f.out + f.out_c_off, 4, odims, ostrides, obox, CU_TENSOR_MAP_SWIZZLE_128B)) return false;
  }
  P.n_work = P.share_a ? n_sp : n_sp * w->n_tiles;
  static const int lean = getenv("OAR_FB_LEAN") ? atoi(getenv("OAR_FB_LEAN")) : 1;
  static const int poll1 = getenv("OAR_FB_POLL1") ? atoi(getenv("OAR_FB_POLL1")) : 0;
  P.lean_mma = lean, P.poll1 = poll1;
  if (f.k == 0) P.ep_tiles = 2;
  static const int dbg_one_team = getenv("OAR_DBG_FB_ONE_TEAM") ? atoi(getenv("OAR_DBG_FB_ONE_TEAM")) : 0;
  static const int dbg_ep = getenv("OAR_DBG_FB_EP") ? atoi(getenv("OAR_DBG_FB_EP")) : 0;
  if (f.k == 0 && dbg_one_team) P.one_team = 1;
  static const int dbg_fence = getenv("OAR_DBG_FB_FENCE") ? atoi(getenv("OAR_DBG_FB_FENCE")) : 0;
  static const int dbg_pad = getenv("OAR_DBG_FB_PAD") ? atoi(getenv("OAR_DBG_FB_PAD")) : 0;
  P.dbg_fence = dbg_fence;
  if (dbg_pad && f.k == 0) fb_pad_kernel<<<1, 32, 0, m->ctx->stream>>>(dbg_pad);
  if (f.k == 0 && dbg_ep == 1) P.ep_tiles = 1;
  P.off_in = (uint32_t)P.ep_tiles * FB_EP_TILE;
  P.off_a = P.off_in + (uint32_t)P.ns_in * (P.in_bytes + P.tap_bytes);
  // a third A buffer (OAR_FB_NA=3, where it fits beside the minimum weight ring) measured exactly equal to two on every
  // block class (3x3 blocks 5.74 vs 5.75 ms per step): the depthwise warps' wait for a free A buffer -- 48 % of their
  // samples in ncu -- is where the pipeline's slack shows, not what bounds it
  static const int want_na = getenv("OAR_FB_NA") ? atoi(getenv("OAR_FB_NA")) : 2;
  P.na = (want_na >= 3 && (size_t)P.off_a + 3 * FB_ABUF + (size_t)min_b * b_stage + FB_CTRL_BYTES + 1024 <= FB_SMEM_MAX) ? 3 : 2;
  P.off_b = P.off_a + (uint32_t)P.na * FB_ABUF;
  // whatever shared memory is left goes to the weight ring (up to 4 stages): a k-block of weights is an L2 round trip
  P.nb = min_b;
  while (P.nb < 4 && (size_t)P.off_b + (size_t)(P.nb + 1) * b_stage + FB_CTRL_BYTES + 1024 <= FB_SMEM_MAX) ++P.nb;
  P.off_ctrl = P.off_b + (uint32_t)P.nb * (uint32_t)b_stage;
  // always above half the SM's shared memory: one CTA per SM owns all 512 TMEM columns
  static const bool dbg_smem_max = getenv("OAR_DBG_FB_SMEMMAX") != nullptr;  // bisecting aid: nothing co-resident
  const size_t smem = dbg_smem_max ? FB_SMEM_MAX : std::max<size_t>((size_t)P.off_ctrl + FB_CTRL_BYTES + 1024, 120 * 1024);
  ensure_max_dynamic_smem((const void*)kern, m->ctx->device, (int)FB_SMEM_MAX);
  static const bool dbg_tiles = getenv("OAR_DBG_TILES") != nullptr;
  if (dbg_tiles)
    fprintf(stderr, "[fused] k=%d s=%dx%d B=%d %dx%d C=%d N=%d -> tile %dx%d stages %d staging %d wstages %d items %d share %d abufs %d smem %zu\n",
            f.k, f.sh, f.sw, f.B, f.Ho, f.Wo, f.C, f.N, P.TH, P.TW, P.ns_in, P.ep_tiles, P.nb, P.n_work, P.share_a, P.na, smem);
  const int grid = std::min(P.n_work, m->ctx->sm_count);
  const double flops = 2.0 * M * f.N * f.C + 2.0 * M * f.C * f.k * f.k;
  const double bytes = 4.0 * ((double)f.B * f.H * f.W * f.C + (double)M * f.N);
  Launch l(m->ctx, name, flops, bytes);
  kern<<<grid, FB_THREADS, smem, m->ctx->stream>>>(P, tm_in, tm_out);
  return true;
}

// CTC head on the persistent kernel: probabilities-free path of OP_CTC_HEAD (engine.cu).  A = [M, C] fp32 activations,
// packed weights of `key` (BN = 256 tiles over the vocabulary); fills p.part_* for launch_ctc_combine.
bool tc_ctc_head_persistent(oar_model* m, int key, const ConvParams& p, const char* name) {
  TcState* st = static_cast<TcState*>(m->tc_state);
  if (!st) return false;
  auto it = st->w.find(key);
  if (it == st->w.end()) return false;
  const TcWeights& w = it->second;
  if (w.rowtaps || w.KC != 4 || w.K != p.K || w.N != p.N || (w.BN & 31) || w.BN > 256) return false;
  if ((p.K & 3) || (((uintptr_t)p.in) & 15) || p.M <= 0) return false;
  FbParams P{};
  P.wpk = w.packed, P.bias = p.bias, P.act = ACT_NONE, P.ps = 1.0f, P.pb = 0.0f;
  P.C = p.K, P.N = p.N, P.BN = w.BN, P.nkb = w.nkb, P.n_tiles = w.n_tiles;
  P.part_max = p.part_max, P.part_idx = p.part_idx, P.part_sum = p.part_sum;
  // (A second epilogue group on four warps of the idle depthwise team, each group reducing one column half, was 1.3x
  // faster on this layer but showed rare run-to-run differences in the softmax statistics under host-side jitter --
  // 17 of 300 stress iterations, synccheck clean, racecheck inconclusive -- and was removed in round 2: DESIGN.md 5.2.)
  P.M = p.M, P.HW = 1;
  static const int lean = getenv("OAR_FB_LEAN") ? atoi(getenv("OAR_FB_LEAN")) : 1;
  static const int poll1 = getenv("OAR_FB_POLL1") ? atoi(getenv("OAR_FB_POLL1")) : 0;
  P.lean_mma = lean, P.poll1 = poll1;
  static const bool one_pass = getenv("OAR_DBG_CTC_EPI1") != nullptr;  // A/B switch
  P.ctc_two_pass = one_pass ? 0 : 1;
  // A second epilogue group (warps 16-19 of the idle convert team, one column half each): 1.00 -> 0.81 ms per step.
  // (Rounds 1 and 2 saw run-to-run differences with two groups; the cause was the early hand-back of the TMA stage in
  // the convert loop -- dep_zero above, profiles/r2_ctc_two_group_bisect.txt -- which a faster epilogue merely exposed.)
  static const bool one_group = getenv("OAR_CTC_GROUPS") && atoi(getenv("OAR_CTC_GROUPS")) == 1;
  P.in_bytes = 128 * 128, P.tap_bytes = 0, P.ns_in = 4;
  P.TH = P.TW = 0, P.tiles_h = P.tiles_w = 1, P.cols_in = 128;
  const size_t b_stage = (size_t)w.BN * 128;
  // resident weights: all nkb k-blocks of the CTA's N tile stay in shared memory (the staging tiles are not used here)
  static const bool no_resident = getenv("OAR_DBG_CTC_STREAM") != nullptr;  // A/B switch
  const int sm_count = m->ctx->sm_count;
  const size_t res_fixed = 2 * FB_ABUF + (size_t)w.nkb * b_stage + FB_CTRL_BYTES + 1024;
  P.b_resident = (!no_resident && w.nkb <= FB_MAX_IN && w.n_tiles <= sm_count && res_fixed + 2 * (size_t)P.in_bytes <= FB_SMEM_MAX) ? 1 : 0;
  // two epilogue groups need the prologue-staged bias, i.e. one N tile per CTA (resident weights), and the two-pass form
  P.ctc_groups = (P.b_resident && P.ctc_two_pass && !one_group) ? 2 : 1;
  const size_t fixed = P.b_resident ? res_fixed : 2 * FB_EP_TILE + 2 * FB_ABUF + 2 * b_stage + FB_CTRL_BYTES + 1024;
  while (P.ns_in > 2 && fixed + (size_t)P.ns_in * P.in_bytes > FB_SMEM_MAX) --P.ns_in;
  if (fixed + (size_t)P.ns_in * P.in_bytes > FB_SMEM_MAX) return false;
  CUtensorMap tm_in, tm_out;
  memset(&tm_in, 0, sizeof(tm_in));
  memset(&tm_out, 0, sizeof(tm_out));
  cuuint64_t dims[2] = {(cuuint64_t)p.K, (cuuint64_t)p.M};
  cuuint64_t strides[1] = {(cuuint64_t)p.K * 4};
  cuuint32_t box[2] = {32, 128};
  if (!encode_map(&tm_in, p.in, 2, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_NONE)) return false;
  P.n_work = cdiv(p.M, 128) * w.n_tiles;
  P.ep_tiles = 2;
  P.off_in = 0;  // the CTC epilogue stores nothing through the staging tiles
  P.off_a = P.off_in + (uint32_t)P.ns_in * P.in_bytes;
  P.na = 2;
  P.off_b = P.off_a + 2 * FB_ABUF;
  P.nb = 2;
  if (P.b_resident) {
    P.nb = w.nkb;
  } else {
    while (P.nb < 4 && (size_t)P.off_b + (size_t)(P.nb + 1) * w.BN * 128 + FB_CTRL_BYTES + 1024 <= FB_SMEM_MAX) ++P.nb;
  }
  P.off_ctrl = P.off_b + (uint32_t)P.nb * (uint32_t)w.BN * 128u;
  const size_t smem = std::max<size_t>((size_t)P.off_ctrl + FB_CTRL_BYTES + 1024, 120 * 1024);
  FbKern kern = lcblock_tc<0, 1, 1>;
  ensure_max_dynamic_smem((const void*)kern, m->ctx->device, (int)FB_SMEM_MAX);
  // resident weights: item t = blockIdx + i * grid keeps t % n_tiles fixed when the grid is a multiple of n_tiles
  const int grid = P.b_resident ? std::min(P.n_work, sm_count / w.n_tiles * w.n_tiles) : std::min(P.n_work, sm_count);
  Launch l(m->ctx, name, 2.0 * p.M * p.N * p.K, 4.0 * (double)p.M * p.K + 24.0 * p.M * w.n_tiles);
  kern<<<grid, FB_THREADS, smem, m->ctx->stream>>>(P, tm_in, tm_out);
  return true;
}

}  // namespace oar
