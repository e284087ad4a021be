// gemm_tc.cu -- tensor-core engine: implicit-GEMM convolution / linear layers on tcgen05 (sm_100a).
//
// Replaces, for the dense contractions of both networks (1x1 and kxk convs, 2x2-s2 transposed convs, the SVTR
// linears and the CTC head), the arithmetic ONNX Runtime does for the reference
// (oar-ocr-core/src/core/inference/ort_infer_execution.rs:178,281).
//
// Precision.  The parity bar is the reference's fp32 CPU run: identical boxes and CTC labels, probabilities within
// 1e-3.  A single fp16/bf16/tf32 pass misses it (measured on the oracle: fp16 operands move the DB map by 1.8e-3 and
// flip CTC arg-maxes), so every operand is split x = hi + lo into two fp16 values and each k-step issues three
// kind::f16 MMAs into the same fp32 TMEM accumulator:  hi*hi + hi*lo + lo*hi  (the dropped lo*lo term is 2^-22
// relative).  Products are exact in the fp32 accumulator, so the result matches an fp32 GEMM to ~1e-6 while the
// whole contraction still runs on the tensor pipe.  These layers are HBM-bound (mobile channel widths), so the 3x
// MMA count is free; activations stay fp32 in HBM exactly as the reference feeds them.
//
// Kernel shape (one CTA = one 128-row x BN-column output tile, 256 threads):
//   A (activations): threads gather (pixel row, 16-byte k-chunk) pairs from the fp32 NHWC tensor (im2col on the
//       fly, lanes of a warp reading contiguous bytes of the same pixel), split to hi/lo fp16 and store 16-byte
//       core-matrix rows into shared memory in the canonical K-major, no-swizzle UMMA layout
//       [k-chunk][row][8 halfs]  (LBO = 128 rows * 16 B + bank padding, SBO = 128 B);
//   B (weights): pre-split and pre-packed on the host at model load in exactly the shared-memory layout, so a stage
//       is one linear 16-byte-vector copy;
//   two shared-memory stages; tcgen05.commit -> mbarrier releases a stage when its MMAs have read it;
//   D: fp32 accumulators in TMEM (128 lanes x BN columns); epilogue = tcgen05.ld 32x32b, + bias, activation,
//       post-affine, then NHWC stores staged through shared memory so every store is a whole 128-byte row segment
//       (conv), 2x2 scatter (transposed conv) or an online softmax/arg-max reduction (CTC head: the [B,T,V] logits
//       never reach HBM).
#include <cuda.h>  // CUtensorMap (types only; the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_fp16.h>

#include <map>

#include "engine.cuh"
#include "tc_ptx.cuh"
#include "tc_state.cuh"

namespace oar {

constexpr int TC_BM = 128;        // UMMA M
// k elements per shared-memory stage: 64 (8 16-byte chunks, one stage) when 32 < K <= 64, else 32 (4 chunks, two
// stages): small stages keep 4+ CTAs resident per SM, which is what hides the gather latency
constexpr int TC_STAGES = 2;
constexpr int TC_MAX_BN = 256;

struct TcParams {
  ConvParams c;
  const uint4* wpk;
  int BN, nkb, n_tiles, tmem_cols, stages;
  uint32_t ctrl_off;  // byte offset of the mbarriers / TMEM slot behind the stages (and the epilogue staging tile)
  int tma_out;        // 1: the conv epilogue stores through the TMA tensor map passed next to these parameters
};

constexpr int TC_THREADS = 256;  // two threads per output row: each takes half the k-chunks and half the columns

// A-operand gather modes
constexpr int AM_POINTWISE = 0;  // 1x1 stride-1 conv / linear, Cin % 8 == 0: the row is contiguous
constexpr int AM_TAPS = 1;       // kxk conv, Cin % 8 == 0: a 16-byte chunk never straddles a tap
constexpr int AM_SCALAR = 2;     // anything else (stem Cin = 3, odd channel counts)
// epilogues: EPI = mode * 8 + activation (mode 0 conv, 1 transposed conv), EPI_CTC = fused CTC reduction
constexpr int EPI_CTC = 16;

// Per-KC geometry of the A operand.  256 threads cover the 128 x KC (row, 16-byte chunk) pairs of a k-block in
// KC/2 passes: chunk = tid % KC, row = tid / KC + pass * (256 / KC), so KC consecutive lanes read one pixel's
// contiguous KC*32 bytes (4 or 8 rows per warp instruction instead of 32 scattered rows).  The chunk stride in shared
// memory (the descriptor's leading byte offset) is padded by 128/KC bytes so the 16-byte stores of a quarter warp
// (which now differ in chunk as well as row) land in distinct bank groups.
template <int KC>
struct AGeo {
  static constexpr int NP = KC / 2;
  static constexpr int ROWS_PER_PASS = 256 / KC;
  static constexpr uint32_t LBO = TC_BM * 16 + 128 / KC;
  static constexpr uint32_t PART = KC * LBO;  // bytes of one A part (hi or lo) per stage
};

// One thread's share of a k-block of A as raw fp32 in registers, issued one k-block ahead so the HBM/L2 latency
// overlaps the previous block's split, barrier and MMAs.
template <int KC>
struct ARegs {
  float4 v[KC];  // NP pairs x 2 float4
};

struct ARow {  // per (thread, pass): the pixel this thread gathers for
  const float* ptr;  // pointwise: in + m * Cin
  int b, ho, wo;
  bool ok;
};

template <int KC, int A_MODE>
__device__ __forceinline__ void load_a(const ConvParams& p, int kb, int chunk, const ARow (&rows)[KC / 2],
                                       ARegs<KC>& r) {
  const int k0 = kb * (KC * 8) + chunk * 8;
#pragma unroll
  for (int j = 0; j < KC / 2; ++j) {
    const ARow& rw = rows[j];
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f), v = u;
    if (rw.ok && k0 < p.K) {
      if (A_MODE == AM_POINTWISE) {
        u = __ldg(reinterpret_cast<const float4*>(rw.ptr + k0));
        v = __ldg(reinterpret_cast<const float4*>(rw.ptr + k0 + 4));
      } else if (A_MODE == AM_TAPS) {
        int tap = k0 / p.Cin, ci = k0 - tap * p.Cin;
        int ky = tap / p.kw, kx = tap - ky * p.kw;
        int ih = rw.ho * p.sh - p.ph + ky, iw = rw.wo * p.sw - p.pw + kx;
        if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
          const float* src = p.in + (((size_t)rw.b * p.H + ih) * p.W + iw) * p.Cin + ci;
          u = __ldg(reinterpret_cast<const float4*>(src));
          v = __ldg(reinterpret_cast<const float4*>(src + 4));
        }
      } else {
        float x[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          x[i] = 0.0f;
          int k = k0 + i;
          if (k < p.K) {
            int tap = k / p.Cin, ci = k - tap * p.Cin;
            int ky = tap / p.kw, kx = tap - ky * p.kw;
            int ih = rw.ho * p.sh - p.ph + ky, iw = rw.wo * p.sw - p.pw + kx;
            if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
              x[i] = __ldg(p.in + (((size_t)rw.b * p.H + ih) * p.W + iw) * p.Cin + ci);
          }
        }
        u = make_float4(x[0], x[1], x[2], x[3]);
        v = make_float4(x[4], x[5], x[6], x[7]);
      }
    }
    r.v[2 * j] = u;
    r.v[2 * j + 1] = v;
  }
}

constexpr int EP_LD = 36;                          // floats per row of the epilogue staging tile (32 + pad)
constexpr int EP_BYTES = TC_BM * EP_LD * 4;        // 18432

// Conv epilogue: bias + activation in registers, then a 128 x 32 staging tile in (now idle) shared memory so that
// the global stores are whole 128-byte row segments (8 lanes per row) instead of 32 scattered 16-byte pieces.
// Rows m0 .. m0+nrows-1 of the [M, out_ld] output are written.
template <int ACT>
__device__ __forceinline__ void epi_conv_store(const ConvParams& p, uint8_t* smem, uint32_t lane_base, int BN,
                                               int n_base, int m0, int nrows, int tid) {
  const int row = tid & (TC_BM - 1), half = tid >> 7;
  {
    float* ep = reinterpret_cast<float*>(smem);
    const bool vec_ok = ((p.out_ld & 3) == 0) && ((p.out_c_off & 3) == 0) && ((n_base & 3) == 0) &&
                        ((((uintptr_t)p.out) & 15) == 0);
    const float ps = p.post_scale, pb = p.post_bias;
    const bool affine = ps != 1.0f || pb != 0.0f;  // the Act's learnable affine; identity in deploy graphs
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n_base + c0 >= p.N) break;
      const int cols_here = min(min(32, BN - c0), p.N - n_base - c0);  // the tile ends at BN, the tensor at N
      const int cc = c0 + 16 * half;
      if (cc < BN && n_base + cc < p.N) {  // warp-uniform
        float v[16];
        tmem_ld16(lane_base + cc, v);
        if (n_base + cc + 16 <= p.N) {
#pragma unroll
          for (int i = 0; i < 16; i += 4) {
            float4 b;
            if ((n_base & 3) == 0)
              b = __ldg(reinterpret_cast<const float4*>(p.bias + n_base + cc + i));  // weight arrays are 16-B aligned
            else
              b = make_float4(__ldg(p.bias + n_base + cc + i), __ldg(p.bias + n_base + cc + i + 1),
                              __ldg(p.bias + n_base + cc + i + 2), __ldg(p.bias + n_base + cc + i + 3));
            v[i] = act_t<ACT>(v[i] + b.x);
            v[i + 1] = act_t<ACT>(v[i + 1] + b.y);
            v[i + 2] = act_t<ACT>(v[i + 2] + b.z);
            v[i + 3] = act_t<ACT>(v[i + 3] + b.w);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            int n = n_base + cc + i;
            v[i] = n < p.N ? act_t<ACT>(v[i] + __ldg(p.bias + n)) : 0.0f;
          }
        }
        if (affine) {
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] = v[i] * ps + pb;
        }
        float4* dst = reinterpret_cast<float4*>(ep + row * EP_LD + 16 * half);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      __syncthreads();
      float* obase = p.out + (size_t)m0 * p.out_ld + p.out_c_off + n_base + c0;
      if (vec_ok && (cols_here & 3) == 0) {
        const int cpr = cols_here >> 2;  // float4 per row
        if (cpr == 8) {
#pragma unroll
          for (int idx = tid; idx < TC_BM * 8; idx += TC_THREADS) {
            const int r = idx >> 3, c4 = idx & 7;
            if (r < nrows)
              *reinterpret_cast<float4*>(obase + (size_t)r * p.out_ld + 4 * c4) =
                  *reinterpret_cast<const float4*>(ep + r * EP_LD + 4 * c4);
          }
        } else {
          for (int idx = tid; idx < TC_BM * cpr; idx += TC_THREADS) {
            const int r = idx / cpr, c4 = idx - r * cpr;
            if (r < nrows)
              *reinterpret_cast<float4*>(obase + (size_t)r * p.out_ld + 4 * c4) =
                  *reinterpret_cast<const float4*>(ep + r * EP_LD + 4 * c4);
          }
        }
      } else {
        for (int idx = tid; idx < TC_BM * cols_here; idx += TC_THREADS) {
          const int r = idx / cols_here, c = idx - r * cols_here;
          if (r < nrows) obase[(size_t)r * p.out_ld + c] = ep[r * EP_LD + c];
        }
      }
      __syncthreads();
    }
  }
}

// Conv epilogue through TMA: bias + activation in registers, 128 x 32 fp32 staging tiles in shared memory in the
// 128-byte-swizzled layout the tensor map describes (16-byte chunk index XOR row % 8: conflict-free for the
// row-per-thread writes), then ONE cp.async.bulk.tensor store per tile issued by one thread.  The TMA engine clips
// the box at the tensor bounds (rows >= M or pixels >= W, columns >= N), so no per-element predicates remain, and the
// global store costs no LSU instructions at all.  Two staging tiles alternate so the store of one group overlaps the
// TMEM reads of the next.  c1/c2 = tensor coordinates of the tile's first row.
template <int ACT>
__device__ __forceinline__ void epi_conv_store_tma(const ConvParams& p, const CUtensorMap* tm, uint8_t* smem,
                                                   uint32_t lane_base, int BN, int n_base, int c1, int c2, int rank3,
                                                   int tid) {
  const int row = tid & (TC_BM - 1), half = tid >> 7;
  const float ps = p.post_scale, pb = p.post_bias;
  const bool affine = ps != 1.0f || pb != 0.0f;
  int buf = 0;
  for (int c0 = 0; c0 < BN; c0 += 32, buf ^= 1) {
    if (n_base + c0 >= p.N) break;
    uint8_t* ep = smem + buf * (TC_BM * 128);
    const int cc = c0 + 16 * half;
    if (cc < BN && n_base + cc < p.N) {  // warp-uniform
      float v[16];
      tmem_ld16(lane_base + cc, v);
      if (n_base + cc + 16 <= p.N) {
#pragma unroll
        for (int i = 0; i < 16; i += 4) {
          float4 b;
          if ((n_base & 3) == 0)
            b = __ldg(reinterpret_cast<const float4*>(p.bias + n_base + cc + i));
          else
            b = make_float4(__ldg(p.bias + n_base + cc + i), __ldg(p.bias + n_base + cc + i + 1),
                            __ldg(p.bias + n_base + cc + i + 2), __ldg(p.bias + n_base + cc + i + 3));
          v[i] = act_t<ACT>(v[i] + b.x);
          v[i + 1] = act_t<ACT>(v[i + 1] + b.y);
          v[i + 2] = act_t<ACT>(v[i + 2] + b.z);
          v[i + 3] = act_t<ACT>(v[i + 3] + b.w);
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int n = n_base + cc + i;
          v[i] = n < p.N ? act_t<ACT>(v[i] + __ldg(p.bias + n)) : 0.0f;
        }
      }
      if (affine) {
#pragma unroll
        for (int i = 0; i < 16; ++i) v[i] = v[i] * ps + pb;
      }
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int c16 = 4 * half + i;  // 16-byte chunk of this row
        *reinterpret_cast<float4*>(ep + row * 128 + ((c16 ^ (row & 7)) << 4)) =
            make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
    }
    fence_proxy_async_smem();  // generic-proxy writes -> visible to the TMA (async proxy)
    __syncthreads();
    if (tid == 0) {
      const uint32_t src = smem_u32(ep);
      if (rank3)
        asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(tm),
                     "r"(n_base + c0), "r"(c1), "r"(c2), "r"(src)
                     : "memory");
      else
        asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%1, %2}], [%3];" ::"l"(tm),
                     "r"(n_base + c0), "r"(c1), "r"(src)
                     : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      // the buffer written two groups ago must have been read out before the next group reuses it
      asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
    }
    __syncthreads();
  }
  if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // stores complete before the CTA retires
}

template <int KC, int A_MODE, int EPI>
__global__ void __launch_bounds__(TC_THREADS) conv_gemm_tc(const TcParams P, const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ __align__(1024) uint8_t smem[];  // 1024: the swizzled TMA staging tiles start at offset 0
  constexpr int BK = KC * 8;
  using G = AGeo<KC>;
  const ConvParams& p = P.c;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int BN = P.BN;
  constexpr uint32_t a_part = G::PART;
  const uint32_t b_part = KC * BN * 16;
  const uint32_t stage_bytes = 2 * a_part + 2 * b_part;
  const int smask = P.stages - 1;
  uint8_t* ctrl = smem + P.ctrl_off;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(ctrl);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 8 * TC_STAGES);
  const int m0 = blockIdx.x * TC_BM;

  // the pixels this thread gathers A for
  const int chunk = tid % KC;
  ARow rows[G::NP];
#pragma unroll
  for (int j = 0; j < G::NP; ++j) {
    const int m = m0 + tid / KC + j * G::ROWS_PER_PASS;
    rows[j].ok = m < p.M;
    rows[j].ptr = p.in + (size_t)m * p.Cin;
    rows[j].b = rows[j].ho = rows[j].wo = 0;
    if (A_MODE != AM_POINTWISE && rows[j].ok) {
      int b = m / (p.Ho * p.Wo);
      int r = m - b * p.Ho * p.Wo;
      rows[j].b = b;
      rows[j].ho = r / p.Wo;
      rows[j].wo = r - rows[j].ho * p.Wo;
    }
  }
  const int nt = blockIdx.y;
  const uint4* wtile = P.wpk + (size_t)nt * P.nkb * (2 * KC * BN);
  const int nvec_b = 2 * KC * BN;  // 16-byte vectors of one B stage (hi then lo)

  // first two k-blocks of A into registers, B(0) straight into stage 0 (all in flight during the set-up below)
  ARegs<KC> pre0, pre1;
  load_a<KC, A_MODE>(p, 0, chunk, rows, pre0);
  if (KC == 4 && P.nkb > 1) load_a<KC, A_MODE>(p, 1, chunk, rows, pre1);  // KC = 8 tiles have a single k-block
  {
    const uint32_t b_dst = smem_u32(smem + 2 * a_part);
    for (int i = tid; i < nvec_b; i += TC_THREADS) cp_async16(b_dst + i * 16, wtile + i);
    cp_async_commit();
  }

  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(smem_u32(&mbar[s]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  // instruction descriptor: D fp32, A/B fp16, both K-major, N = BN, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);
  const uint32_t a_store_off = (uint32_t)chunk * G::LBO + (uint32_t)(tid / KC) * 16;

  // Software pipeline, per k-block kb (stage s = kb & 1):
  //   a. split the registers loaded two iterations ago into A(kb) -> stage s   (free since MMA(kb-2), waited last time)
  //   b. issue the global loads of A(kb+2) into the same registers
  //   c. wait for MMA(kb-1) (it ran under a/b), which frees stage s^1
  //   d. cp.async B(kb+1) -> stage s^1        e. wait for B(kb), issued one whole iteration ago
  //   f. fence, barrier, one thread issues the MMAs of kb and commits to mbar[s]
  auto step = [&](int kb, ARegs<KC>& pre) {
    const int s = kb & smask;
    uint8_t* st = smem + s * stage_bytes;
#pragma unroll
    for (int j = 0; j < G::NP; ++j) {
      float x[8] = {pre.v[2 * j].x,     pre.v[2 * j].y,     pre.v[2 * j].z,     pre.v[2 * j].w,
                    pre.v[2 * j + 1].x, pre.v[2 * j + 1].y, pre.v[2 * j + 1].z, pre.v[2 * j + 1].w};
      uint4 hi, lo;
      split8(x, hi, lo);
      const uint32_t off = a_store_off + (uint32_t)(j * G::ROWS_PER_PASS) * 16;
      *reinterpret_cast<uint4*>(st + off) = hi;
      *reinterpret_cast<uint4*>(st + a_part + off) = lo;
    }
    if (kb + 2 < P.nkb) load_a<KC, A_MODE>(p, kb + 2, chunk, rows, pre);
    if (kb + 1 < P.nkb) {
      if (kb >= 1) mbar_wait(smem_u32(&mbar[s ^ 1]), (uint32_t)(((kb - 1) >> 1) & 1));
      const uint32_t b_dst = smem_u32(smem + (s ^ 1) * stage_bytes + 2 * a_part);
      const uint4* b_src = wtile + (size_t)(kb + 1) * nvec_b;
      for (int i = tid; i < nvec_b; i += TC_THREADS) cp_async16(b_dst + i * 16, b_src + i);
      cp_async_commit();
      cp_async_wait_1();  // B(kb) has landed; B(kb+1) stays in flight
    } else {
      cp_async_wait_all();
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi_s = smem_u32(st), a_lo_s = a_hi_s + a_part;
      const uint32_t b_hi_s = a_hi_s + 2 * a_part, b_lo_s = b_hi_s + b_part;
      const uint32_t a_lbo = G::LBO, b_lbo = (uint32_t)BN * 16;
#pragma unroll
      for (int j = 0; j < BK / 16; ++j) {
        uint64_t ah = make_desc(a_hi_s + 2 * j * a_lbo, a_lbo, 128);
        uint64_t al = make_desc(a_lo_s + 2 * j * a_lbo, a_lbo, 128);
        uint64_t bh = make_desc(b_hi_s + 2 * j * b_lbo, b_lbo, 128);
        uint64_t bl = make_desc(b_lo_s + 2 * j * b_lbo, b_lbo, 128);
        umma_f16(tmem_base, ah, bh, idesc, (kb | j) ? 1u : 0u);
        umma_f16(tmem_base, ah, bl, idesc, 1u);
        umma_f16(tmem_base, al, bh, idesc, 1u);
      }
      umma_commit(smem_u32(&mbar[s]));
    }
  };
  if (KC == 4) {
    for (int kb = 0; kb < P.nkb; kb += 2) {
      step(kb, pre0);
      if (kb + 1 < P.nkb) step(kb + 1, pre1);
    }
  } else {
    for (int kb = 0; kb < P.nkb; ++kb) step(kb, pre0);  // nkb == 1 by construction: no second register set
  }
  // all MMAs done when the last commit lands (a commit tracks every prior tcgen05 op of the issuing thread)
  {
    const int last = P.nkb - 1;
    mbar_wait(smem_u32(&mbar[last & smask]), (uint32_t)((P.stages == 2 ? (last >> 1) : last) & 1));
    tc_fence_after();
  }

  // ---- epilogue: TMEM -> registers with thread = (row, column half); a warp may only touch TMEM lanes
  // 32*(warp%4)..+31, which is exactly `row` for both halves.
  const int row = tid & (TC_BM - 1), half = tid >> 7;
  const int m = m0 + row;
  const bool row_ok = m < p.M;
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  const int n_base = nt * BN;
  if (EPI == EPI_CTC) {
    // online softmax statistics over this thread's contiguous class range (BN % 32 == 0 for this mode)
    float mx = -INFINITY, sum = 0.0f;
    int mi = 0;
    const int cbeg = half * (BN >> 1), cend = cbeg + (BN >> 1);
    // branch-free per 16 classes: group max -> one rescale of the running sum -> 16 exps; the arg-max keeps the
    // LAST maximal index (simd.rs:194-204) because classes are visited in ascending order with >=
    for (int c0 = cbeg; c0 < cend; c0 += 16) {
      if (n_base + c0 >= p.N) break;
      float v[16];
      tmem_ld16(lane_base + c0, v);
      float gmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int n = n_base + c0 + i;
        v[i] = n < p.N ? v[i] + __ldg(p.bias + n) : -INFINITY;
        gmax = fmaxf(gmax, v[i]);
      }
      const float nmx = fmaxf(mx, gmax);
      float part = 0.0f;
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        part += __expf(v[i] - nmx);  // exp(-inf) = 0 for the padded classes
        if (v[i] >= mx && v[i] == nmx) mi = n_base + c0 + i;
      }
      sum = sum * __expf(mx - nmx) + part;
      mx = nmx;
    }
    if (row_ok) {
      size_t o = (size_t)m * (2 * P.n_tiles) + 2 * nt + half;
      p.part_max[o] = mx;
      p.part_idx[o] = mi;
      p.part_sum[o] = sum;
    }
  } else if (EPI < 8) {
    if (P.tma_out)
      epi_conv_store_tma<EPI & 7>(p, &tm_out, smem, lane_base, BN, n_base, m0, 0, 0, tid);
    else
      epi_conv_store<EPI & 7>(p, smem, lane_base, BN, n_base, m0, min(TC_BM, p.M - m0), tid);
  } else {
    // 2x2 stride-2 transposed conv: column n = (dy*2+dx)*cout + co belongs to output pixel (2y+dy, 2x+dx).  For a
    // fixed dy the 2*cout columns of an input pixel are one contiguous run of the output row 2y+dy, and consecutive
    // input pixels continue that run, so the tile is staged through shared memory like the conv epilogue and written
    // as 16-byte pieces of those runs.
    constexpr int ACT = EPI & 7;
    float* ep = reinterpret_cast<float*>(smem);
    const float ps = p.post_scale, pb = p.post_bias;
    const int run = 2 * p.cout;                     // columns per dy
    const size_t dy_stride = (size_t)2 * p.Wo * p.cout;  // floats between output rows 2y and 2y+1
    const bool vec_ok = (run & 3) == 0 && ((((uintptr_t)p.out) & 15) == 0);
    const int nrows = min(TC_BM, p.M - m0);
    for (int c0 = 0; c0 < BN; c0 += 32) {
      if (n_base + c0 >= p.N) break;
      const int cols_here = min(min(32, BN - c0), p.N - n_base - c0);
      const int cc = c0 + 16 * half;
      if (cc < BN && n_base + cc < p.N) {
        float v[16];
        tmem_ld16(lane_base + cc, v);
        int co = (n_base + cc) % p.cout;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          v[i] = (n_base + cc + i < p.N) ? act_t<ACT>(v[i] + __ldg(p.bias + co)) * ps + pb : 0.0f;
          if (++co == p.cout) co = 0;
        }
        float4* dst = reinterpret_cast<float4*>(ep + row * EP_LD + 16 * half);
#pragma unroll
        for (int i = 0; i < 4; ++i) dst[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
      }
      __syncthreads();
      if (vec_ok && (cols_here & 3) == 0 && ((n_base + c0) & 3) == 0) {
        const int cpr = cols_here >> 2;
        for (int idx = tid; idx < TC_BM * cpr; idx += TC_THREADS) {
          const int r = idx / cpr, c4 = idx - r * cpr;
          if (r < nrows) {
            const int mm = m0 + r;
            const int ob = mm / (p.Ho * p.Wo);
            const int rem = mm - ob * p.Ho * p.Wo;
            const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
            const int n = n_base + c0 + 4 * c4;
            const int dy = n / run, k = n - dy * run;
            float* o = p.out + (((size_t)ob * (2 * p.Ho) + 2 * oy) * (2 * p.Wo) + 2 * ox) * p.cout + dy * dy_stride + k;
            *reinterpret_cast<float4*>(o) = *reinterpret_cast<const float4*>(ep + r * EP_LD + 4 * c4);
          }
        }
      } else {
        for (int idx = tid; idx < TC_BM * cols_here; idx += TC_THREADS) {
          const int r = idx / cols_here, c = idx - r * cols_here;
          if (r < nrows) {
            const int mm = m0 + r;
            const int ob = mm / (p.Ho * p.Wo);
            const int rem = mm - ob * p.Ho * p.Wo;
            const int oy = rem / p.Wo, ox = rem - oy * p.Wo;
            const int n = n_base + c0 + c;
            const int dy = n / run, k = n - dy * run;
            p.out[(((size_t)ob * (2 * p.Ho) + 2 * oy) * (2 * p.Wo) + 2 * ox) * p.cout + dy * dy_stride + k] =
                ep[r * EP_LD + c];
          }
        }
      }
      __syncthreads();
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// ---------------------------------------------------------------------------
// Stride-1 k x k convolution ("row taps"): the generic im2col gather above re-reads every input pixel kh*kw times
// through L1/L2 (a 3x3 over 96 channels moves 9x the tensor).  Here a CTA owns up to 128 consecutive output pixels of
// ONE image row; a k-block is (ky, 32 input channels): the CTA stages the input row ih = h + ky - ph for pixels
// w0 - pw .. w0 + 127 + pw ONCE (130 rows of the K-major no-swizzle layout, whose row dimension is linear at 16 B per
// row because SBO = 8 * 16 B), and the kw taps of that row are issued as MMAs whose A descriptors start kx * 16 bytes
// further: a shifted view of the same shared-memory rows.  Gather traffic drops from kh*kw to ~kh reads per pixel.
// ---------------------------------------------------------------------------
constexpr int RT_KC = 4;            // 32 input channels per k-block
constexpr int RT_MAX_KW = 3;        // halo rows that fit the padded chunk stride (130 rows)
constexpr uint32_t RT_LBO = AGeo<RT_KC>::LBO;
static_assert(RT_LBO >= (TC_BM + RT_MAX_KW - 1) * 16, "chunk stride must hold the halo rows");

template <int EPI>
__global__ void __launch_bounds__(TC_THREADS) conv_rowtaps_tc(const TcParams P,
                                                              const __grid_constant__ CUtensorMap tm_out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const ConvParams& p = P.c;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int BN = P.BN, kw = p.kw;
  constexpr uint32_t a_part = RT_KC * RT_LBO;
  const uint32_t b_part = RT_KC * BN * 16;
  const uint32_t stage_bytes = 2 * a_part + (uint32_t)kw * 2 * b_part;
  const int smask = P.stages - 1;
  uint8_t* ctrl = smem + P.ctrl_off;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(ctrl);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 8 * TC_STAGES);

  // tile -> (image, output row, 128-pixel segment)
  const int segs = (p.Wo + TC_BM - 1) / TC_BM;
  const int seg = blockIdx.x % segs, bh = blockIdx.x / segs;
  const int h = bh % p.Ho, b = bh / p.Ho;
  const int w0 = seg * TC_BM;
  const int nrows = min(TC_BM, p.Wo - w0);
  const int m0 = (b * p.Ho + h) * p.Wo + w0;
  const int ncb = p.Cin / (RT_KC * 8);
  const int halo_rows = TC_BM + kw - 1;

  // this thread's (halo row, chunk) pairs: pair = tid + pass * 256, chunk = pair % 4, row = pair / 4
  const int chunk = tid & 3;
  const float* px_ptr[3];
  bool px_ok[3];
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int rr = (tid >> 2) + j * 64;
    const int w = w0 - p.pw + rr;
    px_ok[j] = rr < halo_rows && w >= 0 && w < p.W;
    px_ptr[j] = p.in + (((size_t)b * p.H) * p.W + (px_ok[j] ? w : 0)) * p.Cin + chunk * 8;
  }
  const int nt = blockIdx.y;
  const int nvec_b = kw * 2 * RT_KC * BN;
  const uint4* wtile = P.wpk + (size_t)nt * P.nkb * nvec_b;

  auto load_a = [&](int kb, float4 (&v)[6]) {
    const int ky = kb / ncb, cb = kb - ky * ncb;
    const int ih = h + ky - p.ph;
    const bool row_in = ih >= 0 && ih < p.H;
    const size_t off = (size_t)ih * p.W * p.Cin + (size_t)cb * (RT_KC * 8);
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      float4 u = make_float4(0.f, 0.f, 0.f, 0.f), q = u;
      if (row_in && px_ok[j]) {
        u = __ldg(reinterpret_cast<const float4*>(px_ptr[j] + off));
        q = __ldg(reinterpret_cast<const float4*>(px_ptr[j] + off + 4));
      }
      v[2 * j] = u;
      v[2 * j + 1] = q;
    }
  };

  float4 pre[6];
  load_a(0, pre);
  {
    const uint32_t b_dst = smem_u32(smem + 2 * a_part);
    for (int i = tid; i < nvec_b; i += TC_THREADS) cp_async16(b_dst + i * 16, wtile + i);
  }
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < TC_STAGES; ++s) mbar_init(smem_u32(&mbar[s]), 1);
    fence_mbar_init();
  }
  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), (uint32_t)P.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  for (int kb = 0; kb < P.nkb; ++kb) {
    const int s = kb & smask;
    uint8_t* st = smem + s * stage_bytes;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const int rr = (tid >> 2) + j * 64;
      if (rr < halo_rows) {
        float x[8] = {pre[2 * j].x,     pre[2 * j].y,     pre[2 * j].z,     pre[2 * j].w,
                      pre[2 * j + 1].x, pre[2 * j + 1].y, pre[2 * j + 1].z, pre[2 * j + 1].w};
        uint4 hi, lo;
        split8(x, hi, lo);
        const uint32_t off = (uint32_t)chunk * RT_LBO + (uint32_t)rr * 16;
        *reinterpret_cast<uint4*>(st + off) = hi;
        *reinterpret_cast<uint4*>(st + a_part + off) = lo;
      }
    }
    cp_async_wait_all();
    if (kb + 1 < P.nkb) load_a(kb + 1, pre);
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi_s = smem_u32(st), a_lo_s = a_hi_s + a_part;
      const uint32_t b_base = a_hi_s + 2 * a_part;
      const uint32_t b_lbo = (uint32_t)BN * 16;
      for (int kx = 0; kx < kw; ++kx) {
        const uint32_t b_hi_s = b_base + (uint32_t)kx * 2 * b_part, b_lo_s = b_hi_s + b_part;
#pragma unroll
        for (int j = 0; j < RT_KC / 2; ++j) {
          // the kx-th tap of output pixel r reads halo row r + kx: same rows, start address kx * 16 bytes on
          uint64_t ah = make_desc(a_hi_s + 2 * j * RT_LBO + kx * 16, RT_LBO, 128);
          uint64_t al = make_desc(a_lo_s + 2 * j * RT_LBO + kx * 16, RT_LBO, 128);
          uint64_t bh = make_desc(b_hi_s + 2 * j * b_lbo, b_lbo, 128);
          uint64_t bl = make_desc(b_lo_s + 2 * j * b_lbo, b_lbo, 128);
          umma_f16(tmem_base, ah, bh, idesc, (kb | kx | j) ? 1u : 0u);
          umma_f16(tmem_base, ah, bl, idesc, 1u);
          umma_f16(tmem_base, al, bh, idesc, 1u);
        }
      }
      umma_commit(smem_u32(&mbar[s]));
    }
    if (kb + 1 < P.nkb) {
      if (kb >= 1) mbar_wait(smem_u32(&mbar[s ^ 1]), (uint32_t)(((kb - 1) >> 1) & 1));
      const uint32_t b_dst = smem_u32(smem + (s ^ 1) * stage_bytes + 2 * a_part);
      const uint4* b_src = wtile + (size_t)(kb + 1) * nvec_b;
      for (int i = tid; i < nvec_b; i += TC_THREADS) cp_async16(b_dst + i * 16, b_src + i);
    }
  }
  {
    const int last = P.nkb - 1;
    mbar_wait(smem_u32(&mbar[last & smask]), (uint32_t)((P.stages == 2 ? (last >> 1) : last) & 1));
    tc_fence_after();
  }
  const uint32_t lane_base = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
  if (P.tma_out)  // 3-D map [N][Wo][B*Ho]: the box is clipped at the end of the image row
    epi_conv_store_tma<EPI & 7>(p, &tm_out, smem, lane_base, BN, nt * BN, w0, b * p.Ho + h, 1, tid);
  else
    epi_conv_store<EPI & 7>(p, smem, lane_base, BN, nt * BN, m0, nrows, tid);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// per-row combine of the CTC head's tile partials: arg-max (last index on ties) and 1 / sum exp(z - zmax).
// One warp per row: lanes stride the row's n_tiles partials (144 for the 18385-class head: a thread per row walked
// three strided arrays in a dependent chain, 38 us per launch), butterfly-reduce the (max, index) pair, then the
// rescaled sums in a fixed lane order (deterministic).
__global__ void __launch_bounds__(256) ctc_combine_kernel(const float* __restrict__ pmax, const int32_t* __restrict__ pidx,
                                                          const float* __restrict__ psum, size_t rows, int n_tiles,
                                                          int32_t* __restrict__ idx, float* __restrict__ prob) {
  const size_t r = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const float* a = pmax + r * n_tiles;
  const int32_t* ai = pidx + r * n_tiles;
  float mx = -INFINITY;
  int mi = -1;
  for (int t = lane; t < n_tiles; t += 32) {
    const float v = a[t];
    const int vi = ai[t];
    if (v > mx || (v == mx && vi > mi)) mx = v, mi = vi;  // last index on ties (class ranges ascend with t)
  }
  for (int o = 16; o; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi > mi)) mx = ov, mi = oi;
  }
  float tot = 0.0f;
  for (int t = lane; t < n_tiles; t += 32) tot += psum[r * n_tiles + t] * expf(a[t] - mx);
  for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
  if (lane == 0) {
    idx[r] = mi < 0 ? 0 : mi;
    prob[r] = 1.0f / tot;
  }
}

void launch_ctc_combine(oar_ctx* ctx, const float* part_max, const int32_t* part_idx, const float* part_sum, size_t rows,
                        int n_tiles, int32_t* idx, float* prob) {
  Launch l(ctx, "ctc_combine", 0, 12.0 * rows * n_tiles);
  ctc_combine_kernel<<<cdiv((long long)rows, 8), 256, 0, ctx->stream>>>(part_max, part_idx, part_sum, rows, n_tiles, idx,
                                                                       prob);
}

// ---------------------------------------------------------------------------
// host: weight packing and launch
// ---------------------------------------------------------------------------
TcWeights pack_weights(const float* w, int N, int K, bool force_kc4) {
  TcWeights t;
  t.N = N, t.K = K;
  t.n_tiles = (N + TC_MAX_BN - 1) / TC_MAX_BN;
  int per = (N + t.n_tiles - 1) / t.n_tiles;
  t.BN = t.n_tiles > 1 ? (per + 31) / 32 * 32 : std::max(16, (per + 15) / 16 * 16);
  t.KC = (!force_kc4 && K > 32 && K <= 64) ? 8 : 4;  // 64-wide single stage for 32 < K <= 64; otherwise 32-wide double-buffered
  const int TC_KC = t.KC, TC_BK = t.KC * 8;
  t.nkb = (K + TC_BK - 1) / TC_BK;
  size_t halfs = (size_t)t.n_tiles * t.nkb * 2 * TC_KC * t.BN * 8;
  std::vector<__half> buf(halfs, __float2half(0.0f));
  for (int nt = 0; nt < t.n_tiles; ++nt)
    for (int kb = 0; kb < t.nkb; ++kb) {
      __half* stage = buf.data() + ((size_t)nt * t.nkb + kb) * (2 * TC_KC * t.BN * 8);
      for (int kc = 0; kc < TC_KC; ++kc)
        for (int r = 0; r < t.BN; ++r) {
          int n = nt * t.BN + r;
          if (n >= N) continue;
          for (int e = 0; e < 8; ++e) {
            int k = kb * TC_BK + kc * 8 + e;
            if (k >= K) continue;
            float x = w[(size_t)n * K + k];
            __half hi = __float2half_rn(x);
            __half lo = __float2half_rn(x - __half2float(hi));
            stage[((size_t)kc * t.BN + r) * 8 + e] = hi;
            stage[((size_t)(TC_KC + kc) * t.BN + r) * 8 + e] = lo;
          }
        }
    }
  OAR_CUDA(cudaMalloc(&t.packed, halfs * sizeof(__half)));
  OAR_CUDA(cudaMemcpy(t.packed, buf.data(), halfs * sizeof(__half), cudaMemcpyHostToDevice));
  return t;
}

// weights of a stride-1 kh x kw conv for conv_rowtaps_tc: k-block = (ky, 32 input channels), all kw taps per stage
static TcWeights pack_weights_rowtaps(const float* w, int N, int kh, int kw, int Cin, int force_bn = 0) {
  TcWeights t;
  t.N = N, t.K = kh * kw * Cin, t.rowtaps = true, t.kh = kh, t.kw = kw, t.KC = 4;
  // two stages of (A halo rows + kw weight taps) must fit the kernel's 200 KB: 2 * (2 * 4 * RT_LBO + kw * 128 * BN)
  const int bn_cap = std::min(TC_MAX_BN, (int)((200 * 1024 - 2048) / 2 - 2 * RT_KC * (int)RT_LBO) / (kw * 128) / 32 * 32);
  t.n_tiles = (N + bn_cap - 1) / bn_cap;
  int per = (N + t.n_tiles - 1) / t.n_tiles;
  t.BN = t.n_tiles > 1 ? (per + 31) / 32 * 32 : std::max(16, (per + 15) / 16 * 16);
  if (force_bn && t.n_tiles == 1 && force_bn >= t.BN) t.BN = force_bn;  // zero rows behind N
  const int ncb = Cin / 32;
  t.nkb = kh * ncb;
  const size_t part = (size_t)4 * t.BN * 8;  // halfs of one (hi or lo) part
  const size_t stage = (size_t)kw * 2 * part;
  std::vector<__half> buf((size_t)t.n_tiles * t.nkb * stage, __float2half(0.0f));
  for (int nt = 0; nt < t.n_tiles; ++nt)
    for (int ky = 0; ky < kh; ++ky)
      for (int cb = 0; cb < ncb; ++cb) {
        __half* st = buf.data() + ((size_t)nt * t.nkb + ky * ncb + cb) * stage;
        for (int kx = 0; kx < kw; ++kx)
          for (int kc = 0; kc < 4; ++kc)
            for (int r = 0; r < t.BN; ++r) {
              int n = nt * t.BN + r;
              if (n >= N) continue;
              for (int e = 0; e < 8; ++e) {
                int ci = cb * 32 + kc * 8 + e;
                float x = w[(size_t)n * t.K + (size_t)(ky * kw + kx) * Cin + ci];
                __half hi = __float2half_rn(x);
                __half lo = __float2half_rn(x - __half2float(hi));
                st[(size_t)kx * 2 * part + ((size_t)kc * t.BN + r) * 8 + e] = hi;
                st[(size_t)kx * 2 * part + part + ((size_t)kc * t.BN + r) * 8 + e] = lo;
              }
            }
      }
  OAR_CUDA(cudaMalloc(&t.packed, buf.size() * sizeof(__half)));
  OAR_CUDA(cudaMemcpy(t.packed, buf.data(), buf.size() * sizeof(__half), cudaMemcpyHostToDevice));
  return t;
}

void tc_model_init(oar_model* m) {
  TcState* st = new TcState();
  m->tc_state = st;
  std::vector<float> host(m->n_weights);
  OAR_CUDA(cudaMemcpy(host.data(), m->d_weights, m->n_weights * sizeof(float), cudaMemcpyDeviceToHost));
  for (size_t oi = 0; oi < m->ops.size(); ++oi) {
    const OpRec& op = m->ops[oi];
    const float* w0 = host.data() + op.w_off[0];
    switch (op.type) {
      case OP_CONV: {
        const int kh = op.p[0], kw = op.p[1], sh = op.p[2], sw = op.p[3], ph = op.p[4], pw = op.p[5], cin = op.p[6];
        const bool rowtaps = kh * kw > 1 && sh == 1 && sw == 1 && (kh & 1) && (kw & 1) && ph == kh / 2 && pw == kw / 2 &&
                             kw <= RT_MAX_KW && cin % 32 == 0;
        st->w[(int)oi * 2] = rowtaps ? pack_weights_rowtaps(w0, op.p[7], kh, kw, cin) : pack_weights(w0, op.p[7], kh * kw * cin);
        if (rowtaps && kh == 3 && kw == 3 && op.p[7] <= 32 && op.p[7] % 4 == 0) {
          const int n = op.p[7];
          std::vector<float> wf((size_t)3 * n * 3 * cin);  // [kx * n + co][ky][0][ci]
          for (int co = 0; co < n; ++co)
            for (int ky = 0; ky < 3; ++ky)
              for (int kx = 0; kx < 3; ++kx)
                for (int ci = 0; ci < cin; ++ci)
                  wf[((size_t)(kx * n + co) * 3 + ky) * cin + ci] = w0[(size_t)co * 9 * cin + (size_t)(ky * 3 + kx) * cin + ci];
          st->wfold[(int)oi * 2] = pack_weights_rowtaps(wf.data(), 3 * n, 3, 1, cin);
        }
        // the fused depthwise->pointwise kernel (fused_tc.cu) streams 32-channel k-blocks: KC = 4 packing
        if (kh == 1 && kw == 1 && st->w[(int)oi * 2].KC != 4) st->wf[(int)oi * 2] = pack_weights(w0, op.p[7], cin, true);
        break;
      }
      case OP_DWCONV: {
        const int k = op.p[0], c = op.p[6];
        if (op.p[1] != k) break;
        const int nkb = (c + 31) / 32, rows = k * k + 1;
        std::vector<float> pk((size_t)nkb * rows * 32, 0.0f);
        const float* b0 = host.data() + op.w_off[1];
        for (int ch = 0; ch < c; ++ch) {
          float* dst = pk.data() + (size_t)(ch / 32) * rows * 32 + (ch & 31);
          for (int i = 0; i < k * k; ++i) dst[(size_t)i * 32] = w0[(size_t)i * c + ch];
          dst[(size_t)k * k * 32] = b0[ch];
        }
        float* d = nullptr;
        OAR_CUDA(cudaMalloc(&d, pk.size() * sizeof(float)));
        OAR_CUDA(cudaMemcpy(d, pk.data(), pk.size() * sizeof(float), cudaMemcpyHostToDevice));
        st->dwp[(int)oi] = d;
        break;
      }
      case OP_DECONV2:
        st->w[(int)oi * 2] = pack_weights(w0, 4 * op.p[1], op.p[0]);
        break;
      case OP_ATTN:
        st->w[(int)oi * 2] = pack_weights(w0, 3 * op.p[0], op.p[0]);
        st->w[(int)oi * 2 + 1] = pack_weights(host.data() + op.w_off[2], op.p[0], op.p[0]);
        break;
      case OP_CTC_HEAD:
        st->w[(int)oi * 2] = pack_weights(w0, op.p[1], op.p[0]);
        break;
      default:
        break;
    }
  }
}

void tc_model_free(oar_model* m) {
  TcState* st = static_cast<TcState*>(m->tc_state);
  if (!st) return;
  for (auto& kv : st->w) cudaFree(kv.second.packed);
  for (auto& kv : st->wf) cudaFree(kv.second.packed);
  for (auto& kv : st->wfold) cudaFree(kv.second.packed);
  for (auto& kv : st->dwp) cudaFree(kv.second);
  delete st;
  m->tc_state = nullptr;
}

int tc_n_tiles(const oar_model* m, int key) {
  const TcState* st = static_cast<const TcState*>(m->tc_state);
  if (!st) return 0;
  auto it = st->w.find(key);
  // mode-2 partial slots: two column halves per N tile; needs BN % 32 == 0 so each half is whole 16-column chunks
  if (it == st->w.end() || (it->second.BN & 31)) return 0;
  return 2 * it->second.n_tiles;
}

// ---- TMA tensor maps (cuTensorMapEncodeTiled through the runtime's driver entry point: no link-time libcuda)
EncodeTiledFn tmap_encoder() {
  // function-local static: initialised once, thread-safe (contexts on different GPUs launch concurrently)
  static const EncodeTiledFn fn = [] {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) != cudaSuccess || !ptr)
      OAR_FAIL(OAR_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    return (EncodeTiledFn)ptr;
  }();
  return fn;
}

// fp32 output map with 32-column x 128-row boxes, 128-byte swizzle.  rank 2: [N][M]; rank 3: [N][Wo][B*Ho].
static bool make_out_map(CUtensorMap* tm, float* base, int N, int out_ld, long long d1, long long d2, int rank) {
  if ((out_ld & 3) || (((uintptr_t)base) & 15)) return false;  // TMA needs 16-byte aligned base and strides
  cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)d1, (cuuint64_t)d2};
  cuuint64_t strides[2] = {(cuuint64_t)out_ld * 4, (cuuint64_t)out_ld * 4 * (cuuint64_t)d1};
  cuuint32_t box[3] = {32, TC_BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = tmap_encoder()(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, base, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS;
}

bool tc_gemm(oar_model* m, int key, const ConvParams& p, const char* name) {
  TcState* st = static_cast<TcState*>(m->tc_state);
  if (!st) return false;
  auto it = st->w.find(key);
  if (it == st->w.end()) return false;
  const TcWeights& w = it->second;
  if (w.N != p.N || w.K != p.K) OAR_FAIL(OAR_E_MODEL, "tensor-core weights for op key %d do not match its GEMM", key);
  if (p.M <= 0) return true;
  TcParams P;
  P.c = p;
  P.wpk = w.packed;
  P.BN = w.BN, P.nkb = w.nkb, P.n_tiles = w.n_tiles;
  P.stages = w.nkb > 1 ? 2 : 1;
  int cols = 32;
  while (cols < w.BN) cols <<= 1;
  P.tmem_cols = cols;
  const bool pointwise = p.kh == 1 && p.kw == 1 && p.sh == 1 && p.sw == 1 && p.ph == 0 && p.pw == 0;
  const bool aligned = (((uintptr_t)p.in) & 15) == 0;
  if (w.rowtaps) {
    if (!aligned || p.mode != 0 || p.kh != w.kh || p.kw != w.kw || p.sh != 1 || p.sw != 1 || p.Ho != p.H || p.Wo != p.W)
      OAR_FAIL(OAR_E_MODEL, "row-taps weights for op key %d do not match its convolution", key);
    using KernRT = void (*)(const TcParams, const CUtensorMap);
    KernRT krt = nullptr;
    switch (p.act) {
      case ACT_NONE: krt = conv_rowtaps_tc<0>; break;
      case ACT_RELU: krt = conv_rowtaps_tc<1>; break;
      case ACT_HSWISH: krt = conv_rowtaps_tc<2>; break;
      case ACT_SWISH: krt = conv_rowtaps_tc<3>; break;
      case ACT_SIGMOID: krt = conv_rowtaps_tc<4>; break;
      default: return false;
    }
    size_t smem_rt = (size_t)P.stages * (2 * (size_t)RT_KC * RT_LBO + (size_t)w.kw * 2 * RT_KC * w.BN * 16);
    if (smem_rt < (size_t)2 * TC_BM * 128) smem_rt = 2 * TC_BM * 128;  // two TMA staging tiles
    CUtensorMap tm_rt;
    memset(&tm_rt, 0, sizeof(tm_rt));
    P.tma_out = make_out_map(&tm_rt, p.out + p.out_c_off, p.N, p.out_ld, p.Wo, (long long)p.B * p.Ho, 3) ? 1 : 0;
    P.ctrl_off = (uint32_t)smem_rt;
    smem_rt += 64;
    ensure_max_dynamic_smem((const void*)krt, m->ctx->device, 200 * 1024);
    const int segs = (p.Wo + TC_BM - 1) / TC_BM;
    dim3 grid_rt((unsigned)(p.B * p.Ho * segs), w.n_tiles);
    Launch l(m->ctx, name, 2.0 * p.M * p.N * p.K, 4.0 * ((double)p.M * p.K / (p.kh * p.kw)) + 4.0 * (double)p.M * p.N);
    krt<<<grid_rt, TC_THREADS, smem_rt + 1024, m->ctx->stream>>>(P, tm_rt);
    return true;
  }
  int a_mode = AM_SCALAR;
  if (pointwise && (p.Cin % 8) == 0 && aligned)
    a_mode = AM_POINTWISE;
  else if ((p.Cin % 8) == 0 && aligned)
    a_mode = AM_TAPS;
  const int epi = p.mode == 2 ? EPI_CTC : p.mode * 8 + p.act;
  using Kern = void (*)(const TcParams, const CUtensorMap);
  Kern kern = nullptr;
#define TC_PICK(KCV, AM, EP) \
  if (w.KC == KCV && a_mode == AM && epi == EP) kern = conv_gemm_tc<KCV, AM, EP>;
#define TC_PICK_CONV(AM)                                                                                     \
  TC_PICK(4, AM, 0) TC_PICK(4, AM, 1) TC_PICK(4, AM, 2) TC_PICK(4, AM, 3) TC_PICK(4, AM, 4) TC_PICK(8, AM, 0) \
      TC_PICK(8, AM, 1) TC_PICK(8, AM, 2) TC_PICK(8, AM, 3) TC_PICK(8, AM, 4)
  TC_PICK_CONV(AM_POINTWISE)
  TC_PICK_CONV(AM_TAPS)
  TC_PICK_CONV(AM_SCALAR)
  TC_PICK(4, AM_POINTWISE, 8) TC_PICK(4, AM_POINTWISE, 9) TC_PICK(4, AM_POINTWISE, 12)
  TC_PICK(8, AM_POINTWISE, 8) TC_PICK(8, AM_POINTWISE, 9) TC_PICK(8, AM_POINTWISE, 12)
  TC_PICK(4, AM_POINTWISE, EPI_CTC) TC_PICK(8, AM_POINTWISE, EPI_CTC)
#undef TC_PICK_CONV
#undef TC_PICK
  if (!kern) return false;  // combination not instantiated: the caller runs the SIMT kernel
  if (p.mode == 2 && (w.BN & 31)) return false;
  const size_t a_part_bytes = w.KC == 8 ? AGeo<8>::PART : AGeo<4>::PART;
  size_t smem = (size_t)P.stages * (2 * a_part_bytes + 2 * (size_t)w.KC * w.BN * 16);
  if (smem < (size_t)2 * TC_BM * 128) smem = 2 * TC_BM * 128;  // the conv epilogue stages outputs through shared memory
  CUtensorMap tm;
  memset(&tm, 0, sizeof(tm));
  P.tma_out = 0;
  // single-tile layers may have any BN; multi-tile layers need whole 32-column boxes inside a tile
  if (p.mode == 0 && (w.n_tiles == 1 || (w.BN & 31) == 0))
    P.tma_out = make_out_map(&tm, p.out + p.out_c_off, p.N, p.out_ld, p.M, 1, 2) ? 1 : 0;
  P.ctrl_off = (uint32_t)smem;
  smem += 64;
  ensure_max_dynamic_smem((const void*)kern, m->ctx->device, 200 * 1024);
  dim3 grid(cdiv(p.M, TC_BM), w.n_tiles);
  double out_bytes = p.mode == 2 ? 24.0 * p.M * w.n_tiles : 4.0 * (double)p.M * p.N;
  Launch l(m->ctx, name, 2.0 * p.M * p.N * p.K, 4.0 * ((double)p.M * p.K / (p.kh * p.kw)) + out_bytes);
  kern<<<grid, TC_THREADS, smem + 1024, m->ctx->stream>>>(P, tm);
  return true;
}

}  // namespace oar
