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
// Kernel shape (one CTA = one 128-row x BN-column output tile, 128 threads):
//   A (activations): thread r gathers its pixel row's k-chunk from the fp32 NHWC tensor (im2col on the fly),
//       splits to hi/lo fp16 and stores 16-byte core-matrix rows into shared memory in the canonical K-major,
//       no-swizzle UMMA layout [k-chunk][row][8 halfs]  (LBO = 128 rows * 16 B, SBO = 128 B);
//   B (weights): pre-split and pre-packed on the host at model load in exactly the shared-memory layout, so a stage
//       is one linear 16-byte-vector copy;
//   two shared-memory stages; tcgen05.commit -> mbarrier releases a stage when its MMAs have read it;
//   D: fp32 accumulators in TMEM (128 lanes x BN columns); epilogue = tcgen05.ld 32x32b, + bias, activation,
//       post-affine, then NHWC stores (conv), 2x2 scatter (transposed conv) or an online softmax/arg-max reduction
//       (CTC head: the [B,T,V] logits never reach HBM).
#include <cuda_fp16.h>

#include <map>

#include "engine.cuh"

namespace oar {

constexpr int TC_BM = 128;        // UMMA M
constexpr int TC_BK = 32;         // k elements per stage
constexpr int TC_KC = TC_BK / 8;  // 16-byte k-chunks per stage
constexpr int TC_STAGES = 2;
constexpr int TC_MAX_BN = 256;

struct TcWeights {
  uint4* packed = nullptr;  // [n_tile][k_block][hi|lo][k-chunk][BN rows][8 halfs]
  int N = 0, K = 0, BN = 0, n_tiles = 0, nkb = 0;
};

struct TcState {
  std::map<int, TcWeights> w;
};

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// D[tmem] (+)= A[smem] * B[smem]^T, one elected thread issues
__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, "
      "[%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
// start address [0,14), leading (k-chunk) byte offset [16,30), stride (8-row group) byte offset [32,46), all >> 4
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46);
}

__device__ __forceinline__ float tc_act(float v, int act) {
  switch (act) {
    case ACT_RELU:
      return fmaxf(v, 0.0f);
    case ACT_HSWISH:
      return v * fminf(fmaxf(v + 3.0f, 0.0f), 6.0f) / 6.0f;
    case ACT_SWISH:
      return v / (1.0f + expf(-v));
    case ACT_SIGMOID:
      return 1.0f / (1.0f + expf(-v));
    case ACT_HSIGMOID:
      return fminf(fmaxf(v / 6.0f + 0.5f, 0.0f), 1.0f);
    default:
      return v;
  }
}

// split 8 floats into hi / lo fp16 vectors (x ~= hi + lo, |x - hi - lo| <= 2^-22 |x|)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  __half2 h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    __half a = __float2half_rn(x[2 * i]), b = __float2half_rn(x[2 * i + 1]);
    h[i] = __halves2half2(a, b);
    l[i] = __halves2half2(__float2half_rn(x[2 * i] - __half2float(a)), __float2half_rn(x[2 * i + 1] - __half2float(b)));
  }
  hi = make_uint4(*reinterpret_cast<uint32_t*>(&h[0]), *reinterpret_cast<uint32_t*>(&h[1]),
                  *reinterpret_cast<uint32_t*>(&h[2]), *reinterpret_cast<uint32_t*>(&h[3]));
  lo = make_uint4(*reinterpret_cast<uint32_t*>(&l[0]), *reinterpret_cast<uint32_t*>(&l[1]),
                  *reinterpret_cast<uint32_t*>(&l[2]), *reinterpret_cast<uint32_t*>(&l[3]));
}

struct TcParams {
  ConvParams c;
  const uint4* wpk;
  int BN, nkb, n_tiles, tmem_cols;
  int a_mode;  // 0 pointwise (row pointer), 1 conv with Cin % 8 == 0 (vector taps), 2 generic scalar gather
};

__global__ void __launch_bounds__(128) conv_gemm_tc(const TcParams P) {
  extern __shared__ __align__(128) uint8_t smem[];
  const ConvParams& p = P.c;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int BN = P.BN;
  const uint32_t a_part = TC_KC * TC_BM * 16;  // bytes of one A part (hi or lo) per stage
  const uint32_t b_part = TC_KC * BN * 16;
  const uint32_t stage_bytes = 2 * a_part + 2 * b_part;
  uint8_t* ctrl = smem + TC_STAGES * stage_bytes;
  uint64_t* mbar = reinterpret_cast<uint64_t*>(ctrl);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ctrl + 8 * TC_STAGES);

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

  // this thread's output row (pixel)
  const int m = blockIdx.x * TC_BM + tid;
  const bool row_ok = m < p.M;
  int ab = 0, aho = 0, awo = 0;
  if (row_ok && P.a_mode != 0) {
    ab = m / (p.Ho * p.Wo);
    int r = m - ab * p.Ho * p.Wo;
    aho = r / p.Wo;
    awo = r - aho * p.Wo;
  }
  const float* arow = p.in + (size_t)m * p.Cin;  // a_mode 0 only
  const int nt = blockIdx.y;
  const uint4* wtile = P.wpk + (size_t)nt * P.nkb * (2 * TC_KC * BN);
  // instruction descriptor: D fp32, A/B fp16, both K-major, N = BN, M = 128
  const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(TC_BM >> 4) << 24);

  for (int kb = 0; kb < P.nkb; ++kb) {
    const int s = kb & 1;
    uint8_t* st = smem + s * stage_bytes;
    if (kb >= TC_STAGES) mbar_wait(smem_u32(&mbar[s]), (uint32_t)(((kb >> 1) - 1) & 1));
    // ---- A: gather, split, store
    uint4* a_hi = reinterpret_cast<uint4*>(st);
    uint4* a_lo = reinterpret_cast<uint4*>(st + a_part);
#pragma unroll
    for (int kc = 0; kc < TC_KC; ++kc) {
      const int k0 = kb * TC_BK + kc * 8;
      float x[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (row_ok && k0 < p.K) {
        if (P.a_mode == 0) {
          float4 u = __ldg(reinterpret_cast<const float4*>(arow + k0));
          float4 v = __ldg(reinterpret_cast<const float4*>(arow + k0 + 4));
          x[0] = u.x, x[1] = u.y, x[2] = u.z, x[3] = u.w, x[4] = v.x, x[5] = v.y, x[6] = v.z, x[7] = v.w;
        } else if (P.a_mode == 1) {
          int tap = k0 / p.Cin, ci = k0 - tap * p.Cin;
          int ky = tap / p.kw, kx = tap - ky * p.kw;
          int ih = aho * p.sh - p.ph + ky, iw = awo * p.sw - p.pw + kx;
          if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W) {
            const float* src = p.in + (((size_t)ab * p.H + ih) * p.W + iw) * p.Cin + ci;
            float4 u = __ldg(reinterpret_cast<const float4*>(src));
            float4 v = __ldg(reinterpret_cast<const float4*>(src + 4));
            x[0] = u.x, x[1] = u.y, x[2] = u.z, x[3] = u.w, x[4] = v.x, x[5] = v.y, x[6] = v.z, x[7] = v.w;
          }
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            int k = k0 + i;
            if (k < p.K) {
              int tap = k / p.Cin, ci = k - tap * p.Cin;
              int ky = tap / p.kw, kx = tap - ky * p.kw;
              int ih = aho * p.sh - p.ph + ky, iw = awo * p.sw - p.pw + kx;
              if (ih >= 0 && ih < p.H && iw >= 0 && iw < p.W)
                x[i] = __ldg(p.in + (((size_t)ab * p.H + ih) * p.W + iw) * p.Cin + ci);
            }
          }
        }
      }
      uint4 hi, lo;
      split8(x, hi, lo);
      a_hi[kc * TC_BM + tid] = hi;
      a_lo[kc * TC_BM + tid] = lo;
    }
    // ---- B: linear copy of the pre-packed stage (hi then lo)
    {
      uint4* b_dst = reinterpret_cast<uint4*>(st + 2 * a_part);
      const uint4* b_src = wtile + (size_t)kb * (2 * TC_KC * BN);
      const int nvec = 2 * TC_KC * BN;
      for (int i = tid; i < nvec; i += 128) b_dst[i] = __ldg(b_src + i);
    }
    fence_proxy_async_smem();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t a_hi_s = smem_u32(st), a_lo_s = a_hi_s + a_part;
      const uint32_t b_hi_s = a_hi_s + 2 * a_part, b_lo_s = b_hi_s + b_part;
      const uint32_t a_lbo = TC_BM * 16, b_lbo = (uint32_t)BN * 16;
#pragma unroll
      for (int j = 0; j < TC_BK / 16; ++j) {
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
  }
  // all MMAs done when the last commit lands (a commit tracks every prior tcgen05 op of the issuing thread)
  {
    const int last = P.nkb - 1;
    mbar_wait(smem_u32(&mbar[last & 1]), (uint32_t)((last >> 1) & 1));
    tc_fence_after();
  }

  // ---- epilogue: thread = row, 16 columns at a time out of TMEM
  const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
  const int n_base = nt * BN;
  if (p.mode == 2) {
    float mx = -INFINITY, sum = 0.0f;
    int mi = 0;
    for (int c0 = 0; c0 < BN; c0 += 16) {
      if (n_base + c0 >= p.N) break;  // uniform across the CTA
      float v[16];
      tmem_ld16(lane_base + c0, v);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        int n = n_base + c0 + i;
        if (n < p.N) {
          float z = v[i] + __ldg(p.bias + n);
          if (z > mx) {
            sum = sum * expf(mx - z) + 1.0f;
            mx = z;
            mi = n;
          } else {
            sum += expf(z - mx);
            if (z == mx) mi = n;  // LAST maximal index wins (simd.rs:194-204)
          }
        }
      }
    }
    if (row_ok) {
      size_t o = (size_t)m * P.n_tiles + nt;
      p.part_max[o] = mx;
      p.part_idx[o] = mi;
      p.part_sum[o] = sum;
    }
  } else {
    int ob = 0, oy = 0, ox = 0;
    if (p.mode == 1 && row_ok) {
      ob = m / (p.Ho * p.Wo);
      int r = m - ob * p.Ho * p.Wo;
      oy = r / p.Wo;
      ox = r - oy * p.Wo;
    }
    const bool vec_ok = p.mode == 0 && ((p.out_ld & 3) == 0) && ((p.out_c_off & 3) == 0) && ((n_base & 3) == 0);
    float* orow = p.out + (size_t)m * p.out_ld + p.out_c_off;
    for (int c0 = 0; c0 < BN; c0 += 16) {
      if (n_base + c0 >= p.N) break;
      float v[16];
      tmem_ld16(lane_base + c0, v);  // warp-collective: every lane takes part, stores are predicated below
      if (!row_ok) continue;
      if (p.mode == 0) {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int n = n_base + c0 + i;
          if (n < p.N) v[i] = tc_act(v[i] + __ldg(p.bias + n), p.act) * p.post_scale + p.post_bias;
        }
        if (vec_ok && n_base + c0 + 16 <= p.N) {
#pragma unroll
          for (int i = 0; i < 16; i += 4)
            *reinterpret_cast<float4*>(orow + n_base + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i)
            if (n_base + c0 + i < p.N) orow[n_base + c0 + i] = v[i];
        }
      } else {
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          int n = n_base + c0 + i;
          if (n < p.N) {
            int q = n / p.cout, co = n - q * p.cout;
            int dy = q >> 1, dx = q & 1;
            float r = tc_act(v[i] + __ldg(p.bias + co), p.act) * p.post_scale + p.post_bias;
            p.out[(((size_t)ob * (2 * p.Ho) + 2 * oy + dy) * (2 * p.Wo) + 2 * ox + dx) * p.cout + co] = r;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, (uint32_t)P.tmem_cols);
}

// per-row combine of the CTC head's tile partials: arg-max (last index on ties) and 1 / sum exp(z - zmax)
__global__ void ctc_combine_kernel(const float* __restrict__ pmax, const int32_t* __restrict__ pidx,
                                   const float* __restrict__ psum, size_t rows, int n_tiles, int32_t* __restrict__ idx,
                                   float* __restrict__ prob) {
  size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* a = pmax + r * n_tiles;
  float mx = a[0];
  int mi = pidx[r * n_tiles];
  for (int t = 1; t < n_tiles; ++t)
    if (a[t] >= mx) mx = a[t], mi = pidx[r * n_tiles + t];  // tiles ascend in class index
  float tot = 0.0f;
  for (int t = 0; t < n_tiles; ++t) tot += psum[r * n_tiles + t] * expf(a[t] - mx);
  idx[r] = mi;
  prob[r] = 1.0f / tot;
}

void launch_ctc_combine(oar_ctx* ctx, const float* part_max, const int32_t* part_idx, const float* part_sum, size_t rows,
                        int n_tiles, int32_t* idx, float* prob) {
  Launch l(ctx, "ctc_combine", 0, 12.0 * rows * n_tiles);
  ctc_combine_kernel<<<cdiv((long long)rows, 128), 128, 0, ctx->stream>>>(part_max, part_idx, part_sum, rows, n_tiles,
                                                                         idx, prob);
}

// ---------------------------------------------------------------------------
// host: weight packing and launch
// ---------------------------------------------------------------------------
static TcWeights pack_weights(const float* w, int N, int K) {
  TcWeights t;
  t.N = N, t.K = K;
  t.n_tiles = (N + TC_MAX_BN - 1) / TC_MAX_BN;
  int per = (N + t.n_tiles - 1) / t.n_tiles;
  t.BN = std::max(16, (per + 15) / 16 * 16);
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

void tc_model_init(oar_model* m) {
  TcState* st = new TcState();
  m->tc_state = st;
  std::vector<float> host(m->n_weights);
  OAR_CUDA(cudaMemcpy(host.data(), m->d_weights, m->n_weights * sizeof(float), cudaMemcpyDeviceToHost));
  for (size_t oi = 0; oi < m->ops.size(); ++oi) {
    const OpRec& op = m->ops[oi];
    const float* w0 = host.data() + op.w_off[0];
    switch (op.type) {
      case OP_CONV:
        st->w[(int)oi * 2] = pack_weights(w0, op.p[7], op.p[0] * op.p[1] * op.p[6]);
        break;
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
  delete st;
  m->tc_state = nullptr;
}

int tc_n_tiles(const oar_model* m, int key) {
  const TcState* st = static_cast<const TcState*>(m->tc_state);
  if (!st) return 0;
  auto it = st->w.find(key);
  return it == st->w.end() ? 0 : it->second.n_tiles;
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
  int cols = 32;
  while (cols < w.BN) cols <<= 1;
  P.tmem_cols = cols;
  const bool pointwise = p.kh == 1 && p.kw == 1 && p.sh == 1 && p.sw == 1 && p.ph == 0 && p.pw == 0;
  const bool aligned = (((uintptr_t)p.in) & 15) == 0;
  if (pointwise && (p.Cin % 8) == 0 && aligned)
    P.a_mode = 0;
  else if ((p.Cin % 8) == 0 && aligned)
    P.a_mode = 1;
  else
    P.a_mode = 2;
  size_t smem = (size_t)TC_STAGES * (2 * TC_KC * TC_BM * 16 + 2 * TC_KC * w.BN * 16) + 64;
  static bool attr_set[64] = {false};
  if (!attr_set[m->ctx->device & 63]) {  // the attribute is per device
    OAR_CUDA(cudaFuncSetAttribute(conv_gemm_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set[m->ctx->device & 63] = true;
  }
  dim3 grid(cdiv(p.M, TC_BM), w.n_tiles);
  double out_bytes = p.mode == 2 ? 12.0 * p.M * w.n_tiles : 4.0 * (double)p.M * p.N;
  Launch l(m->ctx, name, 2.0 * p.M * p.N * p.K, 4.0 * ((double)p.M * p.K / (p.kh * p.kw)) + out_bytes);
  conv_gemm_tc<<<grid, 128, smem, m->ctx->stream>>>(P);
  return true;
}

}  // namespace oar
