// fused_simt.cu -- two SIMT fusions at the thin ends of the detector / recogniser graphs, where the layers are far too
// narrow for a tensor-core tile and the un-fused form spends its time moving fp32 intermediates through HBM.
//
//   stem_u8_kernel      NormalizeImage (normalization.rs:429-482, simd.rs:161-187) or normalize_crnn_chw_into
//                       (simd.rs:248-308) + the 3x3 stride-2 stem convolution (3 -> 16 channels) in ONE pass over the
//                       u8 page / crop: the normalised fp32 [B,H,W,3] tensor (12 B per pixel written, then read) never
//                       exists.  The normalisation arithmetic is the reference's, operation by operation
//                       (separate multiply and add; true divisions for the CRNN form), so the values entering the
//                       convolution are bit-identical to the stand-alone normalise kernels (prepost.cu).
//   deconv_pair_kernel  DBHead tail: ConvTranspose 2x2/s2 (24 -> 24) + ReLU + ConvTranspose 2x2/s2 (24 -> 1) + Sigmoid
//                       (the reference runs them inside ONNX Runtime, ort_infer_execution.rs:178).  Each input pixel
//                       owns a 4x4 output patch, so the 24-channel 480x480 intermediate (708 MB per 32 pages, written
//                       and read back) stays in registers: 96 B read + 64 B written per input pixel.
#include "engine.cuh"
#include "prepost.cuh"

namespace oar {

__device__ __forceinline__ float act_simt(float v, int act) {
  switch (act) {
    case ACT_RELU: return fmaxf(v, 0.0f);
    case ACT_HSWISH: return v * fminf(fmaxf(v + 3.0f, 0.0f), 6.0f) / 6.0f;
    case ACT_SWISH: return v / (1.0f + expf(-v));
    case ACT_SIGMOID: return 1.0f / (1.0f + expf(-v));
    case ACT_HSIGMOID: return fminf(fmaxf(v / 6.0f + 0.5f, 0.0f), 1.0f);
    case ACT_GELU: return 0.5f * v * (1.0f + erff(v * 0.70710678118654752f));
    default: return v;
  }
}

// ---------------------------------------------------------------------------
// u8 -> normalise -> 3x3 stride-2 pad-1 convolution, 3 -> COUT channels.
// Persistent CTAs over (image, output row, segment of 32 * PXT output pixels) items.  The three input rows an item
// needs are normalised ONCE into shared memory ([ky][2 * 32 * PXT + 1 px][3 ch] floats; taps outside the image -- and,
// for crops, right of the resized width rw: the tensor's zero padding -- are stored as 0, which leaves an FMA
// accumulator unchanged).  The CRNN form costs two IEEE divisions per value, so its 256 possible results come from a
// table built once per CTA with exactly that arithmetic; NormalizeImage is one multiply and one add, done in place.
// Thread = (pixel slot, channel quad) x PXT pixels 32 apart: four lanes share a pixel, so a warp's store instruction
// writes 8 pixels x 64 B = one contiguous 512-byte run, a weight read feeds PXT pixels, and the 9 floats of a kernel
// row sit contiguously at 6 * pixel (conflict-free across the pixels of a warp).
// Accumulation per output: bias, then (ky, kx, ci) ascending with FMA, exactly as stem_conv_kernel (engine.cu).
// (Tried and measured slower on B200: float2-paired staging for FFMA2 without register moves -- 4.5 M bank conflicts.
// Tried and measured equal, round 2: the next item's bytes prefetched as aligned words into registers during the compute
// phase and normalised out of shared memory -- 0.83 vs 0.79 ms per step for the two stems: the kernel is bound by the
// ~60 instructions per staged tap and the ~26 per output, not by the latency of its loads.)
// ---------------------------------------------------------------------------
template <int COUT, int PXT>
__global__ void __launch_bounds__(128) stem_u8_kernel(const U8Input S, const float* __restrict__ w,
                                                      const float* __restrict__ bias, float* __restrict__ out, int Ho,
                                                      int Wo, int segs, int n_items, int act, float ps, float pb) {
  static_assert(COUT == 16, "thread = pixel x channel quad");
  constexpr int PX = 32 * PXT, COLS = 2 * PX + 1, ROW = COLS * 3 + 1;  // ROW even: rows stay 8-byte aligned
  __shared__ __align__(16) float ws[27 * COUT];
  __shared__ __align__(16) float xs[3][ROW];
  __shared__ float lut[256];
  for (int i = threadIdx.x; i < 27 * COUT; i += blockDim.x) {
    const int k = i / COUT, co = i - k * COUT;
    ws[i] = w[(size_t)co * 27 + k];
  }
  if (S.mode != 0)  // ((v / 255 - 0.5) / 0.5), simd.rs:248-308
    for (int i = threadIdx.x; i < 256; i += blockDim.x)
      lut[i] = __fdiv_rn(__fsub_rn(__fdiv_rn((float)i, 255.0f), 0.5f), 0.5f);
  // output channel c reads byte sc[c] of its pixel (NormalizeImage: the configured order; CRNN: BGR)
  const int sc0 = S.mode == 0 ? S.src[0] : 2, sc1 = S.mode == 0 ? S.src[1] : 1, sc2 = S.mode == 0 ? S.src[2] : 0;
  const int q = threadIdx.x & 3, pl = threadIdx.x >> 2;
  const float4 bq = __ldg(reinterpret_cast<const float4*>(bias) + q);
  const bool plain = act == ACT_NONE && ps == 1.0f && pb == 0.0f;
  __syncthreads();
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int seg = item % segs, r = item / segs;
    const int ho = r % Ho, b = r / Ho;
    const int wo_base = seg * PX;
    const uint8_t* base;
    int wv;
    if (S.mode == 0) {
      base = S.table[b];
      wv = S.W;
    } else {
      const CrnnJob j = S.jobs[b];
      base = j.src;
      wv = j.rw;
    }
    const size_t stride = (size_t)wv * 3;
    const int ix0 = 2 * wo_base - 1;
    for (int e = threadIdx.x; e < 3 * COLS; e += blockDim.x) {
      const int ky = e / COLS, px = e - ky * COLS;
      const int iy = 2 * ho - 1 + ky, ix = ix0 + px;
      float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
      if (iy >= 0 && iy < S.H && ix >= 0 && ix < wv) {
        const uint8_t* s = base + (size_t)iy * stride + (size_t)ix * 3;
        const int b0 = __ldg(s), b1 = __ldg(s + 1), b2 = __ldg(s + 2);
        const int u0 = sc0 == 0 ? b0 : (sc0 == 1 ? b1 : b2), u1 = sc1 == 0 ? b0 : (sc1 == 1 ? b1 : b2),
                  u2 = sc2 == 0 ? b0 : (sc2 == 1 ? b1 : b2);
        if (S.mode == 0) {
          v0 = __fadd_rn(__fmul_rn((float)u0, S.a[0]), S.b[0]);
          v1 = __fadd_rn(__fmul_rn((float)u1, S.a[1]), S.b[1]);
          v2 = __fadd_rn(__fmul_rn((float)u2, S.a[2]), S.b[2]);
        } else {
          v0 = lut[u0], v1 = lut[u1], v2 = lut[u2];
        }
      }
      float* d = &xs[ky][px * 3];
      d[0] = v0, d[1] = v1, d[2] = v2;
    }
    __syncthreads();
    float2 acc[PXT][2];
#pragma unroll
    for (int pp = 0; pp < PXT; ++pp) acc[pp][0] = make_float2(bq.x, bq.y), acc[pp][1] = make_float2(bq.z, bq.w);
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
      float x[PXT][10];
#pragma unroll
      for (int pp = 0; pp < PXT; ++pp) {
        const float* xp = &xs[ky][6 * (pl + 32 * pp)];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 v = *reinterpret_cast<const float2*>(xp + 2 * t);
          x[pp][2 * t] = v.x, x[pp][2 * t + 1] = v.y;
        }
        x[pp][8] = xp[8];
      }
#pragma unroll
      for (int t = 0; t < 9; ++t) {
        const float4 w4 = *reinterpret_cast<const float4*>(ws + (ky * 9 + t) * COUT + 4 * q);
        const float2 wa = make_float2(w4.x, w4.y), wb = make_float2(w4.z, w4.w);
#pragma unroll
        for (int pp = 0; pp < PXT; ++pp) {
          const float2 xx = make_float2(x[pp][t], x[pp][t]);
          acc[pp][0] = __ffma2_rn(xx, wa, acc[pp][0]);
          acc[pp][1] = __ffma2_rn(xx, wb, acc[pp][1]);
        }
      }
    }
#pragma unroll
    for (int pp = 0; pp < PXT; ++pp) {
      const int wo = wo_base + pl + 32 * pp;
      if (wo >= Wo) continue;
      float4 v = make_float4(acc[pp][0].x, acc[pp][0].y, acc[pp][1].x, acc[pp][1].y);
      if (!plain) {
        v.x = act_simt(v.x, act) * ps + pb, v.y = act_simt(v.y, act) * ps + pb;
        v.z = act_simt(v.z, act) * ps + pb, v.w = act_simt(v.w, act) * ps + pb;
      }
      reinterpret_cast<float4*>(out + (((size_t)b * Ho + ho) * Wo + wo) * COUT)[q] = v;
    }
    __syncthreads();  // the next item's staging overwrites xs
  }
}

bool launch_stem_u8(oar_ctx* ctx, const U8Input& S, const OpRec& op, const float* w, const float* bias, float* out, int Ho,
                    int Wo) {
  // the one shape both networks (and the classifier) start with: 3x3 stride 2 pad 1, 3 -> 16 channels
  if (op.type != OP_CONV || op.p[0] != 3 || op.p[1] != 3 || op.p[2] != 2 || op.p[3] != 2 || op.p[4] != 1 || op.p[5] != 1 ||
      op.p[6] != 3 || op.p[7] != 16 || op.p[11] != 0 || (((uintptr_t)out) & 15))
    return false;
  // 128- or 64-pixel row segments, whichever wastes fewer pixel slots on this width
  const bool wide = cdiv(Wo, 128) * 128 <= cdiv(Wo, 64) * 64;
  const int segs = cdiv(Wo, wide ? 128 : 64);
  const long long items = (long long)S.B * Ho * segs;
  if (items <= 0 || items > 0x7fffffffLL) return false;
  const double px_in = (double)S.B * S.H * S.W, px_out = (double)S.B * Ho * Wo;
  Launch l(ctx, "stem_u8", 2.0 * px_out * 16 * 27, 3.0 * px_in + 4.0 * 16 * px_out);
  const int grid = (int)std::min<long long>(items, (long long)ctx->sm_count * 12);
  if (wide)
    stem_u8_kernel<16, 4><<<grid, 128, 0, ctx->stream>>>(S, w, bias, out, Ho, Wo, segs, (int)items, op.p[8], op.f[0], op.f[1]);
  else
    stem_u8_kernel<16, 2><<<grid, 128, 0, ctx->stream>>>(S, w, bias, out, Ho, Wo, segs, (int)items, op.p[8], op.f[0], op.f[1]);
  OAR_CUDA(cudaGetLastError());
  return true;
}

// ---------------------------------------------------------------------------
// ConvTranspose 2x2/s2 (CIN -> CMID) + act1 + ConvTranspose 2x2/s2 (CMID -> 1) + act2.
// Weights as OP_DECONV2 stores them: w1 [dy][dx][CMID][CIN], w2 [ey][ex][1][CMID].
// Thread = (two input pixels 64 apart, dx): for dy = 0, 1 it produces the CMID intermediate values of position (dy, dx)
// of both pixels (one shared-memory weight read feeds both) and from them the 2x2 final outputs (ey, ex), i.e. columns
// 4x + 2dx + {0, 1} of rows 4y + 2dy + {0, 1}.  Consecutive lanes hold consecutive (x, dx), so every store instruction
// of a warp writes one contiguous 256-byte run of an output row.  Persistent over (image, row, 128-pixel segment)
// items: the 9 KB of re-laid-out weights are staged once per CTA, not once per 64 pixels.
// ---------------------------------------------------------------------------
template <int CIN, int CMID>
__global__ void __launch_bounds__(128) deconv_pair_kernel(const float* __restrict__ in, const float* __restrict__ w1,
                                                          const float* __restrict__ b1, const float* __restrict__ w2,
                                                          const float* __restrict__ b2, float* __restrict__ out, int H,
                                                          int W, int n_items, int segs, int act1, int act2) {
  // [dy][dx][ci][co]; each (dy, dx) block is padded by 4 floats: the lanes of a warp read at TWO addresses (dx = 0 / 1),
  // CIN * CMID floats apart = the same banks -- ncu: 45 % of the kernel's shared-memory wavefronts were bank conflicts
  constexpr int W1Q = CIN * CMID + 4;
  __shared__ __align__(16) float w1s[4 * W1Q];
  __shared__ __align__(16) float b1s[CMID];
  __shared__ __align__(16) float2 w2s[2 * CMID];  // [ey][co] -> (ex = 0, ex = 1)
  for (int i = threadIdx.x; i < 4 * CIN * CMID; i += blockDim.x) {
    const int co = i % CMID, ci = (i / CMID) % CIN, q = i / (CMID * CIN);
    w1s[q * W1Q + ci * CMID + co] = w1[((size_t)q * CMID + co) * CIN + ci];
  }
  for (int i = threadIdx.x; i < CMID; i += blockDim.x) b1s[i] = b1[i];
  for (int i = threadIdx.x; i < 2 * CMID; i += blockDim.x) {
    const int ey = i / CMID, co = i - ey * CMID;
    w2s[i] = make_float2(w2[(ey * 2 + 0) * CMID + co], w2[(ey * 2 + 1) * CMID + co]);
  }
  __syncthreads();
  const int dx = threadIdx.x & 1, xl = threadIdx.x >> 1;
  const float bias2 = __ldg(b2);
  const size_t OW = (size_t)4 * W;
  for (int item = blockIdx.x; item < n_items; item += gridDim.x) {
    const int seg = item % segs, r = item / segs;
    const int y = r % H, b = r / H;
    const int x0 = seg * 128 + xl;
    if (x0 >= W) continue;
    const bool two = x0 + 64 < W;
    float xin[2][CIN];
#pragma unroll
    for (int pp = 0; pp < 2; ++pp) {
      const float4* src = reinterpret_cast<const float4*>(in + (((size_t)b * H + y) * W + x0 + (pp && two ? 64 : 0)) * CIN);
#pragma unroll
      for (int q = 0; q < CIN / 4; ++q) {
        const float4 v = __ldg(src + q);
        xin[pp][4 * q] = v.x, xin[pp][4 * q + 1] = v.y, xin[pp][4 * q + 2] = v.z, xin[pp][4 * q + 3] = v.w;
      }
    }
    float* obase = out + ((size_t)b * 4 * H + 4 * y) * OW + 4 * x0 + 2 * dx;
#pragma unroll
    for (int dy = 0; dy < 2; ++dy) {
      float2 h[2][CMID / 2];
#pragma unroll
      for (int q = 0; q < CMID / 2; ++q) h[0][q] = h[1][q] = make_float2(b1s[2 * q], b1s[2 * q + 1]);
      const float4* wr = reinterpret_cast<const float4*>(w1s + (size_t)(dy * 2 + dx) * W1Q);
#pragma unroll
      for (int ci = 0; ci < CIN; ++ci) {
        const float2 xa = make_float2(xin[0][ci], xin[0][ci]), xb = make_float2(xin[1][ci], xin[1][ci]);
#pragma unroll
        for (int q = 0; q < CMID / 4; ++q) {
          const float4 w4 = wr[ci * (CMID / 4) + q];
          const float2 wa = make_float2(w4.x, w4.y), wb = make_float2(w4.z, w4.w);
          h[0][2 * q] = __ffma2_rn(xa, wa, h[0][2 * q]);
          h[0][2 * q + 1] = __ffma2_rn(xa, wb, h[0][2 * q + 1]);
          h[1][2 * q] = __ffma2_rn(xb, wa, h[1][2 * q]);
          h[1][2 * q + 1] = __ffma2_rn(xb, wb, h[1][2 * q + 1]);
        }
      }
#pragma unroll
      for (int pp = 0; pp < 2; ++pp)
#pragma unroll
        for (int q = 0; q < CMID / 2; ++q) h[pp][q].x = act_simt(h[pp][q].x, act1), h[pp][q].y = act_simt(h[pp][q].y, act1);
#pragma unroll
      for (int ey = 0; ey < 2; ++ey) {
        float2 o0 = make_float2(bias2, bias2), o1 = o0;  // (ex = 0, ex = 1) of pixel 0 / pixel 1
#pragma unroll
        for (int q = 0; q < CMID / 2; ++q) {
          const float2 wx = w2s[ey * CMID + 2 * q], wy = w2s[ey * CMID + 2 * q + 1];
          o0 = __ffma2_rn(make_float2(h[0][q].x, h[0][q].x), wx, o0);
          o0 = __ffma2_rn(make_float2(h[0][q].y, h[0][q].y), wy, o0);
          o1 = __ffma2_rn(make_float2(h[1][q].x, h[1][q].x), wx, o1);
          o1 = __ffma2_rn(make_float2(h[1][q].y, h[1][q].y), wy, o1);
        }
        float* orow = obase + (size_t)(2 * dy + ey) * OW;
        __stcs(reinterpret_cast<float2*>(orow), make_float2(act_simt(o0.x, act2), act_simt(o0.y, act2)));
        if (two) __stcs(reinterpret_cast<float2*>(orow + 256), make_float2(act_simt(o1.x, act2), act_simt(o1.y, act2)));
      }
    }
  }
}

bool launch_deconv_pair(oar_ctx* ctx, const float* in, int B, int H, int W, const OpRec& d1, const float* w1,
                        const float* b1, const OpRec& d2, const float* w2, const float* b2, float* out) {
  if (d1.p[0] != 24 || d1.p[1] != 24 || d2.p[0] != 24 || d2.p[1] != 1) return false;
  if ((((uintptr_t)in) & 15) || (((uintptr_t)out) & 7)) return false;
  const int segs = cdiv(W, 128);
  const long long items = (long long)B * H * segs;
  if (items <= 0 || items > 0x7fffffffLL) return false;
  const double px = (double)B * H * W;
  Launch l(ctx, "deconv_pair", px * (2.0 * 24 * 96 + 2.0 * 16 * 24), px * (4.0 * 24 + 4.0 * 16));
  const int grid = (int)std::min<long long>(items, (long long)ctx->sm_count * 8);
  deconv_pair_kernel<24, 24><<<grid, 128, 0, ctx->stream>>>(in, w1, b1, w2, b2, out, H, W, (int)items, segs, d1.p[2], d2.p[2]);
  OAR_CUDA(cudaGetLastError());
  return true;
}

}  // namespace oar
