// prepost.cu -- HBM-bound image kernels of the hot path (compiled with -fmad=false).
//   normalize      : NormalizeImage::normalize_batch_refs, normalization.rs:429-482 / simd.rs:87-104
//   resize_triangle: image::imageops::resize(Triangle) as called at crnn.rs:104-109, resize_detection.rs:314
//   crop_plan/warp : get_rotate_crop_image, utils/transform.rs:76-191 (+ :212-283 LU, :312-316 inverse, :439-502 bicubic)
//   crnn_normalize : normalize_crnn_chw_into, simd.rs:248-308
//   ctc_argmax/decode: decode.rs:452-614, simd.rs:190-229
#include <algorithm>

#include "prepost.cuh"

namespace oar {

// ---------------------------------------------------------------------------
// normalize: 3 B read + 12 B written per pixel; 4 pixels per thread
// ---------------------------------------------------------------------------
struct NormCoef {
  int src[3];
  float a[3], b[3];
};

__global__ void __launch_bounds__(256) normalize_kernel(const uint8_t* __restrict__ rgb,
                                                        const uint8_t* const* __restrict__ table,
                                                        float* __restrict__ out, long long plane, int B, NormCoef k,
                                                        int layout, int vec_ok) {
  long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // group of 4 pixels
  long long groups = (plane + 3) >> 2;
  int b = blockIdx.y;
  if (q >= groups) return;
  long long p0 = q << 2;
  const uint8_t* src = (table ? table[b] : rgb + (size_t)b * plane * 3) + p0 * 3;
  uint8_t px[12];
  int npx = (int)min((long long)4, plane - p0);
  if (vec_ok && npx == 4) {
    const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
    uint32_t w0 = __ldg(s4), w1 = __ldg(s4 + 1), w2 = __ldg(s4 + 2);
    *reinterpret_cast<uint32_t*>(px) = w0;
    *reinterpret_cast<uint32_t*>(px + 4) = w1;
    *reinterpret_cast<uint32_t*>(px + 8) = w2;
  } else {
    for (int i = 0; i < npx * 3; ++i) px[i] = src[i];
  }
  if (layout == 0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) v[i] = __fadd_rn(__fmul_rn((float)px[i * 3 + k.src[c]], k.a[c]), k.b[c]);
      float* dst = out + ((size_t)b * 3 + c) * plane + p0;
      if (vec_ok && npx == 4) {
        __stcs(reinterpret_cast<float4*>(dst), make_float4(v[0], v[1], v[2], v[3]));
      } else {
        for (int i = 0; i < npx; ++i) dst[i] = v[i];
      }
    }
  } else {
    float* dst = out + ((size_t)b * plane + p0) * 3;
    float v[12];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int c = 0; c < 3; ++c) v[i * 3 + c] = __fadd_rn(__fmul_rn((float)px[i * 3 + k.src[c]], k.a[c]), k.b[c]);
    if (vec_ok && npx == 4) {
      float4* d4 = reinterpret_cast<float4*>(dst);
      d4[0] = make_float4(v[0], v[1], v[2], v[3]);
      d4[1] = make_float4(v[4], v[5], v[6], v[7]);
      d4[2] = make_float4(v[8], v[9], v[10], v[11]);
    } else {
      for (int i = 0; i < npx * 3; ++i) dst[i] = v[i];
    }
  }
}

void launch_normalize(oar_ctx* ctx, const uint8_t* rgb, const uint8_t* const* d_table, bool table_aligned, float* out,
                      int B, int H, int W, const int src[3], const float alpha[3], const float beta[3], int layout) {
  NormCoef k;
  for (int i = 0; i < 3; ++i) k.src[i] = src[i], k.a[i] = alpha[i], k.b[i] = beta[i];
  long long plane = (long long)H * W;
  bool src_ok = d_table ? table_aligned : ((((uintptr_t)rgb & 3) == 0) && ((plane * 3) % 4 == 0 || B == 1));
  int vec_ok = (plane % 4 == 0) && src_ok && (((uintptr_t)out & 15) == 0);
  long long groups = (plane + 3) / 4;
  Launch l(ctx, "normalize", 2.0 * 3 * plane * B, 15.0 * plane * B);
  if (B > 65535) OAR_FAIL(OAR_E_INVALID, "normalize: batch %d exceeds 65535 images per launch", B);
  normalize_kernel<<<dim3(cdiv(groups, 256), B), 256, 0, ctx->stream>>>(rgb, d_table, out, plane, B, k, layout,
                                                                         vec_ok);
  OAR_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// Triangle resize (two passes, f32 intermediate, exactly the crate's order)
// ---------------------------------------------------------------------------
__device__ __forceinline__ float tri_kernel(float x) {
  float a = fabsf(x);
  return a < 1.0f ? 1.0f - a : 0.0f;
}

// image 0.25 sample.rs: bc_cubic_spline(x, b = 0, c = 0.5) = CatmullRom, support 2
__device__ __forceinline__ float catmull_kernel(float x) {
  const float b = 0.0f, c = 0.5f;
  float a = fabsf(x);
  float k;
  if (a < 1.0f)
    k = (12.0f - 9.0f * b - 6.0f * c) * (a * a * a) + (-18.0f + 12.0f * b + 6.0f * c) * (a * a) + (6.0f - 2.0f * b);
  else if (a < 2.0f)
    k = (-b - 6.0f * c) * (a * a * a) + (6.0f * b + 30.0f * c) * (a * a) + (-12.0f * b - 48.0f * c) * a +
        (8.0f * b + 24.0f * c);
  else
    k = 0.0f;
  return k / 6.0f;
}
__device__ __forceinline__ float filter_kernel(int filter, float x) { return filter == 1 ? catmull_kernel(x) : tri_kernel(x); }

struct TapRange {
  int left, right;
  float in, sratio;
};
__device__ __forceinline__ TapRange tap_range(int o, int in_len, int out_len, int filter = 0) {
  TapRange t;
  float ratio = (float)in_len / (float)out_len;
  t.sratio = ratio < 1.0f ? 1.0f : ratio;
  float support = (filter == 1 ? 2.0f : 1.0f) * t.sratio;
  float c = ((float)o + 0.5f) * ratio;
  long long l = (long long)floorf(c - support);
  l = l < 0 ? 0 : (l > (long long)in_len - 1 ? (long long)in_len - 1 : l);
  long long r = (long long)ceilf(c + support);
  r = r < l + 1 ? l + 1 : (r > (long long)in_len ? (long long)in_len : r);
  t.left = (int)l;
  t.right = (int)r;
  t.in = c - 0.5f;
  return t;
}

__global__ void resize_v_kernel(const ResizeJob* __restrict__ jobs) {
  ResizeJob j = jobs[blockIdx.z];
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int oy = blockIdx.y;
  if (x >= j.sw || oy >= j.dh) return;
  float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
  if (j.dh == j.sh && j.dw == j.sw) return;  // identity handled in the h pass
  TapRange tr = tap_range(oy, j.sh, j.dh, j.filter);
  float sum = 0.0f;
  for (int i = tr.left; i < tr.right; ++i) sum += filter_kernel(j.filter, ((float)i - tr.in) / tr.sratio);
  for (int i = tr.left; i < tr.right; ++i) {
    float w = filter_kernel(j.filter, ((float)i - tr.in) / tr.sratio) / sum;
    const uint8_t* p = j.src + ((size_t)i * j.sw + x) * 3;
    t0 += (float)p[0] * w;
    t1 += (float)p[1] * w;
    t2 += (float)p[2] * w;
  }
  float* o = j.tmp + ((size_t)oy * j.sw + x) * 3;
  o[0] = t0, o[1] = t1, o[2] = t2;
}

__global__ void resize_h_kernel(const ResizeJob* __restrict__ jobs) {
  ResizeJob j = jobs[blockIdx.z];
  int ox = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (ox >= j.dw || y >= j.dh) return;
  uint8_t* o = j.dst + ((size_t)y * j.dw + ox) * 3;
  if (j.dh == j.sh && j.dw == j.sw) {  // same size => plain copy (image crate fast path)
    const uint8_t* s = j.src + ((size_t)y * j.sw + ox) * 3;
    o[0] = s[0], o[1] = s[1], o[2] = s[2];
    return;
  }
  TapRange tr = tap_range(ox, j.sw, j.dw, j.filter);
  float sum = 0.0f;
  for (int i = tr.left; i < tr.right; ++i) sum += filter_kernel(j.filter, ((float)i - tr.in) / tr.sratio);
  float t0 = 0.0f, t1 = 0.0f, t2 = 0.0f;
  for (int i = tr.left; i < tr.right; ++i) {
    float w = filter_kernel(j.filter, ((float)i - tr.in) / tr.sratio) / sum;
    const float* p = j.tmp + ((size_t)y * j.sw + i) * 3;
    t0 += p[0] * w;
    t1 += p[1] * w;
    t2 += p[2] * w;
  }
  o[0] = (uint8_t)roundf(fminf(fmaxf(t0, 0.0f), 255.0f));
  o[1] = (uint8_t)roundf(fminf(fmaxf(t1, 0.0f), 255.0f));
  o[2] = (uint8_t)roundf(fminf(fmaxf(t2, 0.0f), 255.0f));
}

// Both passes in one kernel for sources whose rows fit in shared memory: a CTA owns R output rows of one job, builds
// their vertically filtered f32 rows (the crate's intermediate image, never written to HBM) in shared memory and
// filters them horizontally from there.  Every value is produced by the expressions of the two kernels above in the
// same order -- weights = kernel / sum, taps ascending, separate multiply and add (this file is built -fmad=false) --
// so the bytes are identical; the filter weights of a row / column are computed once per CTA instead of once per
// sample.  (The two-pass form launched max_sw/128 x 48 x jobs blocks, most of them past their crop's width, and moved
// 12 B per intermediate sample through HBM both ways: 0.11 ms per 256-crop batch against 0.02 ms here.)
constexpr int RF_THREADS = 256, RF_MAX_TAPS = 48;
__global__ void __launch_bounds__(RF_THREADS) resize_fused_kernel(const ResizeJob* __restrict__ jobs, int R) {
  extern __shared__ float rf_tmp[];  // [R][sw * 3]
  __shared__ float wv[4][RF_MAX_TAPS];
  __shared__ int vleft[4], vn[4];
  __shared__ float vsum[4], vin[4], vsr[4];  // for rows with more taps than the table holds
  const ResizeJob j = jobs[blockIdx.y];
  const int oy0 = blockIdx.x * R;
  if (oy0 >= j.dh) return;
  const int rows = min(R, j.dh - oy0);
  const int row_e = j.sw * 3;
  if (j.dh == j.sh && j.dw == j.sw) {  // same size => plain copy (image crate fast path)
    for (int r = 0; r < rows; ++r) {
      const uint8_t* sp = j.src + (size_t)(oy0 + r) * row_e;
      uint8_t* dp = j.dst + (size_t)(oy0 + r) * row_e;
      for (int e = threadIdx.x; e < row_e; e += RF_THREADS) dp[e] = sp[e];
    }
    return;
  }
  // vertical weights of the CTA's rows: one thread per row, the loops of resize_v_kernel
  if (threadIdx.x < rows) {
    const TapRange tr = tap_range(oy0 + threadIdx.x, j.sh, j.dh, j.filter);
    float sum = 0.0f;
    for (int i = tr.left; i < tr.right; ++i) sum += filter_kernel(j.filter, ((float)i - tr.in) / tr.sratio);
    const int n = min(tr.right - tr.left, RF_MAX_TAPS);
    for (int i = 0; i < n; ++i) wv[threadIdx.x][i] = filter_kernel(j.filter, ((float)(tr.left + i) - tr.in) / tr.sratio) / sum;
    vleft[threadIdx.x] = tr.left, vn[threadIdx.x] = tr.right - tr.left;
    vsum[threadIdx.x] = sum, vin[threadIdx.x] = tr.in, vsr[threadIdx.x] = tr.sratio;
  }
  __syncthreads();
  for (int r = 0; r < rows; ++r) {
    const int left = vleft[r], n = vn[r];
    float* trow = rf_tmp + (size_t)r * row_e;
    const uint8_t* sp = j.src + (size_t)left * row_e;
    if (n <= RF_MAX_TAPS) {
      for (int e = threadIdx.x; e < row_e; e += RF_THREADS) {
        float t = 0.0f;
        for (int i = 0; i < n; ++i) t += (float)sp[(size_t)i * row_e + e] * wv[r][i];
        trow[e] = t;
      }
    } else {  // a steep vertical reduction: weights on the fly, as resize_v_kernel does
      const float sum = vsum[r], in = vin[r], sr = vsr[r];
      for (int e = threadIdx.x; e < row_e; e += RF_THREADS) {
        float t = 0.0f;
        for (int i = 0; i < n; ++i)
          t += (float)sp[(size_t)i * row_e + e] * (filter_kernel(j.filter, ((float)(left + i) - in) / sr) / sum);
        trow[e] = t;
      }
    }
  }
  __syncthreads();
  for (int ox = threadIdx.x; ox < j.dw; ox += RF_THREADS) {
    const TapRange tr = tap_range(ox, j.sw, j.dw, j.filter);
    float sum = 0.0f;
    for (int i = tr.left; i < tr.right; ++i) sum += filter_kernel(j.filter, ((float)i - tr.in) / tr.sratio);
    float acc[4][3];
#pragma unroll
    for (int r = 0; r < 4; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 0.0f;
    for (int i = tr.left; i < tr.right; ++i) {
      const float w = filter_kernel(j.filter, ((float)i - tr.in) / tr.sratio) / sum;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (r < rows) {
          const float* pp = rf_tmp + (size_t)r * row_e + i * 3;
          acc[r][0] += pp[0] * w, acc[r][1] += pp[1] * w, acc[r][2] += pp[2] * w;
        }
      }
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (r < rows) {
        uint8_t* o = j.dst + ((size_t)(oy0 + r) * j.dw + ox) * 3;
        o[0] = (uint8_t)roundf(fminf(fmaxf(acc[r][0], 0.0f), 255.0f));
        o[1] = (uint8_t)roundf(fminf(fmaxf(acc[r][1], 0.0f), 255.0f));
        o[2] = (uint8_t)roundf(fminf(fmaxf(acc[r][2], 0.0f), 255.0f));
      }
    }
  }
}

void launch_resize_triangle(oar_ctx* ctx, const ResizeJob* d_jobs, int n_jobs, int max_sw, int max_dw, int max_dh) {
  if (n_jobs == 0) return;
  if (max_dh > 65535) OAR_FAIL(OAR_E_INVALID, "resize: %d output rows exceed 65535 per launch", max_dh);
  // fused form: the f32 rows of R output rows in shared memory
  static const bool two_pass = getenv("OAR_DBG_RESIZE2") != nullptr;  // A/B switch
  const size_t row_bytes = (size_t)max_sw * 3 * sizeof(float);
  const int R = row_bytes ? (int)std::min<size_t>(4, (96 * 1024) / row_bytes) : 0;
  const bool fused = !two_pass && R >= 1;
  if (fused) ensure_max_dynamic_smem((const void*)resize_fused_kernel, ctx->device, 96 * 1024);
  // gridDim.y / z are limited to 65535: slice the job list
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    const int nj = std::min(n_jobs - j0, 65535);
    if (fused) {
      Launch l(ctx, "resize_fused");
      resize_fused_kernel<<<dim3(cdiv(max_dh, R), nj), RF_THREADS, (size_t)R * row_bytes, ctx->stream>>>(d_jobs + j0, R);
      continue;
    }
    {
      Launch l(ctx, "resize_tri_v");
      resize_v_kernel<<<dim3(cdiv(max_sw, 128), max_dh, nj), 128, 0, ctx->stream>>>(d_jobs + j0);
    }
    {
      Launch l(ctx, "resize_tri_h");
      resize_h_kernel<<<dim3(cdiv(max_dw, 128), max_dh, nj), 128, 0, ctx->stream>>>(d_jobs + j0);
    }
  }
  OAR_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// crop planning: one thread per quad
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned sat_u32(float v) {
  if (!(v > 0.0f)) return 0u;
  if (v >= 4294967296.0f) return 4294967295u;
  return (unsigned)v;
}
__device__ __forceinline__ float hypot_f32(float a, float b) {
  // libm hypotf evaluates in double and rounds once
  return (float)sqrt((double)a * (double)a + (double)b * (double)b);
}

__device__ bool perspective_transform_dev(const float sx[4], const float sy[4], const float dx[4], const float dy[4],
                                          float M[9]) {
  float a[8][8];
  float b[8];
  for (int i = 0; i < 8; ++i)
    for (int j = 0; j < 8; ++j) a[i][j] = 0.0f;
  for (int i = 0; i < 4; ++i) {
    a[2 * i][0] = sx[i], a[2 * i][1] = sy[i], a[2 * i][2] = 1.0f;
    a[2 * i][6] = -sx[i] * dx[i], a[2 * i][7] = -sy[i] * dx[i];
    a[2 * i + 1][3] = sx[i], a[2 * i + 1][4] = sy[i], a[2 * i + 1][5] = 1.0f;
    a[2 * i + 1][6] = -sx[i] * dy[i], a[2 * i + 1][7] = -sy[i] * dy[i];
    b[2 * i] = dx[i];
    b[2 * i + 1] = dy[i];
  }
  int pa[8], pb[8], np = 0;
  for (int i = 0; i < 8; ++i) {
    int piv = i;
    float mx = fabsf(a[i][i]);
    for (int r = i + 1; r < 8; ++r) {
      float v = fabsf(a[r][i]);
      if (v > mx) mx = v, piv = r;
    }
    float diag = a[piv][i];
    if (diag == 0.0f) continue;
    if (piv != i) {
      pa[np] = i, pb[np] = piv, ++np;
      for (int c = 0; c < 8; ++c) {
        float t = a[i][c];
        a[i][c] = a[piv][c];
        a[piv][c] = t;
      }
    }
    float inv_diag = 1.0f / diag;
    for (int r = i + 1; r < 8; ++r) a[r][i] *= inv_diag;
    for (int c = i + 1; c < 8; ++c) {
      float pv = -a[i][c];
      for (int r = i + 1; r < 8; ++r) a[r][c] = pv * a[r][i] + a[r][c];
    }
  }
  for (int k = 0; k < np; ++k) {
    float t = b[pa[k]];
    b[pa[k]] = b[pb[k]];
    b[pb[k]] = t;
  }
  for (int i = 0; i < 7; ++i) {
    float coeff = -b[i];
    for (int r = i + 1; r < 8; ++r) b[r] = coeff * a[r][i] + b[r];
  }
  for (int i = 7; i >= 0; --i) {
    float diag = a[i][i];
    if (diag == 0.0f) return false;
    float coeff = b[i] / diag;
    b[i] = coeff;
    float nc = -coeff;
    for (int r = 0; r < i; ++r) b[r] = nc * a[r][i] + b[r];
  }
  for (int i = 0; i < 8; ++i) M[i] = b[i];
  M[8] = 1.0f;
  return true;
}

__device__ bool invert3_dev(const float m[9], float inv[9]) {
  float m11 = m[0], m12 = m[1], m13 = m[2], m21 = m[3], m22 = m[4], m23 = m[5], m31 = m[6], m32 = m[7], m33 = m[8];
  float minor_m12_m23 = m22 * m33 - m32 * m23;
  float minor_m11_m23 = m21 * m33 - m31 * m23;
  float minor_m11_m22 = m21 * m32 - m31 * m22;
  float det = m11 * minor_m12_m23 - m12 * minor_m11_m23 + m13 * minor_m11_m22;
  if (det == 0.0f) return false;
  inv[0] = minor_m12_m23 / det;
  inv[1] = (m13 * m32 - m33 * m12) / det;
  inv[2] = (m12 * m23 - m22 * m13) / det;
  inv[3] = -minor_m11_m23 / det;
  inv[4] = (m11 * m33 - m31 * m13) / det;
  inv[5] = (m13 * m21 - m23 * m11) / det;
  inv[6] = minor_m11_m22 / det;
  inv[7] = (m12 * m31 - m32 * m11) / det;
  inv[8] = (m11 * m22 - m21 * m12) / det;
  return true;
}

__global__ void crop_plan_kernel(CropPlan* __restrict__ plans, int n, const ImageRef* __restrict__ images) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  CropPlan p = plans[i];
  ImageRef im = images[p.img];
  float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
  for (int k = 0; k < 4; ++k) {
    mnx = fminf(mnx, p.quad[2 * k]);
    mxx = fmaxf(mxx, p.quad[2 * k]);
    mny = fminf(mny, p.quad[2 * k + 1]);
    mxy = fmaxf(mxy, p.quad[2 * k + 1]);
  }
  unsigned left = sat_u32(fmaxf(mnx, 0.0f)), top = sat_u32(fmaxf(mny, 0.0f));
  unsigned right = sat_u32(fminf(mxx, (float)im.w)), bottom = sat_u32(fminf(mxy, (float)im.h));
  p.status = 0;
  p.ow = p.oh = 0;
  p.wh_ratio = 0.0f;
  p.axis_aligned = 0;
  p.rot270 = 0;
  if (right <= left || bottom <= top) {
    p.status = 1;
    plans[i] = p;
    return;
  }
  unsigned cw = right - left, ch = bottom - top;
  p.left = (int)left, p.top = (int)top, p.cw = (int)cw, p.ch = (int)ch;
  float px[4], py[4];
  for (int k = 0; k < 4; ++k) {
    px[k] = p.quad[2 * k] - (float)left;
    py[k] = p.quad[2 * k + 1] - (float)top;
  }
  // stable sort by x (insertion sort == stable, partial_cmp on finite floats)
  for (int a = 1; a < 4; ++a) {
    float kx = px[a], ky = py[a];
    int b = a - 1;
    while (b >= 0 && px[b] > kx) {
      px[b + 1] = px[b], py[b + 1] = py[b];
      --b;
    }
    px[b + 1] = kx, py[b + 1] = ky;
  }
  int ia = 0, id = 1;
  if (py[1] < py[0]) ia = 1, id = 0;
  int ib = 2, ic = 3;
  if (py[3] < py[2]) ib = 3, ic = 2;
  float ox[4] = {px[ia], px[ib], px[ic], px[id]}, oy[4] = {py[ia], py[ib], py[ic], py[id]};
  float fw = (float)cw, fh = (float)ch;
  bool aligned = ox[0] == 0.0f && oy[0] == 0.0f && ox[1] == fw && oy[1] == 0.0f && ox[2] == fw && oy[2] == fh &&
                 ox[3] == 0.0f && oy[3] == fh;
  if (aligned) {
    p.axis_aligned = 1;
    p.rw = (int)cw, p.rh = (int)ch;
  } else {
    float w1 = hypot_f32(ox[0] - ox[1], oy[0] - oy[1]), w2 = hypot_f32(ox[2] - ox[3], oy[2] - oy[3]);
    float h1 = hypot_f32(ox[0] - ox[3], oy[0] - oy[3]), h2 = hypot_f32(ox[1] - ox[2], oy[1] - oy[2]);
    unsigned icw = sat_u32(roundf(fmaxf(w1, w2))), ich = sat_u32(roundf(fmaxf(h1, h2)));
    if (icw == 0 || ich == 0) {
      p.status = 2;
      plans[i] = p;
      return;
    }
    float dx[4] = {0.0f, (float)icw, (float)icw, 0.0f}, dy[4] = {0.0f, 0.0f, (float)ich, (float)ich};
    float M[9];
    if (!perspective_transform_dev(ox, oy, dx, dy, M)) {
      p.status = 3;
      plans[i] = p;
      return;
    }
    if (!invert3_dev(M, p.inv)) {
      p.status = 4;
      plans[i] = p;
      return;
    }
    p.rw = (int)icw, p.rh = (int)ich;
  }
  if ((float)p.rh >= (float)p.rw * 1.5f) {
    p.rot270 = 1;
    p.ow = p.rh, p.oh = p.rw;
  } else {
    p.ow = p.rw, p.oh = p.rh;
  }
  p.wh_ratio = (float)p.ow / (float)max(p.oh, 1);
  plans[i] = p;
}

void launch_crop_plan(oar_ctx* ctx, CropPlan* d_plans, int n, const ImageRef* d_images) {
  if (!n) return;
  Launch l(ctx, "crop_plan");
  crop_plan_kernel<<<cdiv(n, 64), 64, 0, ctx->stream>>>(d_plans, n, d_images);
  OAR_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// perspective warp, bicubic a=-0.5, replicate border of the crop box
// ---------------------------------------------------------------------------
__device__ __forceinline__ float cubic_kernel(float t) {
  const float A = -0.5f;
  float ta = fabsf(t);
  if (ta <= 1.0f) return (A + 2.0f) * ta * ta * ta - (A + 3.0f) * ta * ta + 1.0f;
  if (ta < 2.0f) return A * ta * ta * ta - 5.0f * A * ta * ta + 8.0f * A * ta - 4.0f * A;
  return 0.0f;
}
__device__ __forceinline__ int sat_i32(float v) {
  if (v != v) return 0;
  if (v >= 2147483648.0f) return 2147483647;
  if (v <= -2147483648.0f) return -2147483647 - 1;
  return (int)v;
}

__global__ void __launch_bounds__(256) crop_warp_kernel(const CropPlan* __restrict__ plans,
                                                        const ImageRef* __restrict__ images, uint8_t* __restrict__ pool) {
  const CropPlan& p = plans[blockIdx.y];
  if (p.status != 0) return;
  int npx = p.ow * p.oh;
  ImageRef im = images[p.img];
  uint8_t* out = pool + p.out_off;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < npx; idx += gridDim.x * blockDim.x) {
    int Y = idx / p.ow, X = idx - Y * p.ow;
    int x, y;  // rectified coordinates
    if (p.rot270) {
      y = X;
      x = p.rw - 1 - Y;
    } else {
      x = X;
      y = Y;
    }
    uint8_t r0, r1, r2;
    if (p.axis_aligned) {
      const uint8_t* s = im.p + ((size_t)(p.top + y) * im.w + p.left + x) * 3;
      r0 = s[0], r1 = s[1], r2 = s[2];
    } else {
      float fx = (float)x, fy = (float)y;
      float sx = p.inv[0] * fx;
      sx = p.inv[1] * fy + sx;
      sx = p.inv[2] * 1.0f + sx;
      float sy = p.inv[3] * fx;
      sy = p.inv[4] * fy + sy;
      sy = p.inv[5] * 1.0f + sy;
      float sz = p.inv[6] * fx;
      sz = p.inv[7] * fy + sz;
      sz = p.inv[8] * 1.0f + sz;
      const uint8_t* base = im.p + ((size_t)p.top * im.w + p.left) * 3;
      size_t stride = (size_t)im.w * 3;
      if (fabsf(sz) > 1.1920929e-7f) {
        float u = sx / sz, v = sy / sz;
        float fu = floorf(u), fv = floorf(v);
        int xi = sat_i32(fu), yi = sat_i32(fv);
        float du = u - (float)xi, dv = v - (float)yi;
        float wx[4] = {cubic_kernel(du + 1.0f), cubic_kernel(du), cubic_kernel(du - 1.0f), cubic_kernel(du - 2.0f)};
        float wy[4] = {cubic_kernel(dv + 1.0f), cubic_kernel(dv), cubic_kernel(dv - 1.0f), cubic_kernel(dv - 2.0f)};
        size_t cx[4], cy[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          long long a = (long long)xi - 1 + k, b = (long long)yi - 1 + k;
          a = a < 0 ? 0 : (a > p.cw - 1 ? p.cw - 1 : a);
          b = b < 0 ? 0 : (b > p.ch - 1 ? p.ch - 1 : b);
          cx[k] = (size_t)a * 3;
          cy[k] = (size_t)b * stride;
        }
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            float wgt = wx[k] * wy[j];
            const uint8_t* s = base + cy[j] + cx[k];
            a0 += wgt * (float)s[0];
            a1 += wgt * (float)s[1];
            a2 += wgt * (float)s[2];
          }
        }
        r0 = (uint8_t)fminf(fmaxf(roundf(a0), 0.0f), 255.0f);
        r1 = (uint8_t)fminf(fmaxf(roundf(a1), 0.0f), 255.0f);
        r2 = (uint8_t)fminf(fmaxf(roundf(a2), 0.0f), 255.0f);
      } else {
        r0 = base[0], r1 = base[1], r2 = base[2];
      }
    }
    uint8_t* o = out + (size_t)idx * 3;
    o[0] = r0, o[1] = r1, o[2] = r2;
  }
}

void launch_crop_warp(oar_ctx* ctx, const CropPlan* d_plans, int n, const ImageRef* d_images, uint8_t* pool,
                      long long total_px) {
  if (!n) return;
  Launch l(ctx, "crop_warp", 96.0 * total_px, 6.0 * total_px);
  // gridDim.y is limited to 65535 (the reference has no such limit: it flushes its crop pool every 4096 crops, but one
  // call here plans all boxes of all pages at once): slice the plan list
  for (int j0 = 0; j0 < n; j0 += 65535)
    crop_warp_kernel<<<dim3(32, std::min(n - j0, 65535)), 256, 0, ctx->stream>>>(d_plans + j0, d_images, pool);
  OAR_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// CRNN normalize + pad
// ---------------------------------------------------------------------------
__global__ void crnn_normalize_kernel(const CrnnJob* __restrict__ jobs, int img_h, int tensor_w, float* __restrict__ out,
                                      int layout) {
  CrnnJob j = jobs[blockIdx.z];
  int x = blockIdx.x * blockDim.x + threadIdx.x;
  int y = blockIdx.y;
  if (x >= tensor_w) return;
  float v[3] = {0.0f, 0.0f, 0.0f};
  if (x < j.rw) {
    const uint8_t* s = j.src + ((size_t)y * j.rw + x) * 3;
#pragma unroll
    for (int c = 0; c < 3; ++c) v[c] = ((float)s[2 - c] / 255.0f - 0.5f) / 0.5f;
  }
  size_t n = blockIdx.z;
  if (layout == 0) {
    size_t plane = (size_t)img_h * tensor_w;
#pragma unroll
    for (int c = 0; c < 3; ++c) out[(n * 3 + c) * plane + (size_t)y * tensor_w + x] = v[c];
  } else {
    float* o = out + ((n * img_h + y) * tensor_w + x) * 3;
    o[0] = v[0], o[1] = v[1], o[2] = v[2];
  }
}

void launch_crnn_normalize(oar_ctx* ctx, const CrnnJob* d_jobs, int n, int img_h, int tensor_w, float* out,
                           int layout) {
  if (!n) return;
  Launch l(ctx, "crnn_normalize", 0, 15.0 * n * img_h * tensor_w);
  if (n > 65535) OAR_FAIL(OAR_E_INVALID, "crnn_normalize: %d crops exceed 65535 per launch", n);
  crnn_normalize_kernel<<<dim3(cdiv(tensor_w, 128), img_h, n), 128, 0, ctx->stream>>>(d_jobs, img_h, tensor_w, out,
                                                                                       layout);
  OAR_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// CTC argmax over probabilities: LAST maximal index wins (simd.rs:194-204)
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ctc_argmax_kernel(const float* __restrict__ pred, int V, int32_t* __restrict__ idx,
                                                         float* __restrict__ prob) {
  __shared__ float s_val[8];
  __shared__ int s_idx[8];
  size_t row = blockIdx.x;
  const float* z = pred + row * V;
  float mx = -INFINITY;
  int mi = 0;
  for (int i = threadIdx.x; i < V; i += 256) {
    float v = __ldcs(z + i);
    if (v >= mx) mx = v, mi = i;
  }
  for (int o = 16; o; o >>= 1) {
    float ov = __shfl_xor_sync(0xffffffffu, mx, o);
    int oi = __shfl_xor_sync(0xffffffffu, mi, o);
    if (ov > mx || (ov == mx && oi > mi)) mx = ov, mi = oi;
  }
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) s_val[warp] = mx, s_idx[warp] = mi;
  __syncthreads();
  if (threadIdx.x == 0) {
    mx = s_val[0], mi = s_idx[0];
    for (int w = 1; w < 8; ++w) {
      float ov = s_val[w];
      int oi = s_idx[w];
      if (ov > mx || (ov == mx && oi > mi)) mx = ov, mi = oi;
    }
    idx[row] = mi;
    prob[row] = mx;
  }
}

void launch_ctc_argmax(oar_ctx* ctx, const float* pred, long long rows, int V, int32_t* idx, float* prob) {
  if (rows == 0 || V == 0) return;
  Launch l(ctx, "ctc_argmax", (double)rows * V, 4.0 * rows * V);
  ctc_argmax_kernel<<<(unsigned)rows, 256, 0, ctx->stream>>>(pred, V, idx, prob);
  OAR_CUDA(cudaGetLastError());
}

// decode.rs:505-614: prev = blank; emit when idx != 0 && idx != prev && idx < n_chars; prev = idx always
__global__ void ctc_decode_kernel(const int32_t* __restrict__ idx, const float* __restrict__ prob, int B, int T,
                                  int n_chars, int32_t* __restrict__ labels, int32_t* __restrict__ cols,
                                  int32_t* __restrict__ lens, float* __restrict__ scores) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  int prev = 0, n = 0;
  float sum = 0.0f;
  for (int t = 0; t < T; ++t) {
    int k = idx[(size_t)b * T + t];
    if (k != 0 && k != prev && k >= 0 && k < n_chars) {
      labels[(size_t)b * T + n] = k;
      cols[(size_t)b * T + n] = t;
      sum += prob[(size_t)b * T + t];
      ++n;
    }
    prev = k;
  }
  lens[b] = n;
  scores[b] = n ? sum / (float)n : 0.0f;
}

void launch_ctc_decode(oar_ctx* ctx, const int32_t* idx, const float* prob, int B, int T, int n_chars, int32_t* labels,
                       int32_t* cols, int32_t* lens, float* scores) {
  if (!B) return;
  Launch l(ctx, "ctc_decode", 0, 8.0 * B * T);
  ctc_decode_kernel<<<cdiv(B, 64), 64, 0, ctx->stream>>>(idx, prob, B, T, n_chars, labels, cols, lens, scores);
  OAR_CUDA(cudaGetLastError());
}

// ---------------------------------------------------------------------------
// text-line orientation stage: top-1 class and the in-place 180-degree rotation (ocr.rs:755-792)
// ---------------------------------------------------------------------------
// Topk::extract_topk_from_prediction (utils/topk.rs) sorts (index, score) pairs by score descending with a stable
// sort, so the top-1 is the first index holding the maximum.
__global__ void cls_top1_kernel(const float* __restrict__ probs, int n, int C, int32_t* __restrict__ ids,
                                float* __restrict__ scores) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = probs + (size_t)i * C;
  float best = p[0];
  int id = 0;
  for (int c = 1; c < C; ++c) {
    float v = p[c];
    if (v > best) best = v, id = c;
  }
  ids[i] = id;
  scores[i] = best;
}

void launch_cls_top1(oar_ctx* ctx, const float* probs, int n, int C, int32_t* ids, float* scores) {
  if (n <= 0 || C <= 0) return;
  Launch l(ctx, "cls_top1", 0, 4.0 * n * C + 8.0 * n);
  cls_top1_kernel<<<cdiv(n, 128), 128, 0, ctx->stream>>>(probs, n, C, ids, scores);
  OAR_CUDA(cudaGetLastError());
}

// One thread per pixel pair (i, npix-1-i); 6 B read + 6 B written per pair, HBM/L2-bound.
__global__ void __launch_bounds__(256) rotate180_kernel(const Rot180Job* __restrict__ jobs,
                                                        const int32_t* __restrict__ class_ids) {
  const int j = blockIdx.y;
  if (class_ids && class_ids[j] != 1) return;
  const Rot180Job job = jobs[j];
  const int half = job.npix >> 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < half; i += gridDim.x * blockDim.x) {
    uint8_t* a = job.p + (size_t)i * 3;
    uint8_t* b = job.p + (size_t)(job.npix - 1 - i) * 3;
    uint8_t a0 = a[0], a1 = a[1], a2 = a[2];
    uint8_t b0 = b[0], b1 = b[1], b2 = b[2];
    a[0] = b0, a[1] = b1, a[2] = b2;
    b[0] = a0, b[1] = a1, b[2] = a2;
  }
}

void launch_rotate180(oar_ctx* ctx, const Rot180Job* d_jobs, int n_jobs, int max_npix, const int32_t* class_ids) {
  if (n_jobs <= 0 || max_npix < 2) return;
  Launch l(ctx, "rotate180", 0, 6.0 * max_npix * n_jobs);
  int bx = std::max(1, std::min(cdiv(max_npix / 2, 256), 64));
  for (int j0 = 0; j0 < n_jobs; j0 += 65535) {
    int nj = std::min(n_jobs - j0, 65535);
    rotate180_kernel<<<dim3(bx, nj), 256, 0, ctx->stream>>>(d_jobs + j0, class_ids ? class_ids + j0 : nullptr);
  }
  OAR_CUDA(cudaGetLastError());
}

}  // namespace oar
