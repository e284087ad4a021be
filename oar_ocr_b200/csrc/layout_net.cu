// layout_net.cu -- PP-DocLayout-L (RT-DETR-L) on the device: preprocess, network, exported-model rows.
//
// In the reference this is LayoutDetectionAdapter::execute (layout_detection_adapter.rs:1128-1197) ->
// ScaleAwareDetectorModel::preprocess / infer (scale_aware_detector.rs:150-333): resize_exact to the fixed input with
// FilterType::CatmullRom, scale 1/255, RGB, three inputs (image, im_shape, scale_factor) through ONNX Runtime, and an
// output tensor of [class_id, score, x1, y1, x2, y2] rows that postprocess_pp_doclayout reads (csrc/layout.cu).
// The architecture follows the reference's in-tree description of the family (oar-ocr-vl/src/models/pp_doclayout/:
// hgnetv2.rs, encoder.rs, decoder.rs, model.rs:154-183, 296-345, 459-476).
//
//   encoder model (OARG, KIND_FEAT)  backbone + hybrid encoder + decoder-input projections -> memory [B, N, 256]
//                                    (models.build_layout_encoder; runs on the layer-list executor, engine.cu)
//   head model (OARG, KIND_FEAT)     the decoder's layers as a table called BY POSITION (models.build_layout_head):
//                                    every Linear is a 1x1 convolution on the tcgen05 kernels, LayerNorm and
//                                    self-attention are the executor's own layers
//   here                              anchors, top-300 query selection (stable segmented radix sort: ties by index),
//                                    6 decoder layers with multi-scale deformable attention (8 heads x 3 levels x 4
//                                    points, bilinear, zeros outside), iterative box refinement, sigmoid + top-300
//                                    over (query, class), cxcywh -> xyxy in source pixels
#include <cub/cub.cuh>

#include <cfloat>
#include <cmath>
#include <vector>

#include "engine.cuh"
#include "prepost.cuh"

namespace oar {

namespace {

constexpr int LD = 256, LHEADS = 8, LHD = 32, LLEVELS = 3, LPOINTS = 4, LQUERIES = 300;
// positions in the head model (models.build_layout_head)
enum { LH_ENC_OUTPUT = 0, LH_ENC_OUTPUT_LN, LH_ENC_SCORE, LH_ENC_BBOX0, LH_ENC_BBOX1, LH_ENC_BBOX2, LH_QPOS0, LH_QPOS1,
       LH_FIXED };
enum { LL_SA = 0, LL_LN1, LL_OFFSETS, LL_WEIGHTS, LL_VALUE, LL_OUT, LL_LN2, LL_FC1, LL_FC2, LL_LN3, LL_SCORE, LL_BBOX0,
       LL_BBOX1, LL_BBOX2, LL_PER_LAYER };

__global__ void mask_rows_kernel(const float4* __restrict__ x, const float* __restrict__ valid, float4* __restrict__ out,
                                 int N, int C4, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float v = valid[(i / C4) % N];
  const float4 a = __ldg(x + i);
  out[i] = make_float4(a.x * v, a.y * v, a.z * v, a.w * v);
}

__global__ void row_max_kernel(const float* __restrict__ x, float* __restrict__ out, int C, size_t rows) {
  size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  const float* p = x + r * C;
  float m = p[0];
  for (int c = 1; c < C; ++c) m = fmaxf(m, p[c]);
  out[r] = m;
}

__global__ void iota_segments_kernel(int32_t* __restrict__ idx, int per, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) idx[i] = (int32_t)(i % per);
}

// hidden[b, q, :] = memory[b, top[b, q], :]
__global__ void gather_rows_kernel(const float4* __restrict__ mem, const int32_t* __restrict__ sorted_idx, float4* __restrict__ out,
                                   int N, int Q, int C4, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c4 = (int)(i % C4);
  const size_t r = i / C4;
  const int q = (int)(r % Q), b = (int)(r / Q);
  out[i] = __ldg(mem + ((size_t)b * N + sorted_idx[(size_t)b * N + q]) * C4 + c4);
}

__device__ __forceinline__ float sigmoid_f(float v) { return 1.0f / (1.0f + expf(-v)); }
// decoder.rs:768-779
__device__ __forceinline__ float inverse_sigmoid_f(float x) {
  x = fminf(fmaxf(x, 0.0f), 1.0f);
  return logf(fmaxf(x, 1e-5f) / fmaxf(1.0f - x, 1e-5f));
}

// reference = sigmoid(z + anchors[top])   (first selection)
__global__ void ref_from_anchors_kernel(const float* __restrict__ z, const float* __restrict__ anchors,
                                        const int32_t* __restrict__ sorted_idx, float* __restrict__ ref, int N, int Q, int total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = i & 3, r = i >> 2;
  const int q = r % Q, b = r / Q;
  ref[i] = sigmoid_f(z[i] + anchors[(size_t)sorted_idx[(size_t)b * N + q] * 4 + k]);
}
// reference = sigmoid(z + inverse_sigmoid(reference))   (per-layer refinement)
__global__ void refine_ref_kernel(const float* __restrict__ z, float* __restrict__ ref, int total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) ref[i] = sigmoid_f(z[i] + inverse_sigmoid_f(ref[i]));
}

__global__ void add4_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ o, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 x = __ldg(a + i), y = __ldg(b + i);
  o[i] = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
}

struct LevelShapes {
  int h[LLEVELS], w[LLEVELS], off[LLEVELS];
};

// Multi-scale deformable attention (decoder.rs:212-470).  One block per (query, image): thread = (head, channel).
// value [B, N, 256]; offsets [B, Q, heads, levels, points, 2]; wlogit [B, Q, heads, levels * points] (softmax here);
// ref [B, Q, 4] (cx, cy, w, h in [0,1]).  loc = ref_xy + offset / POINTS * ref_wh * 0.5; bilinear sample at
// loc * size - 0.5 (grid_sample, align_corners = false), zero outside.
__global__ void __launch_bounds__(LD) deform_attn_kernel(const float* __restrict__ value, const float* __restrict__ offsets,
                                                         const float* __restrict__ wlogit, const float* __restrict__ ref,
                                                         float* __restrict__ out, int N, int Q, LevelShapes S) {
  const int q = blockIdx.x, b = blockIdx.y;
  const int h = threadIdx.x >> 5, d = threadIdx.x & 31;
  const size_t row = (size_t)b * Q + q;
  const float* wl = wlogit + row * (LHEADS * LLEVELS * LPOINTS) + h * (LLEVELS * LPOINTS);
  float wv[LLEVELS * LPOINTS];
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < LLEVELS * LPOINTS; ++i) wv[i] = wl[i], mx = fmaxf(mx, wv[i]);
  float den = 0.0f;
#pragma unroll
  for (int i = 0; i < LLEVELS * LPOINTS; ++i) wv[i] = expf(wv[i] - mx), den += wv[i];
  const float inv = 1.0f / den;
  const float rx = ref[row * 4], ry = ref[row * 4 + 1], rw = ref[row * 4 + 2], rh = ref[row * 4 + 3];
  const float* off = offsets + row * (LHEADS * LLEVELS * LPOINTS * 2) + h * (LLEVELS * LPOINTS * 2);
  const float* vb = value + (size_t)b * N * LD + h * LHD + d;
  float acc = 0.0f;
#pragma unroll
  for (int l = 0; l < LLEVELS; ++l) {
    const int H = S.h[l], W = S.w[l];
    const float* vl = vb + (size_t)S.off[l] * LD;
    float lvl = 0.0f;
#pragma unroll
    for (int p = 0; p < LPOINTS; ++p) {
      const float lx = rx + off[(l * LPOINTS + p) * 2] / (float)LPOINTS * rw * 0.5f;
      const float ly = ry + off[(l * LPOINTS + p) * 2 + 1] / (float)LPOINTS * rh * 0.5f;
      const float ix = ((2.0f * lx - 1.0f) + 1.0f) * 0.5f * (float)W - 0.5f;
      const float iy = ((2.0f * ly - 1.0f) + 1.0f) * 0.5f * (float)H - 0.5f;
      const float fx = floorf(ix), fy = floorf(iy);
      const int x0 = (int)fx, y0 = (int)fy;
      const float ax = ix - fx, ay = iy - fy;
      float s = 0.0f;
      if (y0 >= 0 && y0 < H) {
        if (x0 >= 0 && x0 < W) s += (1.0f - ax) * (1.0f - ay) * vl[((size_t)y0 * W + x0) * LD];
        if (x0 + 1 >= 0 && x0 + 1 < W) s += ax * (1.0f - ay) * vl[((size_t)y0 * W + x0 + 1) * LD];
      }
      if (y0 + 1 >= 0 && y0 + 1 < H) {
        if (x0 >= 0 && x0 < W) s += (1.0f - ax) * ay * vl[((size_t)(y0 + 1) * W + x0) * LD];
        if (x0 + 1 >= 0 && x0 + 1 < W) s += ax * ay * vl[((size_t)(y0 + 1) * W + x0 + 1) * LD];
      }
      lvl += s * (wv[l * LPOINTS + p] * inv);
    }
    acc += lvl;
  }
  out[row * LD + h * LHD + d] = acc;
}

// exported-model tail: scores = sigmoid(logits) computed in f64 and rounded (the oracle's rows())
__global__ void sigmoid_scores_kernel(const float* __restrict__ logits, float* __restrict__ s, size_t total) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) s[i] = (float)(1.0 / (1.0 + exp(-(double)logits[i])));
}
// rows[b, r] = [class, score, x1, y1, x2, y2] of the r-th best (query, class) pair of image b
__global__ void rows_kernel(const float* __restrict__ sorted_scores, const int32_t* __restrict__ sorted_flat,
                            const float* __restrict__ ref, const float* __restrict__ src_wh, float* __restrict__ rows, int Q,
                            int C, int total) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int r = i % Q, b = i / Q;
  const size_t seg = (size_t)b * Q * C;
  const int flat = sorted_flat[seg + r];
  const int qi = flat / C, ci = flat - qi * C;
  const float* bx = ref + ((size_t)b * Q + qi) * 4;
  const float cx = bx[0], cy = bx[1], bw = bx[2], bh = bx[3];
  const float w = src_wh[2 * b], h = src_wh[2 * b + 1];
  float* o = rows + (size_t)i * 6;
  o[0] = (float)ci;
  o[1] = sorted_scores[seg + r];
  o[2] = (cx - bw / 2.0f) * w;
  o[3] = (cy - bh / 2.0f) * h;
  o[4] = (cx + bw / 2.0f) * w;
  o[5] = (cy + bh / 2.0f) * h;
}

template <typename T>
T* upload(oar_ctx* ctx, const std::vector<T>& v) {
  T* d = ctx->arena.get<T>(v.size() ? v.size() : 1);
  if (!v.empty()) {
    T* h = (T*)ctx->pinned_get(v.size() * sizeof(T));
    memcpy(h, v.data(), v.size() * sizeof(T));
    OAR_CUDA(cudaMemcpyAsync(d, h, v.size() * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
  }
  return d;
}

// segments of `per` keys each, sorted descending, stable (equal keys keep ascending index order)
void sort_segments_desc(oar_ctx* ctx, const float* keys, float* keys_out, int32_t* vals_out, int n_seg, int per) {
  const size_t total = (size_t)n_seg * per;
  int32_t* vals_in = ctx->arena.get<int32_t>(total);
  iota_segments_kernel<<<cdiv(total, 256), 256, 0, ctx->stream>>>(vals_in, per, total);
  std::vector<int> offs(n_seg + 1);
  for (int i = 0; i <= n_seg; ++i) offs[i] = i * per;
  int* d_offs = upload(ctx, offs);
  size_t tmp_bytes = 0;
  cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, tmp_bytes, keys, keys_out, vals_in, vals_out, (int)total, n_seg,
                                                     d_offs, d_offs + 1, 0, 32, ctx->stream);
  void* tmp = ctx->arena.alloc(tmp_bytes);
  Launch l(ctx, "layout_topk_sort", 0, 16.0 * total);
  OAR_CUDA(cub::DeviceSegmentedRadixSort::SortPairsDescending(tmp, tmp_bytes, keys, keys_out, vals_in, vals_out, (int)total,
                                                              n_seg, d_offs, d_offs + 1, 0, 32, ctx->stream));
}

}  // namespace

// memory [B, N, 256] (the encoder's output) -> rows [B, 300, 6] on the device
float* layout_decode(oar_model* head, const float* source, int B, const int* shape_h, const int* shape_w, const float* d_src_wh) {
  oar_ctx* ctx = head->ctx;
  cudaStream_t st = ctx->stream;
  if ((int)head->ops.size() < LH_FIXED + LL_PER_LAYER || ((int)head->ops.size() - LH_FIXED) % LL_PER_LAYER)
    OAR_FAIL(OAR_E_MODEL, "layout head has %zu layers: not models.build_layout_head's table", head->ops.size());
  const int n_layers = ((int)head->ops.size() - LH_FIXED) / LL_PER_LAYER;
  const int C = head->ops[LH_ENC_SCORE].p[7];
  if (head->ops[LH_ENC_OUTPUT].p[6] != LD || C <= 0) OAR_FAIL(OAR_E_MODEL, "layout head widths do not match");
  LevelShapes S;
  int N = 0;
  for (int l = 0; l < LLEVELS; ++l) S.h[l] = shape_h[l], S.w[l] = shape_w[l], S.off[l] = N, N += shape_h[l] * shape_w[l];
  const int Q = LQUERIES;
  if (N < Q) OAR_FAIL(OAR_E_INVALID, "layout input too small: %d tokens for %d queries", N, Q);
  // anchors (model.rs:154-183): logit-space, FLT_MAX where a component leaves (0.01, 0.99); valid mask
  std::vector<float> anchors((size_t)N * 4), valid(N);
  {
    size_t i = 0;
    for (int l = 0; l < LLEVELS; ++l) {
      const float wh = 0.05f * (float)std::pow(2.0, l);
      for (int y = 0; y < S.h[l]; ++y)
        for (int x = 0; x < S.w[l]; ++x, ++i) {
          const float c[4] = {(float)((x + 0.5) / S.w[l]), (float)((y + 0.5) / S.h[l]), wh, wh};
          bool ok = true;
          for (float v : c) ok = ok && v > 0.01f && v < 0.99f;
          valid[i] = ok ? 1.0f : 0.0f;
          for (int k = 0; k < 4; ++k) anchors[i * 4 + k] = ok ? logf(c[k] / (1.0f - c[k])) : FLT_MAX;
        }
    }
  }
  float* d_anchors = upload(ctx, anchors);
  float* d_valid = upload(ctx, valid);
  const size_t rowsN = (size_t)B * N, rowsQ = (size_t)B * Q;
  // ---- encoder output head: memory = LN(Linear(source * valid)); class logits; top-300 on the max logit
  float* masked = ctx->arena.get<float>(rowsN * LD);
  {
    Launch l(ctx, "layout_mask", 0, 8.0 * rowsN * LD);
    mask_rows_kernel<<<cdiv(rowsN * LD / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(source), d_valid,
                                                                reinterpret_cast<float4*>(masked), N, LD / 4, rowsN * LD / 4);
  }
  float* lin0 = ctx->arena.get<float>(rowsN * LD);
  op_linear(head, LH_ENC_OUTPUT, masked, (int)rowsN, lin0);
  float* memory = masked;  // reuse: the masked copy is dead once projected
  op_layernorm(head, LH_ENC_OUTPUT_LN, lin0, (int)rowsN, memory);
  float* enc_class = ctx->arena.get<float>(rowsN * C);
  op_linear(head, LH_ENC_SCORE, memory, (int)rowsN, enc_class);
  float* best = ctx->arena.get<float>(rowsN);
  {
    Launch l(ctx, "layout_rowmax", 0, 4.0 * rowsN * C);
    row_max_kernel<<<cdiv(rowsN, 256), 256, 0, st>>>(enc_class, best, C, rowsN);
  }
  float* best_sorted = ctx->arena.get<float>(rowsN);
  int32_t* top = ctx->arena.get<int32_t>(rowsN);  // per image: token indices by descending score, ties by index
  sort_segments_desc(ctx, best, best_sorted, top, B, N);
  float* hidden = ctx->arena.get<float>(rowsQ * LD);
  {
    Launch l(ctx, "layout_gather", 0, 8.0 * rowsQ * LD);
    gather_rows_kernel<<<cdiv(rowsQ * LD / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(memory), top,
                                                                  reinterpret_cast<float4*>(hidden), N, Q, LD / 4, rowsQ * LD / 4);
  }
  // scratch for the 300-query rows
  float* t0 = ctx->arena.get<float>(rowsQ * 1024);
  float* t1 = ctx->arena.get<float>(rowsQ * 1024);
  float* z4 = ctx->arena.get<float>(rowsQ * 4);
  float* ref = ctx->arena.get<float>(rowsQ * 4);
  float* qpos = ctx->arena.get<float>(rowsQ * LD);
  float* xq = ctx->arena.get<float>(rowsQ * LD);
  float* blk = ctx->arena.get<float>(rowsQ * LD);
  float* offs = ctx->arena.get<float>(rowsQ * LHEADS * LLEVELS * LPOINTS * 2);
  float* wlog = ctx->arena.get<float>(rowsQ * LHEADS * LLEVELS * LPOINTS);
  float* value = ctx->arena.get<float>(rowsN * LD);
  float* logits = ctx->arena.get<float>(rowsQ * C);
  auto add_rows = [&](const float* a, const float* b, float* o) {
    Launch l(ctx, "add", (double)rowsQ * LD, 12.0 * rowsQ * LD);
    add4_kernel<<<cdiv(rowsQ * LD / 4, 256), 256, 0, st>>>(reinterpret_cast<const float4*>(a), reinterpret_cast<const float4*>(b),
                                                           reinterpret_cast<float4*>(o), rowsQ * LD / 4);
  };
  auto bbox_mlp = [&](int first, const float* x) {  // 3 Linear layers, ReLU after the first two -> z4 [rows, 4]
    op_linear(head, first, x, (int)rowsQ, t0);
    op_linear(head, first + 1, t0, (int)rowsQ, t1);
    op_linear(head, first + 2, t1, (int)rowsQ, z4);
  };
  bbox_mlp(LH_ENC_BBOX0, hidden);
  ref_from_anchors_kernel<<<cdiv(rowsQ * 4, 256), 256, 0, st>>>(z4, d_anchors, top, ref, N, Q, (int)(rowsQ * 4));
  // ---- decoder layers (decoder.rs:696-765)
  for (int li = 0; li < n_layers; ++li) {
    const int base = LH_FIXED + li * LL_PER_LAYER;
    op_linear(head, LH_QPOS0, ref, (int)rowsQ, t0);
    op_linear(head, LH_QPOS1, t0, (int)rowsQ, qpos);
    add_rows(hidden, qpos, xq);
    op_attention(head, base + LL_SA, hidden, xq, B, Q, blk);
    add_rows(hidden, blk, t0);
    op_layernorm(head, base + LL_LN1, t0, (int)rowsQ, hidden);
    add_rows(hidden, qpos, xq);
    op_linear(head, base + LL_OFFSETS, xq, (int)rowsQ, offs);
    op_linear(head, base + LL_WEIGHTS, xq, (int)rowsQ, wlog);
    op_linear(head, base + LL_VALUE, source, (int)rowsN, value);
    {
      Launch l(ctx, "layout_deform_attn", 2.0 * rowsQ * LD * LLEVELS * LPOINTS * 4, 4.0 * rowsQ * LD * (LLEVELS * LPOINTS * 4 + 1));
      if (B > 65535) OAR_FAIL(OAR_E_UNSUPPORTED, "layout batch too large for one launch");
      deform_attn_kernel<<<dim3(Q, B), LD, 0, st>>>(value, offs, wlog, ref, blk, N, Q, S);
    }
    op_linear(head, base + LL_OUT, blk, (int)rowsQ, t0);
    add_rows(hidden, t0, t1);
    op_layernorm(head, base + LL_LN2, t1, (int)rowsQ, hidden);
    op_linear(head, base + LL_FC1, hidden, (int)rowsQ, t0);
    op_linear(head, base + LL_FC2, t0, (int)rowsQ, blk);
    add_rows(hidden, blk, t1);
    op_layernorm(head, base + LL_LN3, t1, (int)rowsQ, hidden);
    bbox_mlp(base + LL_BBOX0, hidden);
    refine_ref_kernel<<<cdiv(rowsQ * 4, 256), 256, 0, st>>>(z4, ref, (int)(rowsQ * 4));
    if (li == n_layers - 1) op_linear(head, base + LL_SCORE, hidden, (int)rowsQ, logits);
  }
  // ---- exported-model rows
  float* sc = ctx->arena.get<float>(rowsQ * C);
  sigmoid_scores_kernel<<<cdiv(rowsQ * C, 256), 256, 0, st>>>(logits, sc, rowsQ * C);
  float* sc_sorted = ctx->arena.get<float>(rowsQ * C);
  int32_t* flat_sorted = ctx->arena.get<int32_t>(rowsQ * C);
  sort_segments_desc(ctx, sc, sc_sorted, flat_sorted, B, Q * C);
  float* rows = ctx->arena.get<float>(rowsQ * 6);
  rows_kernel<<<cdiv(rowsQ, 256), 256, 0, st>>>(sc_sorted, flat_sorted, ref, d_src_wh, rows, Q, C, (int)rowsQ);
  OAR_CUDA(cudaGetLastError());
  return rows;
}

// ScaleAwareDetectorModel::preprocess (pp_doclayout) + the network: u8 pages -> rows [n, 300, 6] on the device
float* layout_rows_device(oar_model* enc, oar_model* head, const uint8_t* const* images, const int32_t* hs, const int32_t* ws,
                          int n, int on_device, int in_h, int in_w) {
  oar_ctx* ctx = enc->ctx;
  cudaStream_t st = ctx->stream;
  if (in_h <= 0 || in_w <= 0 || (in_h % 32) || (in_w % 32)) OAR_FAIL(OAR_E_INVALID, "layout input must be a multiple of 32");
  // pages into HBM, resize_exact to (in_h, in_w) with CatmullRom (scale_aware_detector.rs:66-80, resize_detection.rs:337-366)
  std::vector<const uint8_t*> ptrs(n);
  std::vector<ResizeJob> jobs;
  std::vector<float> src_wh(2 * (size_t)n);
  int max_sw = 0;
  size_t total = 0;
  for (int i = 0; i < n; ++i) {
    if (hs[i] <= 0 || ws[i] <= 0 || !images[i]) OAR_FAIL(OAR_E_INVALID, "image %d is empty", i);
    total += ((size_t)hs[i] * ws[i] * 3 + 15) & ~(size_t)15;
  }
  // host pages: page-locked ones are copied from where they lie (one transfer per page), pageable ones go through one
  // pinned staging buffer and one transfer
  bool pinned = !on_device;
  for (int i = 0; i < n && pinned; ++i) {
    cudaPointerAttributes at{};
    if (cudaPointerGetAttributes(&at, images[i]) != cudaSuccess) {
      cudaGetLastError();
      pinned = false;
    } else {
      pinned = at.type == cudaMemoryTypeHost;
    }
  }
  uint8_t* h_pages = (on_device || pinned) ? nullptr : (uint8_t*)ctx->pinned_get(total);
  uint8_t* d_pages = on_device ? nullptr : ctx->arena.get<uint8_t>(total);
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const size_t bytes = (size_t)hs[i] * ws[i] * 3;
    const uint8_t* d = images[i];
    if (!on_device) {
      if (pinned)
        OAR_CUDA(cudaMemcpyAsync(d_pages + off, images[i], bytes, cudaMemcpyHostToDevice, st));
      else
        memcpy(h_pages + off, images[i], bytes);
      d = d_pages + off;
      off += (bytes + 15) & ~(size_t)15;
    }
    src_wh[2 * i] = (float)ws[i], src_wh[2 * i + 1] = (float)hs[i];
    if (hs[i] == in_h && ws[i] == in_w) {
      ptrs[i] = d;
      continue;
    }
    ResizeJob j;
    j.src = d, j.sw = ws[i], j.sh = hs[i], j.dw = in_w, j.dh = in_h, j.filter = 1;
    j.tmp = ctx->arena.get<float>((size_t)in_h * ws[i] * 3);
    j.dst = ctx->arena.get<uint8_t>((size_t)in_h * in_w * 3);
    jobs.push_back(j);
    max_sw = std::max(max_sw, j.sw);
    ptrs[i] = j.dst;
  }
  if (!on_device && !pinned) OAR_CUDA(cudaMemcpyAsync(d_pages, h_pages, total, cudaMemcpyHostToDevice, st));
  if (!jobs.empty()) {
    ResizeJob* d_jobs = upload(ctx, jobs);
    launch_resize_triangle(ctx, d_jobs, (int)jobs.size(), max_sw, in_w, in_h);
  }
  // NormalizeImage: scale 1/255, mean 0, std 1, RGB (alpha = scale / std, beta = -mean / std, normalization.rs:142-143)
  const uint8_t** d_table = (const uint8_t**)upload(ctx, ptrs);
  Tensor x;
  x.B = n, x.H = in_h, x.W = in_w, x.C = 3;
  x.p = ctx->arena.get<float>(x.numel());
  const int src[3] = {0, 1, 2};
  const float alpha[3] = {1.0f / 255.0f, 1.0f / 255.0f, 1.0f / 255.0f}, beta[3] = {0.0f, 0.0f, 0.0f};
  launch_normalize(ctx, nullptr, d_table, false, x.p, n, in_h, in_w, src, alpha, beta, /*NHWC*/ 1);
  Tensor mem = model_forward(enc, x, false, nullptr);
  const int sh[3] = {in_h / 8, in_h / 16, in_h / 32}, sw[3] = {in_w / 8, in_w / 16, in_w / 32};
  const int N = sh[0] * sw[0] + sh[1] * sw[1] + sh[2] * sw[2];
  if (!mem.p || mem.B != n || mem.H != 1 || mem.W != N || mem.C != LD)
    OAR_FAIL(OAR_E_MODEL, "layout encoder produced %dx%dx%dx%d, expected %dx1x%dx%d (was it built for this input size?)", mem.B,
             mem.H, mem.W, mem.C, n, N, LD);
  float* d_src_wh = upload(ctx, src_wh);
  return layout_decode(head, mem.p, n, sh, sw, d_src_wh);
}

}  // namespace oar
