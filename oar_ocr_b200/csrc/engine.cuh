// engine.cuh -- OARG layer-list executor (declarations)
#pragma once
#include "common.cuh"
#include "prepost.cuh"

namespace oar {

enum OpType {
  OP_CONV = 1,
  OP_DWCONV,
  OP_SE,
  OP_ADD,
  OP_UPADD,
  OP_UPSAMPLE,
  OP_DECONV2,
  OP_AVGPOOL,
  OP_LAYERNORM,
  OP_ATTN,
  OP_CTC_HEAD,
  OP_PAD,      // zero padding p = {top, left, bottom, right}: HGNetV2 stem (hgnetv2.rs:264-348)
  OP_MAXPOOL,  // p = {kh, kw, sh, sw}, no padding
  OP_TOKENS    // a map's pixels as rows [p[0], p[0] + H*W) of a [B, 1, p[1], C] sequence (layout detector memory)
};
enum Act { ACT_NONE = 0, ACT_RELU, ACT_HSWISH, ACT_SWISH, ACT_SIGMOID, ACT_HSIGMOID, ACT_GELU /* exact, erf */ };

struct OpRec {
  int32_t type, in0, in1, out;
  int32_t p[12];
  float f[4];
  int64_t w_off[4];
  int64_t w_len[4];
};
static_assert(sizeof(OpRec) == 144, "OpRec layout");

// NHWC activation.  ld = channel stride of the underlying buffer (>= C when the
// tensor is a channel slice target of a concat).
struct Tensor {
  float* p = nullptr;
  int B = 0, H = 0, W = 0, C = 0;
  size_t numel() const { return (size_t)B * H * W * C; }
};

// Implicit-GEMM description shared by the SIMT (engine.cu) and tcgen05 (gemm_tc.cu) kernels:
//   C[M,N] = A[M,K] * W[N,K]^T,  M = B*Ho*Wo, K = kh*kw*Cin ordered (ky,kx,ci)
struct ConvParams {
  const float* in;
  const float* w;
  const float* bias;
  float* out;
  int B, H, W, Cin, Ho, Wo, kh, kw, sh, sw, ph, pw;
  int N, K, M;
  int out_ld, out_c_off;
  int act;
  float post_scale, post_bias;
  int mode;  // 0 conv, 1 deconv2x2 scatter (N = 4*Cout), 2 CTC partial softmax/argmax (tensor-core engine only)
  int cout;  // deconv: real output channels
  // mode 2 outputs: per (row, n-tile) running max / last arg-max / sum of exp(z - max)
  float* part_max;
  int32_t* part_idx;
  float* part_sum;
};

// One PP-LCNetV3 block for the fused tensor-core kernel (fused_tc.cu): [depthwise k x k -> act] -> 1x1 conv -> act.
// k == 0: no depthwise stage (plain 1x1 conv), optionally with a squeeze-excite multiplier on its input.
struct FusedBlock {
  const float* in;  // NHWC input of the depthwise conv (k > 0) or of the 1x1 conv (k == 0)
  int B, H, W, C;
  int k, sh, sw;  // depthwise kernel size (square, pad k/2) and strides
  int dw_key;  // op index of the depthwise conv (its packed taps live in the tensor-core state)
  int dw_act;
  float dw_ps, dw_pb;
  const float* se_scale;  // k == 0 only: [B][C] multiplier or null
  const float* bias;      // 1x1 conv
  int act;
  float ps, pb;
  int N;
  float* out;
  int out_ld, out_c_off, Ho, Wo;
};

// The network input still as u8 pixels (engines >= 1 fold the normalisation into the stem convolution, fused_simt.cu):
// mode 0 = pages behind a device pointer table with NormalizeImage coefficients (normalization.rs:142-143),
// mode 1 = resized crops with the CRNN normalisation and zero padding right of each crop's rw (simd.rs:248-308).
struct U8Input {
  int mode;
  const uint8_t* const* table;  // mode 0: B image pointers, u8 HWC, row stride 3 * W
  const CrnnJob* jobs;          // mode 1: B crops {src, rw}, row stride 3 * rw
  int B, H, W;                  // the tensor the network sees: [B, H, W, 3]
  int src[3];
  float a[3], b[3];
  int table_aligned;
};

// Output of the fused CTC head: per (b,t) argmax class and its softmax prob.
struct CtcOut {
  int32_t* idx = nullptr;
  float* prob = nullptr;
  int B = 0, T = 0, V = 0;
};

}  // namespace oar

struct oar_model {
  oar_ctx* ctx = nullptr;
  int kind = 0;
  int engine = 0;  // 0 fp32 SIMT, 1 tcgen05 per layer, 2 tcgen05 with fused depthwise->pointwise blocks
  int n_tensors = 0;
  std::vector<oar::OpRec> ops;
  float* d_weights = nullptr;  // all weights, fp32, resident in HBM
  size_t n_weights = 0;
  // tensor-core engine state (gemm_tc.cu): fp16 copies of GEMM weights etc.
  void* tc_state = nullptr;
  uint64_t uid = 0;     // unique per loaded model: names its captured CUDA graphs (an address can be reused, a uid cannot)
  int graph_safe = -1;  // may a walk of this layer list be captured?  -1 = not decided yet (engine.cu)

  const float* w(const oar::OpRec& op, int i) const { return d_weights + op.w_off[i]; }
};

namespace oar {

// Runs the graph on `in` (NHWC f32, C=3).  For det returns the [B,H,W,1]
// probability map.  For rec: when `want_probs` the full [B,1,T,V] softmax is
// materialised in the returned tensor, otherwise only `ctc` is filled (the
// logits never leave the head kernel's launch) and the returned tensor is empty.
// `u8` (optional): the input is given as u8 pixels instead of `in.p` (in.p == nullptr, dims from u8).
//
// CUDA graphs: a walk that reads u8 pixels (the detector, recogniser and classifier of the pipeline) is recorded into a
// CUDA graph the second time its (model, engine, launch lane, batch, height, width) comes up and replayed from then on:
// ~100 kernel launches, their tensor-map encodes and argument marshalling become one cudaGraphLaunch.  A graph owns its
// activations (one cudaMalloc sized by the first, eager walk), so its outputs -- the returned tensor and `ctc` -- stay
// valid until the next walk with the same key on the same lane: consume them in stream order.  OAR_GRAPHS=0 turns it
// off; profiling (per-kernel events) always walks eagerly.
Tensor model_forward(oar_model* m, const Tensor& in, bool want_probs, CtcOut* ctc, const U8Input* u8 = nullptr);
void graph_cache_free(oar_ctx* ctx);                      // every graph of a context (oar_ctx_destroy)
void graph_cache_drop_model(oar_ctx* ctx, uint64_t uid);  // the graphs of one model (oar_model_destroy)

// Single layers of a loaded model, run outside a graph walk (engine.cu; the layout decoder in layout_net.cu calls them by
// position): 1x1 convolution as a Linear over `rows` rows, LayerNorm, multi-head attention over B sequences of T tokens
// (x_qk: the input with positions added for the q / k projections, or null).
void op_linear(oar_model* m, int oi, const float* in, int rows, float* out, int out_ld = 0, int out_off = 0);
void op_layernorm(oar_model* m, int oi, const float* in, int rows, float* out);
void op_attention(oar_model* m, int oi, const float* x, const float* x_qk, int B, int T, float* out);

// layout_net.cu: ScaleAwareDetectorModel::preprocess (pp_doclayout) + RT-DETR-L on the device.  Pages in (host, or device pointers), the
// exported model's rows [n, 300, 6] = [class_id, score, x1, y1, x2, y2] (source pixels) out, in the context's arena.
float* layout_rows_device(oar_model* enc, oar_model* head, const uint8_t* const* images, const int32_t* hs, const int32_t* ws,
                          int n, int on_device, int in_h, int in_w);

// jpeg_ingest.cu: JPEG streams -> u8 HWC RGB pages in the context's arena (nvJPEG, looked up at first use)
void decode_jpegs_to_device(oar_ctx* ctx, const uint8_t* const* data, const size_t* lens, int n, const uint8_t** ptrs,
                            int32_t* hs, int32_t* ws);
bool jpeg_ingest_available();

// fused_simt.cu: false = shape not covered, the caller runs the per-layer path
bool launch_stem_u8(oar_ctx* ctx, const U8Input& S, const OpRec& op, const float* w, const float* bias, float* out, int Ho,
                    int Wo);
bool launch_deconv_pair(oar_ctx* ctx, const float* in, int B, int H, int W, const OpRec& d1, const float* w1,
                        const float* b1, const OpRec& d2, const float* w2, const float* b2, float* out);

// tensor-core engine lifetime hooks (gemm_tc.cu): build fp16 weight copies / release them
void tc_model_init(oar_model* m);
void tc_model_free(oar_model* m);
// Launches the tcgen05 kernel for the GEMM identified by `key` (op index * 2 + sub); false if this
// model has no packed weights for it (the caller then runs the SIMT kernel).
bool tc_gemm(oar_model* m, int key, const ConvParams& p, const char* name);
// number of N tiles the tensor-core engine uses for `key` (0 if absent); sizes the mode-2 partials
int tc_n_tiles(const oar_model* m, int key);
// Fused block on the persistent tcgen05 kernel (fused_tc.cu); false if the shape is not supported (caller falls back to
// the per-layer kernels).  `key` names the 1x1 conv's packed weights.
bool tc_fused_block(oar_model* m, int key, const FusedBlock& f, const char* name);
// Dense stride-1 k x k convolution (k <= 3, Cin % 32 == 0) on the persistent TMA-halo kernel (conv_halo_tc.cu); false ->
// the caller uses tc_gemm (row-taps / im2col kernels)
bool tc_conv_halo(oar_model* m, int key, const ConvParams& p, const char* name);
bool tc_conv_fold(oar_model* m, int key, const ConvParams& p, const char* name);  // narrow 3 x 3: kernel columns folded into N
// Stand-alone depthwise k x k conv on TMA-staged shared-memory tiles (dw_tma.cu); false -> the register-tiled kernel.
// tile_sums (optional): per-(image, tile) channel sums for the squeeze-excite pool, *n_tiles = tiles per image
bool tc_dw_tma(oar_model* m, int dw_key, const float* in, float* out, int B, int H, int W, int C, int Ho, int Wo, int k, int sh,
               int sw, int act, float ps, float pb, float* tile_sums, int* n_tiles, const char* name);
// CTC head (mode-2 ConvParams: part_* outputs) on the persistent kernel; false -> caller uses tc_gemm
bool tc_ctc_head_persistent(oar_model* m, int key, const ConvParams& p, const char* name);
void launch_ctc_combine(oar_ctx* ctx, const float* part_max, const int32_t* part_idx, const float* part_sum, size_t rows,
                        int n_tiles, int32_t* idx, float* prob);

// ONNX ModelProto bytes -> OARG blob (onnx_import.cu; host only).  kind_hint < 0: inferred from the graph's tail.
std::vector<uint8_t> onnx_to_oarg(const void* bytes, size_t len, int kind_hint);

// input layout conversion for the seam-1 API (NCHW f32 host layout -> NHWC)
void launch_nchw_to_nhwc(oar_ctx* ctx, const float* in, float* out, int B, int C, int H, int W);

}  // namespace oar
