// prepost.cuh -- image pre/post-processing kernels (declarations).
// All kernels in prepost.cu / dbpost.cu are compiled with -fmad=false: the
// reference computes `x*a + b` as a separate multiply and add
// (oar-ocr-core/src/processors/simd.rs:11-14) and Rust never contracts f32
// expressions, so bit-exact parity needs un-fused arithmetic.
#pragma once
#include "common.cuh"

namespace oar {

// u8 HWC RGB -> f32, out[c] = rgb[src[c]] * alpha[c] + beta[c].
// layout 0: NCHW (the reference tensor, normalization.rs:429-482); 1: NHWC (engine input)
// Source: either one contiguous batch `rgb`, or (d_table != nullptr) a device table of B image pointers
// (table_aligned: every pointer is 4-byte aligned).
void launch_normalize(oar_ctx* ctx, const uint8_t* rgb, const uint8_t* const* d_table, bool table_aligned, float* out,
                      int B, int H, int W, const int src[3], const float alpha[3], const float beta[3], int layout);

// image 0.25 imageops::resize(Triangle) restated: vertical pass to f32, horizontal pass to u8.
// Batched over jobs; each job has its own src/dst pointers and dims.
struct ResizeJob {
  const uint8_t* src;  // [sh][sw][3]
  float* tmp;          // [dh][sw][3]
  uint8_t* dst;        // [dh][dw][3]
  int sw, sh, dw, dh;
  int filter = 0;  // 0 = Triangle (text detection / recognition), 1 = CatmullRom (PP-DocLayout detectors,
                   // scale_aware_detector.rs:66-80): same two-pass sampler, other kernel and support
};
void launch_resize_triangle(oar_ctx* ctx, const ResizeJob* d_jobs, int n_jobs, int max_sw, int max_dw, int max_dh);

// get_rotate_crop_image (transform.rs:76-191)
struct CropPlan {
  // inputs
  float quad[8];
  int img;  // index into the image table
  // outputs of the planning kernel
  int status;       // 0 ok, else the reference returns Err (crop skipped)
  int left, top;    // crop origin
  int cw, ch;       // crop-box dims
  int axis_aligned; // fast path: plain copy
  int rw, rh;       // rectified dims before the tall-rotation
  int rot270;       // h >= 1.5 w
  int ow, oh;       // final dims
  float inv[9];     // inverse homography
  float wh_ratio;   // ow / max(oh,1)  (ocr.rs:739)
  long long out_off; // byte offset into the crop pool (filled by host)
};
struct ImageRef {
  const uint8_t* p;
  int h, w;
};
void launch_crop_plan(oar_ctx* ctx, CropPlan* d_plans, int n, const ImageRef* d_images);
void launch_crop_warp(oar_ctx* ctx, const CropPlan* d_plans, int n, const ImageRef* d_images, uint8_t* pool,
                      long long total_px);

// CRNN normalize (simd.rs:248-308): resized u8 [48][rw][3] -> f32 (v/255-0.5)/0.5 BGR, right pad 0.
// layout 0: NCHW [n,3,48,tw]; 1: NHWC [n,48,tw,3]
struct CrnnJob {
  const uint8_t* src;  // resized crop [48][rw][3]
  int rw;
};
void launch_crnn_normalize(oar_ctx* ctx, const CrnnJob* d_jobs, int n, int img_h, int tensor_w, float* out, int layout);

// CTC: argmax over materialised probabilities (decode.rs:452-501) and collapse (decode.rs:505-614)
void launch_ctc_argmax(oar_ctx* ctx, const float* pred, long long rows, int V, int32_t* idx, float* prob);
void launch_ctc_decode(oar_ctx* ctx, const int32_t* idx, const float* prob, int B, int T, int n_chars, int32_t* labels,
                       int32_t* cols, int32_t* lens, float* scores);

// Text-line orientation stage (src/oarocr/ocr.rs:755-792).
// top-1 of Topk::process (utils/topk.rs): stable descending sort = the FIRST maximal class; probs [n][C].
void launch_cls_top1(oar_ctx* ctx, const float* probs, int n, int C, int32_t* ids, float* scores);
// image::imageops::rotate180 in place on u8 HWC images: pixel i <-> pixel npix-1-i.  class_ids (device, one per job,
// may be null = rotate every job): only jobs whose class id is 1 ("180") are rotated.
struct Rot180Job {
  uint8_t* p;
  int npix;
};
void launch_rotate180(oar_ctx* ctx, const Rot180Job* d_jobs, int n_jobs, int max_npix, const int32_t* class_ids);

// DB post-process (db_postprocess.rs:100-179 + db_bitmap.rs:84-150), batched over images.
// pred [B][H][W] device.  Outputs device arrays: boxes [B][max_cand][8], scores [B][max_cand], counts [B].
struct DbPostOut {
  float* boxes;
  float* scores;
  int32_t* counts;
};
struct DbPostStatus {
  int* h_counters = nullptr;  // pinned: n_comps, n_recs, err (valid after stream sync)
  int launch_comps = 0, sort_n = 0;
};
DbPostStatus db_postprocess_device(oar_ctx* ctx, const float* pred, int B, int H, int W, const int32_t* h_src_h,
                                   const int32_t* h_src_w, const oar_det_config& cfg, DbPostOut out,
                                   int launch_comps_hint);
int db_postprocess_check(const DbPostStatus& s, int* need_comps);

}  // namespace oar
