//! `include/oar_b200.h`, declaration by declaration.  Every function returns an `OAR_*` status code; the message of the
//! last failure on the calling thread is `oar_last_error()`.  No exception or panic crosses this boundary.
#![allow(non_camel_case_types, clippy::too_many_arguments)]
use std::os::raw::{c_char, c_void};

#[repr(C)]
pub struct oar_ctx {
    _p: [u8; 0],
}
#[repr(C)]
pub struct oar_model {
    _p: [u8; 0],
}

pub const OAR_OK: i32 = 0;
pub const OAR_E_INVALID: i32 = -1;
pub const OAR_E_NO_DEVICE: i32 = -2;
pub const OAR_E_CUDA: i32 = -3;
pub const OAR_E_MODEL: i32 = -4;
pub const OAR_E_CAPACITY: i32 = -5;
pub const OAR_E_UNSUPPORTED: i32 = -6;

pub const OAR_KIND_DET: i32 = 0;
pub const OAR_KIND_REC: i32 = 1;
pub const OAR_KIND_CLS: i32 = 2;

/// `oar_det_config` = TextDetectionConfig + DBPreprocessConfig (domain/tasks/text_detection.rs:34-67, db.rs:409-415)
#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct oar_det_config {
    pub thresh: f32,
    pub box_thresh: f32,
    pub unclip_ratio: f32,
    pub max_candidates: i32,
    pub min_size: f32,
    pub limit_side_len: i32,
    /// 0 = LimitType::Max, 1 = Min, 2 = ResizeLong (processors/types.rs:50-62)
    pub limit_type: i32,
    pub max_side_limit: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct oar_pipeline_config {
    pub det: oar_det_config,
    pub image_batch_size: i32,
    pub region_batch_size: i32,
    pub rec_score_thresh: f32,
    pub n_chars: i32,
}

/// Caller-owned result buffers of `oar_pipeline_run*` (see the header for the layout).
#[repr(C)]
pub struct oar_ocr_result {
    pub cap_regions: i32,
    pub cap_labels: i32,
    pub region_off: *mut i32,
    pub boxes: *mut f32,
    pub scores: *mut f32,
    pub det_index: *mut i32,
    pub label_off: *mut i32,
    pub labels: *mut i32,
    pub ms_h2d: f32,
    pub ms_det: f32,
    pub ms_post: f32,
    pub ms_crop: f32,
    pub ms_rec: f32,
    pub ms_total: f32,
    pub h2d_bytes: i64,
    pub d2h_bytes: i64,
    pub cols: *mut i32,
    pub seq_len: *mut i32,
    pub wh_ratio: *mut f32,
    pub max_wh_ratio: *mut f32,
    pub line_angle: *mut f32,
    pub ms_cls: f32,
}

#[link(name = "oar_b200")]
unsafe extern "C" {
    pub fn oar_last_error() -> *const c_char;
    pub fn oar_version() -> i32;
    /// kernels launched by the library in this process / host-visible submissions among them (a replayed CUDA graph of a
    /// network batch counts one)
    pub fn oar_launch_count() -> i64;
    pub fn oar_submit_count() -> i64;
    pub fn oar_det_config_default(cfg: *mut oar_det_config);
    pub fn oar_pipeline_config_default(cfg: *mut oar_pipeline_config);

    pub fn oar_ctx_create(device_id: i32, out: *mut *mut oar_ctx) -> i32;
    pub fn oar_ctx_destroy(ctx: *mut oar_ctx);
    pub fn oar_ctx_synchronize(ctx: *mut oar_ctx) -> i32;

    /// `OrtInfer::new(ModelSource::Memory)`: ONNX ModelProto bytes, converted to the layer list behind the ABI.
    /// kind: OAR_KIND_* as the caller's role, or -1 to infer it from the graph.
    pub fn oar_model_load_onnx(ctx: *mut oar_ctx, bytes: *const c_void, len: usize, kind: i32, out: *mut *mut oar_model) -> i32;
    pub fn oar_model_load_blob(ctx: *mut oar_ctx, bytes: *const c_void, len: usize, out: *mut *mut oar_model) -> i32;
    pub fn oar_model_destroy(m: *mut oar_model);
    pub fn oar_model_kind(m: *const oar_model) -> i32;

    /// seam 1: `OrtInfer::infer_first_output_f32` (ort_infer_execution.rs:121-306); host f32 NCHW in, f32 out
    pub fn oar_infer_f32(m: *mut oar_model, input: *const f32, in_shape: *const i64, out: *mut f32, out_cap: usize, out_shape: *mut i64) -> i32;

    /// seam 2: `TextDetectionAdapter::execute` -> `DBModel::forward` (db.rs:281-335)
    pub fn oar_det_run(det: *mut oar_model, images: *const *const u8, hs: *const i32, ws: *const i32, n: i32,
                       cfg: *const oar_det_config, boxes: *mut f32, scores: *mut f32, counts: *mut i32) -> i32;
    /// seam 2: `TextRecognitionAdapter::execute` -> `CRNNModel::forward_refs` (crnn.rs:247-293); ONE batch
    pub fn oar_rec_run(rec: *mut oar_model, crops: *const *const u8, hs: *const i32, ws: *const i32, n: i32, n_chars: i32,
                       labels: *mut i32, cols: *mut i32, lens: *mut i32, scores: *mut f32, t_cap: i32, t_out: *mut i32) -> i32;
    /// pages + boxes -> crops -> recognize_global (transform.rs:76-502, ocr.rs:802-897)
    pub fn oar_crop_rec_run(rec: *mut oar_model, images: *const *const u8, hs: *const i32, ws: *const i32, n: i32,
                            images_on_device: i32, boxes: *const f32, img_index: *const i32, n_boxes: i32,
                            region_batch_size: i32, n_chars: i32, rec_score_thresh: f32, status: *mut i32,
                            labels: *mut i32, cols: *mut i32, lens: *mut i32, scores: *mut f32, seq_len: *mut i32, t_cap: i32) -> i32;
    /// `OAROCR::predict` (src/oarocr/ocr.rs:518-659)
    pub fn oar_pipeline_run(det: *mut oar_model, rec: *mut oar_model, images: *const *const u8, hs: *const i32, ws: *const i32,
                            n: i32, images_on_device: i32, cfg: *const oar_pipeline_config, out: *mut oar_ocr_result) -> i32;
    /// with `with_text_line_orientation_classification` (ocr.rs:197-203, 615, 755-792); `cls` may be null
    pub fn oar_pipeline_run_cls(det: *mut oar_model, rec: *mut oar_model, cls: *mut oar_model, images: *const *const u8,
                                hs: *const i32, ws: *const i32, n: i32, images_on_device: i32,
                                cfg: *const oar_pipeline_config, out: *mut oar_ocr_result) -> i32;
    /// the same predict() over several GPUs of one box: dets[g] / recs[g] on context g
    pub fn oar_pipeline_run_multi(dets: *const *mut oar_model, recs: *const *mut oar_model, n_ctx: i32,
                                  images: *const *const u8, hs: *const i32, ws: *const i32, n: i32,
                                  cfg: *const oar_pipeline_config, out: *mut oar_ocr_result) -> i32;
}
