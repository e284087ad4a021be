//! `OAROCR::predict` (src/oarocr/ocr.rs:518-659) on the B200 executor: ONE call of `oar_pipeline_run` for the whole
//! image list.  Detection chunks of `image_batch_size`, `sort_quad_boxes`, `get_rotate_crop_image`, crop pooling with
//! the 4096-crop flush, the stable wh-ratio sort and `region_batch_size` chunking of `recognize_global`
//! (ocr.rs:550-633, 802-897) all happen behind the ABI; this file only rebuilds `Vec<OAROCRResult>`.
//!
//! With several contexts (one per GPU of the box) the same call goes through `oar_pipeline_run_multi`: one host thread
//! and CUDA context per GPU inside the library, a single global `recognize_global` plan, results identical to one GPU.
use crate::adapters::{
    det_config_to_ffi, B200TextDetectionAdapter, B200TextDetectionAdapterBuilder, B200TextRecognitionAdapter,
    B200TextRecognitionAdapterBuilder, ImageTable,
};
use crate::model::{check, B200Context};
use crate::sys;
use image::RgbImage;
use oar_ocr_core::core::inference::ModelSource;
use oar_ocr_core::core::traits::adapter::AdapterBuilder;
use oar_ocr_core::core::OCRError;
use oar_ocr_core::domain::tasks::{TextDetectionConfig, TextRecognitionConfig};
use oar_ocr_core::domain::text_region::TextRegion;
use oar_ocr_core::processors::{BoundingBox, Point};
use std::sync::Arc;

/// What `OAROCR::predict` returns per image (src/oarocr/result.rs:34-49), minus the optional document stages.
#[derive(Debug, Clone)]
pub struct B200OcrResult {
    pub input_path: Arc<str>,
    pub index: usize,
    pub input_img: Arc<RgbImage>,
    pub text_regions: Vec<TextRegion>,
}

#[derive(Debug)]
struct Replica {
    det: B200TextDetectionAdapter,
    rec: B200TextRecognitionAdapter,
}

#[derive(Debug)]
pub struct B200Ocr {
    replicas: Vec<Replica>, // one per GPU
    image_batch_size: usize,
    region_batch_size: usize,
}

pub struct B200OcrBuilder {
    det: ModelSource,
    rec: ModelSource,
    dict: Vec<String>,
    det_cfg: Option<TextDetectionConfig>,
    rec_cfg: Option<TextRecognitionConfig>,
    devices: Vec<i32>,
    image_batch_size: Option<usize>,
    region_batch_size: Option<usize>,
}

impl B200OcrBuilder {
    /// `OAROCRBuilder::new(det, rec, dict)` (ocr.rs:70-110): `dict` = the lines of the character dictionary file.
    pub fn new(det: impl Into<ModelSource>, rec: impl Into<ModelSource>, dict: Vec<String>) -> Self {
        Self { det: det.into(), rec: rec.into(), dict, det_cfg: None, rec_cfg: None, devices: vec![0],
               image_batch_size: None, region_batch_size: None }
    }
    pub fn text_detection_config(mut self, c: TextDetectionConfig) -> Self {
        self.det_cfg = Some(c);
        self
    }
    pub fn text_recognition_config(mut self, c: TextRecognitionConfig) -> Self {
        self.rec_cfg = Some(c);
        self
    }
    /// GPUs of this box to spread one predict() over (default: device 0 only)
    pub fn devices(mut self, ids: &[i32]) -> Self {
        self.devices = ids.to_vec();
        self
    }
    pub fn image_batch_size(mut self, n: usize) -> Self {
        self.image_batch_size = Some(n);
        self
    }
    pub fn region_batch_size(mut self, n: usize) -> Self {
        self.region_batch_size = Some(n);
        self
    }
    pub fn build(self) -> Result<B200Ocr, OCRError> {
        // ocr.rs:1168-1195
        if self.image_batch_size == Some(0) || self.region_batch_size == Some(0) {
            return Err(OCRError::ConfigError { message: "batch sizes must be at least 1".into() });
        }
        if self.devices.is_empty() {
            return Err(OCRError::ConfigError { message: "B200 provider: no device given".into() });
        }
        // no explicit config -> thresh .3 / box .6 / unclip 2.0 / limit 960 Max 4000 (ocr.rs:351-364)
        let det_cfg = self.det_cfg.unwrap_or(TextDetectionConfig {
            unclip_ratio: 2.0,
            limit_side_len: Some(960),
            max_side_len: Some(4000),
            ..TextDetectionConfig::default()
        });
        let rec_cfg = self.rec_cfg.unwrap_or_default();
        let mut replicas = Vec::new();
        for &d in &self.devices {
            let ctx = B200Context::new(d)?;
            let det = B200TextDetectionAdapterBuilder::new().context(Arc::clone(&ctx)).with_config(det_cfg.clone())
                .build(self.det.clone())?;
            let rec = B200TextRecognitionAdapterBuilder::new().context(ctx).with_config(rec_cfg.clone())
                .character_dict(self.dict.clone()).build(self.rec.clone())?;
            replicas.push(Replica { det, rec });
        }
        // an accelerator provider: adapter defaults 8 / 64 (src/oarocr/builder_utils.rs:86-125)
        Ok(B200Ocr { replicas, image_batch_size: self.image_batch_size.unwrap_or(8),
                     region_batch_size: self.region_batch_size.unwrap_or(64) })
    }
}

impl B200Ocr {
    pub fn predict(&self, images: Vec<RgbImage>) -> Result<Vec<B200OcrResult>, OCRError> {
        if images.is_empty() {
            // ocr.rs:525-532
            return Err(OCRError::invalid_input("images: expected non-empty slice, got empty slice"));
        }
        let images: Vec<Arc<RgbImage>> = images.into_iter().map(Arc::new).collect();
        let table = ImageTable::new(images.iter().map(|a| a.as_ref()));
        let n = table.len();
        let head = &self.replicas[0];
        let det_cfg = head.det.config();
        let chars = head.rec.characters();
        let cfg = sys::oar_pipeline_config {
            det: det_config_to_ffi(det_cfg),
            image_batch_size: self.image_batch_size as i32,
            region_batch_size: self.region_batch_size as i32,
            rec_score_thresh: head.rec.config().score_threshold,
            n_chars: chars.len() as i32,
        };
        // every detection may become a region; a text line has at most 3200 / 8 timesteps
        let cap_regions = n * det_cfg.max_candidates;
        let cap_labels = cap_regions * 64;
        let mut region_off = vec![0i32; n + 1];
        let mut boxes = vec![0f32; cap_regions * 8];
        let mut scores = vec![0f32; cap_regions];
        let mut det_index = vec![0i32; cap_regions];
        let mut label_off = vec![0i32; cap_regions + 1];
        let mut labels = vec![0i32; cap_labels];
        let mut out = sys::oar_ocr_result {
            cap_regions: cap_regions as i32,
            cap_labels: cap_labels as i32,
            region_off: region_off.as_mut_ptr(),
            boxes: boxes.as_mut_ptr(),
            scores: scores.as_mut_ptr(),
            det_index: det_index.as_mut_ptr(),
            label_off: label_off.as_mut_ptr(),
            labels: labels.as_mut_ptr(),
            ms_h2d: 0.0, ms_det: 0.0, ms_post: 0.0, ms_crop: 0.0, ms_rec: 0.0, ms_total: 0.0,
            h2d_bytes: 0, d2h_bytes: 0,
            cols: std::ptr::null_mut(), seq_len: std::ptr::null_mut(), wh_ratio: std::ptr::null_mut(),
            max_wh_ratio: std::ptr::null_mut(), line_angle: std::ptr::null_mut(), ms_cls: 0.0,
        };
        let rc = if self.replicas.len() == 1 {
            unsafe {
                sys::oar_pipeline_run(head.det.model().raw(), head.rec.model().raw(), table.ptrs.as_ptr(), table.hs.as_ptr(),
                                      table.ws.as_ptr(), n as i32, 0, &cfg, &mut out)
            }
        } else {
            let dets: Vec<_> = self.replicas.iter().map(|r| r.det.model().raw()).collect();
            let recs: Vec<_> = self.replicas.iter().map(|r| r.rec.model().raw()).collect();
            unsafe {
                sys::oar_pipeline_run_multi(dets.as_ptr(), recs.as_ptr(), dets.len() as i32, table.ptrs.as_ptr(),
                                            table.hs.as_ptr(), table.ws.as_ptr(), n as i32, &cfg, &mut out)
            }
        };
        check(rc, "PP-OCRv5 det+rec", "oar_pipeline_run")?;

        let mut results = Vec::with_capacity(n);
        for (i, img) in images.into_iter().enumerate() {
            let mut text_regions = Vec::new();
            for r in region_off[i] as usize..region_off[i + 1] as usize {
                let p = &boxes[r * 8..r * 8 + 8];
                let bbox = BoundingBox::new((0..4).map(|j| Point::new(p[2 * j], p[2 * j + 1])).collect());
                let lab = &labels[label_off[r] as usize..label_off[r + 1] as usize];
                let text: String = lab.iter().filter_map(|&k| chars.get(k as usize)).collect();
                // ocr.rs:879-892
                text_regions.push(TextRegion {
                    bounding_box: bbox.clone(),
                    dt_poly: Some(bbox.clone()),
                    rec_poly: Some(bbox),
                    text: Some(Arc::from(text)),
                    confidence: Some(scores[r]),
                    orientation_angle: None,
                    word_boxes: None,
                    label: None,
                });
            }
            results.push(B200OcrResult { input_path: Arc::from(format!("image_{i}")), index: i, input_img: img, text_regions });
        }
        Ok(results)
    }
}
