//! `ModelAdapter` implementations (core/traits/adapter.rs:42-81) over the seam-2 entry points.
//!
//! Same contract as the reference adapters they stand in for:
//! * detection  — text_detection_adapter.rs:36-79: `execute` runs `DBModel::forward` with the effective config's three
//!   thresholds and returns `TextDetectionOutput { detections }`, boxes in source-image coordinates, discovery order;
//! * recognition — text_recognition_adapter.rs:35-111: the whole input is ONE batch; a text whose score is below the
//!   threshold keeps its slot with empty text, empty positions and empty column indices, but keeps score and length.
use crate::model::{check, B200Context, B200Model, ModelKind};
use crate::sys;
use image::RgbImage;
use oar_ocr_core::core::inference::ModelSource;
use oar_ocr_core::core::traits::adapter::{AdapterBuilder, AdapterInfo, ModelAdapter};
use oar_ocr_core::core::traits::task::{ImageTaskInput, Task, TaskType};
use oar_ocr_core::core::OCRError;
use oar_ocr_core::domain::tasks::{
    Detection, TextDetectionConfig, TextDetectionOutput, TextDetectionTask, TextRecognitionConfig,
    TextRecognitionOutput, TextRecognitionTask,
};
use oar_ocr_core::processors::{BoundingBox, LimitType, Point};
use std::sync::Arc;

/// (pointer, height, width) tables of a batch of RGB8 images — what every `*_run` entry point takes.
pub(crate) struct ImageTable {
    pub ptrs: Vec<*const u8>,
    pub hs: Vec<i32>,
    pub ws: Vec<i32>,
}
impl ImageTable {
    pub fn new<'a>(images: impl Iterator<Item = &'a RgbImage>) -> Self {
        let (mut ptrs, mut hs, mut ws) = (Vec::new(), Vec::new(), Vec::new());
        for im in images {
            ptrs.push(im.as_raw().as_ptr()); // RgbImage: contiguous u8 HWC
            hs.push(im.height() as i32);
            ws.push(im.width() as i32);
        }
        Self { ptrs, hs, ws }
    }
    pub fn len(&self) -> usize {
        self.ptrs.len()
    }
}

pub(crate) fn det_config_to_ffi(cfg: &TextDetectionConfig) -> sys::oar_det_config {
    let mut c = std::mem::MaybeUninit::<sys::oar_det_config>::uninit();
    let mut c = unsafe {
        sys::oar_det_config_default(c.as_mut_ptr()); // thresh .3, box .6, unclip 1.5, 1000, min 3, 960 Max 4000
        c.assume_init()
    };
    c.thresh = cfg.score_threshold;
    c.box_thresh = cfg.box_threshold;
    c.unclip_ratio = cfg.unclip_ratio;
    c.max_candidates = cfg.max_candidates as i32;
    if let Some(l) = cfg.limit_side_len {
        c.limit_side_len = l as i32;
    }
    if let Some(t) = &cfg.limit_type {
        c.limit_type = match t {
            LimitType::Max => 0,
            LimitType::Min => 1,
            LimitType::ResizeLong => 2,
        };
    }
    if let Some(m) = cfg.max_side_len {
        c.max_side_limit = m as i32;
    }
    c
}

// ---------------------------------------------------------------------------------------------- detection
#[derive(Debug)]
pub struct B200TextDetectionAdapter {
    model: B200Model,
    info: AdapterInfo,
    config: TextDetectionConfig,
}

impl ModelAdapter for B200TextDetectionAdapter {
    type Task = TextDetectionTask;

    fn info(&self) -> AdapterInfo {
        self.info.clone()
    }

    fn execute(
        &self,
        input: <Self::Task as Task>::Input,
        config: Option<&<Self::Task as Task>::Config>,
    ) -> Result<<Self::Task as Task>::Output, OCRError> {
        let cfg = config.unwrap_or(&self.config);
        let table = ImageTable::new(input.images.iter().map(|a| a.as_ref()));
        let n = table.len();
        if n == 0 {
            return Ok(TextDetectionOutput { detections: Vec::new() });
        }
        let c = det_config_to_ffi(cfg);
        let mc = c.max_candidates.max(1) as usize;
        let mut boxes = vec![0f32; n * mc * 8];
        let mut scores = vec![0f32; n * mc];
        let mut counts = vec![0i32; n];
        let rc = unsafe {
            sys::oar_det_run(self.model.raw(), table.ptrs.as_ptr(), table.hs.as_ptr(), table.ws.as_ptr(), n as i32, &c,
                             boxes.as_mut_ptr(), scores.as_mut_ptr(), counts.as_mut_ptr())
        };
        check(rc, &self.model.name, "oar_det_run").map_err(|e| {
            OCRError::adapter_execution_error(
                "B200TextDetectionAdapter",
                format!("failed to detect text (score_threshold={}, box_threshold={}, unclip_ratio={})",
                        cfg.score_threshold, cfg.box_threshold, cfg.unclip_ratio),
                e,
            )
        })?;
        let detections = (0..n)
            .map(|i| {
                (0..counts[i] as usize)
                    .map(|k| {
                        let p = &boxes[(i * mc + k) * 8..(i * mc + k) * 8 + 8];
                        let quad = (0..4).map(|j| Point::new(p[2 * j], p[2 * j + 1])).collect();
                        Detection::new(BoundingBox::new(quad), scores[i * mc + k])
                    })
                    .collect()
            })
            .collect();
        Ok(TextDetectionOutput { detections })
    }

    fn supports_batching(&self) -> bool {
        true
    }
    fn recommended_batch_size(&self) -> usize {
        8 // text_detection_adapter.rs:85-87
    }
}

#[derive(Debug, Clone)]
pub struct B200TextDetectionAdapterBuilder {
    config: TextDetectionConfig,
    ctx: Option<Arc<B200Context>>,
    device_id: i32,
    model_name: String,
}
impl B200TextDetectionAdapterBuilder {
    pub fn new() -> Self {
        Self { config: TextDetectionConfig::default(), ctx: None, device_id: 0, model_name: "PP-OCRv5_mobile_det".into() }
    }
    /// share one context (stream + arena) with the other adapters of a pipeline
    pub fn context(mut self, ctx: Arc<B200Context>) -> Self {
        self.ctx = Some(ctx);
        self
    }
    pub fn device_id(mut self, id: i32) -> Self {
        self.device_id = id;
        self
    }
    pub fn model_name(mut self, name: impl Into<String>) -> Self {
        self.model_name = name.into();
        self
    }
}
impl Default for B200TextDetectionAdapterBuilder {
    fn default() -> Self {
        Self::new()
    }
}
impl AdapterBuilder for B200TextDetectionAdapterBuilder {
    type Config = TextDetectionConfig;
    type Adapter = B200TextDetectionAdapter;

    fn build(self, model_source: impl Into<ModelSource>) -> Result<Self::Adapter, OCRError> {
        let ctx = match self.ctx {
            Some(c) => c,
            None => B200Context::new(self.device_id)?,
        };
        let model = B200Model::load(&ctx, &model_source.into(), ModelKind::Detection, &self.model_name)?;
        let info = AdapterInfo::new(self.model_name, TaskType::TextDetection,
                                    "Detects text regions in images with bounding boxes (B200 executor)");
        Ok(B200TextDetectionAdapter { model, info, config: self.config })
    }
    fn with_config(mut self, config: Self::Config) -> Self {
        self.config = config;
        self
    }
    fn adapter_type(&self) -> &str {
        "text_detection"
    }
}

// ---------------------------------------------------------------------------------------------- recognition
/// index -> character, exactly as `CTCLabelDecode::from_string_list(dict, use_space_char = true, false)` builds it
/// (processors/decode.rs:118-141, 392-423): blank '\0', the FIRST char of every non-empty dictionary line, ' '.
pub fn ctc_character_list(dict: &[String]) -> Vec<char> {
    let mut chars = vec!['\0'];
    chars.extend(dict.iter().filter_map(|s| s.chars().next()));
    chars.push(' ');
    chars
}

#[derive(Debug)]
pub struct B200TextRecognitionAdapter {
    model: B200Model,
    info: AdapterInfo,
    config: TextRecognitionConfig,
    characters: Vec<char>,
    return_word_box: bool,
}

impl B200TextRecognitionAdapter {
    pub fn characters(&self) -> &[char] {
        &self.characters
    }
    pub(crate) fn model(&self) -> &B200Model {
        &self.model
    }
}

impl ModelAdapter for B200TextRecognitionAdapter {
    type Task = TextRecognitionTask;

    fn info(&self) -> AdapterInfo {
        self.info.clone()
    }

    fn execute(
        &self,
        input: <Self::Task as Task>::Input,
        config: Option<&<Self::Task as Task>::Config>,
    ) -> Result<<Self::Task as Task>::Output, OCRError> {
        let cfg = config.unwrap_or(&self.config);
        let table = ImageTable::new(input.images.iter().map(|a| a.as_ref()));
        let n = table.len();
        if n == 0 {
            return Ok(TextRecognitionOutput::empty());
        }
        // T = tensor width / 8; the tensor is at most 3200 wide (crnn.rs:71-125), so 402 covers every batch
        let t_cap: usize = 3200 / 8 + 2;
        let mut labels = vec![0i32; n * t_cap];
        let mut cols = vec![0i32; n * t_cap];
        let mut lens = vec![0i32; n];
        let mut scores = vec![0f32; n];
        let mut t_out = 0i32;
        let rc = unsafe {
            sys::oar_rec_run(self.model.raw(), table.ptrs.as_ptr(), table.hs.as_ptr(), table.ws.as_ptr(), n as i32,
                             self.characters.len() as i32, labels.as_mut_ptr(), cols.as_mut_ptr(), lens.as_mut_ptr(),
                             scores.as_mut_ptr(), t_cap as i32, &mut t_out)
        };
        check(rc, &self.model.name, "oar_rec_run").map_err(|e| {
            OCRError::adapter_execution_error(
                "B200TextRecognitionAdapter",
                format!("forward (batch_size={}, return_word_box={})", n, self.return_word_box),
                e,
            )
        })?;
        let seq_len = t_out as usize;
        let mut out = TextRecognitionOutput::with_capacity(n);
        for i in 0..n {
            let len = lens[i] as usize;
            let row = &labels[i * t_cap..i * t_cap + len];
            let col = &cols[i * t_cap..i * t_cap + len];
            out.scores.push(scores[i]);
            if scores[i] >= cfg.score_threshold {
                // decode.rs:503-614: indices outside the dictionary are skipped, never a panic
                out.texts.push(row.iter().filter_map(|&k| self.characters.get(k as usize)).filter(|&&c| c != '\0').collect());
                if self.return_word_box {
                    // decode_argmax_with_positions (decode.rs:541-614): timestep / T
                    out.char_positions.push(col.iter().map(|&c| c as f32 / seq_len.max(1) as f32).collect());
                    out.char_col_indices.push(col.iter().map(|&c| c as usize).collect());
                    out.sequence_lengths.push(seq_len);
                } else {
                    // the reference's position-free decode returns no columns and a zero length (adapter :66-85)
                    out.char_positions.push(Vec::new());
                    out.char_col_indices.push(Vec::new());
                    out.sequence_lengths.push(0);
                }
            } else {
                // below the threshold: the slot stays, the text goes (text_recognition_adapter.rs:88-102)
                out.texts.push(String::new());
                out.char_positions.push(Vec::new());
                out.char_col_indices.push(Vec::new());
                out.sequence_lengths.push(if self.return_word_box { seq_len } else { 0 });
            }
        }
        Ok(out)
    }

    fn supports_batching(&self) -> bool {
        true
    }
    fn recommended_batch_size(&self) -> usize {
        64 // text_recognition_adapter.rs:117-129
    }
}

#[derive(Debug, Clone)]
pub struct B200TextRecognitionAdapterBuilder {
    config: TextRecognitionConfig,
    character_dict: Option<Vec<String>>,
    return_word_box: bool,
    ctx: Option<Arc<B200Context>>,
    device_id: i32,
    model_name: String,
}
impl B200TextRecognitionAdapterBuilder {
    pub fn new() -> Self {
        Self { config: TextRecognitionConfig::default(), character_dict: None, return_word_box: false, ctx: None,
               device_id: 0, model_name: "PP-OCRv5_mobile_rec".into() }
    }
    pub fn character_dict(mut self, dict: Vec<String>) -> Self {
        self.character_dict = Some(dict);
        self
    }
    pub fn score_thresh(mut self, t: f32) -> Self {
        self.config.score_threshold = t;
        self
    }
    pub fn return_word_box(mut self, enable: bool) -> Self {
        self.return_word_box = enable;
        self
    }
    pub fn context(mut self, ctx: Arc<B200Context>) -> Self {
        self.ctx = Some(ctx);
        self
    }
    pub fn device_id(mut self, id: i32) -> Self {
        self.device_id = id;
        self
    }
    pub fn model_name(mut self, name: impl Into<String>) -> Self {
        self.model_name = name.into();
        self
    }
}
impl Default for B200TextRecognitionAdapterBuilder {
    fn default() -> Self {
        Self::new()
    }
}
impl AdapterBuilder for B200TextRecognitionAdapterBuilder {
    type Config = TextRecognitionConfig;
    type Adapter = B200TextRecognitionAdapter;

    fn build(self, model_source: impl Into<ModelSource>) -> Result<Self::Adapter, OCRError> {
        let dict = self.character_dict.ok_or_else(|| OCRError::ConfigError {
            message: "Character dictionary is required for text recognition".into(), // crnn.rs builder
        })?;
        let ctx = match self.ctx {
            Some(c) => c,
            None => B200Context::new(self.device_id)?,
        };
        let model = B200Model::load(&ctx, &model_source.into(), ModelKind::Recognition, &self.model_name)?;
        let info = AdapterInfo::new(self.model_name, TaskType::TextRecognition,
                                    "Recognizes text content from image regions (B200 executor)");
        Ok(B200TextRecognitionAdapter { model, info, config: self.config, characters: ctc_character_list(&dict),
                                        return_word_box: self.return_word_box })
    }
    fn with_config(mut self, config: Self::Config) -> Self {
        self.config = config;
        self
    }
    fn adapter_type(&self) -> &str {
        "text_recognition"
    }
}

impl B200TextDetectionAdapter {
    pub(crate) fn model(&self) -> &B200Model {
        &self.model
    }
    pub(crate) fn config(&self) -> &TextDetectionConfig {
        &self.config
    }
}
impl B200TextRecognitionAdapter {
    pub(crate) fn config(&self) -> &TextRecognitionConfig {
        &self.config
    }
}

#[cfg(test)]
mod tests {
    use super::*;

    #[test]
    fn character_list_follows_from_string_list() {
        // decode.rs:118-141: first char of each line, empty lines dropped; blank first, space last
        let dict = vec!["a".to_string(), "".to_string(), "bc".to_string(), "\u{2028}".to_string()];
        assert_eq!(ctc_character_list(&dict), vec!['\0', 'a', 'b', '\u{2028}', ' ']);
    }
}
