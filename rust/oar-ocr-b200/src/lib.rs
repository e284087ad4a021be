//! B200 executor for oar-ocr.
//!
//! `liboar_b200.so` replaces, for the det+rec hot path, what `ort::Session` + the CPU pre/post-processing do in
//! `oar-ocr-core`: the crate's traits stay as they are and this crate implements them.
//!
//! * [`sys`]      — the `extern "C"` declarations of `include/oar_b200.h`, one to one.
//! * [`model`]    — `B200Context` / `B200Model`: RAII over `oar_ctx` / `oar_model`; a model is built from the same
//!                  `ModelSource` (`Path` or `Memory` ONNX bytes) `OrtInfer::new` takes
//!                  (`core/inference/ort_infer_builders.rs:9-70`, `model_source.rs:20-28`).
//! * [`adapters`] — `B200TextDetectionAdapter`, `B200TextRecognitionAdapter`: `ModelAdapter` implementations
//!                  (`core/traits/adapter.rs:42-81`) with the semantics of `text_detection_adapter.rs:36-79` and
//!                  `text_recognition_adapter.rs:35-111`; they drop into `TaskPredictorCore`
//!                  (`predictors/core.rs:19-26`) unchanged.
//! * [`pipeline`] — `B200Ocr::predict`: `OAROCR::predict` (`src/oarocr/ocr.rs:518-659`) as ONE call of
//!                  `oar_pipeline_run` (or `oar_pipeline_run_multi` over several GPUs).
//!
//! The provider switch a maintainer adds to `oar-ocr-core` is in `patches/`.
pub mod adapters;
pub mod model;
pub mod pipeline;
pub mod sys;

pub use adapters::{
    B200TextDetectionAdapter, B200TextDetectionAdapterBuilder, B200TextRecognitionAdapter,
    B200TextRecognitionAdapterBuilder,
};
pub use model::{B200Context, B200Model, ModelKind};
pub use pipeline::{B200Ocr, B200OcrBuilder};
