//! RAII over `oar_ctx` / `oar_model`, and the status-code -> `OCRError` mapping.
use crate::sys;
use oar_ocr_core::core::OCRError;
use oar_ocr_core::core::inference::ModelSource;
use std::ffi::CStr;
use std::sync::Arc;

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum ModelKind {
    Detection = sys::OAR_KIND_DET as isize,
    Recognition = sys::OAR_KIND_REC as isize,
    Classification = sys::OAR_KIND_CLS as isize,
}

/// Message of the last failure on this thread (thread-local inside the library).
pub(crate) fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::oar_last_error()) }.to_string_lossy().into_owned()
}

/// status code -> OCRError (core/errors/types.rs:110-214)
pub(crate) fn check(rc: i32, model_name: &str, context: &str) -> Result<(), OCRError> {
    if rc == sys::OAR_OK {
        return Ok(());
    }
    let msg = last_error();
    Err(match rc {
        sys::OAR_E_INVALID => OCRError::invalid_input(msg),
        sys::OAR_E_MODEL => OCRError::model_load_error(model_name, msg, Some("check that the ONNX file is a PP-OCR det/rec export"), None::<std::io::Error>),
        sys::OAR_E_NO_DEVICE => OCRError::ConfigError { message: format!("B200 provider: {msg}") },
        _ => OCRError::inference_error(model_name, context, std::io::Error::other(msg)),
    })
}

/// One CUDA context + launch stream + activation arena (`oar_ctx`).  Calls on one context are serialised inside the
/// library (as `Mutex<Session>` serialises ORT runs); different contexts run concurrently.
#[derive(Debug)]
pub struct B200Context {
    raw: *mut sys::oar_ctx,
    pub device_id: i32,
}
unsafe impl Send for B200Context {}
unsafe impl Sync for B200Context {}

impl B200Context {
    pub fn new(device_id: i32) -> Result<Arc<Self>, OCRError> {
        let mut raw = std::ptr::null_mut();
        check(unsafe { sys::oar_ctx_create(device_id, &mut raw) }, "B200", "oar_ctx_create")?;
        Ok(Arc::new(Self { raw, device_id }))
    }
    pub(crate) fn raw(&self) -> *mut sys::oar_ctx {
        self.raw
    }
}
impl Drop for B200Context {
    fn drop(&mut self) {
        unsafe { sys::oar_ctx_destroy(self.raw) }
    }
}

/// One network resident in HBM (`oar_model`); stands where `OrtInfer` stands in the reference.
#[derive(Debug)]
pub struct B200Model {
    raw: *mut sys::oar_model,
    ctx: Arc<B200Context>,
    pub name: String,
}
unsafe impl Send for B200Model {}
unsafe impl Sync for B200Model {}

impl B200Model {
    /// `OrtInfer::new` for this provider: `ModelSource::Path` is read, `ModelSource::Memory` is used as is; the bytes
    /// are ONNX (converted behind the ABI) or a pre-converted "OARG" layer list.
    pub fn load(ctx: &Arc<B200Context>, source: &ModelSource, kind: ModelKind, name: &str) -> Result<Self, OCRError> {
        let owned;
        let bytes: &[u8] = match source {
            ModelSource::Memory(b) => b,
            ModelSource::Path(p) => {
                owned = std::fs::read(p).map_err(|e| {
                    OCRError::model_load_error(p, format!("cannot read model file: {e}"), Some("check the path"), Some(e))
                })?;
                &owned
            }
        };
        let mut raw = std::ptr::null_mut();
        let rc = if bytes.starts_with(b"OARG") {
            unsafe { sys::oar_model_load_blob(ctx.raw(), bytes.as_ptr().cast(), bytes.len(), &mut raw) }
        } else {
            unsafe { sys::oar_model_load_onnx(ctx.raw(), bytes.as_ptr().cast(), bytes.len(), kind as i32, &mut raw) }
        };
        check(rc, &source.display_path().to_string_lossy(), "load")?;
        let got = unsafe { sys::oar_model_kind(raw) };
        if got != kind as i32 {
            unsafe { sys::oar_model_destroy(raw) };
            return Err(OCRError::ConfigError { message: format!("model '{name}' is of kind {got}, expected {}", kind as i32) });
        }
        Ok(Self { raw, ctx: Arc::clone(ctx), name: name.to_string() })
    }
    pub(crate) fn raw(&self) -> *mut sys::oar_model {
        self.raw
    }
    pub fn context(&self) -> &Arc<B200Context> {
        &self.ctx
    }
}
impl Drop for B200Model {
    fn drop(&mut self) {
        unsafe { sys::oar_model_destroy(self.raw) }
    }
}
