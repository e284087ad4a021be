"""ctypes binding of liboar_b200.so (include/oar_b200.h).

This is the same C ABI a Rust shim binds; Python is used here
because the test and bench harnesses are Python.  There is no fallback: if the
library is missing, or no sm_100 device is usable, calls raise OCRError.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboar_b200.so")

OAR_OK, OAR_E_INVALID, OAR_E_NO_DEVICE, OAR_E_CUDA, OAR_E_MODEL, OAR_E_CAPACITY, OAR_E_UNSUPPORTED = 0, -1, -2, -3, -4, -5, -6
KIND_DET, KIND_REC, KIND_CLS, KIND_FEAT = 0, 1, 2, 3

# every symbol include/oar_b200.h declares (tests check the built library exports all of them)
SYMBOLS = [
    "oar_det_config_default", "oar_pipeline_config_default", "oar_last_error", "oar_version", "oar_launch_count",
    "oar_submit_count",
    "oar_ctx_create", "oar_ctx_destroy", "oar_ctx_synchronize", "oar_model_load_blob", "oar_model_destroy",
    "oar_model_kind", "oar_model_set_engine", "oar_infer_f32", "oar_normalize_chw", "oar_db_postprocess",
    "oar_det_run", "oar_sort_quad_boxes", "oar_rotate_crop", "oar_crnn_preprocess", "oar_ctc_decode", "oar_rec_run",
    "oar_pipeline_run", "oar_cls_run", "oar_rotate180", "oar_pipeline_run_cls", "oar_layout_config_default",
    "oar_layout_postprocess", "oar_device_alloc", "oar_device_free", "oar_memcpy_h2d", "oar_profile_enable",
    "oar_profile_read", "oar_timer_start", "oar_timer_stop", "oar_l2_flush", "oar_model_validate_blob",
    "oar_model_load_onnx", "oar_onnx_to_oarg", "oar_crop_rec_run", "oar_rec_run_ex", "oar_pipeline_run_multi",
    "oar_layout_rows", "oar_layout_run", "oar_pipeline_run_encoded", "oar_decode_jpeg",
]


class OCRError(RuntimeError):
    """Mirror of oar-ocr-core's OCRError (core/errors/types.rs:110-214): `kind` names the variant."""

    def __init__(self, kind: str, message: str, code: int = 0):
        super().__init__(f"{kind}: {message}")
        self.kind = kind
        self.code = code


_KINDS = {
    OAR_E_INVALID: "InvalidInput",
    OAR_E_NO_DEVICE: "Inference",
    OAR_E_CUDA: "Inference",
    OAR_E_MODEL: "ModelLoad",
    OAR_E_CAPACITY: "Inference",
    OAR_E_UNSUPPORTED: "ConfigError",
}


class DetConfig(C.Structure):
    _fields_ = [("thresh", C.c_float), ("box_thresh", C.c_float), ("unclip_ratio", C.c_float),
                ("max_candidates", C.c_int32), ("min_size", C.c_float), ("limit_side_len", C.c_int32),
                ("limit_type", C.c_int32), ("max_side_limit", C.c_int32)]


class PipelineConfig(C.Structure):
    _fields_ = [("det", DetConfig), ("image_batch_size", C.c_int32), ("region_batch_size", C.c_int32),
                ("rec_score_thresh", C.c_float), ("n_chars", C.c_int32)]


class OcrResult(C.Structure):
    _fields_ = [("cap_regions", C.c_int32), ("cap_labels", C.c_int32), ("region_off", C.POINTER(C.c_int32)),
                ("boxes", C.POINTER(C.c_float)), ("scores", C.POINTER(C.c_float)),
                ("det_index", C.POINTER(C.c_int32)), ("label_off", C.POINTER(C.c_int32)),
                ("labels", C.POINTER(C.c_int32)), ("ms_h2d", C.c_float), ("ms_det", C.c_float),
                ("ms_post", C.c_float), ("ms_crop", C.c_float), ("ms_rec", C.c_float), ("ms_total", C.c_float),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("cols", C.POINTER(C.c_int32)),
                ("seq_len", C.POINTER(C.c_int32)), ("wh_ratio", C.POINTER(C.c_float)),
                ("max_wh_ratio", C.POINTER(C.c_float)), ("line_angle", C.POINTER(C.c_float)),
                ("ms_cls", C.c_float)]


class LayoutConfig(C.Structure):
    """oar_layout_config = LayoutDetectionConfig with its label-keyed maps resolved to class ids"""
    _fields_ = [("score_threshold", C.c_float), ("max_elements", C.c_int32), ("layout_nms", C.c_int32),
                ("num_classes", C.c_int32), ("class_thresholds", C.POINTER(C.c_float)),
                ("class_merge_modes", C.POINTER(C.c_int32)), ("image_class_id", C.c_int32),
                ("formula_class_id", C.c_int32), ("unclip_mode", C.c_int32), ("unclip_w", C.c_float),
                ("unclip_h", C.c_float), ("class_unclip", C.POINTER(C.c_float))]


MERGE_UNSET, MERGE_LARGE, MERGE_SMALL, MERGE_UNION = -1, 0, 1, 2
UNCLIP_NONE, UNCLIP_RATIO, UNCLIP_PER_CLASS = 0, 1, 2


class KernelRecord(C.Structure):
    _fields_ = [("name", C.c_char_p), ("ms", C.c_float), ("flops", C.c_double), ("bytes", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        path = os.environ.get("OAR_B200_LIB", LIB_PATH)  # development: A/B a differently built library
        if not os.path.exists(path):
            raise OCRError("ConfigError", f"{path} is not built; run `python -m oar_ocr_b200.build` "
                           "(there is no CPU fallback)")
        L = C.CDLL(path)
        L.oar_last_error.restype = C.c_char_p
        L.oar_launch_count.restype = C.c_int64
        L.oar_submit_count.restype = C.c_int64
        L.oar_ctx_destroy.restype = None
        L.oar_model_destroy.restype = None
        L.oar_det_config_default.restype = None
        L.oar_pipeline_config_default.restype = None
        L.oar_ctx_destroy.argtypes = [C.c_void_p]
        L.oar_model_destroy.argtypes = [C.c_void_p]
        L.oar_ctx_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
        L.oar_ctx_synchronize.argtypes = [C.c_void_p]
        L.oar_model_load_blob.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.oar_model_validate_blob.argtypes = [C.c_void_p, C.c_size_t]
        L.oar_model_load_onnx.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int32, C.POINTER(C.c_void_p)]
        L.oar_onnx_to_oarg.argtypes = [C.c_void_p, C.c_size_t, C.c_int32, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.c_size_t)]
        L.oar_model_kind.argtypes = [C.c_void_p]
        L.oar_model_set_engine.argtypes = [C.c_void_p, C.c_int32]
        L.oar_infer_f32.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_int64), C.c_void_p, C.c_size_t,
                                    C.POINTER(C.c_int64)]
        L.oar_normalize_chw.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                        C.c_void_p, C.c_void_p, C.c_void_p]
        L.oar_db_postprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p,
                                         C.c_void_p, C.POINTER(DetConfig), C.c_void_p, C.c_void_p, C.c_void_p]
        L.oar_det_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(DetConfig),
                                  C.c_void_p, C.c_void_p, C.c_void_p]
        L.oar_sort_quad_boxes.argtypes = [C.c_void_p, C.c_int32, C.c_void_p]
        L.oar_rotate_crop.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                      C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.oar_crnn_preprocess.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p,
                                          C.c_size_t, C.POINTER(C.c_int32)]
        L.oar_ctc_decode.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32] + \
            [C.c_void_p] * 6
        L.oar_rec_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.oar_rec_run_ex.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.POINTER(C.c_int32)]
        L.oar_crop_rec_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p,
                                       C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float] + [C.c_void_p] * 6 + \
            [C.c_int32]
        L.oar_layout_rows.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                      C.c_int32, C.c_int32, C.c_void_p, C.c_size_t]
        L.oar_layout_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32,
                                     C.c_int32, C.POINTER(LayoutConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oar_pipeline_run_encoded.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                               C.POINTER(PipelineConfig), C.POINTER(OcrResult), C.c_void_p, C.c_void_p]
        L.oar_decode_jpeg.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.POINTER(C.c_int32),
                                      C.POINTER(C.c_int32)]
        L.oar_pipeline_run_multi.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p,
                                             C.c_int32, C.POINTER(PipelineConfig), C.POINTER(OcrResult)]
        L.oar_pipeline_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32,
                                       C.c_int32, C.POINTER(PipelineConfig), C.POINTER(OcrResult)]
        L.oar_cls_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32,
                                  C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_int32)]
        L.oar_rotate180.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.oar_pipeline_run_cls.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                           C.c_int32, C.c_int32, C.POINTER(PipelineConfig), C.POINTER(OcrResult)]
        L.oar_layout_config_default.restype = None
        L.oar_layout_config_default.argtypes = [C.POINTER(LayoutConfig)]
        L.oar_layout_postprocess.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p,
                                             C.POINTER(LayoutConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oar_device_alloc.argtypes = [C.c_void_p, C.c_size_t, C.POINTER(C.c_void_p)]
        L.oar_device_free.argtypes = [C.c_void_p, C.c_void_p]
        L.oar_memcpy_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        L.oar_timer_start.argtypes = [C.c_void_p]
        L.oar_timer_stop.argtypes = [C.c_void_p, C.POINTER(C.c_float)]
        L.oar_l2_flush.argtypes = [C.c_void_p]
        L.oar_profile_enable.argtypes = [C.c_void_p, C.c_int32]
        L.oar_profile_read.argtypes = [C.c_void_p, C.POINTER(KernelRecord), C.c_int32]
        _lib = L
    return _lib


def check(rc: int):
    if rc != OAR_OK:
        msg = lib().oar_last_error().decode("utf-8", "replace")
        raise OCRError(_KINDS.get(rc, "Inference"), msg, rc)


def launch_count() -> int:
    return int(lib().oar_launch_count())


def submit_count() -> int:
    """Host-visible submissions (a replayed CUDA graph counts one)."""
    return int(lib().oar_submit_count())


def det_config(**kw) -> DetConfig:
    cfg = DetConfig()
    lib().oar_det_config_default(C.byref(cfg))
    for k, v in kw.items():
        setattr(cfg, k, v)
    return cfg


def pipeline_config(**kw) -> PipelineConfig:
    cfg = PipelineConfig()
    lib().oar_pipeline_config_default(C.byref(cfg))
    for k, v in kw.items():
        if hasattr(cfg.det, k) and not hasattr(cfg, k):
            setattr(cfg.det, k, v)
        else:
            setattr(cfg, k, v)
    return cfg


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _image_table(images):
    """images: list of HxWx3 u8 arrays -> (kept arrays, pointer array, hs, ws)"""
    arrs = []
    for im in images:
        a = np.ascontiguousarray(im, dtype=np.uint8)
        if a.ndim != 3 or a.shape[2] != 3:
            raise OCRError("InvalidInput", f"expected HxWx3 u8 image, got shape {a.shape}", OAR_E_INVALID)
        arrs.append(a)
    ptrs = (C.c_void_p * max(len(arrs), 1))(*[a.ctypes.data for a in arrs])
    hs = np.array([a.shape[0] for a in arrs], np.int32)
    ws = np.array([a.shape[1] for a in arrs], np.int32)
    return arrs, ptrs, hs, ws


class Context:
    """One device + stream + arena = oar_ctx.  One per GPU (one per process under torchrun)."""

    def __init__(self, device_id: int = 0):
        self.handle = C.c_void_p()
        check(lib().oar_ctx_create(device_id, C.byref(self.handle)))
        self.device_id = device_id

    def close(self):
        if self.handle:
            lib().oar_ctx_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def synchronize(self):
        check(lib().oar_ctx_synchronize(self.handle))

    # ---- row 2
    def normalize_chw(self, rgb: np.ndarray, alpha, beta, src=(2, 1, 0)) -> np.ndarray:
        rgb = np.ascontiguousarray(rgb, np.uint8)
        if rgb.ndim == 3:
            rgb = rgb[None]
        b, h, w, _ = rgb.shape
        out = np.empty((b, 3, h, w), np.float32)
        src = np.asarray(src, np.int32)
        alpha = np.asarray(alpha, np.float32)
        beta = np.asarray(beta, np.float32)
        check(lib().oar_normalize_chw(self.handle, _ptr(rgb), b, h, w, _ptr(src), _ptr(alpha), _ptr(beta), _ptr(out)))
        return out

    # ---- rows 4-9
    def db_postprocess(self, pred: np.ndarray, src_hw=None, cfg: DetConfig | None = None):
        pred = np.ascontiguousarray(pred, np.float32)
        if pred.ndim == 2:
            pred = pred[None]
        b, h, w = pred.shape
        cfg = cfg or det_config()
        if src_hw is None:
            src_hw = [(h, w)] * b
        sh = np.array([s[0] for s in src_hw], np.int32)
        sw = np.array([s[1] for s in src_hw], np.int32)
        mc = cfg.max_candidates
        boxes = np.zeros((b, mc, 4, 2), np.float32)
        scores = np.zeros((b, mc), np.float32)
        counts = np.zeros(b, np.int32)
        check(lib().oar_db_postprocess(self.handle, _ptr(pred), b, h, w, _ptr(sh), _ptr(sw), C.byref(cfg), _ptr(boxes),
                                       _ptr(scores), _ptr(counts)))
        return [(boxes[i, :counts[i]].copy(), scores[i, :counts[i]].copy()) for i in range(b)]

    # ---- row 11
    def rotate_crop(self, image: np.ndarray, quads: np.ndarray):
        """returns a list with one HxWx3 crop per quad, None where the reference returns Err"""
        image = np.ascontiguousarray(image, np.uint8)
        quads = np.ascontiguousarray(quads, np.float32).reshape(-1, 8)
        n = len(quads)
        if n == 0:
            return []
        h, w, _ = image.shape
        ow = np.zeros(n, np.int32)
        oh = np.zeros(n, np.int32)
        st = np.zeros(n, np.int32)
        check(lib().oar_rotate_crop(self.handle, _ptr(image), h, w, _ptr(quads), n, _ptr(ow), _ptr(oh), _ptr(st), None,
                                    0))
        total = int((ow.astype(np.int64) * oh).sum()) * 3
        buf = np.empty(max(total, 1), np.uint8)
        check(lib().oar_rotate_crop(self.handle, _ptr(image), h, w, _ptr(quads), n, _ptr(ow), _ptr(oh), _ptr(st),
                                    _ptr(buf), buf.size))
        out, off = [], 0
        for i in range(n):
            if st[i] != 0:
                out.append(None)
                continue
            sz = int(ow[i]) * int(oh[i]) * 3
            out.append(buf[off:off + sz].reshape(oh[i], ow[i], 3).copy())
            off += sz
        return out

    # ---- row 13
    def crnn_preprocess(self, crops) -> np.ndarray:
        if not crops:
            return np.zeros((0, 0, 0, 0), np.float32)
        arrs, ptrs, hs, ws = _image_table(crops)
        tw = C.c_int32()
        check(lib().oar_crnn_preprocess(self.handle, ptrs, _ptr(hs), _ptr(ws), len(arrs), None, 0, C.byref(tw)))
        out = np.empty((len(arrs), 3, 48, tw.value), np.float32)
        check(lib().oar_crnn_preprocess(self.handle, ptrs, _ptr(hs), _ptr(ws), len(arrs), _ptr(out), out.size,
                                        C.byref(tw)))
        return out

    # ---- rows 15-16
    def ctc_decode(self, pred: np.ndarray, n_chars: int):
        pred = np.ascontiguousarray(pred, np.float32)
        b, t, v = pred.shape
        if pred.size == 0:  # decode.rs:464-476: no batch entries when any dimension is zero
            return dict(idx=np.zeros((0, 0), np.int32), prob=np.zeros((0, 0), np.float32), labels=[], cols=[],
                        scores=np.zeros(0, np.float32), T=0)
        idx = np.zeros((b, t), np.int32)
        prob = np.zeros((b, t), np.float32)
        labels = np.zeros((b, max(t, 1)), np.int32)
        cols = np.zeros((b, max(t, 1)), np.int32)
        lens = np.zeros(max(b, 1), np.int32)
        scores = np.zeros(max(b, 1), np.float32)
        check(lib().oar_ctc_decode(self.handle, _ptr(pred), b, t, v, n_chars, _ptr(idx), _ptr(prob), _ptr(labels),
                                   _ptr(cols), _ptr(lens), _ptr(scores)))
        return dict(idx=idx, prob=prob, labels=[labels[i, :lens[i]].copy() for i in range(b)],
                    cols=[cols[i, :lens[i]].copy() for i in range(b)], scores=scores[:b].copy(), T=t)

    # ---- line orientation: image::imageops::rotate180 (ocr.rs:785-788)
    def rotate180(self, image: np.ndarray) -> np.ndarray:
        image = np.ascontiguousarray(image, np.uint8)
        if image.ndim != 3 or image.shape[2] != 3:
            raise OCRError("InvalidInput", f"expected HxWx3 u8 image, got shape {image.shape}", OAR_E_INVALID)
        out = np.empty_like(image)
        check(lib().oar_rotate180(self.handle, _ptr(image), image.shape[0], image.shape[1], _ptr(out)))
        return out

    def profile(self, on: bool):
        check(lib().oar_profile_enable(self.handle, 1 if on else 0))

    def profile_read(self, cap: int = 65536):
        recs = (KernelRecord * cap)()
        n = lib().oar_profile_read(self.handle, recs, cap)
        return [dict(name=recs[i].name.decode(), ms=recs[i].ms, flops=recs[i].flops, bytes=recs[i].bytes)
                for i in range(n)]

    def timer_start(self):
        check(lib().oar_timer_start(self.handle))

    def timer_stop(self) -> float:
        ms = C.c_float()
        check(lib().oar_timer_stop(self.handle, C.byref(ms)))
        return ms.value

    def l2_flush(self):
        check(lib().oar_l2_flush(self.handle))

    def device_alloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(lib().oar_device_alloc(self.handle, nbytes, C.byref(p)))
        return p.value

    def device_free(self, p: int):
        check(lib().oar_device_free(self.handle, C.c_void_p(p)))

    def memcpy_h2d(self, dst: int, src: np.ndarray):
        check(lib().oar_memcpy_h2d(self.handle, C.c_void_p(dst), _ptr(src), src.nbytes))


def layout_postprocess(pred: np.ndarray, src_wh, num_classes: int, score_threshold=0.5, max_elements=100,
                       layout_nms=True, class_thresholds=None, class_merge_modes=None, image_class_id=-1,
                       formula_class_id=-1, unclip=None):
    """oar_layout_postprocess (host only, no device).  pred [B,N,F]; src_wh: (w, h) per image; class_thresholds /
    class_merge_modes: {class_id: value}; unclip: None | (w, h) | {class_id: (w, h)}.
    Returns per image (boxes [n,4] x1 y1 x2 y2, classes [n], scores [n])."""
    pred = np.ascontiguousarray(pred, np.float32)
    if pred.ndim != 3:
        raise OCRError("InvalidInput", "predictions must be [batch, boxes, features]", OAR_E_INVALID)
    b, n, f = pred.shape
    cfg = LayoutConfig()
    lib().oar_layout_config_default(C.byref(cfg))
    cfg.score_threshold, cfg.max_elements, cfg.layout_nms, cfg.num_classes = score_threshold, max_elements, \
        1 if layout_nms else 0, num_classes
    cfg.image_class_id, cfg.formula_class_id = image_class_id, formula_class_id
    keep = []  # arrays the struct points into
    if class_thresholds is not None:
        a = np.full(num_classes, np.nan, np.float32)
        for k, v in class_thresholds.items():
            if 0 <= k < num_classes:
                a[k] = v
        keep.append(a)
        cfg.class_thresholds = a.ctypes.data_as(C.POINTER(C.c_float))
    if class_merge_modes is not None:
        a = np.full(num_classes, MERGE_UNSET, np.int32)
        for k, v in class_merge_modes.items():
            if 0 <= k < num_classes:
                a[k] = v
        keep.append(a)
        cfg.class_merge_modes = a.ctypes.data_as(C.POINTER(C.c_int32))
    if isinstance(unclip, dict):
        a = np.full((num_classes, 2), np.nan, np.float32)
        for k, v in unclip.items():
            if 0 <= k < num_classes:
                a[k] = v
        keep.append(a)
        cfg.unclip_mode = UNCLIP_PER_CLASS
        cfg.class_unclip = a.ctypes.data_as(C.POINTER(C.c_float))
    elif unclip is not None:
        cfg.unclip_mode, cfg.unclip_w, cfg.unclip_h = UNCLIP_RATIO, float(unclip[0]), float(unclip[1])
    sw = np.array([w for w, _ in src_wh], np.float32)
    sh = np.array([h for _, h in src_wh], np.float32)
    me = max(int(max_elements), 1)
    boxes = np.zeros((max(b, 1), me, 4), np.float32)
    classes = np.zeros((max(b, 1), me), np.int32)
    scores = np.zeros((max(b, 1), me), np.float32)
    counts = np.zeros(max(b, 1), np.int32)
    check(lib().oar_layout_postprocess(_ptr(pred), b, n, f, _ptr(sw), _ptr(sh), C.byref(cfg), _ptr(boxes),
                                       _ptr(classes), _ptr(scores), _ptr(counts)))
    return [(boxes[i, :counts[i]].copy(), classes[i, :counts[i]].copy(), scores[i, :counts[i]].copy())
            for i in range(b)]


def sort_quad_boxes(boxes: np.ndarray):
    """sort_quad_boxes (sorting.rs:35-84): returns (sorted boxes, order)"""
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4, 2).copy()
    order = np.zeros(max(len(boxes), 1), np.int32)
    check(lib().oar_sort_quad_boxes(_ptr(boxes), len(boxes), _ptr(order)))
    return boxes, order[:len(boxes)].copy()


def validate_blob(blob: bytes) -> None:
    """oar_model_validate_blob: structural check of an OARG blob, no device needed; raises OCRError("ModelLoad")"""
    buf = (C.c_char * max(len(blob), 1)).from_buffer_copy(blob or b"\0")
    check(lib().oar_model_validate_blob(buf, len(blob)))


def onnx_to_oarg(data: bytes, kind_hint: int = -1) -> bytes:
    """oar_onnx_to_oarg: the library's own ONNX -> OARG conversion (csrc/onnx_import.cu), no device needed"""
    buf = (C.c_char * max(len(data), 1)).from_buffer_copy(data or b"\0")
    n = C.c_size_t(0)
    check(lib().oar_onnx_to_oarg(buf, len(data), kind_hint, None, 0, C.byref(n)))
    out = (C.c_char * max(n.value, 1))()
    check(lib().oar_onnx_to_oarg(buf, len(data), kind_hint, out, n.value, C.byref(n)))
    return bytes(out[:n.value])


class Model:
    """One network resident in HBM = oar_model (stands where OrtInfer stands in the reference).  `data` is what the
    reference's ModelSource carries: ONNX ModelProto bytes, or an OARG layer-list blob; ONNX is converted behind the
    C ABI (oar_model_load_onnx).  `kind` states the caller's role (KIND_DET / KIND_REC / KIND_CLS) or -1."""

    def __init__(self, ctx: Context, data: bytes, kind: int = -1):
        self.ctx = ctx
        self.handle = C.c_void_p()
        buf = (C.c_char * max(len(data), 1)).from_buffer_copy(data or b"\0")
        if bytes(data[:4]) == b"OARG":
            check(lib().oar_model_load_blob(ctx.handle, buf, len(data), C.byref(self.handle)))
        else:
            check(lib().oar_model_load_onnx(ctx.handle, buf, len(data), kind, C.byref(self.handle)))
        self.kind = lib().oar_model_kind(self.handle)

    def close(self):
        if self.handle:
            lib().oar_model_destroy(self.handle)
            self.handle = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_engine(self, engine: int):
        check(lib().oar_model_set_engine(self.handle, engine))

    def infer(self, x: np.ndarray, vocab_hint: int = 18385, out_cap: int | None = None) -> np.ndarray:
        """OrtInfer::infer: x f32 [B,3,H,W] -> det [B,1,H,W] / rec [B,T,V] / cls [B,classes] / feature extractor
        [B,C,H,W] (out_cap: capacity in floats for a feature extractor, default = the input's size)"""
        x = np.ascontiguousarray(x, np.float32)
        if x.ndim != 4:
            raise OCRError("InvalidInput", "input must be 4-D", OAR_E_INVALID)
        b, c, h, w = x.shape
        cap = b * h * w if self.kind == KIND_DET else (b * 4096 if self.kind == KIND_CLS else
                                                       b * (w // 8 + 2) * vocab_hint)
        if self.kind == KIND_FEAT:
            cap = out_cap or b * h * w * 16
        out = np.empty(max(cap, 1), np.float32)
        ishape = (C.c_int64 * 4)(b, c, h, w)
        oshape = (C.c_int64 * 4)()
        check(lib().oar_infer_f32(self.handle, _ptr(x), ishape, _ptr(out), out.size, oshape))
        if self.kind == KIND_DET:
            return out[:oshape[0] * oshape[1] * oshape[2] * oshape[3]].reshape(oshape[0], oshape[1], oshape[2],
                                                                               oshape[3])
        if self.kind == KIND_FEAT:  # stored NHWC; returned in the reference's NCHW
            return np.ascontiguousarray(out[:oshape[0] * oshape[1] * oshape[2] * oshape[3]].reshape(
                oshape[0], oshape[1], oshape[2], oshape[3]).transpose(0, 3, 1, 2))
        y = out[:oshape[0] * oshape[1] * oshape[2]].reshape(oshape[0], oshape[1], oshape[2])
        return y.reshape(oshape[0], oshape[2]) if self.kind == KIND_CLS else y

    def det_run(self, images, cfg: DetConfig | None = None):
        cfg = cfg or det_config()
        if not images:
            return []
        arrs, ptrs, hs, ws = _image_table(images)
        n, mc = len(arrs), cfg.max_candidates
        boxes = np.zeros((n, mc, 4, 2), np.float32)
        scores = np.zeros((n, mc), np.float32)
        counts = np.zeros(n, np.int32)
        check(lib().oar_det_run(self.handle, ptrs, _ptr(hs), _ptr(ws), n, C.byref(cfg), _ptr(boxes), _ptr(scores),
                                _ptr(counts)))
        return [(boxes[i, :counts[i]].copy(), scores[i, :counts[i]].copy()) for i in range(n)]

    def rec_run(self, crops, n_chars: int):
        if not crops:
            return dict(labels=[], cols=[], scores=np.zeros(0, np.float32), T=0)
        arrs, ptrs, hs, ws = _image_table(crops)
        n = len(arrs)
        t_cap = int(max(320, min(3200, max(48.0 * a.shape[1] / a.shape[0] for a in arrs) + 48))) // 8 + 2
        labels = np.zeros((n, t_cap), np.int32)
        cols = np.zeros((n, t_cap), np.int32)
        lens = np.zeros(n, np.int32)
        scores = np.zeros(n, np.float32)
        t_out = C.c_int32()
        check(lib().oar_rec_run(self.handle, ptrs, _ptr(hs), _ptr(ws), n, n_chars, _ptr(labels), _ptr(cols),
                                _ptr(lens), _ptr(scores), t_cap, C.byref(t_out)))
        return dict(labels=[labels[i, :lens[i]].copy() for i in range(n)],
                    cols=[cols[i, :lens[i]].copy() for i in range(n)], scores=scores, T=t_out.value)


    def rec_run_device(self, crop_ptrs, hs: np.ndarray, ws: np.ndarray, n_chars: int, t_cap: int):
        """oar_rec_run_ex with crops already resident in HBM (bench `value` leg of the recognizer-only workload);
        returns (lens, scores, T) -- labels stay in the caller-sized scratch"""
        n = len(hs)
        labels = np.zeros((n, t_cap), np.int32)
        cols = np.zeros((n, t_cap), np.int32)
        lens = np.zeros(n, np.int32)
        scores = np.zeros(n, np.float32)
        t_out = C.c_int32()
        check(lib().oar_rec_run_ex(self.handle, crop_ptrs, _ptr(hs), _ptr(ws), n, 1, n_chars, _ptr(labels), _ptr(cols),
                                   _ptr(lens), _ptr(scores), t_cap, C.byref(t_out)))
        return dict(labels=[labels[i, :lens[i]].copy() for i in range(n)], scores=scores, T=t_out.value)

    def crop_rec_run(self, images, boxes: np.ndarray, img_index, n_chars: int, region_batch_size: int = 64,
                     rec_score_thresh: float = 0.0):
        """oar_crop_rec_run: pages + boxes -> crops -> recognize_global.  Returns dict(status, labels, cols, scores,
        seq_len), one entry per box."""
        arrs, ptrs, hs, ws = _image_table(images)
        boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 8)
        idx = np.ascontiguousarray(img_index, np.int32)
        nb = len(boxes)
        t_cap = 3200 // 8 + 2
        status = np.zeros(nb, np.int32)
        labels = np.zeros((nb, t_cap), np.int32)
        cols = np.zeros((nb, t_cap), np.int32)
        lens = np.zeros(nb, np.int32)
        scores = np.zeros(nb, np.float32)
        seq = np.zeros(nb, np.int32)
        check(lib().oar_crop_rec_run(self.handle, ptrs, _ptr(hs), _ptr(ws), len(arrs), 0, _ptr(boxes), _ptr(idx), nb,
                                     region_batch_size, n_chars, rec_score_thresh, _ptr(status), _ptr(labels),
                                     _ptr(cols), _ptr(lens), _ptr(scores), _ptr(seq), t_cap))
        return dict(status=status, labels=[labels[i, :lens[i]].copy() for i in range(nb)],
                    cols=[cols[i, :lens[i]].copy() for i in range(nb)], scores=scores, seq_len=seq)

    def cls_run(self, crops, input_shape=(80, 160), want_probs=True):
        """TextLineOrientationAdapter::execute: returns dict(class_ids [n], scores [n], probs [n,C] or None)"""
        if not crops:
            return dict(class_ids=np.zeros(0, np.int32), scores=np.zeros(0, np.float32),
                        probs=np.zeros((0, 0), np.float32) if want_probs else None)
        arrs, ptrs, hs, ws = _image_table(crops)
        n = len(arrs)
        ids = np.zeros(n, np.int32)
        scores = np.zeros(n, np.float32)
        nc = C.c_int32()
        cap = n * 1024
        probs = np.zeros(cap, np.float32) if want_probs else None
        check(lib().oar_cls_run(self.handle, ptrs, _ptr(hs), _ptr(ws), n, int(input_shape[0]), int(input_shape[1]),
                                _ptr(ids), _ptr(scores), _ptr(probs) if want_probs else None, cap, C.byref(nc)))
        return dict(class_ids=ids, scores=scores,
                    probs=probs[:n * nc.value].reshape(n, nc.value).copy() if want_probs else None)


class PipelineBuffers:
    """Caller-owned result buffers for oar_pipeline_run, reusable across calls."""

    def __init__(self, n_images: int, cap_regions: int = 0, cap_labels: int = 0):
        cap_regions = cap_regions or n_images * 1000
        cap_labels = cap_labels or cap_regions * 64
        self.region_off = np.zeros(n_images + 1, np.int32)
        self.boxes = np.zeros((cap_regions, 4, 2), np.float32)
        self.scores = np.zeros(cap_regions, np.float32)
        self.det_index = np.zeros(cap_regions, np.int32)
        self.label_off = np.zeros(cap_regions + 1, np.int32)
        self.labels = np.zeros(cap_labels, np.int32)
        self.res = OcrResult()
        self.res.cap_regions = cap_regions
        self.res.cap_labels = cap_labels
        P = C.POINTER
        self.res.region_off = self.region_off.ctypes.data_as(P(C.c_int32))
        self.res.boxes = self.boxes.ctypes.data_as(P(C.c_float))
        self.res.scores = self.scores.ctypes.data_as(P(C.c_float))
        self.res.det_index = self.det_index.ctypes.data_as(P(C.c_int32))
        self.res.label_off = self.label_off.ctypes.data_as(P(C.c_int32))
        self.res.labels = self.labels.ctypes.data_as(P(C.c_int32))
        # word-box inputs (ocr.rs:827-868): CTC column per emitted character, T / wh_ratio / batch max ratio per region
        self.cols = np.zeros(cap_labels, np.int32)
        self.seq_len = np.zeros(cap_regions, np.int32)
        self.wh_ratio = np.zeros(cap_regions, np.float32)
        self.max_wh_ratio = np.zeros(cap_regions, np.float32)
        self.res.cols = self.cols.ctypes.data_as(P(C.c_int32))
        self.res.seq_len = self.seq_len.ctypes.data_as(P(C.c_int32))
        self.res.wh_ratio = self.wh_ratio.ctypes.data_as(P(C.c_float))
        self.res.max_wh_ratio = self.max_wh_ratio.ctypes.data_as(P(C.c_float))
        # TextRegion.orientation_angle of the line-orientation stage: 0 / 180, -1 = None (no classifier)
        self.line_angle = np.full(cap_regions, -1.0, np.float32)
        self.res.line_angle = self.line_angle.ctypes.data_as(P(C.c_float))


def layout_rows(encoder: "Model", head: "Model", images, input_hw=(640, 640), device_table=None) -> np.ndarray:
    """oar_layout_rows: pages -> the layout detector's output tensor [n, 300, 6] = [class_id, score, x1, y1, x2, y2].
    device_table = (ptrs, hs, ws): pages already resident in HBM instead of `images`."""
    if device_table is not None:
        ptrs, hs, ws = device_table
        on_device = 1
    else:
        arrs, ptrs, hs, ws = _image_table(images)
        on_device = 0
    n = len(hs)
    rows = np.zeros((n, 300, 6), np.float32)
    check(lib().oar_layout_rows(encoder.handle, head.handle, ptrs, _ptr(hs), _ptr(ws), n, on_device, int(input_hw[0]),
                                int(input_hw[1]), _ptr(rows), rows.size))
    return rows


def pipeline_run_encoded(det: "Model", rec: "Model", jpegs: list, cfg: PipelineConfig, bufs: "PipelineBuffers"):
    """oar_pipeline_run_encoded: JPEG byte strings in, decoded by nvJPEG into HBM; returns (result, [(h, w), ...])"""
    n = len(jpegs)
    keep = [(C.c_char * len(j)).from_buffer_copy(j) for j in jpegs]
    ptrs = (C.c_void_p * n)(*[C.addressof(k) for k in keep])
    lens = (C.c_size_t * n)(*[len(j) for j in jpegs])
    hs = np.zeros(n, np.int32)
    ws = np.zeros(n, np.int32)
    check(lib().oar_pipeline_run_encoded(det.handle, rec.handle, ptrs, lens, n, C.byref(cfg), C.byref(bufs.res), _ptr(hs),
                                         _ptr(ws)))
    return bufs.res, list(zip(hs.tolist(), ws.tolist()))


def decode_jpeg(ctx: "Context", jpeg: bytes) -> np.ndarray:
    """oar_decode_jpeg: one JPEG stream decoded on the device (nvJPEG), returned as u8 [h, w, 3] RGB"""
    buf = (C.c_char * len(jpeg)).from_buffer_copy(jpeg)
    h, w = C.c_int32(), C.c_int32()
    check(lib().oar_decode_jpeg(ctx.handle, buf, len(jpeg), None, 0, C.byref(h), C.byref(w)))
    out = np.empty((h.value, w.value, 3), np.uint8)
    check(lib().oar_decode_jpeg(ctx.handle, buf, len(jpeg), _ptr(out), out.size, C.byref(h), C.byref(w)))
    return out


def pipeline_run_multi(dets, recs, image_ptrs, hs: np.ndarray, ws: np.ndarray, cfg: PipelineConfig,
                       bufs: PipelineBuffers):
    """oar_pipeline_run_multi: dets[g] / recs[g] live on context g (one host thread + CUDA context per GPU inside the
    library); host pages only; equal to one un-sharded pipeline_run"""
    n_ctx = len(dets)
    d = (C.c_void_p * n_ctx)(*[m.handle for m in dets])
    r = (C.c_void_p * n_ctx)(*[m.handle for m in recs])
    check(lib().oar_pipeline_run_multi(d, r, n_ctx, image_ptrs, _ptr(hs), _ptr(ws), len(hs), C.byref(cfg),
                                       C.byref(bufs.res)))
    return bufs.res


def pipeline_run(det: Model, rec: Model, image_ptrs, hs: np.ndarray, ws: np.ndarray, on_device: bool,
                 cfg: PipelineConfig, bufs: PipelineBuffers, cls: Model | None = None):
    """image_ptrs: ctypes array of c_void_p (host or device addresses); cls: optional line-orientation classifier"""
    if cls is None:
        check(lib().oar_pipeline_run(det.handle, rec.handle, image_ptrs, _ptr(hs), _ptr(ws), len(hs),
                                     1 if on_device else 0, C.byref(cfg), C.byref(bufs.res)))
    else:
        check(lib().oar_pipeline_run_cls(det.handle, rec.handle, cls.handle, image_ptrs, _ptr(hs), _ptr(ws), len(hs),
                                         1 if on_device else 0, C.byref(cfg), C.byref(bufs.res)))
    return bufs.res
