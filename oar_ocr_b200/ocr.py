"""Host-side mirror of the reference's public API for the det+rec path.

Same names, argument meaning and error behaviour as the Rust crate so parity
tests read like the reference's own:

  OAROCRBuilder / OAROCR.predict          src/oarocr/ocr.rs:66-417, 518-659
  TextDetectionPredictor(+Builder)        oar-ocr-core/src/predictors/text_detection.rs:23-112
  TextRecognitionPredictor(+Builder)      oar-ocr-core/src/predictors/text_recognition.rs:19-110
  TextLineOrientationPredictor(+Builder)  oar-ocr-core/src/predictors/text_line_orientation.rs,
                                          domain/adapters/text_line_orientation_adapter.rs:17-121
  OAROCRResult / TextRegion               src/oarocr/result.rs:33-49, oar-ocr-core/src/domain/text_region.rs:10-27
  Detection / BoundingBox                 oar-ocr-core/src/domain/tasks/text_detection.rs:15-21, processors/geometry.rs:15-70

Everything numeric happens behind the C ABI (liboar_b200.so); this module only
converts between Python objects and flat buffers.  There is no CPU fallback.
"""
from __future__ import annotations

import os
from dataclasses import dataclass, field

import numpy as np

from . import ffi
from .ffi import OCRError

MAX_BATCH_SIZE = 4096  # OAROCRBuilder::MAX_BATCH_SIZE, ocr.rs:93


@dataclass
class BoundingBox:
    points: np.ndarray  # [N,2] f32

    def x_min(self):
        return float(self.points[:, 0].min())

    def x_max(self):
        return float(self.points[:, 0].max())

    def y_min(self):
        return float(self.points[:, 1].min())

    def y_max(self):
        return float(self.points[:, 1].max())

    @staticmethod
    def from_coords(x1, y1, x2, y2) -> "BoundingBox":
        """BoundingBox::from_coords (geometry.rs): the four corners clockwise from the top-left"""
        return BoundingBox(np.array([[x1, y1], [x2, y1], [x2, y2], [x1, y2]], np.float32))


@dataclass
class Detection:
    bbox: BoundingBox
    score: float


@dataclass
class TextDetectionResult:
    detections: list  # list[list[Detection]]


@dataclass
class TextRecognitionResult:
    texts: list
    scores: list
    char_col_indices: list = field(default_factory=list)
    sequence_lengths: list = field(default_factory=list)
    label_indices: list = field(default_factory=list)  # CTC-collapsed class indices (not in the reference struct)


@dataclass
class TextRegion:
    bounding_box: BoundingBox
    dt_poly: BoundingBox
    rec_poly: BoundingBox
    text: str
    confidence: float
    orientation_angle: float | None = None
    word_boxes: list | None = None
    label: str | None = None
    detection_index: int = 0
    label_indices: np.ndarray | None = None


@dataclass
class OAROCRResult:
    input_path: str
    index: int
    input_img: np.ndarray
    text_regions: list
    orientation_angle: float | None = None
    rectified_img: np.ndarray | None = None


@dataclass
class TextDetectionConfig:
    """tasks/text_detection.rs:33-53"""
    score_threshold: float = 0.3
    box_threshold: float = 0.6
    unclip_ratio: float = 1.5
    max_candidates: int = 1000
    limit_side_len: int | None = None
    limit_type: str | None = None  # "max" | "min" | "resize_long"
    max_side_len: int | None = None

    def to_ffi(self) -> ffi.DetConfig:
        lt = {"max": 0, "min": 1, "resize_long": 2}
        # adapter preprocessing defaults: 960 / Max / 4000 (adapters/preprocessing.rs:73-90)
        return ffi.det_config(thresh=self.score_threshold, box_thresh=self.box_threshold,
                              unclip_ratio=self.unclip_ratio, max_candidates=self.max_candidates,
                              limit_side_len=self.limit_side_len or 960,
                              limit_type=lt[(self.limit_type or "max").lower()],
                              max_side_limit=self.max_side_len or 4000)

    def validate(self):
        for name in ("score_threshold", "box_threshold"):
            v = getattr(self, name)
            if not (0.0 <= v <= 1.0):
                raise OCRError("ConfigError", f"{name} must be in [0,1], got {v}")
        if self.unclip_ratio <= 0 or self.max_candidates <= 0:
            raise OCRError("ConfigError", "unclip_ratio and max_candidates must be positive")


@dataclass
class TextRecognitionConfig:
    score_threshold: float = 0.0


def dict_lines(content: str) -> list[str]:
    """Rust `str::lines()` as the reference applies it to the dictionary file (ocr.rs:386, utils/dict.rs:43,
    predictors/text_recognition.rs:89): split on '\\n' only, a '\\r' directly before it belongs to the terminator, no
    final empty piece (split_inclusive('\\n') + strip_suffix, library/core/src/str/mod.rs).  (Python's str.splitlines() also breaks on \\x0b \\x0c \\x1c-\\x1e \\x85 U+2028 U+2029, which would shift
    every later class index.)"""
    parts = content.split("\n")
    last = parts.pop()  # the piece after the final '\n': a line only if non-empty, and its bare '\r' is preserved
    lines = [p[:-1] if p.endswith("\r") else p for p in parts]  # '\r' is stripped only as part of a '\r\n' terminator
    if last:
        lines.append(last)
    return lines


def character_list(lines: list[str]) -> list[str]:
    """CTCLabelDecode::from_string_list(dict, use_space_char=true, has_explicit_blank=false)
    (decode.rs:392-423, 118-141): ['\\0' blank] + the FIRST char of every non-empty dictionary line
    (`filter_map(|s| s.chars().next())`: empty lines are dropped, longer lines truncated) + [' ']"""
    return ["\0"] + [ln[0] for ln in lines if ln] + [" "]


_KIND_ID = {"det": ffi.KIND_DET, "rec": ffi.KIND_REC, "cls": ffi.KIND_CLS}


def _resolve_model(source, kind: str) -> bytes:
    """ModelSource (core/config/model_source.rs:20-28: ModelSource::Path / Memory): OARG or ONNX bytes, a path to an
    .oarg or .onnx file, or 'synthetic[:seed]'.  Returns the bytes to hand to the C ABI: ONNX models go through
    unchanged and are converted to the layer list inside the library (oar_model_load_onnx), exactly where the reference
    hands them to ONNX Runtime."""
    if isinstance(source, (bytes, bytearray)):
        return bytes(source)
    if isinstance(source, str) and source.startswith("synthetic"):
        from . import models
        seed = int(source.split(":")[1]) if ":" in source else 42
        return models.get_blob(kind, seed)
    if isinstance(source, (str, os.PathLike)):
        path = os.fspath(source)
        if not os.path.exists(path):
            raise OCRError("ModelLoad", f"model file '{path}' does not exist")
        with open(path, "rb") as f:
            return f.read()
    raise OCRError("InvalidInput", f"unsupported model source {type(source)}")


def _load_model(ctx, source, kind: str) -> "ffi.Model":
    return ffi.Model(ctx, _resolve_model(source, kind), _KIND_ID[kind])


def _decode_texts(chars, label_lists):
    return ["".join(chars[k] for k in lab if 0 <= k < len(chars)) for lab in label_lists]


_contexts: dict = {}


def is_cjk(ch: str) -> bool:
    """OAROCR::is_cjk (src/oarocr/ocr.rs:1065-1084)"""
    u = ord(ch)
    return (0x4E00 <= u <= 0x9FFF or 0x3400 <= u <= 0x4DBF or 0x20000 <= u <= 0x2A6DF or 0x2A700 <= u <= 0x2B73F
            or 0x2B740 <= u <= 0x2B81F)


def ctc_word_boxes(line_bbox: BoundingBox, text: str, col_indices, seq_len: int, wh_ratio: float,
                   max_wh_ratio: float) -> list:
    """OAROCR::ctc_word_boxes (src/oarocr/ocr.rs:949-1022) in f32: one axis-aligned box per character, from the CTC
    timestep it was emitted at.  CJK characters get a box of the line's average character width around their cell
    centre; others span to the midpoints between neighbouring centres."""
    f = np.float32
    if len(col_indices) == 0 or seq_len == 0 or not text:
        return []
    effective = f(seq_len) * (f(wh_ratio) / f(max_wh_ratio))
    eps = f(np.finfo(np.float32).eps)
    if effective <= eps:
        return []
    x_min, y_min = f(line_bbox.x_min()), f(line_bbox.y_min())
    x_max, y_max = f(line_bbox.x_max()), f(line_bbox.y_max())
    width = f(x_max - x_min)
    cell = f(width / max(effective, eps))
    chars = list(text)
    avg = f(width / f(max(len(chars), 1)))
    centers = [f(x_min + f(f(f(int(i)) + f(0.5)) * cell)) for i in col_indices]
    out = []
    for i in range(len(col_indices)):
        ch = chars[i] if i < len(chars) else "?"
        c = centers[i]
        if is_cjk(ch):
            half = f(avg / f(2.0))
            lo, hi = max(f(c - half), x_min), min(f(c + half), x_max)
        else:
            lo = x_min if i == 0 else f(f(centers[i - 1] + c) / f(2.0))
            lo = max(lo, x_min)
            hi = x_max if i == len(col_indices) - 1 else f(f(c + centers[i + 1]) / f(2.0))
            hi = min(hi, x_max)
        out.append(BoundingBox.from_coords(lo, y_min, hi, y_max))
    return out


def default_context(device_id: int = 0) -> ffi.Context:
    if device_id not in _contexts:
        _contexts[device_id] = ffi.Context(device_id)
    return _contexts[device_id]


def _validate_images(images, what):
    if images is None or len(images) == 0:
        # OCRError::validation_error(..., "non-empty slice", "empty slice"), ocr.rs:525-532 / validation.rs
        raise OCRError("InvalidInput", f"{what}: images: expected non-empty slice, got empty slice", ffi.OAR_E_INVALID)


class TextDetectionPredictorBuilder:
    def __init__(self):
        self._config = TextDetectionConfig()
        self._device = 0

    def score_threshold(self, v):
        self._config.score_threshold = v
        return self

    def box_threshold(self, v):
        self._config.box_threshold = v
        return self

    def unclip_ratio(self, v):
        self._config.unclip_ratio = v
        return self

    def max_candidates(self, v):
        self._config.max_candidates = v
        return self

    def with_config(self, cfg: TextDetectionConfig):
        self._config = cfg
        return self

    def device_id(self, d):
        self._device = d
        return self

    def build(self, model_source) -> "TextDetectionPredictor":
        self._config.validate()
        ctx = default_context(self._device)
        return TextDetectionPredictor(_load_model(ctx, model_source, "det"), self._config)


class TextDetectionPredictor:
    """predict(): validate_input -> TextDetectionAdapter::execute -> validate_output (predictors/core.rs:58-69).
    Boxes come back in discovery order, unsorted, as the adapter returns them."""

    def __init__(self, model: ffi.Model, config: TextDetectionConfig):
        self.model = model
        self.config = config

    @staticmethod
    def builder():
        return TextDetectionPredictorBuilder()

    def predict(self, images) -> TextDetectionResult:
        _validate_images(images, "TextDetection")
        out = self.model.det_run(images, self.config.to_ffi())
        dets = []
        for boxes, scores in out:
            for s in scores:  # validate_output: scores in [0,1] (tasks/text_detection.rs:129-141)
                if not (0.0 <= float(s) <= 1.0):
                    raise OCRError("InvalidInput", f"detection score {s} outside [0,1]")
            dets.append([Detection(BoundingBox(b.copy()), float(s)) for b, s in zip(boxes, scores)])
        return TextDetectionResult(dets)


class TextRecognitionPredictorBuilder:
    def __init__(self):
        self._config = TextRecognitionConfig()
        self._dict = None
        self._device = 0

    def score_threshold(self, v):
        self._config.score_threshold = v
        return self

    def dict_path(self, path):
        self._dict = path
        return self

    def character_dict(self, lines):
        self._dict = list(lines)
        return self

    def device_id(self, d):
        self._device = d
        return self

    def build(self, model_source) -> "TextRecognitionPredictor":
        if self._dict is None:
            raise OCRError("ConfigError", "missing field dict_path for TextRecognitionPredictor")
        if isinstance(self._dict, list):
            lines = self._dict
        else:
            try:
                with open(self._dict, "r", encoding="utf-8", newline="") as f:
                    lines = dict_lines(f.read())
            except OSError as e:
                raise OCRError("InvalidInput", f"Failed to read character dictionary from '{self._dict}': {e}")
        ctx = default_context(self._device)
        return TextRecognitionPredictor(_load_model(ctx, model_source, "rec"), character_list(lines),
                                        self._config)


class TextRecognitionPredictor:
    """predict(): the whole input is ONE batch (predictors/text_recognition.rs:38-45 -> crnn.rs:247-293)"""

    def __init__(self, model: ffi.Model, chars: list[str], config: TextRecognitionConfig):
        self.model = model
        self.chars = chars
        self.config = config

    @staticmethod
    def builder():
        return TextRecognitionPredictorBuilder()

    def predict(self, images) -> TextRecognitionResult:
        _validate_images(images, "TextRecognition")
        r = self.model.rec_run(images, len(self.chars))
        texts = _decode_texts(self.chars, r["labels"])
        scores = [float(s) for s in r["scores"]]
        labels = list(r["labels"])
        for i, s in enumerate(scores):  # text_recognition_adapter.rs:88-102: below threshold -> empty text, slot kept
            if s < self.config.score_threshold:
                texts[i] = ""
                labels[i] = labels[i][:0]
        return TextRecognitionResult(texts, scores, [c.tolist() for c in r["cols"]], [r["T"]] * len(texts), labels)


@dataclass
class Classification:
    """domain/tasks: Classification{class_id, label, score}"""
    class_id: int
    label: str
    score: float


@dataclass
class TextLineOrientationConfig:
    """tasks/text_line_orientation.rs:16-32"""
    score_threshold: float = 0.5
    topk: int = 2

    def validate(self):
        if not (0.0 <= self.score_threshold <= 1.0):
            raise OCRError("ConfigError", f"score_threshold must be in [0,1], got {self.score_threshold}")
        if self.topk < 1:
            raise OCRError("ConfigError", f"topk must be at least 1, got {self.topk}")


@dataclass
class TextLineOrientationResult:
    orientations: list  # list[list[Classification]], one list (top-k, best first) per image


def topk_indices(probs: np.ndarray, k: int):
    """Topk::extract_topk_from_prediction (utils/topk.rs): stable sort by score descending, first k"""
    if k <= 0:
        raise OCRError("InvalidInput", "k must be greater than 0", ffi.OAR_E_INVALID)
    order = sorted(range(len(probs)), key=lambda i: -float(probs[i]))  # sorted() is stable
    return order[:min(k, len(probs))]


class TextLineOrientationPredictorBuilder:
    def __init__(self):
        self._config = TextLineOrientationConfig()
        # the stand-alone predictor's default (predictors/text_line_orientation.rs:59) is (192, 48), read as
        # (height, width) by PPLCNetModel; the pipeline's adapter default is (80, 160)
        self._input_shape = (192, 48)
        self._device = 0

    def score_threshold(self, v):
        self._config.score_threshold = v
        return self

    def topk(self, k):
        self._config.topk = k
        return self

    def input_shape(self, shape):
        self._input_shape = (int(shape[0]), int(shape[1]))
        return self

    def with_config(self, cfg: TextLineOrientationConfig):
        self._config = cfg
        return self

    def device_id(self, d):
        self._device = d
        return self

    def build(self, model_source) -> "TextLineOrientationPredictor":
        self._config.validate()
        ctx = default_context(self._device)
        return TextLineOrientationPredictor(_load_model(ctx, model_source, "cls"), self._config,
                                            self._input_shape)


class TextLineOrientationPredictor:
    """predict(): TextLineOrientationAdapter::execute (text_line_orientation_adapter.rs:63-121): one batch, top-k
    classes per crop with labels "0" / "180"."""

    LABELS = ["0", "180"]  # TextLineOrientationAdapter::labels

    def __init__(self, model: ffi.Model, config: TextLineOrientationConfig, input_shape=(192, 48)):
        self.model, self.config, self.input_shape = model, config, input_shape

    @staticmethod
    def builder():
        return TextLineOrientationPredictorBuilder()

    def predict(self, images) -> TextLineOrientationResult:
        if images is None or len(images) == 0:
            raise OCRError("InvalidInput", "No images provided for text line orientation classification",
                           ffi.OAR_E_INVALID)
        r = self.model.cls_run(images, self.input_shape)
        out = []
        for p in r["probs"]:
            row = []
            for i in topk_indices(p, self.config.topk):
                label = self.LABELS[i] if i < len(self.LABELS) else f"class_{i}"
                s = float(p[i])
                if not (0.0 <= s <= 1.0):  # validate_output, tasks/text_line_orientation.rs:97-110
                    raise OCRError("InvalidInput", f"score {s} outside [0,1]")
                row.append(Classification(int(i), label, s))
            out.append(row)
        return TextLineOrientationResult(out)


# ---- layout detection, host half: LayoutDetectionAdapter::postprocess_pp_doclayout ----
# class labels of LayoutModelConfig::pp_doclayout_l (layout_detection_adapter.rs:314-348)
PP_DOCLAYOUT_L_LABELS = ["paragraph_title", "image", "text", "number", "abstract", "content", "figure_title",
                         "formula", "table", "table_title", "reference", "doc_title", "footnote", "header",
                         "algorithm", "footer", "seal", "chart_title", "chart", "formula_number", "header_image",
                         "footer_image", "aside_text"]


@dataclass
class LayoutDetectionConfig:
    """tasks/layout_detection.rs:45-100.  class_thresholds / class_merge_modes are keyed by label ("large" | "small" |
    "union"); layout_unclip_ratio: None | r | (w, h) | {class_id: (w, h)} (UnclipRatio::Uniform / Separate / PerClass)"""
    score_threshold: float = 0.5
    max_elements: int = 100
    class_thresholds: dict | None = None
    class_merge_modes: dict | None = None
    layout_nms: bool = True
    nms_threshold: float = 0.5
    layout_unclip_ratio: object = None

    def validate(self):
        if not (0.0 <= self.score_threshold <= 1.0):
            raise OCRError("ConfigError", f"score_threshold must be in [0,1], got {self.score_threshold}")
        if self.max_elements < 1:
            raise OCRError("ConfigError", f"max_elements must be at least 1, got {self.max_elements}")

    @staticmethod
    def with_pp_structurev3_thresholds() -> "LayoutDetectionConfig":
        """tasks/layout_detection.rs:102-135"""
        return LayoutDetectionConfig(class_thresholds={"paragraph_title": 0.3, "formula": 0.3, "text": 0.4,
                                                       "seal": 0.45})


@dataclass
class LayoutDetectionElement:
    bbox: BoundingBox
    element_type: str
    score: float


def postprocess_pp_doclayout(predictions: np.ndarray, img_shapes, config: LayoutDetectionConfig | None = None,
                             class_labels=None) -> list:
    """LayoutDetectionAdapter::postprocess_pp_doclayout (layout_detection_adapter.rs:631-846): predictions
    [B, N, 6|7|8] rows [class_id, score, x1, y1, x2, y2, (order keys)], img_shapes = (src_w, src_h) per image.
    The label-keyed maps are resolved to class ids here, as the adapter does; the arithmetic runs behind
    oar_layout_postprocess.  Returns one list of LayoutDetectionElement per image."""
    config = config or LayoutDetectionConfig()
    config.validate()
    labels = list(class_labels) if class_labels is not None else PP_DOCLAYOUT_L_LABELS
    ids = {lab: i for i, lab in enumerate(labels)}
    modes = {"large": ffi.MERGE_LARGE, "small": ffi.MERGE_SMALL, "union": ffi.MERGE_UNION}
    thr = None if config.class_thresholds is None else \
        {ids[k]: v for k, v in config.class_thresholds.items() if k in ids}
    mm = None if config.class_merge_modes is None else \
        {ids[k]: modes[str(v).lower()] for k, v in config.class_merge_modes.items() if k in ids}
    u = config.layout_unclip_ratio
    if isinstance(u, (int, float)):
        u = (float(u), float(u))
    pred = np.asarray(predictions, np.float32)
    if pred.ndim == 4:  # the model's [B, N, 1, F] layout
        pred = pred.reshape(pred.shape[0], pred.shape[1], pred.shape[3])
    out = ffi.layout_postprocess(pred, img_shapes, len(labels), config.score_threshold, config.max_elements,
                                 config.layout_nms, thr, mm, ids.get("image", -1), ids.get("formula", -1), u)
    return [[LayoutDetectionElement(BoundingBox.from_coords(*map(float, b)), labels[c] if c < len(labels) else "unknown",
                                    float(s)) for b, c, s in zip(bs, cs, ss)] for bs, cs, ss in out]


class OAROCRBuilder:
    """OAROCRBuilder::new(det_model, rec_model, dict_path) (ocr.rs:105-128)"""

    def __init__(self, text_detection_model, text_recognition_model, character_dict_path=None):
        self._det = text_detection_model
        self._rec = text_recognition_model
        self._dict_path = character_dict_path
        self._dict_content = None
        self._det_cfg = None
        self._rec_cfg = None
        self._image_bs = None
        self._region_bs = None
        self._device = 0
        self._return_word_box = False
        self._line_ori = None

    def with_text_line_orientation_classification(self, model_source):
        """OAROCRBuilder::with_text_line_orientation_classification (ocr.rs:197-203)"""
        self._line_ori = model_source
        return self

    def character_dict_content(self, content: str):
        self._dict_content = content
        return self

    def text_detection_config(self, cfg: TextDetectionConfig):
        self._det_cfg = cfg
        return self

    def text_recognition_config(self, cfg: TextRecognitionConfig):
        self._rec_cfg = cfg
        return self

    def return_word_box(self, enable: bool):
        """OAROCRBuilder::return_word_box (ocr.rs:241): per-character boxes from the CTC columns"""
        self._return_word_box = bool(enable)
        return self

    def image_batch_size(self, size: int):
        self._image_bs = size
        return self

    def region_batch_size(self, size: int):
        self._region_bs = size
        return self

    def device_id(self, d: int):
        self._device = d
        return self

    @staticmethod
    def validate_batch_size(name: str, size: int):
        if size == 0 or size > MAX_BATCH_SIZE or size < 0:
            raise OCRError("ConfigError", f"{name} must be in 1..={MAX_BATCH_SIZE}, got {size}")

    def build(self) -> "OAROCR":
        if self._image_bs is not None:
            self.validate_batch_size("image_batch_size", self._image_bs)
        if self._region_bs is not None:
            self.validate_batch_size("region_batch_size", self._region_bs)
        if self._dict_content is not None:
            content = self._dict_content
        else:
            if self._dict_path is None:
                raise OCRError("InvalidInput", "Failed to read character dictionary from '': no path given")
            try:
                with open(self._dict_path, "r", encoding="utf-8", newline="") as f:
                    content = f.read()
            except OSError as e:
                raise OCRError("InvalidInput", f"Failed to read character dictionary from '{self._dict_path}': {e}")
        chars = character_list(dict_lines(content))
        # no explicit config -> thresh .3 / box .6 / unclip 2.0 / limit 960 Max 4000 (ocr.rs:351-364)
        det_cfg = self._det_cfg or TextDetectionConfig(unclip_ratio=2.0, limit_side_len=960, limit_type="max",
                                                       max_side_len=4000)
        det_cfg.validate()
        rec_cfg = self._rec_cfg or TextRecognitionConfig()
        ctx = default_context(self._device)
        det = _load_model(ctx, self._det, "det")
        rec = _load_model(ctx, self._rec, "rec")
        # the B200 provider is an accelerator: adapter defaults 8 / 64 (builder_utils.rs:86-125)
        ocr = OAROCR(ctx, det, rec, chars, det_cfg, rec_cfg, self._image_bs or 8, self._region_bs or 64)
        ocr.return_word_box = self._return_word_box
        if self._line_ori is not None:
            ocr.cls = _load_model(ctx, self._line_ori, "cls")
            if ocr.cls.kind != ffi.KIND_CLS:
                raise OCRError("ModelLoad", "text line orientation model is not a classifier")
        return ocr


class OAROCR:
    def __init__(self, ctx, det, rec, chars, det_cfg, rec_cfg, image_bs, region_bs):
        self.ctx, self.det, self.rec, self.chars = ctx, det, rec, chars
        self.det_cfg, self.rec_cfg = det_cfg, rec_cfg
        self.image_batch_size, self.region_batch_size = image_bs, region_bs
        self.last_timing = {}
        self._bufs = None
        self.return_word_box = False  # ocr.rs:441: per-character boxes in TextRegion.word_boxes
        self.cls = None  # optional text-line orientation classifier (ocr.rs:35, 755-792)

    def _config(self) -> ffi.PipelineConfig:
        cfg = ffi.pipeline_config(image_batch_size=self.image_batch_size, region_batch_size=self.region_batch_size,
                                  rec_score_thresh=self.rec_cfg.score_threshold, n_chars=len(self.chars))
        cfg.det = self.det_cfg.to_ffi()
        return cfg

    def predict_raw(self, image_ptrs, hs, ws, on_device=False):
        """flat-buffer form used by the bench: returns the ffi.PipelineBuffers of this call"""
        n = len(hs)
        if self._bufs is None or len(self._bufs.region_off) != n + 1:
            self._bufs = ffi.PipelineBuffers(n, cap_regions=n * self.det_cfg.max_candidates)
        res = ffi.pipeline_run(self.det, self.rec, image_ptrs, hs, ws, on_device, self._config(), self._bufs, self.cls)
        self.last_timing = dict(ms_h2d=res.ms_h2d, ms_det=res.ms_det, ms_post=res.ms_post, ms_crop=res.ms_crop,
                                ms_rec=res.ms_rec, ms_total=res.ms_total, ms_cls=res.ms_cls, h2d_bytes=res.h2d_bytes,
                                d2h_bytes=res.d2h_bytes)
        return self._bufs

    def predict(self, images) -> list[OAROCRResult]:
        _validate_images(images, "OCR Pipeline")
        arrs, ptrs, hs, ws = ffi._image_table(images)
        b = self.predict_raw(ptrs, hs, ws, False)
        results = []
        for i, img in enumerate(arrs):
            regions = []
            for r in range(b.region_off[i], b.region_off[i + 1]):
                lab = b.labels[b.label_off[r]:b.label_off[r + 1]].copy()
                bbox = BoundingBox(b.boxes[r].copy())
                text = _decode_texts(self.chars, [lab])[0]
                word_boxes = None
                if self.return_word_box:  # ocr.rs:860-868
                    cols = b.cols[b.label_off[r]:b.label_off[r + 1]]
                    if len(cols) and b.seq_len[r] > 0:
                        word_boxes = ctc_word_boxes(bbox, text, cols, int(b.seq_len[r]), float(b.wh_ratio[r]),
                                                    float(b.max_wh_ratio[r]))
                angle = float(b.line_angle[r]) if self.cls is not None and b.line_angle[r] >= 0 else None
                regions.append(TextRegion(bbox, bbox, bbox, text, float(b.scores[r]), orientation_angle=angle,
                                          word_boxes=word_boxes, detection_index=int(b.det_index[r]),
                                          label_indices=lab))
            results.append(OAROCRResult(f"image_{i}", i, img, regions))
        return results
