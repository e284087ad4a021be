"""ctypes binding of the CPU oracle (oracle/oar_oracle.cpp).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  The product package
(oar_ocr_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboar_oracle.so")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "oar_oracle.cpp")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = C.CDLL(_SO)
        _lib.oracle_box_score_fast.restype = C.c_float
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


u8p = lambda a: _p(a, C.c_uint8)
f32p = lambda a: _p(a, C.c_float)
i32p = lambda a: _p(a, C.c_int32)


def det_resize_dims(h, w, limit=960, limit_type=0, max_side=4000):
    oh, ow = C.c_uint32(), C.c_uint32()
    lib().oracle_det_resize_dims(C.c_uint32(h), C.c_uint32(w), C.c_uint32(limit), limit_type, C.c_uint32(max_side),
                                 C.byref(oh), C.byref(ow))
    return oh.value, ow.value


def resize_triangle(img: np.ndarray, nw: int, nh: int) -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    out = np.empty((nh, nw, 3), np.uint8)
    lib().oracle_resize_triangle(u8p(img), C.c_uint32(w), C.c_uint32(h), C.c_uint32(nw), C.c_uint32(nh), u8p(out))
    return out


def norm_coeffs(scale, mean, std):
    mean = np.asarray(mean, np.float32)
    std = np.asarray(std, np.float32)
    a = np.empty(3, np.float32)
    b = np.empty(3, np.float32)
    lib().oracle_norm_coeffs(C.c_float(scale), f32p(mean), f32p(std), f32p(a), f32p(b))
    return a, b


def normalize(img: np.ndarray, alpha, beta, src=(2, 1, 0), layout="chw") -> np.ndarray:
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    src_a = np.asarray(src, np.int32)
    alpha = np.asarray(alpha, np.float32)
    beta = np.asarray(beta, np.float32)
    if layout == "chw":
        out = np.empty((3, h, w), np.float32)
        lib().oracle_normalize_chw(u8p(img), w, h, i32p(src_a), f32p(alpha), f32p(beta), f32p(out))
    else:
        out = np.empty((h, w, 3), np.float32)
        lib().oracle_normalize_hwc(u8p(img), w, h, i32p(src_a), f32p(alpha), f32p(beta), f32p(out))
    return out


# DB detector normalisation exactly as db.rs:409-415 configures it
DET_SCALE = np.float32(1.0) / np.float32(255.0)
DET_MEAN = (0.485, 0.456, 0.406)
DET_STD = (0.229, 0.224, 0.225)


def det_normalize(img: np.ndarray) -> np.ndarray:
    a, b = norm_coeffs(DET_SCALE, DET_MEAN, DET_STD)
    return normalize(img, a, b, (2, 1, 0), "chw")


def threshold_mask(pred: np.ndarray, thresh: float) -> np.ndarray:
    pred = np.ascontiguousarray(pred, np.float32)
    out = np.empty(pred.shape, np.uint8)
    lib().oracle_threshold_mask(f32p(pred), pred.size, C.c_float(thresh), u8p(out))
    return out


def find_contours(mask: np.ndarray):
    mask = np.ascontiguousarray(mask, np.uint8)
    h, w = mask.shape
    tot = C.c_int64()
    n = lib().oracle_find_contours(u8p(mask), w, h, None, C.c_int64(0), None, None, 0, C.byref(tot))
    xy = np.empty((max(tot.value, 1), 2), np.int32)
    off = np.empty(n + 1, np.int32)
    bt = np.empty(max(n, 1), np.int32)
    n2 = lib().oracle_find_contours(u8p(mask), w, h, i32p(xy), C.c_int64(xy.size), i32p(off), i32p(bt), n,
                                    C.byref(tot))
    assert n2 == n
    return [xy[off[i]:off[i + 1]].copy() for i in range(n)], bt[:n].copy()


def simplify_chain(pts):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.empty_like(pts)
    n = lib().oracle_simplify_chain(f32p(pts), len(pts), f32p(out))
    return out[:n]


def order_mini_box(pts):
    pts = np.ascontiguousarray(pts, np.float32).copy()
    lib().oracle_order_mini_box(f32p(pts))
    return pts


def mini_boxes_from_points(pts):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.empty((4, 2), np.float32)
    ms = C.c_float()
    ok = lib().oracle_mini_boxes_from_points(f32p(pts), len(pts), f32p(out), C.byref(ms))
    return (out, ms.value) if ok else None


def min_area_rect(pts):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.empty(5, np.float32)
    lib().oracle_min_area_rect(f32p(pts), len(pts), f32p(out))
    return out


def box_score_fast(pred, pts):
    pred = np.ascontiguousarray(pred, np.float32)
    pts = np.ascontiguousarray(pts, np.float32)
    h, w = pred.shape
    return float(lib().oracle_box_score_fast(f32p(pred), w, h, f32p(pts), len(pts)))


def unclip(pts, ratio):
    pts = np.ascontiguousarray(pts, np.float32)
    out = np.empty((4096, 2), np.float32)
    n = lib().oracle_unclip(f32p(pts), len(pts), C.c_float(ratio), f32p(out), 4096)
    return out[:n].copy()


def db_postprocess(pred, dest_w, dest_h, thresh=0.3, box_thresh=0.6, unclip_ratio=2.0, max_candidates=1000,
                   min_size=3.0, with_raw=False):
    pred = np.ascontiguousarray(pred, np.float32)
    h, w = pred.shape
    cap = max_candidates
    boxes = np.empty((cap, 4, 2), np.float32)
    scores = np.empty(cap, np.float32)
    raw = np.empty((cap, 4, 2), np.float32)
    n = lib().oracle_db_postprocess(f32p(pred), w, h, C.c_uint32(dest_w), C.c_uint32(dest_h), C.c_float(thresh),
                                    C.c_float(box_thresh), C.c_float(unclip_ratio), max_candidates,
                                    C.c_float(min_size), f32p(boxes), f32p(scores), f32p(raw), cap)
    if with_raw:
        return boxes[:n].copy(), scores[:n].copy(), raw[:n].copy()
    return boxes[:n].copy(), scores[:n].copy()


def sort_quad_boxes(boxes):
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4, 2).copy()
    order = np.empty(max(len(boxes), 1), np.int32)
    lib().oracle_sort_quad_boxes(f32p(boxes), len(boxes), i32p(order))
    return boxes, order[:len(boxes)].copy()


def is_exact_axis_aligned(pts, w, h):
    pts = np.ascontiguousarray(pts, np.float32)
    return bool(lib().oracle_is_exact_axis_aligned(f32p(pts), C.c_uint32(w), C.c_uint32(h)))


def perspective_transform(src, dst):
    src = np.ascontiguousarray(src, np.float32)
    dst = np.ascontiguousarray(dst, np.float32)
    m = np.empty(9, np.float32)
    ok = lib().oracle_perspective_transform(f32p(src), f32p(dst), f32p(m))
    return m.reshape(3, 3) if ok else None


def bicubic(img, x, y):
    img = np.ascontiguousarray(img, np.uint8)
    h, w, _ = img.shape
    out = np.empty(3, np.uint8)
    lib().oracle_bicubic(u8p(img), w, h, C.c_float(x), C.c_float(y), u8p(out))
    return out


def rotate_crop(img, quad):
    """get_rotate_crop_image; returns None where the reference returns Err."""
    img = np.ascontiguousarray(img, np.uint8)
    quad = np.ascontiguousarray(quad, np.float32)
    h, w, _ = img.shape
    ow, oh = C.c_int(), C.c_int()
    rc = lib().oracle_rotate_crop(u8p(img), w, h, f32p(quad), None, C.byref(ow), C.byref(oh))
    if rc != 0:
        return None
    out = np.empty((oh.value, ow.value, 3), np.uint8)
    lib().oracle_rotate_crop(u8p(img), w, h, f32p(quad), u8p(out), C.byref(ow), C.byref(oh))
    return out


REC_H, REC_W, REC_MAX_W = 48, 320, 3200


def crnn_tensor_width(crops):
    ws = np.array([c.shape[1] for c in crops], np.int32)
    hs = np.array([c.shape[0] for c in crops], np.int32)
    return lib().oracle_crnn_tensor_width(i32p(ws), i32p(hs), len(crops), REC_H, REC_W, REC_MAX_W)


def crnn_preprocess(crops):
    """crnn.rs:71-125 -> f32 [B,3,48,tensor_w]"""
    if not crops:
        return np.zeros((0, 0, 0, 0), np.float32)
    tw = crnn_tensor_width(crops)
    out = np.zeros((len(crops), 3, REC_H, tw), np.float32)
    for i, c in enumerate(crops):
        c = np.ascontiguousarray(c, np.uint8)
        lib().oracle_crnn_preprocess_one(u8p(c), c.shape[1], c.shape[0], REC_H, tw, f32p(out[i]))
    return out


def ctc_argmax(pred):
    pred = np.ascontiguousarray(pred, np.float32)
    b, t, v = pred.shape
    if pred.size == 0:
        return np.zeros((0, 0), np.int32), np.zeros((0, 0), np.float32)
    idx = np.empty((b, t), np.int32)
    prob = np.empty((b, t), np.float32)
    lib().oracle_ctc_argmax(f32p(pred), C.c_int64(b * t), v, i32p(idx), f32p(prob))
    return idx, prob


def ctc_decode(idx, prob, n_chars):
    """returns (list of kept-index arrays, scores, list of column arrays, T)"""
    idx = np.ascontiguousarray(idx, np.int32)
    prob = np.ascontiguousarray(prob, np.float32)
    b, t = idx.shape
    if b == 0:
        return [], np.zeros(0, np.float32), [], t
    oi = np.empty((b, max(t, 1)), np.int32)
    oc = np.empty((b, max(t, 1)), np.int32)
    ol = np.empty(b, np.int32)
    osc = np.empty(b, np.float32)
    lib().oracle_ctc_decode(i32p(idx), f32p(prob), b, t, n_chars, i32p(oi), i32p(oc), i32p(ol), f32p(osc))
    return [oi[i, :ol[i]].copy() for i in range(b)], osc, [oc[i, :ol[i]].copy() for i in range(b)], t


# ---- text-line orientation stage ----
CLS_INPUT_SHAPE = (80, 160)  # TextLineOrientationAdapter::DEFAULT_INPUT_SHAPE (h, w), text_line_orientation_adapter.rs:48


def rotate180(img: np.ndarray) -> np.ndarray:
    """image::imageops::rotate180 as classify_line_orientations applies it (ocr.rs:785-788)"""
    img = np.ascontiguousarray(img, np.uint8)
    out = np.empty_like(img)
    lib().oracle_rotate180(u8p(img), img.shape[1], img.shape[0], u8p(out))
    return out


def cls_preprocess(crops, input_shape=CLS_INPUT_SHAPE) -> np.ndarray:
    """PPLCNetModel::preprocess_refs with resize_short = None (pp_lcnet.rs:139-196): direct Triangle resize to
    (w, h) = (input_shape[1], input_shape[0]), then NormalizeImage(scale 1/255, ImageNet mean/std, CHW, RGB order)
    (pp_lcnet.rs:400-412).  Zero-sized images are dropped, as the reference's filter_map does."""
    ih, iw = input_shape
    a, b = norm_coeffs(DET_SCALE, DET_MEAN, DET_STD)
    out = []
    for c in crops:
        if c.shape[0] == 0 or c.shape[1] == 0:
            continue
        out.append(normalize(resize_triangle(c, iw, ih), a, b, (0, 1, 2), "chw"))
    return np.stack(out) if out else np.zeros((0, 3, ih, iw), np.float32)


def topk(pred, k):
    """Topk::process for one prediction row (utils/topk.rs): (indexes, scores) of the k best, ties in index order"""
    pred = np.ascontiguousarray(pred, np.float32).ravel()
    if k <= 0:
        raise ValueError("k must be greater than 0")
    idx = np.empty(max(len(pred), 1), np.int32)
    sc = np.empty(max(len(pred), 1), np.float32)
    m = lib().oracle_topk(f32p(pred), len(pred), k, i32p(idx), f32p(sc))
    return idx[:m].copy(), sc[:m].copy()


# ---- layout detection post-process (host half) ----
def layout_nms(boxes, classes, scores, compacting=False):
    """paddlex_layout_nms (layout_detection_adapter.rs:884-933), or with compacting=True the reference's own test
    oracle compacting_nms_reference (:1667-1697).  boxes [n,4] = from_coords arguments.  Returns kept indices."""
    boxes = np.ascontiguousarray(boxes, np.float32).reshape(-1, 4)
    classes = np.ascontiguousarray(classes, np.int32)
    scores = np.ascontiguousarray(scores, np.float32)
    keep = np.empty(max(len(boxes), 1), np.int32)
    fn = lib().oracle_layout_nms_compacting if compacting else lib().oracle_layout_nms
    m = fn(f32p(boxes), i32p(classes), f32p(scores), len(boxes), i32p(keep))
    return keep[:m].copy()


def layout_postprocess(pred, src_w, src_h, num_classes, score_threshold=0.5, max_elements=100, layout_nms=True,
                       class_thresholds=None, class_merge_modes=None, image_class_id=-1, formula_class_id=-1,
                       unclip=None):
    """postprocess_pp_doclayout for one image (layout_detection_adapter.rs:674-840).  pred [N,F]; dict-valued
    options are keyed by class id; unclip: None | (w, h) | {class_id: (w, h)}.  Returns (boxes [n,4], classes, scores)."""
    pred = np.ascontiguousarray(pred, np.float32)
    n, f = pred.shape
    thr = mm = cu = None
    if class_thresholds is not None:
        thr = np.full(num_classes, np.nan, np.float32)
        for k, v in class_thresholds.items():
            if 0 <= k < num_classes:
                thr[k] = v
    if class_merge_modes is not None:
        mm = np.full(num_classes, -1, np.int32)
        for k, v in class_merge_modes.items():
            if 0 <= k < num_classes:
                mm[k] = v
    mode, uw, uh = 0, 1.0, 1.0
    if isinstance(unclip, dict):
        mode = 2
        cu = np.full((num_classes, 2), np.nan, np.float32)
        for k, v in unclip.items():
            if 0 <= k < num_classes:
                cu[k] = v
    elif unclip is not None:
        mode, uw, uh = 1, float(unclip[0]), float(unclip[1])
    me = max(int(max_elements), 1)
    ob = np.zeros((me, 4), np.float32)
    oc = np.zeros(me, np.int32)
    os_ = np.zeros(me, np.float32)
    m = lib().oracle_layout_postprocess(
        f32p(pred), n, f, C.c_float(src_w), C.c_float(src_h), C.c_float(score_threshold), me, 1 if layout_nms else 0,
        num_classes, f32p(thr) if thr is not None else None, i32p(mm) if mm is not None else None, image_class_id,
        formula_class_id, mode, C.c_float(uw), C.c_float(uh), f32p(cu) if cu is not None else None, f32p(ob), i32p(oc),
        f32p(os_))
    return ob[:m].copy(), oc[:m].copy(), os_[:m].copy()


FILTERS = {"triangle": 0, "catmullrom": 1, "lanczos3": 2}


def resize_filter(img: np.ndarray, nw: int, nh: int, filter: str = "catmullrom") -> np.ndarray:
    """image::imageops::resize / DynamicImage::resize_exact with FilterType::{Triangle, CatmullRom, Lanczos3}"""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    h, w, _ = img.shape
    out = np.empty((nh, nw, 3), np.uint8)
    lib().oracle_resize_filter(u8p(img), C.c_uint32(w), C.c_uint32(h), C.c_uint32(nw), C.c_uint32(nh), u8p(out),
                               FILTERS[filter])
    return out


def layout_preprocess(images, image_shape=(640, 640)):
    """ScaleAwareDetectorModel::preprocess with ScaleAwareDetectorPreprocessConfig::pp_doclayout
    (scale_aware_detector.rs:66-80, 195-229; resize_image_type1 with keep_ratio = false, resize_detection.rs:337-366):
    resize_exact to image_shape (h, w) with CatmullRom, scale 1/255, mean 0 / std 1, RGB, CHW.
    Returns (tensor [B,3,h,w], scale factors [B,2] = resized / original (h, w))."""
    th, tw = image_shape
    a, b = norm_coeffs(DET_SCALE, (0.0, 0.0, 0.0), (1.0, 1.0, 1.0))
    out, scale = [], []
    for img in images:
        h, w, _ = img.shape
        r = img if (h, w) == (th, tw) else resize_filter(img, tw, th, "catmullrom")
        out.append(normalize(r, a, b, (0, 1, 2), "chw"))
        scale.append((np.float32(th) / np.float32(h), np.float32(tw) / np.float32(w)))
    return np.stack(out), np.array(scale, np.float32)
