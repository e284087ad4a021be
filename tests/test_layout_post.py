"""Layout-detection post-process, the host half of SURVEY.md 8f item 1 (runs without a GPU: host code in the reference
and in the library).

oracle  = oracle/oar_oracle.cpp: step-by-step restatement of LayoutDetectionAdapter::postprocess_pp_doclayout
          (oar-ocr-core/src/domain/adapters/layout_detection_adapter.rs:631-1116) and unclip_boxes
          (processors/layout_postprocess.rs:636-681), pinned by the reference's own NMS test vector (:1699-1727);
product = oar_layout_postprocess in liboar_b200.so (csrc/layout.cu), compared with the oracle bit for bit."""
import numpy as np
import pytest


def _reference_nms_vector():
    """the input of paddlex_layout_nms_matches_compacting_reference_on_dense_input, :1700-1721"""
    boxes = []
    for i in range(256):
        x, y, size = float((i * 37) % 80), float((i * 53) % 80), 18.0 + float(i % 11)
        boxes.append((x, y, x + size, y + size))
    boxes.append((float("nan"), 0.0, 10.0, 10.0))
    n = len(boxes)
    classes = np.array([i % 7 for i in range(n)], np.int32)
    scores = np.array([((i * 97) % 1000) / 1000.0 for i in range(n)], np.float32)
    return np.array(boxes, np.float32), classes, scores


def test_oracle_nms_reference_vector():
    """the reference asserts paddlex_layout_nms == compacting_nms_reference on this input; both restatements agree,
    and an independent numpy NMS with the PaddleX +1 IoU gives the same indices"""
    from oracle import cpu
    boxes, classes, scores = _reference_nms_vector()
    got = cpu.layout_nms(boxes, classes, scores)
    want = cpu.layout_nms(boxes, classes, scores, compacting=True)
    # (on the reference's lattice no pair reaches its threshold: the vector pins the ordering and the NaN handling)
    assert got.tolist() == want.tolist() and len(got) == 257
    assert got[:4].tolist() == [134, 103, 237, 72]  # scores 0.998, 0.991, 0.989, 0.984: stable descending order
    # a denser variant of the same generator, where suppression does happen
    boxes[:, :2] = boxes[:, :2] % 24
    boxes[:256, 2:] = boxes[:256, :2] + (18.0 + (np.arange(256) % 11)[:, None]).astype(np.float32)
    got = cpu.layout_nms(boxes, classes, scores)
    want = cpu.layout_nms(boxes, classes, scores, compacting=True)
    assert got.tolist() == want.tolist() and 10 < len(got) < 200

    def corners(b):  # x_min()/x_max() of from_coords: a NaN corner falls back to the other one
        x = np.array([b[0], b[2]])
        y = np.array([b[1], b[3]])
        return np.nanmin(x), np.nanmin(y), np.nanmax(x), np.nanmax(y)

    def iou(a, b):
        ax1, ay1, ax2, ay2 = corners(a)
        bx1, by1, bx2, by2 = corners(b)
        iw = max(np.float32(min(ax2, bx2) - max(ax1, bx1) + 1), 0)
        ih = max(np.float32(min(ay2, by2) - max(ay1, by1) + 1), 0)
        inter = np.float32(iw * ih)
        uni = np.float32((ax2 - ax1 + 1) * (ay2 - ay1 + 1) + (bx2 - bx1 + 1) * (by2 - by1 + 1) - inter)
        return inter / uni if uni > 0 else 0.0

    order = sorted(range(len(boxes)), key=lambda i: -scores[i])
    keep = []
    while order:
        cur = order.pop(0)
        keep.append(cur)
        order = [i for i in order if iou(boxes[cur], boxes[i]) < (0.6 if classes[i] == classes[cur] else 0.98)]
    assert keep == got.tolist()


def _random_predictions(rng, n, fdim, w, h, num_classes, normalised):
    """rows [class, score, x1, y1, x2, y2, (keys)] with nested, overlapping, degenerate and out-of-range boxes"""
    p = np.zeros((n, fdim), np.float32)
    p[:, 0] = rng.integers(-1, num_classes + 1, n)  # some ids out of range
    p[:, 1] = rng.random(n)
    sx, sy = (1.0, 1.0) if normalised else (w, h)
    cx, cy = rng.random(n) * sx, rng.random(n) * sy
    bw, bh = rng.random(n) * 0.5 * sx, rng.random(n) * 0.5 * sy
    p[:, 2], p[:, 3], p[:, 4], p[:, 5] = cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2
    for i in range(0, n - 1, 5):  # a box nested inside its predecessor, often of another class
        p[i + 1, 2:6] = p[i, 2:6] + np.array([1, 1, -1, -1], np.float32) * (0.01 * sx)
    for i in range(0, n - 1, 7):  # near duplicates (NMS)
        p[i + 1, 2:6] = p[i, 2:6] + np.float32(0.002 * sx)
        p[i + 1, 0] = p[i, 0]
    p[n // 2, 2:6] = (0.0, 0.0, sx, sy)  # a page-sized box
    p[n // 2, 0] = 1
    p[n // 3, 4] = p[n // 3, 2]  # zero width -> invalid
    p[n // 4, 3] = np.nan
    if fdim >= 7:
        p[:, 6] = rng.integers(0, 6, n)
    if fdim >= 8:
        p[:, 7] = rng.integers(0, 9, n)
    return p


CASES = [
    dict(),  # LayoutDetectionConfig::default
    dict(layout_nms=False, max_elements=7),
    dict(score_threshold=0.2, class_thresholds={0: 0.3, 7: 0.3, 2: 0.4, 16: 0.45}),  # with_pp_structurev3_thresholds
    dict(score_threshold=0.1, class_merge_modes={0: 0, 1: 0, 2: 2, 7: 0, 18: 0, 8: 2}, unclip=(1.0, 1.0)),  # v3 defaults
    dict(score_threshold=0.1, class_merge_modes={2: 1, 5: 1, 7: 0}),  # Small
    dict(score_threshold=0.3, unclip=(1.1, 0.9)),
    dict(score_threshold=0.3, unclip={2: (1.2, 1.2), 8: (1.0, 1.05)}, class_merge_modes={3: 2}),
    dict(score_threshold=-1.0, image_class_id=-1, formula_class_id=-1),
]


@pytest.mark.parametrize("case", range(len(CASES)))
@pytest.mark.parametrize("fdim", [6, 7, 8])
def test_product_equals_oracle(built_lib, case, fdim):
    """oar_layout_postprocess == the oracle restatement, bit for bit: kept rows, their order, coordinates, scores"""
    from oracle import cpu
    from oar_ocr_b200 import ffi
    kw = dict(image_class_id=1, formula_class_id=7)
    kw.update(CASES[case])
    rng = np.random.default_rng(100 * case + fdim)
    sizes = [(1024.0, 1024.0), (800.0, 1200.0), (1600.0, 900.0)]
    preds = np.stack([_random_predictions(rng, 300, fdim, w, h, 23, normalised=(b == 1))
                      for b, (w, h) in enumerate(sizes)])
    got = ffi.layout_postprocess(preds, sizes, 23, **kw)
    kept = 0
    for b, (w, h) in enumerate(sizes):
        ob, oc, os_ = cpu.layout_postprocess(preds[b], w, h, 23, **kw)
        gb, gc, gs = got[b]
        assert gc.tolist() == oc.tolist()
        assert np.array_equal(gb.view(np.uint32), ob.view(np.uint32))
        assert np.array_equal(gs.view(np.uint32), os_.view(np.uint32))
        kept += len(oc)
    assert kept > 10


def test_postprocess_semantics(built_lib):
    """hand-checkable page: threshold, normalised coordinates, same-class NMS, page-sized image box, containment
    (Large), reading-order keys, max_elements"""
    from oar_ocr_b200 import ffi
    W, H = 1000.0, 800.0
    rows = np.array([
        # class score x1    y1    x2    y2    key
        [2, 0.90, 0.10, 0.10, 0.50, 0.30, 3],   # text (normalised page)
        [2, 0.80, 0.10, 0.10, 0.50, 0.31, 1],   # near duplicate of the first: suppressed (same class, IoU > 0.6)
        [1, 0.95, 0.00, 0.00, 1.00, 1.00, 0],   # "image" covering the whole page: dropped
        [0, 0.70, 0.12, 0.12, 0.30, 0.20, 2],   # paragraph_title inside the text box
        [2, 0.40, 0.60, 0.60, 0.90, 0.90, 1],   # below the 0.5 threshold
        [8, 0.60, 0.55, 0.55, 0.95, 0.95, 0],   # table
        [30, 0.99, 0.1, 0.1, 0.2, 0.2, 0],      # class id out of range
    ], np.float32)
    (b, c, s), = ffi.layout_postprocess(rows[None], [(W, H)], 23, image_class_id=1, formula_class_id=7)
    assert c.tolist() == [8, 0, 2]  # sorted by the key column: 0, 2, 3
    assert np.allclose(b[2], [100.0, 80.0, 500.0, 240.0]) and s.tolist() == pytest.approx([0.6, 0.7, 0.9])
    # Large merge on "text": the title mostly inside the text box goes away
    (b, c, s), = ffi.layout_postprocess(rows[None], [(W, H)], 23, image_class_id=1, formula_class_id=7,
                                        class_merge_modes={2: ffi.MERGE_LARGE})
    assert c.tolist() == [8, 2]
    (b, c, s), = ffi.layout_postprocess(rows[None], [(W, H)], 23, image_class_id=1, max_elements=1)
    assert c.tolist() == [8]
    # pixel coordinates are clamped to the page, 6 columns = no reordering
    px = np.array([[2, 0.9, -20.0, 10.0, 400.0, 300.0], [8, 0.8, 500.0, 500.0, 5000.0, 700.0]], np.float32)
    (b, c, s), = ffi.layout_postprocess(px[None], [(W, H)], 23)
    assert b.tolist() == [[0.0, 10.0, 400.0, 300.0], [500.0, 500.0, 1000.0, 700.0]] and c.tolist() == [2, 8]
    # unclip about the centre
    (b, c, s), = ffi.layout_postprocess(px[None], [(W, H)], 23, unclip=(1.5, 1.0))
    assert b[0].tolist() == [-100.0, 10.0, 500.0, 300.0]


def test_postprocess_edges_and_errors(built_lib):
    from oar_ocr_b200 import ffi
    assert ffi.layout_postprocess(np.zeros((0, 300, 6), np.float32), [], 23) == []
    (b, c, s), = ffi.layout_postprocess(np.zeros((1, 0, 6), np.float32), [(10.0, 10.0)], 23)
    assert len(c) == 0
    with pytest.raises(ffi.OCRError):
        ffi.layout_postprocess(np.zeros((1, 3, 5), np.float32), [(10.0, 10.0)], 23)  # fewer than 6 columns
    with pytest.raises(ffi.OCRError):
        ffi.layout_postprocess(np.zeros((1, 3, 6), np.float32), [(10.0, 10.0)], 0)
    cfg = ffi.LayoutConfig()
    ffi.lib().oar_layout_config_default(cfg)
    assert (cfg.score_threshold, cfg.max_elements, cfg.layout_nms, cfg.num_classes) == (0.5, 100, 1, 23)
    assert (cfg.image_class_id, cfg.formula_class_id, cfg.unclip_mode) == (1, 7, ffi.UNCLIP_NONE)


def test_mirror_resolves_labels(built_lib):
    """the Python mirror of the adapter: label-keyed thresholds / merge modes -> class ids, elements carry labels"""
    from oar_ocr_b200.ocr import LayoutDetectionConfig, OCRError, postprocess_pp_doclayout
    rows = np.array([[2, 0.45, 0.1, 0.1, 0.5, 0.3], [0, 0.35, 0.12, 0.12, 0.3, 0.2], [8, 0.45, 0.6, 0.6, 0.9, 0.9]],
                    np.float32)
    els, = postprocess_pp_doclayout(rows[None, :, None, :], [(1000.0, 800.0)],
                                    LayoutDetectionConfig.with_pp_structurev3_thresholds())
    assert [e.element_type for e in els] == ["text", "paragraph_title"]  # table 0.45 < 0.5, text 0.45 >= 0.4
    assert els[0].bbox.x_min() == 100.0 and els[0].bbox.y_max() == pytest.approx(240.0, abs=1e-3)  # f32: 0.3 * 800
    cfg = LayoutDetectionConfig(score_threshold=0.3, class_merge_modes={"text": "large", "no_such_label": "small"})
    els, = postprocess_pp_doclayout(rows[None], [(1000.0, 800.0)], cfg)
    assert [e.element_type for e in els] == ["text", "table"]
    with pytest.raises(OCRError):
        postprocess_pp_doclayout(rows[None], [(1000.0, 800.0)], LayoutDetectionConfig(max_elements=0))
