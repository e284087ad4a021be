"""Pins the CPU oracle against the reference's own known-answer tests.

Every test names the reference unit test whose vectors it restates (paths under the
reference repo, GreatV/oar-ocr v0.9.3).  No GPU, no product code: oracle only.
"""
import numpy as np

from oracle import cpu


def _px(r, g, b):
    return np.array([[[r, g, b]]], np.uint8)


def _norm(img, scale, mean, std, order, layout="chw"):
    """NormalizeImage::with_color_order(scale, mean, std, layout, order).apply: mean/std are given in
    OUTPUT channel order (normalization.rs:36-41, 142-143)."""
    a, b = cpu.norm_coeffs(np.float32(scale), mean, std)
    src = (0, 1, 2) if order == "rgb" else (2, 1, 0)
    return cpu.normalize(img, a, b, src, layout)


# normalization.rs:498-526 test_normalize_image_color_order_rgb_vs_bgr_chw
def test_color_order_rgb_vs_bgr_chw():
    img = _px(10, 20, 30)
    assert _norm(img, 1.0, [0, 0, 0], [1, 1, 1], "rgb").ravel().tolist() == [10.0, 20.0, 30.0]
    assert _norm(img, 1.0, [0, 0, 0], [1, 1, 1], "bgr").ravel().tolist() == [30.0, 20.0, 10.0]


# normalization.rs:528-555 test_normalize_image_mean_std_applied_in_output_channel_order
def test_mean_std_in_output_channel_order():
    img = _px(11, 22, 33)
    assert _norm(img, 1.0, [1, 2, 3], [2, 4, 5], "rgb").ravel().tolist() == [5.0, 5.0, 6.0]
    assert _norm(img, 1.0, [3, 2, 1], [5, 4, 2], "bgr").ravel().tolist() == [6.0, 5.0, 5.0]


# normalization.rs:621-683 test_normalize_batch_to_preserves_batch_and_layout_semantics
def test_batch_and_layout_semantics():
    def mk(base0):
        img = np.zeros((2, 2, 3), np.uint8)
        for y in range(2):
            for x in range(2):
                base = (y * 2 + x) * 3 + base0
                img[y, x] = (base, base + 1, base + 2)
        return img

    a, b = mk(1), mk(21)
    chw = np.stack([_norm(a, 1.0, [0, 0, 0], [1, 1, 1], "rgb"), _norm(b, 1.0, [0, 0, 0], [1, 1, 1], "rgb")])
    assert chw.shape == (2, 3, 2, 2)
    assert chw.ravel().tolist() == [1, 4, 7, 10, 2, 5, 8, 11, 3, 6, 9, 12, 21, 24, 27, 30, 22, 25, 28, 31, 23, 26, 29,
                                    32]
    hwc = np.stack([_norm(a, 1.0, [0, 0, 0], [1, 1, 1], "rgb", "hwc"),
                    _norm(b, 1.0, [0, 0, 0], [1, 1, 1], "rgb", "hwc")])
    assert hwc.shape == (2, 2, 2, 3)
    assert hwc.ravel().tolist() == list(range(1, 13)) + list(range(21, 33))


def _make_rgb(w, h):
    # simd.rs tests' make_rgb is a deterministic byte ramp; any deterministic bytes pin scalar == vector
    i = np.arange(w * h * 3, dtype=np.uint32)
    return ((i * 31 + 7) % 256).astype(np.uint8).reshape(h, w, 3)


# simd.rs:356-372 chw_simd_matches_scalar_rgb_and_bgr: separate multiply then add, no FMA
def test_normalize_is_mul_then_add_f32():
    w, h = 37, 19
    rgb = _make_rgb(w, h)
    alpha = np.array([1.0 / 255.0, 0.5, 2.0], np.float32)
    beta = np.array([-0.485, 0.1, -1.0], np.float32)
    for src in ([0, 1, 2], [2, 1, 0]):
        got = cpu.normalize(rgb, alpha, beta, src, "chw")
        for c in range(3):
            want = rgb[:, :, src[c]].astype(np.float32) * alpha[c] + beta[c]  # numpy f32: two roundings
            assert np.array_equal(got[c], want)
        got_hwc = cpu.normalize(rgb, alpha, beta, src, "hwc")
        assert np.array_equal(got_hwc.transpose(2, 0, 1), got)


# normalization.rs:685-709 normalize_batch_refs_matches_owned_path_bit_exact: DB constants
def test_det_normalize_constants():
    x, y = np.meshgrid(np.arange(96), np.arange(64))
    img = np.stack([x % 251, y % 241, (x + y) % 239], -1).astype(np.uint8)
    got = cpu.det_normalize(img)
    scale = np.float32(1.0) / np.float32(255.0)
    mean = np.array([0.485, 0.456, 0.406], np.float32)
    std = np.array([0.229, 0.224, 0.225], np.float32)
    for c in range(3):
        want = img[:, :, 2 - c].astype(np.float32) * (scale / std[c]) + (-mean[c] / std[c])
        assert np.array_equal(got[c], want)


# simd.rs:389-403 argmax_matches_scalar_including_ties
def test_argmax_last_max_wins():
    row = np.array([((i * 17) % 13) * 0.5 for i in range(101)], np.float32)
    idx, prob = cpu.ctc_argmax(row[None, None, :])
    want = max(range(101), key=lambda i: (row[i], i))
    assert idx[0, 0] == want and prob[0, 0] == row[want]
    tied = np.array([1, 2, 5, 9, 4, 8, 9, 0], np.float32)
    idx, prob = cpu.ctc_argmax(tied[None, None, :])
    assert (idx[0, 0], prob[0, 0]) == (6, 9.0)
    idx, prob = cpu.ctc_argmax(np.array([[[42.0]]], np.float32))
    assert (idx[0, 0], prob[0, 0]) == (0, 42.0)


# simd.rs:405-429 crnn_simd_matches_scalar_with_padding
def test_crnn_normalize_with_padding():
    rw, h = 21, 48
    rgb = _make_rgb(rw, h)
    # a crop that is already 48 high resizes to itself (image::imageops::resize same-size = copy)
    wide = _make_rgb(400, 48)  # forces tensor_w = 400 > 21
    x = cpu.crnn_preprocess([rgb, wide])
    assert x.shape == (2, 3, 48, 400)
    v = rgb.astype(np.float32)
    want = (v / np.float32(255.0) - np.float32(0.5)) / np.float32(0.5)
    for c in range(3):
        assert np.array_equal(x[0, c, :, :rw], want[:, :, 2 - c])
    assert np.all(x[0, :, :, rw:] == 0.0)


# decode.rs:692-745 compact_argmax_preserves_ctc_text_scores_and_positions
def test_ctc_decode_texts_scores_columns():
    winners = [[(0, 0.9), (1, 0.8), (1, 0.7), (0, 0.6), (1, 0.5), (2, 0.4), (2, 0.3)],
               [(3, 0.95), (3, 0.85), (4, 0.75), (3, 0.65), (0, 0.55), (2, 0.45), (0, 0.35)]]
    logits = np.full((2, 7, 5), -10.0, np.float32)
    for b, seq in enumerate(winners):
        for t, (k, p) in enumerate(seq):
            logits[b, t, k] = p
    chars = ["blank", "a", "b", "c"]  # from_string_list(dict, use_space_char=false): slot 4 is out of dictionary
    idx, prob = cpu.ctc_argmax(logits)
    labels, scores, cols, T = cpu.ctc_decode(idx, prob, len(chars))
    texts = ["".join(chars[k] for k in lab) for lab in labels]
    assert texts == ["aab", "ccb"]
    f = np.float32
    assert scores[0] == (f(0.8) + f(0.5) + f(0.4)) / f(3.0)
    assert scores[1] == (f(0.95) + f(0.65) + f(0.45)) / f(3.0)
    assert [c.tolist() for c in cols] == [[1, 4, 5], [0, 3, 5]]
    assert T == 7
    positions = [[f(c) / f(T) for c in cc] for cc in cols]
    assert positions[0] == [f(1.0) / f(7.0), f(4.0) / f(7.0), f(5.0) / f(7.0)]


# decode.rs:747-758 compact_argmax_preserves_empty_tensor_behavior
def test_ctc_empty_tensor():
    idx, prob = cpu.ctc_argmax(np.zeros((2, 0, 5), np.float32))
    assert idx.size == 0 and prob.size == 0
    labels, scores, cols, T = cpu.ctc_decode(idx, prob, 5)
    assert labels == [] and len(scores) == 0 and cols == []


# db_bitmap.rs:375-391 test_paddlex_order_mini_box_points
def test_order_mini_box_points():
    got = cpu.order_mini_box([[20, 20], [10, 10], [20, 10], [10, 20]])
    assert got.tolist() == [[10, 10], [20, 10], [20, 20], [10, 20]]


# db_bitmap.rs:393-407 test_get_mini_boxes_from_points_returns_min_side
def test_mini_boxes_min_side():
    box, min_side = cpu.mini_boxes_from_points([[0, 0], [10, 0], [10, 5], [0, 5]])
    assert abs(min_side - 5.0) < 1e-3
    assert sorted(map(tuple, np.round(box).tolist())) == [(0, 0), (0, 5), (10, 0), (10, 5)]


# db_bitmap.rs:409-423 test_simplify_chain_points_removes_straight_segment_points
def test_simplify_chain_points():
    pts = [[0, 0], [1, 0], [2, 0], [2, 1], [2, 2], [1, 2], [0, 2], [0, 1]]
    assert len(cpu.simplify_chain(pts)) == 4


def _from_coords(x1, y1, x2, y2):
    return [[x1, y1], [x2, y1], [x2, y2], [x1, y2]]


# sorting.rs:740-753 / 755-766 / 768-786 / 800-808
def test_sort_quad_boxes_reference_cases():
    boxes, _ = cpu.sort_quad_boxes([_from_coords(10, 50, 50, 70), _from_coords(10, 10, 50, 30),
                                    _from_coords(10, 30, 50, 50)])
    assert [b[:, 1].min() for b in boxes] == [10.0, 30.0, 50.0]
    boxes, _ = cpu.sort_quad_boxes([_from_coords(60, 10, 100, 30), _from_coords(10, 12, 50, 32)])
    assert boxes[0][:, 0].min() < boxes[1][:, 0].min()
    boxes, order = cpu.sort_quad_boxes([_from_coords(60, 10, 100, 30), _from_coords(10, 11, 50, 31),
                                        _from_coords(10, 50, 50, 70), _from_coords(60, 52, 100, 72)])
    assert order.tolist() == [1, 0, 2, 3]
    boxes, order = cpu.sort_quad_boxes(np.zeros((0, 4, 2), np.float32))
    assert len(boxes) == 0


# transform.rs:699-716 exact_axis_aligned_fast_path_is_deliberately_strict
def test_axis_aligned_predicate_is_strict():
    exact = np.array([[0, 0], [50, 0], [50, 30], [0, 30]], np.float32)
    assert cpu.is_exact_axis_aligned(exact, 50, 30)
    skew = exact.copy()
    skew[1, 1] = 0.001
    assert not cpu.is_exact_axis_aligned(skew, 50, 30)
    frac = exact.copy()
    frac[0, 0] = 0.5
    assert not cpu.is_exact_axis_aligned(frac, 50, 30)


# transform.rs:618-640 test_get_perspective_transform (finite) + :718-728 singular matrix -> error
def test_perspective_transform_finite_and_singular():
    m = cpu.perspective_transform([[0, 0], [1, 0], [1, 1], [0, 1]], [[0, 0], [2, 0], [2, 2], [0, 2]])
    assert m is not None and np.all(np.isfinite(m))
    assert np.allclose(m, np.diag([2, 2, 1]), atol=1e-5)
    # four collinear source points: no homography exists
    assert cpu.rotate_crop(np.zeros((8, 8, 3), np.uint8), [[1, 1], [2, 2], [3, 3], [4, 4]]) is None


def _cubic(t):
    f = np.float32
    a = f(-0.5)
    t = abs(f(t))
    if t <= 1:
        return (a + f(2)) * t * t * t - (a + f(3)) * t * t + f(1)
    if t < 2:
        return a * t * t * t - f(5) * a * t * t + f(8) * a * t - f(4) * a
    return f(0)


def _bicubic_reference(img, x, y):
    """transform.rs:543-577 bicubic_reference (the test-local ground truth of the reference)"""
    f = np.float32
    h, w, _ = img.shape
    xi, yi = int(np.floor(f(x))), int(np.floor(f(y)))
    dx, dy = f(x) - f(xi), f(y) - f(yi)
    wx = [_cubic(dx + f(1)), _cubic(dx), _cubic(dx - f(1)), _cubic(dx - f(2))]
    wy = [_cubic(dy + f(1)), _cubic(dy), _cubic(dy - f(1)), _cubic(dy - f(2))]
    res = [f(0), f(0), f(0)]
    for j in range(4):
        sy = min(max(yi - 1 + j, 0), h - 1)
        for i in range(4):
            sx = min(max(xi - 1 + i, 0), w - 1)
            wgt = f(wx[i] * wy[j])
            for c in range(3):
                res[c] = f(res[c] + f(wgt * f(img[sy, sx, c])))
    out = []
    for c in range(3):
        r = np.float32(np.floor(abs(res[c]) + f(0.5)) * np.sign(res[c]))  # f32::round = half away from zero
        out.append(int(min(max(r, 0), 255)))
    return out


# transform.rs:579-608 bicubic_raw_buffer_matches_reference_bit_exact (17x11 LCG image, incl. out of bounds)
def test_bicubic_matches_reference_formula():
    w, h = 17, 11
    i = np.arange(w * h, dtype=np.uint64).reshape(h, w)
    img = np.stack([(i * 37 + 11) % 256, (i * 59 + 7) % 256, (i * 101 + 3) % 256], -1).astype(np.uint8)
    for yi in range(-3, h + 3, 2):
        for xi in range(-3, w + 3, 2):
            for fx in (0.0, 0.25, 0.5, 0.75):
                for fy in (0.0, 0.33, 0.66):
                    x, y = np.float32(xi) + np.float32(fx), np.float32(yi) + np.float32(fy)
                    assert cpu.bicubic(img, x, y).tolist() == _bicubic_reference(img, x, y), (x, y)


# processors.rs:282-302 test_parallel_text_cropping_preserves_detection_order (axis-aligned fast path)
def test_axis_aligned_crops_in_order():
    img = np.zeros((4, 64, 3), np.uint8)
    img[:, :, 0] = (np.arange(64) // 4)[None, :]
    for index in range(16):
        x = index * 4
        crop = cpu.rotate_crop(img, _from_coords(x, 0, x + 4, 4))
        assert crop.shape == (4, 4, 3)
        assert crop[0, 0].tolist() == [index, 0, 0]


# transform.rs:671-697 test_get_rotate_crop_image_success
def test_rotate_crop_square():
    img = np.zeros((4, 4, 3), np.uint8)
    for y in range(4):
        for x in range(4):
            img[y, x] = (x * 64, y * 64, (x + y) * 32)
    crop = cpu.rotate_crop(img, _from_coords(1, 1, 3, 3))
    assert crop is not None and crop.shape[0] > 0 and crop.shape[1] > 0
    assert np.array_equal(crop, img[1:3, 1:3])


# resize_detection.rs:243-319: 960/Max/4000 -> multiples of 32, identity at 960x960 and 640x640
def test_det_resize_dims():
    assert cpu.det_resize_dims(960, 960) == (960, 960)
    assert cpu.det_resize_dims(640, 640) == (640, 640)
    assert cpu.det_resize_dims(1920, 1080) == (960, 544)  # ratio .5 -> 960x540 -> (540+16)//32*32 = 544
    assert cpu.det_resize_dims(100, 30) == (96, 32)
    assert cpu.det_resize_dims(10, 10) == (32, 32)


# ocr.rs:1197-1232 test_ctc_word_boxes_logic, for the oracle restatement and the product-side mirror alike
def test_ctc_word_boxes_kat_and_mirror_agreement():
    from oar_ocr_b200.ocr import BoundingBox, ctc_word_boxes as product_word_boxes
    from oracle.pipeline import ctc_word_boxes as oracle_word_boxes
    line = np.array([[0, 0], [100, 0], [100, 20], [0, 20]], np.float32)
    got = oracle_word_boxes(line, "ABC", [1, 4, 7], 10, 5.0, 5.0)
    assert got.shape == (3, 4)
    assert np.allclose(got[:, 0], [0.0, 30.0, 60.0], atol=1e-5) and np.allclose(got[:, 2], [30.0, 60.0, 100.0], atol=1e-5)
    assert np.all(got[:, 1] == 0.0) and np.all(got[:, 3] == 20.0)
    # CJK: a box of the average character width around the cell centre, clamped to the line
    cjk = oracle_word_boxes(line, "一二", [0, 9], 10, 5.0, 5.0)
    assert np.allclose(cjk, [[0.0, 0, 30.0, 20], [70.0, 0, 100.0, 20]], atol=1e-5)
    # padding undone through wh_ratio / max_wh_ratio: half-width crop in a batch of full-width ones
    half = oracle_word_boxes(line, "AB", [1, 3], 10, 2.5, 5.0)  # effective columns = 5, cell = 20
    assert np.allclose(half[:, 0], [0.0, 50.0], atol=1e-5) and np.allclose(half[:, 2], [50.0, 100.0], atol=1e-5)
    assert len(oracle_word_boxes(line, "", [1], 10, 5.0, 5.0)) == 0 and len(oracle_word_boxes(line, "A", [], 10, 5, 5)) == 0
    # the product-side mirror (oar_ocr_b200/ocr.py) computes the same f32 values
    rng = np.random.default_rng(5)
    for _ in range(50):
        n = int(rng.integers(1, 12))
        cols = np.sort(rng.choice(40, n, replace=False))
        text = "".join(rng.choice(["A", "b", "中", "文", "7"]) for _ in range(n))
        x0, y0 = float(rng.uniform(0, 50)), float(rng.uniform(0, 50))
        box = np.array([[x0, y0], [x0 + 200.5, y0 + 3], [x0 + 199, y0 + 40], [x0 - 2, y0 + 37]], np.float32)
        r, mr = float(rng.uniform(1, 12)), 12.5
        a = oracle_word_boxes(box, text, cols, 40, r, mr)
        b = product_word_boxes(BoundingBox(box), text, cols, 40, r, mr)
        assert len(a) == len(b) == n
        for row, bb in zip(a, b):
            # from_coords keeps the corners as computed (a clamped box may have x0 > x1): compare the corners
            assert row[0] == bb.points[0, 0] and row[2] == bb.points[1, 0]
            assert row[1] == bb.points[0, 1] and row[3] == bb.points[2, 1]
