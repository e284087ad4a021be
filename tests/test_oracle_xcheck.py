"""Cross-checks of the oracle's restatements of un-vendored third-party arithmetic against independent
implementations available offline (OpenCV, torch).  The crates themselves (imageproc 0.27, clipper2-rust 1.0.3,
image 0.25, nalgebra 0.35) are absent from /root/reference, so these are consistency checks, not golden vectors:
parity for those four stays "unpinned" (DESIGN.md).
"""
import numpy as np
import pytest

from oracle import cpu

cv2 = pytest.importorskip("cv2")


def _blobs(seed, h=96, w=128, n=7, holes=True):
    rng = np.random.default_rng(seed)
    m = np.zeros((h, w), np.uint8)
    for _ in range(n):
        cx, cy = int(rng.integers(12, w - 12)), int(rng.integers(12, h - 12))
        ax, ay = int(rng.integers(3, 11)), int(rng.integers(3, 9))
        cv2.ellipse(m, (cx, cy), (ax, ay), float(rng.uniform(0, 180)), 0, 360, 255, -1)
        if holes and ax > 5 and ay > 4:
            cv2.circle(m, (cx, cy), 2, 0, -1)
    m[[0, -1], :] = 0
    m[:, [0, -1]] = 0
    return m


@pytest.mark.parametrize("seed", range(6))
def test_find_contours_matches_opencv_border_sets(seed):
    """Suzuki-Abe is also what cv2.findContours implements: same borders (outer + hole), same pixels."""
    m = _blobs(seed)
    ours, btypes = cpu.find_contours(m)
    theirs, _ = cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    a = sorted((len(c), tuple(sorted(map(tuple, c.tolist())))) for c in ours)
    b = sorted((len(c), tuple(sorted(map(tuple, c.reshape(-1, 2).tolist())))) for c in theirs)
    assert a == b
    # discovery order = raster order of each border's first pixel, which is also its start point
    starts = [(c[0][1], c[0][0]) for c in ours]
    assert starts == sorted(starts)
    n_outer = 0
    for c in ours:  # an outer border starts at its raster-first pixel; a hole border starts left of the hole
        if (c[0][1], c[0][0]) == min((p[1], p[0]) for p in c.tolist()):
            n_outer += 1
    assert n_outer >= len(cv2.findContours(m, cv2.RETR_EXTERNAL, cv2.CHAIN_APPROX_NONE)[0])


def _noise_map(seed, h=120, w=160):
    rng = np.random.default_rng(seed)
    m = (rng.random((h, w)) > 0.55).astype(np.uint8) * 255
    return cv2.morphologyEx(m, cv2.MORPH_OPEN, np.ones((2, 2), np.uint8))


@pytest.mark.parametrize("seed", range(6))
def test_find_contours_equals_opencv_point_for_point(seed):
    """Stronger than the border sets: every contour is the SAME SEQUENCE of points as OpenCV's (same start pixel, same
    direction, same repeated pixels on one-pixel-wide parts), and the contours come in the same discovery order
    (cv2.findContours with RETR_LIST prepends, so its list is ours reversed)."""
    m = _blobs(seed)
    ours, _ = cpu.find_contours(m)
    theirs, _ = cv2.findContours(m, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    assert [c.tolist() for c in ours] == [c.reshape(-1, 2).tolist() for c in theirs][::-1]


@pytest.mark.parametrize("seed", range(4))
def test_find_contours_equals_opencv_on_noise_away_from_the_frame(seed):
    """~400 contours of thresholded noise (holes inside holes, one-pixel bridges, diagonal contacts): identical to
    OpenCV point for point and in order once no component touches the image frame.  With components ON the frame the
    two differ, and only there: imageproc 0.27 starts no outer border at x = 0 and applies its right-edge rule at
    x = W - 1 (db_bitmap.rs:84-150 calls it on the un-padded mask; the oracle restates that), OpenCV pads the image
    with a background frame first."""
    raw = _noise_map(seed)
    framed = raw.copy()
    framed[[0, -1], :] = 0
    framed[:, [0, -1]] = 0
    ours, _ = cpu.find_contours(framed)
    theirs, _ = cv2.findContours(framed, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    assert len(ours) >= 300
    assert [c.tolist() for c in ours] == [c.reshape(-1, 2).tolist() for c in theirs][::-1]
    ours, _ = cpu.find_contours(raw)
    theirs, _ = cv2.findContours(raw, cv2.RETR_LIST, cv2.CHAIN_APPROX_NONE)
    assert len(ours) == len(theirs)
    h, w = raw.shape
    differing = 0
    for a, b in zip([c.tolist() for c in ours], [c.reshape(-1, 2).tolist() for c in theirs][::-1]):
        if a != b:
            differing += 1
            xs, ys = [p[0] for p in a], [p[1] for p in a]
            assert min(xs) == 0 or max(xs) == w - 1 or min(ys) == 0 or max(ys) == h - 1, "differs away from the frame"
    assert differing >= 1  # the frame rule is exercised


def test_find_contours_single_pixel_and_order():
    m = np.zeros((8, 8), np.uint8)
    m[2, 5] = 255
    m[4, 1:4] = 255
    cs, _ = cpu.find_contours(m)
    assert [c.tolist() for c in cs][0] == [[5, 2]]
    assert cs[1][0].tolist() == [1, 4]


@pytest.mark.parametrize("dims", [((60, 200), (48, 160)), ((31, 700), (48, 1084)), ((48, 320), (48, 320)),
                                  ((100, 90), (50, 45))])
def test_resize_triangle_close_to_antialiased_bilinear(dims):
    """image 0.25 Triangle = separable tent filter whose support grows with the down-scale ratio: the same
    definition torch uses for bilinear(antialias=True).  Rounding differs by at most one level."""
    import torch
    import torch.nn.functional as F
    (h, w), (nh, nw) = dims
    rng = np.random.default_rng(h * 1000 + w)
    img = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
    ours = cpu.resize_triangle(img, nw, nh).astype(np.int32)
    t = torch.from_numpy(img).permute(2, 0, 1)[None].float()
    ref = F.interpolate(t, size=(nh, nw), mode="bilinear", antialias=True, align_corners=False)
    ref = ref[0].permute(1, 2, 0).round().clamp(0, 255).numpy().astype(np.int32)
    if (h, w) == (nh, nw):
        assert np.array_equal(ours, img)
    assert np.abs(ours - ref).max() <= 1
    assert (ours != ref).mean() < 0.02


@pytest.mark.parametrize("dims", [((60, 200), (48, 160)), ((31, 700), (48, 1084)), ((100, 37), (48, 18)),
                                  ((20, 90), (48, 216)), ((200, 800), (48, 192)), ((48, 320), (48, 320))])
def test_resize_triangle_within_one_level_of_pillow(dims):
    """Pillow's BILINEAR resize is the same algorithm family as image 0.25's FilterType::Triangle (separable, the
    filter's support scaled by the reduction factor), with fixed-point coefficients and an 8-bit intermediate where
    the crate keeps f32: the two may differ by one grey level, never by more, shrinking or enlarging."""
    PIL = pytest.importorskip("PIL.Image")
    (sh, sw), (dh, dw) = dims
    img = np.random.default_rng(sh * 1000 + sw).integers(0, 256, (sh, sw, 3), dtype=np.uint8)
    ours = cpu.resize_triangle(img, dw, dh)
    theirs = np.asarray(PIL.fromarray(img).resize((dw, dh), PIL.BILINEAR))
    d = np.abs(ours.astype(np.int32) - theirs.astype(np.int32))
    assert d.max() <= 1
    if (sh, sw) == (dh, dw):
        assert d.max() == 0  # identity resize copies


def test_perspective_transform_matches_opencv():
    rng = np.random.default_rng(3)
    for _ in range(20):
        src = np.array([[0, 0], [200, 0], [200, 40], [0, 40]], np.float32) + rng.uniform(-6, 6, (4, 2)).astype(
            np.float32)
        dst = np.array([[0, 0], [203, 0], [203, 41], [0, 41]], np.float32)
        ours = cpu.perspective_transform(src, dst)
        ref = cv2.getPerspectiveTransform(src, dst)
        assert np.allclose(ours, ref, rtol=2e-3, atol=2e-3)


def test_min_area_rect_matches_opencv_area():
    rng = np.random.default_rng(5)
    for _ in range(20):
        pts = rng.uniform(0, 100, (12, 2)).astype(np.float32)
        cx, cy, w, h, _ang = cpu.min_area_rect(pts)
        (_, (rw, rh), _) = cv2.minAreaRect(pts)
        assert abs(w * h - rw * rh) <= 1e-2 * rw * rh


def test_unclip_grows_rect_by_delta():
    """Clipper2 round-join offset of a convex quad: its min-area rect is the quad grown by delta per side."""
    for (w, h, ang) in [(200, 30, 0.0), (120, 24, 3.0), (300, 40, -2.5), (50, 50, 0.0)]:
        c, s = np.cos(np.deg2rad(ang)), np.sin(np.deg2rad(ang))
        base = np.array([[-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2], [-w / 2, h / 2]])
        pts = (base @ np.array([[c, s], [-s, c]]) + [400, 300]).astype(np.float32)
        delta = (w * h) * 2.0 / (2 * (w + h))
        poly = cpu.unclip(pts, 2.0)
        assert len(poly) > 8
        _, _, rw, rh, _ = cpu.min_area_rect(poly)
        lo, hi = sorted([rw, rh])
        assert abs(lo - (min(w, h) + 2 * delta)) < 0.05
        assert abs(hi - (max(w, h) + 2 * delta)) < 0.05
        # every offset vertex is at distance ~delta from the original polygon (0.01 grid)
        d = [cv2.pointPolygonTest(pts.reshape(-1, 1, 2), (float(x), float(y)), True) for x, y in poly]
        assert np.all(np.abs(np.abs(d) - delta) < 0.02)


def test_box_score_fast_matches_polygon_mask_mean():
    rng = np.random.default_rng(9)
    pred = rng.random((80, 120), dtype=np.float32)
    box = np.array([[10.3, 12.2], [90.7, 15.9], [89.1, 50.4], [9.2, 47.0]], np.float32)
    got = cpu.box_score_fast(pred, box)
    ys, xs = np.mgrid[0:80, 0:120]
    inside = np.array([[cv2.pointPolygonTest(box.reshape(-1, 1, 2), (float(x), y + 0.5), False) >= 0 for x in
                        range(120)] for y in range(80)])
    assert abs(got - pred[inside].mean()) < 2e-2
