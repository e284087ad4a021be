"""Round-2 parity additions ("Tighten parity where it is cheap"):
  * the recognition score filter actually taken on the CUDA path (text_recognition_adapter.rs:88-102),
  * the bench configuration itself (32 / 256 batch sizes, 960x960 bench pages) against the oracle on 8 pages,
  * stage-by-stage edge cases of SURVEY.md App. C: blobs touching x = 0 / x = W-1 / y = 0 / y = H-1 (the border
    ambiguity of imageproc's find_contours), tall boxes that get_rotate_crop_image turns by 270 degrees
    (transform.rs:76-502), and more than 4096 crops in one call (the flush of recognize_global, ocr.rs:802-897).
Everything goes through the C ABI (ctypes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-3


@pytest.fixture(scope="module")
def nets(ctx, det_blob, rec_blob):
    from oar_ocr_b200 import ffi
    return ffi.Model(ctx, det_blob), ffi.Model(ctx, rec_blob)


@pytest.fixture(scope="module")
def oracle_nets(det_blob, rec_blob):
    from oracle.net import OracleNet
    return OracleNet(det_blob), OracleNet(rec_blob)


def _compare(got, want, tol=TOL, max_tie_frac=0.0):
    """boxes, reading order and CTC label sequences identical to the oracle, confidences within `tol`.
    max_tie_frac > 0 (the large runs only): a label sequence may differ in a region where the ORACLE's own smallest
    top-1 / top-2 probability gap is below `tol` -- an arg-max tie at the float tolerance north_star states, which any
    fp32 run with another summation order (ONNX Runtime included) can resolve either way -- and at most that fraction
    of the regions may be such ties."""
    n = ties = 0
    for g, w in zip(got, want):
        assert len(g.text_regions) == len(w), (len(g.text_regions), len(w))
        for r, o in zip(g.text_regions, w):
            assert np.array_equal(r.bounding_box.points, o["box"])
            assert r.detection_index == o["det_index"]
            if not np.array_equal(r.label_indices, o["labels"]):
                assert max_tie_frac > 0 and o["margin"] < tol, (r.label_indices, o["labels"], o["margin"])
                ties += 1
                n += 1
                continue  # the confidence averages over the emitted characters, which differ in a tie
            assert abs(r.confidence - o["score"]) <= tol
            n += 1
    assert ties <= max_tie_frac * n, (ties, n)
    return n


@pytest.mark.parametrize("thresh", [0.15, 0.28, 0.45])
def test_rec_score_filter_taken(nets, oracle_nets, thresh):
    """a17: regions whose mean CTC confidence is below the threshold keep their box and score but lose their text;
    the thresholds are chosen so that the filter drops some regions and keeps others on these pages"""
    from oracle import pipeline
    from oar_ocr_b200 import synth
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    imgs = [synth.page(70 + i, 480) for i in range(4)]
    ocr = OAROCR(det.ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0),
                 TextRecognitionConfig(score_threshold=thresh), 4, 16)
    got = ocr.predict(imgs)
    want = pipeline.predict(oracle_nets[0], oracle_nets[1], imgs, 18385, image_batch_size=4, region_batch_size=16,
                            rec_score_thresh=thresh)
    # regions within the tolerance of the threshold could legitimately differ: none on these pages
    margin = min(abs(o["score"] - thresh) for w in want for o in w)
    assert margin > 2 * TOL, margin
    assert _compare(got, want) >= 20
    scores = [o["score"] for w in want for o in w]
    dropped = sum(1 for w in want for o in w if o["score"] < thresh)
    kept = sum(1 for w in want for o in w if o["score"] >= thresh and len(o["labels"]))
    assert dropped >= 10 and kept >= 5, (dropped, kept, min(scores), max(scores))
    for g in got:
        for r in g.text_regions:
            if r.confidence < thresh:
                assert len(r.label_indices) == 0


def test_bench_config_pages_match_oracle(nets, oracle_nets):
    """configs[1] at the bench's own settings: 8 of the bench pages (seeds 0..7 = bench.make_pages(0, 32)[:8]),
    image_batch_size 32, region_batch_size 256 -- boxes, reading order and CTC labels identical, scores within 1e-3"""
    import bench
    from oracle import pipeline
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    pages = bench.make_pages(0, 8)
    ocr = OAROCR(det.ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 32, 256)
    got = ocr.predict(pages)
    want = pipeline.predict(oracle_nets[0], oracle_nets[1], pages, 18385, image_batch_size=32, region_batch_size=256,
                            with_margin=True)
    assert _compare(got, want, max_tie_frac=0.01) >= 300


def _border_maps():
    h, w = 96, 160
    maps = {}
    m = np.zeros((h, w), np.float32)
    m[20:40, 0:50] = 0.9            # touches x = 0
    m[50:70, w - 45:w] = 0.9        # touches x = W-1
    maps["left_right"] = m
    m = np.zeros((h, w), np.float32)
    m[0:18, 30:100] = 0.9           # touches y = 0
    m[h - 16:h, 40:130] = 0.9       # touches y = H-1
    maps["top_bottom"] = m
    m = np.zeros((h, w), np.float32)
    m[0:22, 0:60] = 0.9             # corner blobs
    m[h - 20:h, w - 70:w] = 0.9
    m[0:14, w - 40:w] = 0.85
    maps["corners"] = m
    m = np.zeros((h, w), np.float32)
    m[10:90, 0:12] = 0.9            # tall, on the left border: h >= 1.5 w
    m[5:80, w - 10:w] = 0.9
    m[8:88, 70:84] = 0.9
    maps["tall_on_border"] = m
    m = np.zeros((h, w), np.float32)
    m[30:60, :] = 0.9               # spans the full width
    maps["full_width"] = m
    m = np.zeros((h, w), np.float32)
    m[0:h, 40:70] = 0.9             # spans the full height
    m[40:56, 0:30] = 0.8            # and a neighbour touching it and the left border
    maps["full_height"] = m
    return maps


@pytest.mark.parametrize("name", sorted(_border_maps()))
def test_db_postprocess_border_blobs(ctx, name):
    """components on the image border: boxes (clipped), scores and discovery order bit-identical to the oracle"""
    from oracle import cpu
    from oar_ocr_b200 import ffi
    m = _border_maps()[name]
    for unclip, bt in ((2.0, 0.6), (1.5, 0.3)):
        got = ctx.db_postprocess(m, cfg=ffi.det_config(unclip_ratio=unclip, box_thresh=bt))[0]
        want = cpu.db_postprocess(m, m.shape[1], m.shape[0], 0.3, bt, unclip)
        assert len(want[0]) >= 1 or name == "full_width"  # (its unclipped box leaves the size limits: no box at all)
        assert len(got[0]) == len(want[0])
        assert np.array_equal(np.asarray(got[0]), np.asarray(want[0]))
        assert np.array_equal(np.asarray(got[1], np.float32), np.asarray(want[1], np.float32))
        # dest size != map size: the rescale + clip path (db_postprocess.rs:100-221)
        got2 = ctx.db_postprocess(m, src_hw=[(m.shape[0] * 2 + 1, m.shape[1] * 3 - 2)],
                                  cfg=ffi.det_config(unclip_ratio=unclip, box_thresh=bt))[0]
        want2 = cpu.db_postprocess(m, m.shape[1] * 3 - 2, m.shape[0] * 2 + 1, 0.3, bt, unclip)
        assert np.array_equal(np.asarray(got2[0]), np.asarray(want2[0]))


def test_rotate_crop_tall_boxes_turn_270(ctx):
    """h >= 1.5 w: the crop is rotated by 270 degrees (transform.rs); bytes identical to the oracle"""
    from oracle import cpu
    rng = np.random.default_rng(11)
    img = rng.integers(0, 256, (300, 260, 3), dtype=np.uint8)
    boxes = []
    for (x0, y0, bw, bh) in [(10, 10, 20, 30), (40, 5, 20, 31), (100, 20, 12, 200), (0, 0, 8, 12), (200, 90, 40, 61),
                             (150, 100, 40, 59), (230, 10, 30, 280), (5, 250, 3, 48)]:
        boxes.append(np.array([[x0, y0], [x0 + bw, y0], [x0 + bw, y0 + bh], [x0, y0 + bh]], np.float32))
    # a tilted tall quad as well
    boxes.append(np.array([[60, 40], [82, 46], [70, 120], [48, 114]], np.float32))
    got = ctx.rotate_crop(img, np.stack(boxes))
    n_rot = 0
    for g, b in zip(got, boxes):
        want = cpu.rotate_crop(img, b)
        assert (g is None) == (want is None)
        if want is None:
            continue
        assert g.shape == want.shape, (g.shape, want.shape)
        assert np.array_equal(g, want)
        w_crop = max(np.linalg.norm(b[0] - b[1]), np.linalg.norm(b[2] - b[3]))
        h_crop = max(np.linalg.norm(b[0] - b[3]), np.linalg.norm(b[1] - b[2]))
        if h_crop >= 1.5 * w_crop:
            n_rot += 1
            assert g.shape[1] >= g.shape[0]  # turned: now wider than tall
    assert n_rot >= 6


def test_more_than_4096_crops_flush(nets, oracle_nets):
    """> 4096 crops in one predict(): recognize_global flushes the pool at 4096 (ocr.rs:802-897), so the chunk
    composition changes at the flush boundary.  Pages are dense grids of tiny dark marks; the product's result must
    equal the oracle's (which restates the flush) region by region."""
    from oracle import pipeline
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    rng = np.random.default_rng(5)
    pages = []
    for p in range(5):
        img = np.full((960, 960, 3), 235, np.uint8)
        for r in range(36):
            for c in range(24):
                y, x = 10 + 26 * r, 8 + 39 * c
                wlen = 18 + int(rng.integers(0, 14))
                img[y:y + 9, x:x + wlen] = rng.integers(10, 60, (9, wlen, 1), dtype=np.uint8)
        pages.append(img)
    ocr = OAROCR(det.ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 8, 256)
    got = ocr.predict(pages)
    n_regions = sum(len(g.text_regions) for g in got)
    assert n_regions > 4096, n_regions
    want = pipeline.predict(oracle_nets[0], oracle_nets[1], pages, 18385, image_batch_size=8, region_batch_size=256,
                            with_margin=True)
    assert _compare(got, want, max_tie_frac=0.01) == n_regions


def test_crop_rec_run_equals_pipeline(nets):
    """oar_crop_rec_run (pages + boxes -> crops -> recognize_global) fed with the pipeline's own boxes reproduces the
    pipeline's labels and scores bit for bit, including the boxes whose crop fails"""
    from oar_ocr_b200 import synth
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    imgs = [synth.page(120 + i, 480) for i in range(3)] + [synth.page(130, 320)]
    ocr = OAROCR(det.ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 4, 16)
    want = ocr.predict(imgs)
    boxes, idx = [], []
    for i, res in enumerate(want):
        for r in res.text_regions:
            boxes.append(r.bounding_box.points)
            idx.append(i)
    # a degenerate box in the middle: its crop fails and it must not shift anybody else's result
    boxes.insert(5, np.zeros((4, 2), np.float32))
    idx.insert(5, idx[5])
    got = rec.crop_rec_run(imgs, np.stack(boxes), idx, 18385, region_batch_size=16)
    assert got["status"][5] == 1 and got["status"].sum() == 1
    flat = [r for res in want for r in res.text_regions]
    k = 0
    for j in range(len(boxes)):
        if j == 5:
            continue
        assert np.array_equal(got["labels"][j], flat[k].label_indices)
        assert got["scores"][j] == np.float32(flat[k].confidence)
        k += 1
    assert k == len(flat) and k >= 20


def test_rec_run_device_crops(nets):
    """oar_rec_run_ex with crops resident in HBM == oar_rec_run with host crops"""
    import ctypes as C
    from oar_ocr_b200 import synth
    det, rec = nets
    ctx = rec.ctx
    crops = [synth.crop(900 + j, 48, 200 + 8 * j) for j in range(12)]
    want = rec.rec_run(crops, 18385)
    total = sum(c.size for c in crops)
    base = ctx.device_alloc(total)
    ptrs, off = [], 0
    for c in crops:
        ctx.memcpy_h2d(base + off, np.ascontiguousarray(c))
        ptrs.append(base + off)
        off += c.size
    hs = np.array([c.shape[0] for c in crops], np.int32)
    ws = np.array([c.shape[1] for c in crops], np.int32)
    got = rec.rec_run_device((C.c_void_p * len(crops))(*ptrs), hs, ws, 18385, 64)
    assert got["T"] == want["T"]
    assert np.array_equal(got["scores"], want["scores"])
    assert all(np.array_equal(a, b) for a, b in zip(got["labels"], want["labels"]))


def test_pipeline_run_multi_two_contexts(det_blob, rec_blob):
    """oar_pipeline_run_multi with two contexts (two host threads inside the library, concurrently; both on device 0
    here -- on a multi-GPU box one per device) equals ONE un-sharded oar_pipeline_run: boxes, labels, scores, order.
    Page blocks of unequal size, more chunks than contexts, crops fetched across contexts."""
    from oar_ocr_b200 import ffi, synth
    ctxs = [ffi.Context(0), ffi.Context(0)]
    dets = [ffi.Model(c, det_blob) for c in ctxs]
    recs = [ffi.Model(c, rec_blob) for c in ctxs]
    imgs = [synth.page(140 + i, 480) for i in range(5)]
    arrs, ptrs, hs, ws = ffi._image_table(imgs)
    cfg = ffi.pipeline_config(image_batch_size=2, region_batch_size=16, rec_score_thresh=0.0, n_chars=18385)
    cfg.det = ffi.det_config(unclip_ratio=2.0)
    one = ffi.PipelineBuffers(5)
    ffi.pipeline_run(dets[0], recs[0], ptrs, hs, ws, False, cfg, one)
    for rep in range(3):
        two = ffi.PipelineBuffers(5)
        ffi.pipeline_run_multi(dets, recs, ptrs, hs, ws, cfg, two)
        n = int(one.region_off[5])
        assert n >= 30 and np.array_equal(one.region_off, two.region_off)
        assert np.array_equal(one.boxes[:n], two.boxes[:n])
        assert np.array_equal(one.det_index[:n], two.det_index[:n])
        assert np.array_equal(one.label_off[:n + 1], two.label_off[:n + 1])
        nl = int(one.label_off[n])
        assert np.array_equal(one.labels[:nl], two.labels[:nl])
        assert np.array_equal(one.scores[:n], two.scores[:n])
        assert np.array_equal(one.seq_len[:n], two.seq_len[:n])
        assert np.array_equal(one.max_wh_ratio[:n], two.max_wh_ratio[:n])


def test_db_postprocess_border_longer_than_the_staging_run(ctx):
    """dbpost.cu walks every border once into a per-block staging run of 8192 points; a longer border is walked a second
    time straight into the pool.  A comb with 150 teeth has an outer border of > 20000 points: boxes and scores must
    still be the oracle's, next to ordinary blobs and with holes inside the comb's spine."""
    from oracle import cpu
    h, w = 640, 960
    m = np.zeros((h, w), np.float32)
    m[40:60, 20:940] = 0.9                      # spine
    for k in range(150):
        m[60:130, 22 + 6 * k:25 + 6 * k] = 0.9  # teeth, 3 px wide, 3 px apart
    m[46:54, 100:120] = 0.1                     # holes in the spine
    m[46:54, 400:460] = 0.1
    m[300:340, 100:500] = 0.8                   # ordinary text-line blobs
    m[400:430, 300:900] = 0.85
    from oar_ocr_b200 import ffi
    got = ctx.db_postprocess(m, cfg=ffi.det_config(box_thresh=0.3))[0]
    want = cpu.db_postprocess(m, w, h, 0.3, 0.3, 2.0)
    assert len(want[0]) >= 3  # the comb itself is accepted at this box threshold
    assert np.array_equal(got[0], want[0]) and np.array_equal(got[1], want[1])
