"""Text-line orientation stage: OAROCR::classify_line_orientations (src/oarocr/ocr.rs:755-792),
TextLineOrientationAdapter::execute (domain/adapters/text_line_orientation_adapter.rs:63-121), PPLCNetModel
(models/classification/pp_lcnet.rs:139-300) and Topk (utils/topk.rs).

CPU part: the oracle restatement against the reference's own Topk vectors and numpy, the synthetic classifier's
conditioning, the host mirror.  GPU part (`-m gpu`): the CUDA path through the C ABI against the oracle -- rotate180 and
class ids bit-exact, probabilities within 1e-3, the whole pipeline with the classifier switched on."""
import numpy as np
import pytest

LOGIT_TOL = 1e-3  # north_star: float outputs within 1e-3 of the fp32 CPU run
PAGES = [(50, 480), (51, 480), (52, 320), (53, 480), (54, 320)]


@pytest.fixture(scope="module")
def cls_blob():
    from oar_ocr_b200 import models
    return models.get_blob("cls")


@pytest.fixture(scope="module")
def oracle_cls(cls_blob):
    from oracle.net import OracleNet
    return OracleNet(cls_blob)


def _pages():
    from oar_ocr_b200 import synth
    return [synth.page(s, n) for s, n in PAGES]


def _mixed_crops():
    """crops of many sizes: recogniser-style lines, page crops (tall, wide, tiny)"""
    from oar_ocr_b200 import synth
    rng = np.random.default_rng(11)
    crops = [synth.crop(j) for j in range(6)]
    page = synth.page(3, 480)
    for (y, x, h, w) in [(40, 40, 30, 200), (100, 60, 17, 333), (200, 10, 64, 48), (5, 5, 3, 7), (300, 100, 80, 160),
                         (0, 0, 1, 1), (50, 50, 96, 31)]:
        crops.append(np.ascontiguousarray(page[y:y + h, x:x + w]))
    crops.append(rng.integers(0, 256, size=(23, 41, 3), dtype=np.uint8))
    return crops


# --------------------------------------------------------------------------------------------------------------
# CPU: oracle pinned by the reference's vectors; host mirror
# --------------------------------------------------------------------------------------------------------------
def test_topk_reference_vectors():
    """utils/topk.rs:294-373"""
    from oracle import cpu
    idx, sc = cpu.topk([0.1, 0.8, 0.1], 2)  # test_topk_without_class_names, test_process_single
    assert idx.tolist() == [1, 0] and sc.tolist() == pytest.approx([0.8, 0.1])
    assert cpu.topk([0.7, 0.2, 0.1], 2)[0].tolist() == [0, 1]
    assert len(cpu.topk([0.1, 0.8], 5)[0]) == 2  # test_topk_k_larger_than_classes
    with pytest.raises(ValueError):  # test_topk_invalid_k
        cpu.topk([0.1, 0.8, 0.1], 0)
    # stable sort: equal scores keep index order
    assert cpu.topk([0.5, 0.5], 1)[0].tolist() == [0]
    assert cpu.topk([0.2, 0.3, 0.3, 0.1], 3)[0].tolist() == [1, 2, 0]


def test_mirror_topk_equals_oracle():
    from oracle import cpu
    from oar_ocr_b200.ocr import OCRError, topk_indices
    rng = np.random.default_rng(5)
    for _ in range(50):
        p = rng.integers(0, 4, size=int(rng.integers(1, 7))).astype(np.float32) / 4
        k = int(rng.integers(1, 8))
        assert topk_indices(p, k) == cpu.topk(p, k)[0].tolist()
    with pytest.raises(OCRError):
        topk_indices(np.array([0.1, 0.9], np.float32), 0)


@pytest.mark.parametrize("shape", [(1, 1), (1, 2), (3, 5), (48, 320), (7, 2)])
def test_oracle_rotate180(shape):
    from oracle import cpu
    img = np.random.default_rng(shape[0]).integers(0, 256, size=shape + (3,), dtype=np.uint8)
    out = cpu.rotate180(img)
    assert np.array_equal(out, img[::-1, ::-1])  # image::imageops::rotate180: out(w-1-x, h-1-y) = in(x, y)
    assert np.array_equal(cpu.rotate180(out), img)


def test_oracle_cls_preprocess():
    """direct resize to (w, h) = (160, 80), scale 1/255, ImageNet mean/std, RGB order, CHW (pp_lcnet.rs:168-196,
    400-412); an image that already has the input size is only normalised"""
    from oracle import cpu
    img = np.random.default_rng(1).integers(0, 256, size=(80, 160, 3), dtype=np.uint8)
    x = cpu.cls_preprocess([img, img[:40, :50]])
    assert x.shape == (2, 3, 80, 160) and x.dtype == np.float32
    mean = np.array([0.485, 0.456, 0.406], np.float32)
    std = np.array([0.229, 0.224, 0.225], np.float32)
    alpha = (np.float32(1.0) / np.float32(255.0)) / std
    beta = -mean / std
    want = (img.astype(np.float32) * alpha + beta).transpose(2, 0, 1)
    assert np.array_equal(x[0], want)
    assert cpu.cls_preprocess([np.zeros((0, 5, 3), np.uint8)]).shape == (0, 3, 80, 160)  # filter_map drops it
    assert cpu.cls_preprocess([img], (192, 48)).shape == (1, 3, 192, 48)


def test_cls_blob_is_well_formed(cls_blob):
    from oar_ocr_b200 import models
    from oracle.net import parse
    kind, n_tensors, ops, weights = parse(cls_blob)
    assert kind == models.KIND_CLS == 2
    assert sum(o["type"] == models.OP_DWCONV for o in ops) == 13  # PP-LCNet x1.0: 13 depthwise-separable blocks
    assert sum(o["type"] == models.OP_SE for o in ops) == 2
    assert ops[-1]["type"] == models.OP_CTC_HEAD and tuple(ops[-1]["p"][:2]) == (1280, 2)
    assert ops[-3]["type"] == models.OP_AVGPOOL and tuple(ops[-3]["p"][:2]) == (0, 0)  # global pool


def test_oracle_pipeline_with_line_orientation(det_blob, rec_blob, oracle_cls):
    """classify_line_orientations inside predict: every region gets an angle, class-1 crops are rotated (their text
    changes), boxes / order / wh_ratio-driven batching stay as without the classifier.  Also the conditioning the GPU
    parity test relies on: both classes occur and no crop sits within 0.05 of the decision boundary."""
    from oracle import cpu, pipeline
    from oracle.net import OracleNet
    det, rec = OracleNet(det_blob), OracleNet(rec_blob)
    imgs = _pages()
    base = pipeline.predict(det, rec, imgs, 18385, image_batch_size=2, region_batch_size=8)
    want = pipeline.predict(det, rec, imgs, 18385, image_batch_size=2, region_batch_size=8, cls_net=oracle_cls)
    n = n180 = changed = 0
    crops = []
    for img, b, w in zip(imgs, base, want):
        assert len(b) == len(w)
        for rb, rw in zip(b, w):
            assert rb["angle"] is None and rw["angle"] in (0.0, 180.0)
            assert np.array_equal(rb["box"], rw["box"]) and rb["det_index"] == rw["det_index"] and rb["T"] == rw["T"]
            n += 1
            n180 += rw["angle"] == 180.0
            same = np.array_equal(rb["labels"], rw["labels"])
            if rw["angle"] == 0.0:
                assert same and rb["score"] == rw["score"]
            changed += not same
            crops.append(cpu.rotate_crop(img, rw["box"]))
    assert n >= 20 and 5 <= n180 <= n - 5 and changed >= 5
    tops, probs = pipeline.cls_forward(oracle_cls, crops)
    assert [float(t[0][0]) * 180.0 for t in tops] == [r["angle"] for w in want for r in w]
    assert np.abs(probs[:, 1] - probs[:, 0]).min() > 0.05
    assert np.allclose(probs.sum(1), 1.0, atol=1e-6)


def test_mirror_builders_and_validation():
    from oar_ocr_b200.ocr import (OAROCRBuilder, OCRError, TextLineOrientationConfig,
                                  TextLineOrientationPredictorBuilder)
    b = TextLineOrientationPredictorBuilder()
    assert b._input_shape == (192, 48) and b._config.topk == 2 and b._config.score_threshold == 0.5
    assert b.topk(1).score_threshold(0.9).input_shape((80, 160))._input_shape == (80, 160)
    with pytest.raises(OCRError):
        TextLineOrientationConfig(topk=0).validate()
    with pytest.raises(OCRError):
        TextLineOrientationConfig(score_threshold=1.5).validate()
    ob = OAROCRBuilder("synthetic", "synthetic").with_text_line_orientation_classification("synthetic")
    assert ob._line_ori == "synthetic"  # ocr.rs:1108-1124


# --------------------------------------------------------------------------------------------------------------
# GPU: the CUDA path through the C ABI against the oracle
# --------------------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def cls_model(ctx, cls_blob):
    from oar_ocr_b200 import ffi
    return ffi.Model(ctx, cls_blob)


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(1, 1), (1, 2), (3, 5), (48, 320), (7, 2), (33, 1001)])
def test_rotate180_vs_oracle(ctx, shape):
    from oracle import cpu
    img = np.random.default_rng(shape[1]).integers(0, 256, size=shape + (3,), dtype=np.uint8)
    assert np.array_equal(ctx.rotate180(img), cpu.rotate180(img))


@pytest.mark.gpu
@pytest.mark.parametrize("engine", [0, 1, 2])
def test_cls_net_vs_oracle(cls_model, oracle_cls, engine):
    """seam 1 (OrtInfer::infer) on the classifier graph: [B,3,80,160] -> [B,2] probabilities within 1e-3, and on the
    stand-alone predictor's (192, 48) input"""
    from oracle import cpu
    cls_model.set_engine(engine)
    for shape in ((80, 160), (192, 48)):
        x = cpu.cls_preprocess(_mixed_crops()[:5], shape)
        got = cls_model.infer(x)
        want = oracle_cls.forward(x).reshape(len(x), -1)
        assert got.shape == want.shape == (5, 2)
        assert np.abs(got - want).max() <= LOGIT_TOL
    cls_model.set_engine(2)


@pytest.mark.gpu
def test_cls_run_vs_oracle(cls_model, oracle_cls):
    """TextLineOrientationAdapter::execute on crops of many sizes: same classes, probabilities within 1e-3"""
    from oracle import pipeline
    from oar_ocr_b200 import ffi
    crops = _mixed_crops()
    got = cls_model.cls_run(crops)
    tops, probs = pipeline.cls_forward(oracle_cls, crops)
    assert got["probs"].shape == probs.shape == (len(crops), 2)
    assert np.abs(got["probs"] - probs).max() <= LOGIT_TOL
    for i, (ids, sc) in enumerate(tops):
        if abs(probs[i, 0] - probs[i, 1]) > 4 * LOGIT_TOL:  # away from the decision boundary: identical class
            assert got["class_ids"][i] == ids[0]
            assert abs(got["scores"][i] - sc[0]) <= LOGIT_TOL
        assert got["scores"][i] == got["probs"][i].max()
    assert len(set(got["class_ids"].tolist())) == 2
    # more crops than one classifier chunk (256): chunking does not change a crop's result
    many = cls_model.cls_run([crops[i % len(crops)] for i in range(300)], want_probs=True)
    for i in range(300):
        assert np.abs(many["probs"][i] - got["probs"][i % len(crops)]).max() <= 1e-6
    # empty input / empty crop
    assert len(cls_model.cls_run([])["class_ids"]) == 0
    with pytest.raises(ffi.OCRError):
        cls_model.cls_run([np.zeros((0, 4, 3), np.uint8)])


@pytest.mark.gpu
def test_predictor_mirror(ctx, cls_blob, oracle_cls):
    from oracle import pipeline
    from oar_ocr_b200.ocr import OCRError, TextLineOrientationPredictor
    pred = TextLineOrientationPredictor.builder().input_shape((80, 160)).build(cls_blob)
    crops = _mixed_crops()[:6]
    res = pred.predict(crops)
    tops, _ = pipeline.cls_forward(oracle_cls, crops, topk=2)
    assert len(res.orientations) == 6
    for row, (ids, sc) in zip(res.orientations, tops):
        assert [c.class_id for c in row] == ids.tolist()
        assert [c.label for c in row] == [str(int(i) * 180) for i in ids]
        assert np.abs(np.array([c.score for c in row]) - sc).max() <= LOGIT_TOL
    with pytest.raises(OCRError) as e:
        pred.predict([])
    assert "No images provided" in str(e.value)


@pytest.mark.gpu
@pytest.mark.parametrize("engine", [0, 2])
def test_pipeline_line_orientation_vs_oracle(ctx, det_blob, rec_blob, cls_blob, oracle_cls, engine):
    """OAROCR::predict with with_text_line_orientation_classification: angles, boxes and CTC label sequences identical
    to the oracle (class-1 crops are rotated in HBM before recognition), confidences within 1e-3"""
    from oracle import pipeline
    from oracle.net import OracleNet
    from oar_ocr_b200 import ffi
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec, cls = ffi.Model(ctx, det_blob), ffi.Model(ctx, rec_blob), ffi.Model(ctx, cls_blob)
    for m in (det, rec, cls):
        m.set_engine(engine)
    imgs = _pages()
    ocr = OAROCR(ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 2, 8)
    plain = ocr.predict(imgs)
    ocr.cls = cls
    got = ocr.predict(imgs)
    want = pipeline.predict(OracleNet(det_blob), OracleNet(rec_blob), imgs, 18385, image_batch_size=2,
                            region_batch_size=8, cls_net=oracle_cls)
    total = n180 = changed = 0
    for g, p, w in zip(got, plain, want):
        assert len(g.text_regions) == len(w) == len(p.text_regions)
        for r, q, o in zip(g.text_regions, p.text_regions, w):
            assert q.orientation_angle is None
            assert r.orientation_angle == o["angle"]
            assert np.array_equal(r.bounding_box.points, o["box"])
            assert r.detection_index == o["det_index"]
            assert np.array_equal(r.label_indices, o["labels"])
            assert abs(r.confidence - o["score"]) <= LOGIT_TOL
            total += 1
            n180 += r.orientation_angle == 180.0
            changed += not np.array_equal(r.label_indices, q.label_indices)
    assert total >= 20 and n180 >= 5 and changed >= 5
    assert ocr.last_timing["ms_cls"] > 0
