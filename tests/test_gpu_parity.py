"""GPU parity tests: every C-ABI entry point of the hot path against the CPU oracle and the committed golden
vectors.  Integer / byte / index outputs are compared bit-exact; network outputs within 1e-3 (north_star).
Run with `pytest -m gpu` on a B200.
"""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu

DEFAULT_ENGINE = 2  # tcgen05 engine with fused depthwise->pointwise blocks (oar_model_set_engine)
LOGIT_TOL = 1e-3  # BASELINE.json north_star: "float logits within 1e-3"


@pytest.fixture(scope="module")
def G():
    return np.load(os.path.join(GOLDEN, "hotpath_v1.npz"))


@pytest.fixture(scope="module")
def nets(ctx, det_blob, rec_blob):
    from oar_ocr_b200 import ffi
    return ffi.Model(ctx, det_blob), ffi.Model(ctx, rec_blob)


@pytest.fixture(scope="module")
def oracle_nets(det_blob, rec_blob):
    from oracle.net import OracleNet
    return OracleNet(det_blob), OracleNet(rec_blob)


def _boxes_equal(got, want):
    gb, gs = got
    wb, ws = want
    assert gb.shape == wb.shape, (gb.shape, wb.shape)
    assert np.array_equal(gb, wb)
    assert np.array_equal(gs, ws)  # same f32 summation order -> identical bits


# ---------------------------------------------------------------- row 2
def test_normalize_golden(ctx, G):
    from oracle import cpu
    a, b = cpu.norm_coeffs(cpu.DET_SCALE, cpu.DET_MEAN, cpu.DET_STD)
    got = ctx.normalize_chw(G["norm_in"], a, b, (2, 1, 0))
    assert np.array_equal(got, G["norm_out"])


@pytest.mark.parametrize("shape", [(1, 1, 1), (3, 37, 19), (2, 64, 96), (1, 640, 640), (2, 33, 1)])
def test_normalize_vs_oracle(ctx, shape):
    from oracle import cpu
    rng = np.random.default_rng(sum(shape))
    b, h, w = shape
    img = rng.integers(0, 256, (b, h, w, 3), dtype=np.uint8)
    alpha = np.array([1.0 / 255.0, 0.5, 2.0], np.float32)
    beta = np.array([-0.485, 0.1, -1.0], np.float32)
    for src in ((0, 1, 2), (2, 1, 0)):
        got = ctx.normalize_chw(img, alpha, beta, src)
        want = np.stack([cpu.normalize(i, alpha, beta, src) for i in img])
        assert np.array_equal(got, want)


def test_normalize_empty(ctx):
    out = ctx.normalize_chw(np.zeros((0, 4, 4, 3), np.uint8), [1, 1, 1], [0, 0, 0])
    assert out.shape == (0, 3, 4, 4)


# ---------------------------------------------------------------- rows 4-9
def test_db_postprocess_golden(ctx, G):
    out = ctx.db_postprocess(G["db_pred"])
    for i in range(2):
        _boxes_equal(out[i], (G[f"db_boxes{i}"], G[f"db_scores{i}"]))
    from oar_ocr_b200 import ffi
    out = ctx.db_postprocess(G["db_pred"][0], src_hw=[(240, 384)], cfg=ffi.det_config(unclip_ratio=1.5))
    _boxes_equal(out[0], (G["db_boxes_scaled"], G["db_scores_scaled"]))


def _edge_maps():
    h, w = 96, 128
    maps = {}
    m = np.zeros((h, w), np.float32)
    maps["empty"] = m.copy()
    maps["full"] = np.full((h, w), 0.9, np.float32)
    m = np.zeros((h, w), np.float32)
    m[0:20, 0:50] = 0.9          # touches the top-left corner
    m[70:96, 90:128] = 0.95      # touches the bottom-right corner
    m[40:60, 40:90] = 0.8
    m[46:54, 50:60] = 0.1        # hole -> hole border is a candidate too
    m[30, 100] = 0.99            # single pixel
    m[10:13, 100:103] = 0.99     # 3x3
    maps["edges_holes"] = m
    m = np.zeros((h, w), np.float32)
    for k in range(12):          # diagonal staircase: 8-connectivity joins it into one component
        m[10 + 4 * k:14 + 4 * k, 10 + 6 * k:17 + 6 * k] = 0.9
    maps["staircase"] = m
    m = np.zeros((h, w), np.float32)
    m[20:70, 60:64] = 0.9        # tall thin
    m[10:12, 5:120] = 0.9        # wide thin: min_side < 3
    maps["thin"] = m
    m = np.full((h, w), 0.3, np.float32)  # exactly at the threshold: strict > keeps it background
    m[30:50, 30:90] = np.float32(0.3) + np.float32(1e-6)
    maps["at_threshold"] = m
    return maps


@pytest.mark.parametrize("name", ["empty", "full", "edges_holes", "staircase", "thin", "at_threshold"])
def test_db_postprocess_edge_cases(ctx, name):
    from oracle import cpu
    from oar_ocr_b200 import ffi
    m = _edge_maps()[name]
    for unclip, bt in ((2.0, 0.6), (1.5, 0.3)):
        got = ctx.db_postprocess(m, cfg=ffi.det_config(unclip_ratio=unclip, box_thresh=bt))[0]
        want = cpu.db_postprocess(m, m.shape[1], m.shape[0], 0.3, bt, unclip)
        _boxes_equal(got, want)


def test_db_postprocess_max_candidates_cut(ctx):
    """> max_candidates blobs: only the first N in discovery (raster) order are considered"""
    from oracle import cpu
    from oar_ocr_b200 import ffi
    m = np.zeros((200, 400), np.float32)
    for r in range(12):
        for c in range(24):
            m[8 + 16 * r:16 + 16 * r, 6 + 16 * c:18 + 16 * c] = 0.9
    for mc in (1000, 50, 7):
        got = ctx.db_postprocess(m, cfg=ffi.det_config(max_candidates=mc))[0]
        want = cpu.db_postprocess(m, 400, 200, max_candidates=mc)
        assert len(want[0]) == min(mc, 288)
        _boxes_equal(got, want)


@pytest.mark.parametrize("seed", [0, 1])
def test_db_postprocess_random_fields_batch(ctx, seed):
    """batched maps (B=3) of noisy blobs incl. regions >= 8000 px (the reference's parallel scoring branch)"""
    from oracle import cpu
    rng = np.random.default_rng(seed)
    B, h, w = 3, 256, 320
    pred = rng.random((B, h, w), dtype=np.float32) * 0.25
    for b in range(B):
        for _ in range(10):
            y, x = int(rng.integers(0, h - 40)), int(rng.integers(0, w - 120))
            bh, bw = int(rng.integers(5, 40)), int(rng.integers(10, 120))
            pred[b, y:y + bh, x:x + bw] = rng.random((bh, bw), dtype=np.float32) * 0.5 + 0.5
        pred[b, 100:200, 100:300] = np.maximum(pred[b, 100:200, 100:300], 0.7)  # 20000 px region
    got = ctx.db_postprocess(pred, src_hw=[(h, w), (2 * h, w), (h, 3 * w)])
    for b, (sh, sw) in enumerate([(h, w), (2 * h, w), (h, 3 * w)]):
        _boxes_equal(got[b], cpu.db_postprocess(pred[b], sw, sh))


# ---------------------------------------------------------------- row 10
def test_sort_quad_boxes(built_lib):
    from oracle import cpu
    from oar_ocr_b200 import ffi
    rng = np.random.default_rng(4)
    for n in (0, 1, 2, 17, 200):
        tl = np.stack([rng.integers(0, 900, n), rng.integers(0, 60, n) * 7], -1).astype(np.float32)
        boxes = tl[:, None, :] + np.array([[0, 0], [80, 0], [80, 20], [0, 20]], np.float32)[None]
        gb, go = ffi.sort_quad_boxes(boxes)
        wb, wo = cpu.sort_quad_boxes(boxes)
        assert np.array_equal(go, wo) and np.array_equal(gb, wb)


# ---------------------------------------------------------------- row 11
def test_rotate_crop_golden(ctx, G):
    crops = ctx.rotate_crop(G["crop_page"], G["crop_quads"])
    for i, c in enumerate(crops):
        want = G[f"crop{i}"]
        if want.size == 0:
            assert c is None
        else:
            assert c is not None and c.shape == want.shape
            assert np.array_equal(c, want)


def test_rotate_crop_vs_oracle_random(ctx):
    from oracle import cpu
    from oar_ocr_b200 import synth
    page = synth.page(21, 480)
    rng = np.random.default_rng(2)
    quads = []
    for _ in range(40):
        cx, cy = rng.uniform(-20, 500), rng.uniform(-20, 500)
        w, h = rng.uniform(4, 300), rng.uniform(4, 80)
        if rng.random() < 0.2:
            w, h = h, w * 1.2
        a = np.deg2rad(rng.uniform(-12, 12))
        base = np.array([[-w / 2, -h / 2], [w / 2, -h / 2], [w / 2, h / 2], [-w / 2, h / 2]])
        q = base @ np.array([[np.cos(a), np.sin(a)], [-np.sin(a), np.cos(a)]]) + [cx, cy]
        quads.append(np.round(q) if rng.random() < 0.7 else q)  # DB boxes are integer valued
    quads = np.array(quads, np.float32)
    got = ctx.rotate_crop(page, quads)
    n_ok = 0
    for q, g in zip(quads, got):
        want = cpu.rotate_crop(page, q)
        if want is None:
            assert g is None
            continue
        n_ok += 1
        assert g is not None and g.shape == want.shape
        assert np.array_equal(g, want)
    assert n_ok >= 25


# ---------------------------------------------------------------- row 13
def test_crnn_preprocess_golden(ctx, G):
    got = ctx.crnn_preprocess([G["crnn_in0"], G["crnn_in1"], G["crnn_in2"]])
    assert got.shape == G["crnn_out"].shape
    assert np.array_equal(got, G["crnn_out"])


def test_crnn_preprocess_vs_oracle(ctx):
    from oracle import cpu
    rng = np.random.default_rng(8)
    crops = [rng.integers(0, 256, (int(rng.integers(8, 90)), int(rng.integers(8, 700)), 3), dtype=np.uint8)
             for _ in range(9)]
    crops.append(rng.integers(0, 256, (10, 900, 3), dtype=np.uint8))  # ratio 90 -> tensor_w capped at 3200
    got = ctx.crnn_preprocess(crops)
    want = cpu.crnn_preprocess(crops)
    assert got.shape == want.shape and got.shape[3] == 3200
    assert np.array_equal(got, want)
    assert ctx.crnn_preprocess([]).shape == (0, 0, 0, 0)


# ---------------------------------------------------------------- rows 15-16
def test_ctc_decode_golden(ctx, G):
    r = ctx.ctc_decode(G["ctc_pred"], 36)
    assert np.array_equal(r["idx"], G["ctc_idx"]) and np.array_equal(r["prob"], G["ctc_prob"])
    assert np.array_equal(r["scores"], G["ctc_scores"])
    for i in range(3):
        assert np.array_equal(r["labels"][i], G[f"ctc_labels{i}"])
        assert np.array_equal(r["cols"][i], G[f"ctc_cols{i}"])


def test_ctc_reference_vectors(ctx):
    """decode.rs:692-758 through the CUDA path"""
    winners = [[(0, 0.9), (1, 0.8), (1, 0.7), (0, 0.6), (1, 0.5), (2, 0.4), (2, 0.3)],
               [(3, 0.95), (3, 0.85), (4, 0.75), (3, 0.65), (0, 0.55), (2, 0.45), (0, 0.35)]]
    logits = np.full((2, 7, 5), -10.0, np.float32)
    for b, seq in enumerate(winners):
        for t, (k, p) in enumerate(seq):
            logits[b, t, k] = p
    r = ctx.ctc_decode(logits, 4)
    chars = ["", "a", "b", "c"]
    assert ["".join(chars[k] for k in lab) for lab in r["labels"]] == ["aab", "ccb"]
    f = np.float32
    assert r["scores"][0] == (f(0.8) + f(0.5) + f(0.4)) / f(3.0)
    assert r["scores"][1] == (f(0.95) + f(0.65) + f(0.45)) / f(3.0)
    assert [c.tolist() for c in r["cols"]] == [[1, 4, 5], [0, 3, 5]]
    tied = np.array([[[1, 2, 5, 9, 4, 8, 9, 0]]], np.float32)
    assert ctx.ctc_decode(tied, 8)["idx"][0, 0] == 6
    assert ctx.ctc_decode(np.zeros((2, 0, 5), np.float32), 5)["labels"] == []


def test_ctc_full_vocab_vs_oracle(ctx):
    from oracle import cpu
    rng = np.random.default_rng(1)
    pred = rng.random((4, 40, 18385), dtype=np.float32)
    pred[:, ::3, 0] = 2.0
    r = ctx.ctc_decode(pred, 18385)
    idx, prob = cpu.ctc_argmax(pred)
    labels, scores, cols, _ = cpu.ctc_decode(idx, prob, 18385)
    assert np.array_equal(r["idx"], idx) and np.array_equal(r["prob"], prob)
    assert np.array_equal(r["scores"], scores)
    assert all(np.array_equal(a, b) for a, b in zip(r["labels"], labels))


# ---------------------------------------------------------------- rows 3 / 14: the networks
@pytest.mark.parametrize("engine", [0, 1, 2])
def test_det_net_golden_and_oracle(nets, oracle_nets, G, engine):
    from oracle import cpu
    det, _ = nets
    det.set_engine(engine)
    x = cpu.det_normalize(G["det_in"])[None]
    got = det.infer(x)
    assert got.shape == (1, 1, 64, 96)
    assert np.abs(got[0, 0] - G["det_pred"]).max() <= LOGIT_TOL
    from oar_ocr_b200 import synth
    imgs = [synth.page(30 + i, 160) for i in range(2)]
    x = np.stack([cpu.det_normalize(i) for i in imgs])
    got = det.infer(x)
    want = oracle_nets[0].forward(x)
    assert np.abs(got - want).max() <= LOGIT_TOL
    det.set_engine(DEFAULT_ENGINE)


@pytest.mark.parametrize("engine", [0, 1, 2])
def test_rec_net_golden_and_oracle(nets, oracle_nets, G, engine):
    from oracle import cpu
    _, rec = nets
    rec.set_engine(engine)
    crops = [G["rec_in0"], G["rec_in1"]]
    x = cpu.crnn_preprocess(crops)
    probs = rec.infer(x)
    assert probs.shape[0] == 2 and probs.shape[2] == 18385
    want = oracle_nets[1].forward(x)
    assert probs.shape == want.shape
    assert np.abs(probs - want).max() <= LOGIT_TOL
    idx, prob = cpu.ctc_argmax(probs)
    assert np.array_equal(idx, G["rec_idx"])
    assert np.abs(prob - G["rec_prob"]).max() <= LOGIT_TOL
    r = rec.rec_run(crops, 18385)  # fused head: argmax + softmax-max without materialising [B,T,V]
    lab, sc, cols, T = cpu.ctc_decode(G["rec_idx"], G["rec_prob"], 18385)
    assert r["T"] == T
    assert all(np.array_equal(a, b) for a, b in zip(r["labels"], lab))
    assert np.abs(r["scores"] - sc).max() <= LOGIT_TOL
    rec.set_engine(DEFAULT_ENGINE)


# ---------------------------------------------------------------- seam 2: adapters and the whole path
def test_det_run_vs_oracle_mixed_shapes(nets, oracle_nets):
    """TextDetectionAdapter::execute on a batch with two shape groups and one image that needs resizing"""
    from oracle import pipeline
    from oar_ocr_b200 import synth, ffi
    det, _ = nets
    imgs = [synth.page(40, 320), synth.page(41, 256), synth.page(42, 320), synth.page(43, 1200)[:600]]
    cfg = ffi.det_config(limit_side_len=640)
    got = det.det_run(imgs, cfg)
    want = pipeline.det_forward(oracle_nets[0], imgs, limit=640)
    assert sum(len(w[0]) for w in want) >= 8
    for g, w in zip(got, want):
        _boxes_equal((g[0], np.zeros(0)), (w[0], np.zeros(0)))  # boxes bit-exact
        assert len(w[1]) == 0 or np.abs(g[1] - w[1]).max() <= LOGIT_TOL  # scores are means of net outputs
    assert det.det_run([], cfg) == []


def test_pipeline_golden(nets, G):
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    ocr = OAROCR(det.ctx, det, rec, None, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 8, 4)
    ocr.chars = [""] * 18385
    res = ocr.predict([G["pipe_page"]])[0]
    assert len(res.text_regions) == len(G["pipe_boxes"])
    off = G["pipe_label_off"]
    for i, r in enumerate(res.text_regions):
        assert np.array_equal(r.bounding_box.points, G["pipe_boxes"][i])
        assert np.array_equal(r.label_indices, G["pipe_labels"][off[i]:off[i + 1]])
        assert abs(r.confidence - G["pipe_scores"][i]) <= LOGIT_TOL


@pytest.mark.parametrize("engine", [0, 1, 2])
def test_pipeline_vs_oracle_batch(nets, oracle_nets, engine):
    """OAROCR::predict on 5 pages of two sizes, image_batch_size 2, region_batch_size 8: boxes and CTC label
    sequences identical to the CPU oracle, confidences within 1e-3"""
    from oracle import pipeline
    from oar_ocr_b200 import synth
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    det.set_engine(engine)
    rec.set_engine(engine)
    imgs = [synth.page(50, 480), synth.page(51, 480), synth.page(52, 320), synth.page(53, 480), synth.page(54, 320)]
    ocr = OAROCR(det.ctx, det, rec, [""] * 18385, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 2, 8)
    got = ocr.predict(imgs)
    want = pipeline.predict(oracle_nets[0], oracle_nets[1], imgs, 18385, image_batch_size=2, region_batch_size=8)
    total = 0
    for g, w in zip(got, want):
        assert len(g.text_regions) == len(w)
        for r, o in zip(g.text_regions, w):
            assert np.array_equal(r.bounding_box.points, o["box"])
            assert r.detection_index == o["det_index"]
            assert np.array_equal(r.label_indices, o["labels"])
            assert abs(r.confidence - o["score"]) <= LOGIT_TOL
            total += 1
    assert total >= 20
    assert ocr.last_timing["ms_total"] > 0
    det.set_engine(DEFAULT_ENGINE)
    rec.set_engine(DEFAULT_ENGINE)


def test_pipeline_errors(nets):
    from oar_ocr_b200.ocr import OAROCR, OCRError, TextDetectionConfig, TextRecognitionConfig
    det, rec = nets
    ocr = OAROCR(det.ctx, det, rec, [""] * 18385, TextDetectionConfig(), TextRecognitionConfig(), 8, 64)
    with pytest.raises(OCRError) as e:
        ocr.predict([])
    assert "non-empty slice" in str(e.value)  # ocr.rs:525-532
    blank = np.full((64, 64, 3), 240, np.uint8)
    assert ocr.predict([blank])[0].text_regions == []


def test_launch_counter_moves(nets, G):
    from oar_ocr_b200 import ffi
    det, _ = nets
    n0 = ffi.launch_count()
    det.det_run([G["det_in"]])
    assert ffi.launch_count() - n0 > 50


def test_pipeline_word_boxes_vs_oracle(nets, oracle_nets):
    """return_word_box (ocr.rs:241, 860-868): the pipeline hands back CTC columns, T, wh_ratio and the batch's max
    ratio per region; the per-character boxes computed from them equal the oracle's bit for bit."""
    from oracle import pipeline
    from oar_ocr_b200 import models, synth
    from oar_ocr_b200.ocr import OAROCR, TextDetectionConfig, TextRecognitionConfig, character_list
    det, rec = nets
    chars = character_list(models.synthetic_dict())
    imgs = [synth.page(60, 480), synth.page(61, 320), synth.page(62, 480)]
    ocr = OAROCR(det.ctx, det, rec, chars, TextDetectionConfig(unclip_ratio=2.0), TextRecognitionConfig(), 2, 5)
    ocr.return_word_box = True
    got = ocr.predict(imgs)
    want = pipeline.predict(oracle_nets[0], oracle_nets[1], imgs, len(chars), image_batch_size=2, region_batch_size=5,
                            chars=chars)
    n_boxes = 0
    for g, w in zip(got, want):
        assert len(g.text_regions) == len(w)
        for r, o in zip(g.text_regions, w):
            assert np.array_equal(r.label_indices, o["labels"])
            wb = r.word_boxes or []
            assert len(wb) == len(o["word_boxes"]) == len(o["labels"])
            for b, row in zip(wb, o["word_boxes"]):
                assert (b.points[0, 0], b.points[0, 1], b.points[1, 0], b.points[2, 1]) == tuple(row)
                n_boxes += 1
    assert n_boxes >= 100


# ---------------------------------------------------------------- fused block kernel on shapes the two nets do not hit
def _block_graph(spec, seed):
    """stem conv 3 -> c0, then [depthwise k, stride -> (SE) -> 1x1 conv -> hardswish] blocks, then a 1x1 conv to 1 channel"""
    from oar_ocr_b200 import models
    g = models.GraphBuilder(models.KIND_DET, seed)
    g.base_gain = 0.8
    x = g.conv(0, spec["c0"], (3, 3), spec.get("stem_stride", (1, 1)), act=models.ACT_NONE)
    for (k, cout, stride, use_se) in spec["blocks"]:
        x = models._lcnet_block(g, x, k, cout, stride, use_se, plant=0)
    for cout in spec.get("pw", []):
        x = g.conv(x, cout, (1, 1), act=models.ACT_HSWISH)
    g.conv(x, 1, (1, 1), act=models.ACT_NONE)
    return g.serialize()


_BLOCK_SPECS = [
    # odd image size, C = 16 block, stride-2 3x3, 5x5 with a 96-channel tail (3 k-blocks)
    dict(hw=(37, 53), c0=16, blocks=[(3, 32, (1, 1), False), (3, 48, (2, 2), False), (5, 96, (1, 1), False)]),
    # rec-like strides, 240 channels (K tail of 16), N = 480 (two N tiles)
    dict(hw=(24, 50), c0=64, blocks=[(3, 240, (1, 2), False), (5, 240, (1, 1), False), (5, 480, (2, 1), False)]),
    # N = 200 (not a multiple of 16 or 32), 5x5 stride 2 on an odd size
    dict(hw=(31, 31), c0=96, blocks=[(5, 200, (2, 2), False), (3, 200, (1, 1), False)]),
    # squeeze-excite blocks: depthwise with per-tile sums -> pool -> FCs -> 1x1 conv with the scale folded into A
    dict(hw=(20, 44), c0=128, blocks=[(5, 256, (1, 1), True), (3, 256, (2, 1), True)]),
    # plain 1x1 convs: C = 12 (TMA rows of 48 bytes), wide N, a 3-row image (tile taller than the image)
    dict(hw=(3, 80), c0=12, blocks=[], pw=[96, 480, 24]),
]


@pytest.mark.parametrize("spec_i", range(len(_BLOCK_SPECS)))
def test_fused_block_kernel_odd_shapes(ctx, spec_i):
    """engine 2 (persistent fused kernel: tile choice, halo boxes, channel / class tails, clipped stores) against the
    CPU oracle and the fp32 SIMT engine on small graphs with shapes the PP-OCRv5 nets never produce"""
    from oar_ocr_b200 import ffi
    from oracle.net import OracleNet
    spec = _BLOCK_SPECS[spec_i]
    blob = _block_graph(spec, 100 + spec_i)
    rng = np.random.default_rng(spec_i)
    h, w = spec["hw"]
    x = rng.standard_normal((3, 3, h, w)).astype(np.float32)
    want = OracleNet(blob).forward(x)
    m = ffi.Model(ctx, blob)
    scale = max(1.0, float(np.abs(want).max()))
    outs = {}
    for eng in (0, 1, 2):
        m.set_engine(eng)
        outs[eng] = m.infer(x)
        assert outs[eng].shape == want.shape
        assert np.abs(outs[eng] - want).max() <= 2e-4 * scale, (eng, np.abs(outs[eng] - want).max(), scale)
    # same batch again: bit-stable
    assert np.array_equal(m.infer(x), outs[2])
    m.close()
