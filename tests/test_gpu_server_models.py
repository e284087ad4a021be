"""SURVEY.md 8f item 4 (server-size models) and the network half of item 1 (layout detector) on the CUDA engine:
the HGNetV2 feature extractor (zero pad, 2x2 max pool, dense 3x3 stacks writing concat slices, 1x1 + depthwise 5x5
light layers, identity residuals), the server-size recogniser built on it, and the layout detector's backbone +
hybrid encoder (exact GELU, attention at head width 32 with 2-D sine positions on q / k, token rows) -- each against
the CPU oracle's interpreter of the same layer list (oracle/net.py), through the C ABI (oar_infer_f32 / oar_rec_run)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rel(got, want):
    return float(np.abs(got - want).max() / max(float(np.abs(want).max()), 1.0))


def test_new_ops_on_device(ctx):
    """OP_PAD / OP_MAXPOOL alone: bit-exact against the oracle's definition"""
    from oar_ocr_b200 import ffi, models
    from oracle.net import OracleNet
    g = models.GraphBuilder(models.KIND_FEAT, 0)
    x0 = g.conv(0, 8, (1, 1), act=models.ACT_NONE)  # 3 -> 8 channels (the engine's maps are /4 channels)
    g.maxpool(g.pad(x0, 1, 2, 1, 3), (2, 2), (1, 1))
    blob = g.serialize()
    x = np.random.default_rng(0).standard_normal((2, 3, 9, 11)).astype(np.float32)
    want = OracleNet(blob).forward(x)
    got = ffi.Model(ctx, blob).infer(x, out_cap=want.size + 64)
    assert got.shape == want.shape
    assert _rel(got, want) <= 1e-5


@pytest.mark.parametrize("engine", [0, 2])
@pytest.mark.parametrize("stage", [0, 3])
def test_hgnetv2_l_features(ctx, stage, engine):
    """HGNetV2-L stage outputs (strides 4 and 32, 128 and 2048 channels) equal the oracle to 1e-4 of their range"""
    from oar_ocr_b200 import ffi, models
    from oracle.net import OracleNet
    blob = models.build_hgnetv2_l(return_idx=(stage,))
    x = np.random.default_rng(1).standard_normal((2, 3, 96, 128)).astype(np.float32)
    want = OracleNet(blob).forward(x)
    m = ffi.Model(ctx, blob)
    assert m.kind == ffi.KIND_FEAT
    m.set_engine(engine)
    got = m.infer(x, out_cap=want.size + 64)
    assert got.shape == want.shape == (2, (128, 512, 1024, 2048)[stage], 96 // (4, 8, 16, 32)[stage], 128 // (4, 8, 16, 32)[stage])
    assert _rel(got, want) <= 1e-4


def test_server_recogniser_matches_oracle(ctx):
    """PP-OCRv5_server_rec shaped graph through oar_rec_run: CTC label sequences identical to the oracle, confidences
    within 1e-3 (crops of different widths in one batch, vocabulary 18385)"""
    from oar_ocr_b200 import ffi, models, synth
    from oracle import pipeline
    from oracle.net import OracleNet
    blob = models.build_rec_server()
    rec = ffi.Model(ctx, blob)
    assert rec.kind == ffi.KIND_REC
    crops = [synth.crop(40 + j, 48, 96 + 32 * j) for j in range(6)]
    got = rec.rec_run(crops, 18385)
    want = pipeline.rec_forward(OracleNet(blob), crops, 18385)
    assert got["T"] == want["T"]
    assert all(np.array_equal(a, b) for a, b in zip(got["labels"], want["labels"]))
    assert np.abs(got["scores"] - want["scores"]).max() <= 1e-3
    # and the probabilities themselves through seam 1
    x = np.random.default_rng(2).standard_normal((2, 3, 48, 160)).astype(np.float32)
    small = models.build_rec_server(vocab=97)
    p_want = OracleNet(small).forward(x)
    p_got = ffi.Model(ctx, small).infer(x, vocab_hint=97)
    assert p_got.shape == p_want.shape == (2, 20, 97)
    assert np.abs(p_got - p_want).max() <= 1e-3


@pytest.mark.parametrize("engine", [0, 2])
def test_layout_encoder_memory(ctx, engine):
    """RT-DETR-L backbone + hybrid encoder + decoder-input projections as one layer list on the CUDA engine: the
    decoder memory [B, N, 256] equals the oracle's (oracle/net.py interpreter == oracle/rtdetr.py, test_rtdetr_oracle)"""
    from oar_ocr_b200 import ffi, models
    from oracle.net import OracleNet
    shapes = [(24, 32), (12, 16), (6, 8)]
    blob = models.build_layout_encoder(models.layout_weights(42), seed=42, shapes_hw=shapes)
    x = np.random.default_rng(3).standard_normal((2, 3, 192, 256)).astype(np.float32)
    want = OracleNet(blob).forward(x)  # [B, 256, 1, N]
    m = ffi.Model(ctx, blob)
    m.set_engine(engine)
    got = m.infer(x, out_cap=want.size + 64)
    assert got.shape == want.shape == (2, 256, 1, sum(h * w for h, w in shapes))
    assert _rel(got, want) <= 2e-4


def _layout_models(ctx, in_hw):
    from oar_ocr_b200 import ffi, models
    w = models.layout_weights(42)
    shapes = [(in_hw[0] // s, in_hw[1] // s) for s in (8, 16, 32)]
    enc = ffi.Model(ctx, models.build_layout_encoder(w, seed=42, shapes_hw=shapes))
    head = ffi.Model(ctx, models.build_layout_head(w))
    return enc, head


def test_layout_rows_match_oracle(ctx):
    """BASELINE.json configs[4] at a reduced input size, whole network on the device: pages -> CatmullRom resize ->
    RT-DETR-L -> the exported model's rows, against oracle/rtdetr.py fed by the oracle's preprocessing.  Rows are
    score-ordered; near-equal scores may swap places, so every oracle row is matched to a product row of the same class
    (score within 2e-3, box corners within 0.25 px of a page that is hundreds of pixels wide)."""
    from oar_ocr_b200 import ffi, synth
    from oracle import cpu
    from oracle.rtdetr import RTDetrL
    in_hw = (256, 256)
    pages = [synth.page(400, 480), synth.page(401, 320), synth.page(402, 256)]
    enc, head = _layout_models(ctx, in_hw)
    got = ffi.layout_rows(enc, head, pages, in_hw)
    x, _ = cpu.layout_preprocess(pages, in_hw)
    want = RTDetrL(42).rows(x, [(p.shape[1], p.shape[0]) for p in pages])
    assert got.shape == want.shape == (3, 300, 6)
    for b in range(3):
        g, w = got[b], want[b]
        assert np.all(np.diff(g[:, 1]) <= 1e-7)  # descending scores
        assert abs(g[0, 1] - w[0, 1]) <= 2e-3
        used = np.zeros(300, bool)
        for r in range(100):  # the 100 best oracle rows
            cand = np.where((g[:, 0] == w[r, 0]) & (np.abs(g[:, 1] - w[r, 1]) <= 2e-3) & ~used)[0]
            d = [np.abs(g[c, 2:] - w[r, 2:]).max() for c in cand]
            assert len(d) and min(d) <= 0.25, (b, r, w[r], len(d), min(d) if d else None)
            used[cand[int(np.argmin(d))]] = True


def test_layout_run_end_to_end(ctx):
    """oar_layout_run = rows + postprocess_pp_doclayout: the kept elements equal the oracle's on its own rows
    (same classes, scores within 2e-3, boxes within 0.25 px)"""
    import ctypes as C
    from oar_ocr_b200 import ffi, synth
    from oracle import cpu
    from oracle.rtdetr import RTDetrL
    in_hw = (256, 256)
    pages = [synth.page(410, 480), synth.page(411, 320)]
    enc, head = _layout_models(ctx, in_hw)
    arrs, ptrs, hs, ws = ffi._image_table(pages)
    cfg = ffi.LayoutConfig()
    ffi.lib().oar_layout_config_default(C.byref(cfg))
    cfg.score_threshold = 0.3
    cfg.num_classes = 23
    me = cfg.max_elements
    boxes = np.zeros((2, me, 4), np.float32)
    classes = np.zeros((2, me), np.int32)
    scores = np.zeros((2, me), np.float32)
    counts = np.zeros(2, np.int32)
    ffi.check(ffi.lib().oar_layout_run(enc.handle, head.handle, ptrs, ffi._ptr(hs), ffi._ptr(ws), 2, in_hw[0], in_hw[1],
                                       C.byref(cfg), ffi._ptr(boxes), ffi._ptr(classes), ffi._ptr(scores), ffi._ptr(counts)))
    x, _ = cpu.layout_preprocess(pages, in_hw)
    rows = RTDetrL(42).rows(x, [(p.shape[1], p.shape[0]) for p in pages])
    total = 0
    for b, p in enumerate(pages):
        ob, oc, os_ = cpu.layout_postprocess(rows[b], p.shape[1], p.shape[0], 23, score_threshold=0.3,
                                             image_class_id=cfg.image_class_id, formula_class_id=cfg.formula_class_id)
        n = int(counts[b])
        assert n == len(ob), (n, len(ob))
        for k in range(n):
            j = [i for i in range(n) if classes[b, i] == oc[k] and abs(scores[b, i] - os_[k]) <= 2e-3 and
                 np.abs(boxes[b, i] - ob[k]).max() <= 0.25]
            assert j, (b, k, oc[k], os_[k], ob[k])
        total += n
    assert total >= 1
